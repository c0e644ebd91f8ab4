"""Sweep the band height R and the z segment count of the symmetric-storage SpMV (k_spmv_sym).
MACROC_SYM_R / MACROC_SYM_NSEG are read when a context is created."""
import os
import statistics
import sys
sys.path.insert(0, ".")
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
combos = [(0, 0), (16, 9), (16, 18), (16, 27), (8, 5), (8, 9), (8, 18), (12, 7), (12, 14), (4, 9), (16, 4)]
if "full" in sys.argv:
    m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, op=M.OP_ASSEMBLED))
    m.assembly_jac()
    m.time_kernel(0, 3)
    print("full-storage spmv", statistics.median(m.time_kernel(0, 1) for _ in range(10)),
          "pcg", statistics.median(m.time_kernel(2, 1) for _ in range(10)), flush=True)
    m.close()
for R, nseg in combos:
    os.environ["MACROC_SYM_R"] = str(R); os.environ["MACROC_SYM_NSEG"] = str(nseg)
    m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, op=M.OP_ASSEMBLED_SYM))
    m.assembly_jac()
    m.time_kernel(8, 3)
    t = statistics.median(m.time_kernel(8, 1) for _ in range(10))
    tc = statistics.median(m.time_kernel(9, 1) for _ in range(10))
    print(f"sym R={R} nseg={nseg}: spmv {t:.3f} ms  pcg iteration {tc:.3f} ms", flush=True)
    m.close()
