"""One line per (variant, R, nseg) of the symmetric-storage SpMV; the MACROC_SYM_* knobs are read at create."""
import os, statistics, sys
sys.path.insert(0, ".")
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for variant, R, nseg in [(0, 0, 0), (1, 0, 0), (1, 6, 4), (1, 5, 3), (1, 4, 2), (7, 0, 0), (7, 8, 4), (7, 8, 5), (7, 7, 4), (5, 10, 6), (5, 9, 5)]:
    os.environ["MACROC_SYM_VARIANT"] = str(variant); os.environ["MACROC_SYM_R"] = str(R); os.environ["MACROC_SYM_NSEG"] = str(nseg)
    m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, op=M.OP_ASSEMBLED_SYM))
    m.assembly_jac()
    m.time_kernel(8, 3)
    t = statistics.median(m.time_kernel(8, 1) for _ in range(10))
    print(f"variant {variant} R {R} nseg {nseg}: spmv {t:.3f} ms", flush=True)
    m.close()
