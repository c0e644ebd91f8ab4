"""Time the symmetric-storage SpMV (and, with `full`, the full-storage one on the same box).
MACROC_SYM_VARIANT / MACROC_SYM_HINT / MACROC_SYM_NSEG select the kernel variant."""
import sys
sys.path.insert(0, ".")
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
if "full" in sys.argv:
    m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, op=M.OP_ASSEMBLED))
    m.assembly_jac()
    print("full", m.time_kernel(0, 8), "cg", m.time_kernel(2, 8))
    del m
m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, op=M.OP_ASSEMBLED_SYM))
m.assembly_jac()
print("sym", m.time_kernel(8, 8), "cg", m.time_kernel(9, 8) if "cg" in sys.argv else "")
