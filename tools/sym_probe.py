"""Times the symmetric-storage SpMV / PCG iteration against the full-storage ones."""
import sys
sys.path.insert(0, ".")
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for op, a, b in ((M.OP_ASSEMBLED, 0, 2), (M.OP_ASSEMBLED_SYM, 8, 9)):
    m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, op=op))
    m.assembly_jac()
    m.time_kernel(a, 3)
    print("op", op, "spmv ms", m.time_kernel(a, 10), "pcg iteration ms", m.time_kernel(b, 10), flush=True)
    m.close()
