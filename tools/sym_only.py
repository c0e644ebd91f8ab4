import sys
sys.path.insert(0, ".")
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, op=M.OP_ASSEMBLED_SYM))
m.assembly_jac()
print(m.time_kernel(8, 4))
