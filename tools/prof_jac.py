"""Element-Jacobian kernel alone (uniform tangent, then per-GP) for an ncu capture."""
import sys
sys.path.insert(0, ".")
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, jac_mode=M.JAC_ELEMENT))
m.apply_bc_on_u(-1e-3); m.set_strains()
print("uniform ms", m.time_kernel(7, 2))
m.close()
if "pergp" in sys.argv:
    m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, material=M.MAT_PER_GP))
    m.apply_bc_on_u(-1e-3); m.set_strains(); m.homogenize()
    print("per-GP ms", m.time_kernel(7, 2))
    m.close()
