"""Times the per-element Jacobian kernel (uniform D from constant memory, then per-GP tangents)."""
import sys
sys.path.insert(0, ".")
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nb = (3 * N - 2) ** 3; ne = (N - 1) ** 3
for mat in (M.MAT_UNIFORM, M.MAT_PER_GP):
    m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, material=mat))
    m.apply_bc_on_u(-1e-3); m.set_strains(); m.homogenize()
    m.time_kernel(7, 1)
    ms = m.time_kernel(7, 3)
    byt = 72 * nb + (2304 * ne if mat == M.MAT_PER_GP else 0)
    print("material", mat, "element-kernel Jacobian ms", ms, "GB/s algorithmic", byt / ms / 1e6, "GFMA/s", 17.3e3 * ne / ms / 1e6, flush=True)
    if mat == M.MAT_PER_GP:
        print("residual (per-GP stresses) ms", m.time_kernel(4, 3))
    m.close()
