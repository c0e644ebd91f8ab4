"""Times the per-element Jacobian kernel (uniform D from constant memory, then per-GP tangents) into the full
(what 7) and the symmetric (what 17) operator layout; `check`: both kernels and layouts against the class-stencil fill
on three grids."""
import sys
sys.path.insert(0, ".")
import numpy as np
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nb = (3 * N - 2) ** 3; ne = (N - 1) ** 3; nn = N ** 3
if "check" in sys.argv:                      # element kernel against the uniform-tangent fill (bitwise the oracle's operator)
    for n3 in ((33, 5, 4), (40, 7, 6), (64, 64, 8)):
        ref = M.MacroC(M.Config(NX=n3[0], NY=n3[1], NZ=n3[2], bc_type=M.BC_BENDING)); ref.assembly_jac(); Ar = ref.get_matrix_blocks(); ref.close()
        for mat in (M.MAT_UNIFORM, M.MAT_PER_GP):
            for op in (M.OP_ASSEMBLED, M.OP_ASSEMBLED_SYM):
                m = M.MacroC(M.Config(NX=n3[0], NY=n3[1], NZ=n3[2], bc_type=M.BC_BENDING, material=mat, jac_mode=M.JAC_ELEMENT, op=op))
                m.set_strains(); m.homogenize(); m.assembly_jac(); A = m.get_matrix_blocks(); m.close()
                lo = 13 if op == M.OP_ASSEMBLED_SYM else 0
                err = np.abs(A[:, lo:] - Ar[:, lo:]).max() / np.abs(Ar).max()
                print("check", n3, "material", mat, "op", op, "max rel err", err, flush=True)
for mat in (M.MAT_UNIFORM, M.MAT_PER_GP):
    m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, material=mat))
    m.apply_bc_on_u(-1e-3); m.set_strains(); m.homogenize()
    for what, name, byt in ((7, "full", 72 * nb), (17, "sym", 36 * (nb + nn))):
        m.time_kernel(what, 1)
        ms = m.time_kernel(what, 3)
        byt += 2304 * ne if mat == M.MAT_PER_GP else 0
        print("material", mat, "layout", name, "element-kernel Jacobian ms", round(ms, 3), "GB/s algorithmic", round(byt / ms / 1e6, 1), flush=True)
    m.close()
