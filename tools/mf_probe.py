"""Matrix-free apply: the z-marching kernel (default) against the assembled operator, then timings at N^3
(MACROC_MF_VARIANT=1: the patch-form kernel of round 1; MACROC_MF_NSEG overrides the z-segment count)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
if "check" in sys.argv:
    rng = np.random.default_rng(3)
    for n3, bc in (((33, 5, 4), M.BC_BENDING), ((40, 17, 9), M.BC_BENDING), ((70, 9, 35), M.BC_CIRCLE), ((4, 4, 2), M.BC_BENDING), ((64, 64, 20), M.BC_BENDING)):
        m = M.MacroC(M.Config(NX=n3[0], NY=n3[1], NZ=n3[2], bc_type=bc)); m.assembly_jac()
        x = rng.standard_normal(3 * n3[0] * n3[1] * n3[2])
        ya = m.matmult(x, M.OP_ASSEMBLED)
        m.set_operator(M.OP_MATRIX_FREE); m.assembly_jac()
        ym = m.matmult(x, M.OP_MATRIX_FREE)
        print("check", n3, "bc", bc, "max rel err", np.abs(ya - ym).max() / np.abs(ya).max(), flush=True)
        m.close()
m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, op=M.OP_MATRIX_FREE))
m.apply_bc_on_u(-1e-3); m.set_strains(); m.assembly_res(); m.assembly_jac()
m.time_kernel(1, 3)
print("apply_matrix_free ms", m.time_kernel(1, 20), " pcg_iteration_matrix_free ms", m.time_kernel(5, 20), flush=True)
m.close()
