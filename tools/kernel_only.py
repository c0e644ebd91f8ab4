"""Runs one kernel family a few times (for ncu captures): python tools/kernel_only.py <what> [N] [reps]"""
import sys
sys.path.insert(0, ".")
import macroc_b200 as M
what = int(sys.argv[1]); N = int(sys.argv[2]) if len(sys.argv) > 2 else 256; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING))
m.apply_bc_on_u(-1e-3); m.set_strains(); m.assembly_res()
if what in (0, 2, 3):
    m.assembly_jac()
print(what, m.time_kernel(what, reps))
