"""Quick kernel timings on one GPU (development aid; bench.py is the contract)."""
import json
import sys
import time

sys.path.insert(0, ".")
import macroc_b200 as M

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
NY = int(sys.argv[2]) if len(sys.argv) > 2 else N
NZ = int(sys.argv[3]) if len(sys.argv) > 3 else N
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
m = M.MacroC(M.Config(NX=N, NY=NY, NZ=NZ, bc_type=M.BC_BENDING, lx=1.0, ly=1.0, lz=1.0))
nn = N * NY * NZ
nd = 3 * nn
nb = (3 * N - 2) * (3 * NY - 2) * (3 * NZ - 2)
m.apply_bc_on_u(-1e-3)
m.set_strains()
t0 = time.time(); norm = m.assembly_res(); t_res = time.time() - t0
t0 = time.time(); m.assembly_jac(); m.synchronize(); t_jac = time.time() - t0
out = {"grid": [N, NY, NZ], "ndof": nd, "res_first_call_s": t_res, "jac_first_call_s": t_jac, "norm": norm}
names = {0: "spmv", 1: "apply_mf", 2: "cg_iter", 5: "cg_iter_mf", 3: "jac_fill", 4: "residual"}
bytes_ = {0: 72 * nb + 16 * nd, 1: 16 * nd, 2: 72 * nb + 104 * nd, 5: 104 * nd, 3: 72 * nb, 4: 16 * nd}
for what in (0, 1, 3, 4, 2, 5):
    m.time_kernel(what, 3)
    ms = m.time_kernel(what, reps)
    out[names[what]] = {"ms": ms, "GBps_algorithmic": bytes_[what] / ms / 1e6, "dof_per_s": nd / ms * 1e3}
print(json.dumps(out, indent=1))
