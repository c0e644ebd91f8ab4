"""Times the assembled-SpMV variants (MACROC_SPMV_VARIANT) on one GPU and checks them against variant 0."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import macroc_b200 as M

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 10, 11, 14]
nd = 3 * N ** 3
nb = (3 * N - 2) ** 3
bytes_spmv = 72 * nb + 16 * nd
# correctness on a ragged grid first
ref = None
xs = np.sin(0.37 * np.arange(3 * 37 * 11 * 9)) + 0.1
for v in variants:
    os.environ["MACROC_SPMV_VARIANT"] = str(v)
    m = M.MacroC(M.Config(NX=37, NY=11, NZ=9, bc_type=0))
    m.assembly_jac()
    y = m.matmult(xs)
    its, rn = (m.apply_bc_on_u(-1e-3), m.set_strains(), m.assembly_res(), m.solve_Ax())[3]
    if ref is None:
        ref = (y, its)
    print("variant", v, "matmult equal to variant 0:", bool(np.array_equal(y, ref[0])), "cg its", its, flush=True)
    m.close()
out = {}
for v in variants:
    os.environ["MACROC_SPMV_VARIANT"] = str(v)
    m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=0))
    m.assembly_jac()
    m.time_kernel(0, 3)
    ms = m.time_kernel(0, 10)
    ms_it = m.time_kernel(2, 10)
    out[v] = {"spmv_ms": ms, "GBps": bytes_spmv / ms / 1e6, "cg_iter_ms": ms_it}
    print(v, out[v], flush=True)
    m.close()
print(json.dumps(out))
