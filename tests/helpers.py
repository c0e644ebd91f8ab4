"""Shared helpers for the parity tests (test infrastructure)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    args = [str(a) for a in z["args"]]
    kv = {args[i]: args[i + 1] for i in range(0, len(args), 2)}
    return z, kv


def cfg_kwargs_from_flags(kv):
    """MacroC flags -> the keyword names shared by oracle.Config and macroc_b200.Config."""
    out = dict(NX=int(kv["-da_grid_x"]), NY=int(kv["-da_grid_y"]), NZ=int(kv["-da_grid_z"]),
               ts=int(kv["-ts"]), bc_type=int(kv["-bc_type"]))
    for k in ("lx", "ly", "lz"):
        if "-" + k in kv:
            out[k] = float(kv["-" + k])
    return out


def csr_to_block_stencil(rowptr, col, val, NX, NY, NZ):
    """Scalar CSR in natural ordering -> [node, 27, 3, 3] (absent slots = 0)."""
    nn = NX * NY * NZ
    out = np.zeros((nn, 27, 3, 3))
    rows = np.repeat(np.arange(3 * nn), np.diff(rowptr))
    rn, rd = rows // 3, rows % 3
    cn, cd = col // 3, col % 3
    def ijk(n):
        return n % NX, (n // NX) % NY, n // (NX * NY)
    ri, rj, rk = ijk(rn)
    ci, cj, ck = ijk(cn)
    slot = (ck - rk + 1) * 9 + (cj - rj + 1) * 3 + (ci - ri + 1)
    out[rn, slot, rd, cd] = val
    return out


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    return float(np.abs(a - b).max() / scale) if b.size else 0.0


def slab_of(v_natural, NX, NY, zs, nz, ncomp=3):
    """Rows of a natural-order nodal array that belong to planes [zs, zs+nz)."""
    npl = NX * NY
    return v_natural.reshape(-1, ncomp)[zs * npl:(zs + nz) * npl].reshape(-1)


def newton_driver(p, ts, newton_max_its=5, min_tol=1e-1, rel_tol=1e-4, solve=None, capture=None):
    """The reference's time/Newton loop (src/main.c:49-82) over any object with
    the reference's function names (the oracle or macroc_b200.MacroC).
    capture(time_s, newton_it, stage) is called with stage in {"pre_solve", "post_solve"}."""
    solve = solve or (p.solve_Ax if hasattr(p, "solve_Ax") else p.solve)
    log = []
    for time_s in range(ts):
        p.apply_bc_on_u(p.get_displacement(time_s))
        it, res, kits = 0, [], []
        norm0 = 0.0
        while it < newton_max_its:
            p.set_strains()
            if hasattr(p, "homogenize"):
                p.homogenize()
            norm = p.assembly_res()
            res.append(norm)
            if it == 0:
                norm0 = norm
            if norm < min_tol or norm < norm0 * rel_tol:
                break
            p.assembly_jac()
            if capture:
                capture(time_s, it, "pre_solve")
            its, rn = solve()
            kits.append(its)
            if capture:
                capture(time_s, it, "post_solve")
            p.update_u()
            it += 1
        log.append({"newton_its": it, "res_norm": res, "ksp_its": kits})
    return log
