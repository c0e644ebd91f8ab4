"""CPU tests of the oracle (the parity checker): known answers, an independent
numpy/scipy restatement, the reference binary built over the PETSc shim, and the
committed golden fixtures that binary produced."""
import os
import re
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import oracle as O
from helpers import (cfg_kwargs_from_flags, csr_to_block_stencil, golden_cases, load_golden, newton_driver,
                     rel_err)

SGN = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float)
XG = 0.577350269189626


def np_B(gp):
    """Independent restatement of calc_B (reference src/assembly.c:195-254), unit cube."""
    xi = SGN[gp] * XG
    B = np.zeros((6, 24))
    for n in range(8):
        f = 1 + SGN[n] * xi
        dsh = np.array([SGN[n, 0] * f[1] * f[2], SGN[n, 1] * f[0] * f[2], SGN[n, 2] * f[0] * f[1]]) / 8 * 2
        B[0, 3 * n] = dsh[0]; B[1, 3 * n + 1] = dsh[1]; B[2, 3 * n + 2] = dsh[2]
        B[3, 3 * n] = dsh[1]; B[3, 3 * n + 1] = dsh[0]
        B[4, 3 * n] = dsh[2]; B[4, 3 * n + 2] = dsh[0]
        B[5, 3 * n + 1] = dsh[2]; B[5, 3 * n + 2] = dsh[1]
    return B


def np_D(E=1e7, nu=0.25):
    lam = E * nu / ((1 + nu) * (1 - 2 * nu)); mu = E / (2 * (1 + nu))
    D = np.zeros((6, 6)); D[:3, :3] = lam; D[:3, :3] += 2 * mu * np.eye(3); D[3:, 3:] = mu * np.eye(3)
    return D


def np_assemble(NX, NY, NZ, wg, D):
    """Independent global assembly in natural ordering with scipy COO."""
    Ke = sum(np_B(g).T @ D @ np_B(g) for g in range(8)) * wg
    pos = ((SGN + 1) // 2).astype(int)
    rows, cols, vals = [], [], []
    for k in range(NZ - 1):
        for j in range(NY - 1):
            for i in range(NX - 1):
                nodes = [(i + p[0]) + NX * ((j + p[1]) + NY * (k + p[2])) for p in pos]
                dof = np.array([[3 * n + d for d in range(3)] for n in nodes]).reshape(-1)
                rows.append(np.repeat(dof, 24)); cols.append(np.tile(dof, 24)); vals.append(Ke.reshape(-1))
    n = 3 * NX * NY * NZ
    return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr(), Ke


def test_calc_B_matches_independent_restatement():
    for gp in range(8):
        assert np.array_equal(O.calc_B(gp), np_B(gp))
    # partition of unity: derivatives sum to zero; 9 non-zeros per node
    B = O.calc_B(5)
    assert np.count_nonzero(B) == 72
    assert abs(B[0, 0::3].sum()) < 1e-15


def test_element_matrix_known_answers():
    """SURVEY.md 8a 'derived known-answer values' for sum_gp B^T D B."""
    D = O.isotropic_D()
    assert np.array_equal(D, np_D())
    Ke = O.elem_jac(np.tile(D.reshape(-1), (8, 1)), 1.0)
    assert Ke[0, 0] == pytest.approx(17777777.78, rel=1e-9)
    assert Ke[0, 1] == pytest.approx(5333333.33, rel=1e-9)
    assert Ke[0, 3] == pytest.approx(-7111111.11, rel=1e-9)
    assert np.trace(Ke) == pytest.approx(426666666.67, rel=1e-10)
    assert np.abs(Ke - Ke.T).max() < 1e-7
    assert np.abs(Ke.sum(axis=1)).max() < 1e-6
    ev = np.linalg.eigvalsh((Ke + Ke.T) / 2)
    assert np.sum(np.abs(ev) < 1e-3) == 6                      # rigid-body modes
    assert ev.max() == pytest.approx(8.0e7, rel=1e-9)
    Ke_np = sum(np_B(g).T @ np_D() @ np_B(g) for g in range(8))
    assert rel_err(Ke, Ke_np) < 1e-14


def test_element_residual_is_Ke_times_u():
    rng = np.random.default_rng(0)
    ue = rng.standard_normal(24)
    D = np_D()
    stress = np.array([D @ (np_B(g) @ ue) for g in range(8)])
    be = O.elem_res(stress, 0.37)
    Ke = O.elem_jac(np.tile(D.reshape(-1), (8, 1)), 0.37)
    assert rel_err(be, Ke @ ue) < 1e-13


@pytest.mark.parametrize("grid,bc", [((4, 4, 2), 0), ((5, 3, 4), 1), ((6, 5, 4), 0)])
def test_assembly_matches_scipy(grid, bc):
    NX, NY, NZ = grid
    o = O.Oracle(O.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, ts=2, lx=3.0, ly=2.0, lz=4.0))
    u = np.random.default_rng(1).standard_normal(o.ndof)
    o.set_vec("u", u)
    o.set_strains(); o.homogenize()
    norm = o.assembly_res()
    o.assembly_jac()
    K, _ = np_assemble(NX, NY, NZ, o.wg, np_D())
    mask = o.dirichlet_mask_natural()
    b_np = -(K @ u); b_np[mask] = 0
    assert rel_err(o.get_vec("b"), b_np) < 1e-12
    assert norm == pytest.approx(np.linalg.norm(b_np), rel=1e-12)
    M = sp.diags((~mask).astype(float))
    A_np = (M @ K @ M + sp.diags(mask.astype(float))).tocsr()
    rowptr, col, val = o.csr()
    A_or = sp.csr_matrix((val, col, rowptr), shape=(o.ndof, o.ndof))
    assert abs(A_or - A_np).max() / abs(A_np).max() < 1e-14
    # block-stencil export agrees with the CSR
    blocks = o.block_stencil()
    assert np.array_equal(blocks, csr_to_block_stencil(rowptr, col, val, NX, NY, NZ))
    # y = A x
    x = np.random.default_rng(2).standard_normal(o.ndof)
    assert rel_err(o.matmult(x), A_np @ x) < 1e-13


def test_patch_test_linear_field_has_zero_interior_residual():
    NX, NY, NZ = 6, 5, 4
    o = O.Oracle(O.Config(NX=NX, NY=NY, NZ=NZ, bc_type=0))
    i, j, k = np.meshgrid(np.arange(NX), np.arange(NY), np.arange(NZ), indexing="ij")
    def nat(a):
        return a.transpose(2, 1, 0).reshape(-1)
    ux = 1e-3 * nat(i) + 2e-3 * nat(j); uy = -1e-3 * nat(k); uz = 5e-4 * nat(i)
    o.set_vec("u", np.stack([ux, uy, uz], axis=1).reshape(-1))
    o.set_strains(); o.homogenize()
    eps = o.strain(0)
    assert np.abs(eps - eps[0, 0]).max() < 1e-15             # constant strain (unit-cube B)
    assert eps[0, 0] == pytest.approx([1e-3, 0, 0, 2e-3, 5e-4, -1e-3], abs=1e-16)
    o.assembly_res()
    b = o.get_vec("b").reshape(NZ, NY, NX, 3)
    assert np.abs(b[1:-1, 1:-1, 1:-1]).max() < 1e-9 * np.abs(b).max()


@pytest.mark.parametrize("grid,bc", [((5, 2, 2), 0), ((4, 4, 4), 1), ((16, 6, 6), 0)])
def test_pcg_solution_matches_direct_solve(grid, bc):
    NX, NY, NZ = grid
    o = O.Oracle(O.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, ts=2, rtol=1e-13))
    o.apply_bc_on_u(o.get_displacement(1))
    o.set_strains(); o.homogenize(); o.assembly_res(); o.assembly_jac()
    its, rn = o.solve()
    rowptr, col, val = o.csr()
    A = sp.csr_matrix((val.copy(), col.copy(), rowptr.copy()), shape=(o.ndof, o.ndof))
    x = spla.spsolve(A.tocsc(), o.get_vec("b"))
    assert rel_err(o.get_vec("du"), x) < 1e-9
    assert 0 < its < 200


@pytest.mark.parametrize("grid,bc", [((5, 2, 2), 0), ((5, 3, 5), 0), ((4, 4, 4), 1), ((9, 7, 8), 0)])
def test_decomposition_independence(grid, bc):
    """The reference's implied invariant (tests/CMakeLists.txt:21-28: same grid at -np 1..8)."""
    NX, NY, NZ = grid
    base = None
    for nr, pg in [(1, (0, 0, 0)), (2, (0, 0, 0)), (3, (0, 0, 0)), (4, (0, 0, 0)), (8, (0, 0, 0)),
                   (2, (1, 1, 2)), (4, (2, 2, 1))]:
        try:
            o = O.Oracle(O.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, ts=3, nranks=nr, px=pg[0], py=pg[1], pz=pg[2]))
        except ValueError:
            continue
        logs = o.run()
        cur = (o.get_vec("u"), o.block_stencil(), [l.newton_its for l in logs], [l.ksp_its for l in logs])
        if base is None:
            base = cur
            continue
        assert rel_err(cur[1], base[1]) < 1e-14
        assert cur[2] == base[2] and cur[3] == base[3]
        assert rel_err(cur[0], base[0]) < 1e-7      # tiny grids: CG terminates at machine-level residuals
    assert base is not None


def test_time_step_zero_is_a_noop_and_one_newton_iteration_after():
    o = O.Oracle(O.Config(NX=6, NY=4, NZ=4, bc_type=0, ts=3))
    logs = o.run()
    assert logs[0].newton_its == 0 and logs[0].res_norm == [0.0]
    assert [l.newton_its for l in logs[1:]] == [1, 1]


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_reproduces_reference_golden(name):
    """Fixtures were produced by the reference's own sources (oracle/_ref/macroc_ref)."""
    z, kv = load_golden(name)
    kw = cfg_kwargs_from_flags(kv)
    o = O.Oracle(O.Config(**kw))
    logs = o.run()
    res = [r for l in logs for r in l.res_norm]
    assert np.allclose(res, z["res_norms"], rtol=2e-6, atol=0)      # printed with %e
    assert [i for l in logs for i in l.ksp_its] == list(z["ksp_its"])
    assert np.array_equal(o.get_vec("u"), z["u"])                    # bit for bit
    rowptr, col, val = o.csr()
    assert np.array_equal(rowptr, z["rowptr"]) and np.array_equal(col, z["col"]) and np.array_equal(val, z["val"])
    assert np.array_equal(o.get_vec("du"), z["x"])
    assert np.allclose([l.force for l in logs], z["force"], rtol=2e-6, atol=1e-300)
    # replay through the per-function API and catch the b handed to the last KSPSolve
    o2 = O.Oracle(O.Config(**kw))
    seen = {}
    def cap(time_s, it, stage):
        if stage == "pre_solve":
            seen["b"] = o2.get_vec("b")
    log2 = newton_driver(o2, kw["ts"], capture=cap)
    assert np.array_equal(seen["b"], z["b"])
    assert [l["newton_its"] for l in log2] == [l.newton_its for l in logs]


@pytest.mark.skipif(not (os.path.exists(O.REF_BIN) and os.path.isdir("/root/reference")),
                    reason="reference tree / oracle/_ref not present")
@pytest.mark.parametrize("flags", [
    "-da_grid_x 5 -da_grid_y 2 -da_grid_z 2 -ts 5",                # tests/CMakeLists.txt:21,32 (default BC_CIRCLE)
    "-da_grid_x 4 -da_grid_y 4 -da_grid_z 4 -ts 5",                # :31
    "-da_grid_x 5 -da_grid_y 3 -da_grid_z 5 -ts 5 -bc_type 0",     # :28 + bending
    "-da_grid_x 4 -da_grid_y 4 -da_grid_z 2 -ts 2 -bc_type 0",     # README example
    "-da_grid_x 12 -da_grid_y 5 -da_grid_z 7 -ts 2 -bc_type 0 -lx 10 -ly 1 -lz 1",
])
def test_oracle_log_equals_live_reference_run(flags):
    args = flags.split()
    kv = {args[i]: args[i + 1] for i in range(0, len(args), 2)}
    kv.setdefault("-bc_type", "1")
    with tempfile.TemporaryDirectory() as d:
        ref = O.run_reference(args, d, os.path.join(d, "dump"))
        o = O.Oracle(O.Config(**cfg_kwargs_from_flags(kv)))
        o.run(os.path.join(d, "orc.log"))
        mine = open(os.path.join(d, "orc.log")).read()
        pick = lambda t: [l for l in t.splitlines() if re.match(r"(\|RES\||KSP :|Newton Iteration|Time Step)", l)]
        assert pick(ref) == pick(mine)
        assert np.array_equal(np.fromfile(os.path.join(d, "dump_vec0.bin")), o.get_vec("u"))


def test_openmp_mode_equals_strict_mode():
    """bench.py times the oracle with nthreads > 1 (ranks as threads): same operator and residual,
    solution to rounding (parallel reductions reorder the dot products)."""
    kw = dict(NX=12, NY=6, NZ=9, bc_type=0, ts=2, lx=10., ly=1., lz=1.)
    a = O.Oracle(O.Config(nranks=1, nthreads=1, **kw))
    b = O.Oracle(O.Config(nranks=3, px=1, py=1, pz=3, nthreads=3, **kw))
    la, lb = a.run(), b.run()
    assert [l.newton_its for l in la] == [l.newton_its for l in lb]
    assert rel_err(b.block_stencil(), a.block_stencil()) < 1e-14
    assert rel_err(b.get_vec("u"), a.get_vec("u")) < 1e-8
    assert b.time_cg_iterations(2) > 0


@pytest.mark.parametrize("grid,bc,extra,its", [((5, 2, 2), 0, {}, 9), ((4, 4, 2), 1, {}, 39), ((4, 4, 4), 1, {}, 28),
                                              ((16, 6, 6), 0, dict(lx=10.0), 49)])
def test_cg_iteration_counts_of_the_survey(grid, bc, extra, its):
    """CG iteration counts of the first loaded time step, obtained there with an
    independent numpy restatement of PETSc's KSPCG + PCJACOBI semantics."""
    NX, NY, NZ = grid
    o = O.Oracle(O.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, ts=2, **extra))
    logs = o.run()
    assert logs[0].newton_its == 0 and logs[1].newton_its == 1 and logs[1].ksp_its == [its]
