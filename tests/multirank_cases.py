"""Decomposition-independence cases shared by the NCCL worker (tests/mp_worker.py, one process
per GPU) and the loopback tests (tests/test_loopback.py, one thread per rank on one GPU).

The reference's own invariant (tests/CMakeLists.txt:21-28: the same grid at -np 1,2,3,4,8):
whatever the processor grid, the time loop must reproduce the single-rank CPU oracle --
displacement, operator entries, SpMV, Newton history, reaction force, Gauss-point export.
Test infrastructure; `comm` only has to offer rank, world, gather(obj) and bcast(obj).
"""
import os
import tempfile

import numpy as np

import macroc_b200 as M
from helpers import rel_err


def cases_for(world, which="all"):
    zs = (1, 1, world)
    cases = [(8, 5, 2 * world + 1, M.BC_BENDING, {}, zs), (40, 3, 40, M.BC_CIRCLE, {}, zs),
             (33, 9, 4 * world + 3, M.BC_BENDING, dict(lx=10., ly=1., lz=1.), zs),
             (9, 3, max(9, world), M.BC_CIRCLE, dict(lx=4., lz=4.), zs),
             (8, 5, world, M.BC_BENDING, {}, zs),                       # one plane per rank
             # general DMDA boxes: x split, y split, PETSC_DECIDE
             (2 * world + 3, 6, 5, M.BC_BENDING, {}, (world, 1, 1)),
             (9, 2 * world + 2, 7, M.BC_CIRCLE, dict(lx=4., lz=4.), (1, world, 1)),
             (37, 11, 9, M.BC_BENDING, dict(lx=10., ly=1., lz=1.), (0, 0, 0))]
    if world == 4:
        cases += [(12, 10, 9, M.BC_BENDING, {}, (2, 2, 1)), (11, 5, 10, M.BC_CIRCLE, dict(lx=4., lz=4.), (2, 1, 2))]
    if world == 8:
        cases += [(12, 10, 9, M.BC_BENDING, {}, (2, 2, 2)), (13, 9, 11, M.BC_CIRCLE, dict(lx=4., lz=4.), (0, 0, 0))]
    if which == "boxes":                                     # only the 2-D / 3-D processor grids
        cases = [c for c in cases if sum(1 for q in c[5] if q != 1) >= 2]
    elif which == "slabs":
        cases = [c for c in cases if c[5] == zs]
    elif which == "quick":                                   # one slab case, one split per axis, PETSC_DECIDE
        cases = [cases[2], cases[4], cases[5], cases[6], cases[7]] + cases[8:]
    return cases


VARIANTS = [(M.OP_ASSEMBLED, M.MAT_UNIFORM), (M.OP_MATRIX_FREE, M.MAT_UNIFORM), (M.OP_ASSEMBLED, M.MAT_PER_GP),
            (M.OP_ASSEMBLED_SYM, M.MAT_UNIFORM), (M.OP_ASSEMBLED_SYM, M.MAT_PER_GP)]


def run_cases(comm, device, make_id, cases, variants=VARIANTS, log=None):
    """Every rank calls this with its own `comm`; rank 0 checks against the single-rank oracle."""
    from oracle import oracle as O
    rank, world = comm.rank, comm.world
    for (NX, NY, NZ, bc, extra, pg) in cases:
        for op, material in variants:
            uid = comm.bcast(make_id() if rank == 0 else None)
            ts = 3
            cfg = M.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, px=pg[0], py=pg[1], pz=pg[2], ts=ts, ksp_rtol=1e-12,
                           op=op, device=device, material=material, **extra)
            m = M.MacroC(cfg, rank=rank, nranks=world, unique_id=uid)
            logs = [m.time_step(t) for t in range(ts)]
            u_loc = m.get_vec(M.VEC_U)
            force = m.calc_force()
            x = np.sin(0.37 * np.arange(3 * NX * NY * NZ)) + 0.1
            p = M.partition(cfg, rank, world)
            xs0, ys0, zs0, xm, ym, zm = p["corners"]
            box = np.zeros((NZ, NY, NX), bool); box[zs0:zs0 + zm, ys0:ys0 + ym, xs0:xs0 + xm] = True
            nodes = np.flatnonzero(box.reshape(-1))            # owned nodes, x fastest inside the box
            assembled = op in (M.OP_ASSEMBLED, M.OP_ASSEMBLED_SYM)
            if assembled:
                m.set_strains(); m.homogenize(); m.assembly_jac()
            y_loc = m.matmult(x.reshape(-1, 3)[nodes].reshape(-1), op)
            A_loc = m.get_matrix_blocks() if assembled else None
            eps_loc = None
            if material == M.MAT_UNIFORM and op == M.OP_ASSEMBLED:
                m.set_vec(M.VEC_U, u_loc)                    # strains of the converged displacement
                m.set_strains(materialize=True)
                eps_loc = m.get_strain_stress()
                with tempfile.TemporaryDirectory() as d:     # every rank writes its piece
                    m.write_pvtu(os.path.join(d, f"sol_{NX}_{rank}"))
                    piece = open(os.path.join(d, f"sol_{NX}_{rank}-subdo-{rank}.vtu")).read()
                    gx, gy, gz = p["ghost_corners"][3:]
                    ex, ey, ez = p["elements_sizes"]
                    assert f'NumberOfPoints="{gx * gy * gz}" NumberOfCells="{ex * ey * ez}"' in piece
            got = comm.gather((u_loc, y_loc, A_loc, logs, force, nodes, eps_loc))
            m.close()
            if rank == 0:
                o = O.Oracle(O.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, ts=ts, rtol=1e-12, faithful_ke=0, **extra))
                ologs = o.run()
                u = np.zeros(3 * NX * NY * NZ); y = np.zeros_like(u)
                for g in got:
                    u.reshape(-1, 3)[g[5]] = g[0].reshape(-1, 3); y.reshape(-1, 3)[g[5]] = g[1].reshape(-1, 3)
                eu = rel_err(u, o.get_vec("u"))
                if not eu < 1e-9:                            # say which rank, and how its history went
                    uo = o.get_vec("u").reshape(-1, 3)
                    per_rank = [float(np.abs(g[0].reshape(-1, 3) - uo[g[5]]).max() / np.abs(uo).max()) for g in got]
                    raise AssertionError((NX, NY, NZ, bc, pg, op, material, eu, per_rank, [g[3] for g in got][:2],
                                          [(l.newton_its, l.ksp_its) for l in ologs]))
                o.assembly_jac()
                ey_ = rel_err(y, o.matmult(x))
                assert ey_ < 1e-13, (NX, NY, NZ, bc, pg, op, material, ey_)
                if assembled:
                    A = np.zeros((NX * NY * NZ, 27, 3, 3))
                    for g in got:
                        A[g[5]] = g[2]
                    if material == M.MAT_PER_GP:
                        assert rel_err(A, o.block_stencil()) < 1e-12
                    elif op == M.OP_ASSEMBLED_SYM:           # stored half bitwise, mirrored half to rounding
                        assert np.array_equal(A[:, 13:], o.block_stencil()[:, 13:])
                        assert rel_err(A, o.block_stencil()) < 1e-14
                    else:
                        assert np.array_equal(A, o.block_stencil())
                # reaction force: the reference's own rank logic (forces.c:75,133) depends on the
                # decomposition, so compare with the oracle run on the same processor grid
                om = O.Oracle(O.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, ts=ts, rtol=1e-12, faithful_ke=0, nranks=world,
                                       px=pg[0], py=pg[1], pz=pg[2], **extra))
                f_ref = om.run()[-1].force
                om.set_strains(); om.homogenize()
                for r_, g in enumerate(got):
                    if g[6] is not None:                     # Gauss-point export in DMDA element order
                        assert rel_err(g[6][0], om.strain(r_)) < 1e-7 and rel_err(g[6][1], om.stress(r_)) < 1e-7, (NX, NY, NZ, pg, r_)
                for g in got:
                    assert [l["newton_its"] for l in g[3]] == [l.newton_its for l in ologs]
                    assert g[3] == got[0][3]                 # every rank saw the same history
                    assert abs(g[4] - f_ref) <= 1e-6 * abs(f_ref) + 1e-9, (NX, NY, NZ, bc, pg, g[4], f_ref)
                if log is not None:
                    log.append(dict(grid=(NX, NY, NZ), bc=bc, procs=pg, op=op, material=material, err_u=eu, err_y=ey_))
            comm.barrier()
