"""Multi-rank worker launched by torch.distributed.run (test infrastructure).

--mode host (gloo, CPU): every rank asks the C-ABI library for its z-slab, Dirichlet lists
  and the bench's per-rank byte accounting; the ranks all_gather them and check that the
  slabs tile the grid, that neighbours agree on the halo plane, that the merged Dirichlet
  lists equal the single-rank list (the reference's decomposition-independence invariant).
--mode gpu (nccl): every rank owns one GPU, runs the time loop on its slab (NCCL halos and
  all-reduces inside the library), the slabs are gathered and rank 0 checks them against the
  single-rank CPU oracle.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import macroc_b200 as M  # noqa: E402


def gather_objects(obj):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def host_mode(args):
    rank, world = dist.get_rank(), dist.get_world_size()
    for (NX, NY, NZ) in [(5, 2, 2), (6, 4, 9), (16, 8, 7)]:
        if world > NZ:
            continue
        for bc in (M.BC_BENDING, M.BC_CIRCLE):
            cfg = M.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, lx=5.0, lz=6.0, px=1, py=1, pz=world)
            part = M.partition(cfg, rank, world)
            idx, coef = M.bc_lists(cfg, rank, world)
            parts = gather_objects(part)
            lists = gather_objects((idx, coef))
            # slabs tile [0, NZ) in rank order, element layers tile [0, NZ-1)
            z = 0
            ez = 0
            for r, p in enumerate(parts):
                xs, ys, zs, xm, ym, zm = p["corners"]
                assert (xs, ys, xm, ym) == (0, 0, NX, NY) and zs == z
                z += zm
                Xs, Ys, Zs, Xm, Ym, Zm = p["ghost_corners"]
                assert Zs == max(zs - 1, 0) and Zs + Zm == min(zs + zm + 1, NZ)
                assert p["elements_sizes"][:2] == (NX - 1, NY - 1)
                ez += p["elements_sizes"][2]
            assert z == NZ and ez == NZ - 1
            # halo agreement: my upper ghost plane is my upper neighbour's first owned plane
            if rank + 1 < world:
                up = parts[rank + 1]["corners"]
                me = part["ghost_corners"]
                assert me[2] + me[5] - 1 == up[2]
            # merged Dirichlet lists == single-rank list (values too)
            merged = {}
            for ix, cf in lists:
                for i, c in zip(ix, cf):
                    if i >= 0:
                        assert merged.setdefault(int(i), float(c)) == float(c)
            ix1, cf1 = M.bc_lists(M.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, lx=5.0, lz=6.0), 0, 1)
            single = {int(i): float(c) for i, c in zip(ix1, cf1) if i >= 0}
            assert merged == single, (NX, NY, NZ, bc)
    # bench.py's max/sum-over-ranks helpers on the gloo backend
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    if rank == 0:
        print("HOST-OK")


def gpu_mode(args):
    from oracle import oracle as O
    from helpers import rel_err
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
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
    if os.environ.get("MACROC_TEST_CASES") == "boxes":      # only the 2-D / 3-D processor grids
        cases = [c for c in cases if sum(1 for q in c[5] if q != 1) >= 2]
    for (NX, NY, NZ, bc, extra, pg) in cases:
        variants = [(M.OP_ASSEMBLED, M.MAT_UNIFORM), (M.OP_MATRIX_FREE, M.MAT_UNIFORM), (M.OP_ASSEMBLED, M.MAT_PER_GP)]
        if os.environ.get("MACROC_SYM_MULTIRANK", "0") != "0":   # symmetric storage on several ranks (opt-in)
            variants.append((M.OP_ASSEMBLED_SYM, M.MAT_UNIFORM))
        for op, material in variants:
            box = [M.get_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            ts = 3
            cfg = M.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, px=pg[0], py=pg[1], pz=pg[2], ts=ts, ksp_rtol=1e-12,
                           op=op, device=local, material=material, **extra)
            m = M.MacroC(cfg, rank=rank, nranks=world, unique_id=box[0])
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
                import tempfile
                with tempfile.TemporaryDirectory() as d:     # every rank writes its piece
                    m.write_pvtu(os.path.join(d, f"sol_{NX}_{rank}"))
                    piece = open(os.path.join(d, f"sol_{NX}_{rank}-subdo-{rank}.vtu")).read()
                    gx, gy, gz = p["ghost_corners"][3:]
                    ex, ey, ez = p["elements_sizes"]
                    assert f'NumberOfPoints="{gx * gy * gz}" NumberOfCells="{ex * ey * ez}"' in piece
            got = gather_objects((u_loc, y_loc, A_loc, logs, force, nodes, eps_loc))
            m.close()
            if rank == 0:
                o = O.Oracle(O.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, ts=ts, rtol=1e-12, faithful_ke=0, **extra))
                ologs = o.run()
                u = np.zeros(3 * NX * NY * NZ); y = np.zeros_like(u)
                for g in got:
                    u.reshape(-1, 3)[g[5]] = g[0].reshape(-1, 3); y.reshape(-1, 3)[g[5]] = g[1].reshape(-1, 3)
                assert rel_err(u, o.get_vec("u")) < 1e-9, (NX, NY, NZ, bc, op, rel_err(u, o.get_vec("u")))
                o.assembly_jac()
                assert rel_err(y, o.matmult(x)) < 1e-13
                if assembled:
                    A = np.zeros((NX * NY * NZ, 27, 3, 3))
                    for g in got:
                        A[g[5]] = g[2]
                    if op == M.OP_ASSEMBLED_SYM:             # stored half bitwise, mirrored half to rounding
                        assert np.array_equal(A[:, 13:], o.block_stencil()[:, 13:])
                        assert rel_err(A, o.block_stencil()) < 1e-14
                    elif material == M.MAT_UNIFORM:
                        assert np.array_equal(A, o.block_stencil())
                    else:
                        assert rel_err(A, o.block_stencil()) < 1e-12
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
            dist.barrier()
    if rank == 0:
        print("GPU-OK")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["host", "gpu"], required=True)
    args = ap.parse_args()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.mode == "host":
        dist.init_process_group("gloo")
        host_mode(args)
    else:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gpu_mode(args)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
