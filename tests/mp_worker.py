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


class TorchComm:
    def __init__(self):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def gather(self, obj):
        return gather_objects(obj)

    def bcast(self, obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def barrier(self):
        dist.barrier()


def gpu_mode(args):
    from multirank_cases import cases_for, run_cases
    comm = TorchComm()
    local = int(os.environ.get("LOCAL_RANK", comm.rank))
    torch.cuda.set_device(local)
    cases = cases_for(comm.world, os.environ.get("MACROC_TEST_CASES", "all"))
    run_cases(comm, local, M.get_unique_id, cases)
    if comm.rank == 0:
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
