"""Decomposition independence on ONE GPU: N contexts of one process, one host thread each, joined
by the in-process loopback communicator (macroc_loopback_id, csrc/loopback.h).  The multi-rank code
paths -- z-slabs, x / y / PETSC_DECIDE boxes, the three-phase halo, the Gauss-point halo, the
ghost-plane tiles of the symmetric operator -- run exactly as under NCCL and are checked against
the single-rank CPU oracle (the reference's invariant: tests/CMakeLists.txt:21-28, -np 1,2,3,4,8)."""
import threading

import numpy as np
import pytest

import macroc_b200 as M
from multirank_cases import VARIANTS, cases_for, run_cases

pytestmark = pytest.mark.gpu


class ThreadComm:
    """gather / bcast / barrier between the rank threads of one process."""

    class Shared:
        def __init__(self, world):
            self.bar = threading.Barrier(world, timeout=600)
            self.slots = [None] * world

    def __init__(self, shared, rank, world):
        self.sh, self.rank, self.world = shared, rank, world

    def gather(self, obj):
        self.sh.slots[self.rank] = obj
        self.sh.bar.wait()
        out = list(self.sh.slots)
        self.sh.bar.wait()
        return out

    def bcast(self, obj):
        return self.gather(obj)[0]

    def barrier(self):
        self.sh.bar.wait()


def run_world(world, fn):
    """fn(comm) on `world` threads; re-raises the first failure."""
    shared = ThreadComm.Shared(world)
    errors = []

    def body(rank):
        try:
            fn(ThreadComm(shared, rank, world))
        except BaseException as exc:                      # noqa: BLE001 -- reported below
            errors.append((rank, exc))
            shared.bar.abort()

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=1800)
    assert not any(t.is_alive() for t in threads), "a rank thread hangs"
    real = [e for e in errors if not isinstance(e[1], threading.BrokenBarrierError)] or errors
    if real:
        raise AssertionError(f"rank {real[0][0]}: {real[0][1]!r}") from real[0][1]


@pytest.mark.parametrize("world,which", [(2, "all"), (3, "quick"), (4, "quick"), (8, "quick")])
def test_decomposition_independence_loopback(world, which):
    log = []
    run_world(world, lambda comm: run_cases(comm, 0, lambda: M.loopback_id(world), cases_for(world, which), VARIANTS, log))
    assert log, "rank 0 checked nothing"
    worst = max(e["err_u"] for e in log)
    print(f"\nloopback world={world}: {len(log)} (grid, operator) cases, worst displacement error {worst:.2e}")
    assert worst < 1e-9


def test_loopback_matches_single_rank_bitwise_operator():
    """Same grid on 1, 2 and 4 loopback ranks: the assembled operator rows are bitwise equal and the
    CG histories identical to within +-1 (the dot products are summed in a different order)."""
    kw = dict(NX=17, NY=6, NZ=12, bc_type=M.BC_BENDING, lx=10., ly=1., lz=1., ts=2)
    ref = M.MacroC(M.Config(**kw))
    r1 = [ref.time_step(t) for t in range(2)]
    ref.assembly_jac()
    A1 = ref.get_matrix_blocks(); u1 = ref.get_vec(M.VEC_U)
    ref.close()
    for world in (2, 4):
        out = {}

        def fn(comm, world=world):
            uid = comm.bcast(M.loopback_id(world) if comm.rank == 0 else None)
            m = M.MacroC(M.Config(pz=world, px=1, py=1, **kw), rank=comm.rank, nranks=world, unique_id=uid)
            logs = [m.time_step(t) for t in range(2)]
            m.assembly_jac()
            got = comm.gather((m.get_matrix_blocks(), m.get_vec(M.VEC_U), logs))
            m.close()
            if comm.rank == 0:
                out["A"] = np.concatenate([g[0] for g in got]); out["u"] = np.concatenate([g[1] for g in got])
                out["logs"] = got[0][2]

        run_world(world, fn)
        assert np.array_equal(out["A"], A1)
        for a, b in zip(out["logs"], r1):
            assert a["newton_its"] == b["newton_its"]
            assert all(abs(x - y) <= 1 for x, y in zip(a["ksp_its"], b["ksp_its"]))
        assert np.abs(out["u"] - u1).max() <= 1e-5 * np.abs(u1).max()
