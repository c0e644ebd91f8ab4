"""Debug helper: time loop of one multi-rank case on loopback ranks, printing the per-step logs."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import macroc_b200 as M
from oracle import oracle as O
from test_loopback import run_world

def case(NX, NY, NZ, bc, pg, world, op, extra={}):
    o = O.Oracle(O.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, ts=3, rtol=1e-12, faithful_ke=0, **extra))
    ologs = o.run()
    def fn(comm):
        uid = comm.bcast(M.loopback_id(world) if comm.rank == 0 else None)
        cfg = M.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, px=pg[0], py=pg[1], pz=pg[2], ts=3, ksp_rtol=1e-12, op=op, device=0, **extra)
        m = M.MacroC(cfg, rank=comm.rank, nranks=world, unique_id=uid)
        logs = [m.time_step(t) for t in range(3)]
        reason = m.ksp_reason()
        u = m.get_vec(M.VEC_U)
        got = comm.gather((logs, reason, float(np.abs(u).max())))
        if comm.rank == 0:
            print("op", op, "oracle", [(l.newton_its, l.ksp_its) for l in ologs], flush=True)
            for r, g in enumerate(got):
                print(" rank", r, [(l["newton_its"], l["ksp_its"], ["%.3e" % x for x in l["res_norm"]]) for l in g[0]], "reason", g[1], "umax", g[2], flush=True)
        m.close()
    run_world(world, fn)

for rep in range(3):
    for op in (0, 2):
        case(12, 10, 9, 0, (2, 2, 2), 8, op)
