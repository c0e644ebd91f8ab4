"""Debug helper: the symmetric operator on loopback ranks vs the oracle, SpMV only, per rank."""
import sys, threading
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import macroc_b200 as M
from oracle import oracle as O
from test_loopback import run_world

def case(NX, NY, NZ, bc, pg, world, extra={}):
    o = O.Oracle(O.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, faithful_ke=0, **extra))
    o.assembly_jac()
    x = np.sin(0.37 * np.arange(3 * NX * NY * NZ)) + 0.1
    y_ref = o.matmult(x).reshape(-1, 3)
    def fn(comm):
        uid = comm.bcast(M.loopback_id(world) if comm.rank == 0 else None)
        cfg = M.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, px=pg[0], py=pg[1], pz=pg[2], op=M.OP_ASSEMBLED_SYM, device=0, **extra)
        m = M.MacroC(cfg, rank=comm.rank, nranks=world, unique_id=uid)
        p = M.partition(cfg, comm.rank, world)
        xs0, ys0, zs0, xm, ym, zm = p["corners"]
        box = np.zeros((NZ, NY, NX), bool); box[zs0:zs0 + zm, ys0:ys0 + ym, xs0:xs0 + xm] = True
        nodes = np.flatnonzero(box.reshape(-1))
        m.assembly_jac()
        for rep in range(3):
            y = m.matmult(x.reshape(-1, 3)[nodes].reshape(-1), M.OP_ASSEMBLED_SYM).reshape(-1, 3)
            err = np.abs(y - y_ref[nodes]).max(axis=1) / np.abs(y_ref).max()
            bad = np.flatnonzero(err > 1e-12)
            got = comm.gather((comm.rank, p["corners"], float(err.max()), [(int(nodes[b] % NX), int(nodes[b] // NX % NY), int(nodes[b] // (NX * NY))) for b in bad[:6]]))
            if comm.rank == 0:
                for g in got:
                    print(rep, g, flush=True)
        m.close()
    run_world(world, fn)

case(12, 10, 9, 0, (2, 2, 2), 8)
case(13, 9, 11, 1, (0, 0, 0), 8, dict(lx=4., lz=4.))
