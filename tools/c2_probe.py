"""BASELINE configs[1] (cantilever 128x32x32) with the symmetric, the full and the matrix-free operator: ms per Newton
step and per CG iteration.  ncu --metrics gpu__time_duration.sum -s 3000 -c 60 shows the per-launch times."""
import sys, time
sys.path.insert(0, ".")
import macroc_b200 as M
ops = {"sym": M.OP_ASSEMBLED_SYM, "full": M.OP_ASSEMBLED, "mf": M.OP_MATRIX_FREE}
for name in (sys.argv[1:] or ["sym", "full", "mf"]):
    m = M.MacroC(M.Config(NX=128, NY=32, NZ=32, lx=10., ly=1., lz=1., bc_type=M.BC_BENDING, op=ops[name]))
    for t in (1, 2):
        m.time_step(t)
    m.event_record(0)
    rs = [m.time_step(t) for t in (3, 4, 5, 6, 7)]
    m.event_record(1)
    ms = m.event_elapsed_ms(0, 1) / len(rs)
    its = sum(sum(r["ksp_its"]) for r in rs) / len(rs)
    print(name, "ms per Newton step", round(ms, 3), "CG iterations", its, "us per iteration (whole step / its)", round(1e3 * ms / its, 2), flush=True)
    m.close()
