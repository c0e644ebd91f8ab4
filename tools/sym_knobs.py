"""One line per (variant, hint) of the symmetric-storage SpMV: MACROC_SYM_VARIANT / MACROC_SYM_HINT are read at create."""
import os, statistics, sys
sys.path.insert(0, ".")
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for variant, hint in [(0, 0), (0, 1), (0, 2), (0, 4), (0, 5), (6, 0), (6, 1), (7, 0), (1, 0), (3, 1)]:
    os.environ["MACROC_SYM_VARIANT"] = str(variant); os.environ["MACROC_SYM_HINT"] = str(hint)
    m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, op=M.OP_ASSEMBLED_SYM))
    m.assembly_jac()
    m.time_kernel(8, 3)
    t = statistics.median(m.time_kernel(8, 1) for _ in range(10))
    print(f"variant {variant} hint {hint}: spmv {t:.3f} ms", flush=True)
    m.close()
