"""A/B of the element contraction (north_star: DMMA "only if ncu shows a win over FFMA"): measured DFMA and DMMA
rates, then Ke = sum_gp B^T C B of all elements of an N^3 grid in the sparsity-aware DFMA form and as dense
mma.m8n8k4.f64 tiles.   ncu: -k regex:'k_ab_dfma|k_ab_dmma'"""
import sys
sys.path.insert(0, ".")
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, material=M.MAT_PER_GP))
print("DFMA rate TFLOP/s", m.fp64_probe(), " DMMA rate TFLOP/s", m.dmma_probe(), flush=True)
m.apply_bc_on_u(-1e-3); m.set_strains(); m.homogenize()
ne = (N - 1) ** 3
for v, name, fma in ((0, "DFMA sparsity-aware (2160 FMA/gp)", 17280), (1, "DMMA dense (24 DMMA/gp)", 192 * 256), (2, "DMMA upper tiles (18 DMMA/gp)", 144 * 256)):
    ms, _ = m.contraction_ab(v, reps=3)
    print(f"{name}: {ms:.3f} ms, {ne / ms / 1e3:.1f} M elements/s, executed {2 * fma * ne / ms / 1e9:.2f} TFLOP/s, useful {2 * 17280 * ne / ms / 1e9:.2f} TFLOP/s", flush=True)
m.close()
