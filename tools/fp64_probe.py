"""DFMA rate with register operands and (MACROC_FP64_PROBE_CONST=1) with a constant-bank multiplier."""
import sys
sys.path.insert(0, ".")
import macroc_b200 as M
m = M.MacroC(M.Config(NX=8, NY=8, NZ=8, bc_type=M.BC_BENDING))
m.apply_bc_on_u(-1e-3); m.set_strains(); m.assembly_res()      # binds the constants
print("TFLOP/s", m.fp64_probe())
