"""Launches the kernels whose ncu captures are kept under profiles/ (round 2): the symmetric-storage SpMV, the
full-storage SpMV, the matrix-free apply, the per-element Jacobian with a uniform and with per-Gauss-point
tangents (full and symmetric layout).
  ncu --set full --clock-control none --import-source on -k regex:'k_spmv_sym|k_spmv_tma|k_apply_mf_march|k_assemble_nodes' ...
"""
import sys
sys.path.insert(0, ".")
import macroc_b200 as M
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
R = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, op=M.OP_ASSEMBLED_SYM))
m.apply_bc_on_u(-1e-3); m.set_strains(); m.assembly_res(); m.assembly_jac()
print("spmv_sym ms", m.time_kernel(8, R))
m.set_operator(M.OP_ASSEMBLED); m.assembly_jac()
print("spmv_full ms", m.time_kernel(0, R))
m.set_operator(M.OP_MATRIX_FREE); m.assembly_jac()
print("apply_matrix_free ms", m.time_kernel(1, R))
print("jacobian_per_element (uniform tangent, full layout) ms", m.time_kernel(7, R))
print("jacobian_per_element (uniform tangent, symmetric layout) ms", m.time_kernel(17, R))
m.close()
m = M.MacroC(M.Config(NX=N, NY=N, NZ=N, bc_type=M.BC_BENDING, material=M.MAT_PER_GP))
m.apply_bc_on_u(-1e-3); m.set_strains(); m.homogenize()
print("jacobian_per_element (per-GP tangents, full layout) ms", m.time_kernel(7, R))
print("jacobian_per_element (per-GP tangents, symmetric layout) ms", m.time_kernel(17, R))
m.close()
