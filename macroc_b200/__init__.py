"""macroc_b200 -- B200 (sm_100a) implementation of MacroC's macro-scale FE hot path.

Hex8 residual/Jacobian assembly on the 3-D DMDA grid (reference src/assembly.c),
Dirichlet treatment (src/bcs.c) and the CG+Jacobi solve inside the Newton loop
(src/main.c), as hand-written fp64 CUDA behind the C ABI of include/macroc_b200.h.
"""
from .capi import (  # noqa: F401
    BC_BENDING, BC_CIRCLE, JAC_AUTO, JAC_ELEMENT, MAT_PER_GP, MAT_UNIFORM, OP_ASSEMBLED, OP_ASSEMBLED_SYM, OP_MATRIX_FREE,
    VEC_B, VEC_DU, VEC_U,
    Config, MacroC, MacrocError, bc_lists, build, calc_B, get_unique_id, lib, loopback_id, partition,
)
