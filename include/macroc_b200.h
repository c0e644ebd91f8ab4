/*
 * macroc_b200.h -- C ABI of the B200 (sm_100a) implementation of MacroC's
 * macro-scale FE hot path.
 *
 * The reference (GG1991/macroc) has no plugin/FFI layer: its hot path is a
 * handful of C functions that take PETSc handles and read file-scope globals
 * (reference include/macroc.h:71-155).  This header is the drop-in boundary
 * for that path: one opaque context replaces the globals, and there is one
 * entry point per reference function.  Plain C types only; every function
 * returns a PetscErrorCode-style int (0 = success) and never throws.  One host
 * thread/process per GPU, like one MPI rank per DMDA sub-box in the reference.
 *
 * Vector layout at the boundary is the reference's: the rank's part of the
 * DMDA global vector, i.e. its owned box (macroc_partition) x fastest,
 * dof-interleaved:
 *     v[3*((i - xs) + xm*((j - ys) + ym*(k - zs))) + d].
 * (With -da_processors_x 1 -da_processors_y 1 -da_processors_z P this is the
 * natural ordering restricted to the slab.)
 *
 * There is no CPU fallback: every compute entry point fails with
 * MACROC_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef MACROC_B200_H
#define MACROC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MACROC_OK                0
#define MACROC_ERR_ARG          62   /* PETSC_ERR_ARG_WRONG  */
#define MACROC_ERR_UNSUPPORTED  56   /* PETSC_ERR_SUP        */
#define MACROC_ERR_MEM          55   /* PETSC_ERR_MEM        */
#define MACROC_ERR_NO_DEVICE    97
#define MACROC_ERR_CUDA         98
#define MACROC_ERR_NCCL         99

enum { MACROC_BC_BENDING = 0, MACROC_BC_CIRCLE = 1 };          /* include/macroc.h:58 */
enum { MACROC_VEC_U = 0, MACROC_VEC_DU = 1, MACROC_VEC_B = 2 }; /* include/macroc.h:128 */
enum { MACROC_OP_ASSEMBLED = 0, MACROC_OP_MATRIX_FREE = 1,
       MACROC_OP_ASSEMBLED_SYM = 2 };  /* assembled, symmetric storage: 14 of the 27 slots */
/* where the Gauss-point stress / tangent come from (the MicroPP boundary) */
enum { MACROC_MAT_UNIFORM = 0,   /* sigma = D eps, C = D in registers (north_star's fixed D)    */
       MACROC_MAT_PER_GP = 1 };  /* device arrays strain/stress[ngp*6], ctan[ngp*36], gpi=ie*8+gp */
enum { MACROC_JAC_AUTO = 0,      /* uniform D: class-stencil fill; per-GP: element kernel        */
       MACROC_JAC_ELEMENT = 1 }; /* always the per-element kernel                                 */
/* KSPConvergedReason values used */
enum {
    MACROC_KSP_CONVERGED_RTOL = 2, MACROC_KSP_CONVERGED_ATOL = 3,
    MACROC_KSP_DIVERGED_ITS = -3, MACROC_KSP_DIVERGED_DTOL = -4,
    MACROC_KSP_DIVERGED_INDEFINITE_PC = -8, MACROC_KSP_DIVERGED_NANORINF = -9,
    MACROC_KSP_DIVERGED_INDEFINITE_MAT = -10
};

/* Everything init() fixes (reference src/init.c:47-64,66-83,85-94,137-157). */
typedef struct {
    int32_t NX, NY, NZ;            /* -da_grid_x/y/z          (macroc.h:44-46: 40, 3, 40)   */
    int32_t px, py, pz;            /* -da_processors_x/y/z; 0 = PETSC_DECIDE (PETSc's squarish
                                      factorisation of the rank count).  Any px*py*pz = nranks
                                      is accepted; z-slabs (1,1,P) are the fast path (halo
                                      overlapped with the SpMV)                               */
    double  lx, ly, lz;            /* -lx -ly -lz             (macroc.h:47-49: 50, 1, 50)   */
    int32_t bc_type;               /* -bc_type                (init.c:64: BC_CIRCLE)        */
    int32_t ts;                    /* -ts                     (macroc.h:41: 1)              */
    int32_t vtu_freq;              /* -vtu_freq               (macroc.h:42: -1 = never)     */
    int32_t pad0;
    double  dt, final_time;        /* -dt                     (macroc.h:43,40)              */
    int32_t newton_max_its;        /* -newton_max_its | -new_its (macroc.h:38: 5)           */
    double  newton_min_tol;        /* -newton_min_tol | -new_tol (macroc.h:37: 1e-1)        */
    double  newton_rel_tol;        /* -newton_rel_tol         (macroc.h:36: 1e-4)           */
    double  ksp_rtol, ksp_abstol, ksp_dtol;   /* init.c:147: 1e-5, 1e-50, 1e4               */
    int32_t ksp_maxits;            /* init.c:148: 10000                                      */
    double  E, nu;                 /* -micro_mat_1 E,nu,..    (init.c:31: 1e7, 0.25)        */
    double  D[36];                 /* homogenised tangent, row-major 6x6, Voigt order
                                      (e11 e22 e33 g12 g13 g23); used if use_D != 0          */
    int32_t use_D;
    int32_t op;                    /* MACROC_OP_ASSEMBLED (reference: MATAIJ + MatMult) or
                                      MACROC_OP_MATRIX_FREE for solve_Ax                     */
    int32_t device;                /* CUDA device ordinal, -1 = current                      */
    int32_t material;              /* MACROC_MAT_*                                           */
    int32_t jac_mode;              /* MACROC_JAC_*                                           */
    int32_t physical_B;            /* 0 (default): the reference's calc_B, whose local dx=dy=dz=1
                                      makes B that of a unit cube (assembly.c:198); 1: B of the
                                      physical element (-physical_B 1)                       */
    int32_t strict_fp;             /* 0 (default): production kernels (FMA, tree reductions).
                                      1: verification mode -- the reference's rounding: no FMA contraction,
                                      CSR-order row sums, sequential dots (csrc/strict_fp.cuh).  One rank,
                                      uniform tangent, MACROC_OP_ASSEMBLED; reproduces the reference binary
                                      bit for bit and is orders of magnitude slower (-strict_fp 1)          */
    int32_t reserved[4];
} macroc_config;

typedef struct macroc_ctx macroc_ctx;

/* ---- setup (init.c:25-219 / finish :222-237) ------------------------------ */
int macroc_default_config(macroc_config *cfg);
/* Parses MacroC's/PETSc's command-line keys into cfg (init.c:66-83 and the
 * -da_* / -ksp_* keys DMSetFromOptions/KSPSetFromOptions read, init.c:93,156).
 * README's -new_its/-new_tol are accepted as aliases. Unknown keys are ignored,
 * like the PETSc options database does. */
int macroc_config_from_args(macroc_config *cfg, int argc, const char *const *argv);
/* 128-byte NCCL unique id for multi-rank contexts; rank 0 creates it, the host
 * ships it to the other ranks by any means (file, torch.distributed, MPI). */
int macroc_get_unique_id(void *id128);
/* In-process ranks instead of NCCL (csrc/loopback.h): returns an id that makes the `nranks`
 * contexts created with it -- each by its own host thread, on one device or several -- talk
 * through device-to-device copies and a rank-ordered host sum.  Every multi-rank code path
 * (slabs, general boxes, halos, ghost-plane tiles) runs unchanged; used to check decomposition
 * independence (reference tests/CMakeLists.txt:21-28) on a single GPU.  Not a performance path. */
int macroc_loopback_id(int nranks, void *id128);
int macroc_create(const macroc_config *cfg, int rank, int nranks, const void *id128,
                  macroc_ctx **out);
int macroc_destroy(macroc_ctx *ctx);
const char *macroc_last_error(const macroc_ctx *ctx);   /* ctx may be NULL */

/* ---- host-only partition queries (no GPU needed) --------------------------- */
/* DMDA ownership of `rank` (PETSc: M/m + ((M%m) > i)), its ghost corners and
 * element counts (DMDAGetCorners / GetGhostCorners / GetElementsSizes as used
 * at init.c:167-171).  out[15] = xs,ys,zs,xm,ym,zm, Xs,Ys,Zs,Xm,Ym,Zm, nex,ney,nez */
int macroc_partition(const macroc_config *cfg, int rank, int nranks, int32_t out[15]);
/* bc_init (bcs.c:154-338): Dirichlet GLOBAL dof ids of this rank's ghosted box,
 * -1 padded exactly like index_dirichlet; coef[i]*U is the value
 * apply_bc_on_u (bcs.c:61-146) inserts.  Call with idx == NULL to get *n. */
int macroc_bc_lists(const macroc_config *cfg, int rank, int nranks, int32_t *idx,
                    double *coef, int32_t *n);

/* ---- the hot path: one entry point per reference function ------------------ */
double macroc_get_displacement(const macroc_ctx *ctx, int time_s);            /* bcs.c:52-58     */
int macroc_apply_bc_on_u(macroc_ctx *ctx, double U);                          /* bcs.c:29-146    */
/* set_strains (assembly.c:25-66).  Always refreshes the halo of u.  If
 * materialize != 0 the Gauss-point strain and stress = D strain arrays
 * (gpi = ie*8 + gp, assembly.c:58) are written to device memory for a
 * constitutive plug-in / export; the residual kernel recomputes them in
 * registers either way. */
int macroc_set_strains(macroc_ctx *ctx, int materialize);
/* micropp_C_homogenize (main.c:62) stand-in for MACROC_MAT_PER_GP: stress = D strain, ctan = D
 * for every owned Gauss point, on the device.  A GPU material model replaces this call by
 * writing the arrays of macroc_gp_arrays itself.  No-op for MACROC_MAT_UNIFORM. */
int macroc_homogenize(macroc_ctx *ctx);
/* DEVICE pointers to the Gauss-point arrays of the rank's DMDA-owned elements (the data the
 * reference exchanges with MicroPP at gpi = ie*8+gp, assembly.c:58,91,148).  On the device they
 * are SoA over elements so that consecutive lanes touch consecutive doubles:
 *     strain/stress[(gp*6 + i) * pitch + ie],   ctan[((gp*6 + k)*6 + l) * pitch + ie],
 * ie = element of the rank's local box, ie = ex + lnex*(ey + lney*ez) with lnex, lney the local
 * box's elements per row / rows per layer (ghosted node extents of macroc_partition minus one);
 * the DMDA-owned elements are ex < nex, ey < ney, ez < nez (for z-slabs: all of them, in
 * DMDAGetElements order); the extra layers belong to the upper neighbours and are refreshed by
 * Gauss-point halos.  pitch is returned in *pitch, *n_gp = 8 * lnex*lney*nez.  A GPU material
 * model reads strain and writes stress and ctan of the owned elements in place. */
int macroc_gp_arrays(macroc_ctx *ctx, double **strain, double **stress, double **ctan, int64_t *n_gp,
                     int64_t *pitch);
/* host -> device copies into those arrays in the reference's AoS view, stress[gpi*6 + i],
 * ctan[gpi*36 + 6k + l] (tests, CPU material models); NULL = leave as is */
int macroc_set_gp_data(macroc_ctx *ctx, const double *stress_host, const double *ctan_host);
int macroc_assembly_res(macroc_ctx *ctx, double *norm);    /* assembly.c:120-176 + VecNorm main.c:67 */
int macroc_assembly_jac(macroc_ctx *ctx);                  /* assembly.c:69-117 + bcs.c:341-347      */
int macroc_solve_Ax(macroc_ctx *ctx, int *its, double *rnorm);   /* assembly.c:179-192 (KSPCG+PCJACOBI) */
int macroc_ksp_reason(const macroc_ctx *ctx, int *reason);
/* switch solve_Ax between the assembled and the matrix-free operator (cfg.op) at run time;
 * call assembly_jac again before the next solve */
int macroc_set_operator(macroc_ctx *ctx, int op);
int macroc_update_u(macroc_ctx *ctx);                      /* VecAXPY(u,1,du) main.c:79 */
int macroc_calc_B(int gp, double *B /* [6][24] */);        /* assembly.c:195-254 (host, constants) */
/* forces.c:25-166.  Deviation: with the uniform tangent the Gauss-point stresses are re-evaluated from the
 * CURRENT u (after a halo refresh); the reference reads the stresses of the last micropp_C_homogenize,
 * i.e. of the u before the last Newton update.  The two coincide whenever the Newton loop ended on its
 * residual test (it then breaks right after a homogenize with no update in between); they differ if it ran
 * out of newton_max_its.  With MACROC_MAT_PER_GP the stored stresses are used, exactly like the reference. */
int macroc_calc_force(macroc_ctx *ctx, double *force);

/* One Newton loop of one time step (main.c:53-82) with the reference's
 * control flow; res_norms gets the |RES| of every iteration (n_res of them),
 * ksp_its the CG iteration count of every solve.  Arrays may be NULL; otherwise
 * they must hold newton_max_its + 1 entries. */
int macroc_time_step(macroc_ctx *ctx, int time_s, int *newton_its, double *res_norms,
                     int *n_res, int *ksp_its, double *ksp_rnorms);

/* ---- vectors / matrix across the boundary ---------------------------------- */
int64_t macroc_local_ndof(const macroc_ctx *ctx);          /* 3 * owned nodes            */
int64_t macroc_global_ndof(const macroc_ctx *ctx);
int macroc_set_vec(macroc_ctx *ctx, int which, const double *host);   /* H2D, owned slab */
int macroc_get_vec(macroc_ctx *ctx, int which, double *host);         /* D2H, owned slab */
/* Assembled operator as 27 3x3 blocks per owned node:
 * out[((node*27 + slot)*9) + 3*r + c], slot = (dz+1)*9 + (dy+1)*3 + (dx+1). */
int macroc_get_matrix_blocks(macroc_ctx *ctx, double *host);
/* y = A x with the assembled (op = 0) or matrix-free (op = 1) operator; x, y on
 * the host in boundary layout (halo exchanged internally). */
int macroc_matmult(macroc_ctx *ctx, int op, const double *x_host, double *y_host);
/* Gauss-point strain / stress in the reference's AoS view ([gpi*6 + i], gpi = ie*8+gp) */
int macroc_get_strain_stress(macroc_ctx *ctx, double *strain, double *stress, int64_t *n_gp);

/* write_pvtu (src/output.c:25-267): <prefix>.pvtu (rank 0) and <prefix>-subdo-<rank>.vtu with the
 * ghosted box's points, the rank's hex cells (type 12, local ghosted numbering), displ, part,
 * cost / non-linear (0: no MicroPP) and the element integrals of strain and stress
 * (sum_gp value*wg, output.c:230,247), same ASCII formats as the reference. */
int macroc_write_pvtu(macroc_ctx *ctx, const char *file_prefix);

/* ---- measurement hooks ------------------------------------------------------ */
/* Runs `reps` launches of one kernel family on the context's stream with data
 * already resident in HBM and returns the mean device time per launch in ms
 * (CUDA events on that stream).  what: 0 SpMV assembled, 1 apply matrix-free,
 * 2 one full PCG iteration (assembled), 3 Jacobian assembly, 4 residual,
 * 5 one full PCG iteration (matrix-free), 7 per-element Jacobian kernel,
 * 8 SpMV with symmetric storage, 9 one full PCG iteration with it.
 * flush_l2 != 0 writes a >L2 buffer between launches.  Clobbers b, du and the
 * KSP work vectors (2, 5) -- a measurement hook, not part of the solve path. */
int macroc_time_kernel(macroc_ctx *ctx, int what, int reps, int flush_l2, double *ms_mean);
/* Measured DFMA rate of the device (register-resident FMA chains, best of 5 launches), in
 * TFLOP/s: the denominator of the FP64-pipe fractions quoted for the matrix-free apply and the
 * element assembly (no FP64 figure is in MEASURED_PEAKS.json). */
int macroc_fp64_probe(macroc_ctx *ctx, double *tflops);
/* Measured rate of the fp64 tensor-core instruction (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4), 8 independent
 * accumulator tiles per warp, in TFLOP/s (512 flop per instruction). */
int macroc_dmma_probe(macroc_ctx *ctx, double *tflops);
/* A/B of the element contraction Ke = sum_gp B^T C_gp B wg (src/assembly.c:94-99) over all stored elements,
 * tangents from the per-Gauss-point arrays (needs MACROC_MAT_PER_GP + macroc_homogenize): variant 0 = the
 * sparsity-aware DFMA form the product kernels use, 1 = dense mma.m8n8k4.f64 tiles (24 DMMA per Gauss point),
 * 2 = the same with the upper tiles only (18 DMMA).  Mean milliseconds of `reps` launches after one warm-up;
 * the matrices of the first n_full elements are copied to full_host[n_full][24][24] (row-major).
 * A measurement hook (north_star: "DMMA ... only if ncu shows a win over FFMA"), not part of the solve path. */
int macroc_contraction_ab(macroc_ctx *ctx, int variant, int reps, int n_full, double *full_host, double *ms_mean);
/* how the CG dot products are reduced over the ranks: 0 one rank, 1 ncclAllReduce, 2 peer-mapped
 * mailboxes inside the reduction kernels (csrc/cg_mbox.cuh), 3 loopback host sum */
int macroc_allreduce_path(const macroc_ctx *ctx);
uint64_t macroc_launch_count(const macroc_ctx *ctx);       /* kernels launched so far */
/* CUDA-event stopwatch on the context's stream (slots 0..7): record, then
 * elapsed(a, b) synchronises on b and returns the device time between them. */
int macroc_event_record(macroc_ctx *ctx, int slot);
int macroc_event_elapsed_ms(macroc_ctx *ctx, int slot_a, int slot_b, double *ms);
/* Live profile of the dominant kernel: while enabled, every `stride`-th operator
 * application inside solve_Ax is bracketed by CUDA events on the launch stream;
 * get returns the mean device time per application and the sample count. */
int macroc_profile_enable(macroc_ctx *ctx, int enable, int stride);
int macroc_profile_get(macroc_ctx *ctx, double *apply_ms_mean, int64_t *samples);
/* ... and the device time of whole solve_Ax calls with their CG iteration count (the extra iterations
 * launched after convergence are no-ops but are inside the time): ms per PCG iteration = ratio. */
int macroc_profile_get_solve(macroc_ctx *ctx, double *solve_ms_total, int64_t *iterations);
int macroc_device_synchronize(macroc_ctx *ctx);
int macroc_version(void);

#ifdef __cplusplus
}
#endif
#endif
