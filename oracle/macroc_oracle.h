/*
 * macroc_oracle.h -- CPU oracle for the MacroC macro-scale FE hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under macroc_b200/ (the product) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * It restates, line by line, the arithmetic of the reference
 *   src/assembly.c (set_strains :25-66, assembly_jac :69-117,
 *                   assembly_res :120-176, solve_Ax :179-192, calc_B :195-254)
 *   src/bcs.c      (:29-362)
 *   src/main.c     (:49-82  time loop + Newton loop)
 *   src/init.c     (:47-64 defaults, :137-157 wg + KSP parameters)
 *   src/forces.c   (:58-166 reaction force)
 * and the PETSc behaviour those calls rely on (DMDA box decomposition, element
 * ownership, local->global map, AIJ 27-point pattern, MatZeroRowsColumns,
 * KSPCG + PCJACOBI with the preconditioned-norm test).  MicroPP is replaced by
 * sigma = D eps, C = D (isotropic linear elasticity) as BASELINE.json's
 * north_star prescribes.
 *
 * PARITY PIN.  The reference ships no golden vectors and
 * PETSc / MicroPP are not installable here, so: "parity unpinned" at the
 * PETSc/MicroPP boundary.  What IS pinned: oracle/Makefile compiles the
 * reference's own src/ C files, unmodified and in place, over a serial
 * PETSc-semantics shim (oracle/shim/) into oracle/_ref/macroc_ref; tests check
 * that this oracle reproduces that binary's |RES| / KSP lines and exported
 * A, b, u bit for bit (tests/test_oracle_vs_ref.py, tests/golden/).
 */
#ifndef MACROC_ORACLE_H
#define MACROC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NGP 8
#define ORC_NPE 8
#define ORC_NVOI 6
#define ORC_DIM 3

enum { ORC_BC_BENDING = 0, ORC_BC_CIRCLE = 1 };   /* include/macroc.h:58 */

typedef struct {
    int NX, NY, NZ;            /* -da_grid_x/y/z        (macroc.h:44-46: 40,3,40) */
    int px, py, pz;            /* -da_processors_x/y/z  (0 = PETSC_DECIDE)        */
    int nranks;                /* number of simulated MPI ranks                   */
    double lx, ly, lz;         /* macroc.h:47-49: 50,1,50                         */
    int bc_type;               /* init.c:64  default BC_CIRCLE                    */
    double E, nu;              /* init.c:31  1e7, 0.25 (MicroPP material 1)       */
    double rtol, abstol, dtol; /* init.c:147 1e-5, 1e-50, 1e4                     */
    int maxits;                /* init.c:148 10000                                */
    double newton_min_tol;     /* macroc.h:37 1e-1                                */
    double newton_rel_tol;     /* macroc.h:36 1e-4                                */
    int newton_max_its;        /* macroc.h:38 5                                   */
    double dt, final_time;     /* macroc.h:43,40 1e-3, 1.0                        */
    int ts;                    /* macroc.h:41 1                                   */
    int faithful_ke;           /* 1: recompute Ae per element with the reference's
                                  4-deep loop (assembly.c:94-99); 0: reuse Ae while
                                  the 8 tangents are bitwise identical (same bits) */
    int nthreads;              /* 1: strict reference order; >1: OpenMP (timing)   */
    int physical_B;            /* 0: reference quirk, calc_B's local dx=dy=dz=1 (assembly.c:198);
                                  1: B of the physical element (process-wide while the context lives) */
} orc_config;

typedef struct orc_ctx orc_ctx;

void orc_default_config(orc_config *cfg);
orc_ctx *orc_create(const orc_config *cfg);
void orc_destroy(orc_ctx *c);

/* --- element level (assembly.c:195-254, :94-99, :151-153) ------------------ */
void orc_calc_B(int gp, double *B /* [6][24] row-major */);
void orc_isotropic_D(double E, double nu, double *D /* [36] row-major */);
void orc_elem_jac(const double *ctan /* [8][36] */, double wg, double *Ae /* [576] */);
void orc_elem_res(const double *stress /* [8][6] */, double wg, double *be /* [24] */);

/* --- DMDA restatement (PETSc semantics) ------------------------------------ */
void orc_proc_grid(const orc_ctx *c, int out[3]);
void orc_corners(const orc_ctx *c, int rank, int out[6]);        /* xs,ys,zs,xm,ym,zm */
void orc_ghost_corners(const orc_ctx *c, int rank, int out[6]);
void orc_elements_sizes(const orc_ctx *c, int rank, int out[3]); /* nex,ney,nez */
int  orc_nelem(const orc_ctx *c, int rank);
const int *orc_elements(const orc_ctx *c, int rank);             /* nelem*8 local ids */
const int *orc_l2g(const orc_ctx *c, int rank);                  /* local dof -> global dof */
int  orc_bc_list(const orc_ctx *c, int rank, const int **idx);   /* returns nbcs; idx incl. -1 */
int  orc_bc_list_positive(const orc_ctx *c, int rank, const int **idx);

int64_t orc_ndof(const orc_ctx *c);
int64_t orc_nnz(const orc_ctx *c);
double  orc_wg(const orc_ctx *c);

/* --- hot path, one call per reference call --------------------------------- */
double orc_get_displacement(const orc_ctx *c, int time_s);        /* bcs.c:52-58  */
int orc_apply_bc_on_u(orc_ctx *c, double U);                      /* bcs.c:29-146 */
int orc_set_strains(orc_ctx *c);                                  /* assembly.c:25-66 */
int orc_homogenize(orc_ctx *c);                                   /* main.c:62 -> sigma=D eps */
int orc_assembly_res(orc_ctx *c, double *norm);                   /* assembly.c:120-176 + main.c:67 */
int orc_assembly_jac(orc_ctx *c);                                 /* assembly.c:69-117 + bcs.c:341-347 */
int orc_solve(orc_ctx *c, int *its, double *rnorm);               /* assembly.c:179-192 (KSPCG+PCJACOBI) */
int orc_update_u(orc_ctx *c);                                     /* main.c:79 */
double orc_calc_force(orc_ctx *c);                                /* forces.c:25-166 */

/* main.c:49-109.  Writes the reference's stdout lines to `log` (may be NULL).
 * Per time step records newton its, KSP its, last |RES|, force. */
typedef struct {
    int newton_its;
    int ksp_its[8];
    double res_norm[8];     /* |RES| printed at each Newton iteration */
    double ksp_rnorm[8];
    int n_res;              /* number of |RES| lines printed           */
    double U, force;
} orc_step_log;
int orc_run(orc_ctx *c, orc_step_log *steps /* [ts] or NULL */, const char *log_path);

/* --- export (natural ordering: dof = 3*(i + NX*(j + NY*k)) + d) ------------- */
enum { ORC_VEC_U = 0, ORC_VEC_DU = 1, ORC_VEC_B = 2 };
void orc_get_vec(const orc_ctx *c, int which, double *out);
void orc_set_vec(orc_ctx *c, int which, const double *in);
/* matrix as 27-slot 3x3 block stencil, natural node order:
 * out[((node*27 + slot)*9) + 3*r + cc], slot = (dz+1)*9 + (dy+1)*3 + (dx+1);
 * slots that fall outside the grid are returned as 0. */
void orc_get_block_stencil(const orc_ctx *c, double *out);
/* raw CSR in PETSc global ordering */
void orc_get_csr(const orc_ctx *c, const int64_t **rowptr, const int32_t **col, const double **val);
void orc_natural_to_petsc(const orc_ctx *c, int32_t *perm /* [nnodes] natural node -> petsc node */);
/* y = A x on natural-order vectors (through the CSR) */
void orc_matmult(const orc_ctx *c, const double *x, double *y);
/* strain / stress of rank r, gpi = ie*8+gp (assembly.c:58) */
const double *orc_strain(const orc_ctx *c, int rank);
const double *orc_stress(const orc_ctx *c, int rank);

/* --- timing helpers for bench.py's cpu_baseline / --impl reference ---------- */
/* n CG iterations' worth of PETSc-shaped work (MatMult, 2 Dot, Norm, 2 AXPY,
 * AYPX, PointwiseMult) on the assembled operator; returns seconds. */
double orc_time_cg_iterations(orc_ctx *c, int n);
double orc_wtime(void);

#ifdef __cplusplus
}
#endif
#endif
