/*
 * macroc_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see macroc_oracle.h).
 *
 * Every function cites the reference file:line it restates.  Compile with
 * -ffp-contract=off so the arithmetic is plain IEEE fp64 in source order (the
 * reference's CMake default build has no optimisation flags).
 *
 * Parity pin: reference sources compiled over oracle/shim (serial PETSc
 * semantics) -> oracle/_ref/macroc_ref; PETSc/MicroPP internals themselves
 * are not available here: "parity unpinned" at that boundary.
 */
#include "macroc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NGP ORC_NGP
#define NPE ORC_NPE
#define NVOI ORC_NVOI
#define DIM ORC_DIM

/* include/macroc.h:51-52 */
static const double U_MAX = -1.0;
static const double CONSTXG = 0.577350269189626;

/* Natural-coordinate signs of the 8 hex nodes; the Gauss points use the same
 * order (include/macroc.h:61-69: xg[k] = CONSTXG * sign[k]). */
static const int SGN[8][3] = {
    {-1, -1, -1}, {+1, -1, -1}, {+1, +1, -1}, {-1, +1, -1},
    {-1, -1, +1}, {+1, -1, +1}, {+1, +1, +1}, {-1, +1, +1}};

typedef struct {
    int pi, pj, pk;
    int xs, ys, zs, xm, ym, zm;      /* DMDAGetCorners       */
    int Xs, Ys, Zs, Xm, Ym, Zm;      /* DMDAGetGhostCorners  */
    int nex, ney, nez, nelem;        /* DMDAGetElementsSizes */
    int *eix;                        /* DMDAGetElements: nelem*8 local ghosted node ids */
    int *l2g;                        /* ISLocalToGlobalMapping: local dof -> global dof */
    int64_t node_off;                /* first PETSc-global node id owned */
    int nbcs, nbcs_positive;
    int *index_dirichlet, *index_dirichlet_positive;
    double *strain, *stress;         /* nelem*8*6, gpi = ie*8+gp (assembly.c:58) */
} orc_rank;

struct orc_ctx {
    orc_config cfg;
    int px, py, pz, nranks;
    int *lx_, *ly_, *lz_;            /* nodes per process along each axis */
    int *ox, *oy, *oz;               /* start node per process            */
    int *ownx, *owny, *ownz;         /* node coordinate -> process coord  */
    orc_rank *rk;
    double dx, dy, dz, wg, rad;      /* init.c:137-141 */
    double D[36];
    int64_t nnodes, ndof, nnz;
    int64_t *rowptr;
    int32_t *col;
    double *val;
    double *u, *du, *b;              /* PETSc global ordering */
    double *dinv, *r, *z, *p, *w;    /* KSP work vectors      */
};

double orc_wtime(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

void orc_default_config(orc_config *cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->NX = 40; cfg->NY = 3; cfg->NZ = 40;            /* macroc.h:44-46 */
    cfg->lx = 50.0; cfg->ly = 1.0; cfg->lz = 50.0;      /* macroc.h:47-49 */
    cfg->px = cfg->py = cfg->pz = 0; cfg->nranks = 1;
    cfg->bc_type = ORC_BC_CIRCLE;                       /* init.c:64      */
    cfg->E = 1.0e7; cfg->nu = 0.25;                     /* init.c:31      */
    cfg->rtol = 1.0e-5; cfg->abstol = 1.0e-50; cfg->dtol = 1.0e4; cfg->maxits = 10000; /* init.c:147-148 */
    cfg->newton_min_tol = 1.0e-1; cfg->newton_rel_tol = 1.0e-4; cfg->newton_max_its = 5; /* macroc.h:36-38 */
    cfg->dt = 0.001; cfg->final_time = 1.0; cfg->ts = 1; /* macroc.h:40-43 */
    cfg->faithful_ke = 1; cfg->nthreads = 1;
}

/* ------------------------------------------------------------------------- */
/* Element level                                                             */
/* ------------------------------------------------------------------------- */

/* assembly.c:195-254.  Note the local dx=dy=dz=1 at :198: B is that of a unit
 * cube whatever lx/NX is. */
/* element edge lengths seen by calc_B: 1,1,1 = the reference (assembly.c:198); the physical
 * dx,dy,dz when a context is created with physical_B (named switch) */
static double g_elem_h[3] = {1., 1., 1.};

void orc_calc_B(int gp, double *B)
{
    const double hx = g_elem_h[0], hy = g_elem_h[1], hz = g_elem_h[2];
    double xi = SGN[gp][0] * CONSTXG, eta = SGN[gp][1] * CONSTXG, zeta = SGN[gp][2] * CONSTXG;
    double dsh[NPE][DIM];
    for (int n = 0; n < NPE; ++n) {
        double fx = 1 + SGN[n][0] * xi, fy = 1 + SGN[n][1] * eta, fz = 1 + SGN[n][2] * zeta;
        dsh[n][0] = SGN[n][0] * fy * fz / 8. * 2. / hx;
        dsh[n][1] = SGN[n][1] * fx * fz / 8. * 2. / hy;
        dsh[n][2] = SGN[n][2] * fx * fy / 8. * 2. / hz;
    }
    memset(B, 0, sizeof(double) * NVOI * NPE * DIM);
#define Bm(r, c) B[(r) * (NPE * DIM) + (c)]
    for (int n = 0; n < NPE; ++n) {        /* rows: e11 e22 e33 g12 g13 g23 (:234-253) */
        Bm(0, n * DIM + 0) = dsh[n][0];
        Bm(1, n * DIM + 1) = dsh[n][1];
        Bm(2, n * DIM + 2) = dsh[n][2];
        Bm(3, n * DIM + 0) = dsh[n][1]; Bm(3, n * DIM + 1) = dsh[n][0];
        Bm(4, n * DIM + 0) = dsh[n][2]; Bm(4, n * DIM + 2) = dsh[n][0];
        Bm(5, n * DIM + 1) = dsh[n][2]; Bm(5, n * DIM + 2) = dsh[n][1];
    }
#undef Bm
}

/* MicroPP stand-in (north_star): isotropic linear elasticity in the Voigt order
 * the B rows define, engineering shears. */
void orc_isotropic_D(double E, double nu, double *D)
{
    double lambda = E * nu / ((1. + nu) * (1. - 2. * nu));
    double mu = E / (2. * (1. + nu));
    memset(D, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            D[i * 6 + j] = lambda + (i == j ? 2. * mu : 0.);
    for (int i = 3; i < 6; ++i) D[i * 6 + i] = mu;
}

/* assembly.c:87-101: Ae[24 i + j] += B[k][i] * ctan[6k+l] * B[l][j] * wg */
void orc_elem_jac(const double *ctan, double wg, double *Ae)
{
    double B[NVOI * NPE * DIM];
    const int N = NPE * DIM;
    memset(Ae, 0, sizeof(double) * N * N);
    for (int gp = 0; gp < NGP; ++gp) {
        const double *C = ctan + gp * NVOI * NVOI;
        orc_calc_B(gp, B);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
                for (int k = 0; k < NVOI; ++k)
                    for (int l = 0; l < NVOI; ++l)
                        Ae[N * i + j] += B[k * N + i] * C[k * NVOI + l] * B[l * N + j] * wg;
    }
}

/* assembly.c:144-154: be[i] += B[j][i] * stress[j] * wg */
void orc_elem_res(const double *stress, double wg, double *be)
{
    double B[NVOI * NPE * DIM];
    const int N = NPE * DIM;
    memset(be, 0, sizeof(double) * N);
    for (int gp = 0; gp < NGP; ++gp) {
        orc_calc_B(gp, B);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < NVOI; ++j)
                be[i] += B[j * N + i] * stress[gp * NVOI + j] * wg;
    }
}

/* ------------------------------------------------------------------------- */
/* DMDA restatement (PETSc semantics)        */
/* ------------------------------------------------------------------------- */

/* PETSc's "squarish" PETSC_DECIDE factorisation of the communicator size for a
 * 3-D DMDA (restated from memory of DMSetUp_DA_3D; UNVERIFIED against a PETSc
 * build -- the z-slab configurations set px,py,pz explicitly). */
static int squarish2(double A, double B, int fixed, int size, int *a, int *b)
{
    int aa = (int)(0.5 + sqrt(A * (double)size / (B * (double)fixed))), bb = 0;
    if (!aa) aa = 1;
    while (aa > 0) { bb = size / (aa * fixed); if (aa * bb * fixed == size) break; aa--; }
    if (!aa) return 0;
    if (A > B && aa < bb) { int t = aa; aa = bb; bb = t; }
    *a = aa; *b = bb;
    return 1;
}

static void decide_proc_grid(int M, int N, int P, int size, int *m_, int *n_, int *p_)
{
    int m = *m_ > 0 ? *m_ : 0, n = *n_ > 0 ? *n_ : 0, p = *p_ > 0 ? *p_ : 0;
    if (m && n && p) return;
    if (!m && n && p) m = size / (n * p);
    else if (m && !n && p) n = size / (m * p);
    else if (m && n && !p) p = size / (m * n);
    else if (!m && !n && p) squarish2(M, N, p, size, &m, &n);
    else if (!m && n && !p) squarish2(M, P, n, size, &m, &p);
    else if (m && !n && !p) squarish2(N, P, m, size, &n, &p);
    else {
        int pm;
        n = (int)(0.5 + pow(((double)N * N) * ((double)size) / ((double)P * M), 1. / 3.));
        if (!n) n = 1;
        while (n > 0) { pm = size / n; if (n * pm == size) break; n--; }
        if (!n) n = 1;
        m = (int)(0.5 + sqrt(((double)M) * ((double)size) / ((double)P * n)));
        if (!m) m = 1;
        while (m > 0) { p = size / (m * n); if (m * n * p == size) break; m--; }
        if (M > P && m < p) { int t = m; m = p; p = t; }
    }
    *m_ = m; *n_ = n; *p_ = p;
}

static void split_axis(int M, int m, int *cnt, int *off, int *own)
{
    int o = 0;
    for (int i = 0; i < m; ++i) {
        cnt[i] = M / m + ((M % m) > i);      /* PETSc ownership rule */
        off[i] = o;
        for (int q = 0; q < cnt[i]; ++q) own[o + q] = i;
        o += cnt[i];
    }
}

static inline int64_t node_petsc(const orc_ctx *c, int i, int j, int k)
{
    int pi = c->ownx[i], pj = c->owny[j], pk = c->ownz[k];
    const orc_rank *r = &c->rk[pi + pj * c->px + pk * c->px * c->py];
    return r->node_off + (i - r->xs) + (int64_t)(j - r->ys) * r->xm + (int64_t)(k - r->zs) * r->xm * r->ym;
}

static void setup_rank(orc_ctx *c, int rank)
{
    orc_rank *r = &c->rk[rank];
    int NX = c->cfg.NX, NY = c->cfg.NY, NZ = c->cfg.NZ;
    r->pi = rank % c->px; r->pj = (rank / c->px) % c->py; r->pk = rank / (c->px * c->py);
    r->xs = c->ox[r->pi]; r->xm = c->lx_[r->pi];
    r->ys = c->oy[r->pj]; r->ym = c->ly_[r->pj];
    r->zs = c->oz[r->pk]; r->zm = c->lz_[r->pk];
    /* ghost corners: stencil width 1, DM_BOUNDARY_NONE (init.c:85-90) */
    r->Xs = r->xs > 0 ? r->xs - 1 : 0;
    r->Ys = r->ys > 0 ? r->ys - 1 : 0;
    r->Zs = r->zs > 0 ? r->zs - 1 : 0;
    int Xe = r->xs + r->xm < NX ? r->xs + r->xm + 1 : NX;
    int Ye = r->ys + r->ym < NY ? r->ys + r->ym + 1 : NY;
    int Ze = r->zs + r->zm < NZ ? r->zs + r->zm + 1 : NZ;
    r->Xm = Xe - r->Xs; r->Ym = Ye - r->Ys; r->Zm = Ze - r->Zs;

    /* DMDAGetElements, Q1 hexes (assembly.c:42,83,140): a rank owns the cells
     * whose upper corner node it owns -> its range starts one node into the
     * lower ghost layer when one exists; x fastest; node order matches SGN. */
    int exs = r->xs, eys = r->ys, ezs = r->zs;
    int exe = r->xs + r->xm, eye = r->ys + r->ym, eze = r->zs + r->zm;
    if (exs != r->Xs) exs -= 1;
    if (eys != r->Ys) eys -= 1;
    if (ezs != r->Zs) ezs -= 1;
    r->nex = exe - exs - 1; r->ney = eye - eys - 1; r->nez = eze - ezs - 1;
    if (r->nex < 0) r->nex = 0;
    if (r->ney < 0) r->ney = 0;
    if (r->nez < 0) r->nez = 0;
    r->nelem = r->nex * r->ney * r->nez;
    r->eix = (int *)malloc(sizeof(int) * (size_t)(r->nelem > 0 ? r->nelem : 1) * 8);
    int cnt = 0;
    for (int k = ezs; k < eze - 1; ++k)
        for (int j = eys; j < eye - 1; ++j)
            for (int i = exs; i < exe - 1; ++i) {
                int a = i - r->Xs, bb = j - r->Ys, cc = k - r->Zs;
                int sx = 1, sy = r->Xm, sz = r->Xm * r->Ym;
                int base = a + bb * sy + cc * sz;
                int *e = &r->eix[cnt * 8];
                e[0] = base;           e[1] = base + sx;
                e[2] = base + sx + sy; e[3] = base + sy;
                e[4] = e[0] + sz; e[5] = e[1] + sz; e[6] = e[2] + sz; e[7] = e[3] + sz;
                cnt++;
            }
    r->strain = (double *)calloc((size_t)(r->nelem > 0 ? r->nelem : 1) * NGP * NVOI, sizeof(double));
    r->stress = (double *)calloc((size_t)(r->nelem > 0 ? r->nelem : 1) * NGP * NVOI, sizeof(double));
}

static void setup_l2g(orc_ctx *c, int rank)
{
    orc_rank *r = &c->rk[rank];
    size_t nl = (size_t)r->Xm * r->Ym * r->Zm;
    r->l2g = (int *)malloc(sizeof(int) * nl * DIM);
    for (int k = 0; k < r->Zm; ++k)
        for (int j = 0; j < r->Ym; ++j)
            for (int i = 0; i < r->Xm; ++i) {
                int64_t g = node_petsc(c, r->Xs + i, r->Ys + j, r->Zs + k);
                size_t l = (size_t)i + (size_t)j * r->Xm + (size_t)k * r->Xm * r->Ym;
                for (int d = 0; d < DIM; ++d) r->l2g[l * DIM + d] = (int)(g * DIM + d);
            }
}

/* bcs.c:198-251 */
static void bc_init_bending(orc_ctx *c, orc_rank *r)
{
    int nx_ghost = r->Xm, ny_ghost = r->Ym, nz_ghost = r->Zm;
    int nbcs = 2 * ny_ghost * nz_ghost * DIM;
    int *ix = (int *)malloc(sizeof(int) * (size_t)(nbcs > 0 ? nbcs : 1));
    for (int q = 0; q < nbcs; ++q) ix[q] = -1;
    int index = 0;
    for (int face = 0; face < 2; ++face) {
        int on = face == 0 ? (r->Xs == 0) : (r->Xs + nx_ghost == c->cfg.NX);
        if (!on) continue;
        int i = face == 0 ? 0 : nx_ghost - 1;
        for (int k = 0; k < nz_ghost; ++k)
            for (int j = 0; j < ny_ghost; ++j)
                for (int d = 0; d < DIM; ++d) {
                    int local_id = i + j * nx_ghost + k * nx_ghost * ny_ghost;
                    ix[index++] = r->l2g[local_id * DIM + d];
                }
    }
    r->index_dirichlet = ix; r->nbcs = nbcs;
}

/* the cell-centre-like circle test shared by bcs.c:132-134, :324-327 and
 * forces.c:138-141 (note the + d/2 offsets) */
static inline int in_circle(const orc_ctx *c, int gi, int gk)
{
    double x = c->cfg.lx / 2. - (gi * c->dx + c->dx / 2.);
    double z = c->cfg.lz / 2. - (gk * c->dz + c->dz / 2.);
    return (x * x + z * z) < (c->rad * c->rad);
}

/* bcs.c:254-338 */
static void bc_init_circle(orc_ctx *c, orc_rank *r)
{
    int nx_ghost = r->Xm, ny_ghost = r->Ym, nz_ghost = r->Zm;
    int si = r->Xs, sj = r->Ys, sk = r->Zs;
    int nbcs = (2 * nx_ghost + 2 * nz_ghost) * DIM + nx_ghost * nz_ghost;
    int *ix = (int *)malloc(sizeof(int) * (size_t)nbcs);
    for (int q = 0; q < nbcs; ++q) ix[q] = -1;
    int index = 0;
#define LID(i, j, k) ((i) + (j) * nx_ghost + (k) * nx_ghost * ny_ghost)
    if (si == 0 && sj == 0)                               /* X=0 & Y=0 along z  (:276-284) */
        for (int k = 0; k < nz_ghost; ++k)
            for (int d = 0; d < DIM; ++d) ix[index++] = r->l2g[LID(0, 0, k) * DIM + d];
    if (si + nx_ghost == c->cfg.NX && sj == 0)            /* X=LX & Y=0 along z (:287-295) */
        for (int k = 0; k < nz_ghost; ++k)
            for (int d = 0; d < DIM; ++d) ix[index++] = r->l2g[LID(nx_ghost - 1, 0, k) * DIM + d];
    if (sk == 0 && sj == 0)                               /* Z=0 & Y=0 along x  (:298-306) */
        for (int i = 1; i < nx_ghost - 1; ++i)
            for (int d = 0; d < DIM; ++d) ix[index++] = r->l2g[LID(i, 0, 0) * DIM + d];
    if (sk + nz_ghost == c->cfg.NZ && sj == 0)            /* Z=LZ & Y=0 along x (:309-317) */
        for (int i = 1; i < nx_ghost - 1; ++i)
            for (int d = 0; d < DIM; ++d) ix[index++] = r->l2g[LID(i, 0, nz_ghost - 1) * DIM + d];
    if (sj + ny_ghost == c->cfg.NY)                       /* circle on Y=LY, dof y (:320-333) */
        for (int i = 0; i < nx_ghost; ++i)
            for (int k = 0; k < nz_ghost; ++k)
                if (in_circle(c, si + i, sk + k))
                    ix[index++] = r->l2g[LID(i, ny_ghost - 1, k) * DIM + 1];
#undef LID
    r->index_dirichlet = ix; r->nbcs = nbcs;
}

/* bcs.c:154-195.  The reference copies positives with the *uncompacted* index
 * (:187-189); valid entries always form a prefix, so this is the same list. */
static void bc_init(orc_ctx *c, orc_rank *r)
{
    if (c->cfg.bc_type == ORC_BC_BENDING) bc_init_bending(c, r);
    else bc_init_circle(c, r);
    int np = 0;
    for (int i = 0; i < r->nbcs; ++i) if (r->index_dirichlet[i] >= 0) np++;
    r->index_dirichlet_positive = (int *)malloc(sizeof(int) * (size_t)(np > 0 ? np : 1));
    int q = 0;
    for (int i = 0; i < r->nbcs; ++i)
        if (r->index_dirichlet[i] >= 0) r->index_dirichlet_positive[q++] = r->index_dirichlet[i];
    r->nbcs_positive = np;
}

static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* DMCreateMatrix for MATAIJ on a box-stencil dof-3 DMDA (init.c:92-95): every
 * node couples with each in-grid neighbour of its 27-point box, 3x3 scalars
 * each, zeros stored explicitly. */
static void build_pattern(orc_ctx *c)
{
    int NX = c->cfg.NX, NY = c->cfg.NY, NZ = c->cfg.NZ;
    c->rowptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(c->ndof + 1));
    int64_t *cnt = (int64_t *)calloc((size_t)c->nnodes, sizeof(int64_t));
    for (int k = 0; k < NZ; ++k)
        for (int j = 0; j < NY; ++j)
            for (int i = 0; i < NX; ++i) {
                int nx_ = (i > 0) + 1 + (i < NX - 1), ny_ = (j > 0) + 1 + (j < NY - 1), nz_ = (k > 0) + 1 + (k < NZ - 1);
                cnt[node_petsc(c, i, j, k)] = (int64_t)nx_ * ny_ * nz_ * DIM;
            }
    c->rowptr[0] = 0;
    for (int64_t n = 0; n < c->nnodes; ++n)
        for (int d = 0; d < DIM; ++d)
            c->rowptr[n * DIM + d + 1] = c->rowptr[n * DIM + d] + cnt[n];
    free(cnt);
    c->nnz = c->rowptr[c->ndof];
    c->col = (int32_t *)malloc(sizeof(int32_t) * (size_t)c->nnz);
    c->val = (double *)calloc((size_t)c->nnz, sizeof(double));
#pragma omp parallel for collapse(2) if (c->cfg.nthreads > 1)
    for (int k = 0; k < NZ; ++k)
        for (int j = 0; j < NY; ++j)
            for (int i = 0; i < NX; ++i) {
                int32_t nb[27 * 3];
                int n = 0;
                for (int dk = -1; dk <= 1; ++dk)
                    for (int dj = -1; dj <= 1; ++dj)
                        for (int di = -1; di <= 1; ++di) {
                            int ii = i + di, jj = j + dj, kk = k + dk;
                            if (ii < 0 || ii >= NX || jj < 0 || jj >= NY || kk < 0 || kk >= NZ) continue;
                            int64_t g = node_petsc(c, ii, jj, kk);
                            for (int d = 0; d < DIM; ++d) nb[n++] = (int32_t)(g * DIM + d);
                        }
                qsort(nb, (size_t)n, sizeof(int32_t), cmp_i32);
                int64_t row0 = node_petsc(c, i, j, k) * DIM;
                for (int d = 0; d < DIM; ++d)
                    memcpy(&c->col[c->rowptr[row0 + d]], nb, sizeof(int32_t) * (size_t)n);
            }
}

orc_ctx *orc_create(const orc_config *cfg)
{
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    c->cfg = *cfg;
    int NX = cfg->NX, NY = cfg->NY, NZ = cfg->NZ;
    c->nranks = cfg->nranks > 0 ? cfg->nranks : 1;
    c->px = cfg->px; c->py = cfg->py; c->pz = cfg->pz;
    decide_proc_grid(NX, NY, NZ, c->nranks, &c->px, &c->py, &c->pz);
    if (c->px * c->py * c->pz != c->nranks || c->px > NX || c->py > NY || c->pz > NZ) {
        free(c);
        return NULL;
    }
    c->lx_ = (int *)malloc(sizeof(int) * c->px); c->ox = (int *)malloc(sizeof(int) * c->px);
    c->ly_ = (int *)malloc(sizeof(int) * c->py); c->oy = (int *)malloc(sizeof(int) * c->py);
    c->lz_ = (int *)malloc(sizeof(int) * c->pz); c->oz = (int *)malloc(sizeof(int) * c->pz);
    c->ownx = (int *)malloc(sizeof(int) * NX); c->owny = (int *)malloc(sizeof(int) * NY); c->ownz = (int *)malloc(sizeof(int) * NZ);
    split_axis(NX, c->px, c->lx_, c->ox, c->ownx);
    split_axis(NY, c->py, c->ly_, c->oy, c->owny);
    split_axis(NZ, c->pz, c->lz_, c->oz, c->ownz);
    c->nnodes = (int64_t)NX * NY * NZ;
    c->ndof = c->nnodes * DIM;
    c->rk = (orc_rank *)calloc((size_t)c->nranks, sizeof(orc_rank));
    int64_t off = 0;
    for (int r = 0; r < c->nranks; ++r) {
        setup_rank(c, r);
        c->rk[r].node_off = off;                 /* rank-contiguous global numbering */
        off += (int64_t)c->rk[r].xm * c->rk[r].ym * c->rk[r].zm;
    }
    /* init.c:137-141 */
    c->dx = cfg->lx / (NX - 1);
    c->dy = cfg->ly / (NY - 1);
    c->dz = cfg->lz / (NZ - 1);
    c->wg = c->dx * c->dy * c->dz / NPE;
    c->rad = 1.;
    orc_isotropic_D(cfg->E, cfg->nu, c->D);
    g_elem_h[0] = cfg->physical_B ? c->dx : 1.; g_elem_h[1] = cfg->physical_B ? c->dy : 1.; g_elem_h[2] = cfg->physical_B ? c->dz : 1.;
    for (int r = 0; r < c->nranks; ++r) { setup_l2g(c, r); bc_init(c, &c->rk[r]); }
    build_pattern(c);
    c->u = (double *)calloc((size_t)c->ndof, sizeof(double));
    c->du = (double *)calloc((size_t)c->ndof, sizeof(double));
    c->b = (double *)calloc((size_t)c->ndof, sizeof(double));
    c->dinv = (double *)calloc((size_t)c->ndof, sizeof(double));
    c->r = (double *)calloc((size_t)c->ndof, sizeof(double));
    c->z = (double *)calloc((size_t)c->ndof, sizeof(double));
    c->p = (double *)calloc((size_t)c->ndof, sizeof(double));
    c->w = (double *)calloc((size_t)c->ndof, sizeof(double));
#ifdef _OPENMP
    if (cfg->nthreads > 1) omp_set_num_threads(cfg->nthreads);
#endif
    return c;
}

void orc_destroy(orc_ctx *c)
{
    if (!c) return;
    for (int r = 0; r < c->nranks; ++r) {
        orc_rank *k = &c->rk[r];
        free(k->eix); free(k->l2g); free(k->index_dirichlet); free(k->index_dirichlet_positive);
        free(k->strain); free(k->stress);
    }
    free(c->rk);
    free(c->lx_); free(c->ly_); free(c->lz_); free(c->ox); free(c->oy); free(c->oz);
    free(c->ownx); free(c->owny); free(c->ownz);
    free(c->rowptr); free(c->col); free(c->val);
    free(c->u); free(c->du); free(c->b); free(c->dinv); free(c->r); free(c->z); free(c->p); free(c->w);
    free(c);
}

void orc_proc_grid(const orc_ctx *c, int out[3]) { out[0] = c->px; out[1] = c->py; out[2] = c->pz; }
void orc_corners(const orc_ctx *c, int rank, int out[6])
{
    const orc_rank *r = &c->rk[rank];
    out[0] = r->xs; out[1] = r->ys; out[2] = r->zs; out[3] = r->xm; out[4] = r->ym; out[5] = r->zm;
}
void orc_ghost_corners(const orc_ctx *c, int rank, int out[6])
{
    const orc_rank *r = &c->rk[rank];
    out[0] = r->Xs; out[1] = r->Ys; out[2] = r->Zs; out[3] = r->Xm; out[4] = r->Ym; out[5] = r->Zm;
}
void orc_elements_sizes(const orc_ctx *c, int rank, int out[3])
{
    out[0] = c->rk[rank].nex; out[1] = c->rk[rank].ney; out[2] = c->rk[rank].nez;
}
int orc_nelem(const orc_ctx *c, int rank) { return c->rk[rank].nelem; }
const int *orc_elements(const orc_ctx *c, int rank) { return c->rk[rank].eix; }
const int *orc_l2g(const orc_ctx *c, int rank) { return c->rk[rank].l2g; }
int orc_bc_list(const orc_ctx *c, int rank, const int **idx)
{
    *idx = c->rk[rank].index_dirichlet;
    return c->rk[rank].nbcs;
}
int orc_bc_list_positive(const orc_ctx *c, int rank, const int **idx)
{
    *idx = c->rk[rank].index_dirichlet_positive;
    return c->rk[rank].nbcs_positive;
}
int64_t orc_ndof(const orc_ctx *c) { return c->ndof; }
int64_t orc_nnz(const orc_ctx *c) { return c->nnz; }
double orc_wg(const orc_ctx *c) { return c->wg; }
const double *orc_strain(const orc_ctx *c, int rank) { return c->rk[rank].strain; }
const double *orc_stress(const orc_ctx *c, int rank) { return c->rk[rank].stress; }

/* ------------------------------------------------------------------------- */
/* Hot path                                                                   */
/* ------------------------------------------------------------------------- */

/* bcs.c:52-58 (the function has no return statement; this is the intended value) */
double orc_get_displacement(const orc_ctx *c, int time_s)
{
    double time = time_s * c->cfg.dt;
    return U_MAX * (time / c->cfg.final_time);
}

/* bcs.c:29-45 -> :61-91 (bending) / :94-146 (circle).  VecSetValues(INSERT)
 * with VEC_IGNORE_NEGATIVE_INDICES (init.c:100). */
int orc_apply_bc_on_u(orc_ctx *c, double U)
{
    for (int rank = 0; rank < c->nranks; ++rank) {
        orc_rank *r = &c->rk[rank];
        double *vals = (double *)malloc(sizeof(double) * (size_t)(r->nbcs > 0 ? r->nbcs : 1));
        int index = 0;
        int nx_ghost = r->Xm, ny_ghost = r->Ym, nz_ghost = r->Zm;
        if (c->cfg.bc_type == ORC_BC_BENDING) {
            if (r->Xs == 0)
                for (int q = 0; q < nz_ghost * ny_ghost * DIM; ++q) vals[index++] = 0.;
            if (r->Xs + nx_ghost == c->cfg.NX)
                for (int q = 0; q < nz_ghost * ny_ghost; ++q)
                    for (int d = 0; d < DIM; ++d) vals[index++] = (d == 1) ? U : 0.;
        } else {
            if (r->Xs == 0 && r->Ys == 0)
                for (int q = 0; q < nz_ghost * DIM; ++q) vals[index++] = 0.;
            if (r->Xs + nx_ghost == c->cfg.NX && r->Ys == 0)
                for (int q = 0; q < nz_ghost * DIM; ++q) vals[index++] = 0.;
            if (r->Zs == 0 && r->Ys == 0)
                for (int i = 1; i < nx_ghost - 1; ++i)
                    for (int d = 0; d < DIM; ++d) vals[index++] = 0.;
            if (r->Zs + nz_ghost == c->cfg.NZ && r->Ys == 0)
                for (int i = 1; i < nx_ghost - 1; ++i)
                    for (int d = 0; d < DIM; ++d) vals[index++] = 0.;
            if (r->Ys + ny_ghost == c->cfg.NY)
                for (int i = 0; i < nx_ghost; ++i)
                    for (int k = 0; k < nz_ghost; ++k)
                        if (in_circle(c, r->Xs + i, r->Zs + k)) vals[index++] = U;
        }
        for (int q = 0; q < r->nbcs; ++q)
            if (r->index_dirichlet[q] >= 0) c->u[r->index_dirichlet[q]] = vals[q];
        free(vals);
    }
    return 0;
}

/* assembly.c:25-66 */
int orc_set_strains(orc_ctx *c)
{
    double Bg[NGP][NVOI * NPE * DIM];
    for (int gp = 0; gp < NGP; ++gp) orc_calc_B(gp, Bg[gp]);
#pragma omp parallel for if (c->cfg.nthreads > 1)
    for (int rank = 0; rank < c->nranks; ++rank) {
        orc_rank *r = &c->rk[rank];
        size_t nl = (size_t)r->Xm * r->Ym * r->Zm * DIM;
        double *u_arr = (double *)malloc(sizeof(double) * nl);
        for (size_t l = 0; l < nl; ++l) u_arr[l] = c->u[r->l2g[l]];     /* DMGlobalToLocal :40-41 */
        for (int ie = 0; ie < r->nelem; ++ie) {
            double u_e[NPE * DIM];
            for (int n = 0; n < NPE; ++n)
                for (int d = 0; d < DIM; ++d) u_e[n * DIM + d] = u_arr[r->eix[ie * NPE + n] * DIM + d];
            for (int gp = 0; gp < NGP; ++gp) {
                double *strain = &r->strain[((size_t)ie * NGP + gp) * NVOI];
                for (int i = 0; i < NVOI; ++i) {
                    double s = 0.;
                    for (int j = 0; j < NPE * DIM; ++j) s += Bg[gp][i * NPE * DIM + j] * u_e[j];
                    strain[i] = s;
                }
            }
        }
        free(u_arr);
    }
    return 0;
}

/* main.c:62 micropp_C_homogenize() -> sigma = D eps (C = D is used directly) */
int orc_homogenize(orc_ctx *c)
{
#pragma omp parallel for if (c->cfg.nthreads > 1)
    for (int rank = 0; rank < c->nranks; ++rank) {
        orc_rank *r = &c->rk[rank];
        for (size_t g = 0; g < (size_t)r->nelem * NGP; ++g) {
            const double *e = &r->strain[g * NVOI];
            double *s = &r->stress[g * NVOI];
            for (int i = 0; i < NVOI; ++i) {
                double t = 0.;
                for (int j = 0; j < NVOI; ++j) t += c->D[i * NVOI + j] * e[j];
                s[i] = t;
            }
        }
    }
    return 0;
}

static double vec_norm2(const orc_ctx *c, const double *x)
{
    double s = 0.;
#pragma omp parallel for reduction(+ : s) if (c->cfg.nthreads > 1)
    for (int64_t i = 0; i < c->ndof; ++i) s += x[i] * x[i];
    return sqrt(s);
}

static double vec_dot(const orc_ctx *c, const double *x, const double *y)
{
    double s = 0.;
#pragma omp parallel for reduction(+ : s) if (c->cfg.nthreads > 1)
    for (int64_t i = 0; i < c->ndof; ++i) s += x[i] * y[i];
    return s;
}

/* assembly.c:120-176, then VecNorm main.c:67 */
int orc_assembly_res(orc_ctx *c, double *norm)
{
    memset(c->b, 0, sizeof(double) * (size_t)c->ndof);                    /* :130 */
    for (int rank = 0; rank < c->nranks; ++rank) {
        orc_rank *r = &c->rk[rank];
        size_t nl = (size_t)r->Xm * r->Ym * r->Zm * DIM;
        double *b_arr = (double *)calloc(nl, sizeof(double));             /* :134-136 */
        for (int ie = 0; ie < r->nelem; ++ie) {
            double be[NPE * DIM];
            orc_elem_res(&r->stress[(size_t)ie * NGP * NVOI], c->wg, be);  /* :144-154 */
            for (int n = 0; n < NPE; ++n)                                   /* :156-161 */
                for (int d = 0; d < DIM; ++d) b_arr[r->eix[ie * NPE + n] * DIM + d] += be[n * DIM + d];
        }
        for (size_t l = 0; l < nl; ++l) c->b[r->l2g[l]] += b_arr[l];      /* DMLocalToGlobal ADD :164-165 */
        free(b_arr);
    }
    for (int rank = 0; rank < c->nranks; ++rank) {                         /* apply_bc_on_res bcs.c:350-362 */
        orc_rank *r = &c->rk[rank];
        for (int q = 0; q < r->nbcs; ++q)
            if (r->index_dirichlet[q] >= 0) c->b[r->index_dirichlet[q]] = 0.;
    }
    for (int64_t i = 0; i < c->ndof; ++i) c->b[i] = -1. * c->b[i];         /* VecScale :173 */
    if (norm) *norm = vec_norm2(c, c->b);
    return 0;
}

static inline int64_t csr_find(const orc_ctx *c, int64_t row, int32_t colv)
{
    int64_t lo = c->rowptr[row], hi = c->rowptr[row + 1] - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) >> 1;
        int32_t v = c->col[mid];
        if (v == colv) return mid;
        if (v < colv) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* assembly.c:69-117 + apply_bc_on_jac bcs.c:341-347 */
int orc_assembly_jac(orc_ctx *c)
{
    const int N = NPE * DIM;
    memset(c->val, 0, sizeof(double) * (size_t)c->nnz);                   /* MatZeroEntries :80 */
    double ctan[NGP * NVOI * NVOI];
    for (int gp = 0; gp < NGP; ++gp) memcpy(&ctan[gp * 36], c->D, sizeof(double) * 36);  /* get_ctan3 :92 */
    double Ae_cached[24 * 24];
    if (!c->cfg.faithful_ke) orc_elem_jac(ctan, c->wg, Ae_cached);
    /* ranks stand in for MPI processes; with nthreads>1 they run concurrently
     * and rows shared across slab interfaces are added atomically */
#pragma omp parallel for schedule(static, 1) if (c->cfg.nthreads > 1)
    for (int rank = 0; rank < c->nranks; ++rank) {
        orc_rank *r = &c->rk[rank];
        double Ae_loc[24 * 24];
        for (int ie = 0; ie < r->nelem; ++ie) {
            const double *Ae = Ae_cached;
            if (c->cfg.faithful_ke) { orc_elem_jac(ctan, c->wg, Ae_loc); Ae = Ae_loc; }   /* :87-101 */
            int ix[24];
            for (int n = 0; n < NPE; ++n)                                                    /* :102-104 */
                for (int d = 0; d < DIM; ++d) ix[n * DIM + d] = r->l2g[r->eix[ie * NPE + n] * DIM + d];
            for (int i = 0; i < N; ++i)                                                      /* MatSetValuesLocal ADD :106 */
                for (int j = 0; j < N; ++j) {
                    int64_t pos = csr_find(c, ix[i], ix[j]);
                    if (c->cfg.nthreads > 1) {
#pragma omp atomic
                        c->val[pos] += Ae[N * i + j];
                    } else
                        c->val[pos] += Ae[N * i + j];
                }
        }
    }
    /* MatZeroRowsColumns(A, n, idx, 1.0, NULL, NULL): A <- M A M + (I - M) */
    unsigned char *mask = (unsigned char *)calloc((size_t)c->ndof, 1);
    for (int rank = 0; rank < c->nranks; ++rank) {
        orc_rank *r = &c->rk[rank];
        for (int q = 0; q < r->nbcs_positive; ++q) mask[r->index_dirichlet_positive[q]] = 1;
    }
#pragma omp parallel for if (c->cfg.nthreads > 1)
    for (int64_t row = 0; row < c->ndof; ++row)
        for (int64_t q = c->rowptr[row]; q < c->rowptr[row + 1]; ++q) {
            int32_t cc = c->col[q];
            if (mask[row] || mask[cc]) c->val[q] = (row == cc) ? 1.0 : 0.0;
        }
    free(mask);
    return 0;
}

static void csr_matmult(const orc_ctx *c, const double *x, double *y)
{
#pragma omp parallel for schedule(static) if (c->cfg.nthreads > 1)
    for (int64_t row = 0; row < c->ndof; ++row) {
        double s = 0.;
        for (int64_t q = c->rowptr[row]; q < c->rowptr[row + 1]; ++q) s += c->val[q] * x[c->col[q]];
        y[row] = s;
    }
}

/* assembly.c:179-192: KSPSolve with KSPCG + PCJACOBI as configured at
 * init.c:146-157.  PETSc semantics restated: left
 * preconditioning, preconditioned residual norm, zero initial guess,
 * KSPConvergedDefault: ttol = max(rtol*dp0, abstol); converged if dp <= ttol,
 * diverged if dp >= dtol*dp0. */
int orc_solve(orc_ctx *c, int *its_out, double *rnorm_out)
{
    const int64_t n = c->ndof;
    const int par = c->cfg.nthreads > 1;
    double *x = c->du, *r = c->r, *z = c->z, *p = c->p, *w = c->w, *dinv = c->dinv;
    /* PCJACOBI setup: inverse diagonal (zero diagonal -> 1) */
#pragma omp parallel for if (par)
    for (int64_t row = 0; row < n; ++row) {
        int64_t pos = csr_find(c, row, (int32_t)row);
        double d = c->val[pos];
        dinv[row] = d != 0. ? 1. / d : 1.;
    }
#pragma omp parallel for if (par)
    for (int64_t i = 0; i < n; ++i) { x[i] = 0.; r[i] = c->b[i]; z[i] = r[i] * dinv[i]; }
    double dp = vec_norm2(c, z), dp0 = dp;
    double ttol = fmax(c->cfg.rtol * dp0, c->cfg.abstol);
    int its = 0;
    if (!(dp <= ttol)) {
        double beta = vec_dot(c, z, r), betaold = 1., dpi;
        for (int i = 0; i < c->cfg.maxits; ++i) {
            its = i + 1;
            if (beta == 0.0) break;                           /* KSP_CONVERGED_ATOL */
            if (i == 0) {
#pragma omp parallel for if (par)
                for (int64_t q = 0; q < n; ++q) p[q] = z[q];
            } else {
                double bb = beta / betaold;
#pragma omp parallel for if (par)
                for (int64_t q = 0; q < n; ++q) p[q] = z[q] + bb * p[q];   /* VecAYPX */
            }
            csr_matmult(c, p, w);
            dpi = vec_dot(c, p, w);
            betaold = beta;
            if (dpi == 0.0) break;                            /* KSP_DIVERGED_INDEFINITE_MAT */
            double a = beta / dpi;
#pragma omp parallel for if (par)
            for (int64_t q = 0; q < n; ++q) x[q] += a * p[q];
#pragma omp parallel for if (par)
            for (int64_t q = 0; q < n; ++q) r[q] += -a * w[q];
#pragma omp parallel for if (par)
            for (int64_t q = 0; q < n; ++q) z[q] = r[q] * dinv[q];
            dp = vec_norm2(c, z);
            if (dp <= ttol) break;                            /* converged */
            if (dp >= c->cfg.dtol * dp0) break;               /* KSP_DIVERGED_DTOL */
            beta = vec_dot(c, z, r);
        }
    }
    if (its_out) *its_out = its;
    if (rnorm_out) *rnorm_out = dp;
    return 0;
}

/* main.c:79 VecAXPY(u, 1., du) */
int orc_update_u(orc_ctx *c)
{
    for (int64_t i = 0; i < c->ndof; ++i) c->u[i] += 1. * c->du[i];
    return 0;
}

/* forces.c:25-50 -> :58-106 (bending) / :115-166 (circle) */
double orc_calc_force(orc_ctx *c)
{
    double force = 0.;
    for (int rank = 0; rank < c->nranks; ++rank) {
        orc_rank *r = &c->rk[rank];
        double mpi_force = 0.;
        if (c->cfg.bc_type == ORC_BC_BENDING) {
            if (r->xs + r->xm == c->cfg.NX)
                for (int ey = 0; ey < r->ney; ++ey)
                    for (int ez = 0; ez < r->nez; ++ez) {
                        int e = (r->nex - 1) + ey * r->nex + ez * (r->nex * r->ney);
                        double ave[NVOI] = {0};
                        for (int gp = 0; gp < NGP; ++gp)
                            for (int i = 0; i < NVOI; ++i) ave[i] += r->stress[((size_t)e * NGP + gp) * NVOI + i];
                        mpi_force += ave[3] * c->dy * c->dz;
                    }
        } else {
            if (r->Ys + r->ym == c->cfg.NY)      /* forces.c:130-133 mixes ghost start with owned count */
                for (int ex = 0; ex < r->nex; ++ex)
                    for (int ez = 0; ez < r->nez; ++ez)
                        if (in_circle(c, r->Xs + ex, r->Zs + ez)) {
                            int e = ex + (r->ney - 1) * r->nex + ez * (r->nex * r->ney);
                            double ave[NVOI] = {0};
                            for (int gp = 0; gp < NGP; ++gp)
                                for (int i = 0; i < NVOI; ++i) ave[i] += r->stress[((size_t)e * NGP + gp) * NVOI + i];
                            mpi_force += ave[1] * c->dx * c->dz;
                        }
        }
        force += mpi_force;
    }
    return force;
}

/* main.c:49-109 */
int orc_run(orc_ctx *c, orc_step_log *steps, const char *log_path)
{
    FILE *f = NULL;
    if (log_path) f = strcmp(log_path, "-") == 0 ? stdout : fopen(log_path, "w");
    double norm = 0., norm_0 = 0.;
    for (int time_s = 0; time_s < c->cfg.ts; ++time_s) {
        if (f) fprintf(f, "\n\nTime Step = %d\n", time_s);
        double U = orc_get_displacement(c, time_s);
        orc_apply_bc_on_u(c, U);
        orc_step_log sl;
        memset(&sl, 0, sizeof(sl));
        sl.U = U;
        int newton_it = 0;
        while (newton_it < c->cfg.newton_max_its) {
            if (f) fprintf(f, "\nNewton Iteration = %d\nHomogenizing MicroPP\n", newton_it);
            orc_set_strains(c);
            orc_homogenize(c);
            if (f) fprintf(f, "Assemblying RHS\n");
            orc_assembly_res(c, &norm);
            if (f) fprintf(f, "|RES| = %e\n", norm);
            if (sl.n_res < 8) sl.res_norm[sl.n_res++] = norm;
            if (newton_it == 0) norm_0 = norm;
            if (norm < c->cfg.newton_min_tol || norm < norm_0 * c->cfg.newton_rel_tol) break;
            orc_assembly_jac(c);
            int its; double rnorm;
            orc_solve(c, &its, &rnorm);
            if (f) fprintf(f, "KSP : |Ax - b|/|Ax| = %e\tIts = %d\n", rnorm, its);
            if (newton_it < 8) { sl.ksp_its[newton_it] = its; sl.ksp_rnorm[newton_it] = rnorm; }
            orc_update_u(c);
            newton_it++;
        }
        sl.newton_its = newton_it;
        sl.force = orc_calc_force(c);
        if (steps) steps[time_s] = sl;
    }
    if (f && f != stdout) fclose(f);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Export                                                                     */
/* ------------------------------------------------------------------------- */

static double *vec_ptr(const orc_ctx *c, int which)
{
    return which == ORC_VEC_U ? c->u : which == ORC_VEC_DU ? c->du : c->b;
}

void orc_natural_to_petsc(const orc_ctx *c, int32_t *perm)
{
    int NX = c->cfg.NX, NY = c->cfg.NY, NZ = c->cfg.NZ;
    for (int k = 0; k < NZ; ++k)
        for (int j = 0; j < NY; ++j)
            for (int i = 0; i < NX; ++i)
                perm[i + (int64_t)NX * (j + (int64_t)NY * k)] = (int32_t)node_petsc(c, i, j, k);
}

void orc_get_vec(const orc_ctx *c, int which, double *out)
{
    const double *v = vec_ptr(c, which);
    int NX = c->cfg.NX, NY = c->cfg.NY, NZ = c->cfg.NZ;
    for (int k = 0; k < NZ; ++k)
        for (int j = 0; j < NY; ++j)
            for (int i = 0; i < NX; ++i) {
                int64_t nat = i + (int64_t)NX * (j + (int64_t)NY * k), g = node_petsc(c, i, j, k);
                for (int d = 0; d < DIM; ++d) out[nat * DIM + d] = v[g * DIM + d];
            }
}

void orc_set_vec(orc_ctx *c, int which, const double *in)
{
    double *v = vec_ptr(c, which);
    int NX = c->cfg.NX, NY = c->cfg.NY, NZ = c->cfg.NZ;
    for (int k = 0; k < NZ; ++k)
        for (int j = 0; j < NY; ++j)
            for (int i = 0; i < NX; ++i) {
                int64_t nat = i + (int64_t)NX * (j + (int64_t)NY * k), g = node_petsc(c, i, j, k);
                for (int d = 0; d < DIM; ++d) v[g * DIM + d] = in[nat * DIM + d];
            }
}

void orc_get_block_stencil(const orc_ctx *c, double *out)
{
    int NX = c->cfg.NX, NY = c->cfg.NY, NZ = c->cfg.NZ;
    memset(out, 0, sizeof(double) * (size_t)c->nnodes * 27 * 9);
#pragma omp parallel for collapse(2) if (c->cfg.nthreads > 1)
    for (int k = 0; k < NZ; ++k)
        for (int j = 0; j < NY; ++j)
            for (int i = 0; i < NX; ++i) {
                int64_t nat = i + (int64_t)NX * (j + (int64_t)NY * k), g = node_petsc(c, i, j, k);
                for (int dk = -1; dk <= 1; ++dk)
                    for (int dj = -1; dj <= 1; ++dj)
                        for (int di = -1; di <= 1; ++di) {
                            int ii = i + di, jj = j + dj, kk = k + dk;
                            if (ii < 0 || ii >= NX || jj < 0 || jj >= NY || kk < 0 || kk >= NZ) continue;
                            int slot = (dk + 1) * 9 + (dj + 1) * 3 + (di + 1);
                            int64_t gn = node_petsc(c, ii, jj, kk);
                            for (int rr = 0; rr < 3; ++rr)
                                for (int cc = 0; cc < 3; ++cc) {
                                    int64_t pos = csr_find(c, g * 3 + rr, (int32_t)(gn * 3 + cc));
                                    out[(nat * 27 + slot) * 9 + rr * 3 + cc] = c->val[pos];
                                }
                        }
            }
}

void orc_get_csr(const orc_ctx *c, const int64_t **rowptr, const int32_t **col, const double **val)
{
    *rowptr = c->rowptr; *col = c->col; *val = c->val;
}

void orc_matmult(const orc_ctx *c, const double *x, double *y)
{
    double *xp = (double *)malloc(sizeof(double) * (size_t)c->ndof);
    double *yp = (double *)malloc(sizeof(double) * (size_t)c->ndof);
    int NX = c->cfg.NX, NY = c->cfg.NY, NZ = c->cfg.NZ;
    for (int k = 0; k < NZ; ++k)
        for (int j = 0; j < NY; ++j)
            for (int i = 0; i < NX; ++i) {
                int64_t nat = i + (int64_t)NX * (j + (int64_t)NY * k), g = node_petsc(c, i, j, k);
                for (int d = 0; d < DIM; ++d) xp[g * DIM + d] = x[nat * DIM + d];
            }
    csr_matmult(c, xp, yp);
    for (int k = 0; k < NZ; ++k)
        for (int j = 0; j < NY; ++j)
            for (int i = 0; i < NX; ++i) {
                int64_t nat = i + (int64_t)NX * (j + (int64_t)NY * k), g = node_petsc(c, i, j, k);
                for (int d = 0; d < DIM; ++d) y[nat * DIM + d] = yp[g * DIM + d];
            }
    free(xp); free(yp);
}

/* n iterations of the un-fused PETSc CG body on the current operator with a
 * fixed synthetic right-hand side; the iterate is discarded (timing only). */
double orc_time_cg_iterations(orc_ctx *c, int niter)
{
    const int64_t n = c->ndof;
    const int par = c->cfg.nthreads > 1;
    double *x = c->du, *r = c->r, *z = c->z, *p = c->p, *w = c->w, *dinv = c->dinv;
#pragma omp parallel for if (par)
    for (int64_t i = 0; i < n; ++i) {
        int64_t pos = csr_find(c, i, (int32_t)i);
        double d = c->val[pos];
        dinv[i] = d != 0. ? 1. / d : 1.;
        x[i] = 0.; r[i] = sin(0.37 * (double)i) + 0.1; z[i] = r[i] * dinv[i]; p[i] = z[i];
    }
    double beta = vec_dot(c, z, r), betaold = beta;
    double t0 = orc_wtime();
    for (int it = 0; it < niter; ++it) {
        double bb = beta / betaold;
#pragma omp parallel for if (par)
        for (int64_t q = 0; q < n; ++q) p[q] = z[q] + bb * p[q];
        csr_matmult(c, p, w);
        double dpi = vec_dot(c, p, w);
        betaold = beta;
        double a = dpi != 0. ? beta / dpi : 0.;
#pragma omp parallel for if (par)
        for (int64_t q = 0; q < n; ++q) x[q] += a * p[q];
#pragma omp parallel for if (par)
        for (int64_t q = 0; q < n; ++q) r[q] += -a * w[q];
#pragma omp parallel for if (par)
        for (int64_t q = 0; q < n; ++q) z[q] = r[q] * dinv[q];
        double dp = vec_norm2(c, z);
        beta = vec_dot(c, z, r);
        if (dp == 0. || beta == 0.) { beta = betaold = 1.; }
    }
    return orc_wtime() - t0;
}
