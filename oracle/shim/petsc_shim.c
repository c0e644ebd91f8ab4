/*
 * oracle/shim/petsc_shim.c -- one-process implementation of the PETSc calls
 * the reference makes (petscksp.h in this directory).  TEST INFRASTRUCTURE.
 *
 * Semantics implemented (PETSc's published behaviour):
 *  - DMDA 3-D, dof interleaved, natural ordering on one process; Q1 hex
 *    elements x-fastest with the counter-clockwise node order;
 *  - DMCreateMatrix(AIJ): full 27-point x dof x dof pattern, zeros stored;
 *  - MatZeroRowsColumns(diag): rows and columns zeroed, diag on the diagonal;
 *  - KSPCG + PCJACOBI, left preconditioning, preconditioned-norm test,
 *    KSPConvergedDefault with rtol/abstol/dtol.
 * Dump hook (the reference has none): if MACROC_SHIM_DUMP=<prefix> is set,
 * every KSPSolve writes <prefix>_A<n>.bin, _b<n>.bin, _x<n>.bin and every
 * destroyed global vector writes <prefix>_vec<id>.bin (id = creation order:
 * 0=u, 1=b, 2=du per init.c:96-98).
 */
#include "petscksp.h"

#include <stdarg.h>
#include <time.h>

struct _p_L2G { PetscInt n; PetscInt *idx; };
struct _p_DM {
    PetscInt M, N, P, dof, s;
    PetscInt nel; PetscInt *elems;
    struct _p_L2G l2g;
    int nglobal_vecs;
};
struct _p_Vec { PetscInt n; double *a; int global_id; int ignore_neg; };
struct _p_Mat { PetscInt n; int64_t *rowptr; PetscInt *col; double *val; };
struct _p_PC { const char *type; };
struct _p_KSP {
    Mat A; double rtol, abstol, dtol; PetscInt maxits;
    PetscInt its; double rnorm; const char *type; struct _p_PC pc; int nsolves;
};

static int g_argc;
static char **g_argv;
static const char *dump_prefix(void) { return getenv("MACROC_SHIM_DUMP"); }

/* ---- sys ---------------------------------------------------------------- */
PetscErrorCode PetscInitialize(int *argc, char ***args, const char *file, const char *help)
{
    (void)file; (void)help;
    g_argc = *argc; g_argv = *args;
    return 0;
}
PetscErrorCode PetscFinalize(void) { return 0; }

PetscErrorCode PetscPrintf(MPI_Comm comm, const char *fmt, ...)
{
    (void)comm;
    va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap);
    return 0;
}
PetscErrorCode PetscSynchronizedPrintf(MPI_Comm comm, const char *fmt, ...)
{
    (void)comm;
    va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap);
    return 0;
}
PetscErrorCode PetscSynchronizedFlush(MPI_Comm comm, FILE *f) { (void)comm; fflush(f); return 0; }
PetscErrorCode PetscFOpen(MPI_Comm comm, const char *name, const char *mode, FILE **f)
{
    (void)comm; *f = fopen(name, mode); return *f ? 0 : 65;
}
PetscErrorCode PetscFClose(MPI_Comm comm, FILE *f) { (void)comm; if (f) fclose(f); return 0; }
PetscErrorCode PetscFPrintf(MPI_Comm comm, FILE *f, const char *fmt, ...)
{
    (void)comm;
    va_list ap; va_start(ap, fmt); vfprintf(f, fmt, ap); va_end(ap);
    return 0;
}
PetscErrorCode PetscSNPrintf(char *str, size_t len, const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(str, len, fmt, ap); va_end(ap);
    return 0;
}

static const char *opt_find(const char *name)
{
    for (int i = 1; i + 1 < g_argc; ++i)
        if (strcmp(g_argv[i], name) == 0) return g_argv[i + 1];
    return NULL;
}
PetscErrorCode PetscOptionsGetReal(void *o, const char *pre, const char *name, PetscReal *v, PetscBool *set)
{
    (void)o; (void)pre;
    const char *s = opt_find(name);
    if (s) *v = atof(s);
    if (set) *set = s != NULL;
    return 0;
}
PetscErrorCode PetscOptionsGetInt(void *o, const char *pre, const char *name, PetscInt *v, PetscBool *set)
{
    (void)o; (void)pre;
    const char *s = opt_find(name);
    if (s) *v = atoi(s);
    if (set) *set = s != NULL;
    return 0;
}
PetscErrorCode PetscOptionsGetRealArray(void *o, const char *pre, const char *name, PetscReal *v, PetscInt *n, PetscBool *set)
{
    (void)o; (void)pre;
    const char *s = opt_find(name);
    if (set) *set = s != NULL;
    if (!s) { *n = 0; return 0; }
    char *buf = strdup(s), *tok, *save = NULL;
    PetscInt k = 0;
    for (tok = strtok_r(buf, ",", &save); tok && k < *n; tok = strtok_r(NULL, ",", &save)) v[k++] = atof(tok);
    *n = k;
    free(buf);
    return 0;
}

/* ---- MPI, one process --------------------------------------------------- */
static size_t mpi_size(MPI_Datatype t) { return t == MPI_INT ? 4 : 8; }
int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
int MPI_Gather(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, int root, MPI_Comm c)
{
    (void)rc; (void)rt; (void)root; (void)c;
    memcpy(r, s, (size_t)sc * mpi_size(st));
    return 0;
}
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{
    (void)op; (void)root; (void)c;
    memcpy(r, s, (size_t)n * mpi_size(t));
    return 0;
}
double MPI_Wtime(void)
{
    struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* ---- DMDA --------------------------------------------------------------- */
PetscErrorCode DMDACreate3d(MPI_Comm comm, DMBoundaryType bx, DMBoundaryType by, DMBoundaryType bz,
                            DMDAStencilType st, PetscInt M, PetscInt N, PetscInt P,
                            PetscInt m, PetscInt n, PetscInt p, PetscInt dof, PetscInt s,
                            const PetscInt *lx, const PetscInt *ly, const PetscInt *lz, DM *da)
{
    (void)comm; (void)bx; (void)by; (void)bz; (void)st; (void)m; (void)n; (void)p; (void)lx; (void)ly; (void)lz;
    DM d = (DM)calloc(1, sizeof(*d));
    d->M = M; d->N = N; d->P = P; d->dof = dof; d->s = s;
    *da = d;
    return 0;
}
PetscErrorCode DMSetMatType(DM dm, MatType t) { (void)dm; (void)t; return 0; }
PetscErrorCode DMSetFromOptions(DM dm)
{
    PetscOptionsGetInt(NULL, NULL, "-da_grid_x", &dm->M, NULL);
    PetscOptionsGetInt(NULL, NULL, "-da_grid_y", &dm->N, NULL);
    PetscOptionsGetInt(NULL, NULL, "-da_grid_z", &dm->P, NULL);
    return 0;
}
PetscErrorCode DMSetUp(DM dm)
{
    PetscInt M = dm->M, N = dm->N, P = dm->P;
    dm->nel = (M - 1) * (N - 1) * (P - 1);
    dm->elems = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(dm->nel > 0 ? dm->nel : 1) * 8);
    PetscInt c = 0;
    for (PetscInt k = 0; k < P - 1; ++k)
        for (PetscInt j = 0; j < N - 1; ++j)
            for (PetscInt i = 0; i < M - 1; ++i) {
                PetscInt n0 = i + j * M + k * M * N;
                PetscInt *e = dm->elems + 8 * c++;
                e[0] = n0; e[1] = n0 + 1; e[2] = n0 + 1 + M; e[3] = n0 + M;
                for (int q = 0; q < 4; ++q) e[4 + q] = e[q] + M * N;
            }
    dm->l2g.n = M * N * P * dm->dof;
    dm->l2g.idx = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)dm->l2g.n);
    for (PetscInt i = 0; i < dm->l2g.n; ++i) dm->l2g.idx[i] = i;
    return 0;
}
PetscErrorCode DMDestroy(DM *dm) { if (*dm) { free((*dm)->elems); free((*dm)->l2g.idx); free(*dm); *dm = NULL; } return 0; }

static Vec vec_new(PetscInt n)
{
    Vec v = (Vec)calloc(1, sizeof(*v));
    v->n = n; v->a = (double *)calloc((size_t)n, sizeof(double)); v->global_id = -1;
    return v;
}
PetscErrorCode DMCreateGlobalVector(DM dm, Vec *v)
{
    *v = vec_new(dm->M * dm->N * dm->P * dm->dof);
    (*v)->global_id = dm->nglobal_vecs++;
    return 0;
}
PetscErrorCode DMCreateLocalVector(DM dm, Vec *v) { *v = vec_new(dm->M * dm->N * dm->P * dm->dof); return 0; }
PetscErrorCode DMGetLocalVector(DM dm, Vec *v) { return DMCreateLocalVector(dm, v); }

PetscErrorCode DMDAGetInfo(DM da, PetscInt *dim, PetscInt *M, PetscInt *N, PetscInt *P,
                           PetscInt *m, PetscInt *n, PetscInt *p, PetscInt *dof, PetscInt *s,
                           DMBoundaryType *bx, DMBoundaryType *by, DMBoundaryType *bz, DMDAStencilType *st)
{
    if (dim) *dim = 3;
    if (M) *M = da->M;
    if (N) *N = da->N;
    if (P) *P = da->P;
    if (m) *m = 1;
    if (n) *n = 1;
    if (p) *p = 1;
    if (dof) *dof = da->dof;
    if (s) *s = da->s;
    if (bx) *bx = DM_BOUNDARY_NONE;
    if (by) *by = DM_BOUNDARY_NONE;
    if (bz) *bz = DM_BOUNDARY_NONE;
    if (st) *st = DMDA_STENCIL_BOX;
    return 0;
}
PetscErrorCode DMDAGetElementsSizes(DM da, PetscInt *mx, PetscInt *my, PetscInt *mz)
{
    if (mx) *mx = da->M - 1;
    if (my) *my = da->N - 1;
    if (mz) *mz = da->P - 1;
    return 0;
}
PetscErrorCode DMDAGetCorners(DM da, PetscInt *x, PetscInt *y, PetscInt *z, PetscInt *m, PetscInt *n, PetscInt *p)
{
    if (x) *x = 0;
    if (y) *y = 0;
    if (z) *z = 0;
    if (m) *m = da->M;
    if (n) *n = da->N;
    if (p) *p = da->P;
    return 0;
}
PetscErrorCode DMDAGetGhostCorners(DM da, PetscInt *x, PetscInt *y, PetscInt *z, PetscInt *m, PetscInt *n, PetscInt *p)
{
    return DMDAGetCorners(da, x, y, z, m, n, p);
}
PetscErrorCode DMDAGetElements(DM da, PetscInt *nel, PetscInt *nen, const PetscInt **e)
{
    *nel = da->nel; *nen = 8; *e = da->elems;
    return 0;
}
PetscErrorCode DMGlobalToLocalBegin(DM dm, Vec g, InsertMode mode, Vec l)
{
    (void)dm; (void)mode;
    memcpy(l->a, g->a, sizeof(double) * (size_t)g->n);
    return 0;
}
PetscErrorCode DMGlobalToLocalEnd(DM dm, Vec g, InsertMode mode, Vec l) { (void)dm; (void)g; (void)mode; (void)l; return 0; }
PetscErrorCode DMLocalToGlobalBegin(DM dm, Vec l, InsertMode mode, Vec g)
{
    (void)dm;
    for (PetscInt i = 0; i < g->n; ++i) {
        if (mode == ADD_VALUES) g->a[i] += l->a[i];
        else g->a[i] = l->a[i];
    }
    return 0;
}
PetscErrorCode DMLocalToGlobalEnd(DM dm, Vec l, InsertMode mode, Vec g) { (void)dm; (void)g; (void)mode; (void)l; return 0; }
PetscErrorCode DMGetLocalToGlobalMapping(DM dm, ISLocalToGlobalMapping *l2g) { *l2g = &dm->l2g; return 0; }
PetscErrorCode ISLocalToGlobalMappingGetIndices(ISLocalToGlobalMapping l2g, const PetscInt **idx) { *idx = l2g->idx; return 0; }
PetscErrorCode ISLocalToGlobalMappingRestoreIndices(ISLocalToGlobalMapping l2g, const PetscInt **idx) { (void)l2g; *idx = NULL; return 0; }

/* ---- Vec ---------------------------------------------------------------- */
static void dump_doubles(const char *suffix, int id, const double *a, size_t n)
{
    const char *pre = dump_prefix();
    if (!pre) return;
    char name[4096];
    snprintf(name, sizeof(name), "%s_%s%d.bin", pre, suffix, id);
    FILE *f = fopen(name, "wb");
    if (!f) return;
    fwrite(a, sizeof(double), n, f);
    fclose(f);
}
PetscErrorCode VecSetOption(Vec v, VecOption op, PetscBool flg) { if (op == VEC_IGNORE_NEGATIVE_INDICES) v->ignore_neg = flg; return 0; }
PetscErrorCode VecZeroEntries(Vec v) { memset(v->a, 0, sizeof(double) * (size_t)v->n); return 0; }
PetscErrorCode VecGetArray(Vec v, PetscScalar **a) { *a = v->a; return 0; }
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a) { (void)v; *a = NULL; return 0; }
PetscErrorCode VecDestroy(Vec *v)
{
    if (!*v) return 0;
    if ((*v)->global_id >= 0) dump_doubles("vec", (*v)->global_id, (*v)->a, (size_t)(*v)->n);
    free((*v)->a); free(*v); *v = NULL;
    return 0;
}
PetscErrorCode VecSetValues(Vec v, PetscInt n, const PetscInt *ix, const PetscScalar *y, InsertMode mode)
{
    for (PetscInt i = 0; i < n; ++i) {
        if (ix[i] < 0) { if (v->ignore_neg) continue; return 63; }
        if (mode == ADD_VALUES) v->a[ix[i]] += y[i];
        else v->a[ix[i]] = y[i];
    }
    return 0;
}
PetscErrorCode VecAssemblyBegin(Vec v) { (void)v; return 0; }
PetscErrorCode VecAssemblyEnd(Vec v) { (void)v; return 0; }
PetscErrorCode VecNorm(Vec v, NormType t, PetscReal *val)
{
    double s = 0.;
    if (t == NORM_2) { for (PetscInt i = 0; i < v->n; ++i) s += v->a[i] * v->a[i]; s = sqrt(s); }
    else for (PetscInt i = 0; i < v->n; ++i) s += fabs(v->a[i]);
    *val = s;
    return 0;
}
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x)
{
    for (PetscInt i = 0; i < y->n; ++i) y->a[i] += a * x->a[i];
    return 0;
}
PetscErrorCode VecScale(Vec v, PetscScalar a)
{
    for (PetscInt i = 0; i < v->n; ++i) v->a[i] = a * v->a[i];
    return 0;
}

/* ---- Mat ---------------------------------------------------------------- */
PetscErrorCode DMCreateMatrix(DM dm, Mat *Aout)
{
    PetscInt M = dm->M, N = dm->N, P = dm->P, dof = dm->dof;
    Mat A = (Mat)calloc(1, sizeof(*A));
    A->n = M * N * P * dof;
    A->rowptr = (int64_t *)malloc(sizeof(int64_t) * ((size_t)A->n + 1));
    int64_t nnz = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int64_t q = 0;
        for (PetscInt k = 0; k < P; ++k)
            for (PetscInt j = 0; j < N; ++j)
                for (PetscInt i = 0; i < M; ++i)
                    for (PetscInt d = 0; d < dof; ++d) {
                        PetscInt row = (i + j * M + k * M * N) * dof + d;
                        if (pass) A->rowptr[row] = q;
                        for (PetscInt kk = k - 1; kk <= k + 1; ++kk)
                            for (PetscInt jj = j - 1; jj <= j + 1; ++jj)
                                for (PetscInt ii = i - 1; ii <= i + 1; ++ii) {
                                    if (ii < 0 || ii >= M || jj < 0 || jj >= N || kk < 0 || kk >= P) continue;
                                    for (PetscInt e = 0; e < dof; ++e) {
                                        if (pass) A->col[q] = (ii + jj * M + kk * M * N) * dof + e;
                                        q++;
                                    }
                                }
                    }
        if (!pass) {
            nnz = q;
            A->col = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)nnz);
            A->val = (double *)calloc((size_t)nnz, sizeof(double));
        } else
            A->rowptr[A->n] = q;
    }
    *Aout = A;
    return 0;
}
PetscErrorCode MatZeroEntries(Mat A) { memset(A->val, 0, sizeof(double) * (size_t)A->rowptr[A->n]); return 0; }
static int64_t mat_find(Mat A, PetscInt row, PetscInt col)
{
    int64_t lo = A->rowptr[row], hi = A->rowptr[row + 1] - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) / 2;
        if (A->col[mid] == col) return mid;
        if (A->col[mid] < col) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}
PetscErrorCode MatSetValuesLocal(Mat A, PetscInt nr, const PetscInt *ir, PetscInt nc, const PetscInt *ic,
                                 const PetscScalar *v, InsertMode mode)
{
    for (PetscInt r = 0; r < nr; ++r)
        for (PetscInt c = 0; c < nc; ++c) {
            int64_t q = mat_find(A, ir[r], ic[c]);
            if (q < 0) return 63;
            if (mode == ADD_VALUES) A->val[q] += v[r * nc + c];
            else A->val[q] = v[r * nc + c];
        }
    return 0;
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode MatZeroRowsColumns(Mat A, PetscInt n, const PetscInt *rows, PetscScalar diag, Vec x, Vec b)
{
    (void)x; (void)b;
    char *flag = (char *)calloc((size_t)A->n, 1);
    for (PetscInt i = 0; i < n; ++i) flag[rows[i]] = 1;
    for (PetscInt r = 0; r < A->n; ++r)
        for (int64_t q = A->rowptr[r]; q < A->rowptr[r + 1]; ++q)
            if (flag[r] || flag[A->col[q]]) A->val[q] = (A->col[q] == r) ? diag : 0.;
    free(flag);
    return 0;
}
PetscErrorCode MatDestroy(Mat *A) { if (*A) { free((*A)->rowptr); free((*A)->col); free((*A)->val); free(*A); *A = NULL; } return 0; }

/* ---- KSP ---------------------------------------------------------------- */
PetscErrorCode KSPCreate(MPI_Comm comm, KSP *ksp) { (void)comm; *ksp = (KSP)calloc(1, sizeof(**ksp)); (*ksp)->type = "gmres"; return 0; }
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P) { (void)P; ksp->A = A; return 0; }
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits)
{
    ksp->rtol = rtol; ksp->abstol = abstol; ksp->dtol = dtol; ksp->maxits = maxits;
    return 0;
}
PetscErrorCode KSPGetTolerances(KSP ksp, PetscReal *rtol, PetscReal *abstol, PetscReal *dtol, PetscInt *maxits)
{
    *rtol = ksp->rtol; *abstol = ksp->abstol; *dtol = ksp->dtol; *maxits = ksp->maxits;
    return 0;
}
PetscErrorCode KSPSetType(KSP ksp, KSPType t) { ksp->type = t; return 0; }
PetscErrorCode KSPGetType(KSP ksp, KSPType *t) { *t = ksp->type; return 0; }
PetscErrorCode KSPGetPC(KSP ksp, PC *pc) { *pc = &ksp->pc; return 0; }
PetscErrorCode PCSetType(PC pc, PCType t) { pc->type = t; return 0; }
PetscErrorCode KSPSetFromOptions(KSP ksp)
{
    PetscOptionsGetReal(NULL, NULL, "-ksp_rtol", &ksp->rtol, NULL);
    PetscOptionsGetReal(NULL, NULL, "-ksp_atol", &ksp->abstol, NULL);
    PetscOptionsGetReal(NULL, NULL, "-ksp_divtol", &ksp->dtol, NULL);
    PetscOptionsGetInt(NULL, NULL, "-ksp_max_it", &ksp->maxits, NULL);
    return 0;
}
PetscErrorCode KSPSetUp(KSP ksp) { (void)ksp; return 0; }
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its) { *its = ksp->its; return 0; }
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal *rnorm) { *rnorm = ksp->rnorm; return 0; }
PetscErrorCode KSPDestroy(KSP *ksp) { free(*ksp); *ksp = NULL; return 0; }

static void mat_mult(Mat A, const double *x, double *y)
{
    for (PetscInt r = 0; r < A->n; ++r) {
        double s = 0.;
        for (int64_t q = A->rowptr[r]; q < A->rowptr[r + 1]; ++q) s += A->val[q] * x[A->col[q]];
        y[r] = s;
    }
}
static double dot(PetscInt n, const double *x, const double *y)
{
    double s = 0.;
    for (PetscInt i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

/* KSPSolve_CG, KSP_NORM_PRECONDITIONED, PCJACOBI, zero initial guess. */
PetscErrorCode KSPSolve(KSP ksp, Vec B, Vec X)
{
    Mat A = ksp->A;
    PetscInt n = A->n;
    if (strcmp(ksp->type, KSPCG) != 0 || strcmp(ksp->pc.type ? ksp->pc.type : "", PCJACOBI) != 0) {
        fprintf(stderr, "petsc_shim: only -ksp_type cg -pc_type jacobi is implemented\n");
        return 56;
    }
    if (dump_prefix()) {
        const char *pre = dump_prefix();
        char name[4096];
        snprintf(name, sizeof(name), "%s_A%d.bin", pre, ksp->nsolves);
        FILE *f = fopen(name, "wb");
        if (f) {
            int64_t hdr[2] = {n, A->rowptr[n]};
            fwrite(hdr, sizeof(int64_t), 2, f);
            fwrite(A->rowptr, sizeof(int64_t), (size_t)n + 1, f);
            fwrite(A->col, sizeof(PetscInt), (size_t)A->rowptr[n], f);
            fwrite(A->val, sizeof(double), (size_t)A->rowptr[n], f);
            fclose(f);
        }
        dump_doubles("b", ksp->nsolves, B->a, (size_t)n);
    }
    double *dinv = (double *)malloc(sizeof(double) * (size_t)n);
    double *R = (double *)malloc(sizeof(double) * (size_t)n);
    double *Z = (double *)malloc(sizeof(double) * (size_t)n);
    double *Pv = (double *)malloc(sizeof(double) * (size_t)n);
    double *W = (double *)malloc(sizeof(double) * (size_t)n);
    double *x = X->a;
    for (PetscInt i = 0; i < n; ++i) {
        double d = A->val[mat_find(A, i, i)];
        dinv[i] = (d != 0.) ? 1. / d : 1.;
    }
    for (PetscInt i = 0; i < n; ++i) { x[i] = 0.; R[i] = B->a[i]; }
    for (PetscInt i = 0; i < n; ++i) Z[i] = R[i] * dinv[i];
    double dp = sqrt(dot(n, Z, Z)), dp0 = dp;
    double ttol = ksp->rtol * dp0 > ksp->abstol ? ksp->rtol * dp0 : ksp->abstol;
    ksp->its = 0; ksp->rnorm = dp;
    if (!(dp <= ttol)) {
        double beta = dot(n, Z, R), betaold = 1.;
        PetscInt i = 0;
        do {
            ksp->its = i + 1;
            if (beta == 0.0) break;
            if (!i) memcpy(Pv, Z, sizeof(double) * (size_t)n);
            else {
                double b = beta / betaold;
                for (PetscInt q = 0; q < n; ++q) Pv[q] = Z[q] + b * Pv[q];
            }
            mat_mult(A, Pv, W);
            double dpi = dot(n, Pv, W);
            betaold = beta;
            if (dpi == 0.0) break;
            double a = beta / dpi;
            for (PetscInt q = 0; q < n; ++q) x[q] += a * Pv[q];
            for (PetscInt q = 0; q < n; ++q) R[q] += -a * W[q];
            for (PetscInt q = 0; q < n; ++q) Z[q] = R[q] * dinv[q];
            dp = sqrt(dot(n, Z, Z));
            ksp->rnorm = dp;
            if (dp <= ttol) break;
            if (dp >= ksp->dtol * dp0) break;
            beta = dot(n, Z, R);
            i++;
        } while (i < ksp->maxits);
    }
    dump_doubles("x", ksp->nsolves, x, (size_t)n);
    ksp->nsolves++;
    free(dinv); free(R); free(Z); free(Pv); free(W);
    return 0;
}
