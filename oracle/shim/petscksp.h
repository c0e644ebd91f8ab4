/*
 * oracle/shim/petscksp.h -- serial stand-in for the slice of the PETSc C API
 * that the reference (GG1991/macroc, src/*.c) calls.  TEST INFRASTRUCTURE.
 *
 * Purpose: let the reference's own sources compile UNMODIFIED, in place under
 * /root/reference, into oracle/_ref/macroc_ref (see oracle/Makefile), so the
 * oracle restatement can be pinned against the reference's own loops
 * (calc_B, assembly_jac, assembly_res, bc_init_*, the Newton loop in main).
 * One process only: every MPI call degenerates to a copy.  PETSc's real
 * internals are not available in this image; the semantics implemented in
 * petsc_shim.c are the published ones.
 */
#ifndef ORACLE_SHIM_PETSCKSP_H
#define ORACLE_SHIM_PETSCKSP_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef int PetscBool;
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

#define PETSC_TRUE 1
#define PETSC_FALSE 0
#define PETSC_DECIDE (-1)
#define PETSC_COMM_WORLD 0
#define PETSC_COMM_SELF 1
#define MPI_COMM_WORLD 0
#define PETSC_MAX_PATH_LEN 4096
#define PETSC_STDOUT stdout
#define PETSC_ERR_ARG_WRONG 62

#define MPI_INT 4
#define MPI_LONG 8
#define MPI_DOUBLE 108
#define MPI_SUM 0
#define MPI_MAX 1

#define CHKERRQ(ierr) do { if (ierr) return (ierr); } while (0)

typedef struct _p_DM *DM;
typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_KSP *KSP;
typedef struct _p_PC *PC;
typedef struct _p_L2G *ISLocalToGlobalMapping;
typedef const char *KSPType;
typedef const char *PCType;
typedef const char *MatType;

typedef enum { DM_BOUNDARY_NONE } DMBoundaryType;
typedef enum { DMDA_STENCIL_STAR, DMDA_STENCIL_BOX } DMDAStencilType;
typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { NORM_1, NORM_2 } NormType;
typedef enum { MAT_FLUSH_ASSEMBLY, MAT_FINAL_ASSEMBLY } MatAssemblyType;
typedef enum { VEC_IGNORE_OFF_PROC_ENTRIES, VEC_IGNORE_NEGATIVE_INDICES } VecOption;

#define MATAIJ "aij"
#define KSPCG "cg"
#define PCJACOBI "jacobi"

/* sys */
PetscErrorCode PetscInitialize(int *argc, char ***args, const char *file, const char *help);
PetscErrorCode PetscFinalize(void);
PetscErrorCode PetscPrintf(MPI_Comm comm, const char *fmt, ...);
PetscErrorCode PetscSynchronizedPrintf(MPI_Comm comm, const char *fmt, ...);
PetscErrorCode PetscSynchronizedFlush(MPI_Comm comm, FILE *f);
PetscErrorCode PetscFOpen(MPI_Comm comm, const char *name, const char *mode, FILE **f);
PetscErrorCode PetscFClose(MPI_Comm comm, FILE *f);
PetscErrorCode PetscFPrintf(MPI_Comm comm, FILE *f, const char *fmt, ...);
PetscErrorCode PetscSNPrintf(char *str, size_t len, const char *fmt, ...);
PetscErrorCode PetscOptionsGetReal(void *opts, const char *pre, const char *name, PetscReal *v, PetscBool *set);
PetscErrorCode PetscOptionsGetInt(void *opts, const char *pre, const char *name, PetscInt *v, PetscBool *set);
PetscErrorCode PetscOptionsGetRealArray(void *opts, const char *pre, const char *name, PetscReal *v, PetscInt *n, PetscBool *set);

/* MPI, one process */
int MPI_Comm_rank(MPI_Comm c, int *r);
int MPI_Comm_size(MPI_Comm c, int *s);
int MPI_Gather(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
double MPI_Wtime(void);

/* DM / DMDA */
PetscErrorCode DMDACreate3d(MPI_Comm comm, DMBoundaryType bx, DMBoundaryType by, DMBoundaryType bz,
                            DMDAStencilType st, PetscInt M, PetscInt N, PetscInt P,
                            PetscInt m, PetscInt n, PetscInt p, PetscInt dof, PetscInt s,
                            const PetscInt *lx, const PetscInt *ly, const PetscInt *lz, DM *da);
PetscErrorCode DMSetMatType(DM dm, MatType t);
PetscErrorCode DMSetFromOptions(DM dm);
PetscErrorCode DMSetUp(DM dm);
PetscErrorCode DMCreateMatrix(DM dm, Mat *A);
PetscErrorCode DMCreateGlobalVector(DM dm, Vec *v);
PetscErrorCode DMCreateLocalVector(DM dm, Vec *v);
PetscErrorCode DMGetLocalVector(DM dm, Vec *v);
PetscErrorCode DMDestroy(DM *dm);
PetscErrorCode DMDAGetInfo(DM da, PetscInt *dim, PetscInt *M, PetscInt *N, PetscInt *P,
                           PetscInt *m, PetscInt *n, PetscInt *p, PetscInt *dof, PetscInt *s,
                           DMBoundaryType *bx, DMBoundaryType *by, DMBoundaryType *bz, DMDAStencilType *st);
PetscErrorCode DMDAGetElementsSizes(DM da, PetscInt *mx, PetscInt *my, PetscInt *mz);
PetscErrorCode DMDAGetGhostCorners(DM da, PetscInt *x, PetscInt *y, PetscInt *z, PetscInt *m, PetscInt *n, PetscInt *p);
PetscErrorCode DMDAGetCorners(DM da, PetscInt *x, PetscInt *y, PetscInt *z, PetscInt *m, PetscInt *n, PetscInt *p);
PetscErrorCode DMDAGetElements(DM da, PetscInt *nel, PetscInt *nen, const PetscInt **e);
PetscErrorCode DMGlobalToLocalBegin(DM dm, Vec g, InsertMode mode, Vec l);
PetscErrorCode DMGlobalToLocalEnd(DM dm, Vec g, InsertMode mode, Vec l);
PetscErrorCode DMLocalToGlobalBegin(DM dm, Vec l, InsertMode mode, Vec g);
PetscErrorCode DMLocalToGlobalEnd(DM dm, Vec l, InsertMode mode, Vec g);
PetscErrorCode DMGetLocalToGlobalMapping(DM dm, ISLocalToGlobalMapping *l2g);
PetscErrorCode ISLocalToGlobalMappingGetIndices(ISLocalToGlobalMapping l2g, const PetscInt **idx);
PetscErrorCode ISLocalToGlobalMappingRestoreIndices(ISLocalToGlobalMapping l2g, const PetscInt **idx);

/* Vec */
PetscErrorCode VecSetOption(Vec v, VecOption op, PetscBool flg);
PetscErrorCode VecZeroEntries(Vec v);
PetscErrorCode VecGetArray(Vec v, PetscScalar **a);
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a);
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode VecSetValues(Vec v, PetscInt n, const PetscInt *ix, const PetscScalar *y, InsertMode mode);
PetscErrorCode VecAssemblyBegin(Vec v);
PetscErrorCode VecAssemblyEnd(Vec v);
PetscErrorCode VecNorm(Vec v, NormType t, PetscReal *val);
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x);
PetscErrorCode VecScale(Vec v, PetscScalar a);

/* Mat */
PetscErrorCode MatZeroEntries(Mat A);
PetscErrorCode MatSetValuesLocal(Mat A, PetscInt nr, const PetscInt *ir, PetscInt nc, const PetscInt *ic,
                                 const PetscScalar *v, InsertMode mode);
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t);
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t);
PetscErrorCode MatZeroRowsColumns(Mat A, PetscInt n, const PetscInt *rows, PetscScalar diag, Vec x, Vec b);
PetscErrorCode MatDestroy(Mat *A);

/* KSP / PC */
PetscErrorCode KSPCreate(MPI_Comm comm, KSP *ksp);
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P);
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits);
PetscErrorCode KSPGetTolerances(KSP ksp, PetscReal *rtol, PetscReal *abstol, PetscReal *dtol, PetscInt *maxits);
PetscErrorCode KSPSetType(KSP ksp, KSPType t);
PetscErrorCode KSPGetType(KSP ksp, KSPType *t);
PetscErrorCode KSPGetPC(KSP ksp, PC *pc);
PetscErrorCode PCSetType(PC pc, PCType t);
PetscErrorCode KSPSetFromOptions(KSP ksp);
PetscErrorCode KSPSetUp(KSP ksp);
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x);
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its);
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal *rnorm);
PetscErrorCode KSPDestroy(KSP *ksp);

#endif
