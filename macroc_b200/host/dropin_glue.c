/*
 * dropin_glue.c -- the reference-side binding of include/macroc_b200.h.
 *
 * Compiled TOGETHER WITH the reference's own, unmodified src/main.c, init.c, forces.c, output.c
 * and util.c (the `dropin` build recipe): this file takes the place of src/assembly.c and
 * src/bcs.c and forwards every function they define (include/macroc.h:130-155) to the C ABI.
 * main.c's Newton loop (src/main.c:53-82), init.c's option parsing and printing, forces.c's
 * reaction force and output.c's VTU writer run as they are; MicroPP (here: its linear-elastic
 * stand-in) stays the constitutive model and is fed through the reference's own
 * micropp_C_set_strain3 / get_stress3 / get_ctan3 calls.
 *
 * Host Vecs stay the reference's: u, b, du are copied across the boundary where the reference
 * reads or writes them (main.c does VecNorm(b) and VecAXPY(u, 1, du) on the host).
 *
 * MACROC_DROPIN_MATERIAL = uniform (default): the device evaluates sigma = D eps, C = D itself
 *                                             (north_star's fixed homogenised D)
 *                        = per_gp           : stresses and tangents are pulled from MicroPP for
 *                                             every Gauss point and uploaded (MACROC_MAT_PER_GP)
 * One process / one GPU: the PETSc stand-in this is linked against is serial.
 */
#include "macroc.h"

#include "macroc_b200.h"

static macroc_ctx *g_ctx = NULL;
static int g_per_gp = 0;
static double *g_strain = NULL, *g_stress = NULL, *g_ctan = NULL;
static int64_t g_ngp = 0;

#define B200(call)                                                                              \
	do {                                                                                    \
		int _rc = (call);                                                               \
		if (_rc) {                                                                      \
			fprintf(stderr, "macroc_b200: %s -> %d: %s\n", #call, _rc, macroc_last_error(g_ctx)); \
			exit(_rc);      /* main.c ignores every ierr: stop instead of computing on */ \
		}                                                                               \
	} while (0)

/* everything init() has fixed by the time it calls bc_init (src/init.c:47-171) */
static int glue_context(void)
{
	if (g_ctx)
		return 0;
	macroc_config cfg;
	macroc_default_config(&cfg);
	cfg.NX = NX; cfg.NY = NY; cfg.NZ = NZ;
	cfg.px = cfg.py = cfg.pz = 1;
	cfg.lx = lx; cfg.ly = ly; cfg.lz = lz;
	cfg.bc_type = bc_type;
	cfg.ts = ts; cfg.dt = dt; cfg.final_time = final_time; cfg.vtu_freq = vtu_freq;
	cfg.newton_max_its = newton_max_its;
	cfg.newton_min_tol = newton_min_tol;
	cfg.newton_rel_tol = newton_rel_tol;
	PetscReal rtol, abstol, dtol;
	PetscInt maxits;
	KSPGetTolerances(ksp, &rtol, &abstol, &dtol, &maxits);               /* init.c:146-157 */
	cfg.ksp_rtol = rtol; cfg.ksp_abstol = abstol; cfg.ksp_dtol = dtol; cfg.ksp_maxits = maxits;
	PetscReal mat[4] = { 1.0e7, 0.25, 1.0e4, 1.0e7 };                    /* init.c:31 */
	PetscInt nmax = 4;
	PetscOptionsGetRealArray(NULL, NULL, "-micro_mat_1", mat, &nmax, NULL);
	cfg.E = mat[0]; cfg.nu = mat[1];
	const char *m = getenv("MACROC_DROPIN_MATERIAL");
	g_per_gp = m && strcmp(m, "per_gp") == 0;
	cfg.material = g_per_gp ? MACROC_MAT_PER_GP : MACROC_MAT_UNIFORM;
	cfg.op = MACROC_OP_ASSEMBLED;
	const char *op = getenv("MACROC_DROPIN_OPERATOR");
	if (op && strcmp(op, "matrix_free") == 0) cfg.op = MACROC_OP_MATRIX_FREE;
	if (op && strcmp(op, "sym") == 0) cfg.op = MACROC_OP_ASSEMBLED_SYM;
	int rc = macroc_create(&cfg, 0, 1, NULL, &g_ctx);
	if (rc) {
		fprintf(stderr, "macroc_b200: macroc_create -> %d: %s\n", rc, macroc_last_error(NULL));
		exit(rc);               /* no CPU fallback */
	}
	g_ngp = (int64_t)(NX - 1) * (NY - 1) * (NZ - 1) * NGP;
	g_strain = malloc(sizeof(double) * NVOI * g_ngp);
	if (g_per_gp) {
		g_stress = malloc(sizeof(double) * NVOI * g_ngp);
		g_ctan = malloc(sizeof(double) * NVOI * NVOI * g_ngp);
	}
	return 0;
}

/* ---- src/bcs.c ---------------------------------------------------------------------------- */

PetscErrorCode bc_init(DM da, PetscInt **_index_dirichlet, PetscInt *_nbcs,
		       PetscInt **_index_dirichlet_positive, PetscInt *_nbcs_positive)
{
	(void)da;
	int rc = glue_context();
	if (rc)
		return rc;
	macroc_config cfg;
	macroc_default_config(&cfg);
	cfg.NX = NX; cfg.NY = NY; cfg.NZ = NZ; cfg.px = cfg.py = cfg.pz = 1;
	cfg.lx = lx; cfg.ly = ly; cfg.lz = lz; cfg.bc_type = bc_type;
	int32_t n = 0;
	B200(macroc_bc_lists(&cfg, 0, 1, NULL, NULL, &n));
	PetscInt *idx = malloc(sizeof(PetscInt) * (n > 0 ? n : 1));
	B200(macroc_bc_lists(&cfg, 0, 1, idx, NULL, &n));
	PetscInt npos = 0;
	for (int i = 0; i < n; ++i)
		if (idx[i] >= 0)
			npos++;
	PetscInt *pos = malloc(sizeof(PetscInt) * (npos > 0 ? npos : 1));
	for (int i = 0, k = 0; i < n; ++i)
		if (idx[i] >= 0)
			pos[k++] = idx[i];
	*_index_dirichlet = idx; *_nbcs = n;
	*_index_dirichlet_positive = pos; *_nbcs_positive = npos;
	return 0;
}

PetscErrorCode bc_finish(PetscInt *idx)
{
	free(idx);
	free(index_dirichlet_positive);
	free(g_strain); free(g_stress); free(g_ctan);
	int rc = macroc_destroy(g_ctx);
	g_ctx = NULL;
	return rc;
}

double get_displacement(int time_s)
{
	return macroc_get_displacement(g_ctx, time_s);
}

PetscErrorCode apply_bc_on_u(double U, Vec u)
{
	PetscScalar *a;
	VecGetArray(u, &a);
	B200(macroc_set_vec(g_ctx, MACROC_VEC_U, a));
	B200(macroc_apply_bc_on_u(g_ctx, U));
	B200(macroc_get_vec(g_ctx, MACROC_VEC_U, a));
	VecRestoreArray(u, &a);
	return 0;
}

/* ---- src/assembly.c ----------------------------------------------------------------------- */

PetscErrorCode set_strains()
{
	PetscScalar *a;
	VecGetArray(u, &a);                          /* main.c updated u on the host (VecAXPY) */
	B200(macroc_set_vec(g_ctx, MACROC_VEC_U, a));
	VecRestoreArray(u, &a);
	B200(macroc_set_strains(g_ctx, 1));
	int64_t ngp = 0;
	B200(macroc_get_strain_stress(g_ctx, g_strain, NULL, &ngp));
	for (int64_t gpi = 0; gpi < ngp; ++gpi)      /* assembly.c:58-59 */
		micropp_C_set_strain3((int)gpi, &g_strain[gpi * NVOI]);
	return 0;
}

PetscErrorCode assembly_res(Vec b)
{
	if (g_per_gp) {                              /* assembly.c:148-149 */
		for (int64_t gpi = 0; gpi < g_ngp; ++gpi)
			micropp_C_get_stress3((int)gpi, &g_stress[gpi * NVOI]);
		B200(macroc_set_gp_data(g_ctx, g_stress, NULL));
	}
	double norm;
	B200(macroc_assembly_res(g_ctx, &norm));
	PetscScalar *a;
	VecGetArray(b, &a);
	B200(macroc_get_vec(g_ctx, MACROC_VEC_B, a));
	VecRestoreArray(b, &a);
	return 0;
}

PetscErrorCode assembly_jac(Mat A)
{
	(void)A;                                     /* the operator lives in HBM */
	if (g_per_gp) {                              /* assembly.c:91-92 */
		for (int64_t gpi = 0; gpi < g_ngp; ++gpi)
			micropp_C_get_ctan3((int)gpi, &g_ctan[gpi * NVOI * NVOI]);
		B200(macroc_set_gp_data(g_ctx, NULL, g_ctan));
	}
	B200(macroc_assembly_jac(g_ctx));
	return 0;
}

PetscErrorCode solve_Ax(KSP ksp, Vec b, Vec x)
{
	(void)ksp; (void)b;                          /* b is the residual assembly_res left on the device */
	int its = 0;
	double rnorm = 0.;
	B200(macroc_solve_Ax(g_ctx, &its, &rnorm));
	PetscScalar *a;
	VecGetArray(x, &a);
	B200(macroc_get_vec(g_ctx, MACROC_VEC_DU, a));
	VecRestoreArray(x, &a);
	PetscPrintf(PETSC_COMM_WORLD, "KSP : |Ax - b|/|Ax| = %e\tIts = %d\n", rnorm, its);   /* assembly.c:188-189 */
	return 0;
}

void calc_B(int gp, double B[6][NPE * DIM])
{
	macroc_calc_B(gp, &B[0][0]);
}
