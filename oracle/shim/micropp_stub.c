/*
 * oracle/shim/micropp_stub.c -- MicroPP stand-in (TEST INFRASTRUCTURE): the
 * fixed homogenised linear-elastic law of BASELINE.json's north_star.
 * stress = D strain, ctan = D, D = isotropic(E, nu) of material 0
 * (init.c:31: E = 1e7, nu = 0.25), Voigt order (e11 e22 e33 g12 g13 g23).
 */
#include "micropp_c_wrapper.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static double mat_E[2], mat_nu[2];
static double D[36];
static int ngp_total;
static double *strain_gp, *stress_gp;

void micropp_C_material_set(int id, double E, double nu, double Ka, double Sy, int type)
{
    (void)Ka; (void)Sy; (void)type;
    if (id >= 0 && id < 2) { mat_E[id] = E; mat_nu[id] = nu; }
}
void micropp_C_material_print(int id) { printf("material %d : E = %e nu = %e (linear-elastic stand-in)\n", id, mat_E[id], mat_nu[id]); }
void micropp_C_create3(int ngp, int size[3], int type, double *params)
{
    (void)size; (void)type; (void)params;
    double E = mat_E[0], nu = mat_nu[0];
    double lambda = E * nu / ((1. + nu) * (1. - 2. * nu));
    double mu = E / (2. * (1. + nu));
    memset(D, 0, sizeof(D));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) D[i * 6 + j] = lambda + (i == j ? 2. * mu : 0.);
    for (int i = 3; i < 6; ++i) D[i * 6 + i] = mu;
    ngp_total = ngp;
    strain_gp = (double *)calloc((size_t)ngp * 6, sizeof(double));
    stress_gp = (double *)calloc((size_t)ngp * 6, sizeof(double));
}
void micropp_C_print_info(void) { printf("MicroPP stand-in : %d Gauss points, sigma = D eps\n", ngp_total); }
void micropp_C_set_strain3(int gp, double *strain) { memcpy(&strain_gp[(size_t)gp * 6], strain, 6 * sizeof(double)); }
void micropp_C_get_stress3(int gp, double *stress) { memcpy(stress, &stress_gp[(size_t)gp * 6], 6 * sizeof(double)); }
void micropp_C_get_ctan3(int gp, double *ctan) { (void)gp; memcpy(ctan, D, sizeof(D)); }
void micropp_C_homogenize(void)
{
    for (int g = 0; g < ngp_total; ++g)
        for (int i = 0; i < 6; ++i) {
            double t = 0.;
            for (int j = 0; j < 6; ++j) t += D[i * 6 + j] * strain_gp[(size_t)g * 6 + j];
            stress_gp[(size_t)g * 6 + i] = t;
        }
}
void micropp_C_update_vars(void) {}
int micropp_C_get_non_linear_gps(void) { return 0; }
double micropp_C_get_f_trial_max(void) { return 0.; }
int micropp_C_get_sigma_cost3(int gp) { (void)gp; return 0; }
int micropp_C_is_non_linear(int gp) { (void)gp; return 0; }
