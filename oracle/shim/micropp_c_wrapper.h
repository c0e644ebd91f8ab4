/*
 * oracle/shim/micropp_c_wrapper.h -- stand-in for MicroPP's C wrapper (not
 * shipped with the reference; signatures inferred from the call sites,
 * SURVEY.md section 2.4).  TEST INFRASTRUCTURE.  The "micro problem" is the
 * fixed homogenised linear-elastic law the north_star prescribes:
 * stress = D strain, ctan = D, D = isotropic(E, nu) of material 0.
 */
#ifndef ORACLE_SHIM_MICROPP_H
#define ORACLE_SHIM_MICROPP_H
void micropp_C_material_set(int id, double E, double nu, double Ka, double Sy, int type);
void micropp_C_material_print(int id);
void micropp_C_create3(int ngp, int size[3], int type, double *params);
void micropp_C_print_info(void);
void micropp_C_set_strain3(int gp, double *strain);
void micropp_C_get_stress3(int gp, double *stress);
void micropp_C_get_ctan3(int gp, double *ctan);
void micropp_C_homogenize(void);
void micropp_C_update_vars(void);
int micropp_C_get_non_linear_gps(void);
double micropp_C_get_f_trial_max(void);
int micropp_C_get_sigma_cost3(int gp);
int micropp_C_is_non_linear(int gp);
#endif
