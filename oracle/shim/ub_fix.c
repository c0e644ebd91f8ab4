/*
 * oracle/shim/ub_fix.c -- TEST INFRASTRUCTURE.
 * The reference's get_displacement() (src/bcs.c:52-58) computes
 * U = U_MAX * (time_s*dt / final_time) and then falls off the end without a
 * return statement: undefined behaviour.  Compilers of the reference's era
 * left U in xmm0 so the caller saw it; gcc 13 -O0 returns garbage
 * (movq %rax,%xmm0).  oracle/Makefile therefore compiles bcs.c with
 * -Dget_displacement=get_displacement_ub (renaming the broken definition, the
 * source file itself untouched) and links this definition of the intended
 * value instead ("get_displacement has no return").
 */
extern double dt, final_time;       /* include/macroc.h:76 (PetscReal) */

double get_displacement(int time_s)
{
    double time = time_s * dt;
    return -1.0 * (time / final_time);      /* U_MAX = -1.0, macroc.h:51 */
}
