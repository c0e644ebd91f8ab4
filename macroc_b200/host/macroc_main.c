/*
 * macroc_main.c -- MacroC's host driver over the B200 C ABI.
 *
 * Keeps the reference's command-line surface (-da_grid_x/y/z, -da_processors_z, -ts, -dt,
 * -lx/-ly/-lz, -bc_type, -newton_max_its|-new_its, -newton_min_tol|-new_tol, -newton_rel_tol,
 * -ksp_rtol ...; reference src/init.c:66-83,93,156), its time/Newton loop (src/main.c:49-109)
 * and its stdout lines ("Time Step", "Newton Iteration", "|RES| = %e", "KSP : ... Its = %d",
 * "Elapsed time") so that logs diff against the reference's.  All numerics happen in
 * libmacroc_b200.so (CUDA, sm_100a); this file contains none.
 *
 * Ranks: one process per GPU.  Without MPI in the image the rank comes from the environment
 * (RANK / WORLD_SIZE / LOCAL_RANK, as torchrun or any launcher sets them) and the 128-byte
 * NCCL id travels through the file named by MACROC_ID_FILE (rank 0 writes it).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../../include/macroc_b200.h"

static double wtime(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

#define CHK(call)                                                                        \
    do {                                                                                 \
        int _rc = (call);                                                                \
        if (_rc) {                                                                       \
            fprintf(stderr, "macroc: %s failed (%d): %s\n", #call, _rc, macroc_last_error(ctx)); \
            exit(_rc ? _rc : 1);   /* the process exit tears the communicator down: peers fail instead of waiting */ \
        }                                                                                \
    } while (0)

int main(int argc, char **argv)
{
    macroc_ctx *ctx = NULL;
    macroc_config cfg;
    macroc_default_config(&cfg);
    if (macroc_config_from_args(&cfg, argc, (const char *const *)argv)) {
        fprintf(stderr, "macroc: only -ksp_type cg -pc_type jacobi are implemented\n");
        return 56;
    }
    int rank = env_int("RANK", 0), nranks = env_int("WORLD_SIZE", 1);
    cfg.device = env_int("LOCAL_RANK", 0);
    unsigned char id[128];
    memset(id, 0, sizeof(id));
    const char *id_path = NULL;
    if (nranks > 1) {
        /* The 128-byte NCCL id travels through a file (no MPI in the image).  The file also carries a
         * per-launch nonce (MACROC_LAUNCH_NONCE, else MASTER_PORT): a file left behind by an earlier run
         * at the same path is recognised as stale instead of sending the ranks into a hang inside
         * ncclCommInitRank.  Rank 0 removes the file before writing (tmp + rename) and after the
         * communicator exists. */
        const char *path = id_path = getenv("MACROC_ID_FILE");
        if (!path) { fprintf(stderr, "macroc: WORLD_SIZE > 1 needs MACROC_ID_FILE\n"); return 62; }
        const char *nv = getenv("MACROC_LAUNCH_NONCE") ? getenv("MACROC_LAUNCH_NONCE") : getenv("MASTER_PORT");
        char nonce[64];
        memset(nonce, 0, sizeof(nonce));
        if (nv) strncpy(nonce, nv, sizeof(nonce) - 1);
        char tmp[4096];
        snprintf(tmp, sizeof(tmp), "%s.tmp", path);
        if (rank == 0) {
            unlink(path);
            if (macroc_get_unique_id(id)) { fprintf(stderr, "macroc: %s\n", macroc_last_error(NULL)); return 99; }
            FILE *f = fopen(tmp, "wb");
            if (!f || fwrite(id, 1, 128, f) != 128 || fwrite(nonce, 1, sizeof(nonce), f) != sizeof(nonce)) {
                fprintf(stderr, "macroc: cannot write %s\n", tmp);
                return 65;
            }
            fclose(f);
            if (rename(tmp, path)) { fprintf(stderr, "macroc: cannot create %s\n", path); return 65; }
        } else {
            int ok = 0;
            for (int tries = 0; tries < 600 && !ok; ++tries) {
                FILE *f = fopen(path, "rb");
                char got[64];
                if (f && fread(id, 1, 128, f) == 128 && fread(got, 1, sizeof(got), f) == sizeof(got) &&
                    memcmp(got, nonce, sizeof(nonce)) == 0)
                    ok = 1;
                if (f) fclose(f);
                if (!ok) usleep(100000);
            }
            if (!ok) { fprintf(stderr, "macroc: no NCCL id for this launch in %s (stale file or rank 0 missing)\n", path); return 65; }
        }
    }
    FILE *file_out = rank == 0 ? fopen("info.dat", "w") : NULL;
    if (rank == 0 && !file_out) { fprintf(stderr, "macroc: cannot open info.dat for writing\n"); return 65; }
    if (rank == 0) {
        printf("\nMacroC : A HPC for FE2 Multi-scale Simulations\n\n");
        printf("Boundary Condition : %s\n", cfg.bc_type == MACROC_BC_BENDING ? "BC_BENDING" : "BC_CIRCLE");
        printf("Number of GPUs     : %d\n", nranks);
        printf("Number of Elements : %ld\n", (long)(cfg.NX - 1) * (cfg.NY - 1) * (cfg.NZ - 1));
        printf("Number of Nodes    : %ld\n", (long)cfg.NX * cfg.NY * cfg.NZ);
        printf("Number of DOFs     : %ld\n\n", (long)cfg.NX * cfg.NY * cfg.NZ * 3);
        printf("NX   : %d\tNY   : %d\tNZ   : %d\n\n", cfg.NX, cfg.NY, cfg.NZ);
        printf("KSP Info: type = cg\trtol = %e\tabstol = %e\tdtol = %e\tmaxits = %d\n\n", cfg.ksp_rtol, cfg.ksp_abstol,
               cfg.ksp_dtol, cfg.ksp_maxits);
    }
    int rc = macroc_create(&cfg, rank, nranks, nranks > 1 ? id : NULL, &ctx);
    if (rank == 0 && id_path) unlink(id_path);          /* every rank has read it once the communicator exists */
    if (rc) { fprintf(stderr, "macroc: macroc_create failed (%d): %s\n", rc, macroc_last_error(NULL)); return rc; }
    if (rank == 0)
        printf("------------------------------------------------------------\n"
               "STARTING CALCULATION...\n"
               "------------------------------------------------------------\n");
    double t1 = wtime(), norm = 0., norm_0 = 0.;
    for (int time_s = 0; time_s < cfg.ts; ++time_s) {               /* main.c:49 */
        if (rank == 0) printf("\n\nTime Step = %d\n", time_s);
        double U = macroc_get_displacement(ctx, time_s);
        CHK(macroc_apply_bc_on_u(ctx, U));
        int newton_it = 0;
        while (newton_it < cfg.newton_max_its) {                     /* main.c:57 */
            if (rank == 0) printf("\nNewton Iteration = %d\nHomogenizing MicroPP\n", newton_it);
            CHK(macroc_set_strains(ctx, 0));
            if (rank == 0) printf("Assemblying RHS\n");
            CHK(macroc_assembly_res(ctx, &norm));
            if (rank == 0) printf("|RES| = %e\n", norm);
            if (newton_it == 0) norm_0 = norm;
            if (norm < cfg.newton_min_tol || norm < norm_0 * cfg.newton_rel_tol) break;
            CHK(macroc_assembly_jac(ctx));
            int its = 0;
            double rnorm = 0.;
            CHK(macroc_solve_Ax(ctx, &its, &rnorm));
            if (rank == 0) printf("KSP : |Ax - b|/|Ax| = %e\tIts = %d\n", rnorm, its);   /* assembly.c:188 */
            CHK(macroc_update_u(ctx));
            newton_it++;
        }
        double force = 0.;
        CHK(macroc_calc_force(ctx, &force));
        if (rank == 0) {
            printf("Non-Linear Gauss points : %ld\n", 0L);
            printf("F_trial_max             : %e\n", 0.);
            fprintf(file_out, "%d\t%e\t%e\t%e\t%e\t%d\n", time_s, time_s * cfg.dt, U, force, 0., 0);   /* main.c:96 */
        }
        if (cfg.vtu_freq > 0 && time_s % cfg.vtu_freq == 0) {     /* main.c:100-108 */
            char file_prefix[4096];
            snprintf(file_prefix, sizeof(file_prefix), "solution_%d", time_s);
            CHK(macroc_write_pvtu(ctx, file_prefix));
        }
    }
    CHK(macroc_device_synchronize(ctx));
    double t2 = wtime();
    if (rank == 0) {
        printf("\n\n------------------------------------------------------------\n"
               "FINISHING CALCULATION...\n"
               "------------------------------------------------------------\n");
        printf("Elapsed time : %f\n", t2 - t1);
        fclose(file_out);
    }
    macroc_destroy(ctx);
    return 0;
}
