/*
 * abi_smoke.c -- a C-only consumer of include/dmcmc.h (no Python, no ctypes mirror in between).
 *
 * Drives the drop-in boundary the way the reference's Julia methods would through ccall:
 *   init_global_workspace  (src/workspaces.jl:145-191)  -> create, set_grid, set_obs, set_start, set_theta, set_rho
 *   build_guid_prop / backward_filter!                   -> set_layout, backward_filter
 *   init_paths! + init_ll!  (src/workspaces.jl:80-92, :425-431) -> init_paths
 *   update!(::PathImputation) (src/run!_alterations.jl:5-26)    -> impute (Philox noise), twice, second time blocked
 *   set_proposal! + accept   (src/run!_alterations.jl:177-211)  -> param_loglik, param_commit
 *   save!                    (src/path_saving_buffer.jl:52-64)   -> download_path, save_path_async / _wait
 * and prints every result with %.17g so that tests/test_abi_smoke.py can compare them bit for bit with the same
 * sequence issued through the Python mirror.  Exit code 0 = every call returned 0.
 *
 * Build: gcc -std=c99 -Wall -Wextra -Werror -I include tests/abi_smoke.c -L<dir> -l:libdmcmc_b200.so -Wl,-rpath,<dir> -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dmcmc.h"

#define NOBS 6
#define NSTEP 10
#define NCH 4

#define CHECK(call)                                                                        \
    do {                                                                                   \
        if ((call) != 0) {                                                                 \
            fprintf(stderr, "FAILED %s: %s\n", #call, dmcmc_last_error(ctx));              \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

static void print_vec(const char *name, const double *v, int n) {
    printf("%s", name);
    for (int i = 0; i < n; ++i) printf(" %.17g", v[i]);
    printf("\n");
}

int main(int argc, char **argv) {
    const int check_only = argc > 1 && strcmp(argv[1], "--symbols") == 0;
    dmcmc_ctx *ctx = NULL;
    dmcmc_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = DMCMC_ABI_VERSION;
    cfg.model = DMCMC_MODEL_FHN;
    cfg.d = 2;
    cfg.n_chains = NCH;
    cfg.chain_offset = 0;
    cfg.interval_offset = 0;
    cfg.device = 0;
    cfg.seed = 20261020u;
    cfg.pin_eps = 1e-4;
    cfg.store_noise = 1;
    printf("sizeof(dmcmc_config) %zu abi %d\n", sizeof cfg, DMCMC_ABI_VERSION);
    if (check_only) { /* no GPU here: the library must load, resolve, and refuse to create a context */
        int rc = dmcmc_create(&cfg, &ctx);
        printf("create rc %d: %s\n", rc, dmcmc_last_error(NULL));
        return rc == 0 ? (dmcmc_destroy(ctx), 0) : 0;
    }
    if (dmcmc_create(&cfg, &ctx) != 0) {
        fprintf(stderr, "dmcmc_create: %s\n", dmcmc_last_error(NULL));
        return 1;
    }

    /* one recording, NOBS observations every 0.1, NSTEP Euler steps per interval */
    int32_t rec_nobs[1] = {NOBS}, n_steps[NOBS];
    double t[NOBS * (NSTEP + 1)], L[NOBS * 2], Sigma[NOBS], v[NOBS];
    static const double obs[NOBS] = {-0.92, -0.71, -0.35, 0.48, 1.12, 1.31};
    for (int j = 0; j < NOBS; ++j) {
        n_steps[j] = NSTEP;
        for (int k = 0; k <= NSTEP; ++k) t[j * (NSTEP + 1) + k] = 0.1 * j + 0.1 * k / NSTEP;
        L[2 * j] = 1.0; L[2 * j + 1] = 0.0;
        Sigma[j] = 0.01;
        v[j] = obs[j];
    }
    CHECK(dmcmc_set_grid(ctx, 1, rec_nobs, n_steps, t));
    CHECK(dmcmc_set_obs(ctx, 1, L, Sigma, v));
    double x0[NCH * 2];
    for (int c = 0; c < NCH; ++c) { x0[2 * c] = -1.0; x0[2 * c + 1] = -1.0; }
    CHECK(dmcmc_set_start(ctx, x0));
    double theta[5] = {0.1, -0.8, 1.5, 0.0, 0.3};
    CHECK(dmcmc_set_theta(ctx, theta, 5));
    double rho[NOBS];
    for (int j = 0; j < NOBS; ++j) rho[j] = 0.9;
    CHECK(dmcmc_set_rho(ctx, rho));

    int32_t i1[NOBS], i2[NOBS], pin[NOBS];
    int nblk = dmcmc_chequerboard_layout(1, rec_nobs, 0, 0, i1, i2, pin);
    CHECK(dmcmc_set_layout(ctx, nblk, i1, i2, pin));
    CHECK(dmcmc_backward_filter(ctx, 0));
    CHECK(dmcmc_init_paths(ctx, NULL));
    double ll[NCH * NOBS], llp[NCH * NOBS];
    uint8_t acc[NCH * NOBS];
    CHECK(dmcmc_get_ll(ctx, ll));
    print_vec("init_ll", ll, NCH * nblk);

    CHECK(dmcmc_impute(ctx, 0, NULL, NULL, llp, ll, acc));
    print_vec("impute0_llp", llp, NCH * nblk);
    print_vec("impute0_ll", ll, NCH * nblk);
    printf("impute0_acc");
    for (int i = 0; i < NCH * nblk; ++i) printf(" %d", acc[i]);
    printf("\n");

    /* chequerboard blocking, half_len = 1, shifted pattern: layout switch + imputation */
    nblk = dmcmc_chequerboard_layout(1, rec_nobs, 1, 1, i1, i2, pin);
    CHECK(dmcmc_relayout(ctx, nblk, i1, i2, pin));
    if (dmcmc_n_blocks(ctx) != nblk) return 2;
    CHECK(dmcmc_impute(ctx, 1, NULL, NULL, llp, ll, acc));
    print_vec("impute1_llp", llp, NCH * nblk);
    print_vec("impute1_ll", ll, NCH * nblk);
    printf("impute1_acc");
    for (int i = 0; i < NCH * nblk; ++i) printf(" %d", acc[i]);
    printf("\n");

    /* parameter step: gamma -> 1.55 (critical), accept */
    double theta_prop[5] = {0.1, -0.8, 1.55, 0.0, 0.3};
    uint8_t ok[NCH];
    CHECK(dmcmc_param_loglik(ctx, theta_prop, 5, 1, llp, ok));
    print_vec("param_llp", llp, NCH * nblk);
    printf("param_ok");
    for (int c = 0; c < NCH; ++c) printf(" %d", ok[c]);
    printf("\n");
    CHECK(dmcmc_param_commit(ctx, 1));
    CHECK(dmcmc_get_ll(ctx, ll));
    print_vec("param_ll_after_commit", ll, NCH * nblk);

    /* glued accepted path of chain 2, synchronously and through the asynchronous save */
    double path[(NOBS * NSTEP + 1) * 2], path2[2 * (NOBS * NSTEP + 1) * 2];
    CHECK(dmcmc_download_path(ctx, 2, 0, path));
    print_vec("path_chain2", path, (NOBS * NSTEP + 1) * 2);
    int32_t chains[2] = {2, 0}, ticket = -1;
    CHECK(dmcmc_save_path_async(ctx, 2, chains, 0, &ticket));
    CHECK(dmcmc_save_path_wait(ctx, ticket, path2));
    if (memcmp(path, path2, sizeof path) != 0) {
        fprintf(stderr, "save_path_async differs from download_path\n");
        return 3;
    }
    print_vec("path_chain0", path2 + (NOBS * NSTEP + 1) * 2, (NOBS * NSTEP + 1) * 2);

    int64_t a[NOBS], p[NOBS];
    CHECK(dmcmc_accept_counts(ctx, a, p));
    printf("accept_counts");
    for (int b = 0; b < nblk; ++b) printf(" %lld/%lld", (long long)a[b], (long long)p[b]);
    printf("\n");
    dmcmc_destroy(ctx);
    printf("abi_smoke ok\n");
    return 0;
}
