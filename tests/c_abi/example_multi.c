/* Plain-C caller of the multi-GPU half of the C ABI (include/cude_b200.h): one process drives n GPUs.
 *   example_multi [n_gpus] [n_individuals] [n_starts]
 * Evaluates the population loss + gradient of a synthetic population
 *   (a) on one context (device 0, the whole population),
 *   (b) on a cude_mctx with the INDIVIDUALS partition (population sharded, per-start sums all-reduced over NCCL inside
 *       the library) and
 *   (c) on the same cude_mctx with the STARTS partition (population replicated, starts split, no communication),
 * and checks (b) and (c) against (a): per-trajectory sse and d loss/d cond bit for bit; loss and d loss/d neural bit for
 * bit in (c) (same summation order) and to 1e-13 in (b) (the shards' partial sums are added in a different order).
 * Exit code 0 = agreed, 3 = no device (CUDE_ENODEVICE: no CPU fallback), 4 = fewer devices than requested, else failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cude_b200.h"

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static double urand(void) {   /* splitmix64 -> [0, 1) */
    unsigned long long z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

static double max_rel(const double* a, const double* b, size_t n, int* bitwise) {
    double scale = 0, err = 0;
    *bitwise = memcmp(a, b, n * sizeof(double)) == 0;
    for (size_t k = 0; k < n; ++k) { if (fabs(b[k]) > scale) scale = fabs(b[k]); if (fabs(a[k] - b[k]) > err) err = fabs(a[k] - b[k]); }
    return scale > 0 ? err / scale : err;
}

int main(int argc, char** argv) {
    const int n_gpus = argc > 1 ? atoi(argv[1]) : 2;
    const int N = argc > 2 ? atoi(argv[2]) : 5000;
    const int S = argc > 3 ? atoi(argv[3]) : 6;
    enum { K = 5, M = 5, P = 37 };
    if (cude_device_count() == 0) {
        cude_mctx* none = NULL;
        const int rc0 = cude_mctx_create(n_gpus, NULL, &none);
        printf("no device: rc %d %s\n", rc0, cude_mlast_error(NULL));
        return rc0 == CUDE_ENODEVICE ? 3 : 1;
    }
    if (cude_device_count() < n_gpus) { printf("only %d device(s), %d requested\n", cude_device_count(), n_gpus); return 4; }

    /* ---- synthetic Ohashi-like population (5 knots on [0, 120] min) ---- */
    int* n_knots = malloc(sizeof(int) * N); int* n_obs = malloc(sizeof(int) * N);
    double* knot_t = malloc(sizeof(double) * N * K); double* knot_g = malloc(sizeof(double) * N * K);
    double* obs_t = malloc(sizeof(double) * N * M); double* obs_y = malloc(sizeof(double) * N * M);
    double* kin = malloc(sizeof(double) * N * 4);
    const double t[K] = {0, 30, 60, 90, 120}, shape[K] = {0.0, 0.8, 1.0, 0.6, 0.3};
    for (int i = 0; i < N; ++i) {
        n_knots[i] = K; n_obs[i] = M;
        const double g0 = 4.0 + 4.0 * urand(), peak = 1.0 + 14.0 * urand(), c0 = 0.3 + 1.0 * urand(), gain = 1.0 + 3.0 * urand();
        for (int k = 0; k < K; ++k) {
            knot_t[i * K + k] = t[k]; knot_g[i * K + k] = g0 + peak * shape[k] * (0.8 + 0.4 * urand());
            obs_t[i * M + k] = t[k]; obs_y[i * M + k] = k == 0 ? c0 : c0 * (1.0 + gain * shape[k] * peak / 8.0) + 0.1 * (urand() - 0.5);
        }
        double k0, k1, k2;
        cude_van_cauter_parameters(20.0 + 60.0 * urand(), urand() < 0.44, &k0, &k1, &k2);
        kin[4 * i] = k0; kin[4 * i + 1] = k1; kin[4 * i + 2] = k2; kin[4 * i + 3] = c0;
    }
    double* neural = malloc(sizeof(double) * P * S); double* cond = malloc(sizeof(double) * (size_t)N * S);
    for (int s = 0; s < S; ++s) for (int p = 0; p < P; ++p) neural[s * P + p] = 0.3 * sin(1.0 + 0.7 * p + 0.1 * s);
    for (size_t k = 0; k < (size_t)N * S; ++k) cond[k] = -2.0 + 2.0 * urand();
    cude_net net = {2, 2, 4};
    cude_opts opts; cude_default_opts(&opts);
    const size_t NS = (size_t)N * S;
    double *sse[3], *gc[3], *loss[3], *gn[3];
    for (int v = 0; v < 3; ++v) { sse[v] = calloc(NS, 8); gc[v] = calloc(NS, 8); loss[v] = calloc(S, 8); gn[v] = calloc((size_t)P * S, 8); }

    /* ---- (a) one context, whole population ---- */
    cude_ctx* ctx = NULL; cude_population* pop = NULL;
    if (cude_ctx_create(0, &ctx)) { printf("ctx_create: %s\n", cude_last_error(NULL)); return 1; }
    if (cude_population_create(ctx, N, K, n_knots, knot_t, knot_g, M, n_obs, obs_t, obs_y, kin, NULL, &pop)) { printf("%s\n", cude_last_error(ctx)); return 1; }
    if (cude_loss_grad(ctx, pop, &net, &opts, S, neural, P, cond, 1, sse[0], loss[0], gn[0], gc[0])) { printf("%s\n", cude_last_error(ctx)); return 1; }

    /* ---- (b), (c) all GPUs from this process ---- */
    cude_mctx* m = NULL;
    int rc = cude_mctx_create(n_gpus, NULL, &m);
    if (rc) { printf("mctx_create: %d %s\n", rc, cude_mlast_error(NULL)); return 1; }
    const int modes[2] = {CUDE_SHARD_INDIVIDUALS, CUDE_SHARD_STARTS};
    int fail = 0;
    for (int v = 1; v <= 2; ++v) {
        cude_mpopulation* mp = NULL;
        rc = cude_mpopulation_create(m, modes[v - 1], N, K, n_knots, knot_t, knot_g, M, n_obs, obs_t, obs_y, kin, NULL, &mp);
        if (rc) { printf("mpopulation_create: %d %s\n", rc, cude_mlast_error(m)); return 1; }
        rc = cude_mloss_grad(m, mp, &net, &opts, S, neural, P, cond, 1, sse[v], loss[v], gn[v], gc[v]);
        if (rc) { printf("mloss_grad: %d %s\n", rc, cude_mlast_error(m)); return 1; }
        cude_stats st; cude_mget_stats(m, &st);
        int b_sse, b_gc, b_loss, b_gn;
        const double e_sse = max_rel(sse[v], sse[0], NS, &b_sse), e_gc = max_rel(gc[v], gc[0], NS, &b_gc);
        const double e_loss = max_rel(loss[v], loss[0], S, &b_loss), e_gn = max_rel(gn[v], gn[0], (size_t)P * S, &b_gn);
        printf("%s x%d: sse bitwise %d (%.1e)  g_cond bitwise %d (%.1e)  loss bitwise %d (%.1e)  g_neural bitwise %d (%.1e)  traj %llu\n",
               v == 1 ? "individuals" : "starts", cude_mctx_size(m), b_sse, e_sse, b_gc, e_gc, b_loss, e_loss, b_gn, e_gn, st.n_traj);
        if (!b_sse || !b_gc || st.n_traj != NS) fail = 1;
        if (v == 1 && (e_loss > 1e-13 || e_gn > 1e-13)) fail = 1;
        if (v == 2 && (!b_loss || !b_gn)) fail = 1;
        /* loss-only call through the same partition */
        double* l2 = calloc(S, 8);
        rc = cude_mloss(m, mp, &net, &opts, S, neural, P, cond, NULL, l2);
        int b;
        if (rc || max_rel(l2, loss[v], S, &b) > 1e-14) { printf("mloss mismatch\n"); fail = 1; }
        free(l2);
        cude_mpopulation_destroy(mp);
    }
    printf("loss[0] %.12g\n", loss[0][0]);
    cude_mctx_destroy(m);
    cude_population_destroy(pop);
    cude_ctx_destroy(ctx);
    return fail;
}
