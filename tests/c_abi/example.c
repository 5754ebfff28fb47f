/* Plain-C caller of the C ABI (include/cude_b200.h): what a cgo / ccall / JNI binding does, without any Python.
 * Two synthetic individuals, one start: loss + gradient.  Exit code 0 = ran on a GPU, 3 = no device (the library
 * reports CUDE_ENODEVICE: there is no CPU fallback), anything else = failure. */
#include <math.h>
#include <stdio.h>
#include "cude_b200.h"

int main(void) {
    cude_ctx* ctx = NULL;
    int rc = cude_ctx_create(0, &ctx);
    if (rc == CUDE_ENODEVICE) {
        printf("no device: %s\n", cude_last_error(NULL));
        return 3;
    }
    if (rc) { printf("ctx_create failed: %d %s\n", rc, cude_last_error(NULL)); return 1; }
    if (cude_abi_version() != CUDE_B200_ABI_VERSION) return 1;

    enum { N = 2, K = 5, M = 5, P = 37 };
    const int n_knots[N] = {K, K}, n_obs[N] = {M, M};
    const double t[K] = {0, 30, 60, 90, 120};
    double knot_t[N * K], knot_g[N * K], obs_t[N * M], obs_y[N * M], kin[N * 4];
    const double g[N][K] = {{5.0, 8.5, 9.0, 7.0, 5.5}, {6.0, 11.0, 13.5, 12.0, 9.0}};
    const double y[N][M] = {{0.5, 1.4, 1.9, 1.7, 1.2}, {0.7, 1.2, 1.8, 2.0, 1.9}};
    const double age[N] = {35.0, 62.0};
    const int t2dm[N] = {0, 1};
    for (int i = 0; i < N; ++i) {
        for (int k = 0; k < K; ++k) { knot_t[i * K + k] = t[k]; knot_g[i * K + k] = g[i][k]; obs_t[i * M + k] = t[k]; obs_y[i * M + k] = y[i][k]; }
        double k0, k1, k2;
        cude_van_cauter_parameters(age[i], t2dm[i], &k0, &k1, &k2);
        kin[4 * i] = k0; kin[4 * i + 1] = k1; kin[4 * i + 2] = k2; kin[4 * i + 3] = y[i][0];
    }
    cude_population* pop = NULL;
    rc = cude_population_create(ctx, N, K, n_knots, knot_t, knot_g, M, n_obs, obs_t, obs_y, kin, NULL, &pop);
    if (rc) { printf("population_create: %s\n", cude_last_error(ctx)); return 1; }

    cude_net net = {2, 2, 4};
    if (cude_net_nparams(&net) != P) return 1;
    cude_opts opts;
    cude_default_opts(&opts);
    double neural[P], cond[N] = {-1.0, -0.5}, loss = 0, g_neural[P], g_cond[N], sse[N];
    for (int p = 0; p < P; ++p) neural[p] = 0.3 * sin(1.0 + 0.7 * p);
    rc = cude_loss_grad(ctx, pop, &net, &opts, 1, neural, P, cond, 1, sse, &loss, g_neural, g_cond);
    if (rc) { printf("loss_grad: %s\n", cude_last_error(ctx)); return 1; }
    cude_stats st;
    cude_get_stats(ctx, &st);
    printf("loss %.12g  sse %.12g %.12g  dloss/dcond %.6g %.6g  |g_neural|_1 ", loss, sse[0], sse[1], g_cond[0], g_cond[1]);
    double s = 0;
    for (int p = 0; p < P; ++p) s += fabs(g_neural[p]);
    printf("%.6g  steps %llu\n", s, st.n_acc);
    if (!(loss > 0) || !isfinite(loss) || fabs(loss - 0.5 * (sse[0] + sse[1])) > 1e-12 * loss || st.n_traj != 2) return 1;
    /* loss-only call: same forward pass */
    double loss2 = 0;
    rc = cude_loss(ctx, pop, &net, &opts, 1, neural, P, cond, NULL, &loss2);
    if (rc || fabs(loss2 - loss) > 1e-14 * loss) return 1;
    cude_population_destroy(pop);
    cude_ctx_destroy(ctx);
    return 0;
}
