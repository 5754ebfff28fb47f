// Host emulation of conditional_ude_b200/csrc/cude_warp.cuh — TEST TOOL ONLY.
// The warp-per-trajectory kernel needs its 32 lanes: here every CUDA thread of a 128-thread block is a host thread,
// warp shuffles and barriers are exchanges through a per-warp slot array between two barrier waits.  Same kernel source,
// compiled with g++; never loaded by the package (the product path is the CUDA library and has no CPU fallback).
#define CUDE_HOST_EMU 1
#define CUDE_HOST_EMU_WARP 1
#include "emu_threads.h"
namespace cude { double smem[1 << 16]; }
#define CUDE_TRACE_STEP(t, dt, eest)
#include "../../conditional_ude_b200/csrc/cude_warp.cuh"

using namespace cude;

// One warp per trajectory, blocks of WARP_TPB warps run one after the other.  rows[N x S][P+1] = {sse, d sse / d neural},
// g_cond[N x S], ovf[N x S] (-1: more than WARP_CAP accepted steps — the fused kernel's on the device), counters[3].
extern "C" int emu_warp_eval(int n_ind, int max_knots, const int* n_knots, const double* knot_t, const double* knot_g,
                             int max_obs, const int* n_obs, const double* obs_t, const double* obs_y, const double* kin, const double* cov,
                             int n_in, int n_starts, const double* neural, const double* cond, double abstol, double reltol, int maxiters, int grad, int lanes,
                             double* sse, double* rows, double* g_cond, int* ovf, unsigned long long* counters) {
    const size_t N = n_ind, K = max_knots, M = max_obs, S = n_starts, NT = N * S;
    std::vector<double> kt(K * N), kg(K * N), sl(K * N, 0.0), ot(M * N), oy(M * N), k0(N), k1(N), k2(N), c0(N), cv(N, 0.0);
    for (size_t i = 0; i < N; ++i) {
        const int nk = n_knots[i], no = n_obs[i];
        for (int k = 0; k < max_knots; ++k) {
            const int kk = k < nk ? k : nk - 1;
            kt[k * N + i] = knot_t[i * K + kk]; kg[k * N + i] = knot_g[i * K + kk];
            if (k + 1 < nk) sl[k * N + i] = (knot_g[i * K + k + 1] - knot_g[i * K + k]) / (knot_t[i * K + k + 1] - knot_t[i * K + k]);
        }
        for (int k = 0; k < max_obs; ++k) { const int kk = k < no ? k : no - 1; ot[k * N + i] = obs_t[i * M + kk]; oy[k * N + i] = obs_y[i * M + kk]; }
        k0[i] = kin[4 * i]; k1[i] = kin[4 * i + 1]; k2[i] = kin[4 * i + 2]; c0[i] = kin[4 * i + 3];
        if (cov) cv[i] = cov[i];
    }
    const int P = (n_in == 2) ? NetShape<2, 2, 4>::P : NetShape<3, 2, 4>::P;
    WarpArgs a{};
    a.pop.n_ind = n_ind; a.pop.max_knots = max_knots; a.pop.max_obs = max_obs;
    a.pop.n_knots = n_knots; a.pop.knot_t = kt.data(); a.pop.knot_g = kg.data(); a.pop.slope = sl.data();
    a.pop.n_obs = n_obs; a.pop.obs_t = ot.data(); a.pop.obs_y = oy.data();
    a.pop.k0 = k0.data(); a.pop.k1 = k1.data(); a.pop.k2 = k2.data(); a.pop.c0 = c0.data(); a.pop.cov = cov ? cv.data() : nullptr;
    a.n_starts = n_starts; a.neural = neural; a.neural_stride = P; a.cond = cond;
    a.abstol = abstol; a.reltol = reltol; a.maxiters = maxiters; a.cond_scale = 1.0;
    a.sse_out = sse; a.g_cond = g_cond; a.rows = rows; a.counters = counters; a.ovf = ovf;
    std::vector<int> flags(NT + 1, 0), list(NT, 0);
    a.blkflag = flags.data(); a.blkcount = flags.data() + NT; a.blklist = list.data();
    a.fb_block = 1; a.nchunks = n_ind;
    if (lanes != 32 && lanes != 8) return 3;
    if (warp_smem_doubles(P, max_knots, max_obs, lanes) > (size_t)(1 << 16)) return 1;
    const int T = 32 * WARP_TPB, tpb = WARP_TPB * (32 / lanes);
    blockDim.x = T; gridDim.x = (int)((NT + tpb - 1) / tpb);
    g_block_threads = T;
    for (int b = 0; b < gridDim.x; ++b) {
        blockIdx.x = b;
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                threadIdx.x = t;
                if (lanes == 8) {
                    if (!grad) { if (n_in == 2) cude_warp_kernel<NetShape<2, 2, 4>, false, 8>(a); else cude_warp_kernel<NetShape<3, 2, 4>, false, 8>(a); }
                    else if (n_in == 2) cude_warp_kernel<NetShape<2, 2, 4>, true, 8>(a); else cude_warp_kernel<NetShape<3, 2, 4>, true, 8>(a);
                } else {
                    if (!grad) { if (n_in == 2) cude_warp_kernel<NetShape<2, 2, 4>, false, 32>(a); else cude_warp_kernel<NetShape<3, 2, 4>, false, 32>(a); }
                    else if (n_in == 2) cude_warp_kernel<NetShape<2, 2, 4>, true, 32>(a); else cude_warp_kernel<NetShape<3, 2, 4>, true, 32>(a);
                }
            });
        for (auto& x : th) x.join();
    }
    return 0;
}
