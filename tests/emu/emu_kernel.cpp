// Host emulation of conditional_ude_b200/csrc/cude_kernels.cuh — TEST TOOL ONLY.
// Compiles the *same kernel source* with g++ behind a minimal CUDA shim and runs it one
// "thread" per block (blockDim = 1), so that the kernel's integrator / adjoint logic can be
// checked against the oracle in the CPU-only test tier (pytest -m "not gpu").  It is never
// loaded by the package; the product path is the CUDA library and has no CPU fallback.
#define CUDE_HOST_EMU 1
#include <math.h>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>
#include <algorithm>
using std::isfinite;

#define __global__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__
#define __constant__ static const
#define __host__
#define CUDART_INF INFINITY
#define CUDART_NAN NAN
struct emu_dim3 { int x, y, z; };
static emu_dim3 blockIdx, threadIdx, blockDim, gridDim;
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline void __syncthreads() {}
static inline void __syncwarp() {}
template <class T> static inline T __shfl_xor_sync(unsigned, T, int) { return T(0); }   // lanes 1..31 are empty
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline int atomicExch(int* p, int v) { int o = *p; *p = v; return o; }
namespace cude { double smem[1 << 16]; }   // the kernel's `extern __shared__ double smem[]`

static double* g_trace_buf = nullptr; static int g_trace_cap = 0, g_trace_n = 0;
#define CUDE_TRACE_STEP(t, dt, eest) if (g_trace_buf && g_trace_n < g_trace_cap) { double* r_ = g_trace_buf + 4 * g_trace_n++; r_[0] = t; r_[1] = dt; r_[2] = eest; r_[3] = (eest <= 1.0) ? 1.0 : 0.0; }
#include "../../conditional_ude_b200/csrc/cude_kernels.cuh"
#include "../../conditional_ude_b200/csrc/cude_split.cuh"
#include "../../conditional_ude_b200/csrc/cude_sup_kernel.cuh"

using namespace cude;

extern "C" int emu_eval(int n_ind, int max_knots, const int* n_knots, const double* knot_t, const double* knot_g,
                        int max_obs, const int* n_obs, const double* obs_t, const double* obs_y,
                        const double* kin, const double* cov,
                        int n_in, int n_starts, const double* neural, long long neural_stride, const double* cond,
                        double abstol, double reltol, int maxiters, int grad, int flat,
                        double* sse, double* g_neural_traj, double* g_cond, unsigned long long* counters, int mixed) {
    const size_t N = n_ind, K = max_knots, M = max_obs;
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
    EvalArgs a{};
    a.pop.n_ind = n_ind; a.pop.max_knots = max_knots; a.pop.max_obs = max_obs;
    a.pop.n_knots = n_knots; a.pop.knot_t = kt.data(); a.pop.knot_g = kg.data(); a.pop.slope = sl.data();
    a.pop.n_obs = n_obs; a.pop.obs_t = ot.data(); a.pop.obs_y = oy.data();
    a.pop.k0 = k0.data(); a.pop.k1 = k1.data(); a.pop.k2 = k2.data(); a.pop.c0 = c0.data(); a.pop.cov = cov ? cv.data() : nullptr;
    a.n_starts = n_starts; a.neural = neural; a.neural_stride = neural_stride; a.cond = cond;
    a.abstol = abstol; a.reltol = reltol; a.maxiters = maxiters;
    a.flat = flat; a.nchunks = n_ind; a.cond_scale = 1.0;
    a.sse_out = sse; a.g_cond = g_cond; a.counters = counters;
    const int P = (n_in == 2) ? NetShape<2, 2, 4>::P : NetShape<3, 2, 4>::P;
    const long long nblocks = (long long)n_ind * n_starts;
    std::vector<double> partials((size_t)nblocks * (P + 1), 0.0);
    a.partials = flat ? nullptr : partials.data();
    blockDim.x = 1; threadIdx.x = 0;
    for (long long b = 0; b < nblocks; ++b) {
        blockIdx.x = (int)b;
        if (grad == 2) {   // d/d cond by forward sensitivity (BSENS)
            if (n_in == 2) cude_eval_kernel<NetShape<2, 2, 4>, false, false, true>(a); else cude_eval_kernel<NetShape<3, 2, 4>, false, false, true>(a);
        }
        else if (mixed == 2 && grad) cude_eval_kernel<NetShape<2, 2, 4>, true, false, false, true>(a);   // FP32 adjoint network
        else if (mixed) { if (grad) cude_eval_kernel<NetShape<2, 2, 4>, true, true>(a); else cude_eval_kernel<NetShape<2, 2, 4>, false, true>(a); }
        else if (n_in == 2) { if (grad) cude_eval_kernel<NetShape<2, 2, 4>, true>(a); else cude_eval_kernel<NetShape<2, 2, 4>, false>(a); }
        else { if (grad) cude_eval_kernel<NetShape<3, 2, 4>, true>(a); else cude_eval_kernel<NetShape<3, 2, 4>, false>(a); }
    }
    if (!flat && grad == 1 && g_neural_traj)
        for (long long b = 0; b < nblocks; ++b)   // block b = s*N + i = trajectory index
            for (int p = 0; p < P; ++p) g_neural_traj[b * P + p] = partials[b * (P + 1) + 1 + p];
    return 0;
}

// The split gradient pipeline (csrc/cude_split.cuh) stage by stage, one thread per block: forward solve with step records
// -> adjoint recursion -> (host) scan -> node kernel -> finish; returns per-start sums {sum sse, d sum sse / d neural} and
// per-trajectory sse / d sse / d cond.  n_in = 2 only.
extern "C" int emu_eval_split(int n_ind, int max_knots, const int* n_knots, const double* knot_t, const double* knot_g,
                              int max_obs, const int* n_obs, const double* obs_t, const double* obs_y, const double* kin,
                              int n_starts, const double* neural, const double* cond, double abstol, double reltol, int maxiters,
                              double* sse, double* sums, double* g_cond, int* n_overflow) {
    typedef NetShape<2, 2, 4> NS;
    const size_t N = n_ind, K = max_knots, M = max_obs, S = n_starts, NT = N * S;
    const int P = NS::P, np1 = P + 1;
    std::vector<double> kt(K * N), kg(K * N), sl(K * N, 0.0), ot(M * N), oy(M * N), k0(N), k1(N), k2(N), c0(N);
    for (size_t i = 0; i < N; ++i) {
        const int nk = n_knots[i], no = n_obs[i];
        for (int k = 0; k < max_knots; ++k) {
            const int kk = k < nk ? k : nk - 1;
            kt[k * N + i] = knot_t[i * K + kk]; kg[k * N + i] = knot_g[i * K + kk];
            if (k + 1 < nk) sl[k * N + i] = (knot_g[i * K + k + 1] - knot_g[i * K + k]) / (knot_t[i * K + k + 1] - knot_t[i * K + k]);
        }
        for (int k = 0; k < max_obs; ++k) { const int kk = k < no ? k : no - 1; ot[k * N + i] = obs_t[i * M + kk]; oy[k * N + i] = obs_y[i * M + kk]; }
        k0[i] = kin[4 * i]; k1[i] = kin[4 * i + 1]; k2[i] = kin[4 * i + 2]; c0[i] = kin[4 * i + 3];
    }
    EvalArgs a{};
    a.pop.n_ind = n_ind; a.pop.max_knots = max_knots; a.pop.max_obs = max_obs;
    a.pop.n_knots = n_knots; a.pop.knot_t = kt.data(); a.pop.knot_g = kg.data(); a.pop.slope = sl.data();
    a.pop.n_obs = n_obs; a.pop.obs_t = ot.data(); a.pop.obs_y = oy.data();
    a.pop.k0 = k0.data(); a.pop.k1 = k1.data(); a.pop.k2 = k2.data(); a.pop.c0 = c0.data(); a.pop.cov = nullptr;
    a.n_starts = n_starts; a.neural = neural; a.neural_stride = P; a.cond = cond;
    a.abstol = abstol; a.reltol = reltol; a.maxiters = maxiters; a.flat = 0; a.nchunks = n_ind; a.cond_scale = 1.0;
    unsigned long long counters[3] = {0, 0, 0};
    a.sse_out = sse; a.counters = counters;
    std::vector<double> rec(NT * SPLIT_CAP * SPLIT_W, 0.0), wrec(NT * SPLIT_CAP * SPLIT_WW, 0.0), res(M * NT, 0.0), beta(NT), spsse(NT), wsum(NT);
    std::vector<int> nrec(NT), flag(NT, 0), blklist(NT + 1, 0);
    a.sp_blklist = blklist.data(); a.sp_blkcount = blklist.data() + NT;
    a.sp_rec = rec.data(); a.sp_res = res.data(); a.sp_nrec = nrec.data(); a.sp_beta = beta.data(); a.sp_sse = spsse.data(); a.sp_blkflag = flag.data();
    blockDim.x = 1; threadIdx.x = 0; gridDim.x = (int)NT; gridDim.y = 1;
    // stage 1 (chunk-major block order: block b = chunk c * S + s, one individual per chunk)
    for (size_t b = 0; b < NT; ++b) { blockIdx.x = (int)b; blockIdx.y = 0; cude_eval_kernel<NS, false, false, false, false, false, true>(a); }
    // stage 2
    RecurArgs ra{};
    ra.pop = a.pop; ra.ntraj = (long long)NT; ra.sp_rec = rec.data(); ra.sp_w = wrec.data(); ra.sp_res = res.data(); ra.sp_nrec = nrec.data(); ra.sp_wsum = wsum.data();
    for (size_t b = 0; b < NT; ++b) { blockIdx.x = (int)b; cude_recur_kernel(ra); }
    // stage 3 (host)
    std::vector<unsigned int> off(NT + 1), map;
    unsigned int o = 0; int novf = 0;
    for (size_t j = 0; j < NT; ++j) { off[j] = o; const int c = nrec[j] > 0 ? nrec[j] : 0; if (nrec[j] < 0) ++novf; for (int q = 0; q < c; ++q) map.push_back((unsigned int)j); o += c; }
    off[NT] = o;
    if (map.empty()) map.push_back(0);
    if (n_overflow) *n_overflow = novf;
    // stage 4: one "block" per start
    std::vector<double> gc(map.size(), 0.0), pA(S * np1, 0.0), pB(NT * np1, 0.0);
    NodeArgs na{};
    na.pop = a.pop; na.neural = neural; na.neural_stride = P; na.wc_base = 0; na.sp_rec = rec.data(); na.sp_w = wrec.data();
    na.off = off.data(); na.map = map.data(); na.sp_beta = beta.data(); na.gc_rec = gc.data(); na.partials = pA.data();
    gridDim.x = 1; gridDim.y = (int)S;
    for (size_t s = 0; s < S; ++s) { blockIdx.x = 0; blockIdx.y = (int)s; cude_node_kernel<NS, double, false>(na); }
    // stage 5
    FinalArgs fa{};
    fa.pop = a.pop; fa.n_starts = n_starts; fa.nchunks = n_ind; fa.neural = neural; fa.neural_stride = P; fa.sp_nrec = nrec.data();
    fa.sp_beta = beta.data(); fa.sp_wsum = wsum.data(); fa.sp_sse = spsse.data(); fa.off = off.data(); fa.gc_rec = gc.data();
    fa.cond_scale = 1.0; fa.g_cond = g_cond; fa.partials = pB.data();
    gridDim.x = (int)NT; gridDim.y = 1;
    for (size_t b = 0; b < NT; ++b) { blockIdx.x = (int)b; blockIdx.y = 0; cude_final_kernel<NS, double>(fa); }
    // second-stage reduction: stage-4 row of the start + the stage-5 rows of its trajectories (row group s * nchunks + c)
    for (size_t s = 0; s < S; ++s)
        for (int q = 0; q < np1; ++q) {
            double v = pA[s * np1 + q];
            for (size_t c = 0; c < N; ++c) v += pB[(s * N + c) * np1 + q];
            sums[s * np1 + q] = v;
        }
    return 0;
}

// The two-kernel gradient (opts.balance = 2): stage 1 as above -> every start's individuals stably sorted by their number of
// accepted steps (host) -> cude_adjoint_kernel in the sorted order, one thread per block.  Returns the same quantities as
// emu_eval_split; trajectories beyond SPLIT_CAP steps (the fused kernel's on the device) are only counted.
extern "C" int emu_eval_exact(int n_ind, int max_knots, const int* n_knots, const double* knot_t, const double* knot_g,
                              int max_obs, const int* n_obs, const double* obs_t, const double* obs_y, const double* kin,
                              int n_starts, const double* neural, const double* cond, double abstol, double reltol, int maxiters,
                              double* sse, double* sums, double* g_cond, int* n_overflow) {
    typedef NetShape<2, 2, 4> NS;
    const size_t N = n_ind, K = max_knots, M = max_obs, S = n_starts, NT = N * S;
    const int P = NS::P, np1 = P + 1;
    std::vector<double> kt(K * N), kg(K * N), sl(K * N, 0.0), ot(M * N), oy(M * N), k0(N), k1(N), k2(N), c0(N);
    for (size_t i = 0; i < N; ++i) {
        const int nk = n_knots[i], no = n_obs[i];
        for (int k = 0; k < max_knots; ++k) {
            const int kk = k < nk ? k : nk - 1;
            kt[k * N + i] = knot_t[i * K + kk]; kg[k * N + i] = knot_g[i * K + kk];
            if (k + 1 < nk) sl[k * N + i] = (knot_g[i * K + k + 1] - knot_g[i * K + k]) / (knot_t[i * K + k + 1] - knot_t[i * K + k]);
        }
        for (int k = 0; k < max_obs; ++k) { const int kk = k < no ? k : no - 1; ot[k * N + i] = obs_t[i * M + kk]; oy[k * N + i] = obs_y[i * M + kk]; }
        k0[i] = kin[4 * i]; k1[i] = kin[4 * i + 1]; k2[i] = kin[4 * i + 2]; c0[i] = kin[4 * i + 3];
    }
    EvalArgs a{};
    a.pop.n_ind = n_ind; a.pop.max_knots = max_knots; a.pop.max_obs = max_obs;
    a.pop.n_knots = n_knots; a.pop.knot_t = kt.data(); a.pop.knot_g = kg.data(); a.pop.slope = sl.data();
    a.pop.n_obs = n_obs; a.pop.obs_t = ot.data(); a.pop.obs_y = oy.data();
    a.pop.k0 = k0.data(); a.pop.k1 = k1.data(); a.pop.k2 = k2.data(); a.pop.c0 = c0.data(); a.pop.cov = nullptr;
    a.n_starts = n_starts; a.neural = neural; a.neural_stride = P; a.cond = cond;
    a.abstol = abstol; a.reltol = reltol; a.maxiters = maxiters; a.flat = 0; a.nchunks = n_ind; a.cond_scale = 1.0;
    unsigned long long counters[3] = {0, 0, 0};
    a.sse_out = sse; a.counters = counters;
    std::vector<double> rec(NT * SPLIT_CAP * SPLIT_W, 0.0), res(M * NT, 0.0), beta(NT), spsse(NT);
    std::vector<int> nrec(NT), flag(NT, 0), blklist(NT + 1, 0);
    std::vector<unsigned int> keys(NT), order(NT);
    std::vector<unsigned short> k16(NT);
    a.sp_blklist = blklist.data(); a.sp_blkcount = blklist.data() + NT;
    a.sp_rec = rec.data(); a.sp_res = res.data(); a.sp_nrec = nrec.data(); a.sp_beta = beta.data(); a.sp_sse = spsse.data(); a.sp_blkflag = flag.data();
    a.keys_out = keys.data(); a.keys16_out = k16.data();
    blockDim.x = 1; threadIdx.x = 0; gridDim.x = (int)NT; gridDim.y = 1;
    for (size_t b = 0; b < NT; ++b) { blockIdx.x = (int)b; blockIdx.y = 0; cude_eval_kernel<NS, false, false, false, false, false, true>(a); }
    int novf = 0;
    for (size_t s = 0; s < S; ++s) {                    // the library's one stable sort on (start, steps), restricted to a start
        std::vector<unsigned int> idx(N);
        for (size_t i = 0; i < N; ++i) idx[i] = keys[s * N + i];
        std::stable_sort(idx.begin(), idx.end(), [](unsigned int x, unsigned int y) { return (x >> 24) < (y >> 24); });
        for (size_t i = 0; i < N; ++i) { order[s * N + i] = idx[i]; if (k16[s * N + i] != (unsigned short)((s << 8) | (keys[s * N + i] >> 24))) return 2; }
    }
    for (size_t j = 0; j < NT; ++j) if (nrec[j] < 0) ++novf;
    if (n_overflow) *n_overflow = novf;
    std::vector<double> pB(NT * np1, 0.0);
    AdjArgs aa{};
    aa.pop = a.pop; aa.n_starts = n_starts; aa.nchunks = n_ind; aa.neural = neural; aa.neural_stride = P; aa.wc_base = 0;
    aa.sp_rec = rec.data(); aa.sp_res = res.data(); aa.sp_nrec = nrec.data(); aa.sp_beta = beta.data(); aa.sp_sse = spsse.data();
    aa.order = order.data(); aa.cond_scale = 1.0; aa.g_cond = g_cond; aa.partials = pB.data();
    for (size_t b = 0; b < NT; ++b) { blockIdx.x = (int)b; cude_adjoint_kernel<NS, double, false>(aa); }
    for (size_t s = 0; s < S; ++s)
        for (int q = 0; q < np1; ++q) {
            double v = 0.0;
            for (size_t c = 0; c < N; ++c) v += pB[(s * N + c) * np1 + q];
            sums[s * np1 + q] = v;
        }
    return 0;
}

extern "C" void emu_set_trace(double* buf, int cap) { g_trace_buf = buf; g_trace_cap = cap; g_trace_n = 0; }
extern "C" int emu_trace_count(void) { return g_trace_n; }
// suppression variant: data in Julia layout [3 x n_obs x n_ind]; one thread per block
extern "C" int emu_sup_eval(int n_ind, int n_obs, const double* obs_t, const double* data, const double* p_true, const double* scale,
                            double t0, double tend, int n_starts, const double* neural, long long neural_stride, const double* theta,
                            double abstol, double reltol, int maxiters, int grad,
                            double* sse, double* g_neural_traj, double* g_theta, unsigned long long* counters) {
    typedef SupNet<5, 3> SN;
    const size_t N = n_ind, M = n_obs;
    std::vector<double> h(M * 3 * N);
    for (size_t i = 0; i < N; ++i)
        for (size_t k = 0; k < M; ++k)
            for (int j = 0; j < 3; ++j) h[(k * 3 + j) * N + i] = data[j + 3 * (k + M * i)];
    SupArgs a{};
    a.n_ind = n_ind; a.n_obs = n_obs; a.n_starts = n_starts; a.nchunks = n_ind;
    a.spb = (n_ind == 1) ? 1 : 0;      // one individual: exercise the flat (small-population) code path, one trajectory per "block"
    const int prow_stride = a.spb ? 2 : 1;   // flat mode keeps two partial rows per start
    a.obs_t = obs_t; a.data = h.data(); a.p1 = p_true[0]; a.p3 = p_true[2];
    for (int j = 0; j < 3; ++j) a.iscale[j] = 1.0 / scale[j];
    a.t0 = t0; a.tend = tend; a.neural = neural; a.neural_stride = neural_stride; a.theta = theta;
    a.abstol = abstol; a.reltol = reltol; a.maxiters = maxiters; a.theta_scale = 1.0;
    a.sse_out = sse; a.g_theta = g_theta; a.counters = counters;
    const long long nblocks = (long long)n_ind * n_starts;
    std::vector<double> partials((size_t)nblocks * prow_stride * (SN::P + 1), 0.0);
    a.partials = partials.data();
    blockDim.x = 1; threadIdx.x = 0; gridDim.x = 1;
    std::vector<double> ring((size_t)SUP_REC_CAP * SUP_REC_ROWS, 0.0);     // one "block" of one thread per launch
    a.ring = ring.data();
    std::vector<double> resg((size_t)3 * M, 0.0), spsse((size_t)nblocks, 0.0);
    std::vector<int> spnacc((size_t)nblocks + 1, 0);
    a.res_g = resg.data(); a.sp_sse = spsse.data(); a.sp_nacc = spnacc.data(); a.ovf_count = spnacc.data() + nblocks;
    for (long long b = 0; b < nblocks; ++b) {
        blockIdx.x = 0; a.blk0 = (int)b;
        if (grad == 2) {            // the two-kernel form: forward solve leaving records, then the adjoint sweep over them
            SupArgs f = a; f.partials = nullptr; f.g_theta = nullptr;
            cude_sup_kernel<SN, SUP_FWD_REC>(f);
            SupArgs g = a; g.counters = nullptr; g.sse_out = nullptr;
            cude_sup_kernel<SN, SUP_ADJ>(g);
        }
        else if (grad) cude_sup_kernel<SN, SUP_FUSED>(a);
        else cude_sup_kernel<SN, SUP_LOSS>(a);
    }
    if (counters && grad == 2) counters[2] += (unsigned long long)spnacc[nblocks] << 32;     // overflow count in the high word of n_fail
    if (grad && g_neural_traj)
        for (long long b = 0; b < nblocks; ++b)
            for (int p = 0; p < SN::P; ++p) g_neural_traj[b * SN::P + p] = partials[b * prow_stride * (SN::P + 1) + 1 + p];
    return 0;
}

// elementary functions of cude_math.cuh (ids as in cude_math_probe_eval)
extern "C" void emu_math(int which, int n, const double* x, double* y) {
    for (int i = 0; i < n; ++i) y[i] = cude_math_probe_eval(which, x[i], EXP_TAB256);
}
extern "C" int emu_rec_cap(void) { return REC_CAP; }
