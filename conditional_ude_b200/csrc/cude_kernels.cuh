// =====================================================================================
// cude_kernels.cuh — sm_100a kernels for the c-peptide conditional-UDE loss and its gradient.
//
// One thread integrates one trajectory (individual i, start s) with an in-kernel adaptive Tsit5
// (the scheme OrdinaryDiffEq's default algorithm runs for `solve(model.problem, p=theta,
// saveat=timepoints, save_idxs=1)`, reference src/parameter-estimation.jl:59), evaluates the SSE of
// :67 and — in the GRAD instantiation — the exact derivative of that discrete solve w.r.t. the MLP
// weights and the conditional parameter by a discrete adjoint over the recorded accepted steps.
//
// Structure exploited (reference src/c-peptide-models.jl:7-14, :86-94, :108-114):
//   u' = A u + b + e1 * prod(t),   prod(t) = NN([dG(t); beta]) - NN([0; beta]),  beta = exp(cond)
// The production term does not depend on the state, so
//   * the forward pass needs only 5 network evaluations per step (c6 = c7 = 1 and FSAL share the
//     node t+dt with the next step's first stage),
//   * the adjoint needs no stored states or stages: only (t_n, dt_n) of the accepted steps; the
//     2-vector adjoint is propagated backwards through the linear stage recursion and the network
//     is re-evaluated (forward + backward) at the 5 nodes of each step with a scalar seed.
//
// Data layout in HBM (struct-of-arrays, individual index fastest => coalesced per warp):
//   knot_t/knot_g/slope [K][N], obs_t/obs_y [M][N], k0,k1,k2,c0,cov [N], cond/sse/g_cond [N x S].
// Per block (one start x 128 individuals, blocks ordered chunk-major): the start's MLP weights and the 256-entry exp
// table are staged once in shared memory, each thread's glucose knots and observations are staged in shared memory
// ([k][tid], conflict-free), state, stages, adjoint and the gradient accumulators live in registers.
// Instantiations: loss only; loss + adjoint gradient (GRAD); loss + d/d cond by forward sensitivity (BSENS, beta-only
// fits); FP32 network throughout (MIXED) or only in the adjoint sweep (FBWD) as optional precisions.
// Reduction: per-thread gradient rows in shared memory -> per warp, lane l sums row l -> one partial row per warp ->
// deterministic second-stage kernel (no atomics on the data path).
// Measured tuning notes (B200, profiles/README.md): unrolling the node loops (CUDE_FWD_UNROLL / CUDE_BWD_UNROLL > 1),
// keeping activations for the adjoint, 2 / 4 resident blocks per SM, 32- / 64- / 96- / 192- / 384-thread blocks and
// warp-persistent scheduling were all measured slower than 3 blocks of 128 threads at 168 registers.
// =====================================================================================
#pragma once
#ifndef CUDE_HOST_EMU   // tests/emu compiles this file with g++ behind a shim (CI without a GPU)
#include <cuda_runtime.h>
#include <math_constants.h>
#endif
#include <type_traits>
#include "cude_math.cuh"

#ifndef CUDE_TRACE_STEP   // host-emulation test hook; compiles to nothing in the CUDA build
#define CUDE_TRACE_STEP(t, dt, eest)
#endif

namespace cude {

// ---------------------------------------------------------------- Tsit5 tableau
// Published coefficients of Tsitouras' 5(4) pair as used by OrdinaryDiffEq.Tsit5.
namespace tab {
#ifdef CUDE_TAB_CONSTMEM   // coefficients as loads from constant memory (LDCU) instead of 64-bit immediates (2 UMOV each)
#define CUDE_TABCONST __constant__
#else
#define CUDE_TABCONST constexpr
#endif
CUDE_TABCONST double c2 = 0.161, c3 = 0.327, c4 = 0.9, c5 = 0.9800255409045097;
CUDE_TABCONST double a21 = 0.161;
CUDE_TABCONST double a31 = -0.008480655492356989, a32 = 0.335480655492357;
CUDE_TABCONST double a41 = 2.8971530571054935, a42 = -6.359448489975075, a43 = 4.3622954328695815;
CUDE_TABCONST double a51 = 5.325864828439257, a52 = -11.748883564062828, a53 = 7.4955393428898365, a54 = -0.09249506636175525;
CUDE_TABCONST double a61 = 5.86145544294642, a62 = -12.92096931784711, a63 = 8.159367898576159, a64 = -0.071584973281401, a65 = -0.028269050394068383;
CUDE_TABCONST double b1 = 0.09646076681806523, b2 = 0.01, b3 = 0.4798896504144996, b4 = 1.379008574103742, b5 = -3.290069515436081, b6 = 2.324710524099774;
CUDE_TABCONST double e1 = -0.00178001105222577714, e2 = -0.0008164344596567469, e3 = 0.007880878010261995, e4 = -0.1447110071732629,
                 e5 = 0.5823571654525552, e6 = -0.45808210592918697, e7 = 0.015151515151515152;
CUDE_TABCONST double r11 = 1.0, r12 = -2.763706197274826, r13 = 2.9132554618219126, r14 = -1.0530884977290216;
CUDE_TABCONST double r22 = 0.13169999999999998, r23 = -0.2234, r24 = 0.1017;
CUDE_TABCONST double r32 = 3.9302962368947516, r33 = -5.941033872131505, r34 = 2.490627285651253;
CUDE_TABCONST double r42 = -12.411077166933676, r43 = 30.33818863028232, r44 = -16.548102889244902;
CUDE_TABCONST double r52 = 37.50931341651104, r53 = -88.1789048947664, r54 = 47.37952196281928;
CUDE_TABCONST double r62 = -27.896526289197286, r63 = 65.09189467479366, r64 = -34.87065786149661;
CUDE_TABCONST double r72 = 1.5, r73 = -4.0, r74 = 2.5;
// PI controller defaults of OrdinaryDiffEq for Tsit5
constexpr double beta1 = 7.0 / 50.0, beta2 = 2.0 / 25.0, gamma = 9.0 / 10.0, qmin = 1.0 / 5.0, qmax = 10.0, qoldinit = 1e-4;
}  // namespace tab

// dense-output weights b_j(theta)
__device__ __forceinline__ void dense_weights(double th, double (&bw)[7]) {
    using namespace tab;
    const double th2 = th * th;
    bw[0] = th * fma(th, fma(th, fma(th, r14, r13), r12), r11);
    bw[1] = th2 * fma(th, fma(th, r24, r23), r22);
    bw[2] = th2 * fma(th, fma(th, r34, r33), r32);
    bw[3] = th2 * fma(th, fma(th, r44, r43), r42);
    bw[4] = th2 * fma(th, fma(th, r54, r53), r52);
    bw[5] = th2 * fma(th, fma(th, r64, r63), r62);
    bw[6] = th2 * fma(th, fma(th, r74, r73), r72);
}

// ---------------------------------------------------------------- device-side problem description
struct PopDev {
    int n_ind, max_knots, max_obs;
    const int* n_knots;
    const double* knot_t;  // [K][N]
    const double* knot_g;  // [K][N]
    const double* slope;   // [K-1][N]  (g[k+1]-g[k])/(t[k+1]-t[k]), DataInterpolations' cached parameter
    const int* n_obs;
    const double* obs_t;   // [M][N]
    const double* obs_y;   // [M][N]
    const double* k0;
    const double* k1;
    const double* k2;
    const double* c0;
    const double* cov;     // [N] or nullptr
};

struct EvalArgs {
    PopDev pop;
    int n_starts;
    const double* neural;       // start s: neural + s*neural_stride
    long long neural_stride;
    const double* cond;         // [N x S]
    double abstol, reltol;
    int maxiters;
    int flat;                   // 1: shared network, trajectories flattened over (i,s)
    int nchunks;                // tiles per start (tile mode)
    double cond_scale;          // g_cond = cond_scale * d sse / d cond
    double* sse_out;            // [N x S] or nullptr
    double* partials;           // [n_blocks * warps_per_block][P+1]: {sum sse, sum d sse/d neural} per warp, or nullptr
    double* g_cond;             // [N x S] or nullptr
    unsigned long long* counters;  // {n_acc, n_rej, n_fail}
    // lane balancing (tile mode, opts.balance): position pos of start s runs individual order[s*N + pos] & 0xffffff —
    // the start's individuals sorted by the step count of a previous call, so that the 32 lanes of a warp (and the 4
    // warps of a block) finish together; keys_out[s*N + i] = (min(steps, 255) << 24) | i feeds the next sort.
    const unsigned int* order;     // [N x S] or nullptr (natural order)
    unsigned int* keys_out;        // [N x S] or nullptr
    unsigned short* keys16_out;    // [N x S] or nullptr: (start << 8) | min(accepted steps, 255) — the key of the two-kernel gradient's one
                                   // sort per group of (at most 256) starts, with keys_out as its payload
    double* yhat_out;              // [M][N x S] -> yhat_out[k + M*j]: the solution at the observation times (cude_simulate;
                                   // loss-only instantiations), or nullptr
    // split gradient pipeline (SPLIT instantiation = stage 1; see "split gradient pipeline" below)
    double* sp_rec;                // [N x S][SPLIT_CAP][SPLIT_W] one record per accepted step: {t, h, dG at the 5 nodes, 0}
    double* sp_res;                // [M][N x S] residuals yhat_k - y_k
    int* sp_nrec;                  // [N x S] accepted steps recorded; 0 = failed trajectory, -1 = more than SPLIT_CAP steps
    double* sp_beta;               // [N x S] beta = exp(cond)
    double* sp_sse;                // [N x S] sse (Inf when failed)
    int* sp_blkflag;               // [blocks] set when a trajectory of the block overflowed SPLIT_CAP (fused-kernel fallback)
    int* sp_blklist;               // [blocks] the flagged blocks (their blockIdx.x, any order), appended by stage 1 ...
    int* sp_blkcount;              // ... and how many: the fallback launch is a few resident blocks walking this list
    long long wc_base;             // offset of the call's first start in the constant weight array (WC instantiations)
    const int* only_flag;          // fused GRAD kernel as that fallback: run only the trajectories with only_flag[j] < 0 of flagged blocks
};

// ---------------------------------------------------------------- network shape
template <int NIN_, int DEPTH_, int WIDTH_>
struct NetShape {
    static constexpr int NIN = NIN_, DEPTH = DEPTH_, W = WIDTH_;
    static constexpr int L1 = W * (NIN + 1);               // first layer block
    static constexpr int LH = W * (W + 1);                 // hidden W->W layer block
    static constexpr int OFF_OUT = L1 + (DEPTH - 1) * LH;  // output layer offset
    static constexpr int P = OFF_OUT + W + 1;
    // compressed accumulator count: first layer keeps only dW1[:,0] and the sum of dz1
    static constexpr int NACC = 2 * W + (DEPTH - 1) * LH + W + 1;
};

#ifndef CUDE_MAX_THREADS
#define CUDE_MAX_THREADS 128  // largest block the kernel is launched with (opts.block <= this)
#endif
#ifndef CUDE_MIN_BLOCKS
#define CUDE_MIN_BLOCKS 3   // resident 128-thread blocks per SM the register allocation is tuned for
#endif
#ifndef CUDE_MIN_BLOCKS_LOSS
#define CUDE_MIN_BLOCKS_LOSS 3   // same for the loss-only instantiation
#endif
#ifndef CUDE_MIN_BLOCKS_LOSS_WC
#define CUDE_MIN_BLOCKS_LOSS_WC 4   // loss-only with the weights in constant memory: 128 registers keep them in uniform registers (5.1e8 vs 4.6e8 evals/s)
#endif
#ifndef CUDE_FWD_UNROLL
#define CUDE_FWD_UNROLL 1   // unroll factor of the forward network-evaluation loop over a step's 5 nodes
#endif
#ifndef CUDE_BWD_UNROLL
#define CUDE_BWD_UNROLL 1   // same for the adjoint's forward+backward evaluation loop
#endif
#define CUDE_PRAGMA(x) _Pragma(#x)
#define CUDE_UNROLL(n) CUDE_PRAGMA(unroll n)
#ifndef CUDE_REC_CAP
#define CUDE_REC_CAP 48
#endif
constexpr int REC_CAP = CUDE_REC_CAP;
// Ring of accepted-step records kept per thread (local memory).  A record is (t, dt): the adjoint recomputes z_out and
// its sigmoid from the activations it recomputes anyway.  Keeping d = 1 + exp(z_out) of the step's 5 nodes as well
// (7 doubles) took the sigmoid off the adjoint's dependent chain for +0.9 %, but the records then no longer stay in L2:
// 76 GB of DRAM traffic per 64 M-trajectory launch instead of 3.6 GB (profiles/README.md).
constexpr int REC_W = 2;   // doubles per record

// ---------------------------------------------------------------- split gradient pipeline
// The adjoint has a cheap sequential part — the 2-vector recursion backwards through the accepted steps, which only
// produces the *weights* of the network's output at the 5 nodes of every step — and an expensive part that is
// embarrassingly parallel once the weights exist: network forward + backward at every (trajectory, step, node), about
// 100 node evaluations per trajectory, whose results are merely *summed*.  The fused kernel does everything in one thread
// per trajectory: 33 FP64 accumulators stay live across 70 KB of code, and a warp waits for its longest trajectory in
// every phase.  The split pipeline separates the phases (cude_split.cuh):
//   stage 1  cude_eval_kernel<.., SPLIT>   the loss-only forward solve, one thread per trajectory; per accepted step it
//                                          leaves a record {t, h, dG at the step's 5 nodes}, per observation a residual
//   stage 2  cude_recur_kernel             adjoint recursion per trajectory (no network): the 5 node weights per record
//   stage 3  cude_scan_*                   exclusive scan of the step counts -> a flat list of (trajectory, step) records
//   stage 4  cude_node_kernel              one thread per RECORD: 5 network forward+backward evaluations into per-thread
//                                          accumulators (all lanes busy, no divergence, a few KB of hot code)
//   stage 5  cude_final_kernel             per trajectory: the NN([0;beta]) node, d sse/d cond, partial rows
// Trajectories with more than SPLIT_CAP accepted steps (tight tolerances) take the fused kernel instead (only_flag).
#ifndef CUDE_SPLIT_CAP
#define CUDE_SPLIT_CAP 32
#endif
constexpr int SPLIT_CAP = CUDE_SPLIT_CAP;
constexpr int SPLIT_W = 8;    // doubles per step record of stage 1 (64 B = 2 sectors, written whole): t, h, dG[5], 0
constexpr int SPLIT_WW = 6;   // doubles per weight record of stage 2 (48 B, written whole): w[5], 0

// per-thread view of the staged glucose knots in shared memory, layout [k][tid]
struct Knots {
    const double* t;
    const double* g;
    const double* sl;
    int nk;
    int stride;
    double g0;
    // DataInterpolations LinearInterpolation: idx = clamp(searchsortedlast(t, tau), 1, n-1)
    __device__ __forceinline__ double dG(double tau) const {
        int idx = 0;
        for (int k = 1; k < nk - 1; ++k) idx = (t[k * stride] <= tau) ? k : idx;
        const double v = fma(sl[idx * stride], tau - t[idx * stride], g[idx * stride]);
        return v - g0;
    }
    // The 5 node times of a step at once (tau increasing): 5 independent select chains, then 15 independent
    // shared-memory loads (a per-node search serialised on the LDS latency, ncu v4).  `base` is an interval index the
    // caller carries along the trajectory with t[base] <= tau[0] (or 0): only the knots in (t[base], tau[4]] are
    // examined — usually none or one instead of all (ncu v9: the full scan was ~160 of the ~1400 instructions a
    // step spends outside the network).  idx_last = interval of tau[4], the next step's base.
    __device__ __forceinline__ void dG5(const double (&tau)[5], double* out, int ostride, int base, int& idx_last) const {
        int idx[5] = {base, base, base, base, base};
        const double tmax = tau[4];
        for (int k = base + 1; k < nk - 1; ++k) {
            const double tk = t[k * stride];
            if (!(tk <= tmax)) break;
#pragma unroll
            for (int q = 0; q < 5; ++q) idx[q] = (tk <= tau[q]) ? k : idx[q];
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const double v = fma(sl[idx[q] * stride], tau[q] - t[idx[q] * stride], g[idx[q] * stride]);
            out[q * ostride] = v - g0;
        }
        idx_last = idx[4];
    }
};

// ---------------------------------------------------------------- MLP
// Node-time coefficient tables: a step evaluates the network at t + c*dt for the 5 new nodes
// (c6 = c7 = 1 share one node with the next step's first stage); the init phase evaluates the two
// nodes t0 (-> NN([0;beta])) and t0 + dt0 (Hairer's probe).
__constant__ double CN_STEP[5] = {0.161, 0.327, 0.9, 0.9800255409045097, 1.0};
__constant__ double CN_INIT[5] = {0.0, 1.0, 1.0, 1.0, 1.0};

// Output-unit pre-activation z_out for input dG; c[] = first-layer pre-activation constant part.  The softplus is
// applied by the caller to the 5 nodes of a step at once (one dependent chain per node: batching them gives the
// scheduler 5-way ILP; inside the rolled loop it was a serial tail with 5x the stall samples per instruction, ncu v6).
// R = double (parity-gated FP64 path) or float (precision = 1: FP32 network, FP64 integrator).
template <class NS>
__device__ __forceinline__ double mlp_zout(const double* __restrict__ sW, const double* __restrict__ tab, const double (&c)[NS::W], double dG) {
    constexpr int W = NS::W;
    int nanmax = 0;
    double a[W], b[W];
#pragma unroll
    for (int j = 0; j < W; ++j) a[j] = t_tanh(fma(sW[j], dG, c[j]), tab, nanmax);
    int off = NS::L1;
#pragma unroll
    for (int l = 1; l < NS::DEPTH; ++l) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
            double z = sW[off + W * W + j];
#pragma unroll
            for (int i = 0; i < W; ++i) z = fma(sW[off + i * W + j], a[i], z);
            b[j] = t_tanh(z, tab, nanmax);
        }
#pragma unroll
        for (int j = 0; j < W; ++j) a[j] = b[j];
        off += NS::LH;
    }
    double z = sW[off + W];
#pragma unroll
    for (int i = 0; i < W; ++i) z = fma(sW[off + i], a[i], z);
    return t_nan_inject(z, nanmax);
}
// z_out and dzb = sum_q (d z_out / d z1_q) * W1[q, beta column]: the derivative of z_out w.r.t. the second network
// input (beta), from the activations in hand — the beta-only fits (reference src/parameter-estimation.jl:272-307) need
// d loss / d cond only, which one forward-sensitivity column delivers without an adjoint sweep.
template <class NS>
__device__ __forceinline__ double mlp_zout_dbeta(const double* __restrict__ sW, const double* __restrict__ tab, const double (&c)[NS::W],
                                                 double dG, double& dzb) {
    constexpr int W = NS::W, D = NS::DEPTH;
    int nanmax = 0;
    double a[D][W];
#pragma unroll
    for (int j = 0; j < W; ++j) a[0][j] = t_tanh(fma(sW[j], dG, c[j]), tab, nanmax);
    int off = NS::L1;
#pragma unroll
    for (int l = 1; l < D; ++l) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
            double z = sW[off + W * W + j];
#pragma unroll
            for (int i = 0; i < W; ++i) z = fma(sW[off + i * W + j], a[l - 1][i], z);
            a[l][j] = t_tanh(z, tab, nanmax);
        }
        off += NS::LH;
    }
    double z = sW[off + W];
    double da[W];
#pragma unroll
    for (int i = 0; i < W; ++i) { z = fma(sW[off + i], a[D - 1][i], z); da[i] = sW[off + i]; }
#pragma unroll
    for (int l = D - 1; l >= 1; --l) {
        off -= NS::LH;
        double dzl[W], dprev[W];
#pragma unroll
        for (int j = 0; j < W; ++j) dzl[j] = da[j] * fma(-a[l][j], a[l][j], 1.0);
#pragma unroll
        for (int i = 0; i < W; ++i) {
            double sacc = 0.0;
#pragma unroll
            for (int j = 0; j < W; ++j) sacc = fma(sW[off + i * W + j], dzl[j], sacc);
            dprev[i] = sacc;
        }
#pragma unroll
        for (int j = 0; j < W; ++j) da[j] = dprev[j];
    }
    double g = 0.0;
#pragma unroll
    for (int j = 0; j < W; ++j) g = fma(da[j] * fma(-a[0][j], a[0][j], 1.0), sW[W + j], g);
    dzb = g;
    return t_nan_inject(z, nanmax);
}
template <class NS>
__device__ __forceinline__ float mlp_zout(const float* __restrict__ sW, const double* __restrict__ tab, const float (&c)[NS::W], float dG) {
    constexpr int W = NS::W;
    float a[W], b[W];
#pragma unroll
    for (int j = 0; j < W; ++j) a[j] = m_tanh(fmaf(sW[j], dG, c[j]), tab);
    int off = NS::L1;
#pragma unroll
    for (int l = 1; l < NS::DEPTH; ++l) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
            float z = sW[off + W * W + j];
#pragma unroll
            for (int i = 0; i < W; ++i) z = fmaf(sW[off + i * W + j], a[i], z);
            b[j] = m_tanh(z, tab);
        }
#pragma unroll
        for (int j = 0; j < W; ++j) a[j] = b[j];
        off += NS::LH;
    }
    float z = sW[off + W];
#pragma unroll
    for (int i = 0; i < W; ++i) z = fmaf(sW[off + i], a[i], z);
    return z;
}
// softplus of z_out and d = 1 + exp(z_out) (kept per accepted step for the adjoint: d softplus = 1 - 1/d)
__device__ __forceinline__ void softplus_d(double z, bool /*mixed*/, const double* __restrict__ tab, double& sp, double& d) { t_softplus_d(z, tab, sp, d); }
__device__ __forceinline__ void softplus_d_mixed(double z, const double* __restrict__ tab, double& sp, double& d) {
    const float x = (float)z;
    sp = (double)m_softplus(x, tab);
    d = 1.0 + (double)f_ex2(f_clamp(x, -30.0f, 30.0f) * 1.4426950408889634f);
}

// Adjoint at one time node: acc += dz * d z_out/d(params), dz = (node weight) * sigmoid(z_out).  The hidden activations
// are recomputed (8 tanh): keeping them from the forward pass (9 doubles per node, ~7 KB per trajectory in local memory)
// was measured slower on B200 (1.05e8 vs 1.30e8 evals/s) — the footprint of all resident threads exceeds L2 and the
// kernel has too few warps to hide the HBM latency (see also REC_W).
// The accumulators g[] stay in registers during the adjoint sweep (updating them in shared memory
// serialised on the LDS latency: ncu v2 short_scoreboard); they are parked in shared memory only around a
// forward replay and for the final block reduction.  Layout:
// [0,W) dW1[:,0]; [W,2W) sum dz1; then per hidden layer LH; then W+1 output.
template <class R>
struct TanhEval;
template <>
struct TanhEval<double> { static __device__ __forceinline__ double f(double x, const double* tab, int& nm) { return t_tanh(x, tab, nm); } };
template <>
struct TanhEval<float> { static __device__ __forceinline__ float f(float x, const double* tab, int&) { return m_tanh(x, tab); } };

template <class R>
struct SigmoidEval;
template <>
struct SigmoidEval<double> { static __device__ __forceinline__ double f(double x, const double* tab) { return t_sigmoid(x, tab); } };
template <>
struct SigmoidEval<float> { static __device__ __forceinline__ float f(float x, const double* tab) { return m_sigmoid(x, tab); } };

template <class NS, class R>
__device__ __forceinline__ void mlp_backward(const R* __restrict__ sW, const double* __restrict__ tab, const R (&c)[NS::W], R dG, R dz,
                                             R (&g)[NS::NACC]) {
    constexpr int W = NS::W, D = NS::DEPTH;
    int nanmax = 0;   // the forward pass succeeded on the same inputs: nothing to re-inject here
    R a[D][W];
#pragma unroll
    for (int j = 0; j < W; ++j) a[0][j] = TanhEval<R>::f(fma(sW[j], dG, c[j]), tab, nanmax);
    int off = NS::L1;
#pragma unroll
    for (int l = 1; l < D; ++l) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
            R z = sW[off + W * W + j];
#pragma unroll
            for (int i = 0; i < W; ++i) z = fma(sW[off + i * W + j], a[l - 1][i], z);
            a[l][j] = TanhEval<R>::f(z, tab, nanmax);
        }
        off += NS::LH;
    }
    {   // dz arrives as the bare node weight: times d softplus(z_out) = sigmoid(z_out)
        R z = sW[off + W];
#pragma unroll
        for (int i = 0; i < W; ++i) z = fma(sW[off + i], a[D - 1][i], z);
        dz *= SigmoidEval<R>::f(z, tab);
    }
    R da[W];
    int aoff = 2 * W + (D - 1) * NS::LH;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        g[aoff + i] = fma(dz, a[D - 1][i], g[aoff + i]);
        da[i] = dz * sW[off + i];
    }
    g[aoff + W] += dz;
#pragma unroll
    for (int l = D - 1; l >= 1; --l) {
        off -= NS::LH;
        aoff -= NS::LH;
        R dzl[W], dprev[W];
#pragma unroll
        for (int j = 0; j < W; ++j) dzl[j] = da[j] * fma(-a[l][j], a[l][j], R(1));
#pragma unroll
        for (int i = 0; i < W; ++i) {
            R s = R(0);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                g[aoff + i * W + j] = fma(dzl[j], a[l - 1][i], g[aoff + i * W + j]);
                s = fma(sW[off + i * W + j], dzl[j], s);
            }
            dprev[i] = s;
        }
#pragma unroll
        for (int j = 0; j < W; ++j) { g[aoff + W * W + j] += dzl[j]; da[j] = dprev[j]; }
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const R dz1 = da[j] * fma(-a[0][j], a[0][j], R(1));
        g[j] = fma(dz1, dG, g[j]);
        g[W + j] += dz1;
    }
}

// ---------------------------------------------------------------- weights in constant memory (WC instantiations)
// Every lane of a block uses the same network, so its weights can be *uniform* operands: from constant memory they are
// loaded into uniform registers (LDCU c[3][UR + off]) and enter the DFMAs as UR operands — no shared-memory load, no
// vector-register operand.  (A DFMA with three vector-register operands issues at 74 % of the rate of one with a
// uniform operand on B200, cude_measure_fp64_peak_rrr.)  cude_eval_dev copies the call's weights device-to-device
// into this array on the stream before the launch; calls with more weights than fit use the shared-memory path.
#ifndef CUDE_WCONST_DOUBLES
#define CUDE_WCONST_DOUBLES 5120            // 40 KB of the 64 KB constant space: 138 starts of the 37-parameter network
#endif
#ifdef CUDE_HOST_EMU
static double CW_CONST[CUDE_WCONST_DOUBLES];
#else
__constant__ double CW_CONST[CUDE_WCONST_DOUBLES];
#endif

// ---------------------------------------------------------------- the kernel
struct Kin { double k0, k1, k2, c0, d00, kc; };  // d00 = -(k0+k2), kc = k0*c0

__device__ __forceinline__ void kinetics(const Kin& K, double u0, double u1, double p, double& f0, double& f1) {
    // c_peptide_kinetics! + production on the plasma compartment
    f0 = fma(K.d00, u0, fma(K.k1, u1, K.kc)) + p;
    f1 = fma(-K.k1, u1, K.k2 * u0);
}

// dynamic shared memory (doubles) needed by cude_eval_kernel
__host__ __device__ inline size_t eval_smem_doubles(int P, int NACC, int K, int M, int B, bool grad, bool mixed = false, bool bsens = false) {
    // exp table (256) + weights (+ float copy) + per-thread rows: knots 3K, observations 2M, node values 5, GRAD: dG 5 + residuals M.
    // Kept small on purpose: what shared memory does not take stays L1, which serves the per-thread step ring and the
    // register spills (ncu v8: with 65 KB per block the L1 hit rate was 12 % and every spill reload went to L2).
    // GRAD: at least NACC rows — the accumulators are parked in the (then dead) rows for the final warp reduction.
    size_t rows = (size_t)(3 * K + 2 * M) + 5 + (grad ? (size_t)(5 + M) : 0) + (bsens ? 5 : 0);
    if (grad && rows < (size_t)NACC) rows = (size_t)NACC;
    if (grad && rows < (size_t)P + 1) rows = (size_t)P + 1;   // expanded rows {sse, d/d neural[0..P)} for the warp reduction
    return (size_t)256 + (size_t)((P + 1) & ~1) * (mixed ? 2 : 1) + rows * B;
}

// BSENS (only with GRAD = false, MIXED = false): loss + d sse / d cond by one forward-sensitivity column carried through
// the same steps (frozen step sequence => the same derivative the adjoint produces), no step ring, no backward sweep.
// FBWD (with GRAD): the forward pass — loss, step sequence — stays FP64 bit for bit, only the adjoint's network
// evaluations and gradient accumulators are FP32 (opts.precision = 2): gradients to ~1e-6 instead of ~1e-13.
// The kernel's body for block `bid` of the launch geometry (blockIdx.x, except in the fallback walk below).
template <class NS, bool GRAD, bool MIXED, bool BSENS, bool FBWD, bool WC, bool SPLIT>
__device__ __forceinline__ void cude_eval_block(const EvalArgs& A, const unsigned bid) {
    static_assert(!SPLIT || (!GRAD && !MIXED && !BSENS), "SPLIT (stage 1 of the split gradient pipeline) is a variant of the FP64 loss-only kernel");
    static_assert(!WC || !MIXED, "WC (weights in constant memory) needs an FP64 forward network");
    static_assert(!BSENS || (!GRAD && !MIXED), "BSENS is a variant of the FP64 loss-only kernel");
    static_assert(!FBWD || (GRAD && !MIXED), "FBWD is a variant of the FP64 gradient kernel");
    using namespace tab;
    constexpr int W = NS::W, P = NS::P;
    typedef typename std::conditional<MIXED, float, double>::type R;     // scalar type of the forward network evaluation
    typedef typename std::conditional<MIXED || FBWD, float, double>::type RB;   // ... of the adjoint's network evaluation
    constexpr bool F32COPY = MIXED || FBWD;                              // a float copy of the weights sits behind the doubles
    extern __shared__ double smem[];
    const int B = blockDim.x, tid = threadIdx.x;
    const int N = A.pop.n_ind, K = A.pop.max_knots, M = A.pop.max_obs;

    // ---- shared memory carve-up ----
    double* sTab = smem;                             // [256] 2^(j/256) for the exp core
    double* sWs = sTab + 256;                        // [P] (padded to even) — the start's weights in shared memory (unused with WC)
    double* sKt = sWs + ((P + 1) & ~1) * (F32COPY ? 2 : 1);   // [K][B]  (a float copy of the weights may sit in between)
    double* sKg = sKt + (size_t)K * B;               // [K][B]
    double* sSl = sKg + (size_t)K * B;               // [K][B] (last row unused)
    double* sOt = sSl + (size_t)K * B;               // [M][B] observation times
    double* sOy = sOt + (size_t)M * B;               // [M][B] observed c-peptide
    double* sNode = sOy + (size_t)M * B;             // [5][B] network outputs (forward) / node weights (adjoint)
    double* sDG = sNode + (size_t)5 * B;             // [5][B] dG at the adjoint's nodes (GRAD) / d z_out/d beta at the nodes (BSENS) / dG copy (SPLIT)
    double* sRes = sDG + ((GRAD || BSENS || SPLIT) ? (size_t)5 * B : 0);  // [M][B] residuals (GRAD)

    // ---- which trajectory ----
    long long j, prow = bid;   // prow: this block's row group in `partials`, [start][chunk] order
    int i, s;
    bool active;
    if (A.flat) {
        // flat indexing over the [S x N] batch of one shared network.  flat = 2: individual-major — the 32 lanes of a warp are 32
        // starts of ONE individual, i.e. the same glucose curve and time span, so they take nearly the same number of steps
        // (start-major warps mix individuals: 22.8 of 32 lanes active on the ragged Ohashi + Fujita population of config 2)
        const long long jf = (long long)bid * B + tid;
        active = jf < (long long)N * A.n_starts;
        if (A.flat == 2) { i = active ? (int)(jf / A.n_starts) : 0; s = active ? (int)(jf - (long long)i * A.n_starts) : 0; }
        else { i = active ? (int)(jf % N) : 0; s = active ? (int)(jf / N) : 0; }
        j = (long long)s * N + i;
    } else {
        // chunk-major block order: the blocks resident at any time cover a few chunks of individuals x all starts, so the
        // population data of a chunk is read from HBM once and served from L2 to the other starts (start-major order
        // re-read the whole population per start: 15 GB per launch at 1 M individuals x 64 starts, ncu v11)
        const int c = bid / A.n_starts;
        s = bid - c * A.n_starts;
        prow = (long long)s * A.nchunks + c;
        i = c * B + tid;
        active = i < N;
        if (!active) i = 0;
        else if (A.order) i = (int)(A.order[(size_t)s * N + i] & 0xffffffu);
        j = (long long)s * N + i;
        if constexpr (GRAD) {
            // fallback of the two-kernel / split gradient: only the trajectories stage 1 could not record
            if (A.only_flag) active = active && A.only_flag[j] < 0;
        }
    }
    // ---- the start's weights: staged in shared memory (block-uniform), or read from constant memory (WC) ----
    // (offset from block-uniform values only, so that the compiler keeps it — and the weight loads — on the uniform path)
    const long long wofs = A.flat ? 0 : (long long)(bid % (unsigned)A.n_starts) * A.neural_stride;
    double wuni[WC ? P : 1];         // WC: the weights as block-uniform values, loaded here in convergent code
    if constexpr (WC) {
#pragma unroll
        for (int p = 0; p < P; ++p) wuni[p] = CW_CONST[A.wc_base + wofs + p];
    }
    const double* const sW = WC ? wuni : sWs;
    R* const sWr = MIXED ? reinterpret_cast<R*>(sWs + ((P + 1) & ~1)) : reinterpret_cast<R*>(const_cast<double*>(sW));   // network weights as R
    RB* const sWb = F32COPY ? reinterpret_cast<RB*>(sWs + ((P + 1) & ~1)) : reinterpret_cast<RB*>(const_cast<double*>(sW));   // ... as RB
    {
        const double* gW = A.neural + (A.flat ? 0 : (long long)s * A.neural_stride);
        if (!WC || F32COPY) for (int p = tid; p < P; p += B) { if (!WC) sWs[p] = gW[p]; if (F32COPY) sWb[p] = (RB)gW[p]; }
        for (int p = tid; p < 256; p += B) sTab[p] = EXP_TAB256[p];
    }
    // ---- stage this thread's knots ----
    const int nk = A.pop.n_knots[i];
    for (int k = 0; k < nk; ++k) {
        sKt[k * B + tid] = A.pop.knot_t[(size_t)k * N + i];
        sKg[k * B + tid] = A.pop.knot_g[(size_t)k * N + i];
        if (k < nk - 1) sSl[k * B + tid] = A.pop.slope[(size_t)k * N + i];
    }
    const int nobs = A.pop.n_obs[i];
    for (int k = 0; k < nobs; ++k) {
        sOt[k * B + tid] = A.pop.obs_t[(size_t)k * N + i];
        sOy[k * B + tid] = A.pop.obs_y[(size_t)k * N + i];
    }
    __syncthreads();

    double sse = 0.0, gcond = 0.0;
    int nacc = 0, nrej = 0;
    bool failed = false;
    double beta = 0.0, covv = 0.0;
    double* const myNode = sNode + tid;
    // gradient accumulators (compressed layout, see mlp_backward): registers for the whole adjoint sweep and the
    // warp reduction at the end; zero for inactive / failed lanes
    // After the solve the thread's knot / observation / node rows are dead: its gradient accumulators are parked there
    // ([k][tid], >= NACC rows by eval_smem_doubles) for the rolled warp reduction at the end.
    double* const myAcc = sKt + tid;
    double* const myDG = sDG + tid;

    if (active) {
        Kin Kc;
        Kc.k0 = A.pop.k0[i]; Kc.k1 = A.pop.k1[i]; Kc.k2 = A.pop.k2[i]; Kc.c0 = A.pop.c0[i];
        Kc.d00 = -(Kc.k0 + Kc.k2); Kc.kc = Kc.k0 * Kc.c0;
        Knots kn;
        kn.t = sKt + tid; kn.g = sKg + tid; kn.sl = sSl + tid; kn.nk = nk; kn.stride = B;
        kn.g0 = kn.g[0];
        const double t0 = kn.t[0], tend = kn.t[(nk - 1) * B];
        const double* const obs_t = sOt + tid;
        const double* const obs_y = sOy + tid;

        // conditional_production: beta = exp(cond); first-layer constant part
        beta = m_exp(A.cond[j]);
        covv = (NS::NIN > 2 && A.pop.cov) ? A.pop.cov[i] : 0.0;
        double c[W];
#pragma unroll
        for (int q = 0; q < W; ++q) {
            double z = fma(sW[W + q], beta, sW[NS::NIN * W + q]);
            if (NS::NIN > 2) z = fma(sW[2 * W + q], covv, z);
            c[q] = z;
        }
        R cr[W];
        RB cb[W];
#pragma unroll
        for (int q = 0; q < W; ++q) { cr[q] = (R)c[q]; cb[q] = (RB)c[q]; }

        const double abstol = A.abstol, reltol = A.reltol;
        const double dtmax = tend - t0;
        const double at0 = fabs(t0), at1 = fabs(tend);
        const double dtmin = fmax(nextafter(at0, CUDART_INF) - at0, nextafter(at1, CUDART_INF) - at1);
        const double snap = 100.0 * (nextafter(at1, CUDART_INF) - at1);

        double rec[GRAD ? REC_CAP * REC_W : 1];   // accepted-step records {t, dt}
        RB acc[GRAD ? NS::NACC : 1];          // gradient accumulators (compressed layout, see mlp_backward): registers during the adjoint sweep
        volatile double park[GRAD ? NS::NACC : 1];   // ... local memory across a forward replay (solves longer than REC_CAP
                                                     // steps: rare); volatile keeps it out of the register allocation
        int stop_at = 0x7fffffff;   // replay limit (GRAD)
        // adjoint carry
        double lam0 = 0.0, lam1 = 0.0, wnode = 0.0, wsum = 0.0, t_next = tend;
        int kobs_top = nobs - 1;
        double top_ot = (nobs > 0) ? obs_t[(nobs - 1) * B] : -CUDART_INF;   // time of observation kobs_top
        bool first_pass = true;

        do {
            // =================== forward pass (or replay up to stop_at accepted steps) ===================
            double u0 = Kc.c0, u1 = (Kc.k2 / Kc.k1) * Kc.c0;   // c-peptide-models.jl:185
            double t = t0;
            int iobs = 0, na = 0, nr = 0;
            double fsse = 0.0;
            double next_ot = (nobs > 0) ? obs_t[0] : CUDART_INF;
            int kbase = 0, klast = 0;      // glucose interval of the step's start time / of its end node
            double s0 = 0.0, s1 = 0.0, js10 = 0.0, js11 = 0.0, dnn0 = 0.0, gsens = 0.0;   // BSENS: d u/d cond, its FSAL stage, d NN0/d beta
            // save_start: observations at (or before) t0 see u0
            while (iobs < nobs && next_ot <= t0) {
                const double r = u0 - obs_y[iobs * B];
                if (GRAD) sRes[iobs * B + tid] = r;
                if constexpr (SPLIT) A.sp_res[(size_t)iobs * ((size_t)N * A.n_starts) + j] = r;
                if constexpr (!GRAD && !BSENS) { if (A.yhat_out) A.yhat_out[(size_t)j * M + iobs] = u0; }
                fsse = fma(r, r, fsse);
                ++iobs;
                next_ot = (iobs < nobs) ? obs_t[iobs * B] : CUDART_INF;
            }
            // production at t0 is NN([0;beta]) - NN([0;beta]) = 0 exactly (dG(t0) = 0)
            double k10, k11;
            kinetics(Kc, u0, u1, 0.0, k10, k11);
            // ---- Hairer initial step (ode_determine_initdt), part 1 ----
            const double sk0 = fma(fabs(u0), reltol, abstol), sk1 = fma(fabs(u1), reltol, abstol);
            const double isk0 = 1.0 / sk0, isk1 = 1.0 / sk1;
            double dt0, d1;
            {
                double x0 = u0 * isk0, x1 = u1 * isk1;
                const double d0 = sqrt(m_sumsq(x0, x1) * 0.5);
                x0 = k10 * isk0; x1 = k11 * isk1;
                d1 = sqrt(m_sumsq(x0, x1) * 0.5);
                dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
                dt0 = fmin(dt0, dtmax);
            }
            double dt = dt0, nn0 = 0.0, lnqold = -9.210340371976182;   // ln(qoldinit = 1e-4)
            int ret = 0, iter = 0;
            bool init = true;
            for (;;) {
                // ---- which nodes does this pass of the (single, shared) network-evaluation loop serve? ----
                int nq;
                const double* cn;
                if (init) { nq = 2; cn = CN_INIT; }
                else {
                    if (!(ret == 0 && t < tend && na < stop_at)) break;
                    if (++iter > A.maxiters) { ret = 1; break; }
                    dt = fmin(dt, tend - t);                       // modify_dt_for_tstops!
                    if (!(dt > dtmin)) { ret = (dt != dt) ? 3 : 2; break; }
                    nq = 5; cn = CN_STEP;
                }
                {
                    double tau[5];
#pragma unroll
                    for (int q = 0; q < 5; ++q) tau[q] = fma(cn[q], dt, t);
                    kn.dG5(tau, myNode, B, kbase, klast);         // dG of the node times, staged in myNode
                    if constexpr (SPLIT) {                         // kept for the step record (myNode is overwritten by z_out)
#pragma unroll
                        for (int q = 0; q < 5; ++q) myDG[q * B] = myNode[q * B];
                    }
                }
                if constexpr (BSENS) {
#pragma unroll 1
                    for (int q = 0; q < 5; ++q)
                        if (q < nq) { double dzb; myNode[q * B] = mlp_zout_dbeta<NS>(sW, sTab, c, myNode[q * B], dzb); myDG[q * B] = dzb; }
                } else {
                    CUDE_UNROLL(CUDE_FWD_UNROLL)
                    for (int q = 0; q < 5; ++q)
                        if (q < nq) myNode[q * B] = (double)mlp_zout<NS>(sWr, sTab, cr, (R)myNode[q * B]);
                }
                // softplus of the 5 nodes together (init: entries >= nq hold dG values — evaluated and ignored)
                double sp[5], dd[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    if (MIXED) softplus_d_mixed(myNode[q * B], sTab, sp[q], dd[q]);
                    else t_softplus_d(myNode[q * B], sTab, sp[q], dd[q]);
                }
                if (init) {
                    // ---- Hairer, part 2: probe f(u0 + dt0 f0, t0 + dt0) ----
                    init = false;
                    nn0 = sp[0];                                  // network([0; beta]) — identical at every call
                    if (BSENS) dnn0 = fma(-1.0, m_rcp(dd[0]), 1.0) * myDG[0];   // d network([0; beta]) / d beta
                    const double pe = sp[1] - nn0;
                    double f0, f1;
                    kinetics(Kc, fma(dt0, k10, u0), fma(dt0, k11, u1), pe, f0, f1);
                    const double x0 = (f0 - k10) * isk0, x1 = (f1 - k11) * isk1;
                    const double d2 = sqrt(m_sumsq(x0, x1) * 0.5) / dt0;
                    const double dm = fmax(d1, d2);
                    const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : t_pow10(-(2.0 + m_log10(dm)) / 5.0, sTab);
                    dt = fmax(dtmin, fmin(fmin(100.0 * dt0, dt1), dtmax));
                    if (!(isfinite(dt) && isfinite(k10) && isfinite(k11) && isfinite(f0) && isfinite(nn0))) ret = 3;
                    continue;
                }
                const double p2 = sp[0] - nn0, p3 = sp[1] - nn0, p4 = sp[2] - nn0,
                             p5 = sp[3] - nn0, p6 = sp[4] - nn0;   // p6: stages 6, 7 and the next k1
                // ---- stages: the kinetics are linear, the production enters additively.  "Push" form: as soon as a stage
                //      derivative k_j exists it is added to the partial sums of every later stage, so the dependent chain per
                //      stage is sum-update -> g -> kinetics (5 DFMA) instead of a j-term sum (ncu v9: this region ran at
                //      6.6 cycles per instruction on fixed-latency waits) ----
                double f0, f1, g0, g1;
                double s30 = a31 * k10, s31 = a31 * k11, s40 = a41 * k10, s41 = a41 * k11, s50 = a51 * k10, s51 = a51 * k11,
                       s60 = a61 * k10, s61 = a61 * k11, sb0 = b1 * k10, sb1 = b1 * k11, se0 = e1 * k10, se1 = e1 * k11;
                g0 = fma(dt * a21, k10, u0); g1 = fma(dt * a21, k11, u1);
                double k20, k21; kinetics(Kc, g0, g1, p2, k20, k21);
                s30 = fma(a32, k20, s30); s31 = fma(a32, k21, s31); s40 = fma(a42, k20, s40); s41 = fma(a42, k21, s41);
                s50 = fma(a52, k20, s50); s51 = fma(a52, k21, s51); s60 = fma(a62, k20, s60); s61 = fma(a62, k21, s61);
                sb0 = fma(b2, k20, sb0); sb1 = fma(b2, k21, sb1); se0 = fma(e2, k20, se0); se1 = fma(e2, k21, se1);
                g0 = fma(dt, s30, u0); g1 = fma(dt, s31, u1);
                double k30, k31; kinetics(Kc, g0, g1, p3, k30, k31);
                s40 = fma(a43, k30, s40); s41 = fma(a43, k31, s41); s50 = fma(a53, k30, s50); s51 = fma(a53, k31, s51);
                s60 = fma(a63, k30, s60); s61 = fma(a63, k31, s61);
                sb0 = fma(b3, k30, sb0); sb1 = fma(b3, k31, sb1); se0 = fma(e3, k30, se0); se1 = fma(e3, k31, se1);
                g0 = fma(dt, s40, u0); g1 = fma(dt, s41, u1);
                double k40, k41; kinetics(Kc, g0, g1, p4, k40, k41);
                s50 = fma(a54, k40, s50); s51 = fma(a54, k41, s51); s60 = fma(a64, k40, s60); s61 = fma(a64, k41, s61);
                sb0 = fma(b4, k40, sb0); sb1 = fma(b4, k41, sb1); se0 = fma(e4, k40, se0); se1 = fma(e4, k41, se1);
                g0 = fma(dt, s50, u0); g1 = fma(dt, s51, u1);
                double k50, k51; kinetics(Kc, g0, g1, p5, k50, k51);
                s60 = fma(a65, k50, s60); s61 = fma(a65, k51, s61);
                sb0 = fma(b5, k50, sb0); sb1 = fma(b5, k51, sb1); se0 = fma(e5, k50, se0); se1 = fma(e5, k51, se1);
                g0 = fma(dt, s60, u0); g1 = fma(dt, s61, u1);
                double k60, k61; kinetics(Kc, g0, g1, p6, k60, k61);
                sb0 = fma(b6, k60, sb0); sb1 = fma(b6, k61, sb1); se0 = fma(e6, k60, se0); se1 = fma(e6, k61, se1);
                const double un0 = fma(dt, sb0, u0), un1 = fma(dt, sb1, u1);
                double k70, k71; kinetics(Kc, un0, un1, p6, k70, k71);
                // ---- error estimate ----
                f0 = dt * fma(e7, k70, se0);
                f1 = dt * fma(e7, k71, se1);
                // ---- BSENS: the same stages for s = d u / d cond: s' = A s + e1 d prod/d cond, d prod/d cond =
                //      beta (sigmoid(z_out) dzb - d NN([0;beta])/d beta) at the step's nodes ----
                double sn0 = 0.0, sn1 = 0.0, j20 = 0.0, j30 = 0.0, j40 = 0.0, j50 = 0.0, j60 = 0.0, j70 = 0.0, j71 = 0.0;
                if constexpr (BSENS) {
                    double dq[5];
#pragma unroll
                    for (int q = 0; q < 5; ++q) dq[q] = beta * fma(fma(-1.0, m_rcp(dd[q]), 1.0), myDG[q * B], -dnn0);
                    double j21, j31, j41, j51, j61, h0, h1;
#define CUDE_SKIN(G0, G1, DP, F0, F1) F0 = fma(Kc.d00, G0, fma(Kc.k1, G1, DP)); F1 = fma(-Kc.k1, G1, Kc.k2 * (G0));
                    h0 = fma(dt * a21, js10, s0); h1 = fma(dt * a21, js11, s1);
                    CUDE_SKIN(h0, h1, dq[0], j20, j21)
                    h0 = fma(dt, fma(a31, js10, a32 * j20), s0); h1 = fma(dt, fma(a31, js11, a32 * j21), s1);
                    CUDE_SKIN(h0, h1, dq[1], j30, j31)
                    h0 = fma(dt, fma(a41, js10, fma(a42, j20, a43 * j30)), s0); h1 = fma(dt, fma(a41, js11, fma(a42, j21, a43 * j31)), s1);
                    CUDE_SKIN(h0, h1, dq[2], j40, j41)
                    h0 = fma(dt, fma(a51, js10, fma(a52, j20, fma(a53, j30, a54 * j40))), s0);
                    h1 = fma(dt, fma(a51, js11, fma(a52, j21, fma(a53, j31, a54 * j41))), s1);
                    CUDE_SKIN(h0, h1, dq[3], j50, j51)
                    h0 = fma(dt, fma(a61, js10, fma(a62, j20, fma(a63, j30, fma(a64, j40, a65 * j50)))), s0);
                    h1 = fma(dt, fma(a61, js11, fma(a62, j21, fma(a63, j31, fma(a64, j41, a65 * j51)))), s1);
                    CUDE_SKIN(h0, h1, dq[4], j60, j61)
                    sn0 = fma(dt, fma(b1, js10, fma(b2, j20, fma(b3, j30, fma(b4, j40, fma(b5, j50, b6 * j60))))), s0);
                    sn1 = fma(dt, fma(b1, js11, fma(b2, j21, fma(b3, j31, fma(b4, j41, fma(b5, j51, b6 * j61))))), s1);
                    CUDE_SKIN(sn0, sn1, dq[4], j70, j71)
#undef CUDE_SKIN
                }
                f0 = f0 * m_rcp(fma(fmax(fabs(u0), fabs(un0)), reltol, abstol));   // denominators >= abstol > 0
                f1 = f1 * m_rcp(fma(fmax(fabs(u1), fabs(un1)), reltol, abstol));
                // EEst = sqrt(E2); accept iff EEst <= 1 iff E2 <= 1; the controller only needs ln EEst = ln(E2)/2
                const double E2 = m_sumsq(f0, f1) * 0.5;
                if (!(E2 == E2) || !isfinite(un0) || !isfinite(un1)) { ret = 3; break; }
                CUDE_TRACE_STEP(t, dt, sqrt(E2))
                // ---- PI controller: q = EEst^beta1 / qold^beta2 / gamma, clamped to [1/qmax, 1/qmin];
                //      E2 == 0 gives ln = -690 -> q saturates at 1/qmax like the reference's explicit branch ----
                const double lnE = 0.5 * m_log_pos(E2);
                if (E2 <= 1.0) {
                    const double q = fmax(1.0 / qmax, fmin(1.0 / qmin, t_exp_sat(fma(beta1, lnE, -beta2 * lnqold), sTab) * (1.0 / gamma)));
                    double tnew = t + dt;
                    if (fabs(tnew - tend) < snap) tnew = tend;
                    // saveat by dense output: observation times in (t, tnew]
                    while (iobs < nobs && next_ot <= tnew) {
                        double y;
                        if (next_ot == tnew) y = un0;
                        else {
                            double bw[7];
                            dense_weights((next_ot - t) * m_rcp(dt), bw);
                            const double sdo = fma(bw[0], k10, fma(bw[1], k20, fma(bw[2], k30, fma(bw[3], k40, fma(bw[4], k50, fma(bw[5], k60, bw[6] * k70))))));
                            y = fma(dt, sdo, u0);
                        }
                        const double r = y - obs_y[iobs * B];
                        if (GRAD) sRes[iobs * B + tid] = r;
                        if constexpr (SPLIT) A.sp_res[(size_t)iobs * ((size_t)N * A.n_starts) + j] = r;
                        if constexpr (!GRAD && !BSENS) { if (A.yhat_out) A.yhat_out[(size_t)j * M + iobs] = y; }
                        if constexpr (BSENS) {
                            double dy;
                            if (next_ot == tnew) dy = sn0;
                            else {
                                double bw[7];
                                dense_weights((next_ot - t) * m_rcp(dt), bw);
                                dy = fma(dt, fma(bw[0], js10, fma(bw[1], j20, fma(bw[2], j30, fma(bw[3], j40, fma(bw[4], j50, fma(bw[5], j60, bw[6] * j70)))))), s0);
                            }
                            gsens = fma(2.0 * r, dy, gsens);
                        }
                        fsse = fma(r, r, fsse);
                        ++iobs;
                        next_ot = (iobs < nobs) ? obs_t[iobs * B] : CUDART_INF;
                    }
                    if (BSENS) { s0 = sn0; s1 = sn1; js10 = j70; js11 = j71; }
                    if (GRAD) {
                        double* const r7 = rec + (na % REC_CAP) * REC_W;
                        r7[0] = t; r7[1] = dt;
                    }
                    if constexpr (SPLIT) {
                        if (na < SPLIT_CAP) {      // 64-byte record = 2 whole sectors: {t, h}, {dG0, dG1}, {dG2, dG3}, {dG4, 0}
                            double2* const r = reinterpret_cast<double2*>(A.sp_rec + ((size_t)j * SPLIT_CAP + na) * SPLIT_W);
                            r[0] = make_double2(t, dt);
                            r[1] = make_double2(myDG[0], myDG[B]);
                            r[2] = make_double2(myDG[2 * B], myDG[3 * B]);
                            r[3] = make_double2(myDG[4 * B], 0.0);
                        }
                    }
                    ++na;
                    lnqold = fmax(lnE, -9.210340371976182);        // qold = max(EEst, qoldinit)
                    dt = fmin(dt * m_rcp(q), dtmax);               // q in [1/qmax, 1/qmin]
                    t = tnew; u0 = un0; u1 = un1; k10 = k70; k11 = k71;   // FSAL
                    kbase = klast;
                } else {
                    ++nr;
                    dt = dt * m_rcp(fmin(1.0 / qmin, t_exp_sat(beta1 * lnE, sTab) * (1.0 / gamma)));
                }
            }
            if (first_pass) {
                if (ret == 0 && iobs < nobs) ret = 3;   // observation beyond tend: not produced by saveat
                nacc = na; nrej = nr;
                failed = (ret != 0);
                sse = failed ? CUDART_INF : fsse;
                if (BSENS) gcond = failed ? 0.0 : gsens;
                stop_at = na;
            }
            const bool was_first = first_pass;
            first_pass = false;
            if (!GRAD || failed) break;
            if constexpr (GRAD) {
            // accumulators: zero after the first pass, otherwise back from local memory (parked for the replay)
#pragma unroll
            for (int k = 0; k < NS::NACC; ++k) acc[k] = was_first ? RB(0) : (RB)park[k];
            // =================== adjoint over steps [lo, stop_at) held in the ring ===================
            // The last chunk appends the virtual step n = -1: the NN([0;beta]) term, one node at dG = 0
            // with weight -sum(w) (the node t0 itself has dG = 0 and cancels exactly).
            const int lo = (stop_at > REC_CAP) ? stop_at - REC_CAP : 0;
            const int nlast = (lo == 0) ? -1 : lo;
            int kb = nk - 2 > 0 ? nk - 2 : 0, kdummy;   // glucose interval of the step's start time (times decrease)
            double rn_t = rec[((stop_at - 1) % REC_CAP) * REC_W], rn_h = rec[((stop_at - 1) % REC_CAP) * REC_W + 1];   // (t, dt) fetched one step ahead
            for (int n = stop_at - 1; n >= nlast; --n) {
                int nq;
                const double* cn;
                double tn, h;
                if (n < 0) {
                    nq = 1; cn = CN_INIT; tn = t0; h = 0.0;
                    myNode[0] = -wsum;
                } else {
                    tn = rn_t; h = rn_h;
                    if (n > lo) { rn_t = rec[((n - 1) % REC_CAP) * REC_W]; rn_h = rec[((n - 1) % REC_CAP) * REC_W + 1]; }
                    nq = 5; cn = CN_STEP;
                    double kb[7][2];
#pragma unroll
                    for (int q = 0; q < 7; ++q) { kb[q][0] = 0.0; kb[q][1] = 0.0; }
                    double ub0 = 0.0, ub1 = 0.0;
                    // observations in (tn, t_next]
                    while (kobs_top >= 0) {
                        const double ts = top_ot;                 // kept in a register: no global load per step
                        if (!(ts > tn)) break;
                        const double wr = 2.0 * sRes[kobs_top * B + tid];
                        if (ts == t_next) lam0 += wr;
                        else {
                            double bw[7];
                            dense_weights((ts - tn) * m_rcp(h), bw);
                            ub0 += wr;
                            const double wh = wr * h;
#pragma unroll
                            for (int q = 0; q < 7; ++q) kb[q][0] = fma(wh, bw[q], kb[q][0]);
                        }
                        --kobs_top;
                        top_ot = (kobs_top >= 0) ? obs_t[kobs_top * B] : -CUDART_INF;
                    }
                    // k7 = A un + b + e1 p7 (dense output only): lam += A^T kb7
                    const double pb7 = kb[6][0];
                    lam0 = fma(Kc.d00, kb[6][0], lam0);   // kb7[1] == 0
                    lam1 = fma(Kc.k1, kb[6][0], lam1);
                    // un = u + h sum b_j k_j
                    ub0 += lam0; ub1 += lam1;
                    {
                        const double hl0 = h * lam0, hl1 = h * lam1;
                        kb[0][0] = fma(b1, hl0, kb[0][0]); kb[0][1] = fma(b1, hl1, kb[0][1]);
                        kb[1][0] = fma(b2, hl0, kb[1][0]); kb[1][1] = fma(b2, hl1, kb[1][1]);
                        kb[2][0] = fma(b3, hl0, kb[2][0]); kb[2][1] = fma(b3, hl1, kb[2][1]);
                        kb[3][0] = fma(b4, hl0, kb[3][0]); kb[3][1] = fma(b4, hl1, kb[3][1]);
                        kb[4][0] = fma(b5, hl0, kb[4][0]); kb[4][1] = fma(b5, hl1, kb[4][1]);
                        kb[5][0] = fma(b6, hl0, kb[5][0]); kb[5][1] = fma(b6, hl1, kb[5][1]);
                    }
                    // stage i: k_i = A g_i + b + e1 p_i, g_i = u + h sum_{j<i} a_ij k_j
                    // gb = A^T kb_i = [d00*kb0 + k2*kb1, k1*kb0 - k1*kb1]
                    double gb0, gb1, hg0, hg1;
#define CUDE_STAGE_BACK(I)                                                    \
    gb0 = fma(Kc.d00, kb[I][0], Kc.k2 * kb[I][1]);                            \
    gb1 = Kc.k1 * (kb[I][0] - kb[I][1]);                                      \
    ub0 += gb0; ub1 += gb1; hg0 = h * gb0; hg1 = h * gb1;
#define CUDE_PUSH(J, COEF) kb[J][0] = fma(COEF, hg0, kb[J][0]); kb[J][1] = fma(COEF, hg1, kb[J][1]);
                    const double pb6 = kb[5][0];
                    CUDE_STAGE_BACK(5) CUDE_PUSH(0, a61) CUDE_PUSH(1, a62) CUDE_PUSH(2, a63) CUDE_PUSH(3, a64) CUDE_PUSH(4, a65)
                    const double pb5 = kb[4][0];
                    CUDE_STAGE_BACK(4) CUDE_PUSH(0, a51) CUDE_PUSH(1, a52) CUDE_PUSH(2, a53) CUDE_PUSH(3, a54)
                    const double pb4 = kb[3][0];
                    CUDE_STAGE_BACK(3) CUDE_PUSH(0, a41) CUDE_PUSH(1, a42) CUDE_PUSH(2, a43)
                    const double pb3 = kb[2][0];
                    CUDE_STAGE_BACK(2) CUDE_PUSH(0, a31) CUDE_PUSH(1, a32)
                    const double pb2 = kb[1][0];
                    CUDE_STAGE_BACK(1) CUDE_PUSH(0, a21)
                    const double pb1 = kb[0][0];
                    CUDE_STAGE_BACK(0)
#undef CUDE_STAGE_BACK
#undef CUDE_PUSH
                    (void)hg0; (void)hg1;
                    // node weights, in CN_STEP order; node tn+h serves stages 6, 7 and the next step's stage 1
                    const double w6 = pb6 + pb7 + wnode;
                    // the bare node weights; mlp_backward multiplies by d softplus(z_out)
                    myNode[0] = pb2; myNode[B] = pb3; myNode[2 * B] = pb4; myNode[3 * B] = pb5; myNode[4 * B] = w6;
                    wsum += w6 + pb5 + pb4 + pb3 + pb2;
                    wnode = pb1;
                    lam0 = ub0; lam1 = ub1;
                    t_next = tn;
                }
                // ---- network gradient at the nodes (single, shared forward+backward evaluation loop) ----
                {
                    double tau[5];
#pragma unroll
                    for (int q = 0; q < 5; ++q) tau[q] = fma(cn[q], h, tn);
                    while (kb > 0 && kn.t[kb * B] > tn) --kb;
                    kn.dG5(tau, myDG, B, kb, kdummy);
                }
                CUDE_UNROLL(CUDE_BWD_UNROLL)
                for (int q = 0; q < 5; ++q)
                    if (q < nq) mlp_backward<NS, RB>(sWb, sTab, cb, (RB)myDG[q * B], (RB)myNode[q * B], acc);
            }
            stop_at = lo;   // steps below lo still to do: replay the forward pass up to lo
            if (stop_at > 0) {   // park the accumulators for the replay's register budget
#pragma unroll
                for (int k = 0; k < NS::NACC; ++k) park[k] = (double)acc[k];
            } else {             // done: hand them to the warp reduction through the thread's (now dead) shared rows
#pragma unroll
                for (int k = 0; k < NS::NACC; ++k) myAcc[k * B] = (double)acc[k];
            }
            }  // if constexpr (GRAD)
        } while (stop_at > 0);

        if constexpr (GRAD) if (!failed) {
            // d sse / d cond = (sum_q dz1_q * W1[q,1]) * beta      (beta = exp(cond))
            double db = 0.0;
#pragma unroll
            for (int q = 0; q < W; ++q) db = fma(myAcc[(W + q) * B], sW[W + q], db);
            gcond = db * beta;
            // expand the compressed accumulators in place to the SimpleChains layout, rows 1..P (row 0: sse).
            // Descending p: row 1+p is written after every compressed row it could alias has been read
            // (compressed row of p is <= p, see the index map below).
#pragma unroll
            for (int p = P - 1; p >= 0; --p) {
                double v;
                if (p < W) v = myAcc[p * B];                                              // W1[:,0]  (dG column)
                else if (p < 2 * W) v = myAcc[(W + (p - W)) * B] * beta;                  // W1[:,1]  (beta column)
                else if (NS::NIN > 2 && p < 3 * W) v = myAcc[(W + (p - 2 * W)) * B] * covv;   // W1[:,2] (covariate)
                else if (p < NS::L1) v = myAcc[(W + (p - NS::NIN * W)) * B];              // b1
                else v = myAcc[(2 * W + (p - NS::L1)) * B];                               // hidden + output layers
                myAcc[(1 + p) * B] = v;
            }
        }
    }
    if constexpr (GRAD) {
        if (!active || failed) {
#pragma unroll 1
            for (int p = 0; p < P; ++p) myAcc[(1 + p) * B] = 0.0;
        }
        myAcc[0] = active ? sse : 0.0;
    }

    // ---- outputs ----
    if (active) {
        if (A.keys_out) {
            const int ns = SPLIT ? nacc : nacc + nrej;      // stage 1 of the two-kernel gradient: the adjoint's length (accepted steps)
            A.keys_out[j] = ((unsigned int)(ns < 255 ? ns : 255) << 24) | (unsigned int)i;
            if constexpr (SPLIT) if (A.keys16_out) A.keys16_out[j] = (unsigned short)(((unsigned int)s << 8) | (unsigned int)(ns < 255 ? ns : 255));
        }
        if (A.sse_out) A.sse_out[j] = sse;
        if constexpr (SPLIT) {
            A.sp_nrec[j] = failed ? 0 : (nacc > SPLIT_CAP ? -1 : nacc);
            A.sp_beta[j] = beta;
            A.sp_sse[j] = sse;
            if (!failed && nacc > SPLIT_CAP && atomicExch(&A.sp_blkflag[prow], 1) == 0)
                A.sp_blklist[atomicAdd(A.sp_blkcount, 1)] = (int)bid;
        }
        if constexpr (!GRAD && !BSENS) {
            if (A.yhat_out && failed) for (int k = 0; k < M; ++k) A.yhat_out[(size_t)j * M + k] = CUDART_NAN;   // no solution
        }
        if ((GRAD || BSENS) && A.g_cond) A.g_cond[j] = failed ? 0.0 : gcond * A.cond_scale;
    }
    // ---- warp reduction: {sse, d sse/d neural[0..P)}; one partial row per WARP, no block barrier (warps of a
    //      block finish at different times; a barrier here idled their slots: ncu v4 epilogue 45 % barrier) ----
    const int lane = tid & 31, wid = tid >> 5, nw = (B + 31) >> 5;
    if (A.partials && !SPLIT) {
        double* const row = A.partials + ((size_t)prow * nw + wid) * (P + 1);
        if constexpr (GRAD) {
            // rows [q][tid] of this warp's 32 columns -> lane l sums row l (and row 32 + l): 32 shared loads and adds
            // per lane instead of 5 shuffle stages per row (the shuffle version was 4 % of the kernel, ncu v9).
            // Column order (j + lane) & 31: lanes l and l + 16 share a bank, which 64-bit accesses tolerate.
            __syncwarp();
            const int wbase = tid & ~31;
            const int nl = (B - wbase < 32) ? B - wbase : 32;      // 32 unless the host emulation runs one thread
#pragma unroll 1
            for (int q = lane; q < P + 1; q += nl) {
                const double* const src = smem + (myAcc - tid - smem) + (size_t)q * B + wbase;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
                for (int jj = 0; jj < 32; jj += 4) {
                    const int j0 = (jj + lane) & 31, j1 = (jj + 1 + lane) & 31, j2 = (jj + 2 + lane) & 31, j3 = (jj + 3 + lane) & 31;
                    s0 += (j0 < nl) ? src[j0] : 0.0;
                    s1 += (j1 < nl) ? src[j1] : 0.0;
                    s2 += (j2 < nl) ? src[j2] : 0.0;
                    s3 += (j3 < nl) ? src[j3] : 0.0;
                }
                row[q] = (s0 + s1) + (s2 + s3);
            }
        } else {
            double v = active ? sse : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) row[0] = v;
        }
    }
    if (A.counters) {
        unsigned int ca = active ? (unsigned)nacc : 0u, cr = active ? (unsigned)nrej : 0u, cf = (active && failed) ? 1u : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ca += __shfl_xor_sync(0xffffffffu, ca, o);
            cr += __shfl_xor_sync(0xffffffffu, cr, o);
            cf += __shfl_xor_sync(0xffffffffu, cf, o);
        }
        if (lane == 0) {
            atomicAdd(&A.counters[0], (unsigned long long)ca);
            atomicAdd(&A.counters[1], (unsigned long long)cr);
            if (cf) atomicAdd(&A.counters[2], (unsigned long long)cf);
        }
    }
}

template <class NS, bool GRAD, bool MIXED = false, bool BSENS = false, bool FBWD = false, bool WC = false, bool SPLIT = false>
__global__ void __launch_bounds__(CUDE_MAX_THREADS, GRAD ? CUDE_MIN_BLOCKS : ((WC && !BSENS) ? CUDE_MIN_BLOCKS_LOSS_WC : CUDE_MIN_BLOCKS_LOSS))
cude_eval_kernel(const EvalArgs A) {
    if constexpr (GRAD) {
        if (A.only_flag) {
            // fallback launch: a grid of resident blocks walks the list of blocks stage 1 flagged (usually empty) instead of
            // 62 500 blocks per group looking up their flag (1.7 % of a bench step)
            const unsigned cnt = (unsigned)*A.sp_blkcount;
            for (unsigned v = blockIdx.x; v < cnt; v += gridDim.x) {
                cude_eval_block<NS, GRAD, MIXED, BSENS, FBWD, WC, SPLIT>(A, (unsigned)A.sp_blklist[v]);
                __syncthreads();
            }
            return;
        }
    }
    cude_eval_block<NS, GRAD, MIXED, BSENS, FBWD, WC, SPLIT>(A, blockIdx.x);
}

// Second stage (tile mode): sums[(P+1) x S] column-major, sums[q + (P+1)*s] = sum over the start's
// chunks of partials[(s*nchunks + c)*(P+1) + q], in fixed chunk order (deterministic).
// One warp per (start, q-group); lanes stride over chunks then shuffle.
__global__ void cude_reduce_partials(const double* __restrict__ partials, int nchunks, int n_starts, int np1,
                                     int nred, double* __restrict__ sums) {
    const int warps_per_block = blockDim.x >> 5;
    const long long gw = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)n_starts * np1) return;
    const int s = (int)(gw / np1), q = (int)(gw % np1);
    double v = 0.0;
    if (q < nred) {
        const double* base = partials + (size_t)s * nchunks * np1 + q;
        for (int c = lane; c < nchunks; c += 32) v += base[(size_t)c * np1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    if (lane == 0) sums[(size_t)s * np1 + q] = v;
}

// The same for a few rows per start (small populations x many starts: the suppression example's 10 000 starts x 2 rows took
// 108 us with a warp per (start, q)): one thread per (start, q), rows in order, consecutive threads on consecutive q.
__global__ void cude_reduce_partials_few(const double* __restrict__ partials, int nchunks, int n_starts, int np1,
                                         int nred, double* __restrict__ sums) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)n_starts * np1) return;
    const int s = (int)(g / np1), q = (int)(g - (long long)s * np1);
    double v = 0.0;
    if (q < nred) {
        const double* base = partials + (size_t)s * nchunks * np1 + q;
        for (int c = 0; c < nchunks; ++c) v += base[(size_t)c * np1];
    }
    sums[g] = v;
}

// flat mode: per-start sse sums from the per-trajectory sse array (loss-only / beta-only calls)
__global__ void cude_sum_sse(const double* __restrict__ sse, int n_ind, int n_starts, int np1, double* __restrict__ sums) {
    const int warps_per_block = blockDim.x >> 5;
    const long long s = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= n_starts) return;
    double v = 0.0;
    for (int i = lane; i < n_ind; i += 32) v += sse[(size_t)s * n_ind + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sums[(size_t)s * np1] = v;
}

// elementary-function probe (tests): 0 tanh, 1 softplus, 2 d softplus from d (1 - 1/d, the c-peptide adjoint),
// 3 exp (clamped to +-40), 4 log, 5 rcp, 6 sigmoid (the suppression adjoint)
__device__ __forceinline__ double cude_math_probe_eval(int which, double v, const double* tab) {
    double r, d;
    int nm = 0;
    switch (which) {
        case 0: r = t_tanh(v, tab, nm); r = t_nan_inject(r, nm); break;
        case 1: t_softplus_d(v, tab, r, d); break;
        case 2: t_softplus_d(v, tab, r, d); r = fma(-1.0, m_rcp(d), 1.0); break;
        case 3: r = t_exp_sat(v, tab); break;
        case 4: r = m_log_pos(v); break;
        case 5: r = m_rcp(v); break;
        default: r = t_sigmoid(v, tab); break;
    }
    return r;
}
__global__ void cude_math_probe_kernel(int which, int n, const double* __restrict__ x, double* __restrict__ y) {
    __shared__ double sTab[256];
    for (int p = threadIdx.x; p < 256; p += blockDim.x) sTab[p] = EXP_TAB256[p];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = cude_math_probe_eval(which, x[i], sTab);
}

// Adam update on device-resident parameters (Optimisers.Adam: m, v moments, bias correction through the running
// powers b1t = beta1^t, b2t = beta2^t): x -= lr * (m/(1-b1t)) / (sqrt(v/(1-b2t)) + eps).  `scale` multiplies the raw
// gradient (e.g. 1/N for sums of per-individual gradients).  With `rows` > 0 the gradient of row r is skipped when
// ok[r] == 0 (a start whose loss is Inf keeps its parameters, like the host optimiser).
__global__ void cude_adam_kernel(long long n, double* __restrict__ x, const double* __restrict__ g, double* __restrict__ m,
                                 double* __restrict__ v, double lr, double beta1, double beta2, double eps, double b1t, double b2t,
                                 double scale, const double* __restrict__ row_flag, long long row_len, long long flag_stride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (row_flag) {
        const double f = row_flag[(i / row_len) * flag_stride];
        if (!(f - f == 0.0)) return;           // Inf / NaN loss: leave this start untouched
    }
    const double gi = g[i] * scale;
    const double mi = fma(beta1, m[i], (1.0 - beta1) * gi);
    const double vi = fma(beta2, v[i], (1.0 - beta2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    x[i] -= lr * (mi / (1.0 - b1t)) / (sqrt(vi / (1.0 - b2t)) + eps);
}

// FP64 FMA peak micro-benchmark: 8 independent DFMA chains per thread.
__global__ void cude_dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// Same, but every DFMA takes three distinct *register* operands (the shape of real code: the chains above read
// two of their operands from the uniform/constant path).  Diagnostic for the roofline discussion.
__global__ void cude_dfma_peak_rrr_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    double y0 = a + 1e-9 * threadIdx.x, y1 = y0 + 1e-10, y2 = y0 + 2e-10, y3 = y0 + 3e-10;
    double z0 = b + 1e-12 * threadIdx.x, z1 = z0 + 1e-13, z2 = z0 + 2e-13, z3 = z0 + 3e-13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            x0 = fma(x0, y0, z0); x1 = fma(x1, y1, z1); x2 = fma(x2, y2, z2); x3 = fma(x3, y3, z3);
            x4 = fma(x4, y0, z1); x5 = fma(x5, y1, z2); x6 = fma(x6, y2, z3); x7 = fma(x7, y3, z0);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// FP32 FMA peak (8 independent FFMA chains per thread) and MUFU peak (8 independent ex2.approx chains): the denominators of
// the optional FP32-network modes (cude_opts.precision = 1, 2).
__global__ void cude_ffma_peak_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
__global__ void cude_mufu_peak_kernel(float* out, int iters) {
    float x0 = 1e-3f * threadIdx.x, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f, x4 = x0 + 0.4f, x5 = x0 + 0.5f, x6 = x0 + 0.6f, x7 = x0 + 0.7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {      // ex2 of a value in [0, 1] stays in [1, 2]: subtract 1 on the FMA pipe to keep the chain bounded
            x0 = f_ex2(x0) - 1.0f; x1 = f_ex2(x1) - 1.0f; x2 = f_ex2(x2) - 1.0f; x3 = f_ex2(x3) - 1.0f;
            x4 = f_ex2(x4) - 1.0f; x5 = f_ex2(x5) - 1.0f; x6 = f_ex2(x6) - 1.0f; x7 = f_ex2(x7) - 1.0f;
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace cude
