// =====================================================================================
// cude_sup_kernel.cuh — second kernel variant: the suppression example's state-dependent cUDE.
//
// Reference: suppression/src/suppression_model.jl
//   ude_lsup! :88-95    u_hat = network([u1;u2;u3; exp(theta_i)], neural)[1]
//                       du1 = -p1 u1;  du2 = p1 u1 - u_hat;  du3 = u_hat - p3 u3        (p_true = [0.4, 0.9, 0.3])
//   suppression_loss :117-130   per individual: solve(prob, Tsit5(), saveat=timepoints), u0 = data[:,1,i];
//                       sum(abs2, (sims - data) ./ scale)/N + lambda*sum(abs2, neural)
// One thread = one trajectory (individual, start).  Unlike the c-peptide kernel the network sits inside the state
// feedback, so the discrete adjoint is the genuinely nonlinear one: per accepted step the ring keeps (t, dt, u[3]);
// the backward sweep replays the step's stages (keeping only the stage inputs g_i), then walks the stages in
// reverse, re-evaluating the network (forward + backward) at each g_i with the scalar seed kb_i[3] - kb_i[2].
// Stages, stage inputs and stage adjoints live in shared memory so that the stage loops stay rolled (one inlined
// network site per phase); the 64 compressed gradient accumulators live in REGISTERS during the adjoint sweep (round 2:
// in shared memory every network evaluation re-read and re-wrote all of them — 128 shared accesses per evaluation —
// and their 64 rows held the kernel at one 128-thread block per SM; now 87 rows = 2 blocks per SM) and are parked in
// the dead stage rows only for the final reduction.  Solves with more accepted steps than the step ring holds replay
// the forward pass in chunks, like the c-peptide kernel.
// =====================================================================================
#pragma once
#include "cude_kernels.cuh"
#include "cude_split.cuh"     // cp_async helpers

namespace cude {

struct SupArgs {
    int n_ind, n_obs, n_starts, nchunks;
    int spb;                    // > 0: small population (n_ind <= block): FLAT indexing — the block's B threads are B consecutive
                                // trajectories of the [S x N] batch (all lanes busy; round 1 packed floor(B/N) whole starts per
                                // block: 111 of 128 lanes at N = 37); spb = the most starts a block can touch, (B-1)/N + 2.  A start
                                // lies in at most two blocks: partial rows [start][2][P+1], zero-filled by the host
    const double* obs_t;        // [M] common time grid
    const double* data;         // [M][3][N]: data[(k*3 + j)*N + i]
    double p1, p3;
    double iscale[3];           // 1/scale_j
    double t0, tend;
    const double* neural;       // start s: neural + s*neural_stride
    long long neural_stride;
    const double* theta;        // [N x S]
    double abstol, reltol;
    int maxiters;
    double theta_scale;         // g_theta = theta_scale * d sse / d theta
    double* sse_out;            // [N x S] or nullptr
    double* partials;           // [blocks * warps][P+1]
    double* g_theta;            // [N x S] or nullptr
    unsigned long long* counters;
    double* ring;               // GRAD: step records of this launch's blocks, [gridDim.x][SUP_REC_CAP][SUP_REC_ROWS][B] (coalesced rows)
    int blk0;                   // this launch covers blocks blk0 .. blk0 + gridDim.x - 1 of the batch (the host cuts a batch into
                                // launches whose rings fit its scratch budget)
    // two-kernel form of the gradient (MODE 2 = forward solve leaving records, MODE 3 = adjoint sweep over them):
    double* res_g;              // [gridDim.x][3 M][B] weighted residuals of the launch's blocks
    int* sp_nacc;               // [N x S] accepted steps (0: failed)
    double* sp_sse;             // [N x S] sse (Inf: failed)
    int* ovf_count;             // trajectories with more than SUP_REC_CAP accepted steps: the host redoes the call with the fused kernel
};

template <int DEPTH_, int WIDTH_>
struct SupNet {
    static constexpr int NIN = 4, DEPTH = DEPTH_, W = WIDTH_;
    static constexpr int L1 = W * (NIN + 1);
    static constexpr int LH = W * (W + 1);
    static constexpr int OFF_OUT = L1 + (DEPTH - 1) * LH;
    static constexpr int P = OFF_OUT + W + 1;
    // compressed accumulators: dW1[:,0:3] (3W), sum dz1 (W), hidden layers, output layer
    static constexpr int NACC = 4 * W + (DEPTH - 1) * LH + W + 1;
};

// Tsit5 coefficient rows for the rolled stage loops: AROW[i][j] = a_{i+2,j+1} (i = 0..4), AROW[5] = b (stage 7 / update)
__constant__ double SUP_A[6][6] = {
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
__constant__ double SUP_E[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                                0.5823571654525552, -0.45808210592918697, 0.015151515151515152};

// network forward at state (u0,u1,u2); c[] = first-layer constant part (theta column + bias)
// One copy of the forward network per kernel instead of one per call site (5 in the gradient instantiation): 7056 -> 4904
// SASS instructions and +9 % (ncu round 2: `no_instruction` was the second largest stall, 1.16 per issue).
#ifndef CUDE_SUP_FWD_INLINE
#define CUDE_SUP_FWD_INLINE __noinline__
#endif
template <class SN>
__device__ CUDE_SUP_FWD_INLINE double sup_nn_forward(const double* __restrict__ sW, const double* __restrict__ tab,
                                                 const double (&c)[SN::W], double u0, double u1, double u2) {
    constexpr int W = SN::W;
    int nanmax = 0;
    double a[W], b[W];
#pragma unroll
    for (int j = 0; j < W; ++j) a[j] = t_tanh(fma(sW[2 * W + j], u2, fma(sW[W + j], u1, fma(sW[j], u0, c[j]))), tab, nanmax);
    int off = SN::L1;
#pragma unroll
    for (int l = 1; l < SN::DEPTH; ++l) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
            double z = sW[off + W * W + j];
#pragma unroll
            for (int i = 0; i < W; ++i) z = fma(sW[off + i * W + j], a[i], z);
            b[j] = t_tanh(z, tab, nanmax);
        }
#pragma unroll
        for (int j = 0; j < W; ++j) a[j] = b[j];
        off += SN::LH;
    }
    double z = sW[off + W];
#pragma unroll
    for (int i = 0; i < W; ++i) z = fma(sW[off + i], a[i], z);
    double sp, d;
    t_softplus_d(t_nan_inject(z, nanmax), tab, sp, d);
    return sp;
}

// forward + backward with scalar seed s: returns d(u_hat)/d(u) * s in du[], accumulates parameter gradients
// into acc (registers; every index below is a compile-time constant after unrolling).
template <class SN>
__device__ __forceinline__ void sup_nn_backward(const double* __restrict__ sW, const double* __restrict__ tab,
                                                const double (&c)[SN::W], double u0, double u1, double u2, double s,
                                                double (&acc)[SN::NACC], double (&du)[3]) {
    constexpr int W = SN::W, D = SN::DEPTH;
    int nanmax = 0;   // the forward pass succeeded on the same inputs
    double a[D][W];
#pragma unroll
    for (int j = 0; j < W; ++j) a[0][j] = t_tanh(fma(sW[2 * W + j], u2, fma(sW[W + j], u1, fma(sW[j], u0, c[j]))), tab, nanmax);
    int off = SN::L1;
#pragma unroll
    for (int l = 1; l < D; ++l) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
            double z = sW[off + W * W + j];
#pragma unroll
            for (int i = 0; i < W; ++i) z = fma(sW[off + i * W + j], a[l - 1][i], z);
            a[l][j] = t_tanh(z, tab, nanmax);
        }
        off += SN::LH;
    }
    double z = sW[off + W];
#pragma unroll
    for (int i = 0; i < W; ++i) z = fma(sW[off + i], a[D - 1][i], z);
    const double dz = s * t_sigmoid(z, tab);
    int aoff = 4 * W + (D - 1) * SN::LH;
    double da[W];
#pragma unroll
    for (int i = 0; i < W; ++i) { acc[aoff + i] = fma(dz, a[D - 1][i], acc[aoff + i]); da[i] = dz * sW[off + i]; }
    acc[aoff + W] += dz;
#pragma unroll
    for (int l = D - 1; l >= 1; --l) {
        off -= SN::LH;
        aoff -= SN::LH;
        double dzl[W], dprev[W];
#pragma unroll
        for (int j = 0; j < W; ++j) dzl[j] = da[j] * fma(-a[l][j], a[l][j], 1.0);
#pragma unroll
        for (int i = 0; i < W; ++i) {
            double sum = 0.0;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                acc[aoff + i * W + j] = fma(dzl[j], a[l - 1][i], acc[aoff + i * W + j]);
                sum = fma(sW[off + i * W + j], dzl[j], sum);
            }
            dprev[i] = sum;
        }
#pragma unroll
        for (int j = 0; j < W; ++j) { acc[aoff + W * W + j] += dzl[j]; da[j] = dprev[j]; }
    }
    {   // first layer: dW1[:,0:3] and sum dz1
        double dz1[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
            dz1[j] = da[j] * fma(-a[0][j], a[0][j], 1.0);
            acc[j] = fma(dz1[j], u0, acc[j]);
            acc[W + j] = fma(dz1[j], u1, acc[W + j]);
            acc[2 * W + j] = fma(dz1[j], u2, acc[2 * W + j]);
            acc[3 * W + j] += dz1[j];
        }
        du[0] = du[1] = du[2] = 0.0;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            du[0] = fma(sW[j], dz1[j], du[0]);
            du[1] = fma(sW[W + j], dz1[j], du[1]);
            du[2] = fma(sW[2 * W + j], dz1[j], du[2]);
        }
    }
}

__host__ __device__ inline size_t sup_smem_doubles(int P, int NACC, int M, int B, bool grad, int spb = 0) {
    // (grad = the fused gradient kernel and the adjoint kernel of the two-kernel form; the forward kernels need the stage rows only)
    // exp table, weights (one copy per start of the block), per-thread rows: k[7][3] (grad: + 2 rows = record buffer 0) +
    // (grad: record buffer 1 [23] + kb[7][3] + residuals M*3); the accumulators are parked in these rows for the reduction
    // ([NACC][B], then expanded to [P+1][B])
    size_t rows = grad ? (size_t)(2 * 23 + 21 + 3 * M) : (size_t)21;     // grad: stage rows doubling as record buffer 0, buffer 1, kb, residuals
    if (grad && rows < (size_t)P + 1) rows = (size_t)P + 1;
    if (grad && rows < (size_t)NACC) rows = (size_t)NACC;
    return (size_t)256 + (size_t)(spb > 0 ? spb : 1) * ((P + 1) & ~1) + rows * B;
}

#ifndef CUDE_SUP_REC_CAP
#define CUDE_SUP_REC_CAP 64
#endif
constexpr int SUP_REC_CAP = CUDE_SUP_REC_CAP;  // ring of accepted-step records per thread, in global memory; longer solves replay
                                               // the forward pass in chunks of this many steps
constexpr int SUP_REC_ROWS = 23;               // a record: k1..k6 [18], t, dt, u[3] = 184 B
#ifndef CUDE_SUP_MIN_BLOCKS
#define CUDE_SUP_MIN_BLOCKS 2
#endif

// MODE 0: loss only.  MODE 1: loss + gradient in one kernel (forward solve, then the adjoint sweep; solves longer than the ring
// replay the forward pass).  MODE 2 + MODE 3: the same gradient as two kernels — 2 is the loss kernel that also leaves the step
// records, the weighted residuals, the step count and the sse in global memory (156 registers, 3 blocks per SM instead of the
// fused kernel's 2 x 255 with spills), 3 walks the records backwards with the fused kernel's sweep and reduction code (no
// forward-pass state alive: fewer spills).  Trajectories beyond SUP_REC_CAP steps are counted; the host then redoes the call
// with MODE 1.
constexpr int SUP_LOSS = 0, SUP_FUSED = 1, SUP_FWD_REC = 2, SUP_ADJ = 3;
#ifndef CUDE_SUP_SORT_BLOCK
#define CUDE_SUP_SORT_BLOCK 1
#endif
#ifndef CUDE_SUP_FWD_BLOCKS
#define CUDE_SUP_FWD_BLOCKS 4      // 2 / 3 / 4 blocks per SM: 7.77 / 7.44 / 7.34 ms for the gradient of 37 x 10 000
#endif
template <class SN, int MODE>
__global__ void __launch_bounds__(128, (MODE == SUP_FUSED || MODE == SUP_ADJ) ? CUDE_SUP_MIN_BLOCKS : (MODE == SUP_FWD_REC ? CUDE_SUP_FWD_BLOCKS : 1))
cude_sup_kernel(const SupArgs A) {
    constexpr bool GRAD = (MODE == SUP_FUSED || MODE == SUP_ADJ);     // adjoint sweep + gradient reduction in this kernel
    constexpr bool FWD = (MODE != SUP_ADJ);                           // forward solve in this kernel
    constexpr bool REC = (MODE == SUP_FUSED || MODE == SUP_FWD_REC);  // the forward solve leaves step records
    using namespace tab;
    constexpr int W = SN::W, P = SN::P;
    extern __shared__ double smem[];
    const int B = blockDim.x, tid = threadIdx.x;
    const int N = A.n_ind, M = A.n_obs;
    constexpr int PP = (P + 1) & ~1;
    const int spb = A.spb;
    double* sTab = smem;
    double* sWall = sTab + 256;                       // [max(spb,1)][PP] weights of the block's start(s)
    double* sK = sWall + (size_t)(spb > 0 ? spb : 1) * PP;   // [7][3][B] stages; GRAD: 2 more rows = record buffer 0 of the adjoint sweep
    double* sKb = sK + (size_t)(GRAD ? SUP_REC_ROWS : 21) * B;   // [7][3][B] stage adjoints (GRAD)
    double* sRes = sKb + (GRAD ? (size_t)21 * B : 0); // [M][3][B] weighted residuals (GRAD)
    double* sRec1 = sRes + (GRAD ? (size_t)3 * M * B : 0);       // [23][B] record buffer 1 (GRAD)

    int s, i, sloc = 0;
    bool active;
    const unsigned bid = blockIdx.x + (unsigned)A.blk0;
    const long long j0 = (long long)bid * B, ntot = (long long)N * A.n_starts;   // flat mode: first trajectory of the block
    int s_first = 0, nsl = 1;
    // vt: which of the block's B trajectories this thread works on.  The adjoint kernel of the two-kernel form sorts the
    // block's trajectories by their number of accepted steps (known from the forward kernel), so that the lanes of a warp
    // sweep equally many steps (29.4 of 32 lanes were active in natural order); everything that identifies the trajectory
    // — index, start, weights, record and residual columns, the column its gradient is parked in for the block
    // reduction — follows vt, the per-thread scratch rows stay with tid.  Results do not depend on the permutation.
    int vt = tid;
    if constexpr (MODE == SUP_ADJ && CUDE_SUP_SORT_BLOCK) {
        __shared__ int sBin[SUP_REC_CAP + 2], sPerm[128];
        for (int k = tid; k < SUP_REC_CAP + 2; k += B) sBin[k] = 0;
        __syncthreads();
        long long jn;                                  // this thread's trajectory in natural order
        bool an;
        if (spb > 0) { jn = j0 + tid; an = jn < ntot; }
        else { const int sn = bid / A.nchunks, in_ = (bid - sn * A.nchunks) * B + tid; an = in_ < N; jn = (long long)sn * N + in_; }
        int key = 0;
        if (an) { const int na = A.sp_nacc[jn]; key = na > SUP_REC_CAP ? SUP_REC_CAP + 1 : na; }
        const int rank = atomicAdd(&sBin[key], 1);
        __syncthreads();
        if (tid == 0) {                                // longest sweeps first
            int o = 0;
            for (int k = SUP_REC_CAP + 1; k >= 0; --k) { const int c = sBin[k]; sBin[k] = o; o += c; }
        }
        __syncthreads();
        sPerm[sBin[key] + rank] = tid;
        __syncthreads();
        vt = sPerm[tid];
    }
    if (spb > 0) {                                    // flat: B consecutive trajectories, up to spb starts per block
        const long long jj = j0 + vt;
        active = jj < ntot;
        s_first = (int)(j0 / N);
        s = active ? (int)(jj / N) : s_first;
        i = active ? (int)(jj - (long long)s * N) : 0;
        sloc = s - s_first;
        const long long jlast = (j0 + B < ntot ? j0 + B : ntot) - 1;
        nsl = (int)(jlast / N) - s_first + 1;
        for (int p = tid; p < nsl * P; p += B) {
            const int sl = p / P, pp = p - sl * P;
            sWall[sl * PP + pp] = A.neural[(long long)(s_first + sl) * A.neural_stride + pp];
        }
    } else {
        s = bid / A.nchunks;
        const int ch = bid - s * A.nchunks;
        i = ch * B + vt;
        active = i < N;
        const double* gW = A.neural + (long long)s * A.neural_stride;
        for (int p = tid; p < P; p += B) sWall[p] = gW[p];
    }
    const double* const sW = sWall + (size_t)sloc * PP;
    const long long jt = (long long)s * N + (active ? i : 0);
    for (int p = tid; p < 256; p += B) sTab[p] = EXP_TAB256[p];
    double* const myK = sK + tid;
    double* const myKb = sKb + tid;
    double* const myRes = sRes + tid;
    double* const myAcc = sK + vt;           // the accumulators' parking rows for the reduction (the stage rows, dead by then)
    double acc[GRAD ? SN::NACC : 1];         // gradient accumulators (compressed layout): registers
#pragma unroll
    for (int q = 0; q < (GRAD ? SN::NACC : 1); ++q) acc[q] = 0.0;
    __syncthreads();

    double sse = 0.0, gtheta = 0.0, etheta = 0.0;
    int nacc = 0, nrej = 0;
    bool failed = false;

    if (active) {
        const double p1 = A.p1, p3 = A.p3, abstol = A.abstol, reltol = A.reltol;
        const double t0 = A.t0, tend = A.tend;
        const double* yd = A.data + i;               // y(k, j) = yd[(k*3 + j)*N]
        etheta = m_exp(A.theta[jt]);
        double c[W];
#pragma unroll
        for (int q = 0; q < W; ++q) c[q] = fma(sW[3 * W + q], etheta, sW[4 * W + q]);
        const double dtmax = tend - t0;
        const double at0 = fabs(t0), at1 = fabs(tend);
        const double dtmin = fmax(nextafter(at0, CUDART_INF) - at0, nextafter(at1, CUDART_INF) - at1);
        const double snap = 100.0 * (nextafter(at1, CUDART_INF) - at1);

        // rhs: f(u) = [-p1 u0, p1 u0 - uhat, uhat - p3 u2]
#define SUP_RHS(U0, U1, U2, F0, F1, F2)                                   \
    {                                                                     \
        const double uh__ = sup_nn_forward<SN>(sW, sTab, c, U0, U1, U2);  \
        F0 = -p1 * (U0); F1 = fma(p1, (U0), -uh__); F2 = fma(-p3, (U2), uh__); \
    }
        // accepted-step record: the step's stage derivatives k1..k6 (so that the adjoint rebuilds the stage inputs
        // g_i = u + dt sum a_ij k_j without re-evaluating the network), then (t, dt, u).  Round 2b: the ring is an explicit
        // global array with rows [slot][row][tid] — the forward pass writes coalesced rows, the adjoint sweep fetches the next
        // record by cp.async into a shared-memory double buffer while it works on the current one (as a local-memory array the
        // ring was read at the top of every step and waited for: long_scoreboard 0.81 per issue)
        double* const ringb = (GRAD || REC) ? A.ring + (size_t)blockIdx.x * SUP_REC_CAP * SUP_REC_ROWS * B + vt : nullptr;
        double* const resg = (MODE == SUP_FWD_REC || MODE == SUP_ADJ) ? A.res_g + (size_t)blockIdx.x * 3 * M * B + vt : nullptr;

        // adjoint carry across replay chunks
        double lam0 = 0.0, lam1 = 0.0, lam2 = 0.0, t_next = tend;
        int kobs = M - 1;
        int stop_at = 0x7fffffff;      // accepted steps the (re)played forward pass runs for
        bool first_pass = true;
        do {
        if constexpr (FWD) {
        // ---------------- forward pass (first pass: the solve; later passes: replay up to stop_at accepted steps) ----------------
        double u0 = yd[0], u1 = yd[(size_t)N], u2 = yd[(size_t)2 * N];      // u0 = data[:,1,i]
        double t = t0;
        int iobs = 0, ret = 0, na = 0, nr = 0;
        double fsse = 0.0;
        while (first_pass && iobs < M && A.obs_t[iobs] <= t0) {              // save_start: residual at t0 is 0 by construction
            const double r0 = (u0 - yd[(size_t)(iobs * 3) * N]) * A.iscale[0];
            const double r1 = (u1 - yd[(size_t)(iobs * 3 + 1) * N]) * A.iscale[1];
            const double r2 = (u2 - yd[(size_t)(iobs * 3 + 2) * N]) * A.iscale[2];
            if (GRAD) { myRes[(iobs * 3) * B] = r0 * A.iscale[0]; myRes[(iobs * 3 + 1) * B] = r1 * A.iscale[1]; myRes[(iobs * 3 + 2) * B] = r2 * A.iscale[2]; }
            if constexpr (MODE == SUP_FWD_REC) { resg[(iobs * 3) * B] = r0 * A.iscale[0]; resg[(iobs * 3 + 1) * B] = r1 * A.iscale[1]; resg[(iobs * 3 + 2) * B] = r2 * A.iscale[2]; }
            fsse += m_sumsq(r0, r1, r2);
            ++iobs;
        }
        double k10, k11, k12;
        SUP_RHS(u0, u1, u2, k10, k11, k12)
        double dt;
        {   // Hairer initial step
            const double isk0 = 1.0 / fma(fabs(u0), reltol, abstol), isk1 = 1.0 / fma(fabs(u1), reltol, abstol),
                         isk2 = 1.0 / fma(fabs(u2), reltol, abstol);
            double x0 = u0 * isk0, x1 = u1 * isk1, x2 = u2 * isk2;
            const double d0 = sqrt(m_sumsq(x0, x1, x2) / 3.0);
            x0 = k10 * isk0; x1 = k11 * isk1; x2 = k12 * isk2;
            const double d1 = sqrt(m_sumsq(x0, x1, x2) / 3.0);
            double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
            dt0 = fmin(dt0, dtmax);
            double f0, f1, f2;
            SUP_RHS(fma(dt0, k10, u0), fma(dt0, k11, u1), fma(dt0, k12, u2), f0, f1, f2)
            x0 = (f0 - k10) * isk0; x1 = (f1 - k11) * isk1; x2 = (f2 - k12) * isk2;
            const double d2 = sqrt(m_sumsq(x0, x1, x2) / 3.0) / dt0;
            const double dm = fmax(d1, d2);
            const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : t_pow10(-(2.0 + m_log10(dm)) / 5.0, sTab);
            dt = fmax(dtmin, fmin(fmin(100.0 * dt0, dt1), dtmax));
            if (!(isfinite(dt) && isfinite(k10) && isfinite(k11) && isfinite(k12))) ret = 3;
        }
        myK[0] = k10; myK[B] = k11; myK[2 * B] = k12;
        double lnqold = -9.210340371976182;
        int iter = 0;
        while (ret == 0 && t < tend && na < stop_at) {
            if (++iter > A.maxiters) { ret = 1; break; }
            dt = fmin(dt, tend - t);
            if (!(dt > dtmin)) { ret = (dt != dt) ? 3 : 2; break; }
            // stages 2..7 (rolled): g = u + dt sum_j a_ij k_j ; k_i = f(g)
            double g0 = u0, g1 = u1, g2 = u2;
#pragma unroll 1
            for (int st = 0; st < 6; ++st) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                for (int q = 0; q <= st; ++q) {
                    const double a = SUP_A[st][q];
                    s0 = fma(a, myK[(q * 3) * B], s0); s1 = fma(a, myK[(q * 3 + 1) * B], s1); s2 = fma(a, myK[(q * 3 + 2) * B], s2);
                }
                g0 = fma(dt, s0, u0); g1 = fma(dt, s1, u1); g2 = fma(dt, s2, u2);
                double f0, f1, f2;
                SUP_RHS(g0, g1, g2, f0, f1, f2)
                myK[((st + 1) * 3) * B] = f0; myK[((st + 1) * 3 + 1) * B] = f1; myK[((st + 1) * 3 + 2) * B] = f2;
            }
            const double un0 = g0, un1 = g1, un2 = g2;     // the last stage input (row b) is the new state
            // error estimate
            double e0 = 0.0, e1 = 0.0, e2 = 0.0;
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                e0 = fma(SUP_E[q], myK[(q * 3) * B], e0); e1 = fma(SUP_E[q], myK[(q * 3 + 1) * B], e1); e2 = fma(SUP_E[q], myK[(q * 3 + 2) * B], e2);
            }
            e0 = dt * e0 * m_rcp(fma(fmax(fabs(u0), fabs(un0)), reltol, abstol));
            e1 = dt * e1 * m_rcp(fma(fmax(fabs(u1), fabs(un1)), reltol, abstol));
            e2 = dt * e2 * m_rcp(fma(fmax(fabs(u2), fabs(un2)), reltol, abstol));
            const double E2 = m_sumsq(e0, e1, e2) / 3.0;
            if (!(E2 == E2) || !isfinite(un0) || !isfinite(un1) || !isfinite(un2)) { ret = 3; break; }
            CUDE_TRACE_STEP(t, dt, sqrt(E2))
            const double lnE = 0.5 * m_log_pos(E2);
            if (E2 <= 1.0) {
                const double q = fmax(1.0 / qmax, fmin(1.0 / qmin, t_exp_sat(fma(beta1, lnE, -beta2 * lnqold), sTab) * (1.0 / gamma)));
                double tnew = t + dt;
                if (fabs(tnew - tend) < snap) tnew = tend;
                while (first_pass && iobs < M && A.obs_t[iobs] <= tnew) {
                    const double ts = A.obs_t[iobs];
                    double y0, y1, y2;
                    if (ts == tnew) { y0 = un0; y1 = un1; y2 = un2; }
                    else {
                        double bw[7];
                        dense_weights((ts - t) * m_rcp(dt), bw);
                        double d0 = 0.0, d1 = 0.0, d2 = 0.0;
#pragma unroll
                        for (int q = 0; q < 7; ++q) { d0 = fma(bw[q], myK[(q * 3) * B], d0); d1 = fma(bw[q], myK[(q * 3 + 1) * B], d1); d2 = fma(bw[q], myK[(q * 3 + 2) * B], d2); }
                        y0 = fma(dt, d0, u0); y1 = fma(dt, d1, u1); y2 = fma(dt, d2, u2);
                    }
                    const double r0 = (y0 - yd[(size_t)(iobs * 3) * N]) * A.iscale[0];
                    const double r1 = (y1 - yd[(size_t)(iobs * 3 + 1) * N]) * A.iscale[1];
                    const double r2 = (y2 - yd[(size_t)(iobs * 3 + 2) * N]) * A.iscale[2];
                    if (GRAD) { myRes[(iobs * 3) * B] = r0 * A.iscale[0]; myRes[(iobs * 3 + 1) * B] = r1 * A.iscale[1]; myRes[(iobs * 3 + 2) * B] = r2 * A.iscale[2]; }
                    if constexpr (MODE == SUP_FWD_REC) { resg[(iobs * 3) * B] = r0 * A.iscale[0]; resg[(iobs * 3 + 1) * B] = r1 * A.iscale[1]; resg[(iobs * 3 + 2) * B] = r2 * A.iscale[2]; }
            if constexpr (MODE == SUP_FWD_REC) { resg[(iobs * 3) * B] = r0 * A.iscale[0]; resg[(iobs * 3 + 1) * B] = r1 * A.iscale[1]; resg[(iobs * 3 + 2) * B] = r2 * A.iscale[2]; }
                    fsse += m_sumsq(r0, r1, r2);
                    ++iobs;
                }
                if (REC && (MODE == SUP_FUSED || na < SUP_REC_CAP)) {
                    double* const r = ringb + (size_t)(na % SUP_REC_CAP) * SUP_REC_ROWS * B;
#pragma unroll
                    for (int q = 0; q < 18; ++q) r[q * B] = myK[q * B];
                    r[18 * B] = t; r[19 * B] = dt; r[20 * B] = u0; r[21 * B] = u1; r[22 * B] = u2;
                }
                ++na;
                lnqold = fmax(lnE, -9.210340371976182);
                dt = fmin(dt * m_rcp(q), dtmax);
                t = tnew; u0 = un0; u1 = un1; u2 = un2;
                myK[0] = myK[18 * B]; myK[B] = myK[19 * B]; myK[2 * B] = myK[20 * B];      // FSAL: k1 <- k7
            } else {
                ++nr;
                dt = dt * m_rcp(fmin(1.0 / qmin, t_exp_sat(beta1 * lnE, sTab) * (1.0 / gamma)));
            }
        }
        if (first_pass) {
            if (ret == 0 && iobs < M) ret = 3;
            failed = (ret != 0);
            sse = failed ? CUDART_INF : fsse;
            nacc = na; nrej = nr;
            stop_at = na;
        }
        } else {
            // adjoint kernel of the two-kernel form: what the forward kernel left behind
            sse = A.sp_sse[jt];
            nacc = A.sp_nacc[jt];
            failed = !(sse - sse == 0.0);
            if (!failed && nacc > SUP_REC_CAP) { atomicAdd(A.ovf_count, 1); break; }     // the host redoes the call with the fused kernel
            stop_at = nacc;
            if (!failed)
                for (int k = 0; k < 3 * M; ++k) myRes[k * B] = resg[k * B];
        }
        first_pass = false;
        if (!GRAD || failed) break;

        if constexpr (GRAD) {
            // ---------------- discrete adjoint over the steps [lo, stop_at) held in the ring ----------------
            const int lo = (stop_at > SUP_REC_CAP) ? stop_at - SUP_REC_CAP : 0;
            auto fetch = [&](int b, int n) {            // record n -> buffer b (this thread's column; own copies only: no barrier)
                const double* const g = ringb + (size_t)(n % SUP_REC_CAP) * SUP_REC_ROWS * B;
                double* const d = (b ? sRec1 : sK) + tid;
#pragma unroll
                for (int q = 0; q < SUP_REC_ROWS; ++q) cp_async8(d + (size_t)q * B, g + (size_t)q * B);
            };
            int cur = 0;
            fetch(0, stop_at - 1);
            cp_async_commit();
            for (int n = stop_at - 1; n >= lo; --n, cur ^= 1) {
                if (n > lo) fetch(cur ^ 1, n - 1);
                cp_async_commit();
                cp_async_wait<1>();                      // record n has landed
                const double* const rk = (cur ? sRec1 : sK) + tid;     // rows 0..17: k1..k6
                const double tn = rk[18 * B], h = rk[19 * B], ru0 = rk[20 * B], ru1 = rk[21 * B], ru2 = rk[22 * B];
                // stage adjoints
#pragma unroll 1
                for (int q = 0; q < 21; ++q) myKb[q * B] = 0.0;
                double ub0 = 0.0, ub1 = 0.0, ub2 = 0.0;
                bool kb7 = false;
                while (kobs >= 0) {
                    const double ts = A.obs_t[kobs];
                    if (!(ts > tn)) break;
                    const double w0 = 2.0 * myRes[(kobs * 3) * B], w1 = 2.0 * myRes[(kobs * 3 + 1) * B], w2 = 2.0 * myRes[(kobs * 3 + 2) * B];
                    if (ts == t_next) { lam0 += w0; lam1 += w1; lam2 += w2; }
                    else {
                        double bw[7];
                        dense_weights((ts - tn) * m_rcp(h), bw);
                        ub0 += w0; ub1 += w1; ub2 += w2;
#pragma unroll
                        for (int q = 0; q < 7; ++q) {
                            const double hb = h * bw[q];
                            myKb[(q * 3) * B] = fma(w0, hb, myKb[(q * 3) * B]);
                            myKb[(q * 3 + 1) * B] = fma(w1, hb, myKb[(q * 3 + 1) * B]);
                            myKb[(q * 3 + 2) * B] = fma(w2, hb, myKb[(q * 3 + 2) * B]);
                        }
                        kb7 = true;
                    }
                    --kobs;
                }
                // stage 7 (k7 = f(u_{n+1}), dense output only), then u_{n+1} = u_n + h sum b_j k_j, then stages 6..1
#pragma unroll 1
                for (int st = 6; st >= 0; --st) {
                    if (st == 6 && !kb7) {
                        // no interior observation: k7 carries no adjoint
                    } else {
                        const double v0 = myKb[(st * 3) * B], v1 = myKb[(st * 3 + 1) * B], v2 = myKb[(st * 3 + 2) * B];
                        // the stage input g_{st+1} = u_n + h sum_{q<st} a_{st+1,q+1} k_{q+1} from the record (g_1 = u_n)
                        double g0 = ru0, g1 = ru1, g2 = ru2;
                        if (st > 0) {
                            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                            for (int q = 0; q < st; ++q) {
                                const double a = SUP_A[st - 1][q];
                                s0 = fma(a, rk[(q * 3) * B], s0); s1 = fma(a, rk[(q * 3 + 1) * B], s1); s2 = fma(a, rk[(q * 3 + 2) * B], s2);
                            }
                            g0 = fma(h, s0, ru0); g1 = fma(h, s1, ru1); g2 = fma(h, s2, ru2);
                        }
                        double du[3];
                        sup_nn_backward<SN>(sW, sTab, c, g0, g1, g2, v2 - v1, acc, du);
                        // gb = J^T v = [-p1 v0 + p1 v1, 0, -p3 v2] + grad_u(u_hat) (v2 - v1)
                        const double gb0 = fma(p1, v1 - v0, du[0]), gb1 = du[1], gb2 = fma(-p3, v2, du[2]);
                        if (st == 6) { lam0 += gb0; lam1 += gb1; lam2 += gb2; }
                        else {
                            ub0 += gb0; ub1 += gb1; ub2 += gb2;
                            for (int q = 0; q < st; ++q) {          // g_{st+1} = u + h sum_{q<st} a_{st+1,q+1} k_{q+1}
                                const double ha = h * SUP_A[st - 1][q];
                                myKb[(q * 3) * B] = fma(ha, gb0, myKb[(q * 3) * B]);
                                myKb[(q * 3 + 1) * B] = fma(ha, gb1, myKb[(q * 3 + 1) * B]);
                                myKb[(q * 3 + 2) * B] = fma(ha, gb2, myKb[(q * 3 + 2) * B]);
                            }
                        }
                    }
                    if (st == 6) {
                        // u_{n+1} = u_n + h sum_{j<=6} b_j k_j
                        ub0 += lam0; ub1 += lam1; ub2 += lam2;
                        for (int q = 0; q < 6; ++q) {
                            const double hb = h * SUP_A[5][q];
                            myKb[(q * 3) * B] = fma(hb, lam0, myKb[(q * 3) * B]);
                            myKb[(q * 3 + 1) * B] = fma(hb, lam1, myKb[(q * 3 + 1) * B]);
                            myKb[(q * 3 + 2) * B] = fma(hb, lam2, myKb[(q * 3 + 2) * B]);
                        }
                    }
                }
                lam0 = ub0; lam1 = ub1; lam2 = ub2;
                t_next = tn;
            }
            cp_async_wait<0>();
            stop_at = lo;          // steps below lo still to do: replay the forward pass up to lo
        }
        } while (stop_at > 0);
        if constexpr (GRAD) if (!failed) {
            // d sse / d theta = (sum_j dz1_j W1[j,3]) * exp(theta)
            double db = 0.0;
#pragma unroll
            for (int q = 0; q < W; ++q) db = fma(acc[3 * W + q], sW[3 * W + q], db);
            gtheta = db * etheta;
        }
#undef SUP_RHS
    }
    if constexpr (MODE == SUP_ADJ && CUDE_SUP_SORT_BLOCK) __syncthreads();   // column vt is another thread's scratch until it has finished
    if constexpr (GRAD) {      // park the accumulators in the (dead) stage rows for the reduction below
#pragma unroll
        for (int q = 0; q < SN::NACC; ++q) myAcc[q * B] = acc[q];
    }

    if (active) {
        if (A.sse_out) A.sse_out[jt] = sse;
        if (GRAD && A.g_theta) A.g_theta[jt] = failed ? 0.0 : gtheta * A.theta_scale;
        if constexpr (MODE == SUP_FWD_REC) { A.sp_nacc[jt] = failed ? 0 : nacc; A.sp_sse[jt] = sse; }
    }
    const int lane = tid & 31, wid = tid >> 5, nw = (B + 31) >> 5;
    constexpr int nred = GRAD ? P + 1 : 1;
    if (A.partials && spb > 0) {
        // flat blocks: a warp may hold lanes of two starts, so the rows go through shared memory: every thread writes
        // its expanded values to [q][tid] (its own column of the — now dead — stage rows), then thread (start, q) sums
        // the start's columns of this block in index order (deterministic) into the start's partial row for this block
        double vals_first = active ? sse : 0.0;
        double* const myRow = sK + vt;
        if constexpr (GRAD) {
            double ex[P];
#pragma unroll 1
            for (int p = 0; p < P; ++p) {
                double v = 0.0;
                if (active && !failed) {
                    if (p < 3 * W) v = myAcc[p * B];
                    else if (p < 4 * W) v = myAcc[(3 * W + (p - 3 * W)) * B] * etheta;
                    else if (p < SN::L1) v = myAcc[(3 * W + (p - 4 * W)) * B];
                    else v = myAcc[(4 * W + (p - SN::L1)) * B];
                }
                ex[p] = v;
            }
            // (the expanded values are collected first: rows 1..P overlap the stage-input / adjoint rows, not the accumulators,
            //  but keeping read and write phases apart makes that irrelevant)
#pragma unroll 1
            for (int p = 0; p < P; ++p) myRow[(1 + p) * B] = ex[p];
        }
        myRow[0] = vals_first;
        __syncthreads();
        for (int idx = tid; idx < nsl * nred; idx += B) {
            const int sl = idx / nred, q = idx - sl * nred, ss = s_first + sl;
            long long lo = (long long)ss * N - j0, hi = (long long)(ss + 1) * N - j0;      // the start's columns in this block
            if (lo < 0) lo = 0;
            if (hi > B) hi = B;
            if (hi > ntot - j0) hi = ntot - j0;
            const double* src = sK + (size_t)q * B;
            double v = 0.0;
            for (long long k = lo; k < hi; ++k) v += src[k];
            const long long which = (long long)bid - ((long long)ss * N) / B;        // 0: the start's first block, 1: its second
            A.partials[((size_t)ss * 2 + (size_t)which) * (P + 1) + q] = v;
        }
    } else if (A.partials) {
        double* const row = A.partials + ((size_t)bid * nw + wid) * (P + 1);
#pragma unroll 1
        for (int q = 0; q < nred; ++q) {
            double v = 0.0;
            if (active) {
                if (q == 0) v = sse;
                else if constexpr (GRAD) { if (!failed) {
                    const int p = q - 1;      // expand to the SimpleChains layout: W1 is [W x 4] column-major, then b1
                    if (p < 3 * W) v = myAcc[p * B];                                   // W1[:,0:3]
                    else if (p < 4 * W) v = myAcc[(3 * W + (p - 3 * W)) * B] * etheta; // W1[:,3]  (exp(theta) column)
                    else if (p < SN::L1) v = myAcc[(3 * W + (p - 4 * W)) * B];         // b1
                    else v = myAcc[(4 * W + (p - SN::L1)) * B];                        // hidden + output layers
                } }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) row[q] = v;
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

}  // namespace cude
