// =====================================================================================
// cude_train.cuh — device-resident lock-step optimisers for the multi-start cUDE training
// (`_optimize`, reference src/parameter-estimation.jl:170-183: Optimisers.Adam(lr) for `number_of_iterations_adam`
// iterations, then Optim.LBFGS(linesearch = BackTracking()) for `number_of_iterations_lbfgs`; called for every selected
// start, :374-376).  The reference runs one optimiser at a time, each objective call re-solving every ODE on the CPU;
// here all S starts advance together and nothing but a 16-byte status ever returns to the host:
//     loss+gradient kernel (cude_eval_dev) -> reduction -> one `step` kernel, one block per start.
// Parameters of a start: x = [neural (P); cond (N)], D = P + N.
//
// Adam: Optimisers.Adam, beta = (0.9, 0.999), eps = 1e-8; the best iterate is kept (what Optimization.jl's Optimisers
// wrapper returns).  L-BFGS: m = 10 two-loop recursion, initial scaling gamma = s'y / y'y, BackTracking(order = 3):
// c1 = 1e-4, rho in [0.1, 0.5], quadratic interpolation on the first backtrack and cubic afterwards, alpha0 = 1;
// convergence on |g|_inf <= g_tol (1e-8).  The line search runs as a state machine: every launch sequence evaluates loss
// AND gradient at the trial point; an accepted trial therefore needs no second evaluation (Optim evaluates f during the
// search and the gradient at the accepted point), a rejected one wastes its gradient.  The accepted iterates are the same.
// =====================================================================================
#pragma once
#ifndef CUDE_HOST_EMU
#include <cuda_runtime.h>
#endif

namespace cude {

struct TrainArgs {
    int S, P, N, D, m;                 // starts, network parameters, individuals, D = P + N, history length
    // evaluation interface
    double* xt_n;                      // [S][P]  trial point, network part  (input of the next evaluation)
    double* xt_c;                      // [S][N]  trial point, conditional part
    const double* sums;                // [S][P+1] {sum sse, d sum sse / d neural} of the evaluation at xt
    const double* g_cond;              // [S][N]  d loss / d cond (already scaled by 1/N)
    double scale;                      // 1/N: loss = scale * sums[0], d loss/d neural = scale * sums[1..]
    // state
    double *x, *g, *d, *best_x;        // [S][D]
    double *am, *av;                   // [S][D] Adam moments
    double *hs, *hy;                   // [m][S][D] L-BFGS history
    double *rho;                       // [m][S]
    double *sc;                        // [S][SC_N] per-start scalars
    int* ic;                           // [S][IC_N] per-start integers
    // options
    double lr, b1, b2, eps, b1t, b2t;  // Adam (b1t, b2t: running powers for this iteration)
    double g_tol, c1, rho_hi, rho_lo;
    int ls_maxiter, maxiters;
    int* status_count;                 // [2] {starts still active, total accepted iterations}: zeroed by the host before a step
};
enum { SC_FX = 0, SC_FPREV, SC_ALPHA, SC_APREV, SC_DPHI0, SC_BESTF, SC_N };
enum { IC_NH = 0, IC_HEAD, IC_ITERS, IC_LS, IC_STATUS, IC_N };       // status: 0 active, 1 converged, 2 line search failed, 3 maxiters
#ifndef CUDE_TRAIN_T
#define CUDE_TRAIN_T 128
#endif
constexpr int TRAIN_T = CUDE_TRAIN_T;      // threads per start (the host emulation runs 64: two warps still exercise the cross-warp sums)

__device__ __forceinline__ double train_block_sum(double v, double* sh) {
    // sum over the block, result in every thread (deterministic: fixed tree)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    return t;
}
__device__ __forceinline__ double train_block_max(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double t = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = fmax(t, sh[w]);
    return t;
}

// value and gradient of the evaluation at the trial point, element k of start s
__device__ __forceinline__ double train_gt(const TrainArgs& A, int s, int k) {
    return k < A.P ? A.scale * A.sums[(size_t)s * (A.P + 1) + 1 + k] : A.g_cond[(size_t)s * A.N + (k - A.P)];
}
__device__ __forceinline__ void train_set_xt(const TrainArgs& A, int s, int k, double v) {
    if (k < A.P) A.xt_n[(size_t)s * A.P + k] = v; else A.xt_c[(size_t)s * A.N + (k - A.P)] = v;
}
__device__ __forceinline__ double train_get_xt(const TrainArgs& A, int s, int k) {
    return k < A.P ? A.xt_n[(size_t)s * A.P + k] : A.xt_c[(size_t)s * A.N + (k - A.P)];
}

// ---- Adam: one iteration for every start.  phase 0: regular step; phase 1: after the last step — compare the final
//      point with the best iterate and leave the best one in x / xt (the L-BFGS start) ----
__global__ void __launch_bounds__(TRAIN_T) cude_adam_step_kernel(const TrainArgs A, int phase) {
    __shared__ double sh[TRAIN_T / 32];
    const int s = blockIdx.x, tid = threadIdx.x;
    double* const x = A.x + (size_t)s * A.D;
    double* const bx = A.best_x + (size_t)s * A.D;
    double* const sc = A.sc + (size_t)s * SC_N;
    const double f = A.scale * A.sums[(size_t)s * (A.P + 1)];
    double bad = (f - f == 0.0) ? 0.0 : 1.0;
    for (int k = tid; k < A.D; k += TRAIN_T) { const double gk = train_gt(A, s, k); if (!(gk - gk == 0.0)) bad = 1.0; }
    bad = train_block_max(bad, sh);
    const bool better = f < sc[SC_BESTF];
    __syncthreads();
    if (better) for (int k = tid; k < A.D; k += TRAIN_T) bx[k] = x[k];
    if (tid == 0 && better) sc[SC_BESTF] = f;
    if (phase == 0) {
        if (bad == 0.0) {
            for (int k = tid; k < A.D; k += TRAIN_T) {
                const double gk = train_gt(A, s, k);
                const double mk = A.b1 * A.am[(size_t)s * A.D + k] + (1.0 - A.b1) * gk;
                const double vk = A.b2 * A.av[(size_t)s * A.D + k] + (1.0 - A.b2) * gk * gk;
                A.am[(size_t)s * A.D + k] = mk; A.av[(size_t)s * A.D + k] = vk;
                const double xn = x[k] - A.lr * (mk / (1.0 - A.b1t)) / (sqrt(vk / (1.0 - A.b2t)) + A.eps);
                x[k] = xn;
                train_set_xt(A, s, k, xn);
            }
        } else {
            // a failed evaluation (Inf loss): zero gradient, the moments decay (what the host optimiser does with g = 0)
            for (int k = tid; k < A.D; k += TRAIN_T) {
                A.am[(size_t)s * A.D + k] *= A.b1; A.av[(size_t)s * A.D + k] *= A.b2;
            }
        }
    } else {
        __syncthreads();
        for (int k = tid; k < A.D; k += TRAIN_T) { const double v = bx[k]; x[k] = v; train_set_xt(A, s, k, v); }
    }
}

// ---- L-BFGS + BackTracking state machine: one micro-step for every start after an evaluation at the trial point ----
// first != 0: the evaluation was at the starting point x itself (initialise fx, g and the first direction).
__global__ void __launch_bounds__(TRAIN_T) cude_lbfgs_step_kernel(const TrainArgs A, int first) {
    __shared__ double sh[TRAIN_T / 32];
    __shared__ double s_alpha[16];
    const int s = blockIdx.x, tid = threadIdx.x, D = A.D, S = A.S;
    double* const x = A.x + (size_t)s * D;
    double* const g = A.g + (size_t)s * D;
    double* const d = A.d + (size_t)s * D;
    double* const sc = A.sc + (size_t)s * SC_N;
    int* const ic = A.ic + (size_t)s * IC_N;
    int status = ic[IC_STATUS];
    if (status != 0) {
        for (int k = tid; k < D; k += TRAIN_T) train_set_xt(A, s, k, x[k]);      // idle starts are re-evaluated at their solution
        return;
    }
    const double ft = A.scale * A.sums[(size_t)s * (A.P + 1)];
    double badg = 0.0;
    for (int k = tid; k < D; k += TRAIN_T) { const double gk = train_gt(A, s, k); if (!(gk - gk == 0.0)) badg = 1.0; }
    badg = train_block_max(badg, sh);
    const bool fin = (ft - ft == 0.0);
    double fx = sc[SC_FX], alpha = sc[SC_ALPHA], dphi0 = sc[SC_DPHI0];
    int nh = ic[IC_NH], head = ic[IC_HEAD], ls = ic[IC_LS], iters = ic[IC_ITERS];
    bool accept;
    if (first) {
        accept = true;
        if (!fin || badg != 0.0) status = 2;      // cannot start from a failed point
    } else {
        accept = fin && badg == 0.0 && ft <= fx + A.c1 * alpha * dphi0;
    }
    __syncthreads();
    if (status == 0 && accept) {
        double ys = 0.0, dxmax = 0.0;
        if (!first) {
            // history update: s = xt - x, y = gt - g
            for (int k = tid; k < D; k += TRAIN_T) {
                const double sv = train_get_xt(A, s, k) - x[k], yv = train_gt(A, s, k) - g[k];
                A.hs[((size_t)head * S + s) * D + k] = sv; A.hy[((size_t)head * S + s) * D + k] = yv;
                ys += sv * yv; dxmax = fmax(dxmax, fabs(sv));
            }
            ys = train_block_sum(ys, sh);
            dxmax = train_block_max(dxmax, sh);
            if (ys > 1e-300) {
                if (tid == 0) A.rho[(size_t)head * S + s] = 1.0 / ys;
                head = (head + 1) % A.m;
                nh = nh + 1 < A.m ? nh + 1 : A.m;
            } else nh = 0;
            ++iters;
        }
        double gmax = 0.0;
        for (int k = tid; k < D; k += TRAIN_T) {
            const double xv = train_get_xt(A, s, k), gv = train_gt(A, s, k);
            x[k] = xv; g[k] = gv; gmax = fmax(gmax, fabs(gv));
        }
        gmax = train_block_max(gmax, sh);
        fx = ft;
        if (gmax <= A.g_tol) status = 1;
        else if (!first && dxmax == 0.0) status = 1;              // no change: stop (the host optimiser does the same)
        else if (iters >= A.maxiters) status = 3;
        if (status == 0) {
            // two-loop recursion over the history slots, newest first
            __syncthreads();
            for (int k = tid; k < D; k += TRAIN_T) d[k] = g[k];    // q
            for (int j = 0; j < nh; ++j) {
                const int slot = (head - 1 - j + 2 * A.m) % A.m;
                double a = 0.0;
                for (int k = tid; k < D; k += TRAIN_T) a += A.hs[((size_t)slot * S + s) * D + k] * d[k];
                a = train_block_sum(a, sh) * A.rho[(size_t)slot * S + s];
                if (tid == 0) s_alpha[j] = a;
                for (int k = tid; k < D; k += TRAIN_T) d[k] -= a * A.hy[((size_t)slot * S + s) * D + k];
            }
            double gam = 1.0;
            if (nh > 0) {
                const int slot = (head - 1 + A.m) % A.m;
                double yy = 0.0, sy = 0.0;
                for (int k = tid; k < D; k += TRAIN_T) {
                    const double yv = A.hy[((size_t)slot * S + s) * D + k];
                    yy += yv * yv; sy += yv * A.hs[((size_t)slot * S + s) * D + k];
                }
                yy = train_block_sum(yy, sh); sy = train_block_sum(sy, sh);
                if (yy > 0.0) gam = sy / yy;
            }
            for (int k = tid; k < D; k += TRAIN_T) d[k] *= gam;    // r
            for (int j = nh - 1; j >= 0; --j) {
                const int slot = (head - 1 - j + 2 * A.m) % A.m;
                double b = 0.0;
                for (int k = tid; k < D; k += TRAIN_T) b += A.hy[((size_t)slot * S + s) * D + k] * d[k];
                b = train_block_sum(b, sh) * A.rho[(size_t)slot * S + s];
                const double coef = s_alpha[j] - b;
                for (int k = tid; k < D; k += TRAIN_T) d[k] += coef * A.hs[((size_t)slot * S + s) * D + k];
            }
            double dp = 0.0;
            for (int k = tid; k < D; k += TRAIN_T) { d[k] = -d[k]; dp += g[k] * d[k]; }
            dp = train_block_sum(dp, sh);
            if (!(dp < 0.0)) {                                      // not a descent direction: steepest descent, drop the history
                __syncthreads();
                dp = 0.0;
                for (int k = tid; k < D; k += TRAIN_T) { d[k] = -g[k]; dp -= g[k] * g[k]; }
                dp = train_block_sum(dp, sh);
                nh = 0;
            }
            dphi0 = dp; alpha = 1.0; ls = 0;
            if (tid == 0) { sc[SC_APREV] = 1.0; sc[SC_FPREV] = fx; }
            for (int k = tid; k < D; k += TRAIN_T) train_set_xt(A, s, k, x[k] + d[k]);
        }
    } else if (status == 0) {
        // backtrack: quadratic interpolation on the first, cubic afterwards; clamped to [rho_lo, rho_hi] * alpha
        ++ls;
        if (ls >= A.ls_maxiter) status = 2;
        else {
            const double ap = sc[SC_APREV], fp = sc[SC_FPREV];
            double an;
            if (!fin) an = 0.5 * alpha;
            else {
                if (ls == 1) an = -(dphi0 * alpha * alpha) / (2.0 * (ft - fx - dphi0 * alpha));
                else {
                    const double div = 1.0 / (ap * ap * alpha * alpha * (alpha - ap));
                    const double t1 = ft - fx - dphi0 * alpha, t2 = fp - fx - dphi0 * ap;
                    const double ca = (ap * ap * t1 - alpha * alpha * t2) * div;
                    const double cb = (-ap * ap * ap * t1 + alpha * alpha * alpha * t2) * div;
                    const double disc = cb * cb - 3.0 * ca * dphi0;
                    an = fabs(ca) < 1e-300 ? -dphi0 / (2.0 * cb) : (-cb + sqrt(fmax(disc, 0.0))) / (3.0 * ca);
                }
                if (!(an - an == 0.0)) an = alpha * A.rho_hi;
                an = fmin(fmax(an, alpha * A.rho_lo), alpha * A.rho_hi);
            }
            __syncthreads();
            if (tid == 0) { sc[SC_APREV] = alpha; sc[SC_FPREV] = ft; }
            alpha = an;
            for (int k = tid; k < D; k += TRAIN_T) train_set_xt(A, s, k, x[k] + alpha * d[k]);
        }
    }
    if (status != 0) for (int k = tid; k < D; k += TRAIN_T) train_set_xt(A, s, k, x[k]);
    __syncthreads();
    if (tid == 0) {
        sc[SC_FX] = fx; sc[SC_ALPHA] = alpha; sc[SC_DPHI0] = dphi0;
        ic[IC_NH] = nh; ic[IC_HEAD] = head; ic[IC_LS] = ls; ic[IC_ITERS] = iters; ic[IC_STATUS] = status;
        if (status == 0) atomicAdd(&A.status_count[0], 1);
    }
}

}  // namespace cude
