// =====================================================================================
// cude_warp.cuh — the LATENCY form of loss + gradient: one WARP (or a group of 8 lanes) per trajectory.
//
// Small batches — config 1 (57 individuals, stored weights), the 25 selected starts of `train`
// (reference src/parameter-estimation.jl:374-376: 1425 trajectories per optimiser iteration) — leave most of the GPU idle
// with one thread per trajectory, and what the caller waits for is the latency of one thread walking ~20 adaptive steps and
// their adjoint: ~100 sequential network evaluations forward, ~100 forward + backward in the sweep.  Here the G lanes of
// a group (G = 32: a whole warp, calls of up to 512 trajectories; G = 8: four trajectories per warp, up to 8192) share one trajectory:
//   forward  every lane carries the state redundantly (uniform control flow, no divergence); lane q % 5 evaluates the
//            network at node q of the step (the production term does not depend on the state, so the 5 nodes of a step
//            are independent), 5 shuffles hand the values round, the stage arithmetic is the fused kernel's
//            (cude_kernels.cuh) expression by expression: the per-trajectory sse is BIT-IDENTICAL to the fused kernel's;
//   adjoint  the linear stage recursion (cheap, sequential) leaves a weight per (step, node); the ~100 (step, node)
//            network forward + backward evaluations are then spread over the lanes (3-4 each instead of 100), the
//            33 accumulators of the lanes are summed by shuffles: gradients equal to summation order (1e-16).
// Records {t, h}, {dG[5]}, {w[5]} of up to WARP_CAP accepted steps live in shared memory; longer solves (tight
// tolerances) are handed to the fused kernel through the same flagged-block list as in the two-kernel gradient.
// One row {sse, d sse/d neural} per trajectory goes to global memory; cude_warp_reduce sums each start's rows in
// individual order (deterministic).
// =====================================================================================
#pragma once
#include "cude_kernels.cuh"
#include "cude_split.cuh"   // reduce_region

namespace cude {

#ifndef CUDE_WARP_CAP
#define CUDE_WARP_CAP 64
#endif
constexpr int WARP_CAP = CUDE_WARP_CAP;
constexpr int WARP_TPB = 4;            // warps per block (32 / G trajectories each)

struct WarpArgs {
    PopDev pop;
    int n_starts;
    const double* neural;              // start s: neural + s*neural_stride
    long long neural_stride;
    const double* cond;                // [N x S]
    double abstol, reltol;
    int maxiters;
    double cond_scale;
    double* sse_out;                   // [N x S] or nullptr
    double* g_cond;                    // [N x S]
    double* rows;                      // [N x S][P+1]: {sse, d sse / d neural} per trajectory
    unsigned long long* counters;
    // fallback for solves longer than WARP_CAP accepted steps: the fused kernel's tile geometry
    int* ovf;                          // [N x S]: -1 = handed to the fused kernel, 0 otherwise (its only_flag)
    int* blkflag;                      // [S x nchunks]
    int* blklist;
    int* blkcount;
    int fb_block, nchunks;             // individuals per fused-kernel block, blocks per start
};

__host__ __device__ inline size_t warp_per_warp_doubles(int P, int K, int M) {
    return (size_t)((P + 1) & ~1) + (size_t)3 * K + (size_t)3 * M + (size_t)12 * WARP_CAP;
}
__host__ __device__ inline size_t warp_smem_doubles(int P, int K, int M, int G = 32) {
    return (size_t)256 + (size_t)WARP_TPB * (32 / G) * warp_per_warp_doubles(P, K, M);
}

#if !defined(CUDE_HOST_EMU) || defined(CUDE_HOST_EMU_WARP)      // the host emulation of this kernel needs 32 cooperating lanes (tests/emu/emu_warp.cpp)
// GRAD = false: the forward pass only (loss-only calls of small batches, e.g. the 2 x N solves of a SAEM Metropolis step):
// sse_out[j] per trajectory, no records, no rows.
// G = lanes per trajectory (32, 16 or 8): with G < 32 a warp carries 32 / G trajectories side by side — the groups run the same
// code on their own data and may diverge from each other (step counts, accept / reject); every shuffle and warp barrier names
// only the group's lanes.  Fewer lanes repeat the state arithmetic, so more trajectories fit a wave; the adjoint's (step, node)
// evaluations take 32 / G times as many rounds.
template <class NS, bool GRAD = true, int G = 32>
__global__ void __launch_bounds__(32 * WARP_TPB, 3) cude_warp_kernel(const WarpArgs A) {
    using namespace tab;
    static_assert(G == 32 || G == 16 || G == 8, "lanes per trajectory");
    constexpr int W = NS::W, P = NS::P, PP = (P + 1) & ~1, TPW = 32 / G;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & (G - 1);                 // lane within the trajectory's group
    const int wid = (threadIdx.x >> 5) * TPW + ((threadIdx.x & 31) / G);     // trajectory slot within the block
    const unsigned FULL = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));   // the group's lanes
    const int N = A.pop.n_ind, K = A.pop.max_knots, M = A.pop.max_obs;
    double* const sTab = smem;
    double* const sWs = smem + 256 + (size_t)wid * warp_per_warp_doubles(P, K, M);
    double* const sKt = sWs + PP;
    double* const sKg = sKt + K;
    double* const sSl = sKg + K;
    double* const sOt = sSl + K;
    double* const sOy = sOt + M;
    double* const sRes = sOy + M;
    double* const rTH = sRes + M;                      // [WARP_CAP][2]  {t, h}
    double* const rDG = rTH + 2 * WARP_CAP;            // [WARP_CAP][5]  dG at the step's nodes
    double* const rW = rDG + 5 * WARP_CAP;             // [WARP_CAP][5]  node weights of the adjoint

    for (int p = threadIdx.x; p < 256; p += blockDim.x) sTab[p] = EXP_TAB256[p];
    const long long j = (long long)blockIdx.x * (WARP_TPB * TPW) + wid;
    const bool active = j < (long long)N * A.n_starts;                    // warp-uniform
    int s = 0, i = 0, nk = 2, nobs = 0;
    if (active) {
        s = (int)(j / N);
        i = (int)(j - (long long)s * N);
        const double* gW = A.neural + (long long)s * A.neural_stride;
        for (int p = lane; p < P; p += G) sWs[p] = gW[p];
        nk = A.pop.n_knots[i];
        for (int k = lane; k < nk; k += G) {
            sKt[k] = A.pop.knot_t[(size_t)k * N + i];
            sKg[k] = A.pop.knot_g[(size_t)k * N + i];
            if (k < nk - 1) sSl[k] = A.pop.slope[(size_t)k * N + i];
        }
        nobs = A.pop.n_obs[i];
        for (int k = lane; k < nobs; k += G) {
            sOt[k] = A.pop.obs_t[(size_t)k * N + i];
            sOy[k] = A.pop.obs_y[(size_t)k * N + i];
        }
    }
    __syncthreads();
    if (!active) return;

    const double* const sW = sWs;
    const int q5 = lane % 5;                           // the node of a step this lane evaluates
    Kin Kc;
    Kc.k0 = A.pop.k0[i]; Kc.k1 = A.pop.k1[i]; Kc.k2 = A.pop.k2[i]; Kc.c0 = A.pop.c0[i];
    Kc.d00 = -(Kc.k0 + Kc.k2); Kc.kc = Kc.k0 * Kc.c0;
    Knots kn;
    kn.t = sKt; kn.g = sKg; kn.sl = sSl; kn.nk = nk; kn.stride = 1;
    kn.g0 = kn.g[0];
    const double t0 = sKt[0], tend = sKt[nk - 1];
    const double beta = m_exp(A.cond[j]);
    const double covv = (NS::NIN > 2 && A.pop.cov) ? A.pop.cov[i] : 0.0;
    double c[W];
#pragma unroll
    for (int q = 0; q < W; ++q) {
        double z = fma(sW[W + q], beta, sW[NS::NIN * W + q]);
        if (NS::NIN > 2) z = fma(sW[2 * W + q], covv, z);
        c[q] = z;
    }
    const double abstol = A.abstol, reltol = A.reltol;
    const double dtmax = tend - t0;
    const double at0 = fabs(t0), at1 = fabs(tend);
    const double dtmin = fmax(nextafter(at0, CUDART_INF) - at0, nextafter(at1, CUDART_INF) - at1);
    const double snap = 100.0 * (nextafter(at1, CUDART_INF) - at1);

    // =================== forward pass (the fused kernel's arithmetic; every lane carries the whole state) ===================
    double u0 = Kc.c0, u1 = (Kc.k2 / Kc.k1) * Kc.c0;   // c-peptide-models.jl:185
    double t = t0;
    int iobs = 0, na = 0, nr = 0;
    double fsse = 0.0;
    double next_ot = (nobs > 0) ? sOt[0] : CUDART_INF;
    while (iobs < nobs && next_ot <= t0) {
        const double r = u0 - sOy[iobs];
        if (lane == 0) sRes[iobs] = r;
        fsse = fma(r, r, fsse);
        ++iobs;
        next_ot = (iobs < nobs) ? sOt[iobs] : CUDART_INF;
    }
    double k10, k11;
    kinetics(Kc, u0, u1, 0.0, k10, k11);
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
    double dt = dt0, nn0 = 0.0, lnqold = -9.210340371976182;
    int ret = 0, iter = 0;
    bool init = true;
    for (;;) {
        const double* cn;
        if (init) cn = CN_INIT;
        else {
            if (!(ret == 0 && t < tend)) break;
            if (++iter > A.maxiters) { ret = 1; break; }
            dt = fmin(dt, tend - t);
            if (!(dt > dtmin)) { ret = (dt != dt) ? 3 : 2; break; }
            cn = CN_STEP;
        }
        // this lane's node of the pass: interpolated glucose, network, softplus; then the 5 values to every lane
        const double dg = kn.dG(fma(cn[q5], dt, t));
        double spq, ddq;
        t_softplus_d(mlp_zout<NS>(sW, sTab, c, dg), sTab, spq, ddq);
        double sp[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) sp[q] = __shfl_sync(FULL, spq, q, G);
        if (init) {
            init = false;
            nn0 = sp[0];
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
        const double p2 = sp[0] - nn0, p3 = sp[1] - nn0, p4 = sp[2] - nn0, p5 = sp[3] - nn0, p6 = sp[4] - nn0;
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
        f0 = dt * fma(e7, k70, se0);
        f1 = dt * fma(e7, k71, se1);
        f0 = f0 * m_rcp(fma(fmax(fabs(u0), fabs(un0)), reltol, abstol));
        f1 = f1 * m_rcp(fma(fmax(fabs(u1), fabs(un1)), reltol, abstol));
        const double E2 = m_sumsq(f0, f1) * 0.5;
        if (!(E2 == E2) || !isfinite(un0) || !isfinite(un1)) { ret = 3; break; }
        const double lnE = 0.5 * m_log_pos(E2);
        if (E2 <= 1.0) {
            const double q = fmax(1.0 / qmax, fmin(1.0 / qmin, t_exp_sat(fma(beta1, lnE, -beta2 * lnqold), sTab) * (1.0 / gamma)));
            double tnew = t + dt;
            if (fabs(tnew - tend) < snap) tnew = tend;
            while (iobs < nobs && next_ot <= tnew) {
                double y;
                if (next_ot == tnew) y = un0;
                else {
                    double bw[7];
                    dense_weights((next_ot - t) * m_rcp(dt), bw);
                    const double sdo = fma(bw[0], k10, fma(bw[1], k20, fma(bw[2], k30, fma(bw[3], k40, fma(bw[4], k50, fma(bw[5], k60, bw[6] * k70))))));
                    y = fma(dt, sdo, u0);
                }
                const double r = y - sOy[iobs];
                if (lane == 0) sRes[iobs] = r;
                fsse = fma(r, r, fsse);
                ++iobs;
                next_ot = (iobs < nobs) ? sOt[iobs] : CUDART_INF;
            }
            if (GRAD && na < WARP_CAP) {
                if (lane == 0) { rTH[2 * na] = t; rTH[2 * na + 1] = dt; }
                if (lane < 5) rDG[5 * na + lane] = dg;
            }
            ++na;
            lnqold = fmax(lnE, -9.210340371976182);
            dt = fmin(dt * m_rcp(q), dtmax);
            t = tnew; u0 = un0; u1 = un1; k10 = k70; k11 = k71;
        } else {
            ++nr;
            dt = dt * m_rcp(fmin(1.0 / qmin, t_exp_sat(beta1 * lnE, sTab) * (1.0 / gamma)));
        }
    }
    if (ret == 0 && iobs < nobs) ret = 3;
    const bool failed = (ret != 0);
    const double sse = failed ? CUDART_INF : fsse;
    const bool overflow = GRAD && !failed && na > WARP_CAP;
    __syncwarp(FULL);
    if constexpr (!GRAD) {
        if (lane == 0) {
            A.sse_out[j] = sse;
            if (A.counters) {
                atomicAdd(&A.counters[0], (unsigned long long)na);
                atomicAdd(&A.counters[1], (unsigned long long)nr);
                if (failed) atomicAdd(&A.counters[2], 1ull);
            }
        }
    } else {

    double acc[NS::NACC];
#pragma unroll
    for (int k = 0; k < NS::NACC; ++k) acc[k] = 0.0;
    double gcond = 0.0;
    if (!failed && !overflow) {
        // =================== adjoint, part 1: the linear stage recursion (cude_adjoint_kernel's), weights per node ===================
        const double d00 = Kc.d00, kk1 = Kc.k1, kk2 = Kc.k2;
        double lam0 = 0.0, lam1 = 0.0, wnode = 0.0, wsum = 0.0, t_next = tend;
        int kobs_top = nobs - 1;
        for (int n = na - 1; n >= 0; --n) {
            const double tn = rTH[2 * n], h = rTH[2 * n + 1];
            double kb[7][2];
#pragma unroll
            for (int q = 0; q < 7; ++q) { kb[q][0] = 0.0; kb[q][1] = 0.0; }
            double ub0 = 0.0, ub1 = 0.0;
            while (kobs_top >= 0) {
                const double ts = sOt[kobs_top];
                if (!(ts > tn)) break;
                const double wr = 2.0 * sRes[kobs_top];
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
            }
            const double pb7 = kb[6][0];
            lam0 = fma(d00, kb[6][0], lam0);
            lam1 = fma(kk1, kb[6][0], lam1);
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
            double gb0, gb1, hg0, hg1;
#define CUDE_STAGE_BACK(I)                                                    \
    gb0 = fma(d00, kb[I][0], kk2 * kb[I][1]);                                 \
    gb1 = kk1 * (kb[I][0] - kb[I][1]);                                        \
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
            const double w6 = pb6 + pb7 + wnode;
            if (lane == 0) { rW[5 * n] = pb2; rW[5 * n + 1] = pb3; rW[5 * n + 2] = pb4; rW[5 * n + 3] = pb5; rW[5 * n + 4] = w6; }
            wsum += w6 + pb5 + pb4 + pb3 + pb2;
            wnode = pb1;
            lam0 = ub0; lam1 = ub1;
            t_next = tn;
        }
        __syncwarp(FULL);
        // =================== adjoint, part 2: the (step, node) network evaluations spread over the lanes ===================
        for (int p = lane; p < 5 * na; p += G) mlp_backward<NS, double>(sW, sTab, c, rDG[p], rW[p], acc);
        // the NN([0; beta]) term (c-peptide-models.jl:91): one node at dG = 0 with weight -sum(w), on the least loaded lane
        if (lane == G - 1) mlp_backward<NS, double>(sW, sTab, c, 0.0, -wsum, acc);
#pragma unroll
        for (int k = 0; k < NS::NACC; ++k) {
            double v = acc[k];
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o, G);
            acc[k] = v;
        }
        double db = 0.0;
#pragma unroll
        for (int q = 0; q < W; ++q) db = fma(acc[W + q], sW[W + q], db);
        gcond = db * beta * A.cond_scale;
    }
    // =================== outputs ===================
    double* const row = A.rows + (size_t)j * (P + 1);
    const bool has_grad = !failed && !overflow;               // otherwise the accumulators are zero, but beta may be NaN
#pragma unroll
    for (int p = -1; p < P; ++p) {
        double v;
        if (p < 0) v = overflow ? 0.0 : sse;                  // an overflowed trajectory's sse and gradient come from the fallback
        else if (!has_grad) v = 0.0;
        else if (p < W) v = acc[p];
        else if (p < 2 * W) v = acc[W + (p - W)] * beta;
        else if (NS::NIN > 2 && p < 3 * W) v = acc[W + (p - 2 * W)] * covv;
        else if (p < NS::L1) v = acc[W + (p - NS::NIN * W)];
        else v = acc[2 * W + (p - NS::L1)];
        if (lane == ((p + 1) & (G - 1))) row[p + 1] = v;
    }
    if (lane == 0) {
        if (A.sse_out) A.sse_out[j] = sse;
        if (!overflow) A.g_cond[j] = gcond;
        A.ovf[j] = overflow ? -1 : 0;
        if (overflow) {
            const int cch = i / A.fb_block, prow = s * A.nchunks + cch;
            if (atomicExch(&A.blkflag[prow], 1) == 0) A.blklist[atomicAdd(A.blkcount, 1)] = cch * A.n_starts + s;
        }
        if (A.counters) {
            atomicAdd(&A.counters[0], (unsigned long long)na);
            atomicAdd(&A.counters[1], (unsigned long long)nr);
            if (failed) atomicAdd(&A.counters[2], 1ull);
        }
    }
    }   // GRAD
}

#endif
#ifndef CUDE_HOST_EMU
// sums[(P+1) x S]: start s = its N trajectory rows in individual order + the partial rows of its flagged fallback blocks
// (nw rows per block, block order).  One block per start, thread = column q x row group; fixed order: deterministic.
__global__ void __launch_bounds__(RED_T) cude_warp_reduce(const double* __restrict__ rows, int N, const double* __restrict__ pC,
                                                          const int* __restrict__ blkflag, int nchunks, int nw, int np1,
                                                          double* __restrict__ sums) {
    __shared__ double part[RED_T];
    const int s = blockIdx.x, q = threadIdx.x % RED_Q, rg = threadIdx.x / RED_Q, nrg = RED_T / RED_Q;
    double v = 0.0;
    if (q < np1) {
        v = reduce_region(rows + (size_t)s * N * np1, N, 0, N, np1, q, rg, nrg);
        if (rg == 0)
            for (int cch = 0; cch < nchunks; ++cch)
                if (blkflag[s * nchunks + cch])
                    for (int w = 0; w < nw; ++w) v += pC[(((size_t)s * nchunks + cch) * nw + w) * np1 + q];
    }
    part[threadIdx.x] = v;
    __syncthreads();
    if (rg == 0 && q < np1) {
        for (int g = 1; g < nrg; ++g) v += part[g * RED_Q + q];
        sums[(size_t)s * np1 + q] = v;
    }
}
#endif  // !CUDE_HOST_EMU

}  // namespace cude
