// =====================================================================================
// cude_math.cuh — branch-free FP64 elementary functions for the cUDE kernels.
//
// The MLP (8 tanh + softplus per evaluation) dominates the instruction stream, and CUDA's libm
// versions (branches for special cases, ~4x the instructions, large code footprint) made the v1
// kernel instruction-fetch bound (ncu: "no_instruction" stall 4.5 per issue, profiles/r01_v1_*).
// These versions are straight-line code on the FP64 pipe:
//   exp core   : Cody-Waite reduction (magic-number rounding) to |r| <= ln2/128, a 64-entry 2^(j/64) table in
//                shared memory, a degree-5 near-minimax polynomial (Chebyshev-node fit) + exponent-field add.
//   reciprocal : MUFU.RCP64H seed + 2 Newton steps (4 DFMA).
//   tanh       : 1 - 2/(exp(2x)+1) on x clamped to [-20,20] (tanh == +-1 beyond 19.06 in FP64);
//                absolute error <= ~2e-16 (relative error grows like 1e-16/|x| near 0, irrelevant here).
//   log core   : exponent split + 2*atanh((m-1)/(m+1)) with a degree-6 polynomial in q^2.
// NaN inputs propagate (comparison-based clamps keep NaN; hardware NaNs are canonical, low word 0).
// =====================================================================================
#pragma once
#ifndef CUDE_HOST_EMU
#include <cuda_runtime.h>
#endif

namespace cude {

#ifdef CUDE_HOST_EMU
// host stand-ins (tests/emu): same bit manipulations through memcpy, a float-accurate reciprocal seed
static inline int __double2hiint(double x) { long long b; memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { long long b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) { long long b = ((long long)hi << 32) | (unsigned int)lo; double x; memcpy(&x, &b, 8); return x; }
static inline double rcp_seed(double d) { return (double)(float)(1.0 / d); }
#else
__device__ __forceinline__ double rcp_seed(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));   // MUFU.RCP64H: ~20 correct bits
    return y;
}
#endif

// 1/d for normal positive d (no special cases needed by the callers): seed y0 with relative error
// e (|e| <~ 2^-20), then y0*(1 + e + e^2) leaves e^3 — one third-order step, 3 DFMA.
__device__ __forceinline__ double m_rcp(double d) {
    const double y = rcp_seed(d);
    const double e = fma(-d, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}

// 2^(j/64), j = 0..63 — staged into shared memory by the kernel (per-lane indices: shared memory serves
// them in <= 4 wavefronts, constant memory would serialise 32-way)
__constant__ double EXP_TAB64[64] = {1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284, 1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199, 1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418, 1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812, 1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687, 1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783, 1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303, 1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112, 1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647, 1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384, 1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267, 1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364, 1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062, 1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989, 1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656, 1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};

// exp(x) for |x| <= ~700 (callers clamp); branch-free.  x = (64 k + j) ln2/64 + r, |r| <= ln2/128:
// exp(x) = 2^k * 2^(j/64) * exp(r), exp(r) by a degree-5 near-minimax polynomial (max rel. error 2.2e-18
// before rounding).  10 FP64 instructions instead of 16 for the table-free degree-11 version (kept under
// CUDE_EXP_POLY11): the MLP regions of the kernel are FP64-pipe bound (ncu v4), so this is throughput.
__device__ __forceinline__ double m_exp_core(double x, const double* __restrict__ tab) {
#if !defined(CUDE_EXP_POLY11)
    const double SHIFT = 6755399441055744.0;   // 1.5 * 2^52
    const double t = fma(x, 92.33248261689366, SHIFT);      // 64/ln2
    const int n = __double2loint(t);
    const double nf = t - SHIFT;
    double r = fma(nf, -0.01083042469326756, x);            // ln2/64 hi
    r = fma(nf, -2.9815858269852933e-12, r);                // ln2/64 lo
    const double m = tab[n & 63];
    double p = 0.008333342062620064;
    p = fma(p, r, 0.04166672777168178);
    p = fma(p, r, 0.16666666666657065);
    p = fma(p, r, 0.4999999999993279);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p *= m;                                                  // in [0.99, 1.99]
    return __hiloint2double(__double2hiint(p) + ((n >> 6) << 20), __double2loint(p));
#else
    (void)tab;
    const double L2E = 1.4426950408889634, SHIFT = 6755399441055744.0;   // 1.5 * 2^52
    const double LN2_HI = 0.6931471803691238, LN2_LO = 1.9082149292705877e-10;
    const double t = fma(x, L2E, SHIFT);
    const int n = __double2loint(t);
    const double nf = t - SHIFT;
    double r = fma(nf, -LN2_HI, x);
    r = fma(nf, -LN2_LO, r);
    // degree-11 near-minimax polynomial (max rel. error 1.6e-17), even/odd split: two half-depth chains
    const double r2 = r * r;
    double pe = 2.763265472252779e-07, po = 2.5110049204818658e-08;
    pe = fma(pe, r2, 2.4801485441561313e-05); po = fma(po, r2, 2.755724088722987e-06);
    pe = fma(pe, r2, 0.0013888888952352863);  po = fma(po, r2, 0.00019841269890076403);
    pe = fma(pe, r2, 0.04166666666648795);    po = fma(po, r2, 0.008333333333319589);
    pe = fma(pe, r2, 0.5000000000000019);     po = fma(po, r2, 0.1666666666666668);
    pe = fma(pe, r2, 1.0);                    po = fma(po, r2, 1.0);
    const double p = fma(po, r, pe);
    // p in [0.70, 1.42]; multiply by 2^n through the exponent field
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
#endif
}

__device__ __forceinline__ double m_clamp(double x, double lo, double hi) {
    x = (x < lo) ? lo : x;     // comparisons keep NaN (fmin/fmax would swallow it)
    return (x > hi) ? hi : x;
}

// |x| >= 20 saturates (integer compare on the high word keeps the clamp off the FP64 pipe); NaN propagates
__device__ __forceinline__ double m_tanh(double x, const double* __restrict__ tab) {
    const int hi = __double2hiint(x);
    const int ahi = hi & 0x7fffffff;
    const bool big = ahi >= 0x40340000;                       // |x| >= 20, Inf or NaN
    const double xc = __hiloint2double(big ? ((hi & 0x80000000) | 0x40340000) : hi, big ? 0 : __double2loint(x));
    const double e = m_exp_core(xc + xc, tab);
    const double r = fma(-2.0, m_rcp(e + 1.0), 1.0);
    return (ahi > 0x7ff00000) ? x : r;                        // NaN in -> NaN out
}

// log(s) for normal positive s
__device__ __forceinline__ double m_log_core(double s) {
    int hi = __double2hiint(s);
    const int lo = __double2loint(s);
    int k = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;          // m in [1, 2)
    if (hi >= 0x3ff6a09e) { hi -= 0x00100000; ++k; }   // m in [sqrt(2)/2, sqrt(2))   (predicated)
    const double m = __hiloint2double(hi, lo);
    const double q = (m - 1.0) * m_rcp(m + 1.0);
    const double w = q * q;
    double g = 0.07308903576708663;
    g = fma(g, w, 0.07665805861027504);
    g = fma(g, w, 0.09091446216268435);
    g = fma(g, w, 0.11111105544689748);
    g = fma(g, w, 0.14285714313126918);
    g = fma(g, w, 0.19999999999949444);
    g = fma(g, w, 0.3333333333333335);
    const double LN2_HI = 0.6931471803691238, LN2_LO = 1.9082149292705877e-10;
    const double kf = (double)k;
    // log(s) = k ln2 + 2q + 2q*w*g
    double res = fma(kf, LN2_LO, (q + q) * (w * g));
    res += q + q;
    return fma(kf, LN2_HI, res);
}

// softplus(x) = log(1 + exp(x)), the naive form of reference src/neural-network.jl:13-15, including its
// overflow: exp(x) = Inf for x > 709.78 -> Inf.  For 36.8 < x the naive form equals x in FP64; for
// x < -40 it equals 0 (1 + exp(x) rounds to 1), as here.
__device__ __forceinline__ double m_softplus(double x, const double* __restrict__ tab) {
    const double xc = m_clamp(x, -40.0, 36.8);
    double sp = m_log_core(1.0 + m_exp_core(xc, tab));
    sp = (x > 36.8) ? x : sp;
    sp = (x > 709.782712893384) ? CUDART_INF : sp;
    return (x != x) ? x : sp;                      // the log core does not propagate NaN by itself
}

// d softplus / dx = 1/(1+exp(-x))
__device__ __forceinline__ double m_sigmoid(double x, const double* __restrict__ tab) {
    const double xc = m_clamp(-x, -40.0, 40.0);
    return m_rcp(1.0 + m_exp_core(xc, tab));
}

// natural log for the step controller: EEst^b1 / qold^b2 = exp(b1 ln EEst - b2 ln qold); the controller
// clamps the result to [1/qmax, 1/qmin], so saturating the exponent at +-40 changes nothing.
__device__ __forceinline__ double m_log_pos(double x) {
    const double xs = m_clamp(x, 1e-300, 1e300);
    return m_log_core(xs);
}
__device__ __forceinline__ double m_exp_sat(double x, const double* __restrict__ tab) { return m_exp_core(m_clamp(x, -40.0, 40.0), tab); }
__device__ __forceinline__ double m_log10(double x) {
    const double xs = (x < 1e-300) ? 1e-300 : x;
    return m_log_core(xs) * 0.4342944819032518;
}
__device__ __forceinline__ double m_pow10(double x, const double* __restrict__ tab) { return m_exp_core(m_clamp(x * 2.302585092994046, -700.0, 700.0), tab); }

// ---------------------------------------------------------------- FP32 network math (precision = 1, "mixed")
// Optional mode with a documented looser bound: the MLP is evaluated in FP32 on the FMA / MUFU pipes
// (ex2.approx, rcp.approx, lg2.approx: ~1e-7 relative each), the integrator, adjoint and reductions stay FP64.
#ifdef CUDE_HOST_EMU
static inline float f_ex2(float x) { return exp2f(x); }
static inline float f_rcp(float x) { return 1.0f / x; }
static inline float f_lg2(float x) { return log2f(x); }
#else
__device__ __forceinline__ float f_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif
__device__ __forceinline__ float f_clamp(float x, float lo, float hi) {
    x = (x < lo) ? lo : x;     // comparisons keep NaN
    return (x > hi) ? hi : x;
}
__device__ __forceinline__ float m_tanh(float x, const double* __restrict__) {
    const float e = f_ex2(f_clamp(x, -15.0f, 15.0f) * 2.8853900817779268f);     // exp(2x)
    return fmaf(-2.0f, f_rcp(e + 1.0f), 1.0f);
}
__device__ __forceinline__ float m_softplus(float x, const double* __restrict__) {
    const float e = f_ex2(f_clamp(x, -30.0f, 15.0f) * 1.4426950408889634f);
    float sp = f_lg2(1.0f + e) * 0.6931471805599453f;
    sp = (x > 15.0f) ? x : sp;
    sp = (x > 709.782712893384f) ? (float)CUDART_INF : sp;      // the reference's naive form overflows there
    return (x != x) ? x : sp;
}
__device__ __forceinline__ float m_sigmoid(float x, const double* __restrict__) {
    return f_rcp(1.0f + f_ex2(f_clamp(-x, -30.0f, 30.0f) * 1.4426950408889634f));
}

// exp(cond): once per trajectory, full range semantics (Inf / 0 / NaN) from the CUDA library
__device__ __forceinline__ double m_exp(double x) { return exp(x); }

}  // namespace cude
