// =====================================================================================
// cude_math.cuh — branch-free FP64 elementary functions for the cUDE kernels.
//
// The MLP (8 tanh + softplus per evaluation) dominates the instruction stream, and CUDA's libm
// versions (branches for special cases, ~4x the instructions, large code footprint) made the v1
// kernel instruction-fetch bound (ncu: "no_instruction" stall 4.5 per issue, profiles/r01_v1_*).
// These versions are straight-line code on the FP64 pipe:
//   exp parts  : magic-number rounding + one-constant reduction to |r| <= ln2/512, a 256-entry 2^(j/256) table in
//                shared memory, a degree-4 Taylor polynomial; 2^k is added on the integer pipe (see below).
//   reciprocal : MUFU.RCP64H seed + one third-order step (3 DFMA).
//   tanh       : 1 - 2/(exp(2x)+1) on |x| clamped to 20 (tanh == +-1 beyond 19.06 in FP64);
//                absolute error <= ~2e-16 (relative error grows like 1e-16/|x| near 0, irrelevant here).
//   log core   : exponent split + 2*atanh((m-1)/(m+1)) with a degree-6 polynomial in q^2.
// NaN inputs propagate (comparison-based clamps keep NaN; tanh tracks them on the integer pipe, t_nan_inject).
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
static inline double __hiloint2double(int hi, int lo) { unsigned long long b = ((unsigned long long)(unsigned int)hi << 32) | (unsigned int)lo; double x; memcpy(&x, &b, 8); return x; }
static inline double rcp_seed(double d) { return (double)(float)(1.0 / d); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
#else
__device__ __forceinline__ double rcp_seed(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));   // MUFU.RCP64H: ~20 correct bits
    return y;
}
#endif

// a*a + b*b (+ c*c) with every product and sum rounded: nvcc is free to contract `a*a + b*b` into fma(a,a,b*b) or
// fma(b,b,a*a) and decides per instantiation, which made the forward passes of the kernel variants differ in the last
// bit of the error norm (and, through the adaptive steps, by ~1e-12 in the loss).  This form is also the oracle's.
#ifdef CUDE_HOST_EMU
static inline double m_sumsq(double a, double b) { return a * a + b * b; }
static inline double m_sumsq(double a, double b, double c) { return a * a + b * b + c * c; }
#else
__device__ __forceinline__ double m_sumsq(double a, double b) { return __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)); }
__device__ __forceinline__ double m_sumsq(double a, double b, double c) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c));
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

__device__ __forceinline__ double m_clamp(double x, double lo, double hi) {
    x = (x < lo) ? lo : x;     // comparisons keep NaN (fmin/fmax would swallow it)
    return (x > hi) ? hi : x;
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

// natural log for the step controller: EEst^b1 / qold^b2 = exp(b1 ln EEst - b2 ln qold); the controller
// clamps the result to [1/qmax, 1/qmin], so saturating the exponent at +-40 changes nothing.
__device__ __forceinline__ double m_log_pos(double x) {
    const double xs = m_clamp(x, 1e-300, 1e300);
    return m_log_core(xs);
}
__device__ __forceinline__ double m_log10(double x) {
    const double xs = (x < 1e-300) ? 1e-300 : x;
    return m_log_core(xs) * 0.4342944819032518;
}

// ---------------------------------------------------------------- exp / tanh / softplus / sigmoid (256-entry table)
// The MLP loops are bound by instruction issue and the FP64 pipe together (ncu v6: 45 % of the issued instructions are
// FP64, each occupying the pipe for two cycles), so these versions minimise *both* counts:
//   * 2^(j/256) table => |residual| <= ln2/512: a degree-4 Taylor polynomial suffices (remainder 3.8e-17 relative)
//     and its low coefficients (1, 1/2 resp. 1, 2, 2) are exact immediates;
//   * one-constant argument reduction r = fma(n, -C, x): the FMA is exact up to its final rounding, the only error is
//     C != ln2/256 (relative 1.1e-16), i.e. |x|*1.1e-16 relative in exp(x).  For tanh and softplus this is harmless
//     in the measure that matters (absolute error): d tanh = 2 e/(1+e)^2 * |2x| * 1.1e-16 <= 5e-17;
//   * the scale 2 of tanh's exp(2x) is folded into the constants; 2^k * table value is assembled on the integer
//     pipe, and "e + 1" is one DFMA on the parts: d = fma(M, q, 1);
//   * the clamp is three integer instructions (min on the high word), NaN inputs are tracked by one integer max
//     per call and re-injected once per network evaluation (t_nan_inject).
// tanh: 12 FP64 instructions (was 16), exp core 7 (was 10).
__constant__ double EXP_TAB256[256] = {1.0, 1.0027112750502025, 1.0054299011128027, 1.0081558981184175, 1.0108892860517005, 1.0136300849514894, 1.016378314910953, 1.019133996077738, 1.0218971486541166, 1.0246677928971357, 1.0274459491187637, 1.030231637686041, 1.0330248790212284, 1.0358256936019572, 1.0386341019613787, 1.041450124688316, 1.0442737824274138, 1.0471050958792898, 1.0499440858006872, 1.0527907730046264, 1.0556451783605572, 1.0585073227945128, 1.061377227289262, 1.0642549128844645, 1.0671404006768237, 1.0700337118202419, 1.0729348675259756, 1.075843889062791, 1.0787607977571199, 1.0816856149932152, 1.0846183622133092, 1.0875590609177697, 1.0905077326652577, 1.0934643990728858, 1.0964290818163769, 1.099401802630222, 1.102382583307841, 1.1053714457017412, 1.1083684117236787, 1.1113735033448175, 1.1143867425958924, 1.1174081515673693, 1.1204377524096067, 1.12347556733302, 1.1265216186082418, 1.129575928566288, 1.1326385195987192, 1.1357094141578055, 1.1387886347566916, 1.1418762039695616, 1.1449721444318042, 1.148076478840179, 1.1511892299529827, 1.154310420590216, 1.1574400736337511, 1.1605782120274988, 1.1637248587775775, 1.1668800369524817, 1.1700437696832502, 1.1732160801636373, 1.1763969916502812, 1.1795865274628758, 1.182784710984341, 1.1859915656609938, 1.189207115002721, 1.1924313825831512, 1.1956643920398273, 1.1989061670743806, 1.202156731452703, 1.2054161090051239, 1.2086843236265816, 1.2119613992768012, 1.215247359980469, 1.2185422298274085, 1.2218460329727576, 1.2251587936371455, 1.22848053610687, 1.2318112847340759, 1.2351510639369334, 1.2384998981998165, 1.241857812073484, 1.245224830175258, 1.2486009771892048, 1.2519862778663162, 1.255380757024691, 1.2587844395497165, 1.2621973503942507, 1.2656195145788063, 1.2690509571917332, 1.2724917033894028, 1.275941778396392, 1.2794012075056693, 1.2828700160787783, 1.2863482295460256, 1.2898358734066657, 1.2933329732290895, 1.2968395546510096, 1.3003556433796506, 1.3038812651919358, 1.3074164459346773, 1.3109612115247644, 1.3145155879493546, 1.318079601266064, 1.3216532776031575, 1.3252366431597413, 1.3288297242059544, 1.3324325470831615, 1.3360451382041458, 1.339667524053303, 1.3432997311868353, 1.3469417862329458, 1.3505937158920345, 1.3542555469368927, 1.3579273062129011, 1.3616090206382248, 1.365300717204012, 1.3690024229745905, 1.3727141650876684, 1.3764359707545302, 1.380167867260238, 1.383909881963832, 1.387662042298529, 1.3914243757719262, 1.3951969099662003, 1.3989796725383112, 1.4027726912202048, 1.4065759938190154, 1.4103896082172707, 1.4142135623730951, 1.4180478843204152, 1.4218926021691656, 1.4257477441054942, 1.42961333839197, 1.433489413367789, 1.4373759974489824, 1.4412731191286257, 1.4451808069770467, 1.449099089642035, 1.4530279958490526, 1.4569675544014438, 1.460917794180647, 1.4648787441464057, 1.4688504333369818, 1.4728328908693675, 1.4768261459394993, 1.4808302278224719, 1.4848451658727524, 1.488870989524397, 1.4929077282912648, 1.4969554117672355, 1.5010140696264256, 1.5050837316234065, 1.5091644275934228, 1.5132561874526098, 1.5173590411982147, 1.5214730189088146, 1.5255981507445384, 1.529734466947287, 1.533881997840956, 1.5380407738316568, 1.5422108254079407, 1.5463921831410214, 1.550584877685, 1.5547889397770887, 1.559004400237837, 1.5632312899713576, 1.567469639965553, 1.5717194812923414, 1.5759808451078865, 1.5802537626528246, 1.5845382652524937, 1.588834384317164, 1.593142151342267, 1.597461597908627, 1.6017927556826934, 1.606135656416771, 1.6104903319492543, 1.6148568142048607, 1.6192351351948637, 1.6236253270173289, 1.6280274218573478, 1.632441451987275, 1.6368674497669644, 1.6413054476440063, 1.645755478153965, 1.6502175739206177, 1.6546917676561943, 1.6591780921616162, 1.6636765803267364, 1.6681872651305825, 1.6727101796415966, 1.6772453570178785, 1.681792830507429, 1.6863526334483934, 1.6909247992693053, 1.6955093614893326, 1.7001063537185235, 1.7047158096580513, 1.709337763100463, 1.713972247929926, 1.718619298122478, 1.723278947746274, 1.7279512309618377, 1.732636182022311, 1.7373338352737062, 1.7420442251551564, 1.746767386199169, 1.7515033530318782, 1.7562521603732995, 1.761013843037584, 1.7657884359332727, 1.7705759740635547, 1.7753764925265212, 1.7801900265154245, 1.785016611318935, 1.789856282321401, 1.7947090750031072, 1.7995750249405351, 1.804454167806624, 1.809346539371032, 1.8142521755003989, 1.8191711121586085, 1.8241033854070534, 1.8290490314048973, 1.8340080864093424, 1.8389805867758937, 1.843966568958626, 1.8489660695104508, 1.8539791250833855, 1.8590057724288205, 1.864046048397789, 1.8690999899412386, 1.8741676341103, 1.8792490180565602, 1.8843441790323345, 1.8894531543909392, 1.8945759815869656, 1.8997126981765553, 1.9048633418176741, 1.9100279502703899, 1.9152065613971474, 1.9203992131630474, 1.925605943636125, 1.930826790987627, 1.9360617934922943, 1.9413109895286405, 1.9465744175792332, 1.9518521162309783, 1.9571441241754002, 1.9624504802089273, 1.9677712232331759, 1.9731063922552343, 1.978456026387951, 1.9838201648502194, 1.9891988469672663, 1.9945921121709402};

// exp(SC*x) = M * q for |SC*x| <= ~700 (callers clamp): M = 2^k 2^(j/256) exactly, q = exp(SC*residual).
// LO adds the second reduction constant (relative accuracy independent of |x|).
template <int SC, bool LO>
__device__ __forceinline__ void t_exp_parts(double x, const double* __restrict__ tab, double& M, double& q) {
    const double SHIFT = 6755399441055744.0;                          // 1.5 * 2^52
    const double t = fma(x, SC * 369.3299304675746, SHIFT);           // 256/ln2
    const int n = __double2loint(t);
    const double nf = t - SHIFT;
    double r = fma(nf, -0.0027076061740622863 / SC, x);               // ln2/256 (division by 2 is exact)
    if (LO) r = fma(nf, -9.058776616587108e-20 / SC, r);             // ln2/256 - double(ln2/256)
    const double m = tab[n & 255];
    M = __hiloint2double(__double2hiint(m) + ((n >> 8) * (1 << 20)), __double2loint(m));
    // exp(SC r), |SC r| <= ln2/512: 1 + s + s^2/2 + s^3/6 + s^4/24 with s = SC r
    double p = fma(SC == 1 ? 0.041666666666666664 : 0.6666666666666666, r, SC == 1 ? 0.16666666666666666 : 1.3333333333333333);
    p = fma(p, r, SC == 1 ? 0.5 : 2.0);
    p = fma(p, r, (double)SC);
    q = fma(p, r, 1.0);
}

// tanh(x) = 1 - 2/(exp(2x)+1); |x| clamped to 20 on the integer pipe (tanh == +-1 beyond 19.06 in FP64, +-Inf included).
// nanmax accumulates the largest |high word| seen: > 0x7ff00000 <=> some input was NaN (see t_nan_inject).
__device__ __forceinline__ double t_tanh(double x, const double* __restrict__ tab, int& nanmax) {
    const int hi = __double2hiint(x);
    const int ahi = hi & 0x7fffffff;
    nanmax = max(nanmax, ahi);
    const double xc = __hiloint2double(min(ahi, 0x40340000) | (hi & 0x80000000), __double2loint(x));
    double M, q;
    t_exp_parts<2, false>(xc, tab, M, q);
    return fma(-2.0, m_rcp(fma(M, q, 1.0)), 1.0);
}
__device__ __forceinline__ double t_nan_inject(double z, int nanmax) {
    return (nanmax > 0x7ff00000) ? __hiloint2double(0x7ff80000, 0) : z;
}

// softplus(x) = log(1 + exp(x)) (naive form of reference src/neural-network.jl:13-15, overflow to Inf included) and
// d = 1 + exp(x) as the reference rounds it; the adjoint uses d softplus/dx = 1 - 1/d.
__device__ __forceinline__ void t_softplus_d(double x, const double* __restrict__ tab, double& sp_out, double& d_out) {
    const double xc = m_clamp(x, -40.0, 36.8);
    double M, q;
    t_exp_parts<1, false>(xc, tab, M, q);
    const double d = fma(M, q, 1.0);
    double sp = m_log_core(d);
    sp = (x > 36.8) ? x : sp;
    sp = (x > 709.782712893384) ? CUDART_INF : sp;
    sp_out = (x != x) ? x : sp;
    d_out = d;
}
// d softplus / dx = 1/(1+exp(-x)) = 1 - 1/(1+exp(x)); absolute error <= 1.2e-16
__device__ __forceinline__ double t_sigmoid(double x, const double* __restrict__ tab) {
    double M, q;
    t_exp_parts<1, false>(m_clamp(x, -40.0, 40.0), tab, M, q);
    return fma(-1.0, m_rcp(fma(M, q, 1.0)), 1.0);
}
// exp for the step controller (argument saturated at +-40: the controller clamps the result anyway)
__device__ __forceinline__ double t_exp_sat(double x, const double* __restrict__ tab) {
    double M, q;
    t_exp_parts<1, false>(m_clamp(x, -40.0, 40.0), tab, M, q);
    return M * q;
}
__device__ __forceinline__ double t_pow10(double x, const double* __restrict__ tab) {
    double M, q;
    t_exp_parts<1, true>(m_clamp(x * 2.302585092994046, -700.0, 700.0), tab, M, q);
    return M * q;
}

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
