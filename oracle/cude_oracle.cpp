// =====================================================================================
// cude_oracle.cpp — CPU restatement (FP64) of the c-peptide conditional-UDE loss path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA path in
// conditional_ude_b200/csrc.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
// `--impl reference` legs of bench.py may load it.  The product path never calls it and has
// no CPU fallback.
//
// PARITY STATUS.  The reference (Computational-Biology-TUe/conditional-ude) is pure Julia, has no tests and no
// golden vectors, and Julia is not installed here, so this restatement could not be run against the reference.
//   * PINNED by numbers the reference itself produced: the integrator core (Tsit5 tableau, error norm, PI
//     controller, Hairer initial step, dense-output saveat, scaled-SSE loss) reproduces all 25 stored training
//     losses and 25 stored validation losses of suppression/results/lambda=1.0.jld2 to 1.1e-10 relative
//     (tests/test_suppression_oracle.py); the c-peptide solve goes through a specialised copy of that core which is
//     bit-identical to it on the c-peptide problem (same test file).
//   * PARITY UNPINNED for the c-peptide-specific pieces (van Cauter kinetics, glucose interpolant, MLP layout,
//     input order): no stored loss value of the reference exists for them.  They are pinned only indirectly
//     (tests/test_oracle.py, tests/test_artifacts.py): the stored trained betas are stationary points of this loss,
//     the stored weights a population optimum, tight-tolerance agreement with an independent scipy DOP853 solve.
//
// What is restated (reference file:line):
//   van_cauter_parameters            src/c-peptide-models.jl:30-42
//   c_peptide_kinetics!              src/c-peptide-models.jl:7-14
//   conditional_production           src/c-peptide-models.jl:86-94   (covariate variant :96-104)
//   combine                          src/c-peptide-models.jl:108-114
//   CPeptideConditionalUDEModel ctor src/c-peptide-models.jl:170-194 (u0, tspan, t0)
//   chain / softplus (MLP)           src/neural-network.jl:13-15,42-58  (SimpleChains TurboDense
//                                    layout: per layer W[out x in] column-major, then bias)
//   loss (single / fixed-NN / pop.)  src/parameter-estimation.jl:56-68, 93-99, 126-140
//   loss_sigma                       src/parameter-estimation.jl:70-75, 101-109
//   likelihood_profile               src/likelihood-profiles.jl:4-17 (= loss on a beta grid)
// Third-party arithmetic that is NOT vendored in the reference and is restated from the
// published algorithms (Project.toml:36-57 gives compat ranges only, no lockfile):
//   OrdinaryDiffEq 6.89  `solve(prob; p, saveat, save_idxs=1)` with the default algorithm
//       (-> Tsit5 for this non-stiff problem), abstol=1e-6, reltol=1e-3, PI controller
//       (beta1=7/50, beta2=2/25, gamma=9/10, qmin=1/5, qmax=10, qoldinit=1e-4), Hairer initial
//       dt, dense-output `saveat`.  Upstream uses an approximate `fastpow` in the controller;
//       here exact pow() is used (documented deviation).
//   DataInterpolations 6  LinearInterpolation(u, t)(tau)
//   SimpleChains 0.4.7    TurboDense{true} forward pass
//   ForwardDiff           gradient semantics (mode 1 below)
//
// Gradient modes:
//   mode 0  "frozen primal": tangents are propagated through the discrete Tsit5 recursion with the
//           step sequence chosen by the primal values only.  = exact derivative of the discrete
//           solve the loss-only path performs; the CUDA discrete adjoint must equal this.
//   mode 1  "ForwardDiff twin": the error norm (and Hairer's initial dt) see value + partials
//           (DiffEqBase's ForwardDiff norm), and theta is processed in ForwardDiff chunks, each
//           chunk being a separate adaptive solve with its own step sequence — the semantic twin
//           of `OptimizationFunction(loss, AutoForwardDiff())`, src/parameter-estimation.jl:231.
//           Documentation of the gap only; not what the CUDA path computes.
// =====================================================================================
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int MAXP = 72;   // max partials carried by a Dual (suppression example: 67 weights + theta)
constexpr int MAXW = 16;   // max MLP layer width
constexpr int MAXOBS = 64;

// ---------------------------------------------------------------- dual numbers
struct Dual {
    double v;
    int n;  // number of live partials
    double d[MAXP];
};
static inline Dual dconst(double v, int n) { Dual r; r.v = v; r.n = n; for (int i = 0; i < n; ++i) r.d[i] = 0.0; return r; }

struct RealOps {
    typedef double T;
    int n = 0;
    static double val(double a) { return a; }
    double cst(double v) const { return v; }
    static double add(double a, double b) { return a + b; }
    static double sub(double a, double b) { return a - b; }
    static double mul(double a, double b) { return a * b; }
    static double div(double a, double b) { return a / b; }
    static double smul(double s, double a) { return s * a; }
    static double sadd(double s, double a) { return s + a; }
    static double exp_(double a) { return std::exp(a); }
    static double log_(double a) { return std::log(a); }
    static double tanh_(double a) { return std::tanh(a); }
    static double abs_(double a) { return std::fabs(a); }
    static double max_(double a, double b) { return (a < b) ? b : a; }
    static double sse(double a) { return a * a; }  // contribution to the squared norm
    static int width() { return 1; }
};

struct DualOps {
    typedef Dual T;
    int n = 0;
    bool norm_partials = false;
    static double val(const Dual& a) { return a.v; }
    Dual cst(double v) const { return dconst(v, n); }
    static Dual add(const Dual& a, const Dual& b) { Dual r; r.n = a.n; r.v = a.v + b.v; for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
    static Dual sub(const Dual& a, const Dual& b) { Dual r; r.n = a.n; r.v = a.v - b.v; for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
    static Dual mul(const Dual& a, const Dual& b) { Dual r; r.n = a.n; r.v = a.v * b.v; for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
    static Dual div(const Dual& a, const Dual& b) { Dual r; r.n = a.n; r.v = a.v / b.v; double ib = 1.0 / b.v; for (int i = 0; i < a.n; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
    static Dual smul(double s, const Dual& a) { Dual r; r.n = a.n; r.v = s * a.v; for (int i = 0; i < a.n; ++i) r.d[i] = s * a.d[i]; return r; }
    static Dual sadd(double s, const Dual& a) { Dual r = a; r.v = s + a.v; return r; }
    static Dual chain1(const Dual& a, double f, double df) { Dual r; r.n = a.n; r.v = f; for (int i = 0; i < a.n; ++i) r.d[i] = df * a.d[i]; return r; }
    static Dual exp_(const Dual& a) { double e = std::exp(a.v); return chain1(a, e, e); }
    static Dual log_(const Dual& a) { return chain1(a, std::log(a.v), 1.0 / a.v); }
    static Dual tanh_(const Dual& a) { double t = std::tanh(a.v); return chain1(a, t, 1.0 - t * t); }
    static Dual abs_(const Dual& a) { return (a.v < 0 || (a.v == 0 && std::signbit(a.v))) ? smul(-1.0, a) : a; }
    static Dual max_(const Dual& a, const Dual& b) { return (a.v < b.v) ? b : a; }
};

// ---------------------------------------------------------------- problem description
struct NetDesc { int n_in, depth, width; };
static int net_nparams(const NetDesc& nd) {
    int p = 0, in = nd.n_in;
    for (int l = 0; l < nd.depth; ++l) { p += nd.width * (in + 1); in = nd.width; }
    return p + (in + 1);
}

struct Indiv {
    int nk; const double* kt; const double* kg;   // glucose knots
    int nobs; const double* ot; const double* oy; // observations
    double k0, k1, k2, c0;
    double cov;                                   // covariate input (age) for n_in == 3
};

struct Opts { double abstol, reltol; int maxiters; };

struct Stats { int nacc, nrej, nrhs, retcode; };
// optional per-step trace (t, dt, EEst, accepted) for tests / debugging
struct Trace { double* buf; int cap; int n; };
static thread_local Trace* g_trace = nullptr;
enum { RET_SUCCESS = 0, RET_MAXITERS = 1, RET_DTMIN = 2, RET_UNSTABLE = 3 };

// ---------------------------------------------------------------- Tsit5 tableau
constexpr double C2 = 0.161, C3 = 0.327, C4 = 0.9, C5 = 0.9800255409045097;
constexpr double A21 = 0.161;
constexpr double A31 = -0.008480655492356989, A32 = 0.335480655492357;
constexpr double A41 = 2.8971530571054935, A42 = -6.359448489975075, A43 = 4.3622954328695815;
constexpr double A51 = 5.325864828439257, A52 = -11.748883564062828, A53 = 7.4955393428898365, A54 = -0.09249506636175525;
constexpr double A61 = 5.86145544294642, A62 = -12.92096931784711, A63 = 8.159367898576159, A64 = -0.071584973281401, A65 = -0.028269050394068383;
constexpr double A71 = 0.09646076681806523, A72 = 0.01, A73 = 0.4798896504144996, A74 = 1.379008574103742, A75 = -3.290069515436081, A76 = 2.324710524099774;
constexpr double BT1 = -0.00178001105222577714, BT2 = -0.0008164344596567469, BT3 = 0.007880878010261995, BT4 = -0.1447110071732629,
                 BT5 = 0.5823571654525552, BT6 = -0.45808210592918697, BT7 = 0.015151515151515152;
constexpr double R11 = 1.0, R12 = -2.763706197274826, R13 = 2.9132554618219126, R14 = -1.0530884977290216;
constexpr double R22 = 0.13169999999999998, R23 = -0.2234, R24 = 0.1017;
constexpr double R32 = 3.9302962368947516, R33 = -5.941033872131505, R34 = 2.490627285651253;
constexpr double R42 = -12.411077166933676, R43 = 30.33818863028232, R44 = -16.548102889244902;
constexpr double R52 = 37.50931341651104, R53 = -88.1789048947664, R54 = 47.37952196281928;
constexpr double R62 = -27.896526289197286, R63 = 65.09189467479366, R64 = -34.87065786149661;
constexpr double R72 = 1.5, R73 = -4.0, R74 = 2.5;

// PI controller (OrdinaryDiffEq default for Tsit5)
constexpr double BETA1 = 7.0 / 50.0, BETA2 = 2.0 / 25.0, GAMMA = 9.0 / 10.0, QMIN = 1.0 / 5.0, QMAX = 10.0, QOLDINIT = 1e-4;

// ---------------------------------------------------------------- glucose interpolant
// DataInterpolations v6 LinearInterpolation: idx = clamp(searchsortedlast(t, tau), 1, n-1);
// u[idx] + (u[idx+1]-u[idx])/(t[idx+1]-t[idx]) * (tau - t[idx]).   (c-peptide-models.jl:181, :89)
static inline double glucose_at(const Indiv& I, double tau) {
    int idx = 0;  // 0-based segment
    while (idx + 1 < I.nk - 1 && I.kt[idx + 1] <= tau) ++idx;
    double slope = (I.kg[idx + 1] - I.kg[idx]) / (I.kt[idx + 1] - I.kt[idx]);
    return I.kg[idx] + slope * (tau - I.kt[idx]);
}

// ---------------------------------------------------------------- MLP (neural-network.jl:42-58)
template <class O>
static typename O::T mlp_forward(const O& ops, const NetDesc& nd, const typename O::T* p, const typename O::T* x) {
    typedef typename O::T T;
    T a[MAXW], b[MAXW];
    int in = nd.n_in;
    for (int i = 0; i < in; ++i) a[i] = x[i];
    int off = 0;
    for (int l = 0; l < nd.depth; ++l) {
        const int out = nd.width;
        for (int j = 0; j < out; ++j) {
            T z = p[off + out * in + j];  // bias
            for (int i = 0; i < in; ++i) z = O::add(z, O::mul(p[off + i * out + j], a[i]));
            b[j] = O::tanh_(z);
        }
        off += out * (in + 1);
        in = out;
        for (int j = 0; j < out; ++j) a[j] = b[j];
    }
    T z = p[off + in];
    for (int i = 0; i < in; ++i) z = O::add(z, O::mul(p[off + i], a[i]));
    // softplus(x) = log(1 + exp(x)), the naive form of neural-network.jl:13-15
    return O::log_(O::sadd(1.0, O::exp_(z)));
}

// ---------------------------------------------------------------- RHS
template <class O>
struct Rhs {
    typedef typename O::T T;
    const O& ops; const NetDesc& nd; const Indiv& I; const T* neural;
    T beta, nn0; double g0;
    int* nrhs;
    Rhs(const O& o, const NetDesc& n, const Indiv& i, const T* p, const T& cond, int* cnt)
        : ops(o), nd(n), I(i), neural(p), nrhs(cnt) {
        beta = O::exp_(cond);                 // c-peptide-models.jl:90
        g0 = glucose_at(I, I.kt[0]);          // glucose(t0), :89
        nn0 = net(0.0);                       // network([0; beta]) — same value at every call, hoisted
    }
    T net(double dG) const {
        T x[3]; x[0] = ops.cst(dG); x[1] = beta; x[2] = ops.cst(I.cov);
        return mlp_forward(ops, nd, neural, x);
    }
    void operator()(const T* u, double t, T* du) const {
        ++*nrhs;
        // c_peptide_kinetics!, c-peptide-models.jl:7-14
        T a = O::smul(-(I.k0 + I.k2), u[0]);
        T b = O::smul(I.k1, u[1]);
        T kin1 = O::sadd(I.k0 * I.c0, O::add(a, b));
        T kin2 = O::add(O::smul(-I.k1, u[1]), O::smul(I.k2, u[0]));
        // conditional_production, :86-94
        double dG = glucose_at(I, t) - g0;
        T prod = O::sub(net(dG), nn0);
        du[0] = O::add(kin1, prod);           // combine, :108-114
        du[1] = kin2;
    }
};

// ODE_DEFAULT_NORM over a 2-vector: sqrt(sum(abs2)/length); with duals the ForwardDiff extension
// counts value^2 + sum(partials^2) over length*(1+npartials).
static inline double norm2(const RealOps&, const double* x) { return std::sqrt((x[0] * x[0] + x[1] * x[1]) / 2.0); }
static inline double norm2(const DualOps& o, const Dual* x) {
    if (!o.norm_partials) return std::sqrt((x[0].v * x[0].v + x[1].v * x[1].v) / 2.0);
    double s = 0.0;
    for (int j = 0; j < 2; ++j) { s += x[j].v * x[j].v; for (int i = 0; i < x[j].n; ++i) s += x[j].d[i] * x[j].d[i]; }
    return std::sqrt(s / (2.0 * (1 + o.n)));
}
static inline double scalar_norm(const RealOps&, double x) { return std::fabs(x); }
static inline double scalar_norm(const DualOps& o, const Dual& x) {
    if (!o.norm_partials) return std::fabs(x.v);
    double s = x.v * x.v; for (int i = 0; i < x.n; ++i) s += x.d[i] * x.d[i]; return std::sqrt(s);
}

// ---------------------------------------------------------------- adaptive Tsit5 solve + SSE
// solve(model.problem, p=theta, saveat=timepoints, save_idxs=1) + sum(abs2, sol - data)
// (parameter-estimation.jl:59-67).  Returns +Inf on solver failure (:61-64).
template <class O>
static typename O::T solve_sse(const O& ops, const NetDesc& nd, const Indiv& I, const typename O::T* neural,
                               const typename O::T& cond, const Opts& opt, Stats* st, double* yhat_out) {
    typedef typename O::T T;
    int nrhs = 0, nacc = 0, nrej = 0;
    Rhs<O> f(ops, nd, I, neural, cond, &nrhs);
    const double t0 = I.kt[0], tend = I.kt[I.nk - 1];
    const double dtmax = tend - t0;
    const double dtmin = std::max(std::nextafter(std::fabs(t0), INFINITY) - std::fabs(t0),
                                  std::nextafter(std::fabs(tend), INFINITY) - std::fabs(tend));
    T u[2] = { ops.cst(I.c0), ops.cst((I.k2 / I.k1) * I.c0) };   // c-peptide-models.jl:185
    T k1[2], k2[2], k3[2], k4[2], k5[2], k6[2], k7[2], un[2], g[2];
    T sse = ops.cst(0.0);
    int retcode = RET_SUCCESS;

    auto finish = [&](int rc) {
        st->nacc = nacc; st->nrej = nrej; st->nrhs = nrhs; st->retcode = rc;
    };
    auto nonfinite = [&](const T* x) { return !(std::isfinite(O::val(x[0])) && std::isfinite(O::val(x[1]))); };

    int iobs = 0;
    auto record = [&](const T& y, int k) {
        if (yhat_out) yhat_out[k] = O::val(y);
        T r = O::sadd(-I.oy[k], y);
        sse = O::add(sse, O::mul(r, r));
    };
    while (iobs < I.nobs && I.ot[iobs] <= t0) { record(u[0], iobs); ++iobs; }   // save_start

    f(u, t0, k1);
    // ---- Hairer initial step (ode_determine_initdt) ----
    double dt;
    {
        double sk[2]; T tmp[2];
        for (int j = 0; j < 2; ++j) sk[j] = opt.abstol + scalar_norm(ops, u[j]) * opt.reltol;
        for (int j = 0; j < 2; ++j) tmp[j] = O::smul(1.0 / sk[j], u[j]);
        double d0 = norm2(ops, tmp);
        for (int j = 0; j < 2; ++j) tmp[j] = O::smul(1.0 / sk[j], k1[j]);
        double d1 = norm2(ops, tmp);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
        dt0 = std::min(dt0, dtmax);
        T u1[2], f1[2];
        for (int j = 0; j < 2; ++j) u1[j] = O::add(u[j], O::smul(dt0, k1[j]));
        f(u1, t0 + dt0, f1);
        for (int j = 0; j < 2; ++j) tmp[j] = O::smul(1.0 / sk[j], O::sub(f1[j], k1[j]));
        double d2 = norm2(ops, tmp) / dt0;
        double dm = std::max(d1, d2);
        double dt1 = (dm <= 1e-15) ? std::max(1e-6, dt0 * 1e-3) : std::pow(10.0, -(2.0 + std::log10(dm)) / 5.0);
        dt = std::max(dtmin, std::min(std::min(100.0 * dt0, dt1), dtmax));
    }
    if (!std::isfinite(dt) || nonfinite(k1)) { finish(RET_UNSTABLE); return ops.cst(INFINITY); }

    double t = t0, qold = QOLDINIT;
    int iter = 0;
    while (t < tend) {
        ++iter;
        if (iter > opt.maxiters) { retcode = RET_MAXITERS; break; }
        dt = std::min(dt, tend - t);          // modify_dt_for_tstops!
        if (!(dt > dtmin) ) { retcode = std::isnan(dt) ? RET_UNSTABLE : RET_DTMIN; break; }
        // ---- Tsit5 stages ----
        for (int j = 0; j < 2; ++j) g[j] = O::add(u[j], O::smul(dt * A21, k1[j]));
        f(g, t + C2 * dt, k2);
        for (int j = 0; j < 2; ++j) g[j] = O::add(u[j], O::smul(dt, O::add(O::smul(A31, k1[j]), O::smul(A32, k2[j]))));
        f(g, t + C3 * dt, k3);
        for (int j = 0; j < 2; ++j) g[j] = O::add(u[j], O::smul(dt, O::add(O::add(O::smul(A41, k1[j]), O::smul(A42, k2[j])), O::smul(A43, k3[j]))));
        f(g, t + C4 * dt, k4);
        for (int j = 0; j < 2; ++j) g[j] = O::add(u[j], O::smul(dt, O::add(O::add(O::add(O::smul(A51, k1[j]), O::smul(A52, k2[j])), O::smul(A53, k3[j])), O::smul(A54, k4[j]))));
        f(g, t + C5 * dt, k5);
        for (int j = 0; j < 2; ++j) g[j] = O::add(u[j], O::smul(dt, O::add(O::add(O::add(O::add(O::smul(A61, k1[j]), O::smul(A62, k2[j])), O::smul(A63, k3[j])), O::smul(A64, k4[j])), O::smul(A65, k5[j]))));
        f(g, t + dt, k6);
        for (int j = 0; j < 2; ++j) un[j] = O::add(u[j], O::smul(dt, O::add(O::add(O::add(O::add(O::add(O::smul(A71, k1[j]), O::smul(A72, k2[j])), O::smul(A73, k3[j])), O::smul(A74, k4[j])), O::smul(A75, k5[j])), O::smul(A76, k6[j]))));
        f(un, t + dt, k7);
        // ---- error estimate ----
        T res[2];
        for (int j = 0; j < 2; ++j) {
            T ut = O::smul(dt, O::add(O::add(O::add(O::add(O::add(O::add(O::smul(BT1, k1[j]), O::smul(BT2, k2[j])), O::smul(BT3, k3[j])), O::smul(BT4, k4[j])), O::smul(BT5, k5[j])), O::smul(BT6, k6[j])), O::smul(BT7, k7[j])));
            T den = O::sadd(opt.abstol, O::smul(opt.reltol, O::max_(O::abs_(u[j]), O::abs_(un[j]))));
            res[j] = O::div(ut, den);
        }
        double EEst = norm2(ops, res);
        if (std::isnan(EEst) || nonfinite(un)) { retcode = RET_UNSTABLE; break; }
        // ---- PI controller ----
        double q, q11 = 0.0;
        if (EEst == 0.0) q = 1.0 / QMAX;
        else {
            q11 = std::pow(EEst, BETA1);
            q = q11 / std::pow(qold, BETA2);
            q = std::max(1.0 / QMAX, std::min(1.0 / QMIN, q / GAMMA));
        }
        if (g_trace && g_trace->n < g_trace->cap) {
            double* r = g_trace->buf + 4 * g_trace->n++;
            r[0] = t; r[1] = dt; r[2] = EEst; r[3] = (EEst <= 1.0) ? 1.0 : 0.0;
        }
        if (EEst <= 1.0) {
            ++nacc;
            double tnew = t + dt;
            // fixed_t_for_floatingpoint_error!: 100*eps(max(t, tstop)) with t < tstop
            const double m = std::fabs(tend);
            if (std::fabs(tnew - tend) < 100.0 * (std::nextafter(m, INFINITY) - m)) tnew = tend;
            // saveat by dense output (savevalues!): points in (t, tnew]
            while (iobs < I.nobs && I.ot[iobs] <= tnew) {
                double ts = I.ot[iobs];
                if (ts == tnew) record(un[0], iobs);
                else {
                    double th = (ts - t) / dt;
                    double b1 = th * (R11 + th * (R12 + th * (R13 + th * R14)));
                    double th2 = th * th;
                    double b2 = th2 * (R22 + th * (R23 + th * R24));
                    double b3 = th2 * (R32 + th * (R33 + th * R34));
                    double b4 = th2 * (R42 + th * (R43 + th * R44));
                    double b5 = th2 * (R52 + th * (R53 + th * R54));
                    double b6 = th2 * (R62 + th * (R63 + th * R64));
                    double b7 = th2 * (R72 + th * (R73 + th * R74));
                    T s = O::add(O::add(O::add(O::add(O::add(O::add(O::smul(b1, k1[0]), O::smul(b2, k2[0])), O::smul(b3, k3[0])), O::smul(b4, k4[0])), O::smul(b5, k5[0])), O::smul(b6, k6[0])), O::smul(b7, k7[0]));
                    record(O::add(u[0], O::smul(dt, s)), iobs);
                }
                ++iobs;
            }
            qold = std::max(EEst, QOLDINIT);
            dt = std::min(dt / q, dtmax);
            t = tnew;
            for (int j = 0; j < 2; ++j) { u[j] = un[j]; k1[j] = k7[j]; }   // FSAL
        } else {
            ++nrej;
            dt = dt / std::min(1.0 / QMIN, q11 / GAMMA);
        }
    }
    if (retcode == RET_SUCCESS && iobs < I.nobs) {
        // observation times beyond tend are not produced by saveat; treated as a failure
        retcode = RET_UNSTABLE;
    }
    finish(retcode);
    if (retcode != RET_SUCCESS) return ops.cst(INFINITY);
    return sse;
}


// =====================================================================================
// Suppression example (second variant of the same pattern), reference suppression/src/suppression_model.jl:
//   ude_lsup!            :88-95   u_hat = network([u; exp.(theta_i)], neural)[1]
//                                 du1 = -p1 u1;  du2 = p1 u1 - u_hat;  du3 = u_hat - p3 u3      (p_true = [0.4, 0.9, 0.3])
//   neural_network_model :78-86   input_dims=4 -> `depth` tanh layers of `width` -> 1 softplus (suppression.jl:18: 5 x 3)
//   suppression_loss     :117-130 solve(ensemble, Tsit5(), saveat=timepoints) per individual with u0 = data[:,1,i];
//                                 sum(abs2, (sims - data) ./ scale) / N + lambda * sum(abs2, neural)
// Here: the per-trajectory scaled SSE and its gradient; the mean over individuals and the ridge term are
// assembled by the caller.  Explicit Tsit5 at OrdinaryDiffEq's default tolerances, all 3 states saved.
// =====================================================================================
struct SupIndiv {
    double u0[3];
    int nobs; const double* ot;      // shared time grid
    const double* y;                 // [nobs][3] observations of this individual (state fastest)
    double p1, p3;
    double scale[3];
    double t0, tend;
};

template <class O>
struct SupRhs {
    typedef typename O::T T;
    const O& ops; const NetDesc& nd; const SupIndiv& I; const T* neural; T etheta; int* nrhs;
    SupRhs(const O& o, const NetDesc& n, const SupIndiv& i, const T* p, const T& theta, int* cnt)
        : ops(o), nd(n), I(i), neural(p), nrhs(cnt) { etheta = O::exp_(theta); }
    void operator()(const T* u, double, T* du) const {
        ++*nrhs;
        T x[4] = { u[0], u[1], u[2], etheta };
        T uhat = mlp_forward(ops, nd, neural, x);
        du[0] = O::smul(-I.p1, u[0]);
        du[1] = O::sub(O::smul(I.p1, u[0]), uhat);
        du[2] = O::sub(uhat, O::smul(I.p3, u[2]));
    }
};

template <class O>
static double normD(const O& ops, const typename O::T* x, int D) {
    double s = 0.0;
    int cnt = 0;
    for (int j = 0; j < D; ++j) {
        s += O::val(x[j]) * O::val(x[j]);
        ++cnt;
    }
    (void)ops;
    return std::sqrt(s / cnt);
}

// generic D-state adaptive Tsit5 + scaled SSE over all states (frozen-primal norm: values only)
template <class O, int D, class F>
static typename O::T solve_scaled_sse(const O& ops, const F& f, const double* u0v, double t0, double tend,
                                      int nobs, const double* ot, const double* y, const double* scale,
                                      const Opts& opt, Stats* st, double* yhat_out, int* nrhs, int dobs = D) {
    typedef typename O::T T;
    int nacc = 0, nrej = 0;
    const double dtmax = tend - t0;
    const double dtmin = std::max(std::nextafter(std::fabs(t0), INFINITY) - std::fabs(t0),
                                  std::nextafter(std::fabs(tend), INFINITY) - std::fabs(tend));
    T u[D], k1[D], k2[D], k3[D], k4[D], k5[D], k6[D], k7[D], un[D], g[D];
    for (int j = 0; j < D; ++j) u[j] = ops.cst(u0v[j]);
    T sse = ops.cst(0.0);
    int retcode = RET_SUCCESS, iobs = 0;
    auto record = [&](const T* yv, int k) {
        for (int j = 0; j < dobs; ++j) {          // the first `dobs` states are observed
            if (yhat_out) yhat_out[k * dobs + j] = O::val(yv[j]);
            T r = O::smul(1.0 / scale[j], O::sadd(-y[k * dobs + j], yv[j]));
            sse = O::add(sse, O::mul(r, r));
        }
    };
    auto finish = [&](int rc) { st->nacc = nacc; st->nrej = nrej; st->nrhs = *nrhs; st->retcode = rc; };
    auto nonfinite = [&](const T* x) { for (int j = 0; j < D; ++j) if (!std::isfinite(O::val(x[j]))) return true; return false; };
    while (iobs < nobs && ot[iobs] <= t0) { record(u, iobs); ++iobs; }
    f(u, t0, k1);
    double dt;
    {
        double sk[D]; T tmp[D];
        for (int j = 0; j < D; ++j) sk[j] = opt.abstol + std::fabs(O::val(u[j])) * opt.reltol;
        for (int j = 0; j < D; ++j) tmp[j] = O::smul(1.0 / sk[j], u[j]);
        double d0 = normD(ops, tmp, D);
        for (int j = 0; j < D; ++j) tmp[j] = O::smul(1.0 / sk[j], k1[j]);
        double d1 = normD(ops, tmp, D);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
        dt0 = std::min(dt0, dtmax);
        T u1[D], f1[D];
        for (int j = 0; j < D; ++j) u1[j] = O::add(u[j], O::smul(dt0, k1[j]));
        f(u1, t0 + dt0, f1);
        for (int j = 0; j < D; ++j) tmp[j] = O::smul(1.0 / sk[j], O::sub(f1[j], k1[j]));
        double d2 = normD(ops, tmp, D) / dt0;
        double dm = std::max(d1, d2);
        double dt1 = (dm <= 1e-15) ? std::max(1e-6, dt0 * 1e-3) : std::pow(10.0, -(2.0 + std::log10(dm)) / 5.0);
        dt = std::max(dtmin, std::min(std::min(100.0 * dt0, dt1), dtmax));
    }
    if (!std::isfinite(dt) || nonfinite(k1)) { finish(RET_UNSTABLE); return ops.cst(INFINITY); }
    double t = t0, qold = QOLDINIT;
    int iter = 0;
    const double* A[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    (void)A;
    while (t < tend) {
        if (++iter > opt.maxiters) { retcode = RET_MAXITERS; break; }
        dt = std::min(dt, tend - t);
        if (!(dt > dtmin)) { retcode = std::isnan(dt) ? RET_UNSTABLE : RET_DTMIN; break; }
        for (int j = 0; j < D; ++j) g[j] = O::add(u[j], O::smul(dt * A21, k1[j]));
        f(g, t + C2 * dt, k2);
        for (int j = 0; j < D; ++j) g[j] = O::add(u[j], O::smul(dt, O::add(O::smul(A31, k1[j]), O::smul(A32, k2[j]))));
        f(g, t + C3 * dt, k3);
        for (int j = 0; j < D; ++j) g[j] = O::add(u[j], O::smul(dt, O::add(O::add(O::smul(A41, k1[j]), O::smul(A42, k2[j])), O::smul(A43, k3[j]))));
        f(g, t + C4 * dt, k4);
        for (int j = 0; j < D; ++j) g[j] = O::add(u[j], O::smul(dt, O::add(O::add(O::add(O::smul(A51, k1[j]), O::smul(A52, k2[j])), O::smul(A53, k3[j])), O::smul(A54, k4[j]))));
        f(g, t + C5 * dt, k5);
        for (int j = 0; j < D; ++j) g[j] = O::add(u[j], O::smul(dt, O::add(O::add(O::add(O::add(O::smul(A61, k1[j]), O::smul(A62, k2[j])), O::smul(A63, k3[j])), O::smul(A64, k4[j])), O::smul(A65, k5[j]))));
        f(g, t + dt, k6);
        for (int j = 0; j < D; ++j) un[j] = O::add(u[j], O::smul(dt, O::add(O::add(O::add(O::add(O::add(O::smul(A71, k1[j]), O::smul(A72, k2[j])), O::smul(A73, k3[j])), O::smul(A74, k4[j])), O::smul(A75, k5[j])), O::smul(A76, k6[j]))));
        f(un, t + dt, k7);
        T res[D];
        for (int j = 0; j < D; ++j) {
            T ut = O::smul(dt, O::add(O::add(O::add(O::add(O::add(O::add(O::smul(BT1, k1[j]), O::smul(BT2, k2[j])), O::smul(BT3, k3[j])), O::smul(BT4, k4[j])), O::smul(BT5, k5[j])), O::smul(BT6, k6[j])), O::smul(BT7, k7[j])));
            T den = O::sadd(opt.abstol, O::smul(opt.reltol, O::max_(O::abs_(u[j]), O::abs_(un[j]))));
            res[j] = O::div(ut, den);
        }
        double EEst = normD(ops, res, D);
        if (std::isnan(EEst) || nonfinite(un)) { retcode = RET_UNSTABLE; break; }
        if (g_trace && g_trace->n < g_trace->cap) {
            double* r = g_trace->buf + 4 * g_trace->n++;
            r[0] = t; r[1] = dt; r[2] = EEst; r[3] = (EEst <= 1.0) ? 1.0 : 0.0;
        }
        double q, q11 = 0.0;
        if (EEst == 0.0) q = 1.0 / QMAX;
        else {
            q11 = std::pow(EEst, BETA1);
            q = q11 / std::pow(qold, BETA2);
            q = std::max(1.0 / QMAX, std::min(1.0 / QMIN, q / GAMMA));
        }
        if (EEst <= 1.0) {
            ++nacc;
            double tnew = t + dt;
            const double m = std::fabs(tend);
            if (std::fabs(tnew - tend) < 100.0 * (std::nextafter(m, INFINITY) - m)) tnew = tend;
            while (iobs < nobs && ot[iobs] <= tnew) {
                double ts = ot[iobs];
                if (ts == tnew) record(un, iobs);
                else {
                    double th = (ts - t) / dt, th2 = th * th;
                    double b1 = th * (R11 + th * (R12 + th * (R13 + th * R14)));
                    double b2 = th2 * (R22 + th * (R23 + th * R24)), b3 = th2 * (R32 + th * (R33 + th * R34));
                    double b4 = th2 * (R42 + th * (R43 + th * R44)), b5 = th2 * (R52 + th * (R53 + th * R54));
                    double b6 = th2 * (R62 + th * (R63 + th * R64)), b7 = th2 * (R72 + th * (R73 + th * R74));
                    T yv[D];
                    for (int j = 0; j < D; ++j) {
                        T sdo = O::add(O::add(O::add(O::add(O::add(O::add(O::smul(b1, k1[j]), O::smul(b2, k2[j])), O::smul(b3, k3[j])), O::smul(b4, k4[j])), O::smul(b5, k5[j])), O::smul(b6, k6[j])), O::smul(b7, k7[j]));
                        yv[j] = O::add(u[j], O::smul(dt, sdo));
                    }
                    record(yv, iobs);
                }
                ++iobs;
            }
            qold = std::max(EEst, QOLDINIT);
            dt = std::min(dt / q, dtmax);
            t = tnew;
            for (int j = 0; j < D; ++j) { u[j] = un[j]; k1[j] = k7[j]; }
        } else {
            ++nrej;
            dt = dt / std::min(1.0 / QMIN, q11 / GAMMA);
        }
    }
    if (retcode == RET_SUCCESS && iobs < nobs) retcode = RET_UNSTABLE;
    finish(retcode);
    if (retcode != RET_SUCCESS) return ops.cst(INFINITY);
    return sse;
}

struct Pop {
    int n_ind, max_knots, max_obs;
    const int* n_knots; const double* knot_t; const double* knot_g;
    const int* n_obs; const double* obs_t; const double* obs_y;
    const double* kin;   // [n_ind x 4] row-major: k0,k1,k2,c0
    const double* cov;   // [n_ind] or null
    Indiv get(int i) const {
        Indiv I;
        I.nk = n_knots[i]; I.kt = knot_t + (size_t)i * max_knots; I.kg = knot_g + (size_t)i * max_knots;
        I.nobs = n_obs[i]; I.ot = obs_t + (size_t)i * max_obs; I.oy = obs_y + (size_t)i * max_obs;
        I.k0 = kin[4 * (size_t)i + 0]; I.k1 = kin[4 * (size_t)i + 1]; I.k2 = kin[4 * (size_t)i + 2]; I.c0 = kin[4 * (size_t)i + 3];
        I.cov = cov ? cov[i] : 0.0;
        return I;
    }
};

// ForwardDiff.pickchunksize (DEFAULT_CHUNK_THRESHOLD = 12)
static int pickchunksize(int n) {
    if (n <= 12) return n;
    int nchunks = (n + 11) / 12;
    return (n + nchunks - 1) / nchunks;
}

}  // namespace

extern "C" {

struct cude_oracle_pop {
    int n_ind, max_knots, max_obs;
    const int* n_knots; const double* knot_t; const double* knot_g;   // [n_ind x max_knots] row-major
    const int* n_obs; const double* obs_t; const double* obs_y;       // [n_ind x max_obs] row-major
    const double* kin;                                                // [n_ind x 4]: k0,k1,k2,c0
    const double* cov;                                                // [n_ind] covariate (age) or NULL
};

// van_cauter_parameters(age, t2dm), c-peptide-models.jl:30-42
void cude_oracle_van_cauter(double age, int t2dm, double* k0, double* k1, double* k2) {
    const double ln2 = std::log(2.0);
    double short_half_life = t2dm ? 4.52 : 4.95;
    double fraction = t2dm ? 0.78 : 0.76;
    double long_half_life = 0.14 * age + 29.2;
    *k1 = fraction * (ln2 / long_half_life) + (1 - fraction) * (ln2 / short_half_life);
    *k0 = (ln2 / short_half_life) * (ln2 / long_half_life) / *k1;
    *k2 = (ln2 / short_half_life) + (ln2 / long_half_life) - *k0 - *k1;
}

int cude_oracle_nparams(int n_in, int depth, int width) { NetDesc nd{n_in, depth, width}; return net_nparams(nd); }

// MLP forward for unit tests: out = chain(x, p)
double cude_oracle_mlp(int n_in, int depth, int width, const double* p, const double* x) {
    NetDesc nd{n_in, depth, width}; RealOps ops;
    return mlp_forward(ops, nd, p, x);
}

double cude_oracle_glucose(int nk, const double* kt, const double* kg, double tau) {
    Indiv I; I.nk = nk; I.kt = kt; I.kg = kg; return glucose_at(I, tau);
}

// Per-trajectory evaluation.  Trajectory (i, s): individual i, start s.
//   neural : start s uses neural + s*neural_stride   (neural_stride = 0: one shared network)
//   cond   : [n_ind x n_starts] column-major (individual fastest), cond[i + n_ind*s]
// Outputs (any may be NULL):
//   sse[i + n_ind*s], yhat[(i + n_ind*s)*max_obs + k], stats[(i + n_ind*s)*4 + {nacc,nrej,nrhs,retcode}]
//   grad_mode < 0: no gradient.  0: frozen-primal tangents.  1: ForwardDiff twin (per trajectory,
//   theta = [neural; cond] chunked).
//   g_neural_traj[(i + n_ind*s)*P + p] = d sse / d neural_p,   g_cond[i + n_ind*s] = d sse / d cond
int cude_oracle_eval(const cude_oracle_pop* cp, int n_in, int depth, int width,
                     int n_starts, const double* neural, long neural_stride, const double* cond,
                     double abstol, double reltol, int maxiters, int grad_mode, int n_threads,
                     double* sse, double* yhat, int* stats, double* g_neural_traj, double* g_cond) {
    Pop pop{cp->n_ind, cp->max_knots, cp->max_obs, cp->n_knots, cp->knot_t, cp->knot_g, cp->n_obs, cp->obs_t, cp->obs_y, cp->kin, cp->cov};
    NetDesc nd{n_in, depth, width};
    const int P = net_nparams(nd);
    if (P + 1 > MAXP || width > MAXW || n_in > 3 || pop.max_obs > MAXOBS) return -1;
    Opts opt{abstol, reltol, maxiters};
    const long ntraj = (long)pop.n_ind * n_starts;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (long j = 0; j < ntraj; ++j) {
        const int i = (int)(j % pop.n_ind);
        const long s = j / pop.n_ind;
        Indiv I = pop.get(i);
        const double* p = neural + s * neural_stride;
        Stats st;
        double* yh = yhat ? yhat + j * pop.max_obs : nullptr;
        if (grad_mode < 0) {
            RealOps ops;
            double v = solve_sse(ops, nd, I, p, cond[j], opt, &st, yh);
            if (sse) sse[j] = v;
        } else if (grad_mode == 0) {
            DualOps ops; ops.n = P + 1; ops.norm_partials = false;
            std::vector<Dual> dp(P);
            for (int a = 0; a < P; ++a) { dp[a] = dconst(p[a], P + 1); dp[a].d[a] = 1.0; }
            Dual dc = dconst(cond[j], P + 1); dc.d[P] = 1.0;
            Dual v = solve_sse(ops, nd, I, dp.data(), dc, opt, &st, yh);
            if (sse) sse[j] = v.v;
            const bool ok = std::isfinite(v.v);
            if (g_neural_traj) for (int a = 0; a < P; ++a) g_neural_traj[j * P + a] = ok ? v.d[a] : 0.0;
            if (g_cond) g_cond[j] = ok ? v.d[P] : 0.0;
        } else {
            // ForwardDiff chunk mode over theta = [neural(P); cond(1)]
            const int n = P + 1, N = pickchunksize(n);
            DualOps ops; ops.n = N; ops.norm_partials = true;
            std::vector<Dual> dp(P);
            double val = 0.0;
            for (int c0 = 0; c0 < n; c0 += N) {
                // the last chunk keeps N partials; unused seeds stay zero (ForwardDiff remainder handling)
                for (int a = 0; a < P; ++a) { dp[a] = dconst(p[a], N); if (a >= c0 && a < c0 + N) dp[a].d[a - c0] = 1.0; }
                Dual dc = dconst(cond[j], N); if (P >= c0 && P < c0 + N) dc.d[P - c0] = 1.0;
                Stats stc;
                Dual v = solve_sse(ops, nd, I, dp.data(), dc, opt, &stc, c0 == 0 ? yh : nullptr);
                if (c0 == 0) { st = stc; val = v.v; }
                const bool ok = std::isfinite(v.v);
                for (int a = c0; a < std::min(c0 + N, n); ++a) {
                    double gval = ok ? v.d[a - c0] : 0.0;
                    if (a < P) { if (g_neural_traj) g_neural_traj[j * P + a] = gval; }
                    else if (g_cond) g_cond[j] = gval;
                }
            }
            if (sse) sse[j] = val;
        }
        if (stats) { stats[j * 4 + 0] = st.nacc; stats[j * 4 + 1] = st.nrej; stats[j * 4 + 2] = st.nrhs; stats[j * 4 + 3] = st.retcode; }
    }
    return 0;
}

// Population loss of parameter-estimation.jl:126-140 for each start s:
//   loss[s] = mean_i sse(i, s)  (Inf if any trajectory failed), and with grad_mode 0 its gradient:
//   g_neural[s*P + p] = (1/N) sum_i d sse_i / d neural_p,   g_cond[i + N*s] = (1/N) d sse_i / d cond_i
// This is the function timed as the CPU baseline (OpenMP over trajectories).
int cude_oracle_population_loss(const cude_oracle_pop* cp, int n_in, int depth, int width,
                                int n_starts, const double* neural, long neural_stride, const double* cond,
                                double abstol, double reltol, int maxiters, int with_grad, int n_threads,
                                double* loss, double* g_neural, double* g_cond, long* counters /* nacc,nrej,nrhs sums */) {
    NetDesc nd{n_in, depth, width};
    const int P = net_nparams(nd);
    const int N = cp->n_ind;
    const long ntraj = (long)N * n_starts;
    std::vector<double> sse(ntraj), gn, gc;
    std::vector<int> st(ntraj * 4);
    if (with_grad) { gn.resize(ntraj * P); gc.resize(ntraj); }
    int rc = cude_oracle_eval(cp, n_in, depth, width, n_starts, neural, neural_stride, cond, abstol, reltol, maxiters,
                              with_grad ? 0 : -1, n_threads, sse.data(), nullptr, st.data(),
                              with_grad ? gn.data() : nullptr, with_grad ? gc.data() : nullptr);
    if (rc) return rc;
    long c0 = 0, c1 = 0, c2 = 0;
    for (long j = 0; j < ntraj; ++j) { c0 += st[j * 4]; c1 += st[j * 4 + 1]; c2 += st[j * 4 + 2]; }
    if (counters) { counters[0] = c0; counters[1] = c1; counters[2] = c2; }
    for (int s = 0; s < n_starts; ++s) {
        double acc = 0.0;
        for (int i = 0; i < N; ++i) acc += sse[(long)s * N + i];   // Inf propagates like :134-136
        loss[s] = acc / N;
        if (with_grad) {
            const bool ok = std::isfinite(acc);
            for (int a = 0; a < P; ++a) {
                double g = 0.0;
                for (int i = 0; i < N; ++i) g += gn[((long)s * N + i) * P + a];
                if (g_neural) g_neural[(long)s * P + a] = ok ? g / N : 0.0;
            }
            if (g_cond) for (int i = 0; i < N; ++i) g_cond[(long)s * N + i] = ok ? gc[(long)s * N + i] / N : 0.0;
        }
    }
    return 0;
}

// Step trace of one trajectory (individual i): rows (t, dt, EEst, accepted).  Returns the row count.
int cude_oracle_trace(const cude_oracle_pop* cp, int n_in, int depth, int width, int i, const double* neural, double cond,
                      double abstol, double reltol, int maxiters, double* rows, int cap, double* sse_out) {
    Pop pop{cp->n_ind, cp->max_knots, cp->max_obs, cp->n_knots, cp->knot_t, cp->knot_g, cp->n_obs, cp->obs_t, cp->obs_y, cp->kin, cp->cov};
    NetDesc nd{n_in, depth, width};
    Opts opt{abstol, reltol, maxiters};
    Trace tr{rows, cap, 0};
    g_trace = &tr;
    RealOps ops; Stats st;
    Indiv I = pop.get(i);
    double v = solve_sse(ops, nd, I, neural, cond, opt, &st, nullptr);
    g_trace = nullptr;
    if (sse_out) *sse_out = v;
    return tr.n;
}

// The c-peptide problem through the generic D-state core (the one pinned by the suppression artifacts): used by the
// tests to tie solve_sse (the specialised 2-state restatement) to the pinned core.  sse[i + n_ind*s].
int cude_oracle_eval_generic(const cude_oracle_pop* cp, int n_in, int depth, int width, int n_starts, const double* neural,
                             long neural_stride, const double* cond, double abstol, double reltol, int maxiters,
                             double* sse, int* stats) {
    Pop pop{cp->n_ind, cp->max_knots, cp->max_obs, cp->n_knots, cp->knot_t, cp->knot_g, cp->n_obs, cp->obs_t, cp->obs_y, cp->kin, cp->cov};
    NetDesc nd{n_in, depth, width};
    Opts opt{abstol, reltol, maxiters};
    const long ntraj = (long)pop.n_ind * n_starts;
    for (long j = 0; j < ntraj; ++j) {
        const int i = (int)(j % pop.n_ind);
        Indiv I = pop.get(i);
        RealOps ops; Stats st; int nrhs = 0;
        Rhs<RealOps> f(ops, nd, I, neural + (j / pop.n_ind) * neural_stride, cond[j], &nrhs);
        const double u0[2] = { I.c0, (I.k2 / I.k1) * I.c0 };
        const double one[2] = { 1.0, 1.0 };
        sse[j] = solve_scaled_sse<RealOps, 2>(ops, f, u0, I.kt[0], I.kt[I.nk - 1], I.nobs, I.ot, I.oy, one, opt, &st, nullptr, &nrhs, 1);
        if (stats) { stats[j * 4] = st.nacc; stats[j * 4 + 1] = st.nrej; stats[j * 4 + 2] = st.nrhs; stats[j * 4 + 3] = st.retcode; }
    }
    return 0;
}

// ---- suppression example ----
// data: Julia layout data[state + 3*(k + n_obs*i)] (3 x n_obs x n_ind), theta[i + n_ind*s], neural + s*neural_stride.
// scale[3]: mean over individuals of the per-state maximum over time (suppression_model.jl:125); p_true = {p1, p2, p3}.
// Outputs per trajectory j = i + n_ind*s: sse (scaled), g_neural_traj[j*P + p], g_theta[j], stats[j*4..], yhat[j*n_obs*3..].
int cude_oracle_sup_eval(int n_ind, int n_obs, const double* obs_t, const double* data, const double* p_true,
                         const double* scale, double t0, double tend, int depth, int width,
                         int n_starts, const double* neural, long neural_stride, const double* theta,
                         double abstol, double reltol, int maxiters, int with_grad, int n_threads,
                         double* sse, double* yhat, int* stats, double* g_neural_traj, double* g_theta) {
    NetDesc nd{4, depth, width};
    const int P = net_nparams(nd);
    if (P + 1 > MAXP || width > MAXW) return -1;
    Opts opt{abstol, reltol, maxiters};
    const long ntraj = (long)n_ind * n_starts;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
    for (long j = 0; j < ntraj; ++j) {
        const int i = (int)(j % n_ind);
        const long s = j / n_ind;
        SupIndiv I;
        I.nobs = n_obs; I.ot = obs_t; I.y = data + (size_t)i * n_obs * 3;
        for (int q = 0; q < 3; ++q) { I.u0[q] = I.y[q]; I.scale[q] = scale[q]; }   // u0 = data[:,1,i], :99-104
        I.p1 = p_true[0]; I.p3 = p_true[2]; I.t0 = t0; I.tend = tend;
        const double* p = neural + s * neural_stride;
        Stats st; int nrhs = 0;
        double* yh = yhat ? yhat + j * n_obs * 3 : nullptr;
        if (!with_grad) {
            RealOps ops;
            SupRhs<RealOps> f(ops, nd, I, p, theta[j], &nrhs);
            double v = solve_scaled_sse<RealOps, 3>(ops, f, I.u0, t0, tend, n_obs, obs_t, I.y, I.scale, opt, &st, yh, &nrhs);
            if (sse) sse[j] = v;
        } else {
            DualOps ops; ops.n = P + 1; ops.norm_partials = false;
            std::vector<Dual> dp(P);
            for (int a = 0; a < P; ++a) { dp[a] = dconst(p[a], P + 1); dp[a].d[a] = 1.0; }
            Dual dth = dconst(theta[j], P + 1); dth.d[P] = 1.0;
            SupRhs<DualOps> f(ops, nd, I, dp.data(), dth, &nrhs);
            Dual v = solve_scaled_sse<DualOps, 3>(ops, f, I.u0, t0, tend, n_obs, obs_t, I.y, I.scale, opt, &st, yh, &nrhs);
            if (sse) sse[j] = v.v;
            const bool ok = std::isfinite(v.v);
            if (g_neural_traj) for (int a = 0; a < P; ++a) g_neural_traj[j * P + a] = ok ? v.d[a] : 0.0;
            if (g_theta) g_theta[j] = ok ? v.d[P] : 0.0;
        }
        if (stats) { stats[j * 4 + 0] = st.nacc; stats[j * 4 + 1] = st.nrej; stats[j * 4 + 2] = st.nrhs; stats[j * 4 + 3] = st.retcode; }
    }
    return 0;
}

int cude_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
