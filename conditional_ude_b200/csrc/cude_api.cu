// =====================================================================================
// cude_api.cu — C ABI (include/cude_b200.h) over the sm_100a kernels in cude_kernels.cuh.
// No CPU fallback: every compute entry point requires a CUDA device.
// =====================================================================================
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <new>
#include <mutex>

#include "../../include/cude_b200.h"
#include <cub/device/device_radix_sort.cuh>
#include "cude_kernels.cuh"
#include "cude_split.cuh"
#include "cude_sup_kernel.cuh"
#include "cude_warp.cuh"
#include "cude_train.cuh"

using namespace cude;

// ---------------------------------------------------------------- handles
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

#ifndef CUDE_MAX_CHUNKS
#define CUDE_MAX_CHUNKS 6    // pipeline depth of a host-buffer call; measured at 64 M trajectories: 3 / 4 / 5 / 6 / 8 / 16 / 32 chunks
                             // -> e2e 2.112 / 2.118 / 2.122 / 2.115 / 2.108 / 2.10 / 2.055e8 evals/s (kernel tails vs exposed first/last copy)
#endif
#ifndef CUDE_WEIGHTS_IN_CONSTANT_MEMORY
#define CUDE_WEIGHTS_IN_CONSTANT_MEMORY 1   // weights of the FP64 adjoint kernel as uniform operands from constant memory when they fit (0: shared memory)
#endif
#ifndef CUDE_WC_LOSS
#define CUDE_WC_LOSS 1                        // constant-memory weights in the loss-only kernel too (4 blocks per SM there)
#endif
#ifndef CUDE_WC_BSENS
#define CUDE_WC_BSENS 1                       // ... and in the beta-sensitivity kernel (3.4e8 vs 3.1e8 evals/s)
#endif
#ifndef CUDE_SUP_PACK_DEFAULT
#define CUDE_SUP_PACK_DEFAULT 1            // small suppression populations run several whole starts per 128-thread block (0: one start per block)
#endif
#ifndef CUDE_FLAT_IMAJOR_MIN_STARTS
#define CUDE_FLAT_IMAJOR_MIN_STARTS 4096
#endif
#ifndef CUDE_SUP_TWO_KERNEL
#define CUDE_SUP_TWO_KERNEL 1               // suppression gradient as forward-with-records + adjoint kernels (opts.split = 1: the fused kernel)
#endif
#ifndef CUDE_WARP32_MAX_TRAJ
#define CUDE_WARP32_MAX_TRAJ 512            // largest call that gives a whole warp to each trajectory; above, 8 lanes (kernel ms, 32 / 8 lanes / fused:
                                            // 57 trajectories 0.066 / 0.074 / 0.196, 1425: 0.091 / 0.079 / 0.213, 3990: 0.211 / 0.100 / 0.234, 11 400: 0.500 / 0.216 / 0.238)
#endif
#ifndef CUDE_WARP_MAX_TRAJ
#define CUDE_WARP_MAX_TRAJ 8192             // opts.balance = 0 (automatic): loss + full-gradient calls of at most this many trajectories take the
#endif                                      // warp-per-trajectory latency kernel (cude_warp.cuh): 1776 warps are resident at a time (3 blocks of 4 per SM), a wave takes ~55 us against ~210 us for the fused kernel
#ifndef CUDE_EXACT_MIN_IND
#define CUDE_EXACT_MIN_IND 32768            // opts.balance = 0 (automatic): populations of at least this many individuals take the two-kernel
                                            // gradient with exact lane balance; smaller ones the fused single-kernel adjoint (lower latency)
#endif
#ifndef CUDE_SPLIT_BYTES
#define CUDE_SPLIT_BYTES (64ull << 30)      // device memory the step records may take per group of starts (bench: 20 / 48 / 66 GB = 8 / 4 / 3 groups: 2.461 / 2.479 / 2.486e8 evals/s)
#endif
#ifndef CUDE_BETA_FORWARD_SENSITIVITY
#define CUDE_BETA_FORWARD_SENSITIVITY 1   // 0: beta-only gradients through the adjoint kernel (comparison builds)
#endif
struct cude_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;       // stream in use
    cudaStream_t own_stream = nullptr;   // stream created by the context
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    cude_stats stats{};
    bool stats_pending = false;
    DevBuf neural, cond, sse, partials, sums, gcond, counters, scratch;
    double* h_sums = nullptr;   // pinned
    size_t h_sums_cap = 0;
    unsigned long long* h_counters = nullptr;  // pinned [3]
    // host-buffer calls on large batches run as a pipeline of chunks of starts: H2D on s_in, kernels on `stream`, D2H on s_out
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[CUDE_MAX_CHUNKS] = {}, ev_comp[CUDE_MAX_CHUNKS] = {};
    const unsigned int* bal_order = nullptr;   // lane-balancing pointers of the call in flight (set by balance_prepare)
    unsigned int* bal_keys_out = nullptr;
    int chunk_mode = 0;                 // 0: single call; 1: first chunk of a pipelined call; 2: later chunk (counters/timer accumulate)
    struct KernCfg { const void* kern; size_t smem; int block; };
    std::vector<KernCfg> kern_cfg;      // kernel configurations whose shared-memory attributes have been set
    int sm_count = 0;
    // split gradient pipeline: three sets of {step records, node weights, per-trajectory arrays, record map, per-record
    // d/d beta, partial rows} so that consecutive groups of starts are in different stages at the same time; the
    // memory-bound stages run on the high-priority side stream s_hi next to the compute-bound ones on `stream`
    struct SplitSet { DevBuf rec, w, misc, map, gc, part; cudaEvent_t ev_k1 = nullptr, ev_scan = nullptr, ev_k3 = nullptr, ev_free = nullptr; };
    SplitSet sp[3];
    cudaStream_t s_hi = nullptr;
    cudaEvent_t ev_fork = nullptr;
    size_t split_budget = 0;            // device memory the step records may take (decided once: cudaMemGetInfo synchronises)
    // NCCL communicator of a sharded population (cude_comm_init_rank, or the ranks of a cude_mctx): the per-start sums
    // are all-reduced in place on `stream` (cude_multi.inl)
    void* comm = nullptr;
    int comm_nranks = 1, comm_rank = 0;
};

struct cude_population {
    cude_ctx* ctx = nullptr;
    int n_ind = 0, max_knots = 0, max_obs = 0;
    int* n_knots = nullptr;
    int* n_obs = nullptr;
    double* block = nullptr;   // one allocation holding all double arrays
    PopDev dev{};
    // lane balancing state (opts.balance): keys[0] = sorted order in use, keys[1] = raw keys the kernel writes
    mutable unsigned int* bal_keys[2] = {nullptr, nullptr};
    mutable size_t bal_cap = 0;          // entries per buffer
    mutable int bal_starts = 0;          // n_starts the state belongs to
    mutable bool bal_valid = false;      // keys[0] holds a sorted order
    mutable long long bal_calls = 0;     // balanced calls since (re)allocation
    mutable void* bal_temp = nullptr;
    mutable size_t bal_temp_bytes = 0;
    bool ragged = false;                 // individuals differ in their knot count or time span (e.g. Ohashi + Fujita): see a.flat
};

static thread_local std::string g_err;

// last user of the per-device constant weight array (CW_CONST): uploads are ordered behind its kernel
struct WConstUse { cudaEvent_t ev = nullptr; cudaStream_t stream = nullptr; bool used = false; };
static std::mutex g_wconst_mutex;
static std::vector<WConstUse> g_wconst_use;   // one entry per device, sized on first use

static int fail(cude_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    g_err = msg;
    return code;
}

#define CU_TRY(ctx, call)                                                                          \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char b__[512];                                                                         \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return fail(ctx, (e__ == cudaErrorMemoryAllocation) ? CUDE_ENOMEM : CUDE_ECUDA, b__);  \
        }                                                                                          \
    } while (0)

static int ensure(cude_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return CUDE_OK;
    if (b.p) { CU_TRY(ctx, cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    const size_t head = bytes / 8 < ((size_t)64 << 20) ? bytes / 8 : ((size_t)64 << 20);   // headroom against regrowth, at most 64 MB
    size_t want = bytes + head + 256;
    CU_TRY(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return CUDE_OK;
}

// ---------------------------------------------------------------- small host-side entry points
extern "C" int cude_abi_version(void) { return CUDE_B200_ABI_VERSION; }

extern "C" void cude_default_opts(cude_opts* o) {
    if (!o) return;
    o->abstol = 1e-6;   // OrdinaryDiffEq defaults used by parameter-estimation.jl:59
    o->reltol = 1e-3;
    o->maxiters = 1000000;   // OrdinaryDiffEq's __init: maxiters = anyadaptive(alg) ? 1000000 : typemax(Int)
    o->precision = 0;
    o->block = 0;
    o->balance = 0;
    o->split = 0;
}

extern "C" int cude_net_nparams(const cude_net* net) {
    if (!net || net->n_in < 1 || net->depth < 1 || net->width < 1) return CUDE_EINVAL;
    int p = 0, in = net->n_in;
    for (int l = 0; l < net->depth; ++l) { p += net->width * (in + 1); in = net->width; }
    return p + in + 1;
}

extern "C" void cude_van_cauter_parameters(double age, int t2dm, double* k0, double* k1, double* k2) {
    // src/c-peptide-models.jl:30-42
    const double ln2 = std::log(2.0);
    const double short_half_life = t2dm ? 4.52 : 4.95;
    const double fraction = t2dm ? 0.78 : 0.76;
    const double long_half_life = 0.14 * age + 29.2;
    const double kk1 = fraction * (ln2 / long_half_life) + (1 - fraction) * (ln2 / short_half_life);
    const double kk0 = (ln2 / short_half_life) * (ln2 / long_half_life) / kk1;
    const double kk2 = (ln2 / short_half_life) + (ln2 / long_half_life) - kk0 - kk1;
    if (k0) *k0 = kk0;
    if (k1) *k1 = kk1;
    if (k2) *k2 = kk2;
}

extern "C" int cude_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* cude_last_error(const cude_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

// ---------------------------------------------------------------- context
static void comm_release(cude_ctx* ctx);   // cude_multi.inl
extern "C" int cude_ctx_create(int device, cude_ctx** out) {
    if (!out) return fail(nullptr, CUDE_EINVAL, "cude_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, CUDE_ENODEVICE, std::string("no CUDA device available (there is no CPU fallback): ") +
                                                 (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    if (device < 0 || device >= ndev) return fail(nullptr, CUDE_EINVAL, "cude_ctx_create: bad device index");
    cude_ctx* ctx = new (std::nothrow) cude_ctx();
    if (!ctx) return fail(nullptr, CUDE_ENOMEM, "out of host memory");
    ctx->device = device;
    cudaError_t ce;
    if ((ce = cudaSetDevice(device)) != cudaSuccess ||
        (ce = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (ce = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (ce = cudaMallocHost(&ctx->h_counters, 3 * sizeof(unsigned long long))) != cudaSuccess) {
        const std::string msg = std::string("cude_ctx_create: ") + cudaGetErrorString(ce);
        ctx->stream = ctx->own_stream;
        cude_ctx_destroy(ctx);   // releases whatever was created
        return fail(nullptr, ce == cudaErrorMemoryAllocation ? CUDE_ENOMEM : CUDE_ECUDA, msg);
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return CUDE_OK;
}

extern "C" int cude_ctx_destroy(cude_ctx* ctx) {
    if (!ctx) return CUDE_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    comm_release(ctx);
    DevBuf* bufs[] = {&ctx->neural, &ctx->cond, &ctx->sse, &ctx->partials, &ctx->sums, &ctx->gcond, &ctx->counters, &ctx->scratch};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    if (ctx->s_hi) cudaStreamSynchronize(ctx->s_hi);
    for (auto& st : ctx->sp) {
        DevBuf* sb[] = {&st.rec, &st.w, &st.misc, &st.map, &st.gc, &st.part};
        for (DevBuf* b : sb) if (b->p) cudaFree(b->p);
        cudaEvent_t* ev[] = {&st.ev_k1, &st.ev_scan, &st.ev_k3, &st.ev_free};
        for (cudaEvent_t* e : ev) if (*e) cudaEventDestroy(*e);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->s_hi) cudaStreamDestroy(ctx->s_hi);
    if (ctx->h_sums) cudaFreeHost(ctx->h_sums);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (int k = 0; k < CUDE_MAX_CHUNKS; ++k) {
        if (ctx->ev_in[k]) cudaEventDestroy(ctx->ev_in[k]);
        if (ctx->ev_comp[k]) cudaEventDestroy(ctx->ev_comp[k]);
    }
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return CUDE_OK;
}

extern "C" void* cude_ctx_stream(cude_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int cude_ctx_set_stream(cude_ctx* ctx, void* cuda_stream) {
    if (!ctx) return CUDE_EINVAL;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return CUDE_OK;
}

static void trace_dump();
static int collect_stats(cude_ctx* ctx) {
    if (!ctx->stats_pending) return CUDE_OK;
    CU_TRY(ctx, cudaMemcpyAsync(ctx->h_counters, ctx->counters.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.n_acc = ctx->h_counters[0];
    ctx->stats.n_rej = ctx->h_counters[1];
    ctx->stats.n_fail = ctx->h_counters[2];
    ctx->stats.n_rhs = 6ull * (ctx->stats.n_acc + ctx->stats.n_rej) + 2ull * ctx->stats.n_traj;
    ctx->stats.kernel_ms = ms;
    ctx->stats_pending = false;
    trace_dump();
    return CUDE_OK;
}

extern "C" int cude_sync(cude_ctx* ctx) {
    if (!ctx) return CUDE_EINVAL;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CUDE_OK;
}

extern "C" int cude_get_stats(cude_ctx* ctx, cude_stats* out) {
    if (!ctx || !out) return CUDE_EINVAL;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = collect_stats(ctx);
    if (rc) return rc;
    *out = ctx->stats;
    return CUDE_OK;
}

// ---------------------------------------------------------------- population
extern "C" int cude_population_create(cude_ctx* ctx, int n_ind,
                                      int max_knots, const int* n_knots, const double* knot_t, const double* knot_g,
                                      int max_obs, const int* n_obs, const double* obs_t, const double* obs_y,
                                      const double* kin, const double* covariate, cude_population** out) {
    if (!ctx || !out) return fail(ctx, CUDE_EINVAL, "cude_population_create: NULL ctx/out");
    *out = nullptr;
    if (n_ind < 1 || max_knots < 2 || max_obs < 1 || !n_knots || !knot_t || !knot_g || !n_obs || !obs_t || !obs_y || !kin)
        return fail(ctx, CUDE_EINVAL, "cude_population_create: bad argument");
    const size_t N = (size_t)n_ind, K = (size_t)max_knots, M = (size_t)max_obs;
    // validate + transpose to struct-of-arrays (individual index fastest)
    std::vector<double> h((3 * K + 2 * M + 5) * N, 0.0);
    double* hkt = h.data();
    double* hkg = hkt + K * N;
    double* hsl = hkg + K * N;
    double* hot = hsl + K * N;
    double* hoy = hot + M * N;
    double* hk0 = hoy + M * N;
    double* hk1 = hk0 + N;
    double* hk2 = hk1 + N;
    double* hc0 = hk2 + N;
    double* hcv = hc0 + N;
    bool ragged = false;
    for (size_t i = 0; i < N; ++i) {
        const int nk = n_knots[i], no = n_obs[i];
        if (nk < 2 || nk > max_knots || no < 0 || no > max_obs) return fail(ctx, CUDE_EINVAL, "cude_population_create: n_knots/n_obs out of range");
        for (int k = 0; k < max_knots; ++k) {
            // pad with the last knot so that stray reads stay finite
            const int kk = k < nk ? k : nk - 1;
            hkt[k * N + i] = knot_t[i * K + kk];
            hkg[k * N + i] = knot_g[i * K + kk];
            if (k + 1 < nk) {
                const double dtk = knot_t[i * K + k + 1] - knot_t[i * K + k];
                if (!(dtk > 0.0)) return fail(ctx, CUDE_EINVAL, "cude_population_create: knot times must be strictly increasing");
                hsl[k * N + i] = (knot_g[i * K + k + 1] - knot_g[i * K + k]) / dtk;   // DataInterpolations slope
            }
        }
        const double t0 = knot_t[i * K], t1 = knot_t[i * K + nk - 1];
        if (nk != n_knots[0] || t0 != knot_t[0] || t1 != knot_t[n_knots[0] - 1]) ragged = true;
        for (int k = 0; k < max_obs; ++k) {
            const int kk = k < no ? k : (no > 0 ? no - 1 : 0);
            hot[k * N + i] = no > 0 ? obs_t[i * M + kk] : 0.0;
            hoy[k * N + i] = no > 0 ? obs_y[i * M + kk] : 0.0;
            if (k < no) {
                const double ts = obs_t[i * M + k];
                if (!(ts >= t0 && ts <= t1)) return fail(ctx, CUDE_EINVAL, "cude_population_create: observation time outside the glucose time span");
                if (k > 0 && !(ts > obs_t[i * M + k - 1])) return fail(ctx, CUDE_EINVAL, "cude_population_create: observation times must be strictly increasing");
            }
        }
        hk0[i] = kin[4 * i + 0]; hk1[i] = kin[4 * i + 1]; hk2[i] = kin[4 * i + 2]; hc0[i] = kin[4 * i + 3];
        hcv[i] = covariate ? covariate[i] : 0.0;
    }
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cude_population* pop = new (std::nothrow) cude_population();
    if (pop) pop->ragged = ragged;
    if (!pop) return fail(ctx, CUDE_ENOMEM, "out of host memory");
    pop->ctx = ctx; pop->n_ind = n_ind; pop->max_knots = max_knots; pop->max_obs = max_obs;
    cudaError_t e;
    if ((e = cudaMalloc(&pop->block, h.size() * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&pop->n_knots, N * sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&pop->n_obs, N * sizeof(int))) != cudaSuccess) {
        cude_population_destroy(pop);
        return fail(ctx, CUDE_ENOMEM, std::string("cude_population_create: cudaMalloc failed: ") + cudaGetErrorString(e));
    }
    if ((e = cudaMemcpyAsync(pop->block, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(pop->n_knots, n_knots, N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(pop->n_obs, n_obs, N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
        cude_population_destroy(pop);
        return fail(ctx, CUDE_ECUDA, std::string("cude_population_create: upload failed: ") + cudaGetErrorString(e));
    }
    PopDev& d = pop->dev;
    d.n_ind = n_ind; d.max_knots = max_knots; d.max_obs = max_obs;
    d.n_knots = pop->n_knots; d.n_obs = pop->n_obs;
    d.knot_t = pop->block; d.knot_g = d.knot_t + K * N; d.slope = d.knot_g + K * N;
    d.obs_t = d.slope + K * N; d.obs_y = d.obs_t + M * N;
    d.k0 = d.obs_y + M * N; d.k1 = d.k0 + N; d.k2 = d.k1 + N; d.c0 = d.k2 + N;
    d.cov = covariate ? d.c0 + N : nullptr;
    *out = pop;
    return CUDE_OK;
}

extern "C" int cude_population_destroy(cude_population* pop) {
    if (!pop) return CUDE_OK;
    if (pop->ctx) cudaSetDevice(pop->ctx->device);
    if (pop->block) cudaFree(pop->block);
    if (pop->n_knots) cudaFree(pop->n_knots);
    if (pop->n_obs) cudaFree(pop->n_obs);
    for (int k = 0; k < 2; ++k) if (pop->bal_keys[k]) cudaFree(pop->bal_keys[k]);
    if (pop->bal_temp) cudaFree(pop->bal_temp);
    delete pop;
    return CUDE_OK;
}

extern "C" int cude_population_size(const cude_population* pop) { return pop ? pop->n_ind : CUDE_EINVAL; }

// ---------------------------------------------------------------- launch
typedef void (*eval_kernel_t)(const EvalArgs);

// grad: discrete adjoint (all gradients); bsens: d/d cond only by forward sensitivity (FP64 only)
template <class NS>
static eval_kernel_t pick(bool grad, bool mixed, bool bsens, bool fbwd, bool wc) {
    if (wc && grad && fbwd) return cude_eval_kernel<NS, true, false, false, true, true>;   // FP32 adjoint network, FP64 forward weights in constant memory
    if (wc && grad) return cude_eval_kernel<NS, true, false, false, false, true>;   // FP64 adjoint kernel, weights in constant memory
#if CUDE_WC_LOSS
    if (wc && !bsens) return cude_eval_kernel<NS, false, false, false, false, true>;
#endif
#if CUDE_WC_BSENS
    if (wc && bsens) return cude_eval_kernel<NS, false, false, true, false, true>;
#endif
    if (bsens) return cude_eval_kernel<NS, false, false, true>;
    if (fbwd && grad) return cude_eval_kernel<NS, true, false, false, true>;
    if (mixed) return grad ? cude_eval_kernel<NS, true, true> : cude_eval_kernel<NS, false, true>;
    return grad ? cude_eval_kernel<NS, true, false> : cude_eval_kernel<NS, false, false>;
}

static eval_kernel_t select_kernel(const cude_net* net, bool grad, bool mixed, bool bsens = false, bool fbwd = false, bool wc = false) {
    if (net->depth == 2 && net->width == 4) {
        if (net->n_in == 2) return pick<NetShape<2, 2, 4>>(grad, mixed, bsens, fbwd, wc);   // chain(4, 2, tanh), 02-conditional.jl:22
        if (net->n_in == 3) return pick<NetShape<3, 2, 4>>(grad, mixed, bsens, fbwd, wc);   // covariate net, 07-covariate-inclusion.jl:32
    }
    return nullptr;
}

static int choose_block(const cude_opts* o, int n_ind, bool flat) {
    if (o->block > 0) return o->block;
    if (flat) return 128;
    if (n_ind > 64) return 128;
    if (n_ind > 32) return 64;
    return 32;
}

// ---- lane balancing (cude_opts.balance): per start, the individuals sorted by the step counts of an earlier call.
// The kernel writes keys (steps << 24 | individual) in natural order on refresh calls; a stable radix sort over the top
// 8 bits of each start's segment yields the order the following calls run in.
#ifndef CUDE_BAL_REFRESH
#define CUDE_BAL_REFRESH 8
#endif
static const int BAL_REFRESH = CUDE_BAL_REFRESH;            // refresh the grouping every this many balanced calls
static bool balance_applies(const cude_opts& o, const cude_population* pop, bool grad, bool flat, int n_starts) {
    return o.balance == 1 && grad && !flat && pop->n_ind >= 4096 && pop->n_ind < (1 << 24) && n_starts <= 65536;
}
// sets ctx->bal_order / ctx->bal_keys_out for a call over n_starts starts (pointers to start 0)
static int balance_prepare(cude_ctx* ctx, const cude_population* pop, int n_starts) {
    const size_t need = (size_t)pop->n_ind * n_starts;
    if (pop->bal_starts != n_starts || pop->bal_cap < need) {
        for (int k = 0; k < 2; ++k) { if (pop->bal_keys[k]) cudaFree(pop->bal_keys[k]); pop->bal_keys[k] = nullptr; }
        pop->bal_cap = 0; pop->bal_valid = false; pop->bal_calls = 0; pop->bal_starts = n_starts;
        for (int k = 0; k < 2; ++k) CU_TRY(ctx, cudaMalloc(&pop->bal_keys[k], need * sizeof(unsigned int)));
        pop->bal_cap = need;
    }
    ctx->bal_order = pop->bal_valid ? pop->bal_keys[0] : nullptr;
    ctx->bal_keys_out = (pop->bal_calls % BAL_REFRESH == 0) ? pop->bal_keys[1] : nullptr;
    return CUDE_OK;
}
// after the kernels of the call: on refresh calls sort every start's keys into the order buffer (same stream)
static int balance_finish(cude_ctx* ctx, const cude_population* pop, int n_starts, int* launches) {
    const bool refresh = ctx->bal_keys_out != nullptr;
    ctx->bal_order = nullptr; ctx->bal_keys_out = nullptr;
    ++pop->bal_calls;
    if (!refresh) return CUDE_OK;
    const int N = pop->n_ind;
    size_t bytes = 0;
    CU_TRY(ctx, cub::DeviceRadixSort::SortKeys(nullptr, bytes, pop->bal_keys[1], pop->bal_keys[0], N, 24, 32, ctx->stream));
    if (bytes > pop->bal_temp_bytes) {
        if (pop->bal_temp) cudaFree(pop->bal_temp);
        pop->bal_temp = nullptr; pop->bal_temp_bytes = 0;
        CU_TRY(ctx, cudaMalloc(&pop->bal_temp, bytes));
        pop->bal_temp_bytes = bytes;
    }
    for (int s = 0; s < n_starts; ++s) {
        size_t b = pop->bal_temp_bytes;
        CU_TRY(ctx, cub::DeviceRadixSort::SortKeys(pop->bal_temp, b, pop->bal_keys[1] + (size_t)s * N, pop->bal_keys[0] + (size_t)s * N,
                                                   N, 24, 32, ctx->stream));
    }
    if (launches) *launches += n_starts;
    pop->bal_valid = true;
    return CUDE_OK;
}

static int eval_dev_impl(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts_in,
                         int n_starts, const double* d_neural, long long neural_stride, const double* d_cond,
                         int want_grad, double cond_scale,
                         double* d_sse_out, double* d_sums_out, double* d_g_cond, double* d_yhat);

// Shared-memory attributes of a kernel configuration, set once: dynamic size above 48 KB, and a carve-out of what the
// resident blocks need and no more — the rest of the 256 KB stays L1, which serves the per-thread step ring and the
// register spills of the gradient kernels.
static int prep_kernel(cude_ctx* ctx, const void* kern, int B, size_t smem) {
    for (const auto& k : ctx->kern_cfg) if (k.kern == kern && k.smem == smem && k.block == B) return CUDE_OK;
    if (smem > 227 * 1024) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: too many knots/observations for shared memory; lower opts.block");
    if (smem > 48 * 1024) CU_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, B, smem));
    if (nb < 1) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: kernel does not fit on an SM with this block size");
    const size_t need = (size_t)nb * (smem + 1024);
    int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    CU_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    // entries of a kernel with another configuration are replaced (the attribute is per kernel)
    for (auto& k : ctx->kern_cfg) if (k.kern == kern) { k.smem = smem; k.block = B; return CUDE_OK; }
    ctx->kern_cfg.push_back({kern, smem, B});
    return CUDE_OK;
}

// the per-device constant weight array (CW_CONST): upload ordered behind the last launch that read it, also across contexts
static int wconst_upload(cude_ctx* ctx, const double* d_w, size_t n) {
    std::lock_guard<std::mutex> lk(g_wconst_mutex);
    if ((int)g_wconst_use.size() <= ctx->device) g_wconst_use.resize((size_t)ctx->device + 1);
    WConstUse& u = g_wconst_use[ctx->device];
    if (!u.ev) CU_TRY(ctx, cudaEventCreateWithFlags(&u.ev, cudaEventDisableTiming));
    if (u.used && u.stream != ctx->stream) CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, u.ev, 0));
    CU_TRY(ctx, cudaMemcpyToSymbolAsync(CW_CONST, d_w, n * sizeof(double), 0, cudaMemcpyDeviceToDevice, ctx->stream));
    return CUDE_OK;
}
static int wconst_used(cude_ctx* ctx) {      // after the launches that read the array
    std::lock_guard<std::mutex> lk(g_wconst_mutex);
    WConstUse& u = g_wconst_use[ctx->device];
    CU_TRY(ctx, cudaEventRecord(u.ev, ctx->stream));
    u.used = true; u.stream = ctx->stream;
    return CUDE_OK;
}

// Record-memory budget of the split / two-kernel gradient paths: CUDE_SPLIT_BYTES, at most 40 % of what was free when the
// context first asked (asked once: cudaMemGetInfo synchronises the device, which would serialise a pipelined host call).
static int split_budget(cude_ctx* ctx, size_t* out) {
    if (!ctx->split_budget) {
        size_t free_b = 0, total_b = 0;
        CU_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
        size_t b = CUDE_SPLIT_BYTES;
        if (const char* e = getenv("CUDE_SCRATCH_BYTES")) {      // tests: force the grouped / multi-launch code paths on small batches
            const unsigned long long v = strtoull(e, nullptr, 10);
            if (v) b = (size_t)v;
        }
        if (b > free_b * 2 / 5) b = free_b * 2 / 5;
        ctx->split_budget = b > (1u << 20) ? b : (1u << 20);
    }
    *out = ctx->split_budget;
    return CUDE_OK;
}

// ---------------------------------------------------------------- split gradient pipeline (cude_split.cuh)
typedef void (*node_kernel_t)(const NodeArgs);
typedef void (*final_kernel_t)(const FinalArgs);
typedef void (*adj_kernel_t)(const AdjArgs);
struct SplitKernels { eval_kernel_t k1 = nullptr; node_kernel_t k3 = nullptr; final_kernel_t k4 = nullptr; adj_kernel_t ka = nullptr; };

template <class NS>
static SplitKernels pick_split(bool fbwd, bool wc) {
    SplitKernels k;
    k.k1 = wc ? cude_eval_kernel<NS, false, false, false, false, true, true> : cude_eval_kernel<NS, false, false, false, false, false, true>;
    if (fbwd) { k.k3 = wc ? cude_node_kernel<NS, float, true> : cude_node_kernel<NS, float, false>; k.k4 = cude_final_kernel<NS, float>;
                k.ka = wc ? cude_adjoint_kernel<NS, float, true> : cude_adjoint_kernel<NS, float, false>; }
    else { k.k3 = wc ? cude_node_kernel<NS, double, true> : cude_node_kernel<NS, double, false>; k.k4 = cude_final_kernel<NS, double>;
           k.ka = wc ? cude_adjoint_kernel<NS, double, true> : cude_adjoint_kernel<NS, double, false>; }
    return k;
}
static SplitKernels select_split(const cude_net* net, bool fbwd, bool wc) {
    if (net->depth == 2 && net->width == 4) {
        if (net->n_in == 2) return pick_split<NetShape<2, 2, 4>>(fbwd, wc);
        if (net->n_in == 3) return pick_split<NetShape<3, 2, 4>>(fbwd, wc);
    }
    return SplitKernels();
}

// Loss + full gradient of n_starts starts through the five stages, in groups of starts whose step records fit the
// memory budget.  `a` carries the call's arguments (pointers to start 0); `fused` is the fused gradient kernel, launched
// once per group as the fallback for trajectories with more than SPLIT_CAP accepted steps (its blocks return at once
// when the block has none).  Nothing synchronises with the host.
//
// Two streams: the compute-bound stages 1 and 4 (and the fallback) run back to back on ctx->stream; the memory-bound
// stages 2, 3 and 5 and the second-stage reduction run on a high-priority side stream with small grids, so that they
// share the SMs with the compute stages of the neighbouring groups (three buffer sets keep three groups in flight):
//   ctx->stream : K1(0) K1(1) K4(0) K1(2) K4(1) K1(3) ...
//   side stream :       K2(0) K2(1) K5(0) K2(2) K5(1) ...
#ifndef CUDE_RECUR_BLOCKS
#define CUDE_RECUR_BLOCKS 2      // blocks of 128 threads per SM for the adjoint recursion (stage 2)
#endif
// Diagnostic timeline (environment variable CUDE_SPLIT_TRACE=1): timing events around every stage on both streams, printed
// to stderr by the next cude_get_stats.  Off in normal operation.
struct TraceEv { const char* name; int group; cudaEvent_t a, b; };
static std::vector<TraceEv> g_trace;
static cudaEvent_t g_trace_base = nullptr;
static bool trace_on() { static int on = -1; if (on < 0) { const char* e = getenv("CUDE_SPLIT_TRACE"); on = (e && *e == '1') ? 1 : 0; } return on == 1; }
static void trace_begin(cudaStream_t s, const char* name, int g) {
    if (!trace_on()) return;
    TraceEv t{name, g, nullptr, nullptr};
    cudaEventCreate(&t.a); cudaEventCreate(&t.b);
    cudaEventRecord(t.a, s);
    g_trace.push_back(t);
}
static void trace_end(cudaStream_t s) { if (trace_on() && !g_trace.empty()) cudaEventRecord(g_trace.back().b, s); }
static void trace_dump() {
    if (!trace_on() || g_trace.empty()) return;
    for (auto& t : g_trace) {
        float t0 = 0, t1 = 0;
        cudaEventSynchronize(t.b);
        cudaEventElapsedTime(&t0, g_trace_base, t.a); cudaEventElapsedTime(&t1, g_trace_base, t.b);
        fprintf(stderr, "trace %-8s g%-3d %9.3f -> %9.3f  (%7.3f ms)\n", t.name, t.group, t0, t1, t1 - t0);
        cudaEventDestroy(t.a); cudaEventDestroy(t.b);
    }
    g_trace.clear();
}
// second-stage reduction of the partial rows: a warp per (start, q) striding over many rows, or a thread per (start, q) for few
static void reduce_partials(cudaStream_t st, const double* partials, int nrows, int n_starts, int np1, int nred, double* sums) {
    const long long n = (long long)n_starts * np1;
    if (nrows <= 16) cude_reduce_partials_few<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partials, nrows, n_starts, np1, nred, sums);
    else cude_reduce_partials<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(partials, nrows, n_starts, np1, nred, sums);
}

static int sm_count(cude_ctx* ctx) {
    if (!ctx->sm_count) CU_TRY(ctx, cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
    return CUDE_OK;
}

static int run_split(cude_ctx* ctx, const cude_net* net, EvalArgs a, int B, int nchunks, bool fbwd, bool wc, size_t n_w,
                     eval_kernel_t fused, size_t smem_fused, double* d_sums_out, int* launches) {
    const int P = cude_net_nparams(net), np1 = P + 1, N = a.pop.n_ind, S = a.n_starts, nw = B / 32;
    const int M = a.pop.max_obs, K = a.pop.max_knots;
    const SplitKernels sk = select_split(net, fbwd, wc);
    if (!sk.k1) return fail(ctx, CUDE_EUNSUPPORTED, "cude_eval_dev: network shape not compiled in");
    int rc_sm = sm_count(ctx);
    if (rc_sm) return rc_sm;
    if (!ctx->s_hi) {
        int least = 0, greatest = 0;
        CU_TRY(ctx, cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CU_TRY(ctx, cudaStreamCreateWithPriority(&ctx->s_hi, cudaStreamNonBlocking, greatest));
        CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        for (auto& st : ctx->sp) {
            CU_TRY(ctx, cudaEventCreateWithFlags(&st.ev_k1, cudaEventDisableTiming));
            CU_TRY(ctx, cudaEventCreateWithFlags(&st.ev_scan, cudaEventDisableTiming));
            CU_TRY(ctx, cudaEventCreateWithFlags(&st.ev_k3, cudaEventDisableTiming));
            CU_TRY(ctx, cudaEventCreateWithFlags(&st.ev_free, cudaEventDisableTiming));
        }
    }
    cudaStream_t sM = ctx->stream, sH = ctx->s_hi;
    // ---- groups of starts: three sets of buffers share the budget ----
    const size_t per_traj = (size_t)(SPLIT_W + SPLIT_WW) * SPLIT_CAP * 8 + (size_t)SPLIT_CAP * (4 + 8) + (size_t)M * 8 + 4 * 8 + 8 + 3 * (size_t)np1 * 8 / 32 + 64;
    size_t budget = 0;
    int rc0 = split_budget(ctx, &budget);
    if (rc0) return rc0;
    long long sg = (long long)(budget / 3 / (per_traj * (size_t)N));
    if (sg < 1) sg = 1;
    if (sg > (S + 2) / 3) sg = (S + 2) / 3;                                   // at least three groups when there are three starts
    while ((long long)N * sg * SPLIT_CAP > 0x7fffffffLL && sg > 1) --sg;      // 32-bit record offsets
    if ((long long)N * SPLIT_CAP > 0x7fffffffLL) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: population too large for one call; shard the individuals");
    const int ngroups = (int)((S + sg - 1) / sg);
    const int Sg = (S + ngroups - 1) / ngroups;
    const size_t ntg = (size_t)N * Sg;
    const int nscan = (int)((ntg + SCAN_TILE - 1) / SCAN_TILE);
    int G = (4 * ctx->sm_count * CUDE_NODE_MIN_BLOCKS) / Sg;               // stage-4 blocks per start: at most 4 full waves of resident blocks per group
    const long long max_tiles = ((long long)N * SPLIT_CAP + CUDE_NODE_THREADS - 1) / CUDE_NODE_THREADS;
    if (G > max_tiles) G = (int)max_tiles;
    if (G < 1) G = 1;
    const int nwn = CUDE_NODE_THREADS / 32;
    const int nA = G * nwn, nB = nchunks * nw;
    const size_t rowsA = (size_t)Sg * nA, rowsB = (size_t)Sg * nB;
    int nseg = (nA + 2 * nB + 2047) / 2048;                                 // second-stage reduction: ~2048 rows per block
    if (nseg > 64) nseg = 64;
    // per-trajectory arrays of a set: res [M] , beta, wsum, sse (double), off (uint, +1), nrec (int); then bsum (uint, nscan + 1),
    // blkflag (int), segment sums of the reduction
    const size_t o_res = 0, o_beta = o_res + ntg * M * 8, o_wsum = o_beta + ntg * 8, o_sse = o_wsum + ntg * 8, o_off = o_sse + ntg * 8,
                 o_nrec = o_off + ((ntg + 2) * 4 + 7) / 8 * 8, o_bsum = o_nrec + (ntg * 4 + 7) / 8 * 8,
                 o_flag = o_bsum + (((size_t)nscan + 2) * 4 + 7) / 8 * 8, o_list = o_flag + ((size_t)Sg * nchunks * 4 + 4 + 7) / 8 * 8,
                 o_seg = o_list + ((size_t)Sg * nchunks * 4 + 7) / 8 * 8, misc_bytes = o_seg + (size_t)Sg * nseg * np1 * 8;
    const unsigned fb_grid = (unsigned)(ctx->sm_count * CUDE_MIN_BLOCKS);          // resident blocks walking the fallback list
    int rc;
    const int nsets = ngroups < 3 ? ngroups : 3;
    for (int k = 0; k < nsets; ++k) {
        cude_ctx::SplitSet& st = ctx->sp[k];
        if ((rc = ensure(ctx, st.rec, ntg * SPLIT_W * SPLIT_CAP * sizeof(double)))) return rc;
        if ((rc = ensure(ctx, st.w, ntg * SPLIT_WW * SPLIT_CAP * sizeof(double)))) return rc;
        if ((rc = ensure(ctx, st.misc, misc_bytes))) return rc;
        if ((rc = ensure(ctx, st.map, ntg * SPLIT_CAP * sizeof(unsigned int)))) return rc;
        if ((rc = ensure(ctx, st.gc, ntg * SPLIT_CAP * sizeof(double)))) return rc;
        if ((rc = ensure(ctx, st.part, (rowsA + 2 * rowsB) * np1 * sizeof(double)))) return rc;
    }
    // ---- kernel attributes ----
    const int nacc = 2 * net->width + (net->depth - 1) * net->width * (net->width + 1) + net->width + 1;
    const size_t smem1 = sizeof(double) * eval_smem_doubles(P, nacc, K, M, B, false, false, true);    // loss-only rows + the 5 dG rows
    const size_t smem3 = sizeof(double) * node_smem_doubles(P, CUDE_NODE_THREADS, fbwd, wc);
    if ((rc = prep_kernel(ctx, (const void*)sk.k1, B, smem1))) return rc;
    if ((rc = prep_kernel(ctx, (const void*)sk.k3, CUDE_NODE_THREADS, smem3))) return rc;
    if ((rc = prep_kernel(ctx, (const void*)fused, B, smem_fused))) return rc;
    if (wc && (rc = wconst_upload(ctx, a.neural, n_w))) return rc;          // the whole call's weights; groups index it by wc_base
    CU_TRY(ctx, cudaEventRecord(ctx->ev_fork, sM));
    CU_TRY(ctx, cudaStreamWaitEvent(sH, ctx->ev_fork, 0));
    if (trace_on()) { if (!g_trace_base) cudaEventCreate(&g_trace_base); cudaEventRecord(g_trace_base, sM); }

    struct GroupView { EvalArgs e; double *wsum, *seg, *pA, *pB, *pC; unsigned int *off, *bsum; int ns, g0; size_t nt; unsigned nblk; };
    auto view = [&](int g) {
        cude_ctx::SplitSet& st = ctx->sp[g % 3];
        GroupView v;
        v.g0 = g * Sg;
        v.ns = (S - v.g0 < Sg) ? S - v.g0 : Sg;
        v.nt = (size_t)N * v.ns;
        v.nblk = (unsigned)((size_t)v.ns * nchunks);
        char* const misc = (char*)st.misc.p;
        EvalArgs e = a;
        e.n_starts = v.ns;
        e.neural = a.neural + (size_t)v.g0 * a.neural_stride;
        e.wc_base = a.wc_base + (long long)v.g0 * a.neural_stride;
        e.cond = a.cond + (size_t)v.g0 * N;
        e.sse_out = a.sse_out ? a.sse_out + (size_t)v.g0 * N : nullptr;
        e.g_cond = a.g_cond + (size_t)v.g0 * N;
        e.order = a.order ? a.order + (size_t)v.g0 * N : nullptr;
        e.keys_out = a.keys_out ? a.keys_out + (size_t)v.g0 * N : nullptr;
        e.partials = nullptr;
        e.sp_rec = (double*)st.rec.p;
        e.sp_res = (double*)(misc + o_res); e.sp_beta = (double*)(misc + o_beta); e.sp_sse = (double*)(misc + o_sse);
        e.sp_nrec = (int*)(misc + o_nrec); e.sp_blkflag = (int*)(misc + o_flag);
        e.sp_blklist = (int*)(misc + o_list); e.sp_blkcount = e.sp_blkflag + (size_t)Sg * nchunks;
        v.e = e;
        v.wsum = (double*)(misc + o_wsum); v.off = (unsigned int*)(misc + o_off); v.bsum = (unsigned int*)(misc + o_bsum);
        v.seg = (double*)(misc + o_seg);
        v.pA = (double*)st.part.p; v.pB = v.pA + rowsA * np1; v.pC = v.pB + rowsB * np1;
        return v;
    };
    for (int g = 0; g <= ngroups; ++g) {
        if (g < ngroups) {
            cude_ctx::SplitSet& st = ctx->sp[g % 3];
            const GroupView v = view(g);
            if (g >= 3) CU_TRY(ctx, cudaStreamWaitEvent(sM, st.ev_free, 0));          // the set's previous group has been finished
            CU_TRY(ctx, cudaMemsetAsync(v.e.sp_blkflag, 0, ((size_t)Sg * nchunks + 1) * sizeof(int), sM));
            CU_TRY(ctx, cudaMemsetAsync(v.pC, 0, (size_t)v.ns * nB * np1 * sizeof(double), sM));
            // stage 1: forward solve -> step records {t, h, dG[5]}, residuals
            trace_begin(sM, "fwd", g);
            sk.k1<<<v.nblk, B, smem1, sM>>>(v.e);
            trace_end(sM);
            CU_TRY(ctx, cudaGetLastError());
            CU_TRY(ctx, cudaEventRecord(st.ev_k1, sM));
            CU_TRY(ctx, cudaStreamWaitEvent(sH, st.ev_k1, 0));
            // stage 2: adjoint recursion -> node weights
            RecurArgs ra{};
            ra.pop = a.pop; ra.ntraj = (long long)v.nt; ra.sp_rec = v.e.sp_rec; ra.sp_w = (double*)st.w.p; ra.sp_res = v.e.sp_res;
            ra.sp_nrec = v.e.sp_nrec; ra.sp_wsum = v.wsum;
            long long rblocks = (long long)(v.nt + 127) / 128;
            if (rblocks > (long long)ctx->sm_count * CUDE_RECUR_BLOCKS) rblocks = (long long)ctx->sm_count * CUDE_RECUR_BLOCKS;
            trace_begin(sH, "recur", g);
            cude_recur_kernel<<<(unsigned)rblocks, 128, 0, sH>>>(ra);
            trace_end(sH);
            CU_TRY(ctx, cudaGetLastError());
            trace_begin(sH, "scan", g);
            // stage 3: flat record list
            const int nsc = (int)((v.nt + SCAN_TILE - 1) / SCAN_TILE);
            cude_scan_sums<<<nsc, SCAN_T, 0, sH>>>(v.e.sp_nrec, (long long)v.nt, v.bsum);
            cude_scan_bsums<<<1, 1024, 0, sH>>>(v.bsum, nsc);
            cude_scan_final<<<nsc, SCAN_T, 0, sH>>>(v.e.sp_nrec, (long long)v.nt, v.bsum, v.off, (unsigned int*)st.map.p);
            trace_end(sH);
            CU_TRY(ctx, cudaGetLastError());
            CU_TRY(ctx, cudaEventRecord(st.ev_scan, sH));
            *launches += 5;
        }
        if (g >= 1) {
            cude_ctx::SplitSet& st = ctx->sp[(g - 1) % 3];
            const GroupView v = view(g - 1);
            CU_TRY(ctx, cudaStreamWaitEvent(sM, st.ev_scan, 0));
            // stage 4: network forward + backward, one thread per record
            NodeArgs na{};
            na.pop = a.pop; na.neural = v.e.neural; na.neural_stride = a.neural_stride; na.wc_base = v.e.wc_base;
            na.sp_rec = v.e.sp_rec; na.sp_w = (const double*)st.w.p; na.off = v.off; na.map = (const unsigned int*)st.map.p;
            na.sp_beta = v.e.sp_beta; na.gc_rec = (double*)st.gc.p; na.partials = v.pA;
            trace_begin(sM, "node", g - 1);
            sk.k3<<<dim3((unsigned)G, (unsigned)v.ns), CUDE_NODE_THREADS, smem3, sM>>>(na);
            trace_end(sM);
            CU_TRY(ctx, cudaGetLastError());
            // fallback: trajectories with more than SPLIT_CAP accepted steps through the fused kernel (flagged blocks only)
            EvalArgs f = v.e;
            f.partials = v.pC; f.only_flag = v.e.sp_nrec; f.counters = nullptr; f.keys_out = nullptr; f.sse_out = nullptr;
            fused<<<(v.nblk < fb_grid ? v.nblk : fb_grid), B, smem_fused, sM>>>(f);
            CU_TRY(ctx, cudaGetLastError());
            CU_TRY(ctx, cudaEventRecord(st.ev_k3, sM));
            CU_TRY(ctx, cudaStreamWaitEvent(sH, st.ev_k3, 0));
            // stage 5: NN([0;beta]) node, d/d cond, sse rows; then the second-stage reduction
            FinalArgs fa{};
            fa.pop = a.pop; fa.n_starts = v.ns; fa.nchunks = nchunks; fa.neural = v.e.neural; fa.neural_stride = a.neural_stride;
            fa.sp_nrec = v.e.sp_nrec; fa.sp_beta = v.e.sp_beta; fa.sp_wsum = v.wsum; fa.sp_sse = v.e.sp_sse; fa.off = v.off;
            fa.gc_rec = (const double*)st.gc.p; fa.cond_scale = a.cond_scale; fa.g_cond = v.e.g_cond; fa.partials = v.pB;
            trace_begin(sH, "final", g - 1);
            sk.k4<<<v.nblk, B, 0, sH>>>(fa);
            CU_TRY(ctx, cudaGetLastError());
            *launches += 3;
            if (d_sums_out) {
                cude_reduce_rows<<<dim3((unsigned)nseg, (unsigned)v.ns), RED_T, 0, sH>>>(v.pA, nA, v.pB, v.pC, nB, np1, nseg * np1, v.seg);
                cude_reduce_rows<<<dim3(1, (unsigned)v.ns), RED_T, 0, sH>>>(v.seg, nseg, nullptr, nullptr, 0, np1, np1, d_sums_out + (size_t)v.g0 * np1);
                CU_TRY(ctx, cudaGetLastError());
                *launches += 2;
            }
            trace_end(sH);
            CU_TRY(ctx, cudaEventRecord(st.ev_free, sH));
        }
    }
    CU_TRY(ctx, cudaStreamWaitEvent(sM, ctx->sp[(ngroups - 1) % 3].ev_free, 0));    // join: the side stream is in order
    if (wc && (rc = wconst_used(ctx))) return rc;
    return CUDE_OK;
}

// Two-kernel gradient with exact lane balance (opts.balance = 2): stage 1 (forward solve + step records + keys) -> every
// start's trajectories sorted by their accepted-step count (stable radix sort over the count byte: deterministic) -> the
// adjoint kernel in sorted order -> fused-kernel fallback for trajectories beyond SPLIT_CAP steps -> row reduction.  One
// stream, groups of starts sized by the record memory (set 0 of the split pipeline's buffers).
static int run_exact(cude_ctx* ctx, const cude_net* net, EvalArgs a, int B, int nchunks, bool fbwd, bool wc, size_t n_w,
                     eval_kernel_t fused, size_t smem_fused, double* d_sums_out, int* launches) {
    const int P = cude_net_nparams(net), np1 = P + 1, N = a.pop.n_ind, S = a.n_starts, nw = B / 32;
    const int M = a.pop.max_obs, K = a.pop.max_knots;
    if (N >= (1 << 24)) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: balance = 2 supports up to 2^24 individuals per population");
    const SplitKernels sk = select_split(net, fbwd, wc);
    if (!sk.k1) return fail(ctx, CUDE_EUNSUPPORTED, "cude_eval_dev: network shape not compiled in");
    cudaStream_t st = ctx->stream;
    const size_t per_traj = (size_t)SPLIT_W * SPLIT_CAP * 8 + (size_t)M * 8 + 2 * 8 + 4 + 2 * 4 + 2 * 2 + 2 * (size_t)np1 * 8 / 32 + 64;
    size_t budget = 0;
    int rc0 = split_budget(ctx, &budget);
    if (rc0) return rc0;
    long long sg = (long long)(budget / (per_traj * (size_t)N));
    if (sg < 1) sg = 1;
    if (sg > S) sg = S;
    if (sg > 256) sg = 256;                 // the sort key holds the start in 8 bits
    int ngroups, Sg;
    size_t ntg;
    cude_ctx::SplitSet& set = ctx->sp[0];
    for (;;) {                              // the step records of a group: if the device cannot give that much, halve the groups
        ngroups = (int)((S + sg - 1) / sg);
        Sg = (S + ngroups - 1) / ngroups;
        ntg = (size_t)N * Sg;
        if (ensure(ctx, set.rec, ntg * SPLIT_W * SPLIT_CAP * sizeof(double)) == CUDE_OK) break;
        (void)cudaGetLastError();
        if (sg <= 1) return fail(ctx, CUDE_ECUDA, "cude_eval_dev: no device memory for the step records of one start; use opts.balance = 3");
        sg = (sg + 1) / 2;
        ctx->split_budget = ctx->split_budget / 2 > ((size_t)1 << 20) ? ctx->split_budget / 2 : ((size_t)1 << 20);
    }
    const int nB = nchunks * nw;
    const size_t rowsB = (size_t)Sg * nB;
    int nseg = (2 * nB + 2047) / 2048;
    if (nseg > 64) nseg = 64;
    const size_t o_res = 0, o_beta = o_res + ntg * M * 8, o_sse = o_beta + ntg * 8, o_keys = o_sse + ntg * 8,
                 o_k16 = o_keys + 2 * ((ntg * 4 + 7) / 8 * 8), o_nrec = o_k16 + 2 * ((ntg * 2 + 7) / 8 * 8),
                 o_flag = o_nrec + (ntg * 4 + 7) / 8 * 8,
                 o_list = o_flag + ((size_t)Sg * nchunks * 4 + 4 + 7) / 8 * 8,      // block flags, then the length of the list
                 o_seg = o_list + ((size_t)Sg * nchunks * 4 + 7) / 8 * 8, misc_bytes = o_seg + (size_t)Sg * nseg * np1 * 8;
    int rc;
    if ((rc = sm_count(ctx))) return rc;
    const unsigned fb_grid = (unsigned)(ctx->sm_count * CUDE_MIN_BLOCKS);          // resident blocks walking the fallback list
    if ((rc = ensure(ctx, set.misc, misc_bytes))) return rc;
    if ((rc = ensure(ctx, set.part, 2 * rowsB * np1 * sizeof(double)))) return rc;
    size_t sort_bytes = 0;
    CU_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (unsigned short*)nullptr, (unsigned short*)nullptr,
                                                (unsigned int*)nullptr, (unsigned int*)nullptr, ntg, 0, 16, st));
    if ((rc = ensure(ctx, set.map, sort_bytes))) return rc;        // radix-sort workspace
    char* const misc = (char*)set.misc.p;
    double* const pB = (double*)set.part.p;
    double* const pC = pB + rowsB * np1;
    const int nacc = 2 * net->width + (net->depth - 1) * net->width * (net->width + 1) + net->width + 1;
    const size_t smem1 = sizeof(double) * eval_smem_doubles(P, nacc, K, M, B, false, false, true);
    const size_t smemA = sizeof(double) * adj_smem_doubles(P, B, fbwd, wc);
    if ((rc = prep_kernel(ctx, (const void*)sk.k1, B, smem1))) return rc;
    if ((rc = prep_kernel(ctx, (const void*)sk.ka, B, smemA))) return rc;
    if ((rc = prep_kernel(ctx, (const void*)fused, B, smem_fused))) return rc;
    if (wc && (rc = wconst_upload(ctx, a.neural, n_w))) return rc;
    for (int g0 = 0; g0 < S; g0 += Sg) {
        const int ns = (S - g0 < Sg) ? S - g0 : Sg;
        const unsigned nblk = (unsigned)((size_t)ns * nchunks);
        EvalArgs e = a;
        e.n_starts = ns;
        e.neural = a.neural + (size_t)g0 * a.neural_stride;
        e.wc_base = a.wc_base + (long long)g0 * a.neural_stride;
        e.cond = a.cond + (size_t)g0 * N;
        e.sse_out = a.sse_out ? a.sse_out + (size_t)g0 * N : nullptr;
        e.g_cond = a.g_cond + (size_t)g0 * N;
        e.order = nullptr;
        e.partials = nullptr;
        unsigned int* const keys_raw = (unsigned int*)(misc + o_keys);
        unsigned int* const keys_sorted = keys_raw + (ntg * 4 + 7) / 8 * 2;
        unsigned short* const k16_raw = (unsigned short*)(misc + o_k16);
        unsigned short* const k16_sorted = k16_raw + (ntg * 2 + 7) / 8 * 4;
        e.keys_out = keys_raw;
        e.keys16_out = k16_raw;
        e.sp_rec = (double*)set.rec.p;
        e.sp_res = (double*)(misc + o_res); e.sp_beta = (double*)(misc + o_beta); e.sp_sse = (double*)(misc + o_sse);
        e.sp_nrec = (int*)(misc + o_nrec); e.sp_blkflag = (int*)(misc + o_flag);
        e.sp_blklist = (int*)(misc + o_list); e.sp_blkcount = e.sp_blkflag + (size_t)Sg * nchunks;
        CU_TRY(ctx, cudaMemsetAsync(e.sp_blkflag, 0, ((size_t)Sg * nchunks + 1) * sizeof(int), st));
        CU_TRY(ctx, cudaMemsetAsync(pC, 0, (size_t)ns * nB * np1 * sizeof(double), st));
        sk.k1<<<nblk, B, smem1, st>>>(e);                                     // forward solve, step records, keys
        CU_TRY(ctx, cudaGetLastError());
        {   // every start's individuals by accepted steps: ONE stable sort of the group on the key (start, steps) — the starts stay
            // contiguous blocks of N, ties keep the individuals' order (64 sorts per step cost 5 % of a 1/8 shard's step)
            size_t b = set.map.cap;
            int bits = 8;
            while ((1 << (bits - 8)) < ns) ++bits;
            CU_TRY(ctx, cub::DeviceRadixSort::SortPairs(set.map.p, b, k16_raw, k16_sorted, keys_raw, keys_sorted, (size_t)N * ns, 0, bits, st));
            *launches += 2 + (bits + 7) / 8;                                  // cub one-sweep: histogram, scan, one sweep per 8 key bits
        }
        AdjArgs aa{};
        aa.pop = a.pop; aa.n_starts = ns; aa.nchunks = nchunks; aa.neural = e.neural; aa.neural_stride = a.neural_stride; aa.wc_base = e.wc_base;
        aa.sp_rec = e.sp_rec; aa.sp_res = e.sp_res; aa.sp_nrec = e.sp_nrec; aa.sp_beta = e.sp_beta; aa.sp_sse = e.sp_sse;
        aa.order = keys_sorted; aa.cond_scale = a.cond_scale; aa.g_cond = e.g_cond; aa.partials = pB;
        sk.ka<<<nblk, B, smemA, st>>>(aa);                                    // adjoint sweep in sorted order
        CU_TRY(ctx, cudaGetLastError());
        EvalArgs f = e;
        f.partials = pC; f.only_flag = e.sp_nrec; f.counters = nullptr; f.keys_out = nullptr; f.sse_out = nullptr;
        fused<<<(nblk < fb_grid ? nblk : fb_grid), B, smem_fused, st>>>(f);   // trajectories beyond SPLIT_CAP steps (list of flagged blocks)
        CU_TRY(ctx, cudaGetLastError());
        *launches += 3;
        if (d_sums_out) {
            double* const d_seg = (double*)(misc + o_seg);
            cude_reduce_rows<<<dim3((unsigned)nseg, (unsigned)ns), RED_T, 0, st>>>(nullptr, 0, pB, pC, nB, np1, nseg * np1, d_seg);
            cude_reduce_rows<<<dim3(1, (unsigned)ns), RED_T, 0, st>>>(d_seg, nseg, nullptr, nullptr, 0, np1, np1, d_sums_out + (size_t)g0 * np1);
            CU_TRY(ctx, cudaGetLastError());
            *launches += 2;
        }
    }
    if (wc && (rc = wconst_used(ctx))) return rc;
    return CUDE_OK;
}


// ---- latency form of loss + full gradient: one warp per trajectory (cude_warp.cuh), fused-kernel fallback for solves longer
// than WARP_CAP accepted steps, per-start reduction of the trajectory rows ----
typedef void (*warp_kernel_t)(const WarpArgs);
typedef void (*warp_kernel_t)(const WarpArgs);
static bool smem_fits_warp(int P, int K, int M) { return sizeof(double) * warp_smem_doubles(P, K, M, 8) <= (size_t)200 * 1024; }
// lanes per trajectory of the latency kernels: 32 for the smallest batches (one trajectory per warp: nothing but its own
// latency), 8 above (four trajectories per warp: a quarter of the repeated state arithmetic, four times the trajectories per wave)
static int warp_lanes(size_t ntraj) {
    static const int forced = [] { const char* e = getenv("CUDE_WARP_LANES"); return e ? atoi(e) : 0; }();   // tuning override, read once
    if (forced == 8 || forced == 32) return forced;
    return ntraj <= CUDE_WARP32_MAX_TRAJ ? 32 : 8;
}
template <bool GRAD>
static warp_kernel_t select_warp_kernel(const cude_net* net, int G) {
    if (!(net->depth == 2 && net->width == 4)) return nullptr;
    if (net->n_in == 2) return G == 8 ? cude_warp_kernel<NetShape<2, 2, 4>, GRAD, 8> : cude_warp_kernel<NetShape<2, 2, 4>, GRAD, 32>;
    if (net->n_in == 3) return G == 8 ? cude_warp_kernel<NetShape<3, 2, 4>, GRAD, 8> : cude_warp_kernel<NetShape<3, 2, 4>, GRAD, 32>;
    return nullptr;
}
static int run_warp(cude_ctx* ctx, const cude_net* net, const EvalArgs& a, int B, int nchunks, eval_kernel_t fused, size_t smem_fused,
                    double* d_sums_out, int* launches) {
    const int P = cude_net_nparams(net), np1 = P + 1, N = a.pop.n_ind, S = a.n_starts, nw = B / 32;
    const size_t ntraj = (size_t)N * S, nblk = (size_t)S * nchunks;
    const int G = warp_lanes(ntraj), tpb = WARP_TPB * (32 / G);           // lanes per trajectory, trajectories per block
    warp_kernel_t kw = select_warp_kernel<true>(net, G);
    if (!kw) return fail(ctx, CUDE_EUNSUPPORTED, "cude_eval_dev: network shape not compiled in");
    int rc;
    if ((rc = sm_count(ctx))) return rc;
    cude_ctx::SplitSet& set = ctx->sp[0];
    const size_t o_ovf = 0, o_flag = o_ovf + (ntraj * 4 + 7) / 8 * 8, o_list = o_flag + ((nblk + 1) * 4 + 7) / 8 * 8,
                 misc_bytes = o_list + (nblk * 4 + 7) / 8 * 8;
    if ((rc = ensure(ctx, set.misc, misc_bytes))) return rc;
    if ((rc = ensure(ctx, set.part, nblk * nw * np1 * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->partials, ntraj * np1 * sizeof(double)))) return rc;
    char* const misc = (char*)set.misc.p;
    WarpArgs w{};
    w.pop = a.pop; w.n_starts = S; w.neural = a.neural; w.neural_stride = a.neural_stride; w.cond = a.cond;
    w.abstol = a.abstol; w.reltol = a.reltol; w.maxiters = a.maxiters; w.cond_scale = a.cond_scale;
    w.sse_out = a.sse_out; w.g_cond = a.g_cond; w.rows = (double*)ctx->partials.p; w.counters = a.counters;
    w.ovf = (int*)(misc + o_ovf); w.blkflag = (int*)(misc + o_flag); w.blkcount = w.blkflag + nblk; w.blklist = (int*)(misc + o_list);
    w.fb_block = B; w.nchunks = nchunks;
    const size_t smem_w = sizeof(double) * warp_smem_doubles(P, a.pop.max_knots, a.pop.max_obs, G);
    if (smem_w > 200 * 1024) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: too many knots / observations for the warp kernel");
    if ((rc = prep_kernel(ctx, (const void*)kw, 32 * WARP_TPB, smem_w))) return rc;
    if ((rc = prep_kernel(ctx, (const void*)fused, B, smem_fused))) return rc;
    cudaStream_t st = ctx->stream;
    CU_TRY(ctx, cudaMemsetAsync(w.blkflag, 0, (nblk + 1) * sizeof(int), st));
    kw<<<(unsigned)((ntraj + tpb - 1) / tpb), 32 * WARP_TPB, smem_w, st>>>(w);
    CU_TRY(ctx, cudaGetLastError());
    EvalArgs f = a;                               // trajectories beyond WARP_CAP steps: the fused kernel on the flagged blocks
    f.partials = (double*)set.part.p; f.only_flag = w.ovf; f.sp_blkflag = w.blkflag; f.sp_blklist = w.blklist; f.sp_blkcount = w.blkcount;
    f.counters = nullptr; f.keys_out = nullptr; f.sse_out = nullptr; f.order = nullptr; f.wc_base = 0;
    const unsigned fb_grid = (unsigned)(ctx->sm_count * CUDE_MIN_BLOCKS);
    fused<<<(nblk < fb_grid ? (unsigned)nblk : fb_grid), B, smem_fused, st>>>(f);
    CU_TRY(ctx, cudaGetLastError());
    *launches += 2;
    if (d_sums_out) {
        cude_warp_reduce<<<(unsigned)S, RED_T, 0, st>>>(w.rows, N, (const double*)set.part.p, w.blkflag, nchunks, nw, np1, d_sums_out);
        CU_TRY(ctx, cudaGetLastError());
        ++*launches;
    }
    return CUDE_OK;
}

// the forward pass alone in the latency form (loss-only calls of small batches): per-trajectory sse, then the per-start sums
static int run_warp_loss(cude_ctx* ctx, const cude_net* net, const EvalArgs& a, double* d_sse, double* d_sums_out, int* launches) {
    const int P = cude_net_nparams(net), np1 = P + 1, N = a.pop.n_ind, S = a.n_starts;
    const size_t ntraj = (size_t)N * S;
    const int G = warp_lanes(ntraj), tpb = WARP_TPB * (32 / G);
    warp_kernel_t kw = select_warp_kernel<false>(net, G);
    if (!kw) return fail(ctx, CUDE_EUNSUPPORTED, "cude_eval_dev: network shape not compiled in");
    WarpArgs w{};
    w.pop = a.pop; w.n_starts = S; w.neural = a.neural; w.neural_stride = a.neural_stride; w.cond = a.cond;
    w.abstol = a.abstol; w.reltol = a.reltol; w.maxiters = a.maxiters; w.cond_scale = a.cond_scale;
    w.sse_out = d_sse; w.counters = a.counters;
    const size_t smem_w = sizeof(double) * warp_smem_doubles(P, a.pop.max_knots, a.pop.max_obs, G);
    int rc;
    if ((rc = prep_kernel(ctx, (const void*)kw, 32 * WARP_TPB, smem_w))) return rc;
    kw<<<(unsigned)((ntraj + tpb - 1) / tpb), 32 * WARP_TPB, smem_w, ctx->stream>>>(w);
    CU_TRY(ctx, cudaGetLastError());
    ++*launches;
    if (d_sums_out) {
        const int wpb = 8;
        cude_sum_sse<<<(S + wpb - 1) / wpb, wpb * 32, 0, ctx->stream>>>(d_sse, N, S, np1, d_sums_out);
        CU_TRY(ctx, cudaGetLastError());
        ++*launches;
    }
    return CUDE_OK;
}

extern "C" int cude_eval_dev(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts_in,
                             int n_starts, const double* d_neural, long long neural_stride, const double* d_cond,
                             int want_grad, double cond_scale,
                             double* d_sse_out, double* d_sums_out, double* d_g_cond) {
    return eval_dev_impl(ctx, pop, net, opts_in, n_starts, d_neural, neural_stride, d_cond, want_grad, cond_scale,
                         d_sse_out, d_sums_out, d_g_cond, nullptr);
}

static int eval_dev_impl(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts_in,
                         int n_starts, const double* d_neural, long long neural_stride, const double* d_cond,
                         int want_grad, double cond_scale,
                         double* d_sse_out, double* d_sums_out, double* d_g_cond, double* d_yhat) {
    if (!ctx || !pop || !net || !d_neural || !d_cond) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: NULL argument");
    if (pop->ctx != ctx) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: population belongs to another context");
    if (n_starts < 1) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: n_starts < 1");
    cude_opts o;
    if (opts_in) o = *opts_in; else cude_default_opts(&o);
    if (!(o.abstol > 0.0) || !(o.reltol > 0.0) || o.maxiters < 1) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: bad solver options");
    if (o.balance < 0 || o.balance > 4) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: balance must be 0 (automatic), 1 (history), 2 (exact), 3 (fused kernel) or 4 (warp per trajectory)");
    if (o.precision < 0 || o.precision > 2) return fail(ctx, CUDE_EUNSUPPORTED, "cude_eval_dev: precision must be 0 (FP64), 1 (FP32 network, FP64 integrator) or 2 (FP64 forward pass, FP32 network in the adjoint)");
    const bool mixed = o.precision == 1;
    const bool fbwd = o.precision == 2;
    const int P = cude_net_nparams(net);
    if (P < 0) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: bad network description");
    if (net->n_in == 3 && !pop->dev.cov) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: 3-input network needs a population with a covariate");
    const bool grad = want_grad != 0;
    // beta-only gradient (want_grad == 1: d/d cond, no network gradient): one forward-sensitivity column inside the
    // loss kernel instead of the adjoint sweep (src/parameter-estimation.jl:272-307, evaluate_model :406-433)
    const bool bsens = want_grad == 1 && !mixed && CUDE_BETA_FORWARD_SENSITIVITY;
    const bool adj = grad && !bsens;
    // weights as uniform operands from constant memory when the call's weights fit (see CW_CONST in cude_kernels.cuh)
    const size_t n_w = neural_stride == 0 ? (size_t)P : (size_t)neural_stride * (n_starts - 1) + P;
    // Only the FP64 adjoint kernel: there ptxas keeps the hoisted weights in uniform registers (+2.7 %); in the loss-only
    // instantiation it put them into 74 vector registers and the kernel lost 8 % (profiles/README.md).
    const bool wc = CUDE_WEIGHTS_IN_CONSTANT_MEMORY && (adj || (CUDE_WC_LOSS && !grad) || (CUDE_WC_BSENS && bsens)) && !mixed &&
                    n_w <= CUDE_WCONST_DOUBLES;
    eval_kernel_t kern = select_kernel(net, adj, mixed, bsens, fbwd && adj, wc);
    if (!kern) return fail(ctx, CUDE_EUNSUPPORTED, "cude_eval_dev: network shape not compiled in (available: n_in 2|3, depth 2, width 4)");
    if (neural_stride != 0 && neural_stride < P) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: neural_stride < n_params");
    CU_TRY(ctx, cudaSetDevice(ctx->device));

    const int N = pop->n_ind, np1 = P + 1;
    // flat indexing when every trajectory shares one network and no per-start network gradient is needed
    const bool want_neural_grad = grad && (want_grad & 2);
    const bool flat = (neural_stride == 0) && !want_neural_grad;
    const int B = choose_block(&o, N, flat);
    if (B < 32 || B > CUDE_MAX_THREADS || (B & 31)) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: block must be a multiple of 32, at most 128");
    const long long ntraj = (long long)N * n_starts;
    const int nchunks = (N + B - 1) / B;
    const long long nblocks = flat ? (ntraj + B - 1) / B : (long long)n_starts * nchunks;
    if (nblocks > 0x7fffffffLL) return fail(ctx, CUDE_EINVAL, "cude_eval_dev: too many blocks; split the call");

    int rc;
    const bool bal = balance_applies(o, pop, grad, flat, n_starts);
    const bool bal_own = bal && ctx->chunk_mode == 0;    // a pipelined host call prepares / finishes around its chunks
    if (bal_own && (rc = balance_prepare(ctx, pop, n_starts))) return rc;
    if ((rc = ensure(ctx, ctx->counters, 3 * sizeof(unsigned long long)))) return rc;
    double* d_partials = nullptr;
    if (!flat && d_sums_out) {
        if ((rc = ensure(ctx, ctx->partials, (size_t)nblocks * (B / 32) * np1 * sizeof(double)))) return rc;
        d_partials = (double*)ctx->partials.p;
    }
    double* d_sse = d_sse_out;
    if (flat && d_sums_out && !d_sse) {
        if ((rc = ensure(ctx, ctx->scratch, (size_t)ntraj * sizeof(double)))) return rc;
        d_sse = (double*)ctx->scratch.p;
    }

    EvalArgs a{};
    a.pop = pop->dev;
    a.n_starts = n_starts;
    a.neural = d_neural;
    a.neural_stride = neural_stride;
    a.cond = d_cond;
    a.abstol = o.abstol; a.reltol = o.reltol; a.maxiters = o.maxiters;
    // shared network, flat indexing: individual-major warps (32 starts of one individual per warp, cude_kernels.cuh) when the
    // individuals differ in length (ragged population: config 2, 0.81 -> 0.56 ms) or the starts are dense (profile grids of
    // >= 4096 points: 1.68 -> 1.57 ms; a 1000-point grid over 25 units of log beta is better off start-major: 0.21 vs 0.24 ms)
    a.flat = !flat ? 0 : (((pop->ragged && n_starts >= 32) || n_starts >= CUDE_FLAT_IMAJOR_MIN_STARTS) ? 2 : 1);
    a.nchunks = nchunks;
    a.cond_scale = cond_scale;
    a.sse_out = d_sse;
    a.partials = d_partials;
    a.g_cond = grad ? d_g_cond : nullptr;
    a.counters = (unsigned long long*)ctx->counters.p;
    a.order = bal ? ctx->bal_order : nullptr;
    a.keys_out = bal ? ctx->bal_keys_out : nullptr;
    a.yhat_out = grad ? nullptr : d_yhat;

    const int K = pop->max_knots, M = pop->max_obs;
    const int nacc = 2 * net->width + (net->depth - 1) * net->width * (net->width + 1) + net->width + 1;
    const size_t smem = sizeof(double) * eval_smem_doubles(P, nacc, K, M, B, adj, mixed || (fbwd && adj), bsens);
    if ((rc = prep_kernel(ctx, (const void*)kern, B, smem))) return rc;

    // loss + full gradient through the split pipeline only on request (opts.split = 2): measured 5 - 15 % slower than the
    // fused kernel on B200 (profiles/README.md, round 2), kept as a parity-tested alternative
    const bool use_split = adj && !flat && !mixed && want_neural_grad && d_g_cond && o.split == 2;
    // the two-kernel gradient whose adjoint runs every start's trajectories sorted by their exact step count: on request
    // (opts.balance = 2) or automatically for large populations (opts.balance = 0).  The choice depends on the population
    // only, never on the number of starts, so that a start's sums do not depend on how a batch is split into calls.
    const bool use_exact = !use_split && adj && !flat && !mixed && want_neural_grad && d_g_cond && N < (1 << 24) &&
                           (o.balance == 2 || (o.balance == 0 && N >= CUDE_EXACT_MIN_IND));
    // the latency form (one warp per trajectory): on request (opts.balance = 4) or automatically for small batches
    const bool use_warp = !use_split && !use_exact && adj && !flat && !mixed && !fbwd && want_neural_grad && d_g_cond && !bal &&
                          smem_fits_warp(P, K, M) &&     // very long knot / observation lists: the thread-per-trajectory kernels
                          (o.balance == 4 || (o.balance == 0 && ntraj <= CUDE_WARP_MAX_TRAJ));
    const bool use_warp_loss = !grad && !mixed && !d_yhat && smem_fits_warp(P, K, M) &&
                               (o.balance == 4 || (o.balance == 0 && ntraj <= CUDE_WARP_MAX_TRAJ));
    if (use_warp_loss && !d_sse && d_sums_out) {
        if ((rc = ensure(ctx, ctx->scratch, (size_t)ntraj * sizeof(double)))) return rc;
        d_sse = (double*)ctx->scratch.p;
    }
    if (ctx->chunk_mode != 2) CU_TRY(ctx, cudaMemsetAsync(ctx->counters.p, 0, 3 * sizeof(unsigned long long), ctx->stream));
    if (d_sums_out && (flat || !want_neural_grad))
        CU_TRY(ctx, cudaMemsetAsync(d_sums_out, 0, (size_t)np1 * n_starts * sizeof(double), ctx->stream));
    if (ctx->chunk_mode != 2) CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    (void)cudaGetLastError();   // drop stale non-sticky errors of other runtime users in this process (e.g. torch)
    int launches = 0;
    if (use_split) {
        if ((rc = run_split(ctx, net, a, B, nchunks, fbwd, wc, n_w, kern, smem, d_sums_out, &launches))) return rc;
    } else if (use_exact) {
        if ((rc = run_exact(ctx, net, a, B, nchunks, fbwd, wc, n_w, kern, smem, d_sums_out, &launches))) return rc;
    } else if (use_warp_loss && d_sse) {
        if ((rc = run_warp_loss(ctx, net, a, d_sse, d_sums_out, &launches))) return rc;
    } else if (use_warp) {
        // (fallback through the shared-memory-weights instantiation: no constant-memory upload on the latency path)
        if ((rc = run_warp(ctx, net, a, B, nchunks, select_kernel(net, true, false, false, false, false), smem, d_sums_out, &launches))) return rc;
    } else {
        if (wc && (rc = wconst_upload(ctx, d_neural, n_w))) return rc;
        kern<<<(unsigned)nblocks, B, smem, ctx->stream>>>(a);
        CU_TRY(ctx, cudaGetLastError());
        if (wc && (rc = wconst_used(ctx))) return rc;
        launches = 1;
    }
    if (d_sums_out && !use_split && !use_exact && !use_warp && !(use_warp_loss && d_sse)) {
        if (flat) {
            const int wpb = 8;
            cude_sum_sse<<<(n_starts + wpb - 1) / wpb, wpb * 32, 0, ctx->stream>>>(d_sse, N, n_starts, np1, d_sums_out);
        } else {
            reduce_partials(ctx->stream, d_partials, nchunks * (B / 32), n_starts, np1, want_neural_grad ? np1 : 1, d_sums_out);
        }
        CU_TRY(ctx, cudaGetLastError());
        ++launches;
    }
    if (bal_own && (rc = balance_finish(ctx, pop, n_starts, &launches))) return rc;
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    if (ctx->chunk_mode != 2) ctx->stats = cude_stats{};
    ctx->stats.n_traj += (unsigned long long)ntraj;
    ctx->stats.launches += launches;
    ctx->stats_pending = true;
    return CUDE_OK;
}

// ---------------------------------------------------------------- host-buffer entry points
// Chunks of starts for a host-buffer call: one chunk below ~2 M trajectories, otherwise up to CUDE_MAX_CHUNKS chunks of
// whole starts (>= ~1 M trajectories = ~5 ms of kernel each), so that the H2D of chunk k+1 and the D2H of chunk k-1
// overlap the kernels of chunk k and only the first copy in and the last copy out are exposed.
static int host_chunks(int n_starts, size_t ntraj) {
    const size_t per = (size_t)1 << 20;
    size_t n = ntraj / per;
    if (n < 2) return 1;
    if (n > CUDE_MAX_CHUNKS) n = CUDE_MAX_CHUNKS;
    if (n > (size_t)n_starts) n = (size_t)n_starts;
    return (int)n;
}

// raw_sums != NULL: raw_sums[(P+1) x S] = {sum_i sse, sum_i d sse/d neural} unscaled and g_cond scaled by cond_scale
// (the sharded-population form); otherwise loss_out / g_neural / g_cond as documented for cude_loss / cude_loss_grad.
// Sharded populations (cude_multi.inl): `ld` = column stride of the host matrices cond / sse_out / g_cond (0 = n_ind:
// contiguous; a rank that owns rows [lo, hi) of a global [N_total x S] matrix passes the matrix + lo and ld = N_total),
// `n_total` > 0 = global number of individuals: the per-start sums are all-reduced over the context's communicator
// before they are read back, and means are taken over n_total.
struct HostShard { long long ld = 0; long long n_total = 0; int want_neural = -1; /* -1: network gradient iff g_neural is given */ };
static int comm_allreduce(cude_ctx* ctx, double* d_buf, size_t count);   // cude_multi.inl

static int eval_host(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                     int n_starts, const double* neural, long long neural_stride, const double* cond,
                     int want_grad, int mean_over_individuals,
                     double* sse_out, double* loss_out, double* g_neural, double* g_cond,
                     double* raw_sums = nullptr, double cond_scale = 1.0, HostShard shard = HostShard()) {
    if (!ctx || !pop || !net || !neural || !cond) return fail(ctx, CUDE_EINVAL, "cude_loss: NULL argument");
    if (n_starts < 1) return fail(ctx, CUDE_EINVAL, "cude_loss: n_starts < 1");
    const int P = cude_net_nparams(net);
    if (P < 0) return fail(ctx, CUDE_EINVAL, "cude_loss: bad network description");
    if (neural_stride != 0 && neural_stride < P) return fail(ctx, CUDE_EINVAL, "cude_loss: neural_stride < n_params");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const int N = pop->n_ind, np1 = P + 1;
    const size_t ntraj = (size_t)N * n_starts;
    const size_t n_neural = neural_stride == 0 ? (size_t)P : (size_t)neural_stride * (n_starts - 1) + P;
    int rc;
    if ((rc = ensure(ctx, ctx->neural, n_neural * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->cond, ntraj * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->sums, (size_t)np1 * n_starts * sizeof(double)))) return rc;
    if (sse_out && (rc = ensure(ctx, ctx->sse, ntraj * sizeof(double)))) return rc;
    if (g_cond && (rc = ensure(ctx, ctx->gcond, ntraj * sizeof(double)))) return rc;
    if (ctx->h_sums_cap < (size_t)np1 * n_starts) {
        if (ctx->h_sums) cudaFreeHost(ctx->h_sums);
        ctx->h_sums = nullptr; ctx->h_sums_cap = 0;
        CU_TRY(ctx, cudaMallocHost(&ctx->h_sums, (size_t)np1 * n_starts * sizeof(double)));
        ctx->h_sums_cap = (size_t)np1 * n_starts;
    }
    int wg = 0;
    if (want_grad) wg = ((shard.want_neural >= 0 ? shard.want_neural > 0 : (g_neural || raw_sums)) ? 2 : 0) | 1;
    const size_t ld = shard.ld > 0 ? (size_t)shard.ld : (size_t)N;   // host column stride
    if (ld < (size_t)N) return fail(ctx, CUDE_EINVAL, "cude_loss: host leading dimension smaller than the population");
    const bool sharded = shard.n_total > 0;
    if (sharded && shard.n_total != N && !ctx->comm)
        return fail(ctx, CUDE_EINVAL, "cude_loss: sharded call without a communicator (cude_comm_init_rank / cude_mctx_create)");
    const double n_mean = sharded ? (double)shard.n_total : (double)N;
    const double scale = raw_sums ? 1.0 : (mean_over_individuals ? 1.0 / n_mean : 1.0);
    const double cscale = raw_sums ? cond_scale : scale;
    // column blocks of a host matrix with leading dimension ld <-> the contiguous [N x ns] device block
    auto copy_in = [&](double* d, const double* h, size_t ns, cudaStream_t st) {
        return ld == (size_t)N ? cudaMemcpyAsync(d, h, ns * N * sizeof(double), cudaMemcpyHostToDevice, st)
                               : cudaMemcpy2DAsync(d, (size_t)N * sizeof(double), h, ld * sizeof(double), (size_t)N * sizeof(double), ns,
                                                   cudaMemcpyHostToDevice, st);
    };
    auto copy_out = [&](double* h, const double* d, size_t ns, cudaStream_t st) {
        return ld == (size_t)N ? cudaMemcpyAsync(h, d, ns * N * sizeof(double), cudaMemcpyDeviceToHost, st)
                               : cudaMemcpy2DAsync(h, ld * sizeof(double), d, (size_t)N * sizeof(double), (size_t)N * sizeof(double), ns,
                                                   cudaMemcpyDeviceToHost, st);
    };
    double* const d_neural = (double*)ctx->neural.p;
    double* const d_cond = (double*)ctx->cond.p;
    double* const d_sums = (double*)ctx->sums.p;
    double* const d_sse = sse_out ? (double*)ctx->sse.p : nullptr;
    double* const d_gc = g_cond ? (double*)ctx->gcond.p : nullptr;
    const int nch = host_chunks(n_starts, ntraj);
    CU_TRY(ctx, cudaMemcpyAsync(d_neural, neural, n_neural * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (nch == 1) {
        CU_TRY(ctx, copy_in(d_cond, cond, (size_t)n_starts, ctx->stream));
        rc = cude_eval_dev(ctx, pop, net, opts, n_starts, d_neural, neural_stride, d_cond, wg, cscale, d_sse, d_sums, d_gc);
        if (rc) return rc;
        if (sse_out) CU_TRY(ctx, copy_out(sse_out, d_sse, (size_t)n_starts, ctx->stream));
        if (g_cond) CU_TRY(ctx, copy_out(g_cond, d_gc, (size_t)n_starts, ctx->stream));
    } else {
        if (!ctx->s_in) {
            CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
            CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
            for (int k = 0; k < CUDE_MAX_CHUNKS; ++k) {
                CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_in[k], cudaEventDisableTiming));
                CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_comp[k], cudaEventDisableTiming));
            }
        }
        // software pipeline in launch order H2D(k+1), kernels(k+1), D2H(k): with pageable host memory the copies block
        // the host, but never before the next chunk's kernels are queued behind the running ones
        cude_opts o;
        if (opts) o = *opts; else cude_default_opts(&o);
        const bool bal = balance_applies(o, pop, want_grad != 0, neural_stride == 0 && !(wg & 2), n_starts);
        if (bal && (rc = balance_prepare(ctx, pop, n_starts))) return rc;
        const unsigned int* const ord0 = bal ? ctx->bal_order : nullptr;
        unsigned int* const keys0 = bal ? ctx->bal_keys_out : nullptr;
        auto lo = [&](int k) { return (int)((long long)n_starts * k / nch); };
        auto h2d = [&](int k) -> int {
            const size_t s0 = (size_t)lo(k), ns = (size_t)(lo(k + 1) - lo(k));
            CU_TRY(ctx, copy_in(d_cond + s0 * N, cond + s0 * ld, ns, ctx->s_in));
            CU_TRY(ctx, cudaEventRecord(ctx->ev_in[k], ctx->s_in));
            return CUDE_OK;
        };
        auto run = [&](int k) -> int {
            const int s0 = lo(k), ns = lo(k + 1) - s0;
            CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[k], 0));
            ctx->chunk_mode = k == 0 ? 1 : 2;
            ctx->bal_order = ord0 ? ord0 + (size_t)s0 * N : nullptr;
            ctx->bal_keys_out = keys0 ? keys0 + (size_t)s0 * N : nullptr;
            const int r = cude_eval_dev(ctx, pop, net, opts, ns, d_neural + (size_t)s0 * neural_stride, neural_stride,
                                        d_cond + (size_t)s0 * N, wg, cscale, d_sse ? d_sse + (size_t)s0 * N : nullptr,
                                        d_sums + (size_t)s0 * np1, d_gc ? d_gc + (size_t)s0 * N : nullptr);
            ctx->chunk_mode = 0;
            if (r) return r;
            CU_TRY(ctx, cudaEventRecord(ctx->ev_comp[k], ctx->stream));
            return CUDE_OK;
        };
        auto d2h = [&](int k) -> int {
            const size_t s0 = (size_t)lo(k), ns = (size_t)(lo(k + 1) - lo(k));
            CU_TRY(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[k], 0));
            if (sse_out) CU_TRY(ctx, copy_out(sse_out + s0 * ld, d_sse + s0 * N, ns, ctx->s_out));
            if (g_cond) CU_TRY(ctx, copy_out(g_cond + s0 * ld, d_gc + s0 * N, ns, ctx->s_out));
            return CUDE_OK;
        };
        if ((rc = h2d(0)) || (rc = run(0))) return rc;
        for (int k = 0; k < nch; ++k) {
            if (k + 1 < nch && ((rc = h2d(k + 1)) || (rc = run(k + 1)))) return rc;
            if ((rc = d2h(k))) return rc;
        }
        if (bal) {
            ctx->bal_order = ord0; ctx->bal_keys_out = keys0;
            if ((rc = balance_finish(ctx, pop, n_starts, nullptr))) return rc;
        }
    }
    // sharded population: the only exchange of the path — sum the shards' {sum sse, sum d sse/d neural} rows in place
    if (sharded && ctx->comm && ctx->comm_nranks > 1 && (rc = comm_allreduce(ctx, d_sums, (size_t)np1 * n_starts))) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(ctx->h_sums, d_sums, (size_t)np1 * n_starts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (nch > 1) CU_TRY(ctx, cudaStreamSynchronize(ctx->s_out));
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(ctx, CUDE_ECUDA, std::string("kernel failed: ") + cudaGetErrorString(e));
    }
    if (raw_sums) {
        memcpy(raw_sums, ctx->h_sums, (size_t)np1 * n_starts * sizeof(double));
        return CUDE_OK;
    }
    for (int s = 0; s < n_starts; ++s) {
        const double* row = ctx->h_sums + (size_t)s * np1;
        const bool ok = std::isfinite(row[0]);
        if (loss_out) loss_out[s] = row[0] * scale;   // Inf stays Inf (parameter-estimation.jl:134-136)
        if (g_neural) for (int p = 0; p < P; ++p) g_neural[(size_t)s * P + p] = ok ? row[1 + p] * scale : 0.0;
        if (g_cond && !ok) for (int i = 0; i < N; ++i) g_cond[(size_t)s * ld + i] = 0.0;
    }
    return CUDE_OK;
}

// model prediction at the observation times: the `solve(...; saveat=timepoints, save_idxs=1)` of the loss (:59) by itself
extern "C" int cude_simulate(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                             int n_starts, const double* neural, long long neural_stride, const double* cond,
                             double* yhat_out, double* sse_out) {
    if (!ctx || !pop || !net || !neural || !cond || !yhat_out) return fail(ctx, CUDE_EINVAL, "cude_simulate: NULL argument");
    if (n_starts < 1) return fail(ctx, CUDE_EINVAL, "cude_simulate: n_starts < 1");
    const int P = cude_net_nparams(net);
    if (P < 0) return fail(ctx, CUDE_EINVAL, "cude_simulate: bad network description");
    if (neural_stride != 0 && neural_stride < P) return fail(ctx, CUDE_EINVAL, "cude_simulate: neural_stride < n_params");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t N = (size_t)pop->n_ind, M = (size_t)pop->max_obs, ntraj = N * n_starts;
    const size_t n_neural = neural_stride == 0 ? (size_t)P : (size_t)neural_stride * (n_starts - 1) + P;
    int rc;
    if ((rc = ensure(ctx, ctx->neural, n_neural * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->cond, ntraj * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->sse, ntraj * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->scratch, ntraj * M * sizeof(double)))) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(ctx->neural.p, neural, n_neural * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->cond.p, cond, ntraj * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(ctx->scratch.p, 0xFF, ntraj * M * sizeof(double), ctx->stream));   // NaN where nothing is observed
    rc = eval_dev_impl(ctx, pop, net, opts, n_starts, (const double*)ctx->neural.p, neural_stride, (const double*)ctx->cond.p, 0, 1.0,
                       (double*)ctx->sse.p, nullptr, nullptr, (double*)ctx->scratch.p);
    if (rc) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(yhat_out, ctx->scratch.p, ntraj * M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (sse_out) CU_TRY(ctx, cudaMemcpyAsync(sse_out, ctx->sse.p, ntraj * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CUDE_OK;
}

extern "C" int cude_loss(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                         int n_starts, const double* neural, long long neural_stride, const double* cond,
                         double* sse_out, double* loss_out) {
    return eval_host(ctx, pop, net, opts, n_starts, neural, neural_stride, cond, 0, 1, sse_out, loss_out, nullptr, nullptr);
}

extern "C" int cude_loss_grad(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                              int n_starts, const double* neural, long long neural_stride, const double* cond,
                              int mean_over_individuals,
                              double* sse_out, double* loss_out, double* g_neural, double* g_cond) {
    return eval_host(ctx, pop, net, opts, n_starts, neural, neural_stride, cond, 1, mean_over_individuals,
                     sse_out, loss_out, g_neural, g_cond);
}

extern "C" int cude_loss_grad_sums(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                                   int n_starts, const double* neural, long long neural_stride, const double* cond,
                                   double cond_scale, double* sums_out, double* g_cond) {
    if (!sums_out) return fail(ctx, CUDE_EINVAL, "cude_loss_grad_sums: sums_out is NULL");
    return eval_host(ctx, pop, net, opts, n_starts, neural, neural_stride, cond, 1, 0, nullptr, nullptr, nullptr, g_cond,
                     sums_out, cond_scale);
}

// ---------------------------------------------------------------- suppression variant
struct cude_sup_population {
    cude_ctx* ctx = nullptr;
    int n_ind = 0, n_obs = 0;
    double* d_data = nullptr;    // [M][3][N]
    double* d_obs_t = nullptr;   // [M]
    double p1 = 0, p3 = 0, iscale[3] = {1, 1, 1}, t0 = 0, tend = 0;
};

extern "C" int cude_sup_population_create(cude_ctx* ctx, int n_ind, int n_obs, const double* obs_t, const double* data,
                                          const double* p_true, const double* scale, double t0, double tend,
                                          cude_sup_population** out) {
    if (!ctx || !out || !obs_t || !data || !p_true || n_ind < 1 || n_obs < 1 || !(tend > t0))
        return fail(ctx, CUDE_EINVAL, "cude_sup_population_create: bad argument");
    *out = nullptr;
    for (int k = 0; k < n_obs; ++k) {
        if (!(obs_t[k] >= t0 && obs_t[k] <= tend) || (k > 0 && !(obs_t[k] > obs_t[k - 1])))
            return fail(ctx, CUDE_EINVAL, "cude_sup_population_create: observation times must be increasing inside tspan");
    }
    const size_t N = n_ind, M = n_obs;
    std::vector<double> h(M * 3 * N);
    double sc[3] = {0, 0, 0};
    for (size_t i = 0; i < N; ++i) {
        double mx[3] = {-1e300, -1e300, -1e300};
        for (size_t k = 0; k < M; ++k)
            for (int j = 0; j < 3; ++j) {
                const double v = data[j + 3 * (k + M * i)];
                h[(k * 3 + j) * N + i] = v;
                if (v > mx[j]) mx[j] = v;
            }
        for (int j = 0; j < 3; ++j) sc[j] += mx[j];
    }
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cude_sup_population* pop = new (std::nothrow) cude_sup_population();
    if (!pop) return fail(ctx, CUDE_ENOMEM, "out of host memory");
    pop->ctx = ctx; pop->n_ind = n_ind; pop->n_obs = n_obs; pop->p1 = p_true[0]; pop->p3 = p_true[2]; pop->t0 = t0; pop->tend = tend;
    for (int j = 0; j < 3; ++j) pop->iscale[j] = 1.0 / (scale ? scale[j] : sc[j] / (double)N);   // suppression_model.jl:125
    if (cudaMalloc(&pop->d_data, h.size() * sizeof(double)) != cudaSuccess || cudaMalloc(&pop->d_obs_t, M * sizeof(double)) != cudaSuccess) {
        cude_sup_population_destroy(pop);
        return fail(ctx, CUDE_ENOMEM, "cude_sup_population_create: cudaMalloc failed");
    }
    CU_TRY(ctx, cudaMemcpyAsync(pop->d_data, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(pop->d_obs_t, obs_t, M * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = pop;
    return CUDE_OK;
}

extern "C" int cude_sup_population_destroy(cude_sup_population* pop) {
    if (!pop) return CUDE_OK;
    if (pop->ctx) cudaSetDevice(pop->ctx->device);
    if (pop->d_data) cudaFree(pop->d_data);
    if (pop->d_obs_t) cudaFree(pop->d_obs_t);
    delete pop;
    return CUDE_OK;
}

typedef void (*sup_kernel_t)(const SupArgs);

extern "C" int cude_sup_loss_grad(cude_ctx* ctx, const cude_sup_population* pop, int depth, int width, const cude_opts* opts_in,
                                  int n_starts, const double* neural, long long neural_stride, const double* theta, double lambda,
                                  double* sse_out, double* loss_out, double* g_neural, double* g_theta) {
    if (!ctx || !pop || !neural || !theta || n_starts < 1) return fail(ctx, CUDE_EINVAL, "cude_sup_loss_grad: bad argument");
    if (pop->ctx != ctx) return fail(ctx, CUDE_EINVAL, "cude_sup_loss_grad: population belongs to another context");
    if (!(depth == 5 && width == 3)) return fail(ctx, CUDE_EUNSUPPORTED, "cude_sup_loss_grad: network shape not compiled in (available: depth 5, width 3)");
    typedef SupNet<5, 3> SN;
    cude_opts o;
    if (opts_in) o = *opts_in; else cude_default_opts(&o);
    if (!(o.abstol > 0.0) || !(o.reltol > 0.0) || o.maxiters < 1) return fail(ctx, CUDE_EINVAL, "cude_sup_loss_grad: bad solver options");
    const int P = SN::P, np1 = P + 1, N = pop->n_ind, M = pop->n_obs;
    if (neural_stride != 0 && neural_stride < P) return fail(ctx, CUDE_EINVAL, "cude_sup_loss_grad: neural_stride < n_params");
    const bool grad = (g_neural != nullptr) || (g_theta != nullptr);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    // small populations (the reference's 37 / 30 / 60 individuals): blocks of 128 threads run floor(128 / N) whole
    // starts side by side (87 - 94 % of the lanes busy instead of 58 % with one start per 64-thread block)
    const int B = o.block > 0 ? o.block : (CUDE_SUP_PACK_DEFAULT && N <= 128 ? 128 : (N > 32 ? 64 : 32));
    if (B < 32 || B > 128 || (B & 31)) return fail(ctx, CUDE_EINVAL, "cude_sup_loss_grad: block must be 32, 64, 96 or 128");
    // small populations: flat indexing over the [S x N] batch (every lane busy), at most (B-1)/N + 2 starts per block
    const int spb = ((CUDE_SUP_PACK_DEFAULT || o.block > 0) && N <= B) ? (B - 1) / N + 2 : 0;
    const int nchunks = spb > 0 ? 2 : (N + B - 1) / B, nw = spb > 0 ? 1 : B / 32;      // flat: a start lies in at most 2 blocks = 2 partial rows
    const long long nblocks = spb > 0 ? ((long long)N * n_starts + B - 1) / B : (long long)n_starts * nchunks;
    const size_t ntraj = (size_t)N * n_starts;
    const size_t n_neural = neural_stride == 0 ? (size_t)P : (size_t)neural_stride * (n_starts - 1) + P;
    int rc;
    if ((rc = ensure(ctx, ctx->counters, 3 * sizeof(unsigned long long)))) return rc;
    if ((rc = ensure(ctx, ctx->neural, n_neural * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->cond, ntraj * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->sums, (size_t)np1 * n_starts * sizeof(double)))) return rc;
    const size_t n_prow = spb > 0 ? (size_t)n_starts * 2 : (size_t)nblocks * nw;
    if ((rc = ensure(ctx, ctx->partials, n_prow * np1 * sizeof(double)))) return rc;
    if (spb > 0) CU_TRY(ctx, cudaMemsetAsync(ctx->partials.p, 0, n_prow * np1 * sizeof(double), ctx->stream));   // rows of blocks a start does not reach
    if (sse_out && (rc = ensure(ctx, ctx->sse, ntraj * sizeof(double)))) return rc;
    if (g_theta && (rc = ensure(ctx, ctx->gcond, ntraj * sizeof(double)))) return rc;
    if (ctx->h_sums_cap < (size_t)np1 * n_starts) {
        if (ctx->h_sums) cudaFreeHost(ctx->h_sums);
        ctx->h_sums = nullptr; ctx->h_sums_cap = 0;
        CU_TRY(ctx, cudaMallocHost(&ctx->h_sums, (size_t)np1 * n_starts * sizeof(double)));
        ctx->h_sums_cap = (size_t)np1 * n_starts;
    }
    CU_TRY(ctx, cudaMemcpyAsync(ctx->neural.p, neural, n_neural * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->cond.p, theta, ntraj * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SupArgs a{};
    a.n_ind = N; a.n_obs = M; a.n_starts = n_starts; a.nchunks = nchunks; a.spb = spb;
    a.obs_t = pop->d_obs_t; a.data = pop->d_data; a.p1 = pop->p1; a.p3 = pop->p3;
    for (int j = 0; j < 3; ++j) a.iscale[j] = pop->iscale[j];
    a.t0 = pop->t0; a.tend = pop->tend;
    a.neural = (const double*)ctx->neural.p; a.neural_stride = neural_stride; a.theta = (const double*)ctx->cond.p;
    a.abstol = o.abstol; a.reltol = o.reltol; a.maxiters = o.maxiters;
    a.theta_scale = 1.0 / N;
    a.sse_out = sse_out ? (double*)ctx->sse.p : nullptr;
    a.partials = (double*)ctx->partials.p;
    a.g_theta = g_theta ? (double*)ctx->gcond.p : nullptr;
    a.counters = (unsigned long long*)ctx->counters.p;
    // gradient calls: two kernels by default (opts.split = 0) — the loss kernel leaving step records (3 blocks per SM, no spills),
    // then the adjoint sweep over them; opts.split = 1 or a solve beyond SUP_REC_CAP accepted steps: the fused kernel, which replays
    const size_t smem_loss = sizeof(double) * sup_smem_doubles(P, SN::NACC, M, B, false, spb);
    const size_t smem_grad = sizeof(double) * sup_smem_doubles(P, SN::NACC, M, B, true, spb);
    if ((grad ? smem_grad : smem_loss) > 227 * 1024) return fail(ctx, CUDE_EINVAL, "cude_sup_loss_grad: too many observations for shared memory; lower opts.block");
    int n_launch = 0;
    int h_ovf = 0;
    for (int attempt = (grad && o.split != 1 && CUDE_SUP_TWO_KERNEL) ? 0 : 1; attempt < 2; ++attempt) {
        const bool two = grad && attempt == 0;
        sup_kernel_t k_main = !grad ? cude_sup_kernel<SN, SUP_LOSS> : (two ? cude_sup_kernel<SN, SUP_FWD_REC> : cude_sup_kernel<SN, SUP_FUSED>);
        sup_kernel_t k_adj = cude_sup_kernel<SN, SUP_ADJ>;
        if ((rc = prep_kernel(ctx, (const void*)k_main, B, (grad && !two) ? smem_grad : smem_loss))) return rc;   // dynamic size + a carve-out that fits all resident blocks
        if (two && (rc = prep_kernel(ctx, (const void*)k_adj, B, smem_grad))) return rc;
        if (spb > 0 && attempt == 1 && n_launch > 0) CU_TRY(ctx, cudaMemsetAsync(ctx->partials.p, 0, n_prow * np1 * sizeof(double), ctx->stream));
        CU_TRY(ctx, cudaMemsetAsync(ctx->counters.p, 0, 3 * sizeof(unsigned long long), ctx->stream));
        if (n_launch == 0) CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        (void)cudaGetLastError();   // drop stale non-sticky errors of other runtime users in this process (e.g. torch)
        // the step ring of every block of a launch lives in global scratch ([SUP_REC_CAP][23][B] doubles per block, + [3 M][B]
        // residuals in the two-kernel form), so a large batch is cut into launches whose rings fit the scratch budget
        long long per_launch = nblocks;
        int* d_ovf = nullptr;
        if (grad) {
            const size_t ring_block = (size_t)(SUP_REC_CAP * SUP_REC_ROWS + (two ? 3 * M : 0)) * B * sizeof(double);
            size_t budget = 0;
            if ((rc = split_budget(ctx, &budget))) return rc;
            per_launch = (long long)(budget / ring_block);
            if (per_launch < 1) per_launch = 1;
            if (per_launch > nblocks) per_launch = nblocks;
            if ((rc = ensure(ctx, ctx->sp[0].rec, (size_t)per_launch * ring_block))) return rc;
            a.ring = (double*)ctx->sp[0].rec.p;
            a.res_g = a.ring + (size_t)per_launch * SUP_REC_CAP * SUP_REC_ROWS * B;
            if (two) {
                if ((rc = ensure(ctx, ctx->sp[0].misc, ntraj * 12 + 16))) return rc;
                a.sp_sse = (double*)ctx->sp[0].misc.p;
                a.sp_nacc = (int*)(a.sp_sse + ntraj);
                d_ovf = a.sp_nacc + ntraj;
                a.ovf_count = d_ovf;
                CU_TRY(ctx, cudaMemsetAsync(d_ovf, 0, sizeof(int), ctx->stream));
            }
        }
        for (long long b0 = 0; b0 < nblocks; b0 += per_launch) {
            a.blk0 = (int)b0;
            const long long nb = nblocks - b0 < per_launch ? nblocks - b0 : per_launch;
            if (two) {
                SupArgs f = a;                       // forward solve: loss, records, residuals, step counts; no gradient rows
                f.partials = nullptr; f.g_theta = nullptr;
                k_main<<<(unsigned)nb, B, smem_loss, ctx->stream>>>(f);
                CU_TRY(ctx, cudaGetLastError());
                SupArgs g = a;                       // adjoint sweep: gradient rows (and the sse column), d/d theta
                g.counters = nullptr; g.sse_out = nullptr;
                k_adj<<<(unsigned)nb, B, smem_grad, ctx->stream>>>(g);
                CU_TRY(ctx, cudaGetLastError());
                n_launch += 2;
            } else {
                k_main<<<(unsigned)nb, B, grad ? smem_grad : smem_loss, ctx->stream>>>(a);
                CU_TRY(ctx, cudaGetLastError());
                ++n_launch;
            }
        }
        if (two) {
            CU_TRY(ctx, cudaMemcpyAsync(&h_ovf, d_ovf, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            if (h_ovf == 0) break;                   // otherwise: once more through the fused kernel, which replays long solves
        }
    }
    {
        reduce_partials(ctx->stream, (const double*)ctx->partials.p, nchunks * nw, n_starts, np1, grad ? np1 : 1, (double*)ctx->sums.p);
        CU_TRY(ctx, cudaGetLastError());
    }
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->stats = cude_stats{};
    ctx->stats.n_traj = (unsigned long long)ntraj;
    ctx->stats.launches = n_launch + 1;
    ctx->stats_pending = true;
    CU_TRY(ctx, cudaMemcpyAsync(ctx->h_sums, ctx->sums.p, (size_t)np1 * n_starts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (sse_out) CU_TRY(ctx, cudaMemcpyAsync(sse_out, ctx->sse.p, ntraj * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (g_theta) CU_TRY(ctx, cudaMemcpyAsync(g_theta, ctx->gcond.p, ntraj * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(ctx, CUDE_ECUDA, std::string("kernel failed: ") + cudaGetErrorString(e));
    }
    for (int s = 0; s < n_starts; ++s) {
        const double* row = ctx->h_sums + (size_t)s * np1;
        const double* w = neural + (size_t)s * neural_stride;
        const bool ok = std::isfinite(row[0]);
        double ridge = 0.0;
        for (int p = 0; p < P; ++p) ridge += w[p] * w[p];
        if (loss_out) loss_out[s] = row[0] / N + lambda * ridge;                 // suppression_model.jl:126-128
        if (g_neural) for (int p = 0; p < P; ++p) g_neural[(size_t)s * P + p] = ok ? row[1 + p] / N + 2.0 * lambda * w[p] : 0.0;
        if (g_theta && !ok) for (int i = 0; i < N; ++i) g_theta[(size_t)s * N + i] = 0.0;
    }
    return CUDE_OK;
}

// ---------------------------------------------------------------- device-resident multi-start training
extern "C" void cude_train_default_opts(cude_train_opts* t) {
    if (!t) return;
    t->adam_iters = 1000; t->adam_lr = 1e-2; t->adam_beta1 = 0.9; t->adam_beta2 = 0.999; t->adam_eps = 1e-8;   // :346-348, Optimisers.Adam
    t->lbfgs_iters = 1000; t->lbfgs_m = 10; t->g_tol = 1e-8;                                                 // Optim.LBFGS defaults
    t->c1 = 1e-4; t->rho_hi = 0.5; t->rho_lo = 0.1; t->ls_maxiter = 50;                                      // LineSearches.BackTracking (1000 iterations upstream)
    t->check_every = 16;
}

extern "C" int cude_train(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                          const cude_train_opts* topts, int n_starts, double* neural, double* cond,
                          double* objective_out, int* iters_out, int* status_out, int* evals_out) {
    if (!ctx || !pop || !net || !neural || !cond || n_starts < 1) return fail(ctx, CUDE_EINVAL, "cude_train: bad argument");
    if (pop->ctx != ctx) return fail(ctx, CUDE_EINVAL, "cude_train: population belongs to another context");
    cude_train_opts t;
    if (topts) t = *topts; else cude_train_default_opts(&t);
    if (t.adam_iters < 0 || t.lbfgs_iters < 0 || t.lbfgs_m < 1 || t.lbfgs_m > 16 || t.ls_maxiter < 1 || t.check_every < 1)
        return fail(ctx, CUDE_EINVAL, "cude_train: bad optimiser options (history length 1..16)");
    const int P = cude_net_nparams(net);
    if (P < 0) return fail(ctx, CUDE_EINVAL, "cude_train: bad network description");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t S = (size_t)n_starts, N = (size_t)pop->n_ind, D = (size_t)P + N, m = (size_t)t.lbfgs_m;
    // one allocation: doubles first, then ints
    const size_t n_d = S * P + S * N + S * (P + 1) + S * N + 6 * S * D + 2 * m * S * D + m * S + S * SC_N;
    const size_t n_i = S * IC_N + 2;
    double* base = nullptr;
    CU_TRY(ctx, cudaMalloc(&base, n_d * sizeof(double) + n_i * sizeof(int)));
    struct Free { double* p; ~Free() { if (p) cudaFree(p); } } guard{base};
    TrainArgs A{};
    A.S = n_starts; A.P = P; A.N = (int)N; A.D = (int)D; A.m = t.lbfgs_m;
    double* q = base;
    A.xt_n = q; q += S * P;
    A.xt_c = q; q += S * N;
    double* d_sums = q; q += S * (P + 1);
    double* d_gc = q; q += S * N;
    A.sums = d_sums; A.g_cond = d_gc; A.scale = 1.0 / (double)N;
    A.x = q; q += S * D; A.g = q; q += S * D; A.d = q; q += S * D; A.best_x = q; q += S * D; A.am = q; q += S * D; A.av = q; q += S * D;
    A.hs = q; q += m * S * D; A.hy = q; q += m * S * D; A.rho = q; q += m * S; A.sc = q; q += S * SC_N;
    A.ic = (int*)q; A.status_count = A.ic + S * IC_N;
    A.lr = t.adam_lr; A.b1 = t.adam_beta1; A.b2 = t.adam_beta2; A.eps = t.adam_eps;
    A.g_tol = t.g_tol; A.c1 = t.c1; A.rho_hi = t.rho_hi; A.rho_lo = t.rho_lo; A.ls_maxiter = t.ls_maxiter; A.maxiters = t.lbfgs_iters;
    cudaStream_t st = ctx->stream;
    CU_TRY(ctx, cudaMemsetAsync(base, 0, n_d * sizeof(double) + n_i * sizeof(int), st));
    CU_TRY(ctx, cudaMemcpyAsync(A.xt_n, neural, S * P * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaMemcpyAsync(A.xt_c, cond, S * N * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaMemcpy2DAsync(A.x, D * sizeof(double), A.xt_n, P * sizeof(double), P * sizeof(double), S, cudaMemcpyDeviceToDevice, st));
    CU_TRY(ctx, cudaMemcpy2DAsync(A.x + P, D * sizeof(double), A.xt_c, N * sizeof(double), N * sizeof(double), S, cudaMemcpyDeviceToDevice, st));
    {
        std::vector<double> sc(S * SC_N, 0.0);
        for (size_t s = 0; s < S; ++s) { sc[s * SC_N + SC_BESTF] = INFINITY; sc[s * SC_N + SC_ALPHA] = 1.0; sc[s * SC_N + SC_APREV] = 1.0; }
        CU_TRY(ctx, cudaMemcpyAsync(A.sc, sc.data(), sc.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        CU_TRY(ctx, cudaStreamSynchronize(st));      // sc goes out of scope
    }
    int rc, evals = 0;
    auto eval = [&]() -> int {
        ++evals;
        return eval_dev_impl(ctx, pop, net, opts, n_starts, A.xt_n, P, A.xt_c, 3, A.scale, nullptr, d_sums, d_gc, nullptr);
    };
    // ---- training step 1: Adam (:172-176) ----
    double b1t = 1.0, b2t = 1.0;
    for (int it = 0; it < t.adam_iters; ++it) {
        if ((rc = eval())) return rc;
        b1t *= t.adam_beta1; b2t *= t.adam_beta2;
        A.b1t = b1t; A.b2t = b2t;
        cude_adam_step_kernel<<<n_starts, TRAIN_T, 0, st>>>(A, 0);
        CU_TRY(ctx, cudaGetLastError());
    }
    if (t.adam_iters > 0) {
        if ((rc = eval())) return rc;
        cude_adam_step_kernel<<<n_starts, TRAIN_T, 0, st>>>(A, 1);      // best iterate -> x, xt
        CU_TRY(ctx, cudaGetLastError());
    }
    // ---- training step 2: L-BFGS with BackTracking (:178-182) ----
    int h_count[2] = {0, 0};
    if (t.lbfgs_iters > 0) {
        if ((rc = eval())) return rc;
        cude_lbfgs_step_kernel<<<n_starts, TRAIN_T, 0, st>>>(A, 1);
        CU_TRY(ctx, cudaGetLastError());
        // every start ends by itself (converged, line search exhausted after ls_maxiter trials, or lbfgs_iters accepted
        // iterations): the bound below is the worst case, the status check leaves the loop long before
        const long long max_micro = (long long)t.lbfgs_iters * (t.ls_maxiter + 1) + t.check_every;
        for (long long k = 0; k < max_micro; ++k) {
            if ((rc = eval())) return rc;
            const bool check = ((k + 1) % t.check_every == 0);
            if (check) CU_TRY(ctx, cudaMemsetAsync(A.status_count, 0, 2 * sizeof(int), st));
            cude_lbfgs_step_kernel<<<n_starts, TRAIN_T, 0, st>>>(A, 0);
            CU_TRY(ctx, cudaGetLastError());
            if (check) {       // 8 bytes every `check_every` micro-steps: are any starts still searching?
                CU_TRY(ctx, cudaMemcpyAsync(h_count, A.status_count, sizeof h_count, cudaMemcpyDeviceToHost, st));
                CU_TRY(ctx, cudaStreamSynchronize(st));
                if (h_count[0] == 0) break;
            }
        }
    }
    // ---- results ----
    std::vector<double> h_x(S * D), h_sc(S * SC_N);
    std::vector<int> h_ic(S * IC_N);
    CU_TRY(ctx, cudaMemcpyAsync(h_x.data(), A.x, S * D * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaMemcpyAsync(h_sc.data(), A.sc, S * SC_N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaMemcpyAsync(h_ic.data(), A.ic, S * IC_N * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    for (size_t s = 0; s < S; ++s) {
        memcpy(neural + s * P, h_x.data() + s * D, P * sizeof(double));
        memcpy(cond + s * N, h_x.data() + s * D + P, N * sizeof(double));
        if (objective_out) objective_out[s] = t.lbfgs_iters > 0 ? h_sc[s * SC_N + SC_FX] : h_sc[s * SC_N + SC_BESTF];
        if (iters_out) iters_out[s] = h_ic[s * IC_N + IC_ITERS];
        if (status_out) status_out[s] = h_ic[s * IC_N + IC_STATUS];
    }
    if (evals_out) *evals_out = evals;
    return CUDE_OK;
}

// ---------------------------------------------------------------- elementary-function probe (tests)
extern "C" int cude_math_probe(cude_ctx* ctx, int which, int n, const double* x, double* y) {
    if (!ctx || !x || !y || n < 1 || which < 0 || which > 6) return fail(ctx, CUDE_EINVAL, "cude_math_probe: bad argument");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = ensure(ctx, ctx->scratch, 2 * (size_t)n * sizeof(double));
    if (rc) return rc;
    double* dx = (double*)ctx->scratch.p;
    double* dy = dx + n;
    CU_TRY(ctx, cudaMemcpyAsync(dx, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    (void)cudaGetLastError();
    cude_math_probe_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(which, n, dx, dy);
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaMemcpyAsync(y, dy, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CUDE_OK;
}

// ---------------------------------------------------------------- device-resident Adam step
extern "C" int cude_adam_dev(cude_ctx* ctx, long long n, double* d_x, const double* d_g, double* d_m, double* d_v,
                             double lr, double beta1, double beta2, double eps, int t, double grad_scale,
                             const double* d_row_flag, long long row_len, long long flag_stride) {
    if (!ctx || !d_x || !d_g || !d_m || !d_v || n < 1 || t < 1) return fail(ctx, CUDE_EINVAL, "cude_adam_dev: bad argument");
    if (d_row_flag && (row_len < 1 || flag_stride < 1)) return fail(ctx, CUDE_EINVAL, "cude_adam_dev: bad row description");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const double b1t = std::pow(beta1, t), b2t = std::pow(beta2, t);
    const int threads = 256;
    const long long blocks = (n + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return fail(ctx, CUDE_EINVAL, "cude_adam_dev: too many elements");
    (void)cudaGetLastError();
    cude_adam_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(n, d_x, d_g, d_m, d_v, lr, beta1, beta2, eps, b1t, b2t, grad_scale,
                                                                     d_row_flag, row_len, flag_stride);
    CU_TRY(ctx, cudaGetLastError());
    return CUDE_OK;
}

// ---------------------------------------------------------------- FP64 peak
extern "C" int cude_measure_fp64_peak(cude_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return CUDE_EINVAL;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 4096;
    int rc = ensure(ctx, ctx->scratch, (size_t)threads * blocks * sizeof(double));
    if (rc) return rc;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        (void)cudaGetLastError();
        cude_dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((double*)ctx->scratch.p, iters, 0.999999, 1e-9);
        CU_TRY(ctx, cudaGetLastError());
        CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double flops = 2.0 * 64.0 * (double)iters * threads * blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    ctx->stats_pending = false;
    *tflops = best;
    return CUDE_OK;
}

// FP32 FMA peak (TFLOP/s) and MUFU peak (G transcendental ops / s): denominators of the FP32-network modes
extern "C" int cude_measure_fp32_peak(cude_ctx* ctx, double* tflops, double* mufu_gops) {
    if (!ctx || !tflops) return CUDE_EINVAL;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 4096;
    int rc = ensure(ctx, ctx->scratch, (size_t)threads * blocks * sizeof(double));
    if (rc) return rc;
    double best[2] = {0.0, 0.0};
    for (int which = 0; which < 2; ++which) {
        if (which == 1 && !mufu_gops) break;
        for (int rep = 0; rep < 5; ++rep) {
            CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
            (void)cudaGetLastError();
            if (which == 0) cude_ffma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((float*)ctx->scratch.p, iters, 0.999999f, 1e-6f);
            else cude_mufu_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((float*)ctx->scratch.p, iters);
            CU_TRY(ctx, cudaGetLastError());
            CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
            CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            float ms = 0.f;
            CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            const double ops = 64.0 * (double)iters * threads * blocks;      // FFMA resp. ex2 instructions (thread level)
            const double v = (which == 0 ? 2.0 * ops / 1e12 : ops / 1e9) / (ms * 1e-3);
            if (rep > 0 && v > best[which]) best[which] = v;
        }
    }
    ctx->stats_pending = false;
    *tflops = best[0];
    if (mufu_gops) *mufu_gops = best[1];
    return CUDE_OK;
}

// diagnostic: DFMA rate with three register operands per instruction
extern "C" int cude_measure_fp64_peak_rrr(cude_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return CUDE_EINVAL;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 4096;
    int rc = ensure(ctx, ctx->scratch, (size_t)threads * blocks * sizeof(double));
    if (rc) return rc;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        (void)cudaGetLastError();
        cude_dfma_peak_rrr_kernel<<<blocks, threads, 0, ctx->stream>>>((double*)ctx->scratch.p, iters, 0.999999, 1e-9);
        CU_TRY(ctx, cudaGetLastError());
        CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double tf = 2.0 * 64.0 * (double)iters * threads * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    ctx->stats_pending = false;
    *tflops = best;
    return CUDE_OK;
}

#include "cude_multi.inl"
