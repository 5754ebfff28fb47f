// =====================================================================================
// cude_multi.inl — the multi-GPU half of the C ABI (included at the end of cude_api.cu).
//
// The reference loops over individuals (src/parameter-estimation.jl:126-140), over initial guesses (:362-366), over
// the selected starts (:374-376), over individuals of a beta-only fit (:272-288) and over profile grid points
// (src/likelihood-profiles.jl:11-14) on one CPU thread.  All of these are independent trajectories, so on one B200 box
//   * starts / grid points shard over the GPUs with NO communication           (CUDE_SHARD_STARTS), and
//   * the individuals of a large population shard over the GPUs with ONE exchange: the all-reduce of the per-start
//     sums {sum_i sse_i, sum_i d sse_i / d neural} — (P+1) x S doubles, 19 KB for 64 starts — over NCCL / NVLink,
//     in place on the buffer the reduction kernel wrote and on the stream it ran on    (CUDE_SHARD_INDIVIDUALS).
// Two host models are served:
//   * one process per GPU (torchrun, MPI, Julia Distributed): cude_comm_get_unique_id on rank 0, the host broadcasts
//     the 128 bytes, every rank calls cude_comm_init_rank on its context; cude_loss_sharded / cude_loss_grad_sharded /
//     cude_allreduce_dev are then collective calls;
//   * one process driving all GPUs (a plain Julia session): cude_mctx_create(n_gpus, device_ids) owns one context, one
//     stream and one host worker thread per device (ncclCommInitAll for the communicator) and cude_mloss /
//     cude_mloss_grad take the *global* host matrices.
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy already in the process if the host loaded one — e.g.
// torch's — else the system library), so that single-GPU hosts need no NCCL at all.
// =====================================================================================
#include <dlfcn.h>
#include <condition_variable>
#include <functional>
#include <memory>
#include <thread>

// ---------------------------------------------------------------- NCCL, loaded on demand
namespace cude_nccl {
typedef void* comm_t;
struct unique_id { char internal[CUDE_UNIQUE_ID_BYTES]; };
static_assert(CUDE_UNIQUE_ID_BYTES == 128, "NCCL_UNIQUE_ID_BYTES");
enum { kSuccess = 0, kSum = 0, kFloat64 = 8 };   // ncclSuccess, ncclSum, ncclFloat64 (stable since NCCL 2.0)
struct Api {
    void* handle = nullptr;
    int version = 0;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};
static Api g_api;
static std::mutex g_api_mutex;

static const Api* load(std::string* why) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (g_api.handle) return &g_api;
    if (!g_api.err.empty()) { if (why) *why = g_api.err; return nullptr; }
    const char* names[] = {getenv("CUDE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        g_api.err = std::string("NCCL not found (dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "?") + "); set CUDE_NCCL_LIB";
        if (why) *why = g_api.err;
        return nullptr;
    }
    Api a;
    a.handle = h;
#define CUDE_NCCL_SYM(field, name)                                                     \
    *(void**)(&a.field) = dlsym(h, name);                                               \
    if (!a.field) { g_api.err = std::string("NCCL symbol missing: ") + name; if (why) *why = g_api.err; dlclose(h); return nullptr; }
    CUDE_NCCL_SYM(GetVersion, "ncclGetVersion")
    CUDE_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    CUDE_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    CUDE_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    CUDE_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    CUDE_NCCL_SYM(AllReduce, "ncclAllReduce")
    CUDE_NCCL_SYM(GroupStart, "ncclGroupStart")
    CUDE_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    CUDE_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef CUDE_NCCL_SYM
    a.GetVersion(&a.version);
    g_api = a;
    return &g_api;
}
}  // namespace cude_nccl

#define NCCL_TRY(ctx, api, call)                                                                              \
    do {                                                                                                      \
        const int r__ = (call);                                                                               \
        if (r__ != cude_nccl::kSuccess) {                                                                     \
            char b__[512];                                                                                    \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #call, (api)->GetErrorString(r__), __FILE__, __LINE__); \
            return fail(ctx, CUDE_ENCCL, b__);                                                                \
        }                                                                                                     \
    } while (0)

static void comm_release(cude_ctx* ctx) {
    if (!ctx || !ctx->comm) return;
    const cude_nccl::Api* api = cude_nccl::load(nullptr);
    if (api) api->CommDestroy((cude_nccl::comm_t)ctx->comm);
    ctx->comm = nullptr; ctx->comm_nranks = 1; ctx->comm_rank = 0;
}

// in-place sum over the ranks of the context's communicator, asynchronous on the context's stream
static int comm_allreduce(cude_ctx* ctx, double* d_buf, size_t count) {
    std::string why;
    const cude_nccl::Api* api = cude_nccl::load(&why);
    if (!api) return fail(ctx, CUDE_ENCCL, why);
    NCCL_TRY(ctx, api, api->AllReduce(d_buf, d_buf, count, cude_nccl::kFloat64, cude_nccl::kSum, (cude_nccl::comm_t)ctx->comm, ctx->stream));
    return CUDE_OK;
}

// ---------------------------------------------------------------- one process per GPU
extern "C" int cude_nccl_version(void) {
    const cude_nccl::Api* api = cude_nccl::load(nullptr);
    return api ? api->version : 0;
}

extern "C" int cude_comm_get_unique_id(void* id_out) {
    if (!id_out) return fail(nullptr, CUDE_EINVAL, "cude_comm_get_unique_id: NULL");
    std::string why;
    const cude_nccl::Api* api = cude_nccl::load(&why);
    if (!api) return fail(nullptr, CUDE_ENCCL, why);
    cude_nccl::unique_id id;
    NCCL_TRY(nullptr, api, api->GetUniqueId(&id));
    memcpy(id_out, id.internal, CUDE_UNIQUE_ID_BYTES);
    return CUDE_OK;
}

extern "C" int cude_comm_init_rank(cude_ctx* ctx, int nranks, int rank, const void* id_in) {
    if (!ctx || !id_in || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, CUDE_EINVAL, "cude_comm_init_rank: bad argument");
    if (ctx->comm) return fail(ctx, CUDE_EINVAL, "cude_comm_init_rank: the context already has a communicator");
    std::string why;
    const cude_nccl::Api* api = cude_nccl::load(&why);
    if (!api) return fail(ctx, CUDE_ENCCL, why);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cude_nccl::unique_id id;
    memcpy(id.internal, id_in, CUDE_UNIQUE_ID_BYTES);
    cude_nccl::comm_t c = nullptr;
    NCCL_TRY(ctx, api, api->CommInitRank(&c, nranks, id, rank));
    ctx->comm = c; ctx->comm_nranks = nranks; ctx->comm_rank = rank;
    return CUDE_OK;
}

extern "C" int cude_comm_destroy(cude_ctx* ctx) {
    if (!ctx) return CUDE_EINVAL;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    comm_release(ctx);
    return CUDE_OK;
}

extern "C" int cude_comm_size(const cude_ctx* ctx) { return ctx ? (ctx->comm ? ctx->comm_nranks : 1) : CUDE_EINVAL; }
extern "C" int cude_comm_rank(const cude_ctx* ctx) { return ctx ? (ctx->comm ? ctx->comm_rank : 0) : CUDE_EINVAL; }

extern "C" int cude_allreduce_dev(cude_ctx* ctx, double* d_buf, long long count) {
    if (!ctx || !d_buf || count < 1) return fail(ctx, CUDE_EINVAL, "cude_allreduce_dev: bad argument");
    if (!ctx->comm || ctx->comm_nranks == 1) return CUDE_OK;   // a single rank: the sums are already global
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    return comm_allreduce(ctx, d_buf, (size_t)count);
}

extern "C" int cude_loss_sharded(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                                 int n_starts, const double* neural, long long neural_stride,
                                 const double* cond, long long ld, long long n_total, double* sse_out, double* loss_out) {
    if (n_total < 1) return fail(ctx, CUDE_EINVAL, "cude_loss_sharded: n_total < 1");
    HostShard sh; sh.ld = ld; sh.n_total = n_total;
    return eval_host(ctx, pop, net, opts, n_starts, neural, neural_stride, cond, 0, 1, sse_out, loss_out, nullptr, nullptr, nullptr, 1.0, sh);
}

extern "C" int cude_loss_grad_sharded(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                                      int n_starts, const double* neural, long long neural_stride,
                                      const double* cond, long long ld, long long n_total, int mean_over_individuals,
                                      double* sse_out, double* loss_out, double* g_neural, double* g_cond) {
    if (n_total < 1) return fail(ctx, CUDE_EINVAL, "cude_loss_grad_sharded: n_total < 1");
    HostShard sh; sh.ld = ld; sh.n_total = n_total;
    return eval_host(ctx, pop, net, opts, n_starts, neural, neural_stride, cond, 1, mean_over_individuals, sse_out, loss_out,
                     g_neural, g_cond, nullptr, 1.0, sh);
}

// ---------------------------------------------------------------- one process, all GPUs
namespace {
// one host thread per device: the single-GPU entry points are synchronous (they end with a stream synchronisation), so
// the devices are driven concurrently by handing each worker its slice of the call
struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, stop = false, done = true;
    int rc = 0;
    void loop() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return has_job || stop; });
            if (stop) return;
            std::function<int()> j = std::move(job);
            has_job = false;
            lk.unlock();
            const int r = j();
            lk.lock();
            rc = r; done = true;
            cv.notify_all();
        }
    }
    void submit(std::function<int()> j) {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(j); has_job = true; done = false;
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return done; });
        return rc;
    }
};
}  // namespace

struct cude_mctx {
    int n = 0;
    std::vector<cude_ctx*> ctx;
    std::vector<std::unique_ptr<Worker>> workers;
    bool have_comm = false;
    std::string err;
    cude_stats stats{};
};

struct cude_mpopulation {
    cude_mctx* mctx = nullptr;
    int mode = 0, n_ind = 0;
    std::vector<cude_population*> pop;   // per device (mode STARTS: replicas; mode INDIVIDUALS: shards)
    std::vector<int> lo;                 // mode INDIVIDUALS: device k owns individuals [lo[k], lo[k+1])
};

static int mfail(cude_mctx* m, int code, const std::string& msg) {
    if (m) m->err = msg;
    g_err = msg;
    return code;
}

// run job(k) on every device's worker; first failure wins (its message is kept)
static int mrun(cude_mctx* m, const std::function<int(int)>& job) {
    for (int k = 0; k < m->n; ++k) m->workers[k]->submit([&job, k] { return job(k); });
    int rc = CUDE_OK;
    for (int k = 0; k < m->n; ++k) {
        const int r = m->workers[k]->wait();
        if (r != CUDE_OK && rc == CUDE_OK) { rc = r; m->err = "device " + std::to_string(m->ctx[k]->device) + ": " + m->ctx[k]->err; g_err = m->err; }
    }
    return rc;
}

extern "C" int cude_mctx_destroy(cude_mctx* m) {
    if (!m) return CUDE_OK;
    for (auto& w : m->workers) {
        if (!w) continue;
        { std::lock_guard<std::mutex> lk(w->m); w->stop = true; w->cv.notify_all(); }
        if (w->th.joinable()) w->th.join();
    }
    for (cude_ctx* c : m->ctx) cude_ctx_destroy(c);   // releases the communicators too
    delete m;
    return CUDE_OK;
}

extern "C" int cude_mctx_create(int n_gpus, const int* device_ids, cude_mctx** out) {
    if (!out) return mfail(nullptr, CUDE_EINVAL, "cude_mctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return mfail(nullptr, CUDE_ENODEVICE, std::string("no CUDA device available (there is no CPU fallback): ") +
                                                  (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    if (n_gpus == 0) n_gpus = ndev;      // 0 = all visible devices
    if (n_gpus < 1 || n_gpus > ndev) return mfail(nullptr, CUDE_EINVAL, "cude_mctx_create: n_gpus exceeds the visible devices");
    for (int k = 0; device_ids && k < n_gpus; ++k)
        for (int j = 0; j < k; ++j)
            if (device_ids[j] == device_ids[k]) return mfail(nullptr, CUDE_EINVAL, "cude_mctx_create: duplicate device id");
    cude_mctx* m = new (std::nothrow) cude_mctx();
    if (!m) return mfail(nullptr, CUDE_ENOMEM, "out of host memory");
    m->n = n_gpus;
    for (int k = 0; k < n_gpus; ++k) {
        cude_ctx* c = nullptr;
        const int rc = cude_ctx_create(device_ids ? device_ids[k] : k, &c);
        if (rc) { const std::string msg = g_err; cude_mctx_destroy(m); return mfail(nullptr, rc, msg); }
        m->ctx.push_back(c);
    }
    for (int k = 0; k < n_gpus; ++k) {
        m->workers.emplace_back(new Worker());
        Worker* w = m->workers.back().get();
        w->th = std::thread([w] { w->loop(); });
    }
    *out = m;
    return CUDE_OK;
}

extern "C" int cude_mctx_size(const cude_mctx* m) { return m ? m->n : CUDE_EINVAL; }
extern "C" cude_ctx* cude_mctx_ctx(cude_mctx* m, int k) { return (m && k >= 0 && k < m->n) ? m->ctx[k] : nullptr; }
extern "C" const char* cude_mlast_error(const cude_mctx* m) { return m ? m->err.c_str() : g_err.c_str(); }

// communicator over the context's devices, created when the first individual-sharded population asks for it
static int mctx_ensure_comm(cude_mctx* m) {
    if (m->have_comm || m->n == 1) return CUDE_OK;
    std::string why;
    const cude_nccl::Api* api = cude_nccl::load(&why);
    if (!api) return mfail(m, CUDE_ENCCL, why);
    std::vector<int> devs(m->n);
    std::vector<cude_nccl::comm_t> comms(m->n, nullptr);
    for (int k = 0; k < m->n; ++k) devs[k] = m->ctx[k]->device;
    const int r = api->CommInitAll(comms.data(), m->n, devs.data());
    if (r != cude_nccl::kSuccess) return mfail(m, CUDE_ENCCL, std::string("ncclCommInitAll failed: ") + api->GetErrorString(r));
    for (int k = 0; k < m->n; ++k) { m->ctx[k]->comm = comms[k]; m->ctx[k]->comm_nranks = m->n; m->ctx[k]->comm_rank = k; }
    m->have_comm = true;
    return CUDE_OK;
}

extern "C" int cude_mpopulation_destroy(cude_mpopulation* mp) {
    if (!mp) return CUDE_OK;
    for (cude_population* p : mp->pop) cude_population_destroy(p);
    delete mp;
    return CUDE_OK;
}

extern "C" int cude_mpopulation_create(cude_mctx* m, int mode, int n_ind,
                                       int max_knots, const int* n_knots, const double* knot_t, const double* knot_g,
                                       int max_obs, const int* n_obs, const double* obs_t, const double* obs_y,
                                       const double* kin, const double* covariate, cude_mpopulation** out) {
    if (!m || !out) return mfail(m, CUDE_EINVAL, "cude_mpopulation_create: NULL mctx/out");
    *out = nullptr;
    if (mode != CUDE_SHARD_STARTS && mode != CUDE_SHARD_INDIVIDUALS) return mfail(m, CUDE_EINVAL, "cude_mpopulation_create: bad mode");
    if (n_ind < 1 || max_knots < 2 || max_obs < 1 || !n_knots || !knot_t || !knot_g || !n_obs || !obs_t || !obs_y || !kin)
        return mfail(m, CUDE_EINVAL, "cude_mpopulation_create: bad argument");
    if (mode == CUDE_SHARD_INDIVIDUALS && n_ind < m->n) return mfail(m, CUDE_EINVAL, "cude_mpopulation_create: fewer individuals than devices");
    int rc;
    if (mode == CUDE_SHARD_INDIVIDUALS && (rc = mctx_ensure_comm(m))) return rc;
    cude_mpopulation* mp = new (std::nothrow) cude_mpopulation();
    if (!mp) return mfail(m, CUDE_ENOMEM, "out of host memory");
    mp->mctx = m; mp->mode = mode; mp->n_ind = n_ind;
    mp->pop.assign(m->n, nullptr);
    mp->lo.assign(m->n + 1, 0);
    for (int k = 0; k <= m->n; ++k) mp->lo[k] = mode == CUDE_SHARD_INDIVIDUALS ? (int)((long long)n_ind * k / m->n) : (k == 0 ? 0 : n_ind);
    const size_t K = max_knots, M = max_obs;
    rc = mrun(m, [&](int k) {
        // STARTS: every device holds the whole population; INDIVIDUALS: device k holds the contiguous block [lo, hi)
        const size_t lo = mode == CUDE_SHARD_INDIVIDUALS ? (size_t)mp->lo[k] : 0;
        const int nloc = mode == CUDE_SHARD_INDIVIDUALS ? mp->lo[k + 1] - mp->lo[k] : n_ind;
        return cude_population_create(m->ctx[k], nloc, max_knots, n_knots + lo, knot_t + lo * K, knot_g + lo * K, max_obs, n_obs + lo,
                                      obs_t + lo * M, obs_y + lo * M, kin + lo * 4, covariate ? covariate + lo : nullptr, &mp->pop[k]);
    });
    if (rc) { cude_mpopulation_destroy(mp); return rc; }
    *out = mp;
    return CUDE_OK;
}

extern "C" int cude_mpopulation_size(const cude_mpopulation* mp) { return mp ? mp->n_ind : CUDE_EINVAL; }
extern "C" int cude_mpopulation_mode(const cude_mpopulation* mp) { return mp ? mp->mode : CUDE_EINVAL; }

// per-device statistics are read by the device's worker (the caller's current device is never changed)
static void mstats_sum(cude_mctx* m, const std::vector<cude_stats>& per) {
    cude_stats t{};
    for (int k = 0; k < m->n; ++k) {
        const cude_stats& s = per[k];
        t.n_traj += s.n_traj; t.n_acc += s.n_acc; t.n_rej += s.n_rej; t.n_rhs += s.n_rhs; t.n_fail += s.n_fail;
        t.launches += s.launches;
        if (s.kernel_ms > t.kernel_ms) t.kernel_ms = s.kernel_ms;      // devices run concurrently: the slowest one
    }
    m->stats = t;
}

static int meval(cude_mctx* m, const cude_mpopulation* mp, const cude_net* net, const cude_opts* opts,
                 int n_starts, const double* neural, long long neural_stride, const double* cond,
                 int want_grad, int mean_over_individuals,
                 double* sse_out, double* loss_out, double* g_neural, double* g_cond) {
    if (!m || !mp || mp->mctx != m) return mfail(m, CUDE_EINVAL, "cude_mloss: population belongs to another multi-GPU context");
    if (!net || !neural || !cond || n_starts < 1) return mfail(m, CUDE_EINVAL, "cude_mloss: bad argument");
    const int P = cude_net_nparams(net);
    if (P < 0) return mfail(m, CUDE_EINVAL, "cude_mloss: bad network description");
    const size_t N = (size_t)mp->n_ind;
    std::vector<cude_stats> per(m->n);
    for (auto& s : per) s = cude_stats{};
    const bool want_neural = want_grad && g_neural != nullptr;
    int rc;
    if (mp->mode == CUDE_SHARD_STARTS) {
        // starts [s0, s1) of device k: contiguous column blocks of every host matrix, no communication
        rc = mrun(m, [&](int k) {
            const long long s0 = (long long)n_starts * k / m->n, s1 = (long long)n_starts * (k + 1) / m->n;
            if (s1 == s0) return (int)CUDE_OK;
            const int r = eval_host(m->ctx[k], mp->pop[k], net, opts, (int)(s1 - s0), neural + s0 * neural_stride, neural_stride, cond + s0 * N,
                                    want_grad, mean_over_individuals, sse_out ? sse_out + s0 * N : nullptr, loss_out ? loss_out + s0 : nullptr,
                                    g_neural ? g_neural + s0 * P : nullptr, g_cond ? g_cond + s0 * N : nullptr);
            return r ? r : cude_get_stats(m->ctx[k], &per[k]);
        });
    } else {
        // rows [lo, hi) of device k; the per-start sums are all-reduced inside eval_host, rank 0 writes loss / g_neural
        rc = mrun(m, [&](int k) {
            HostShard sh; sh.ld = (long long)N; sh.n_total = (long long)N; sh.want_neural = want_neural ? 1 : 0;
            const size_t lo = (size_t)mp->lo[k];
            const int r = eval_host(m->ctx[k], mp->pop[k], net, opts, n_starts, neural, neural_stride, cond + lo, want_grad,
                                    mean_over_individuals, sse_out ? sse_out + lo : nullptr, k == 0 ? loss_out : nullptr,
                                    k == 0 ? g_neural : nullptr, g_cond ? g_cond + lo : nullptr, nullptr, 1.0, sh);
            return r ? r : cude_get_stats(m->ctx[k], &per[k]);
        });
    }
    if (rc) return rc;
    mstats_sum(m, per);
    return CUDE_OK;
}

extern "C" int cude_mloss(cude_mctx* m, const cude_mpopulation* mp, const cude_net* net, const cude_opts* opts,
                          int n_starts, const double* neural, long long neural_stride, const double* cond,
                          double* sse_out, double* loss_out) {
    return meval(m, mp, net, opts, n_starts, neural, neural_stride, cond, 0, 1, sse_out, loss_out, nullptr, nullptr);
}

extern "C" int cude_mloss_grad(cude_mctx* m, const cude_mpopulation* mp, const cude_net* net, const cude_opts* opts,
                               int n_starts, const double* neural, long long neural_stride, const double* cond,
                               int mean_over_individuals, double* sse_out, double* loss_out, double* g_neural, double* g_cond) {
    return meval(m, mp, net, opts, n_starts, neural, neural_stride, cond, 1, mean_over_individuals, sse_out, loss_out, g_neural, g_cond);
}

extern "C" int cude_mget_stats(cude_mctx* m, cude_stats* out) {
    if (!m || !out) return CUDE_EINVAL;
    *out = m->stats;
    return CUDE_OK;
}
