/* =====================================================================================
 * cude_b200.h — C ABI of the B200-native c-peptide conditional-UDE loss / gradient path.
 *
 * The reference (Computational-Biology-TUe/conditional-ude) has no FFI: its seam is the Julia
 * method table.  Each entry point below replaces the reference interface cited beside it; the
 * Julia `ccall` stubs a maintainer would add are in INTEGRATION.md and julia/CUDEB200.jl.
 *
 * Conventions
 *   - every function returns 0 on success, a negative CUDE_E* code on error; the message is
 *     available from cude_last_error().  No exceptions cross the boundary.
 *   - the caller owns all host buffers; the library owns device memory behind opaque handles.
 *   - matrices are column-major (Julia): cond[i + n_ind*s], neural[p + neural_stride*s].
 *   - FP64 host interface.  A context is bound to one CUDA device and one host thread at a time.
 *   - there is NO CPU fallback: without a CUDA device cude_ctx_create fails with CUDE_ENODEVICE.
 *   - a trajectory whose integrator fails (maxiters, dt underflow, NaN) has sse = +Inf and zero
 *     gradient; a start containing such a trajectory has loss = +Inf
 *     (src/parameter-estimation.jl:61-64, :134-136).
 * ===================================================================================== */
#ifndef CUDE_B200_H
#define CUDE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CUDE_B200_ABI_VERSION 3

enum {
    CUDE_OK = 0,
    CUDE_EINVAL = -1,     /* bad argument */
    CUDE_ENODEVICE = -2,  /* no usable CUDA device (no CPU fallback) */
    CUDE_ECUDA = -3,      /* CUDA runtime error, see cude_last_error */
    CUDE_ENOMEM = -4,
    CUDE_EUNSUPPORTED = -5, /* network shape / option not compiled in */
    CUDE_ENCCL = -6         /* NCCL missing or a collective failed (multi-GPU entry points only) */
};

typedef struct cude_ctx cude_ctx;
typedef struct cude_population cude_population;

/* Network of src/neural-network.jl:42-58 `chain(width, depth, tanh; input_dims)`:
 * n_in inputs -> `depth` hidden tanh layers of `width` -> 1 softplus output.
 * Parameter layout = SimpleChains TurboDense{true}: per layer W[out x in] column-major, then
 * bias[out].  Input order [dG; beta] (c-peptide-models.jl:91) or [dG; beta; covariate] (:101). */
typedef struct {
    int n_in;   /* 2 (cUDE) or 3 (covariate cUDE) */
    int depth;  /* hidden layers */
    int width;  /* hidden width */
} cude_net;

/* Solver options of `solve(model.problem, p=theta, saveat=timepoints, save_idxs=1)`
 * (src/parameter-estimation.jl:59): OrdinaryDiffEq defaults -> Tsit5, abstol 1e-6, reltol 1e-3,
 * maxiters 1000000 (OrdinaryDiffEq's __init default for adaptive algorithms).  cude_default_opts() fills these. */
typedef struct {
    double abstol;
    double reltol;
    int maxiters;
    int precision; /* 0 = FP64 (the parity-gated mode); 1 = FP32 network evaluation (MUFU ex2/rcp/lg2) with the
                    * integrator, adjoint and reductions in FP64: looser documented bound, see DESIGN.md;
                    * 2 = forward pass (loss, step sequence) in FP64 bit for bit, FP32 network and accumulators only in
                    * the adjoint sweep: loss as in mode 0, gradients to ~1e-6 relative */
    int block;     /* threads per block (individuals per tile); 0 = library default */
    int balance;   /* lane balance of loss + full-gradient calls with per-start networks (a warp runs to its slowest lane and the
                    * trajectories of a start take 17-26 steps, so 12 % of the lane-cycles idle in natural order):
                    * 0 (default) = automatic:
                    *     - calls of <= 8192 trajectories use the WARP-PER-TRAJECTORY latency kernel (mode 4: a whole warp per trajectory
                    *       up to 512 trajectories, 8 lanes per trajectory above; loss-only calls of that size its forward half);
                    *     - populations of >= 32768 individuals use the TWO-KERNEL gradient (mode 2);
                    *     - everything else the fused kernel (mode 3);
                    * 1 = fused kernel, each start's individuals regrouped by the step counts of an EARLIER call on this
                    *     population (refreshed every 8 calls; pays only when the parameters barely move between calls);
                    * 2 = the two-kernel gradient: forward solve leaving a 64-byte record per accepted step, every start's
                    *     trajectories sorted by their accepted-step count (stable radix sort: deterministic), adjoint sweep in
                    *     the sorted order with no idle lanes (+14 % on B200; needs ~2.2 KB of device memory per trajectory of a
                    *     group of starts, the library sizes the groups to min(64 GB, 40 % of the free memory));
                    * 3 = the fused kernel (one thread per trajectory, forward solve + adjoint sweep) in natural order;
                    * 4 = one warp per trajectory (small batches: config 1, the selected starts of `train`): the 5 network nodes
                    *     of a step on 5 lanes, the adjoint's (step, node) evaluations spread over the trajectory's 32 (or 8) lanes — 0.06-0.08 ms
                    *     instead of 0.2 ms per evaluation of 57-1425 trajectories.
                    * Per-trajectory sse is bitwise the same in every mode, d/d cond in modes 0-3 (mode 4: to 1e-13); the
                    * per-start sums differ only in summation order and are run-to-run deterministic in modes 0, 2, 3 and 4. */
    int split;     /* gradient pipeline of loss + full-gradient calls with per-start networks:
                    * 0, 1 = the fused single-kernel adjoint (default);
                    * 2 = the split pipeline: forward solve leaving one record per accepted step -> adjoint recursion
                    *     per trajectory -> one thread per step record for the network's forward + backward evaluations ->
                    *     per-trajectory finish (needs ~2.3 KB of device memory per trajectory of a group of starts; the
                    *     library sizes the groups).  Same discrete adjoint: per-trajectory sse bit-identical, gradients
                    *     equal to summation order (1e-15).  Measured 5-15 % slower than the fused kernel on B200
                    *     (profiles/README.md): kept as a tested alternative, not the default. */
} cude_opts;

typedef struct {
    unsigned long long n_traj;  /* trajectories evaluated by the last call            */
    unsigned long long n_acc;   /* accepted Tsit5 steps, summed over trajectories     */
    unsigned long long n_rej;   /* rejected steps                                     */
    unsigned long long n_rhs;   /* RHS evaluations the reference scheme would perform */
    unsigned long long n_fail;  /* trajectories that returned +Inf                    */
    float kernel_ms;            /* device time of the last call's kernels (CUDA events) */
    int launches;               /* kernels launched by the last call                  */
} cude_stats;

int cude_abi_version(void);
void cude_default_opts(cude_opts* o);
/* number of MLP parameters: 37 for (2,2,4), 41 for (3,2,4) */
int cude_net_nparams(const cude_net* net);

/* van_cauter_parameters(age, t2dm) -> (k0,k1,k2), src/c-peptide-models.jl:30-42 (host arithmetic) */
void cude_van_cauter_parameters(double age, int t2dm, double* k0, double* k1, double* k2);

/* ---- context ---- */
int cude_device_count(void);   /* visible CUDA devices; 0 when there is none (or no driver) */
int cude_ctx_create(int device, cude_ctx** out);
int cude_ctx_destroy(cude_ctx* ctx);
/* message of the last error on this context (ctx may be NULL for creation errors) */
const char* cude_last_error(const cude_ctx* ctx);
int cude_sync(cude_ctx* ctx);
int cude_get_stats(cude_ctx* ctx, cude_stats* out);
/* the CUDA stream (cudaStream_t) the context launches on, for interop / event timing */
void* cude_ctx_stream(cude_ctx* ctx);
/* make the context launch on a caller-owned stream (e.g. the one NCCL all-reduces on);
 * NULL restores the context's own stream */
int cude_ctx_set_stream(cude_ctx* ctx, void* cuda_stream);

/* ---- population: the device-resident image of a vector of CPeptideConditionalUDEModel
 * (ctor src/c-peptide-models.jl:170-194; covariate variant :196-220) plus its data
 * (timepoints, cpeptide_data rows of the loss tuple, parameter-estimation.jl:126).
 *   knot_t, knot_g : [n_ind x max_knots] row-major (one row per individual), n_knots[i] used
 *   obs_t, obs_y   : [n_ind x max_obs]   row-major, n_obs[i] used; obs_t must lie in
 *                    [knot_t[0], knot_t[n_knots-1]] and be increasing
 *   kin            : [n_ind x 4] rows k0,k1,k2,c0   (u0 = [c0, k2/k1*c0], tspan = knot range)
 *   covariate      : [n_ind] third network input (age) or NULL
 * Uploaded once; re-used by every loss / gradient / profile call. */
int cude_population_create(cude_ctx* ctx, int n_ind,
                           int max_knots, const int* n_knots, const double* knot_t, const double* knot_g,
                           int max_obs, const int* n_obs, const double* obs_t, const double* obs_y,
                           const double* kin, const double* covariate, cude_population** out);
int cude_population_destroy(cude_population* pop);
int cude_population_size(const cude_population* pop);

/* ---- loss only: replaces loss(theta, (model,t,y)) :56-68, loss(theta, (model,t,y,nn)) :93-99,
 * the population loss :126-140, the screening loop :362-366 and likelihood_profile
 * (src/likelihood-profiles.jl:4-17).
 *   start s uses network  neural + s*neural_stride   (neural_stride 0 = one shared network, i.e.
 *   the fixed-NN / profile case) and conditional parameters cond[i + n_ind*s].
 *   sse_out [n_ind x n_starts] (may be NULL): per-trajectory sum of squared errors
 *   loss_out[n_starts]         (may be NULL): mean over individuals (:139), Inf if any failed */
int cude_loss(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
              int n_starts, const double* neural, long long neural_stride, const double* cond,
              double* sse_out, double* loss_out);

/* ---- model prediction: the `solve(model.problem, p=theta, saveat=timepoints, save_idxs=1)` inside the loss (:59) by
 * itself — what the scripts plot against the data (e.g. c-peptide/02-conditional.jl:170) and what generates synthetic
 * observations.  yhat_out[k + max_obs*(i + n_ind*s)] = plasma c-peptide of individual i under start s at its k-th
 * observation time; NaN for k >= n_obs[i] and for failed solves.  sse_out [n_ind x n_starts] may be NULL. */
int cude_simulate(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                  int n_starts, const double* neural, long long neural_stride, const double* cond,
                  double* yhat_out, double* sse_out);

/* ---- loss + gradient: replaces OptimizationFunction(loss, AutoForwardDiff()) :231,:281,:299,:370.
 *   g_neural[P x n_starts] (may be NULL: beta-only estimation, :272-288): d loss[s] / d neural
 *   g_cond  [n_ind x n_starts] (may be NULL): d loss[s] / d cond[i,s]   (= (1/N) d sse_i / d cond)
 *   With mean_over_individuals = 0 the outputs are per-trajectory sums instead:
 *   loss_out[s] = sum_i sse, g_neural = sum_i d sse_i, g_cond = d sse_i / d cond (beta-only fits). */
int cude_loss_grad(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                   int n_starts, const double* neural, long long neural_stride, const double* cond,
                   int mean_over_individuals,
                   double* sse_out, double* loss_out, double* g_neural, double* g_cond);

/* ---- sharded-population form of cude_loss_grad with host buffers (one rank's shard of the individuals,
 * parameter-estimation.jl:126-140 split over ranks): sums_out[(P+1) x n_starts] receives for every start the *unscaled*
 * { sum_i sse_i, sum_i d sse_i / d neural[0..P) } of this shard (sum the shards' sums, divide by the global N; a start
 * whose sums_out[0] is not finite failed), g_cond[n_ind x n_starts] receives d sse_i / d cond scaled by cond_scale
 * (pass 1/N_global).  Like cude_loss / cude_loss_grad, a call of more than ~2 M trajectories runs as a pipeline of
 * chunks of starts: the host->device copy of chunk k+1 and the device->host copy of chunk k-1 overlap the kernels
 * of chunk k (fully asynchronous when the host buffers are page-locked). */
int cude_loss_grad_sums(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                        int n_starts, const double* neural, long long neural_stride, const double* cond,
                        double cond_scale, double* sums_out, double* g_cond);

/* ---- device-resident variants (asynchronous on the context stream; all pointers are device
 * pointers on the context's device).  Used by multi-GPU population training: each rank holds a
 * shard of the individuals, `sums_out` [(P+1) x n_starts] receives for every start
 * { sum_i sse_i, sum_i d sse_i / d neural[0..P) } so that it can be all-reduced in place (NCCL)
 * and divided by the global N; g_cond receives d sse_i / d cond scaled by cond_scale
 * (pass 1/N_global).  want_grad: 0 = only the sse sums (rows 1..P are zeroed); bit 0 = d/d cond; bit 1 = d/d neural
 * (3 = both: the adjoint kernel; 1 = beta-only: the cheaper forward-sensitivity kernel). */
int cude_eval_dev(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                  int n_starts, const double* d_neural, long long neural_stride, const double* d_cond,
                  int want_grad, double cond_scale,
                  double* d_sse_out, double* d_sums_out, double* d_g_cond);

/* ---- second variant: the suppression example (suppression/src/suppression_model.jl).
 * State-dependent cUDE: du1 = -p1 u1; du2 = p1 u1 - NN([u; exp(theta_i)]); du3 = NN(.) - p3 u3  (ude_lsup! :88-95),
 * network input_dims 4 -> `depth` tanh layers of `width` -> 1 softplus (neural_network_model :78-86; suppression.jl:18
 * builds depth 5, width 3 = 67 parameters), explicit Tsit5 at default tolerances, all three states observed.
 *   data   : [3 x n_obs x n_ind] column-major as in Julia (state fastest); u0 of individual i = data[:,1,i] (:99-104)
 *   obs_t  : [n_obs] common time grid (saveat); tspan = (t0, tend)
 *   p_true : {p1, p2, p3} (p2 unused by the hybrid model); scale[3] or NULL = mean_i max_t data (:125)            */
typedef struct cude_sup_population cude_sup_population;
int cude_sup_population_create(cude_ctx* ctx, int n_ind, int n_obs, const double* obs_t, const double* data,
                               const double* p_true, const double* scale, double t0, double tend,
                               cude_sup_population** out);
int cude_sup_population_destroy(cude_sup_population* pop);
/* suppression_loss(p, (prob, data, timepoints, lambda)) :117-130 for n_starts parameter sets and its gradient:
 *   loss_out[s] = sum_i sse_i / N + lambda * sum(neural_s .^ 2)   (Inf if a trajectory failed)
 *   g_neural[P x n_starts] (may be NULL), g_theta[n_ind x n_starts] (may be NULL => loss only)
 *   sse_out [n_ind x n_starts] (may be NULL): per-individual scaled SSE
 * Gradient calls run as two kernels (forward solve leaving step records, then the adjoint sweep over them; opts.split = 1: one
 * fused kernel, same results bit for bit); a call with a solve of more than 64 accepted steps is redone by the fused kernel,
 * which replays the forward pass in chunks: any tolerance works. */
int cude_sup_loss_grad(cude_ctx* ctx, const cude_sup_population* pop, int depth, int width, const cude_opts* opts,
                       int n_starts, const double* neural, long long neural_stride, const double* theta, double lambda,
                       double* sse_out, double* loss_out, double* g_neural, double* g_theta);

/* ---- device-resident multi-start training: `_optimize` (src/parameter-estimation.jl:170-183) for all selected starts of
 * `train` (:374-376) in lock-step — Optimisers.Adam(adam_lr) for adam_iters iterations (best iterate kept), then
 * Optim.LBFGS(m, linesearch = BackTracking(order 3: c1, rho_lo, rho_hi)) for at most lbfgs_iters iterations, objective =
 * the population loss (:126-140).  Parameters, gradients, moments and the L-BFGS history never leave the device; per
 * iteration the host only enqueues the loss+gradient kernel, its reduction and one optimiser kernel (one block per
 * start), and reads 8 bytes every `check_every` line-search steps to stop when every start has finished.
 *   neural[P x n_starts], cond[n_ind x n_starts]: initial parameters on entry, solutions on exit (host memory)
 *   objective_out[s]: final loss; iters_out[s]: accepted L-BFGS iterations; status_out[s]: 0 = stopped by the step budget,
 *   1 = converged (|g|_inf <= g_tol, or no change), 2 = line search failed, 3 = lbfgs_iters reached; evals_out: loss+gradient
 *   evaluations of the whole batch. */
typedef struct {
    int adam_iters; double adam_lr, adam_beta1, adam_beta2, adam_eps;
    int lbfgs_iters, lbfgs_m; double g_tol, c1, rho_hi, rho_lo; int ls_maxiter;
    int check_every;
} cude_train_opts;
void cude_train_default_opts(cude_train_opts* t);
int cude_train(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
               const cude_train_opts* topts, int n_starts, double* neural, double* cond,
               double* objective_out, int* iters_out, int* status_out, int* evals_out);

/* Device-resident Adam step (Optimisers.Adam(eta, (beta1, beta2), eps), `_optimize` step 1,
 * src/parameter-estimation.jl:170-183) for population-scale training, where the parameter vector (one beta per
 * individual per start) never leaves HBM:  g <- grad_scale * d_g;  m, v moments;  x -= lr * mhat / (sqrt(vhat) + eps)
 * with bias correction for iteration t >= 1.  Asynchronous on the context stream.  If d_row_flag is given, element i
 * belongs to row i / row_len and is skipped when d_row_flag[row * flag_stride] is not finite (a start whose loss is
 * Inf keeps its parameters). */
int cude_adam_dev(cude_ctx* ctx, long long n, double* d_x, const double* d_g, double* d_m, double* d_v,
                  double lr, double beta1, double beta2, double eps, int t, double grad_scale,
                  const double* d_row_flag, long long row_len, long long flag_stride);

/* =====================================================================================
 * Multi-GPU (one B200 box, NCCL over NVLink / NVSwitch).  The reference runs every loop of the path on one CPU thread;
 * here the independent trajectories are partitioned over the GPUs:
 *   - starts / initial guesses / profile grid points (src/parameter-estimation.jl:362-366, :374-376, :272-288,
 *     src/likelihood-profiles.jl:11-14) shard with NO communication: every device holds the whole population;
 *   - the individuals of a large population (the loop of :126-140) shard with ONE exchange per call: the all-reduce
 *     (sum, FP64) of the per-start rows {sum_i sse_i, sum_i d sse_i/d neural[0..P)} — (P+1) x n_starts doubles —
 *     issued inside the library on the stream and buffer the reduction kernel wrote.  d/d cond needs no exchange.
 * NCCL is loaded at run time (libnccl.so.2; override with the environment variable CUDE_NCCL_LIB); without it the
 * entry points below that need a communicator return CUDE_ENCCL, everything else keeps working.
 * ===================================================================================== */

/* ---- (1) one process per GPU: torchrun, MPI, Julia Distributed workers ----
 * Rank 0 calls cude_comm_get_unique_id and the host broadcasts the CUDE_UNIQUE_ID_BYTES bytes by its own means; then
 * every rank calls cude_comm_init_rank on its context (collective, like ncclCommInitRank). */
#define CUDE_UNIQUE_ID_BYTES 128
int cude_nccl_version(void);                        /* e.g. 22703; 0 when NCCL cannot be loaded */
int cude_comm_get_unique_id(void* id_out);
int cude_comm_init_rank(cude_ctx* ctx, int nranks, int rank, const void* id);
int cude_comm_destroy(cude_ctx* ctx);               /* also done by cude_ctx_destroy */
int cude_comm_size(const cude_ctx* ctx);            /* 1 without a communicator */
int cude_comm_rank(const cude_ctx* ctx);
/* in-place sum of a device buffer over the ranks, asynchronous on the context's stream (what follows cude_eval_dev in
 * device-resident population training); a no-op without a communicator */
int cude_allreduce_dev(cude_ctx* ctx, double* d_buf, long long count);
/* Population loss (:126-140) and its gradient when `pop` holds this rank's contiguous block of the n_total individuals.
 * Collective: every rank calls it with the same n_starts and networks.  cond / sse_out / g_cond are this rank's rows of
 * the global [n_total x n_starts] matrices: pass the address of the rank's first row and ld = the column stride in
 * elements (n_total when the host holds the global matrix, 0 = this rank's n_ind when it holds only its block).
 * loss_out[s] and g_neural[P x n_starts] are the global values on every rank (mean over n_total when
 * mean_over_individuals != 0, sums otherwise); g_cond = d loss[s] / d cond[i,s] for the rank's individuals. */
int cude_loss_sharded(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                      int n_starts, const double* neural, long long neural_stride,
                      const double* cond, long long ld, long long n_total, double* sse_out, double* loss_out);
int cude_loss_grad_sharded(cude_ctx* ctx, const cude_population* pop, const cude_net* net, const cude_opts* opts,
                           int n_starts, const double* neural, long long neural_stride,
                           const double* cond, long long ld, long long n_total, int mean_over_individuals,
                           double* sse_out, double* loss_out, double* g_neural, double* g_cond);

/* ---- (2) one process, all GPUs: what a single Julia session uses ----
 * cude_mctx_create(n_gpus, device_ids) owns one context + stream + host worker thread per device
 * (device_ids NULL = 0..n_gpus-1; n_gpus 0 = all visible devices).  cude_mpopulation_create takes the arguments of
 * cude_population_create plus the partition:
 *   CUDE_SHARD_STARTS       every device holds the whole population; a call's starts are split into contiguous blocks
 *                           (screening, selected starts, beta-only fits, profiles — no communication)
 *   CUDE_SHARD_INDIVIDUALS  device k holds individuals [N k / n, N (k+1) / n); every call ends with the all-reduce
 *                           of the per-start rows (population training; ncclCommInitAll on first use)
 * cude_mloss / cude_mloss_grad take the same GLOBAL host matrices as cude_loss / cude_loss_grad and fill the same
 * outputs; results do not depend on the number of devices beyond the summation order of the per-start sums
 * (per-trajectory sse and d/d cond are bit-identical).  One host thread per cude_mctx at a time. */
typedef struct cude_mctx cude_mctx;
typedef struct cude_mpopulation cude_mpopulation;
enum { CUDE_SHARD_STARTS = 0, CUDE_SHARD_INDIVIDUALS = 1 };
int cude_mctx_create(int n_gpus, const int* device_ids, cude_mctx** out);
int cude_mctx_destroy(cude_mctx* mctx);
int cude_mctx_size(const cude_mctx* mctx);
cude_ctx* cude_mctx_ctx(cude_mctx* mctx, int k);    /* the k-th device's context (stats, stream, device-pointer calls) */
const char* cude_mlast_error(const cude_mctx* mctx);
int cude_mpopulation_create(cude_mctx* mctx, int mode, int n_ind,
                            int max_knots, const int* n_knots, const double* knot_t, const double* knot_g,
                            int max_obs, const int* n_obs, const double* obs_t, const double* obs_y,
                            const double* kin, const double* covariate, cude_mpopulation** out);
int cude_mpopulation_destroy(cude_mpopulation* mpop);
int cude_mpopulation_size(const cude_mpopulation* mpop);
int cude_mpopulation_mode(const cude_mpopulation* mpop);
int cude_mloss(cude_mctx* mctx, const cude_mpopulation* mpop, const cude_net* net, const cude_opts* opts,
               int n_starts, const double* neural, long long neural_stride, const double* cond,
               double* sse_out, double* loss_out);
int cude_mloss_grad(cude_mctx* mctx, const cude_mpopulation* mpop, const cude_net* net, const cude_opts* opts,
                    int n_starts, const double* neural, long long neural_stride, const double* cond,
                    int mean_over_individuals, double* sse_out, double* loss_out, double* g_neural, double* g_cond);
/* statistics of the last cude_mloss / cude_mloss_grad: counts summed over the devices, kernel_ms of the slowest */
int cude_mget_stats(cude_mctx* mctx, cude_stats* out);

/* Test hook: evaluates the kernels' own branch-free FP64 elementary functions on the device
 * (which: 0 tanh, 1 softplus, 2 sigmoid, 3 exp clamped to +-40, 4 log of a positive normal, 5 reciprocal). */
int cude_math_probe(cude_ctx* ctx, int which, int n, const double* x, double* y);

/* Measured FP64 FMA peak of the context's device (dependent-chain-free DFMA micro-benchmark),
 * the denominator of the roofline.  Returns TFLOP/s in *tflops. */
int cude_measure_fp64_peak(cude_ctx* ctx, double* tflops);
/* Same for the optional FP32-network modes (opts.precision = 1, 2): FP32 FMA peak in TFLOP/s and, if mufu_gops is not
 * NULL, the MUFU (ex2.approx) rate in 1e9 operations / s. */
int cude_measure_fp32_peak(cude_ctx* ctx, double* tflops, double* mufu_gops);
/* Diagnostic: the same with three distinct register operands per DFMA (register-file-bandwidth bound shape). */
int cude_measure_fp64_peak_rrr(cude_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* CUDE_B200_H */
