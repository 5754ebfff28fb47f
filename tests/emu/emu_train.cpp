// Host emulation of conditional_ude_b200/csrc/cude_train.cuh — TEST TOOL ONLY.
// The device-resident optimisers (Adam with best-iterate tracking, L-BFGS with BackTracking as a line-search state machine) run
// one block of 128 cooperating threads per start; here every CUDA thread is a host thread (emu_threads.h) and the host loop of
// cude_train (csrc/cude_api.cu) is restated around them with the objective supplied by the caller: objective(x[S x D]) fills
// f[S] and g[S x D].  Same kernel source, compiled with g++; never loaded by the package.
#define CUDE_HOST_EMU 1
#include "emu_threads.h"
#include <functional>
#undef __shared__
#define __shared__ static          // block-shared arrays inside the kernels: one copy for the block's host threads
#include "../../conditional_ude_b200/csrc/cude_train.cuh"

using namespace cude;

typedef void (*objective_t)(const double* x, double* f, double* g);

// the block's TRAIN_T host threads live for the whole emu_train call: a launch hands every block to them in turn
struct BlockPool {
    std::vector<std::thread> th;
    Barrier start, done;
    std::function<void()> work;
    bool quit = false;
    explicit BlockPool(int T) {
        for (int t = 0; t < T; ++t)
            th.emplace_back([this, t, T] {
                threadIdx.x = t;
                for (;;) {
                    start.wait(T + 1);
                    if (quit) return;
                    work();
                    done.wait(T + 1);
                }
            });
    }
    void run_block(int T) { start.wait(T + 1); done.wait(T + 1); }
    void stop(int T) { quit = true; start.wait(T + 1); for (auto& x : th) x.join(); }
};
static BlockPool* g_pool = nullptr;

template <class K>
static void launch(int n_blocks, K kernel) {
    g_block_threads = TRAIN_T;
    blockDim.x = TRAIN_T; gridDim.x = n_blocks;
    g_pool->work = kernel;
    for (int b = 0; b < n_blocks; ++b) {
        blockIdx.x = b;
        g_pool->run_block(TRAIN_T);
    }
}

// x[S x D] (D = P + N: network part first) holds the starting points on entry and the solutions on exit.
extern "C" int emu_train(int S, int P, int N, objective_t objective, int adam_iters, double adam_lr, int lbfgs_iters, int lbfgs_m,
                         double g_tol, double c1, double rho_hi, double rho_lo, int ls_maxiter, int check_every,
                         double* x, double* objective_out, int* iters_out, int* status_out, int* evals_out) {
    const size_t D = (size_t)P + N, m = (size_t)lbfgs_m;
    std::vector<double> xt_n((size_t)S * P), xt_c((size_t)S * N), sums((size_t)S * (P + 1)), gc((size_t)S * N);
    std::vector<double> X((size_t)S * D), G((size_t)S * D, 0.0), Dd((size_t)S * D, 0.0), best((size_t)S * D, 0.0), am((size_t)S * D, 0.0), av((size_t)S * D, 0.0);
    std::vector<double> hs(m * S * D, 0.0), hy(m * S * D, 0.0), rho(m * S, 0.0), sc((size_t)S * SC_N, 0.0);
    std::vector<int> ic((size_t)S * IC_N + 2, 0);
    for (int s = 0; s < S; ++s) {
        for (int k = 0; k < P; ++k) xt_n[(size_t)s * P + k] = x[s * D + k];
        for (int k = 0; k < N; ++k) xt_c[(size_t)s * N + k] = x[s * D + P + k];
        for (size_t k = 0; k < D; ++k) X[s * D + k] = x[s * D + k];
        sc[(size_t)s * SC_N + SC_BESTF] = INFINITY; sc[(size_t)s * SC_N + SC_ALPHA] = 1.0; sc[(size_t)s * SC_N + SC_APREV] = 1.0;
    }
    TrainArgs A{};
    A.S = S; A.P = P; A.N = N; A.D = (int)D; A.m = lbfgs_m;
    A.xt_n = xt_n.data(); A.xt_c = xt_c.data(); A.sums = sums.data(); A.g_cond = gc.data(); A.scale = 1.0 / (double)N;
    A.x = X.data(); A.g = G.data(); A.d = Dd.data(); A.best_x = best.data(); A.am = am.data(); A.av = av.data();
    A.hs = hs.data(); A.hy = hy.data(); A.rho = rho.data(); A.sc = sc.data(); A.ic = ic.data(); A.status_count = ic.data() + (size_t)S * IC_N;
    A.lr = adam_lr; A.b1 = 0.9; A.b2 = 0.999; A.eps = 1e-8;
    A.g_tol = g_tol; A.c1 = c1; A.rho_hi = rho_hi; A.rho_lo = rho_lo; A.ls_maxiter = ls_maxiter; A.maxiters = lbfgs_iters;
    int evals = 0;
    BlockPool pool(TRAIN_T);
    g_pool = &pool;
    std::vector<double> xe((size_t)S * D), fe(S), ge((size_t)S * D);
    auto eval = [&]() {
        // the library's evaluation: loss = scale * sums[0], d loss / d neural = scale * sums[1..], d loss / d cond = g_cond
        ++evals;
        for (int s = 0; s < S; ++s) {
            for (int k = 0; k < P; ++k) xe[s * D + k] = xt_n[(size_t)s * P + k];
            for (int k = 0; k < N; ++k) xe[s * D + P + k] = xt_c[(size_t)s * N + k];
        }
        objective(xe.data(), fe.data(), ge.data());
        for (int s = 0; s < S; ++s) {
            sums[(size_t)s * (P + 1)] = fe[s] / A.scale;
            for (int k = 0; k < P; ++k) sums[(size_t)s * (P + 1) + 1 + k] = ge[s * D + k] / A.scale;
            for (int k = 0; k < N; ++k) gc[(size_t)s * N + k] = ge[s * D + P + k];
        }
    };
    double b1t = 1.0, b2t = 1.0;
    for (int it = 0; it < adam_iters; ++it) {
        eval();
        b1t *= A.b1; b2t *= A.b2; A.b1t = b1t; A.b2t = b2t;
        launch(S, [&] { cude_adam_step_kernel(A, 0); });
    }
    if (adam_iters > 0) { eval(); launch(S, [&] { cude_adam_step_kernel(A, 1); }); }
    if (lbfgs_iters > 0) {
        eval();
        launch(S, [&] { cude_lbfgs_step_kernel(A, 1); });
        const long long max_micro = (long long)lbfgs_iters * (ls_maxiter + 1) + check_every;
        for (long long k = 0; k < max_micro; ++k) {
            eval();
            const bool check = ((k + 1) % check_every == 0);
            if (check) { A.status_count[0] = 0; A.status_count[1] = 0; }
            launch(S, [&] { cude_lbfgs_step_kernel(A, 0); });
            if (check && A.status_count[0] == 0) break;
        }
    }
    for (int s = 0; s < S; ++s) {
        for (size_t k = 0; k < D; ++k) x[s * D + k] = X[s * D + k];
        if (objective_out) objective_out[s] = lbfgs_iters > 0 ? sc[(size_t)s * SC_N + SC_FX] : sc[(size_t)s * SC_N + SC_BESTF];
        if (iters_out) iters_out[s] = ic[(size_t)s * IC_N + IC_ITERS];
        if (status_out) status_out[s] = ic[(size_t)s * IC_N + IC_STATUS];
    }
    if (evals_out) *evals_out = evals;
    pool.stop(TRAIN_T);
    g_pool = nullptr;
    return 0;
}
