"""CPU tier: the device-resident optimisers (csrc/cude_train.cuh — Adam with best-iterate tracking, L-BFGS with BackTracking
as a line-search state machine, one block of 128 threads per start) compiled for the host with every CUDA thread as a host
thread, driven by the library's host loop restated in tests/emu/emu_train.cpp, on analytic objectives against the host
optimisers estimation.adam_batched / lbfgs_batched (restatements of Optimisers.Adam and Optim.LBFGS + BackTracking)."""
import numpy as np

import emu_wrap
from conditional_ude_b200.estimation import adam_batched, lbfgs_batched

S, D, P = 3, 9, 4
RNG = np.random.default_rng(3)
A = RNG.standard_normal((S, D, D))
H = np.einsum("sij,skj->sik", A, A) + 0.5 * np.eye(D)           # one SPD Hessian per start
B = RNG.standard_normal((S, D))


def quad_fg(x):
    r = x - B
    hr = np.einsum("sij,sj->si", H, r)
    return 0.5 * np.einsum("si,si->s", r, hr), hr


def rosen_fg(x):
    a, b = x[:, :-1], x[:, 1:]
    f = np.sum(100.0 * (b - a ** 2) ** 2 + (1.0 - a) ** 2, axis=1)
    g = np.zeros_like(x)
    g[:, :-1] += -400.0 * a * (b - a ** 2) - 2.0 * (1.0 - a)
    g[:, 1:] += 200.0 * (b - a ** 2)
    return f, g


def test_adam_phase_follows_the_host_optimiser_and_keeps_the_best_iterate():
    x0 = RNG.standard_normal((S, D))
    xa, fa = adam_batched(quad_fg, x0, lr=5e-2, maxiters=60)
    xe, fe, _, _, ev = emu_wrap.emu_train(quad_fg, x0, P, adam_iters=60, adam_lr=5e-2)
    assert ev == 61 and np.allclose(xe, xa, rtol=0, atol=1e-12) and np.allclose(fe, fa, rtol=1e-12)
    # an objective that gets worse after some steps (large learning rate): the best iterate is what comes back
    xa, fa = adam_batched(quad_fg, x0, lr=2.0, maxiters=15)
    xe, fe, _, _, _ = emu_wrap.emu_train(quad_fg, x0, P, adam_iters=15, adam_lr=2.0)
    assert np.allclose(xe, xa, atol=1e-12) and np.allclose(fe, fa, rtol=1e-12)
    assert np.all(fe <= quad_fg(x0)[0])


def test_lbfgs_state_machine_walks_the_host_optimisers_path():
    x0 = RNG.standard_normal((S, D))
    for fg, iters in ((quad_fg, 20), (rosen_fg, 12)):
        f = lambda x, fg=fg: fg(x)[0]
        xh, fh, ith, conv = lbfgs_batched(f, fg, x0, maxiters=iters)
        xe, fe, ite, st, ev = emu_wrap.emu_train(fg, x0, P, lbfgs_iters=iters)
        assert np.array_equal(ite, ith), (ite, ith)                       # same accepted iterations
        assert np.allclose(fe, fh, rtol=1e-9, atol=1e-18) and np.allclose(xe, xh, rtol=1e-7, atol=1e-9)
        assert set(st.tolist()) <= {1, 3}
    # the quadratic is solved to the gradient tolerance: status 1, the minimiser B
    xe, fe, ite, st, ev = emu_wrap.emu_train(quad_fg, x0, P, lbfgs_iters=100)
    assert np.all(st == 1) and np.abs(xe - B).max() < 1e-7 and np.abs(quad_fg(xe)[1]).max() <= 1e-8


def test_failed_evaluations_are_survived():
    # Inf objective beyond a wall: Adam skips the step (moments decay), the line search halves its step
    def walled(x):
        f, g = quad_fg(x)
        bad = np.any(x > 3.0, axis=1)
        f = np.where(bad, np.inf, f)
        g = np.where(bad[:, None], np.nan, g)
        return f, g
    x0 = np.full((S, D), 2.5)
    xe, fe, ite, st, ev = emu_wrap.emu_train(walled, x0, P, adam_iters=10, adam_lr=1e-1, lbfgs_iters=10, ls_maxiter=12)
    assert np.all(np.isfinite(fe)) and np.all(fe <= quad_fg(x0)[0]) and np.all(xe <= 3.0)
