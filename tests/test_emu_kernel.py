"""CPU tier: the CUDA kernel *source* (conditional_ude_b200/csrc/cude_kernels.cuh) compiled for the host
behind a CUDA shim (tests/emu) and run one thread per block, against the oracle.

This is a test tool, not a product path: it lets the CPU-only test tier exercise the kernel's
integrator, dense output, discrete adjoint, step-ring replay and gradient expansion.  The GPU tier
(test_gpu_parity.py) runs the same comparisons on the real device through the C ABI.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import conditional_ude_b200 as cu
from oracle import oracle
from helpers import train57, mixed_population, ohashi_models, random_starts, noise_ok, check_math
import emu_wrap

DET = dict(abstol=1e3, reltol=1e3)


def relmax(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_deterministic_regime_exact(fx):
    models, ts, ys = mixed_population(fx)
    pk = cu.pack_models(models, ts, ys)
    rng = np.random.default_rng(0)
    neural, cond = random_starts(rng, pk["chain"], len(models), 3)
    g = oracle.OraclePopulation(pk).eval(neural, cond, grad_mode=0, **DET)
    e = emu_wrap.emu_eval(pk, neural, cond, **DET)
    assert relmax(e["sse"], g["sse"]) < 1e-10
    assert relmax(e["g_cond"], g["g_cond"]) < 1e-9
    assert relmax(e["g_neural"], g["g_neural"]) < 1e-9
    assert e["n_acc"] == g["stats"][..., 0].sum() and e["n_rej"] == g["stats"][..., 1].sum() and e["n_fail"] == 0
    # flat indexing (shared network) gives the same per-trajectory results
    e1 = emu_wrap.emu_eval(pk, neural[0], cond, flat=True, **DET)
    e2 = emu_wrap.emu_eval(pk, neural[0], cond, flat=False, **DET)
    assert np.array_equal(e1["sse"], e2["sse"]) and np.array_equal(e1["g_cond"], e2["g_cond"])
    # loss-only instantiation == forward pass of the gradient instantiation
    e3 = emu_wrap.emu_eval(pk, neural, cond, grad=False, **DET)
    assert np.array_equal(e3["sse"], e["sse"])


def test_default_tolerance_within_noise_floor(fx):
    models, t, c, nn, betas = train57(fx)
    pk = cu.pack_models(models, t, c)
    g = oracle.OraclePopulation(pk).eval(nn, betas, grad_mode=0)
    e = emu_wrap.emu_eval(pk, nn, betas)
    assert noise_ok(np.abs(e["sse"] - g["sse"]) / g["sse"], 1e-5)
    # gradients relative to the scale of the individual terms (at the stored optimum the net d/dbeta is ~0)
    assert noise_ok(np.abs(e["g_cond"] - g["g_cond"]) / np.abs(g["g_cond"]).max(), 1e-4)
    assert noise_ok(np.abs(e["g_neural"] - g["g_neural"]) / np.abs(g["g_neural"]).max(axis=-1, keepdims=True), 1e-4)
    assert abs(e["sse"].mean() - 0.4281389) < 1e-6


def test_covariate_network(fx):
    models, t, c = ohashi_models(fx, "train", covariate=True)
    idx = fx["train_split_idx"][:20]
    pk = cu.pack_models([models[i] for i in idx], t, c[idx])
    nn, betas = fx["cov_neural"][1], fx["cov_betas"][1][:20]
    g = oracle.OraclePopulation(pk).eval(nn, betas, grad_mode=0, **DET)
    e = emu_wrap.emu_eval(pk, nn, betas, **DET)
    assert e["g_neural"].shape[-1] == 41
    assert relmax(e["sse"], g["sse"]) < 1e-10 and relmax(e["g_neural"], g["g_neural"]) < 1e-9
    assert relmax(e["g_cond"], g["g_cond"]) < 1e-9


def test_step_ring_replay(fx, tmp_path):
    """Rebuild the emulation with a 4-entry step ring: the 9-step deterministic solve then needs three
    replays of the forward pass; the gradient must not change."""
    lib = str(tmp_path / "libcude_emu_cap4.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-DCUDE_REC_CAP=4", "-o", lib, os.path.join(emu_wrap._HERE, "emu_kernel.cpp")])
    models, ts, ys = mixed_population(fx)
    pick = [0, 50, 100, 120, 130]
    pk = cu.pack_models([models[i] for i in pick], [ts[i] for i in pick], [ys[i] for i in pick])
    rng = np.random.default_rng(4)
    neural, cond = random_starts(rng, pk["chain"], len(pick), 2)
    g = oracle.OraclePopulation(pk).eval(neural, cond, grad_mode=0, **DET)
    assert g["stats"][..., 0].min() > 4
    old = emu_wrap.LIB
    try:
        emu_wrap.LIB = lib
        emu_wrap.build = lambda: lib
        assert C.CDLL(lib).emu_rec_cap() == 4
        e = emu_wrap.emu_eval(pk, neural, cond, **DET)
    finally:
        emu_wrap.LIB = old
        import importlib
        importlib.reload(emu_wrap)
    assert relmax(e["sse"], g["sse"]) < 1e-10
    assert relmax(e["g_cond"], g["g_cond"]) < 1e-9 and relmax(e["g_neural"], g["g_neural"]) < 1e-9


def test_failures(fx):
    models, t, c, nn, betas = train57(fx)
    pk = cu.pack_models(models, t, c)
    b = betas.copy()
    b[5] = np.nan
    e = emu_wrap.emu_eval(pk, nn, b)
    assert np.isinf(e["sse"][0, 5]) and e["g_cond"][0, 5] == 0 and np.all(e["g_neural"][0, 5] == 0) and e["n_fail"] == 1
    e = emu_wrap.emu_eval(pk, nn, betas, maxiters=5)
    assert np.isinf(e["sse"]).all() and e["n_fail"] == 57


def test_elementary_functions_host_build():
    """cude_math.cuh compiled for the host (the MUFU reciprocal seed is replaced by a float-accurate one)."""
    L = C.CDLL(emu_wrap.build())
    D = C.POINTER(C.c_double)
    L.emu_math.argtypes = [C.c_int, C.c_int, D, D]

    def fn(which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        L.emu_math(which, x.size, x.ctypes.data_as(D), y.ctypes.data_as(D))
        return y

    check_math(fn)


def test_mixed_precision_documented_bound(fx):
    """precision = 1 (FP32 network, FP64 integrator).  FP32 noise in the right-hand side perturbs the step-size
    control, so the loss agrees with the FP64 path only to the *solver's own tolerance* (reltol 1e-3): documented
    bound = median 2e-3, 99 % within 5e-2 per trajectory, 2e-3 on a population loss, 1e-2 on population gradients."""
    models, ts, ys = mixed_population(fx)
    pk = cu.pack_models(models, ts, ys)
    rng = np.random.default_rng(2)
    neural, cond = random_starts(rng, pk["chain"], len(models), 4)
    g = oracle.OraclePopulation(pk).eval(neural, cond, grad_mode=0)
    m = emu_wrap.emu_eval(pk, neural, cond, mixed=True)
    d = np.abs(m["sse"] - g["sse"]) / g["sse"]
    assert np.median(d) < 2e-3 and np.percentile(d, 99) < 5e-2
    assert np.abs(m["sse"].sum(axis=1) / g["sse"].sum(axis=1) - 1).max() < 2e-3
    gp, mp = g["g_neural"].sum(axis=1), m["g_neural"].sum(axis=1)
    assert (np.abs(mp - gp) / np.abs(gp).max(axis=1, keepdims=True)).max() < 1e-2
    assert d.max() > 1e-9          # it really is a different arithmetic


def _ragged_obs_population(fx):
    """Individuals with different observation grids: interior-only points, points not aligned with knots, a single
    observation, observations only at the end points."""
    models, t, c = ohashi_models(fx, "train")
    grids = [np.array([0.0, 30.0, 60.0, 90.0, 120.0]), np.array([15.0, 47.5, 101.0]), np.array([120.0]),
             np.array([0.0, 120.0]), np.array([1e-3, 29.999, 30.0, 30.001, 119.5])]
    ms, ts, ys = [], [], []
    rng = np.random.default_rng(9)
    for i in range(20):
        g = grids[i % len(grids)]
        ms.append(models[i]); ts.append(g); ys.append(np.interp(g, t, c[i]) + 0.05 * rng.standard_normal(g.size))
    return ms, ts, ys


def test_ragged_observation_grids(fx):
    ms, ts, ys = _ragged_obs_population(fx)
    pk = cu.pack_models(ms, ts, ys)
    assert pk["max_obs"] == 5 and set(pk["n_obs"]) == {1, 2, 3, 5}
    rng = np.random.default_rng(1)
    neural, cond = random_starts(rng, pk["chain"], len(ms), 3)
    for tol in (DET, dict(abstol=1e-6, reltol=1e-3)):
        g = oracle.OraclePopulation(pk).eval(neural, cond, grad_mode=0, **tol)
        e = emu_wrap.emu_eval(pk, neural, cond, **tol)
        bound = (1e-10, 1e-9) if tol is DET else (1e-5, 1e-3)
        assert relmax(e["sse"], g["sse"]) < bound[0]
        assert relmax(e["g_cond"], g["g_cond"]) < bound[1] and relmax(e["g_neural"], g["g_neural"]) < bound[1]


def test_beta_forward_sensitivity_equals_the_adjoint(fx):
    """grad=2: the beta-only gradient kernel (one forward-sensitivity column carried through the same steps, BSENS in
    cude_kernels.cuh) must give the adjoint's d sse/d cond — both are the exact derivative of the same discrete solve —
    with the loss kernel's forward pass, at default tolerance (adaptive steps, rejections) and for the covariate net."""
    models, ts, ys = mixed_population(fx)                     # Ohashi (5 knots) + Fujita (14 knots, t0 = -10)
    pk = cu.pack_models(models, ts, ys)
    rng = np.random.default_rng(4)
    neural, cond = random_starts(rng, pk["chain"], len(models), 2)
    for kw in (DET, {}):
        a = emu_wrap.emu_eval(pk, neural, cond, grad=1, **kw)
        b = emu_wrap.emu_eval(pk, neural, cond, grad=2, **kw)
        assert np.array_equal(a["sse"], b["sse"]) and a["n_acc"] == b["n_acc"] and a["n_rej"] == b["n_rej"]
        assert relmax(b["g_cond"], a["g_cond"]) < 1e-10
        bf = emu_wrap.emu_eval(pk, neural[0], cond, grad=2, flat=True, **kw)
        assert np.array_equal(bf["g_cond"][0], b["g_cond"][0])
    g = oracle.OraclePopulation(pk).eval(neural, cond, grad_mode=0, **DET)
    assert relmax(emu_wrap.emu_eval(pk, neural, cond, grad=2, **DET)["g_cond"], g["g_cond"]) < 1e-9
    cm, t, c = ohashi_models(fx, "train", covariate=True)
    idx = fx["train_split_idx"][:12]
    pkc = cu.pack_models([cm[i] for i in idx], t, c[idx])
    nn, betas = fx["cov_neural"][1], fx["cov_betas"][1][:12]
    assert relmax(emu_wrap.emu_eval(pkc, nn, betas, grad=2)["g_cond"], emu_wrap.emu_eval(pkc, nn, betas, grad=1)["g_cond"]) < 1e-10
    # failed trajectory: Inf loss, zero gradient
    bad = cond.copy(); bad[0, 3] = np.nan
    e = emu_wrap.emu_eval(pk, neural, bad, grad=2)
    assert np.isinf(e["sse"][0, 3]) and e["g_cond"][0, 3] == 0 and e["n_fail"] == 1


def test_fp32_adjoint_network_mode(fx):
    """precision = 2 (mixed=2 here): the forward pass stays FP64 bit for bit — same loss, same steps — and only the
    adjoint's network evaluations and accumulators are FP32: gradients to ~1e-6 of their scale."""
    models, t, c, nn, betas = train57(fx)
    pk = cu.pack_models(models, t, c)
    rng = np.random.default_rng(6)
    neural, cond = random_starts(rng, pk["chain"], len(models), 3)
    a = emu_wrap.emu_eval(pk, neural, cond)
    b = emu_wrap.emu_eval(pk, neural, cond, mixed=2)
    assert np.array_equal(a["sse"], b["sse"]) and a["n_acc"] == b["n_acc"] and a["n_rej"] == b["n_rej"]
    assert relmax(b["g_cond"], a["g_cond"]) < 1e-5
    gn_a, gn_b = a["g_neural"].sum(axis=1), b["g_neural"].sum(axis=1)        # population gradient per start
    assert np.abs(gn_b - gn_a).max() < 1e-5 * np.abs(gn_a).max()
    assert np.abs(b["g_neural"] - a["g_neural"]).max() < 2e-5 * np.abs(a["g_neural"]).max()


def test_split_gradient_pipeline_stages(fx):
    """csrc/cude_split.cuh compiled for the host: forward solve with step records -> adjoint recursion -> scan -> node kernel
    (one thread per step record) -> finish.  Same discrete adjoint as the fused kernel: sse bit for bit, gradients to
    summation order; and the oracle's in the deterministic regime.  Ragged population (5 and 14 knots / observations)."""
    models, ts, ys = mixed_population(fx)
    pick = [0, 3, 50, 100, 120, 130, 136]
    pk = cu.pack_models([models[i] for i in pick], [ts[i] for i in pick], [ys[i] for i in pick])
    rng = np.random.default_rng(6)
    neural, cond = random_starts(rng, pk["chain"], len(pick), 3)
    pk5 = cu.pack_models([models[i] for i in pick[:4]], [ts[i] for i in pick[:4]], [ys[i] for i in pick[:4]])   # Ohashi only: <= 32 steps
    for pkx, cx, o in ((pk, cond, DET), (pk5, cond[:, :4], dict())):
        f = emu_wrap.emu_eval(pkx, neural, cx, **o)
        s = emu_wrap.emu_eval_split(pkx, neural, cx, **o)
        assert s["n_overflow"] == 0
        assert np.array_equal(s["sse"], f["sse"])
        assert relmax(s["g_cond"], f["g_cond"]) < 1e-12
        assert relmax(s["sums"][:, 0], f["sse"].sum(axis=1)) < 1e-14
        assert relmax(s["sums"][:, 1:], f["g_neural"].sum(axis=1)) < 1e-12
    g = oracle.OraclePopulation(pk).eval(neural, cond, grad_mode=0, **DET)
    s = emu_wrap.emu_eval_split(pk, neural, cond, **DET)
    assert relmax(s["sse"], g["sse"]) < 1e-10 and relmax(s["g_cond"], g["g_cond"]) < 1e-9
    assert relmax(s["sums"][:, 1:], g["g_neural"].sum(axis=1)) < 1e-9
    # a failed trajectory: Inf in the start's sse sum, zero gradient contribution, the others untouched
    bad = cond.copy(); bad[1, 2] = np.nan
    s2 = emu_wrap.emu_eval_split(pk, neural, bad, **DET)
    assert np.isinf(s2["sums"][1, 0]) and s2["g_cond"][1, 2] == 0 and np.allclose(s2["sums"][[0, 2]], s["sums"][[0, 2]], rtol=1e-14)
    # more accepted steps than a record block holds (tight tolerance): flagged for the fused-kernel fallback, not recorded
    s3 = emu_wrap.emu_eval_split(pk, neural[:1], cond[:1], abstol=1e-10, reltol=1e-8)
    assert s3["n_overflow"] == len(pick)


def test_two_kernel_gradient_stages(fx):
    """The default gradient path of large populations compiled for the host: forward kernel with step records and sort keys ->
    per-start stable sort by accepted steps -> cude_adjoint_kernel (records through the cp.async double buffer, observation
    loads two ahead) in the sorted order.  Per-trajectory sse and d/d cond are the fused kernel's bit for bit, sums to
    summation order; the oracle's in the deterministic regime."""
    models, ts, ys = mixed_population(fx)
    pick = [0, 3, 50, 100, 120, 130, 136]
    pk = cu.pack_models([models[i] for i in pick], [ts[i] for i in pick], [ys[i] for i in pick])
    rng = np.random.default_rng(7)
    neural, cond = random_starts(rng, pk["chain"], len(pick), 3)
    pk5 = cu.pack_models([models[i] for i in pick[:4]], [ts[i] for i in pick[:4]], [ys[i] for i in pick[:4]])   # Ohashi only: <= 32 steps
    for pkx, cx, o in ((pk, cond, DET), (pk5, cond[:, :4], dict())):
        f = emu_wrap.emu_eval(pkx, neural, cx, **o)
        e = emu_wrap.emu_eval_exact(pkx, neural, cx, **o)
        assert e["n_overflow"] == 0
        assert np.array_equal(e["sse"], f["sse"]) and np.array_equal(e["g_cond"], f["g_cond"])
        assert relmax(e["sums"][:, 0], f["sse"].sum(axis=1)) < 1e-14
        assert relmax(e["sums"][:, 1:], f["g_neural"].sum(axis=1)) < 1e-12
    g = oracle.OraclePopulation(pk).eval(neural, cond, grad_mode=0, **DET)
    e = emu_wrap.emu_eval_exact(pk, neural, cond, **DET)
    assert relmax(e["sse"], g["sse"]) < 1e-10 and relmax(e["g_cond"], g["g_cond"]) < 1e-9
    assert relmax(e["sums"][:, 1:], g["g_neural"].sum(axis=1)) < 1e-9
    bad = cond.copy(); bad[1, 2] = np.nan
    e2 = emu_wrap.emu_eval_exact(pk, neural, bad, **DET)
    assert np.isinf(e2["sums"][1, 0]) and e2["g_cond"][1, 2] == 0 and np.allclose(e2["sums"][[0, 2]], e["sums"][[0, 2]], rtol=1e-14)
    e3 = emu_wrap.emu_eval_exact(pk, neural[:1], cond[:1], abstol=1e-10, reltol=1e-8)
    assert e3["n_overflow"] == len(pick)


def test_warp_per_trajectory_kernel_with_host_threads_as_lanes(fx):
    """csrc/cude_warp.cuh with its 128 CUDA threads per block as host threads (shuffles = exchanges between barrier waits): the
    forward pass on 5 lanes per step reproduces the fused kernel's sse bit for bit, the lane-parallel adjoint its gradients to
    summation order; counters, failures and the overflow flag (more than 64 accepted steps) as on the device."""
    models, ts, ys = mixed_population(fx)
    pick = [0, 3, 50, 100, 120, 130, 136]
    pk = cu.pack_models([models[i] for i in pick], [ts[i] for i in pick], [ys[i] for i in pick])
    rng = np.random.default_rng(8)
    neural, cond = random_starts(rng, pk["chain"], len(pick), 2)
    for o in (dict(), DET, dict(lanes=8)):                                                # lanes=8: four trajectories per warp
        f = emu_wrap.emu_eval(pk, neural, cond, **{k: v for k, v in o.items() if k != "lanes"})
        w = emu_wrap.emu_warp_eval(pk, neural, cond, **o)
        assert not w["overflow"].any()
        assert np.array_equal(w["sse"], f["sse"]) and np.array_equal(w["row_sse"], f["sse"])
        assert relmax(w["g_cond"], f["g_cond"]) < 1e-12 and relmax(w["g_neural"], f["g_neural"]) < 1e-12
        assert (w["n_acc"], w["n_rej"], w["n_fail"]) == (f["n_acc"], f["n_rej"], f["n_fail"])
    wl = emu_wrap.emu_warp_eval(pk, neural, cond, grad=False, lanes=8)                   # the forward pass alone (loss-only calls)
    assert np.array_equal(wl["sse"], emu_wrap.emu_eval(pk, neural, cond)["sse"]) and wl["n_acc"] == emu_wrap.emu_eval(pk, neural, cond)["n_acc"]
    bad = cond.copy(); bad[1, 2] = np.nan
    w = emu_wrap.emu_warp_eval(pk, neural, bad, lanes=8)
    assert np.isinf(w["sse"][1, 2]) and w["g_cond"][1, 2] == 0 and np.all(w["g_neural"][1, 2] == 0) and w["n_fail"] == 1
    w = emu_wrap.emu_warp_eval(pk, neural[:1], cond[:1], abstol=1e-10, reltol=1e-8)      # every solve beyond 64 steps
    f = emu_wrap.emu_eval(pk, neural[:1], cond[:1], abstol=1e-10, reltol=1e-8)
    assert w["overflow"].all() and np.array_equal(w["sse"], f["sse"]) and np.all(w["row_sse"] == 0) and np.all(w["g_neural"] == 0)
    models, t, c = ohashi_models(fx, "train", covariate=True)                            # covariate network, 3 individuals
    pkc = cu.pack_models(models[:3], t, c[:3])
    neural, cond = random_starts(rng, pkc["chain"], 3, 2)
    f = emu_wrap.emu_eval(pkc, neural, cond)
    w = emu_wrap.emu_warp_eval(pkc, neural, cond)
    assert np.array_equal(w["sse"], f["sse"]) and relmax(w["g_neural"], f["g_neural"]) < 1e-12 and relmax(w["g_cond"], f["g_cond"]) < 1e-12
