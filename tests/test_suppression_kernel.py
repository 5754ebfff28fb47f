"""The suppression kernel variant (csrc/cude_sup_kernel.cuh): CPU tier through the host-compiled kernel source,
GPU tier through the C ABI, both against the oracle (which the reference's stored losses pin)."""
import numpy as np
import pytest

import conditional_ude_b200 as cu
from oracle import oracle
from helpers import noise_ok
import emu_wrap


@pytest.fixture(scope="module")
def sup():
    import os
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return dict(np.load(os.path.join(here, "tests", "golden", "suppression_fixtures.npz")))


def relmax(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def _starts(sup, n, seed=0):
    rng = np.random.default_rng(seed)
    nns = sup["neural_0p01"][:n] + 0.02 * rng.standard_normal((n, 67))
    th = rng.uniform(-1, 1, (n, 37))
    return nns, th


def test_emulated_kernel_matches_oracle(sup):
    data, t = sup["group_data"], sup["timepoints"]
    nns, th = _starts(sup, 3)
    g = oracle.sup_eval(data, t, nns, th, with_grad=True)
    e = emu_wrap.emu_sup_eval(data, t, nns, th)
    assert noise_ok(np.abs(e["sse"] - g["sse"]) / g["sse"], 1e-5)
    assert noise_ok(np.abs(e["g_theta"] - g["g_theta"]) / np.abs(g["g_theta"]).max(), 1e-4)
    assert noise_ok(np.abs(e["g_neural"] - g["g_neural"]) / np.abs(g["g_neural"]).max(axis=-1, keepdims=True), 1e-4)
    assert e["n_acc"] == g["stats"][..., 0].sum() and e["n_fail"] == 0
    e0 = emu_wrap.emu_sup_eval(data, t, nns, th, grad=False)
    assert np.array_equal(e0["sse"], e["sse"])
    # stored lambda = 1 networks: the reference's own stored losses through the kernel source
    e1 = emu_wrap.emu_sup_eval(data, t, sup["neural_1p0"][:5], np.zeros((5, 37)), grad=False)
    loss = e1["sse"].sum(axis=1) / 37 + (sup["neural_1p0"][:5] ** 2).sum(axis=1)
    assert np.abs(loss / sup["losses_1p0"][:5] - 1).max() < 1e-8


def test_emulated_packed_block_path(sup):
    """Small populations run several whole starts per block and reduce through shared memory (SupArgs.spb); with one
    individual the one-thread emulation takes that path: it must agree with the unpacked path's oracle result."""
    data, t = sup["group_data"][:, :, 3:4], sup["timepoints"]
    nns, th = _starts(sup, 4)
    th = th[:, 3:4]
    g = oracle.sup_eval(data, t, nns, th, with_grad=True, scale=np.ones(3))
    e = emu_wrap.emu_sup_eval(data, t, nns, th, scale=np.ones(3))
    assert noise_ok(np.abs(e["sse"] - g["sse"]) / g["sse"], 1e-5)
    assert noise_ok(np.abs(e["g_theta"] - g["g_theta"]) / np.abs(g["g_theta"]).max(), 1e-4)
    assert noise_ok(np.abs(e["g_neural"] - g["g_neural"]) / np.abs(g["g_neural"]).max(axis=-1, keepdims=True), 1e-4)


def test_emulated_two_kernel_gradient(sup):
    """The default gradient path as two kernels (grad=2 in the emulation): the loss kernel leaving step records, weighted
    residuals, step counts and sse in global memory, then the adjoint sweep over them — bit for bit the fused kernel's results
    (same sweep code, same records); solves beyond the 64-entry record block are only counted (the host redoes those calls)."""
    data, t = sup["group_data"][:, :, :5], sup["timepoints"]
    nns, th = _starts(sup, 3)
    th = th[:, :5]
    f = emu_wrap.emu_sup_eval(data, t, nns, th)
    e = emu_wrap.emu_sup_eval(data, t, nns, th, grad=2)
    assert e["n_overflow"] == 0 and (e["n_acc"], e["n_rej"]) == (f["n_acc"], f["n_rej"])
    assert all(np.array_equal(e[k], f[k]) for k in ("sse", "g_neural", "g_theta"))
    e = emu_wrap.emu_sup_eval(data, t, nns[:1], th[:1], grad=2, abstol=1e-10, reltol=1e-8)
    assert e["n_overflow"] == 5


def test_emulated_step_ring_replay(sup, tmp_path):
    """The gradient pass keeps a ring of accepted-step records and replays the forward pass in chunks when a solve has more
    steps than the ring (round 1 returned Inf beyond 512 steps).  Rebuilt with a 4-entry ring, the ~25-step solves need
    six replays; gradients must equal the 64-entry build's bit for bit (same arithmetic, same order)."""
    import ctypes as C
    import importlib
    import os
    import subprocess
    lib = str(tmp_path / "libcude_emu_supcap4.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-DCUDE_SUP_REC_CAP=4", "-o", lib, os.path.join(emu_wrap._HERE, "emu_kernel.cpp")])
    data, t = sup["group_data"][:, :, :6], sup["timepoints"]
    nns, th = _starts(sup, 2)
    th = th[:, :6]
    sc = oracle.suppression_scale(sup["group_data"])
    ref = emu_wrap.emu_sup_eval(data, t, nns, th, scale=sc)
    assert ref["n_acc"] > 12 * 20
    old = emu_wrap.LIB
    try:
        emu_wrap.LIB = lib
        emu_wrap.build = lambda: lib
        e = emu_wrap.emu_sup_eval(data, t, nns, th, scale=sc)
    finally:
        emu_wrap.LIB = old
        importlib.reload(emu_wrap)
    assert np.array_equal(e["sse"], ref["sse"]) and e["n_acc"] == ref["n_acc"]
    assert np.array_equal(e["g_theta"], ref["g_theta"]) and np.array_equal(e["g_neural"], ref["g_neural"])


@pytest.mark.gpu
def test_gpu_tight_tolerance_replays_instead_of_failing(sup):
    """reltol 1e-9: ~300 accepted steps per trajectory, several ring lengths: finite and equal to the oracle's gradient."""
    data, t = sup["group_data"], sup["timepoints"]
    nns, th = _starts(sup, 2)
    o = dict(abstol=1e-11, reltol=1e-9)
    g = oracle.sup_eval(data, t, nns, th, with_grad=True, **o)
    pop = cu.SuppressionPopulation(data, t, ctx=cu.Context(0))
    loss, gn, gt, sse = pop.loss_grad(nns, th, lam=0.0, opts=cu.SolverOptions(**o), return_sse=True)
    assert g["stats"][..., 0].min() > 64 and np.isfinite(loss).all()
    assert relmax(sse, g["sse"]) < 1e-6
    assert relmax(gt * 37, g["g_theta"]) < 1e-4 and relmax(gn * 37, g["g_neural"].sum(axis=1)) < 1e-4


@pytest.mark.gpu
def test_gpu_reproduces_stored_reference_losses(sup):
    """All 25 stored training and validation losses of suppression/results/lambda=1.0.jld2 on the B200."""
    data, vdata, t, nns = sup["group_data"], sup["validation_data"], sup["timepoints"], sup["neural_1p0"]
    pop = cu.SuppressionPopulation(data, t)
    loss = pop.loss(nns, np.zeros((25, 37)), lam=1.0)
    assert np.abs(loss / sup["losses_1p0"] - 1).max() < 1e-8
    vpop = cu.SuppressionPopulation(vdata, t)
    assert np.abs(vpop.loss(nns, np.zeros((25, 30)), lam=0.0) / sup["losses_valid_1p0"] - 1).max() < 1e-8
    # reference-named entry point
    p = cu.ComponentVector(neural=nns[0], theta=np.zeros(37))
    assert abs(cu.suppression_loss(p, (None, data, t, 1.0)) / sup["losses_1p0"][0] - 1) < 1e-8
    assert cu.neural_network_model(5, 3, input_dims=4).n_params == 67


@pytest.mark.gpu
def test_gpu_loss_and_gradient_match_oracle(sup):
    data, t = sup["group_data"], sup["timepoints"]
    nns, th = _starts(sup, 8, seed=1)
    lam = 0.01
    pop = cu.SuppressionPopulation(data, t)
    g = oracle.sup_eval(data, t, nns, th, with_grad=True)
    loss, gn, gt, sse = pop.loss_grad(nns, th, lam, return_sse=True)
    assert noise_ok(np.abs(sse - g["sse"]) / g["sse"], 1e-5)
    ref_loss = g["sse"].sum(axis=1) / 37 + lam * (nns ** 2).sum(axis=1)
    ref_gn = g["g_neural"].sum(axis=1) / 37 + 2 * lam * nns
    assert relmax(loss, ref_loss) < 1e-5 and relmax(gn, ref_gn) < 1e-4 and relmax(gt, g["g_theta"] / 37) < 1e-4
    # shared network (stride 0) and loss-only path
    l1 = pop.loss(nns[0], th, lam)
    l2 = pop.loss(np.tile(nns[0], (8, 1)), th, lam)
    assert np.array_equal(l1, l2) and l1[0] == loss[0]
    a = pop.loss_grad(nns, th, lam)
    assert all(np.array_equal(x, y) for x, y in zip(a, (loss, gn, gt)))       # run-to-run determinism
    # failure: NaN parameter -> Inf loss, zero gradient
    th2 = th.copy(); th2[3, 5] = np.nan
    loss, gn, gt = pop.loss_grad(nns, th2, lam)
    assert np.isinf(loss[3]) and np.all(gn[3] == 0) and np.all(gt[3] == 0) and np.isfinite(np.delete(loss, 3)).all()


@pytest.mark.gpu
def test_gpu_two_kernel_gradient_equals_the_fused_kernel(sup):
    """The default gradient path (forward kernel leaving records + adjoint kernel with its block-local sort by step count)
    against the fused kernel (opts.split = 1): per-trajectory sse and d/d theta bit for bit, with flat indexing (37
    individuals) and with one start per block row (300 individuals: the tile geometry, where per-start rows are summed by
    warp shuffles in whatever order the sort left, hence 1e-13); a call with solves beyond the 64-record block falls back to
    the fused kernel by itself."""
    data, t = sup["group_data"], sup["timepoints"]
    rng = np.random.default_rng(5)
    nns = sup["neural_0p01"][rng.integers(0, 25, 40)] + 0.02 * rng.standard_normal((40, 67))
    for d in (data, np.concatenate([data] * 9, axis=2)[:, :, :300]):
        n = d.shape[2]
        th = rng.uniform(-1, 1, (40, n))
        pop = cu.SuppressionPopulation(d, t)
        a = pop.loss_grad(nns, th, 0.01, return_sse=True)
        la = pop.ctx.stats()["launches"]
        f = pop.loss_grad(nns, th, 0.01, opts=cu.SolverOptions(split=1), return_sse=True)
        assert la == 3 and pop.ctx.stats()["launches"] == 2
        assert np.array_equal(a[3], f[3]) and np.array_equal(a[2], f[2])
        assert relmax(a[0], f[0]) < 1e-14 and relmax(a[1], f[1]) < 1e-12
    tight = cu.SolverOptions(abstol=1e-10, reltol=1e-7)
    th = rng.uniform(-1, 1, (4, 37))
    pop = cu.SuppressionPopulation(data, t)
    a = pop.loss_grad(nns[:4], th, 0.01, opts=tight, return_sse=True)
    assert pop.ctx.stats()["launches"] == 4                                    # 2 (two-kernel attempt) + 1 (fused) + reduction
    f = pop.loss_grad(nns[:4], th, 0.01, opts=cu.SolverOptions(abstol=1e-10, reltol=1e-7, split=1), return_sse=True)
    assert all(np.array_equal(x, y) for x, y in zip(a, f))


@pytest.mark.gpu
def test_gpu_small_scratch_budget_cuts_the_batch_into_launches(sup, monkeypatch):
    """The step rings of a gradient launch live in global scratch; a batch whose rings exceed the budget is cut into several
    launches over block ranges (here: 8 MB = 5 blocks per launch, 58 blocks) — same results bit for bit."""
    data, t = sup["group_data"], sup["timepoints"]
    rng = np.random.default_rng(4)
    nns = sup["neural_0p01"][rng.integers(0, 25, 200)] + 0.02 * rng.standard_normal((200, 67))
    th = rng.uniform(-1, 1, (200, 37))
    a = cu.SuppressionPopulation(data, t).loss_grad(nns, th, 0.01, return_sse=True)
    monkeypatch.setenv("CUDE_SCRATCH_BYTES", str(8 << 20))
    ctx = cu.Context(0)                                   # a fresh context decides its budget at first use
    b = cu.SuppressionPopulation(data, t, ctx=ctx).loss_grad(nns, th, 0.01, return_sse=True)
    assert ctx.stats()["launches"] > 5
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.gpu
def test_gpu_fit_and_validate(sup):
    """fit_suppression_model / validate_suppression_model in miniature (reference: 10 000 initials, 2000 + 2000 iterations)."""
    data, vdata, t = sup["group_data"], sup["validation_data"], sup["timepoints"]
    rng = np.random.default_rng(0)
    net = cu.neural_network_model(5, 3, input_dims=4)
    p_init = [cu.ComponentVector(theta=rng.standard_normal(37), neural=net.init_params(rng)) for _ in range(64)]
    pop = cu.SuppressionPopulation(data, t)
    sols, traces = cu.fit_suppression_model(p_init, pop, data, t, 0.01, select_best_n=3, adam_iters=40, lbfgs_iters=40)
    assert len(sols) == 3 and all(s.u.neural.shape == (67,) and s.u.theta.shape == (37,) for s in sols)
    for s, tr in zip(sols, traces):
        assert s.objective < tr[0]
        assert abs(cu.suppression_loss(s.u, (pop, data, t, 0.01)) - s.objective) < 1e-9
    th, obj = cu.validate_suppression_model([rng.random(30) for _ in range(16)], None, vdata, t, sols[0].u.neural, lbfgs_iters=30)
    assert th.shape == (30,) and np.isfinite(obj)
