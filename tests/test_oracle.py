"""CPU tests of the oracle (oracle/cude_oracle.cpp) — the checker itself.

The reference has no tests or golden vectors and cannot run here (Julia absent): these tests pin the
restatement indirectly — component formulas against independent numpy code, the integrator against an
independent scipy DOP853 solve at tight tolerance, the published step statistics of SURVEY.md App. D,
the gradient against finite differences, and the Tsit5 order conditions of the tableau constants.
"""
import math
import os
import re

import numpy as np
import pytest
from scipy.integrate import solve_ivp

import conditional_ude_b200 as cu
from oracle import oracle
from helpers import train57, mixed_population, ohashi_models, random_starts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def np_mlp(p, x, n_in=2, depth=2, width=4):
    a, off = np.asarray(x, float), 0
    for _ in range(depth):
        W = p[off:off + width * a.size].reshape(a.size, width).T      # column-major [out x in]
        b = p[off + width * a.size: off + width * (a.size + 1)]
        off += width * (a.size + 1)
        a = np.tanh(W @ a + b)
    z = p[off:off + a.size] @ a + p[off + a.size]
    return math.log(1.0 + math.exp(z))


def test_van_cauter_parameters():
    # src/c-peptide-models.jl:30-42, worked by hand for (age 40, NGT) and (age 60, T2DM)
    for age, t2dm in ((40.0, False), (60.0, True), (29.0, False)):
        short, frac = (4.52, 0.78) if t2dm else (4.95, 0.76)
        long_ = 0.14 * age + 29.2
        k1 = frac * math.log(2) / long_ + (1 - frac) * math.log(2) / short
        k0 = (math.log(2) / short) * (math.log(2) / long_) / k1
        k2 = math.log(2) / short + math.log(2) / long_ - k0 - k1
        for got in (oracle.van_cauter(age, t2dm), cu.van_cauter_parameters(age, t2dm)):
            assert np.allclose(got, (k0, k1, k2), rtol=1e-15)
    # eigenvalues of the kinetic matrix are -ln2/short and -ln2/long (van Cauter's two half-lives)
    k0, k1, k2 = oracle.van_cauter(40.0, False)
    ev = np.sort(np.linalg.eigvals(np.array([[-(k0 + k2), k1], [k2, -k1]])))
    assert np.allclose(ev, [-math.log(2) / 4.95, -math.log(2) / (0.14 * 40 + 29.2)], rtol=1e-12)


def test_mlp_layout_and_value(fx):
    rng = np.random.default_rng(0)
    for n_in, P in ((2, 37), (3, 41)):
        assert oracle.lib().cude_oracle_nparams(n_in, 2, 4) == P == cu.chain(4, 2, "tanh", input_dims=n_in).n_params
        for _ in range(5):
            p = rng.standard_normal(P)
            x = rng.standard_normal(n_in)
            assert abs(oracle.mlp(n_in, 2, 4, p, x) - np_mlp(p, x, n_in)) < 1e-14


def test_linear_interpolation():
    kt = np.array([-10.0, 0, 10, 20, 30, 45, 60])
    kg = np.array([5.0, 5.0, 5.7, 8.4, 10.0, 11.5, 11.7])
    for tau in np.linspace(-10, 60, 141):
        assert abs(oracle.glucose(kt, kg, tau) - np.interp(tau, kt, kg)) < 1e-13
    assert oracle.glucose(kt, kg, 10.0) == 5.7          # knot value itself (searchsortedlast)


def _tableau(path, prefix):
    src = open(path).read()
    vals = {}
    for name, v in re.findall(r"\b((?:%s)\d+)\s*=\s*(-?[0-9][0-9.eE+-]*)" % prefix, src):
        vals[name.lower()] = float(v)
    return vals


def test_tsit5_tableau_order_conditions():
    """The constants typed into the oracle and into the CUDA header are identical and satisfy the
    Tsit5 order conditions (5th order for b, 4th for b - btilde... sum btilde = 0, FSAL row = b)."""
    o = _tableau(os.path.join(ROOT, "oracle", "cude_oracle.cpp"), "[ACR]|BT")
    k = _tableau(os.path.join(ROOT, "conditional_ude_b200", "csrc", "cude_kernels.cuh"), "[abcer]")
    c = np.array([0.0, o["c2"], o["c3"], o["c4"], o["c5"], 1.0, 1.0])
    A = np.zeros((7, 7))
    for i in range(2, 8):
        for j in range(1, i):
            A[i - 1, j - 1] = o[f"a{i}{j}"]
    b = A[6].copy()
    bt = np.array([o[f"bt{i}"] for i in range(1, 8)])
    # kernel constants equal the oracle's
    for i in range(2, 7):
        for j in range(1, i):
            assert k[f"a{i}{j}"] == o[f"a{i}{j}"]
    assert [k[f"b{j}"] for j in range(1, 7)] == list(b[:6])
    assert [k[f"e{j}"] for j in range(1, 8)] == list(bt)
    for j in range(1, 8):
        for m in (2, 3, 4):
            if j == 1 and m == 2:
                assert k["r11"] == o["r11"] and k["r12"] == o["r12"] and k["r13"] == o["r13"] and k["r14"] == o["r14"]
            elif j > 1:
                assert k[f"r{j}{m}"] == o[f"r{j}{m}"]
    assert np.allclose(A.sum(axis=1), c, atol=1e-14)
    # order conditions up to 5 for b
    Ac = A @ c
    conds = [(b.sum(), 1), (b @ c, 1 / 2), (b @ c**2, 1 / 3), (b @ Ac, 1 / 6), (b @ c**3, 1 / 4), (b @ (c * Ac), 1 / 8),
             (b @ (A @ c**2), 1 / 12), (b @ (A @ Ac), 1 / 24), (b @ c**4, 1 / 5), (b @ (c**2 * Ac), 1 / 10),
             (b @ (Ac * Ac), 1 / 20), (b @ (c * (A @ c**2)), 1 / 15), (b @ (A @ c**3), 1 / 20),
             (b @ (c * (A @ Ac)), 1 / 30), (b @ (A @ (c * Ac)), 1 / 40), (b @ (A @ (A @ c**2)), 1 / 60),
             (b @ (A @ (A @ Ac)), 1 / 120)]
    for got, want in conds:
        assert abs(got - want) < 1e-12
    assert abs(bt.sum()) < 1e-15 and abs(bt @ c) < 1e-13 and abs(bt @ c**2) < 1e-13 and abs(bt @ c**3) < 1e-13
    # dense output: b_theta(1) = b (with b7 = 0), and order 4 at theta = 0.3, 0.7
    def btheta(th):
        r = [[o["r11"], o["r12"], o["r13"], o["r14"]]] + [[0.0, o[f"r{j}2"], o[f"r{j}3"], o[f"r{j}4"]] for j in range(2, 8)]
        return np.array([th * (r[0][0] + th * (r[0][1] + th * (r[0][2] + th * r[0][3])))] +
                        [th * th * (r[j][1] + th * (r[j][2] + th * r[j][3])) for j in range(1, 7)])
    assert np.allclose(btheta(1.0), np.append(b[:6], 0.0), atol=1e-13)
    for th in (0.3, 0.7):
        bb = btheta(th)
        assert abs(bb.sum() - th) < 1e-13 and abs(bb @ c - th**2 / 2) < 1e-13
        assert abs(bb @ c**2 - th**3 / 3) < 1e-13 and abs(bb @ Ac - th**3 / 6) < 1e-13
        assert abs(bb @ c**3 - th**4 / 4) < 1e-12


def test_step_statistics_match_survey(fx):
    """SURVEY.md App. D (an independent Python restatement): 57 train individuals at the stored weights:
    accepted steps mean 20.0 (17..22), rejected mean 0.2, 123 RHS evaluations; mean loss 0.4281 at the
    default tolerance vs 0.4273 converged; every trajectory opens with dt = 1e-4, 1e-3, 1e-2, 0.1, 1.0."""
    models, t, c, nn, betas = train57(fx)
    op = oracle.OraclePopulation(cu.pack_models(models, t, c))
    r = op.eval(nn, betas)
    nacc, nrej, nrhs = r["stats"][0, :, 0], r["stats"][0, :, 1], r["stats"][0, :, 2]
    assert abs(nacc.mean() - 20.0) < 0.1 and nacc.min() == 17 and nacc.max() == 22
    assert abs(nrej.mean() - 0.2) < 0.05 and nrej.max() == 3
    assert abs(nrhs.mean() - 123) < 1
    assert abs(r["sse"].mean() - 0.4281) < 1e-4
    tight = op.eval(nn, betas, abstol=1e-12, reltol=1e-10)
    assert abs(tight["sse"].mean() - 0.4273) < 1e-4
    rel = np.abs(r["sse"] - tight["sse"]) / tight["sse"]
    assert abs(np.median(rel) - 6.6e-3) < 5e-4 and abs(rel.max() - 4.0e-2) < 2e-3
    rows, _ = op.trace(0, nn, betas[0])
    assert np.allclose(rows[:5, 1], [1e-4, 1e-3, 1e-2, 0.1, 1.0], rtol=1e-12)
    assert rows[-1, 0] + rows[-1, 1] == pytest.approx(120.0, abs=1e-9)


def _scipy_solution(model, nn, beta, t_eval):
    k0, k1, k2, c0 = model.k0, model.k1, model.k2, model.c0
    kt, kg = model.glucose_timepoints, model.glucose_data
    b = math.exp(beta)
    nn0 = np_mlp(nn, [0.0, b])

    def rhs(t, u):
        dG = np.interp(t, kt, kg) - kg[0]
        prod = np_mlp(nn, [dG, b]) - nn0
        return [-(k0 + k2) * u[0] + k1 * u[1] + k0 * c0 + prod, -k1 * u[1] + k2 * u[0]]

    # integrate segment by segment so that DOP853 never steps across a kink of the glucose interpolant
    u, out = np.array([c0, k2 / k1 * c0]), {kt[0]: c0}
    for a, bnd in zip(kt[:-1], kt[1:]):
        te = [x for x in t_eval if a < x <= bnd]
        sol = solve_ivp(rhs, (a, bnd), u, method="DOP853", rtol=1e-12, atol=1e-14, t_eval=sorted(set(te + [bnd])))
        for tt, yy in zip(sol.t, sol.y[0]):
            out[tt] = yy
        u = sol.y[:, -1]
    return np.array([out[x] for x in t_eval])


def test_converges_to_independent_scipy_solution(fx):
    """Tight-tolerance oracle solution == scipy DOP853 on an independently written RHS (Ohashi 5 knots
    and Fujita 14 knots, t0 = -10)."""
    models, ts, ys = mixed_population(fx)
    nn = fx["cude_neural"][13]
    pick = [0, 40, 90, 117, 125, 136]          # Ohashi train/test and Fujita individuals
    sub = [models[i] for i in pick]
    pk = cu.pack_models(sub, [ts[i] for i in pick], [ys[i] for i in pick])
    betas = np.array([-1.0, -0.5, -1.7, -0.2, -1.2, -0.8])
    r = oracle.OraclePopulation(pk).eval(nn, betas, abstol=1e-13, reltol=1e-11, want_yhat=True)
    for j, i in enumerate(pick):
        want = _scipy_solution(models[i], nn, betas[j], list(ts[i]))
        got = r["yhat"][0, j, :len(ts[i])]
        assert np.allclose(got, want, rtol=2e-8, atol=1e-10), (i, np.abs(got - want).max())


def test_gradient_matches_finite_differences(fx):
    """Frozen-primal tangents are the exact derivative of the discrete solve.  In the deterministic regime
    (abstol = reltol = 1e12: all steps accepted, controller clamped at qmax, step sequence independent of theta)
    the loss is a smooth function and central differences must agree to ~1e-8; at tight tolerance the
    frozen-step gradient converges to the gradient of the true solution (checked to 1e-3: finite
    differences of an adaptive solve are noisy, see test_noise_floor)."""
    models, t, c, nn, betas = train57(fx)
    sub = [3, 11, 30]
    pk = cu.pack_models([models[i] for i in sub], t, c[sub])
    op = oracle.OraclePopulation(pk)

    def fd4(f, h):
        return (-f(2 * h) + 8 * f(h) - 8 * f(-h) + f(-2 * h)) / (12 * h)

    for tol, h, rtol, atol in ((dict(abstol=1e12, reltol=1e12), 1e-3, 1e-7, 1e-9),
                               (dict(abstol=1e-13, reltol=1e-11), 2e-3, 1e-3, 1e-5)):
        g = op.eval(nn, betas[sub], grad_mode=0, **tol)
        for p in (0, 5, 9, 17, 30, 33, 36):
            e = np.zeros(37); e[p] = 1.0
            fd = fd4(lambda d: op.eval(nn + d * e, betas[sub], **tol)["sse"][0], h)
            assert np.allclose(g["g_neural"][0, :, p], fd, rtol=rtol, atol=atol), (tol, p, g["g_neural"][0, :, p], fd)
        fd = fd4(lambda d: op.eval(nn, betas[sub] + d, **tol)["sse"][0], h)
        assert np.allclose(g["g_cond"][0], fd, rtol=rtol, atol=atol)


def test_noise_floor(fx):
    """An adaptive solve is not a smooth function of its inputs at round-off level: moving beta by one
    ulp changes step sizes (the error estimate is cancellation-dominated) and sometimes an accept/reject
    decision.  This measures the floor any second implementation (CUDA, or Julia itself) sits on, and
    justifies the tolerances of the parity tests: typical ~1e-10, tails up to the contract."""
    models, t, c, nn, betas = train57(fx)
    op = oracle.OraclePopulation(cu.pack_models(models, t, c))
    a = op.eval(nn, betas, grad_mode=0)
    b = op.eval(nn, np.nextafter(betas, 1.0), grad_mode=0)
    d = np.abs(a["sse"] - b["sse"]) / a["sse"]
    dg = np.abs(a["g_cond"] - b["g_cond"]) / np.abs(a["g_cond"]).max()
    assert d.max() > 1e-13, "the solve is unexpectedly smooth at round-off level"
    assert np.median(d) < 1e-8 and d.max() < 1e-4
    assert np.median(dg) < 1e-7 and dg.max() < 1e-2
    # in the deterministic regime (everything accepted, controller clamped) it IS smooth
    a = op.eval(nn, betas, grad_mode=0, abstol=1e3, reltol=1e3)
    b = op.eval(nn, np.nextafter(betas, 1.0), grad_mode=0, abstol=1e3, reltol=1e3)
    assert (np.abs(a["sse"] - b["sse"]) / a["sse"]).max() < 1e-10


def test_forwarddiff_twin_mode_documents_the_gap(fx):
    """Mode 1 (error norm sees the partials, theta processed in ForwardDiff chunks — the semantic twin of
    AutoForwardDiff) differs from the frozen-primal gradient only by solver-tolerance effects, and the two
    coincide at tight tolerance."""
    models, t, c, nn, betas = train57(fx)
    sub = list(range(12))
    pk = cu.pack_models([models[i] for i in sub], t, c[sub])
    op = oracle.OraclePopulation(pk)
    g0 = op.eval(nn, betas[sub], grad_mode=0)
    g1 = op.eval(nn, betas[sub], grad_mode=1)
    gt = op.eval(nn, betas[sub], grad_mode=0, abstol=1e-12, reltol=1e-10)     # converged gradient

    def rel(a, b):
        return np.abs(a - b).max() / np.abs(b).max()

    gap = rel(g1["g_neural"], g0["g_neural"])
    assert 1e-9 < gap < 0.5
    # neither semantics is closer to the converged gradient than the solver error (~reltol-driven, few %)
    assert rel(g0["g_neural"], gt["g_neural"]) < 0.3 and rel(g1["g_neural"], gt["g_neural"]) < 0.3
    # the twin takes more steps: its error norm also sees the (larger) partials
    assert g1["stats"][..., 0].mean() > g0["stats"][..., 0].mean()
    tol = dict(abstol=1e-12, reltol=1e-10)
    g0 = op.eval(nn, betas[sub], grad_mode=0, **tol)
    g1 = op.eval(nn, betas[sub], grad_mode=1, **tol)
    assert np.abs(g1["g_neural"] - g0["g_neural"]).max() / np.abs(g0["g_neural"]).max() < 1e-5
    assert np.abs(g1["g_cond"] - g0["g_cond"]).max() / np.abs(g0["g_cond"]).max() < 1e-5


def test_failure_returns_inf(fx):
    models, t, c, nn, betas = train57(fx)
    op = oracle.OraclePopulation(cu.pack_models(models, t, c))
    r = op.population_loss(nn, betas[None], maxiters=5)
    assert np.isinf(r["loss"][0])
    b = betas.copy(); b[3] = np.nan
    r = op.eval(nn, b, grad_mode=0)
    assert np.isinf(r["sse"][0, 3]) and r["stats"][0, 3, 3] != 0 and r["g_cond"][0, 3] == 0
    assert np.isfinite(np.delete(r["sse"][0], 3)).all()
    assert np.isinf(op.population_loss(nn, b[None])["loss"][0])      # parameter-estimation.jl:134-136


def test_population_loss_is_mean(fx):
    models, t, c, nn, betas = train57(fx)
    op = oracle.OraclePopulation(cu.pack_models(models, t, c))
    rng = np.random.default_rng(3)
    neural, cond = random_starts(rng, cu.chain(4, 2, "tanh"), 57, 3)
    r = op.eval(neural, cond, grad_mode=0)
    p = op.population_loss(neural, cond, with_grad=True)
    assert np.allclose(p["loss"], r["sse"].mean(axis=1), rtol=1e-14)
    assert np.allclose(p["g_neural"], r["g_neural"].mean(axis=1), rtol=1e-12, atol=1e-15)
    assert np.allclose(p["g_cond"], r["g_cond"] / 57, rtol=1e-14)
