"""Shared builders for the tests: reference-style model vectors from the golden fixtures."""
import numpy as np

import conditional_ude_b200 as cu


def ohashi_models(fx, split="train", covariate=False):
    net = cu.chain(4, 2, "tanh", input_dims=3 if covariate else 2)
    ctor = cu.CPeptideConditionalCovariateUDEModel if covariate else cu.CPeptideConditionalUDEModel
    g, c = fx[f"ohashi_{split}_glucose"], fx[f"ohashi_{split}_cpeptide"]
    ages, t2dm, t = fx[f"ohashi_{split}_ages"], fx[f"ohashi_{split}_t2dm"], fx["ohashi_timepoints"]
    # 02-conditional.jl:26-28
    models = [ctor(g[i], t, float(ages[i]), net, c[i], bool(t2dm[i])) for i in range(g.shape[0])]
    return models, t, c


def fujita_models(fx):
    net = cu.chain(4, 2, "tanh")
    g, c, t = fx["fujita_glucose"], fx["fujita_cpeptide"], fx["fujita_timepoints"]
    models = [cu.CPeptideConditionalUDEModel(g[i], t, 29.0, net, c[i], False) for i in range(g.shape[0])]
    return models, t, c


def train57(fx):
    """The 57-individual training split with stored weights 14 and betas (config 1)."""
    models, t, c = ohashi_models(fx, "train")
    idx = fx["train_split_idx"]
    best = int(fx["cude_best_model_index"]) - 1
    return [models[i] for i in idx], t, c[idx], fx["cude_neural"][best], fx["cude_betas"][best]


def mixed_population(fx):
    """Ohashi train+test (5 knots) and Fujita (14 knots, t0 = -10): ragged knots and observations."""
    m1, t1, c1 = ohashi_models(fx, "train")
    m2, t2, c2 = ohashi_models(fx, "test")
    m3, t3, c3 = fujita_models(fx)
    models = m1 + m2 + m3
    ts = [t1] * len(m1) + [t2] * len(m2) + [t3] * len(m3)
    ys = [c1[i] for i in range(len(m1))] + [c2[i] for i in range(len(m2))] + [c3[i] for i in range(len(m3))]
    return models, ts, ys


def random_starts(rng, chain, n_ind, n_starts, scale=1.0):
    neural = np.stack([chain.init_params(rng) * scale for _ in range(n_starts)])
    cond = rng.uniform(-2.0, 0.0, size=(n_starts, n_ind))
    return neural, cond


def noise_ok(d, contract):
    """Distribution test for an adaptive solve (see test_oracle.py::test_noise_floor): typical agreement at
    round-off level, the contract for 99 % of the entries, and the rare accept/reject flips bounded by the
    solver tolerance."""
    d = np.asarray(d)
    return bool(np.median(d) < 1e-8 and np.percentile(d, 99) < contract and d.max() < 2e-2)


def _math_cases():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-25, 25, 20000), rng.uniform(-1, 1, 20000), rng.uniform(-1e-3, 1e-3, 2000),
                        [0.0, -0.0, 19.9, 20.0, 20.1, -20.0, 36.7, 36.9, 40.0, -40.0, 700.0, -700.0, 1e300, -1e300, np.inf, -np.inf]])
    return x


def check_math(fn):
    """fn(which, x) -> y: accuracy of the kernels' elementary functions against numpy/mpmath-grade references
    (ids: cude_math_probe_eval in cude_kernels.cuh)."""
    x = _math_cases()
    t = fn(0, x)
    assert np.abs(t - np.tanh(x)).max() < 4e-16                      # absolute (see cude_math.cuh)
    assert np.isnan(fn(0, np.array([np.nan]))[0])
    xs = x[np.isfinite(x)]
    sp = fn(1, xs)
    big = xs > 709.782712893384                                      # exp overflows in the reference's naive form
    assert np.all(np.isinf(sp[big]))
    xs, sp = xs[~big], sp[~big]
    ref = np.where(xs > 36.8, xs, np.log1p(np.exp(np.minimum(xs, 36.8))))
    assert np.all(np.abs(sp - ref) <= 8e-16 * np.maximum(1.0, np.abs(ref)))      # ~2 ulp
    assert np.isinf(fn(1, np.array([710.0]))[0]) and fn(1, np.array([709.0]))[0] == 709.0      # naive-form overflow
    assert np.isnan(fn(1, np.array([np.nan]))[0])
    xs = x[np.isfinite(x)]
    with np.errstate(over="ignore"):
        sig = 1.0 / (1.0 + np.exp(-xs))
    assert np.abs(fn(2, xs) - sig).max() < 4e-16                     # 1 - 1/d from the softplus' d (c-peptide adjoint)
    assert np.abs(fn(6, xs) - sig).max() < 4e-16                     # sigmoid (suppression adjoint)
    xe = np.random.default_rng(1).uniform(-40, 40, 20000)
    assert np.abs(fn(3, xe) / np.exp(xe) - 1).max() < 5e-15          # one-constant reduction: |x| 1.1e-16 relative
    xl = np.exp(np.random.default_rng(2).uniform(-600, 600, 20000))
    assert np.abs(fn(4, xl) - np.log(xl)).max() < 3e-13 and np.abs(fn(4, xl) / np.log(xl) - 1)[np.abs(np.log(xl)) > 1e-3].max() < 1e-15
    xr = np.exp(np.random.default_rng(3).uniform(-30, 30, 20000))
    assert np.abs(fn(5, xr) * xr - 1).max() < 5e-16


class OraclePopulationAdapter:
    """Test double with the Population interface (loss / loss_grad) computed by the CPU oracle — lets the
    CPU-only tier exercise conditional_ude_b200.estimation end to end.  Tests only."""

    def __init__(self, models, timepoints, cpeptide_data):
        from oracle import oracle
        pk = cu.pack_models(models, timepoints, cpeptide_data)
        self.op = oracle.OraclePopulation(pk)
        self.chain, self.n_ind, self.n_params = pk["chain"], pk["n_ind"], pk["chain"].n_params
        self.n_obs = np.asarray(pk["n_obs"]).copy()
        self.t_first = np.asarray(pk["knot_t"], dtype=np.float64)[:, 0].copy()
        self.calls = 0
        self.traj = 0            # trajectories evaluated (start-sharding tests)

    def loss(self, neural, cond, opts=None, return_sse=False):
        self.calls += 1
        self.traj += int(np.asarray(cond).size)
        r = self.op.eval(neural, cond)
        loss = r["sse"].mean(axis=1)
        return (loss, r["sse"]) if return_sse else loss

    def loss_grad(self, neural, cond, opts=None, neural_grad=True, mean=True, return_sse=False):
        self.calls += 1
        self.traj += int(np.asarray(cond).size)
        r = self.op.eval(neural, cond, grad_mode=0)
        sc = 1.0 / self.n_ind if mean else 1.0
        out = (r["sse"].sum(axis=1) * sc, r["g_neural"].sum(axis=1) * sc if neural_grad else None, r["g_cond"] * sc)
        return out + (r["sse"],) if return_sse else out
