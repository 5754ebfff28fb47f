"""Non-conditional UDE (CPeptideUDEModel, src/c-peptide-models.jl:144-168; train :211-247): a 1-input network run through
the conditional kernels by embedding its parameters (zero beta column).  CPU tier: the embedding is exact (numpy MLPs),
the oracle-backed loss is independent of the dummy conditional parameter, train works on a test double; GPU tier: loss,
gradient (against finite differences of the loss at tight tolerance) and `train` on the device."""
import numpy as np
import pytest

import conditional_ude_b200 as cu
from conditional_ude_b200.models import embed_ude_parameters, extract_ude_gradient
from helpers import train57, OraclePopulationAdapter


def _mlp(p, x, n_in, width=4, depth=2):
    """SimpleChains layout: per layer W[out x in] column-major then bias; tanh hidden, softplus output."""
    a, off = np.asarray(x, dtype=np.float64), 0
    for l in range(depth):
        W = p[off:off + width * n_in].reshape(n_in, width).T; off += width * n_in
        b = p[off:off + width]; off += width
        a, n_in = np.tanh(W @ a + b), width
    W = p[off:off + width]; b = p[off + width]
    return np.log1p(np.exp(W @ a + b))


def test_embedding_is_exact():
    rng = np.random.default_rng(0)
    net1 = cu.chain(4, 2, "tanh", input_dims=1)
    assert net1.n_params == 33
    p = rng.standard_normal(33)
    q = embed_ude_parameters(p, 4)
    assert q.shape == (37,) and np.all(q[4:8] == 0)
    for dG, beta in ((0.0, 1.0), (3.7, 0.2), (-0.9, 50.0)):
        assert _mlp(q, [dG, beta], 2) == _mlp(p, [dG], 1)
    g = rng.standard_normal((3, 37))
    assert np.array_equal(extract_ude_gradient(g, 4), np.delete(g, [4, 5, 6, 7], axis=1))
    assert np.array_equal(embed_ude_parameters(np.stack([p, p]), 4)[1], q)


def test_model_constructor_and_oracle_loss(fx):
    models, t, c, nn, betas = train57(fx)
    m = models[0]
    net1 = cu.chain(4, 2, "tanh", input_dims=1)
    ude = cu.CPeptideUDEModel(m.glucose_data, m.glucose_timepoints, m.age, net1, m.cpeptide_data, m.t2dm)
    assert (ude.k0, ude.k1, ude.k2, ude.c0) == (m.k0, m.k1, m.k2, m.c0) and ude.chain.input_dims == 2 and ude.ude_chain is net1
    with pytest.raises(ValueError):
        cu.CPeptideUDEModel(m.glucose_data, m.glucose_timepoints, m.age, cu.chain(4, 2, "tanh"), m.cpeptide_data, m.t2dm)
    p = np.random.default_rng(1).standard_normal(33) * 0.5
    pop = OraclePopulationAdapter([ude], t, c[:1])
    l0 = pop.loss(embed_ude_parameters(p, 4), np.array([[0.0]]))
    l1 = pop.loss(embed_ude_parameters(p, 4), np.array([[1.7]]))           # the dummy conditional parameter is inert
    assert l0[0] == l1[0]
    _, gn, gc = pop.loss_grad(embed_ude_parameters(p, 4), np.array([[0.0]]))
    assert gc[0, 0] == 0.0                     # d/d cond vanishes with a zero beta column (its weight gradient is discarded)


def test_train_ude_on_a_test_double(fx):
    models, t, c, nn, betas = train57(fx)
    m = models[3]
    ude = cu.CPeptideUDEModel(m.glucose_data, m.glucose_timepoints, m.age, cu.chain(4, 2, "tanh", input_dims=1), m.cpeptide_data, m.t2dm)
    pop = OraclePopulationAdapter([ude], t, c[3:4])
    from conditional_ude_b200.estimation import train_ude
    sols = train_ude(ude, t, c[3], np.random.default_rng(2), initial_guesses=40, selected_initials=3,
                     number_of_iterations_adam=15, number_of_iterations_lbfgs=10, population=pop)
    assert len(sols) == 3 and all(s.u.shape == (33,) for s in sols)
    p0 = np.stack(cu.initial_parameters(ude.ude_chain, 40, rng=np.random.default_rng(2)))
    screening = np.sort(pop.loss(embed_ude_parameters(p0, 4), np.zeros((40, 1))))
    assert min(s.objective for s in sols) < screening[0]


@pytest.mark.gpu
def test_ude_model_on_the_device(fx):
    models, t, c, nn, betas = train57(fx)
    m = models[0]
    ude = cu.CPeptideUDEModel(m.glucose_data, m.glucose_timepoints, m.age, cu.chain(4, 2, "tanh", input_dims=1), m.cpeptide_data, m.t2dm)
    p = np.random.default_rng(1).standard_normal(33) * 0.5
    # (tolerances this tight because the *difference quotient* needs them: the adaptive step sequence changes with the
    #  parameters, and at reltol 1e-10 that noise is 1e-3 of the quotient; the analytic gradient is the same at both)
    tight = cu.SolverOptions(abstol=1e-14, reltol=1e-13)
    l = cu.loss(p, (ude, t, c[0]), opts=tight)
    l2, g = cu.loss_and_gradient(p, (ude, t, c[0]), opts=tight)
    assert l == l2 and g.shape == (33,)
    for k in (0, 5, 20, 32):                                   # central differences at tight tolerance
        e = np.zeros(33); e[k] = 1e-5
        fd = (cu.loss(p + e, (ude, t, c[0]), opts=tight) - cu.loss(p - e, (ude, t, c[0]), opts=tight)) / 2e-5
        assert abs(fd - g[k]) < 1e-4 * max(1.0, np.abs(g).max())
    sols = cu.train(ude, t, c[0], np.random.default_rng(3), initial_guesses=2000, selected_initials=4,
                    number_of_iterations_adam=100, number_of_iterations_lbfgs=50)
    assert len(sols) == 4 and min(s.objective for s in sols) < 0.5
