"""Host-side mirror of the reference's model constructors for the cUDE path.

Same names, argument order and meaning as the Julia reference (Julia is not installed in this
image, so the host side above the C ABI is Python; the Julia shim with identical entry points is
in julia/CUDEB200.jl):

  van_cauter_parameters                      src/c-peptide-models.jl:30-42
  chain / softplus                           src/neural-network.jl:13-15, 42-58, 85-87, 105-107
  CPeptideConditionalUDEModel                src/c-peptide-models.jl:170-194, src/types.jl:16-19
  CPeptideConditionalCovariateUDEModel       src/c-peptide-models.jl:196-220

A model here is plain data (kinetic constants, glucose knots, time span): the ODE right-hand side
itself lives in the CUDA kernels.  Nothing in this module computes a trajectory.
"""
import math
from dataclasses import dataclass, field

import numpy as np

LN2 = math.log(2.0)


def softplus(x):
    """softplus(x) = log(1 + exp(x)) (src/neural-network.jl:13-15) — marker for the output activation."""
    return math.log(1.0 + math.exp(x))


def van_cauter_parameters(age, t2dm):
    """(k0, k1, k2) of the van Cauter c-peptide kinetics, src/c-peptide-models.jl:30-42."""
    short_half_life = 4.52 if t2dm else 4.95
    fraction = 0.78 if t2dm else 0.76
    long_half_life = 0.14 * age + 29.2
    k1 = fraction * (LN2 / long_half_life) + (1 - fraction) * (LN2 / short_half_life)
    k0 = (LN2 / short_half_life) * (LN2 / long_half_life) / k1
    k2 = (LN2 / short_half_life) + (LN2 / long_half_life) - k0 - k1
    return k0, k1, k2


def _is_tanh(f):
    return f in ("tanh", math.tanh, np.tanh) or getattr(f, "__name__", None) == "tanh"


@dataclass(frozen=True)
class Chain:
    """Shape of a SimpleChains MLP: input_dims -> widths (tanh) -> 1 (softplus).

    Parameter layout (SimpleChains TurboDense{true}): per layer W[out x in] column-major then
    bias[out]; `chain(4, 2, tanh)` has 37 parameters, input_dims=3 gives 41.
    """
    input_dims: int
    width: int
    depth: int

    @property
    def n_params(self):
        p, n_in = 0, self.input_dims
        for _ in range(self.depth):
            p += self.width * (n_in + 1)
            n_in = self.width
        return p + n_in + 1

    def init_params(self, rng):
        """Glorot-normal weights, zero biases (SimpleChains.init_params default; the Julia RNG
        stream itself cannot be reproduced here)."""
        out, n_in = [], self.input_dims
        for n_out in [self.width] * self.depth + [1]:
            sigma = math.sqrt(2.0 / (n_in + n_out))
            out.append(rng.normal(0.0, sigma, size=n_out * n_in))
            out.append(np.zeros(n_out))
            n_in = n_out
        return np.concatenate(out)

    def init_params_batch(self, rng, n):
        """n Glorot-normal weight vectors at once, [n x n_params] (one normal draw per layer for all n networks; the multi-start
        `train` draws 25 000 — a Python loop over init_params took 0.14 s of its 0.39 s)."""
        out, n_in = [], self.input_dims
        for n_out in [self.width] * self.depth + [1]:
            sigma = math.sqrt(2.0 / (n_in + n_out))
            out.append(rng.normal(0.0, sigma, size=(n, n_out * n_in)))
            out.append(np.zeros((n, n_out)))
            n_in = n_out
        return np.concatenate(out, axis=1)


def chain(*args, input_dims=2, output_dims=1, output_activation=softplus):
    """chain(width, depth, act) | chain(widths, act) | chain(widths, acts); src/neural-network.jl:42,85,105.

    The device path supports equal hidden widths with tanh and a single softplus output (every
    network the reference builds for the cUDE: chain(4, 2, tanh), chain(4, 2, tanh; input_dims=3)).
    """
    if len(args) == 3:
        width, depth, act = args
        widths, acts = [int(width)] * int(depth), [act] * int(depth)
    elif len(args) == 2:
        widths = list(args[0])
        acts = list(args[1]) if isinstance(args[1], (list, tuple)) else [args[1]] * len(widths)
    else:
        raise TypeError("chain(width, depth, act) or chain(widths, act[s])")
    if len(widths) == 0:
        raise ValueError("Input widths must be non-empty.")
    if len(widths) != len(acts):
        raise ValueError("The number of widths must match the number of activation functions.")
    if output_dims != 1 or output_activation is not softplus:
        raise NotImplementedError("device path: single softplus output only")
    if len(set(widths)) != 1 or not all(_is_tanh(a) for a in acts):
        raise NotImplementedError("device path: equal hidden widths with tanh only")
    return Chain(int(input_dims), int(widths[0]), len(widths))


@dataclass
class CPeptideConditionalUDEModel:
    """Data image of the reference's CPeptideConditionalUDEModel (src/types.jl:16-19):
    `problem` = ODEProblem(kinetics + conditional production, u0, tspan) and `chain`.

    CPeptideConditionalUDEModel(glucose_data, glucose_timepoints, age, network, cpeptide_data, t2dm)
    follows src/c-peptide-models.jl:170-194: c0 = cpeptide_data[1]; (k0,k1,k2) = van Cauter;
    glucose = LinearInterpolation(glucose_data, glucose_timepoints); u0 = [c0, k2/k1*c0];
    tspan = (glucose_timepoints[1], glucose_timepoints[end]); t0 = glucose_timepoints[1].
    """
    glucose_data: np.ndarray
    glucose_timepoints: np.ndarray
    age: float
    chain: Chain
    cpeptide_data: np.ndarray
    t2dm: bool
    covariate: float = None          # third network input (age) for the covariate variant
    k0: float = field(init=False)
    k1: float = field(init=False)
    k2: float = field(init=False)
    c0: float = field(init=False)

    def __post_init__(self):
        self.glucose_data = np.ascontiguousarray(self.glucose_data, dtype=np.float64)
        self.glucose_timepoints = np.ascontiguousarray(self.glucose_timepoints, dtype=np.float64)
        self.cpeptide_data = np.ascontiguousarray(self.cpeptide_data, dtype=np.float64)
        if self.glucose_data.ndim != 1 or self.glucose_data.shape != self.glucose_timepoints.shape:
            raise ValueError("glucose_data and glucose_timepoints must be vectors of equal length")
        if self.glucose_data.size < 2:
            raise ValueError("need at least two glucose knots")
        if np.any(np.diff(self.glucose_timepoints) <= 0):
            raise ValueError("glucose_timepoints must be strictly increasing")
        if not isinstance(self.chain, Chain):
            raise TypeError("network must come from chain(...)")
        need = 3 if self.covariate is not None else 2
        if self.chain.input_dims != need:
            raise ValueError(f"network has input_dims={self.chain.input_dims}, model needs {need}")
        self.c0 = float(self.cpeptide_data[0])
        self.k0, self.k1, self.k2 = van_cauter_parameters(float(self.age), bool(self.t2dm))

    @property
    def u0(self):
        return np.array([self.c0, (self.k2 / self.k1) * self.c0])

    @property
    def tspan(self):
        return float(self.glucose_timepoints[0]), float(self.glucose_timepoints[-1])


def CPeptideConditionalCovariateUDEModel(glucose_data, glucose_timepoints, age, network, cpeptide_data, t2dm):
    """src/c-peptide-models.jl:196-220 — returns a CPeptideConditionalUDEModel (:219) whose network
    takes [dG; beta; age]."""
    return CPeptideConditionalUDEModel(glucose_data, glucose_timepoints, age, network, cpeptide_data, t2dm,
                                       covariate=float(age))


def pack_models(models, timepoints, cpeptide_data):
    """Flatten a vector of models + (timepoints, cpeptide_data) of the loss tuple
    (src/parameter-estimation.jl:126) into the row-major arrays cude_population_create takes.

    timepoints: vector shared by all individuals, or a list of per-individual vectors.
    cpeptide_data: matrix [n_individuals x n_timepoints] or list of per-individual vectors.
    """
    n = len(models)
    if n == 0:
        raise ValueError("empty model vector")
    per_ind_t = isinstance(timepoints, (list, tuple)) and len(timepoints) == n and np.ndim(timepoints[0]) == 1
    obs_ts = [np.asarray(timepoints[i] if per_ind_t else timepoints, dtype=np.float64) for i in range(n)]
    obs_ys = [np.asarray(cpeptide_data[i], dtype=np.float64) for i in range(n)]
    max_knots = max(m.glucose_timepoints.size for m in models)
    max_obs = max(t.size for t in obs_ts)
    out = dict(
        n_ind=n, max_knots=max_knots, max_obs=max_obs,
        n_knots=np.zeros(n, np.int32), knot_t=np.zeros((n, max_knots)), knot_g=np.zeros((n, max_knots)),
        n_obs=np.zeros(n, np.int32), obs_t=np.zeros((n, max_obs)), obs_y=np.zeros((n, max_obs)),
        kin=np.zeros((n, 4)), cov=None)
    has_cov = [m.covariate is not None for m in models]
    if any(has_cov) and not all(has_cov):
        raise ValueError("mixing covariate and plain cUDE models")
    if all(has_cov):
        out["cov"] = np.array([m.covariate for m in models], dtype=np.float64)
    ch = models[0].chain
    for i, m in enumerate(models):
        if m.chain != ch:
            raise ValueError("all models must share one network shape")
        if obs_ts[i].shape != obs_ys[i].shape:
            raise ValueError("timepoints and cpeptide_data length mismatch")
        nk, no = m.glucose_timepoints.size, obs_ts[i].size
        out["n_knots"][i] = nk
        out["knot_t"][i, :nk] = m.glucose_timepoints
        out["knot_g"][i, :nk] = m.glucose_data
        out["knot_t"][i, nk:] = m.glucose_timepoints[-1]
        out["knot_g"][i, nk:] = m.glucose_data[-1]
        out["n_obs"][i] = no
        out["obs_t"][i, :no] = obs_ts[i]
        out["obs_y"][i, :no] = obs_ys[i]
        if no:
            out["obs_t"][i, no:] = obs_ts[i][-1]
        out["kin"][i] = (m.k0, m.k1, m.k2, m.c0)
    out["chain"] = ch
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Non-conditional UDE (src/c-peptide-models.jl:144-168, src/types.jl:11-14): production = NN([dG]) - NN([0]) with a
# 1-input network, e.g. chain(4, 2, tanh; input_dims=1) (c-peptide/01-non-conditional.jl:21-23).  The device kernels take
# the conditional form NN([dG; beta]); a 1-input network is *exactly* the 2-input network whose beta column is zero
# (fma(0, beta, b) = b), so the UDE model runs through the same kernels by embedding its parameter vector.
def embed_ude_parameters(p, width):
    """[W1[:,0], b1, rest] of a 1-input chain -> [W1[:,0], 0 (beta column), b1, rest] of the 2-input chain."""
    p = np.asarray(p, dtype=np.float64)
    return np.concatenate([p[..., :width], np.zeros(p.shape[:-1] + (width,)), p[..., width:]], axis=-1)


def extract_ude_gradient(g, width):
    """Inverse of `embed_ude_parameters` for gradients: drop the beta column."""
    g = np.asarray(g, dtype=np.float64)
    return np.concatenate([g[..., :width], g[..., 2 * width:]], axis=-1)


class CPeptideUDEModel(CPeptideConditionalUDEModel):
    """CPeptideUDEModel(glucose_data, glucose_timepoints, age, network, cpeptide_data, t2dm) with a 1-input network
    (src/c-peptide-models.jl:144-168).  `chain` is the user's 1-input chain; `device_chain` the embedded 2-input one."""

    def __init__(self, glucose_data, glucose_timepoints, age, network, cpeptide_data, t2dm):
        if not isinstance(network, Chain) or network.input_dims != 1:
            raise ValueError("CPeptideUDEModel needs a network with input_dims=1")
        super().__init__(glucose_data, glucose_timepoints, age, Chain(2, network.width, network.depth), cpeptide_data, t2dm)
        self.device_chain = self.chain
        self.ude_chain = network
