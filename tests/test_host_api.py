"""CPU tier: host-side mirror of the reference interface and the C ABI surface (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import conditional_ude_b200 as cu
from conditional_ude_b200 import _lib
from helpers import ohashi_models, fujita_models, mixed_population

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cude_b200.h")).read()
    declared = set(re.findall(r"\b(cude_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.cude_abi_version() == 3
    o = _lib.cude_opts()
    lib.cude_default_opts(C.byref(o))
    assert (o.abstol, o.reltol, o.maxiters) == (1e-6, 1e-3, 1000000)    # OrdinaryDiffEq defaults (maxiters: __init, adaptive)
    assert lib.cude_net_nparams(C.byref(_lib.cude_net(2, 2, 4))) == 37
    assert lib.cude_net_nparams(C.byref(_lib.cude_net(3, 2, 4))) == 41
    k = [C.c_double() for _ in range(3)]
    lib.cude_van_cauter_parameters(40.0, 0, *[C.byref(x) for x in k])
    assert np.allclose([x.value for x in k], cu.van_cauter_parameters(40.0, False), rtol=1e-15)


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly (this test tier runs on a CPU-only box;
    on a GPU box it is skipped)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(_lib.CudeError) as ei:
        cu.Context(0)
    assert ei.value.code == _lib.CUDE_ENODEVICE and "no CPU fallback" in str(ei.value)
    models, t, c = ohashi_models({k: v for k, v in np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")).items()})
    with pytest.raises(_lib.CudeError):
        cu.loss(-1.0, (models[0], t, c[0], np.zeros(37)))


def test_multi_gpu_entry_points_fail_loudly_without_a_device():
    """The multi-GPU half of the ABI has no CPU fallback either: cude_mctx_create reports CUDE_ENODEVICE, and the NCCL
    loader (run-time dlopen) finds the library of the image."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    lib = _lib.load()
    assert lib.cude_device_count() == 0
    with pytest.raises(_lib.CudeError) as ei:
        cu.MultiContext(2)
    assert ei.value.code == _lib.CUDE_ENODEVICE and "no CPU fallback" in str(ei.value)
    assert lib.cude_nccl_version() >= 20000            # libnccl.so.2 of the image, loaded on demand
    assert lib.cude_comm_size(None) == _lib.CUDE_EINVAL


def test_default_device_follows_the_local_rank():
    """One process per GPU: a rank that names no device takes the one of its LOCAL_RANK (torchrun does not narrow
    CUDA_VISIBLE_DEVICES), never silently device 0 for everybody; two ranks on one device is an error."""
    from conditional_ude_b200.population import pick_device
    assert [pick_device(8, r) for r in range(8)] == list(range(8))          # distinct devices
    assert pick_device(8, None, None) == 0 and pick_device(8, None, 5) == 5   # no launcher: torch's current device, else 0
    assert pick_device(1, 3, None, one_visible=True) == 0                     # launcher narrowed the visibility per rank
    with pytest.raises(RuntimeError):
        pick_device(2, 3)                                                     # ranks would share a GPU


def test_product_package_does_not_import_the_oracle():
    import sys
    import importlib
    for m in [m for m in sys.modules if m.startswith("oracle")]:
        del sys.modules[m]
    importlib.reload(cu)
    assert not any(m.startswith("oracle") for m in sys.modules)
    src = "".join(open(os.path.join(ROOT, "conditional_ude_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "conditional_ude_b200")) if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src


def test_chain_constructor_mirrors_reference():
    net = cu.chain(4, 2, "tanh")
    assert (net.input_dims, net.width, net.depth, net.n_params) == (2, 4, 2, 37)
    assert cu.chain([4, 4], np.tanh) == net and cu.chain([4, 4], ["tanh", "tanh"]) == net
    assert cu.chain(4, 2, "tanh", input_dims=3).n_params == 41
    with pytest.raises(ValueError):
        cu.chain([], "tanh")                                   # neural-network.jl:44-46
    with pytest.raises(ValueError):
        cu.chain([4, 4], ["tanh"])                             # neural-network.jl:48-50
    with pytest.raises(NotImplementedError):
        cu.chain([4, 3], "tanh")
    p = net.init_params(np.random.default_rng(0))
    assert p.shape == (37,) and np.all(p[8:12] == 0) and np.all(p[28:32] == 0) and p[36] == 0   # zero biases


def test_model_constructor(fx):
    models, t, c = ohashi_models(fx, "train")
    m = models[0]
    k0, k1, k2 = cu.van_cauter_parameters(float(fx["ohashi_train_ages"][0]), bool(fx["ohashi_train_t2dm"][0]))
    assert (m.k0, m.k1, m.k2) == (k0, k1, k2) and m.c0 == c[0, 0]
    assert np.allclose(m.u0, [c[0, 0], k2 / k1 * c[0, 0]]) and m.tspan == (0.0, 120.0)
    fm, ft, fc = fujita_models(fx)
    assert fm[0].tspan == (-10.0, 240.0)
    with pytest.raises(ValueError):
        cu.CPeptideConditionalUDEModel(fx["ohashi_train_glucose"][0], t[::-1].copy(), 30.0, cu.chain(4, 2, "tanh"), c[0], False)
    with pytest.raises(ValueError):
        cu.CPeptideConditionalUDEModel(fx["ohashi_train_glucose"][0], t, 30.0, cu.chain(4, 2, "tanh", input_dims=3), c[0], False)
    cm = cu.CPeptideConditionalCovariateUDEModel(fx["ohashi_train_glucose"][0], t, 34.0, cu.chain(4, 2, "tanh", input_dims=3), c[0], False)
    assert isinstance(cm, cu.CPeptideConditionalUDEModel) and cm.covariate == 34.0     # c-peptide-models.jl:219


def test_pack_models_ragged(fx):
    models, ts, ys = mixed_population(fx)
    pk = cu.pack_models(models, ts, ys)
    assert pk["n_ind"] == 137 and pk["max_knots"] == 14 and pk["max_obs"] == 14
    assert set(pk["n_knots"]) == {5, 14} and pk["kin"].shape == (137, 4)
    assert pk["knot_t"][0, 4] == 120 and pk["knot_t"][0, 13] == 120            # padded with the last knot
    assert pk["knot_t"][136, 0] == -10 and pk["obs_t"][136, 13] == 240
    with pytest.raises(ValueError):
        cu.pack_models(models[:3], ts[:3], [ys[0], ys[1], ys[2][:3]])


def test_find_confidence_intervals():
    x = np.linspace(-3, 3, 601)
    nll = 10.0 * x ** 2
    lo, hi = cu.find_confidence_intervals(nll, 0.0, x, target="cantelli95")   # threshold 7.16
    assert abs(lo + np.sqrt(0.716)) < 0.01 and abs(hi - np.sqrt(0.716)) < 0.01
    lo, hi = cu.find_confidence_intervals(nll, 0.0, x, target="cantelli90")   # 5.24
    assert abs(hi - np.sqrt(0.524)) < 0.01
    lo, hi = cu.find_confidence_intervals(nll, 0.0, x, target="raue95")       # chi2_1(0.95) = 3.8415
    assert abs(hi - np.sqrt(0.38415)) < 0.01
    lo, hi = cu.find_confidence_intervals(0.1 * x ** 2, 0.0, x)               # never crosses: open interval
    assert lo == -np.inf and hi == np.inf


def test_likelihood_profile_generic_method():
    """likelihood_profile(beta, loss_function, args, lb, ub, sigma; steps) — src/likelihood-profiles.jl:19-32."""
    nll, nmin, grid = cu.likelihood_profile(0.3, lambda b, a: (b - a) ** 2, 1.0, -1.0, 2.0, 0.5, steps=4)
    assert np.allclose(grid, [-1, 0, 1, 2]) and np.allclose(nll, 2 * (grid - 1) ** 2) and np.isclose(nmin, 2 * 0.49)


def test_component_vector():
    th = cu.ComponentVector(neural=np.arange(3.0), conditional=[1.0])
    assert th.neural[2] == 2.0 and th["conditional"] == [1.0]
    th.sigma = 0.5
    assert th["sigma"] == 0.5


def test_jld2_reader_on_reference_artifacts(fx):
    """The structural JLD2 reader (conditional_ude_b200/jld2.py) against the reference's own files; the files are
    only present in the build container (skipped on the GPU box), the derived arrays are the golden fixtures."""
    import os
    from conditional_ude_b200 import jld2
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present")
    c = jld2.load(os.path.join(ref, "source_data/cude_neural_parameters.jld2"))
    assert c["best_model_index"] == 14 and (c["width"], c["depth"]) == (4, 2)
    assert np.array_equal(np.stack(c["parameters"]), fx["cude_neural"])
    assert np.array_equal(np.stack(c["betas"]), fx["cude_betas"])
    d = jld2.load(os.path.join(ref, "suppression/results/lambda=0.01.jld2"))
    assert d["group_data"].shape == (3, 8, 37) and d["λ"] == 0.01 and len(d["neural_parameters"]) == 25
    assert np.all(d["group_data"][1:, 0, :] == 0)              # u0 = [x, 0, 0] (multiplicative noise keeps the zeros)
    sup = dict(np.load(os.path.join(ROOT, "tests", "golden", "suppression_fixtures.npz")))
    assert np.array_equal(sup["group_data"], d["group_data"]) and np.array_equal(sup["neural_0p01"], np.stack(d["neural_parameters"]))
    with pytest.raises(ValueError):
        jld2.JLD2File(os.path.join(ref, "data/ohashi.jld2"))["train"]      # NamedTuple: outside the subset


def test_jld2_checksums_and_writer_round_trip(tmp_path, fx):
    """lookup3 is pinned by the reference's own files (superblock + every object header of a JLD2-written file carries
    one); the writer's files pass the same verification and read back bit-identically."""
    import os
    from conditional_ude_b200 import jld2
    assert jld2.lookup3(b"") == 0xDEADBEEF and jld2.lookup3(b"Four score and seven years ago") == 0x17770551   # lookup3.c driver5
    ref = "/root/reference"
    if os.path.isdir(ref):
        for name, n_headers in (("source_data/cude_neural_parameters.jld2", 8), ("suppression/results/lambda=1.0.jld2", 17)):
            assert jld2.JLD2File(os.path.join(ref, name)).verify() == n_headers
    rng = np.random.default_rng(0)
    data = {"a": 3, "b": 2.5, "M": rng.standard_normal((3, 5)), "T": rng.standard_normal((2, 3, 4)), "iv": np.arange(7),
            "λ": 0.1, "empty": np.zeros(0)}
    path = str(tmp_path / "w.jld2")
    jld2.save(path, data)
    f = jld2.JLD2File(path)
    assert f.verify() == len(data) + 1 and f.keys() == list(data)
    for k, v in data.items():
        assert np.array_equal(f[k], v) and type(f[k]) is type(v)
    raw = bytearray(open(path, "rb").read())
    raw[512 + 60] ^= 1                                           # one flipped bit inside the first object header
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        jld2.JLD2File(path).verify()
    with pytest.raises(TypeError):
        jld2.save(path, {"s": "text"})
    # the result file of 02-conditional.jl:44-50, written from the stored arrays and read back
    jld2.save_neural_parameters(path, list(fx["cude_neural"]), list(fx["cude_betas"]), best_model_index=14)
    r = jld2.load_neural_parameters(path)
    assert (r["width"], r["depth"], r["best_model_index"]) == (4, 2, 14)
    assert np.array_equal(r["parameters"], fx["cude_neural"]) and np.array_equal(r["betas"], fx["cude_betas"])
    if os.path.isdir(ref):
        q = jld2.load_neural_parameters(os.path.join(ref, "source_data/cude_neural_parameters.jld2"))
        assert np.array_equal(q["parameters"], r["parameters"]) and np.array_equal(q["betas"], r["betas"])
        # dataset messages byte-identical to what JLD2 itself wrote for the same content (addresses aside)
        g = jld2.JLD2File(os.path.join(ref, "source_data/ude_neural_parameters.jld2"))
        jld2.save(path, {k: g[k] for k in g.keys()})
        h = jld2.JLD2File(path)
        for k in g.keys():
            ma, mb = list(h._messages(h._links(h.root)[k])), list(g._messages(g._links(g.root)[k]))
            assert [(t, bytes(b)) for t, b in ma if t != 8] == [(t, bytes(b)) for t, b in mb if t != 8]
