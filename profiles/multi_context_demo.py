#!/usr/bin/env python
"""ONE process driving all GPUs of the box through the C ABI's multi-GPU context (cude_mctx_*): what a single Julia session
would do.  BASELINE configs 2-4 with the starts / grid points split over the devices (no communication) and config 5 with the
individuals split (NCCL all-reduce of the per-start sums inside the library), host matrices in, host matrices out.
  python profiles/multi_context_demo.py [n_gpus] [individuals_config5]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import conditional_ude_b200 as cu

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n5 = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
nn = bench.stored_network()
mctx = cu.MultiContext(ng)
one = cu.Context(0)
out = {"n_gpus": mctx.n_gpus}


def timeit(fn, reps=3):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), r


def both(name, build, call, ntraj, shard):
    p1 = build(lambda pk: cu.Population(packed=pk, ctx=one))
    pm = build(lambda pk: cu.MultiPopulation(packed=pk, mctx=mctx, shard=shard))
    t1, r1 = timeit(lambda: call(p1))
    tm, rm = timeit(lambda: call(pm))
    same = all(np.allclose(a, b, rtol=1e-12, atol=0, equal_nan=True) for a, b in zip(r1, rm) if a is not None)
    out[name] = {"trajectories": int(ntraj), "partition": shard, "one_gpu_s": t1, "all_gpus_s": tm, "speedup": t1 / tm,
                 "evals_per_s_all_gpus": ntraj / tm, "results_equal_one_gpu": bool(same), "kernel_ms_slowest_device": mctx.stats()["kernel_ms"]}


# config 2: beta-only gradient, 137 individuals x 8000 starts (1000 per GPU)
m, t, y = bench._fixture_models(fx, cu, ["train", "test", "fujita"])
pk137 = cu.pack_models(m, t, y)
S2 = 1000 * max(1, mctx.n_gpus)
cond2 = np.random.default_rng(0).uniform(-4.0, 1.0, size=(S2, len(m)))
both("config2_beta_only_137xS", lambda mk: mk(pk137), lambda p: p.loss_grad(nn, cond2, neural_grad=False, mean=False)[:1] + (p.loss_grad(nn, cond2, neural_grad=False, mean=False)[2],),
     cond2.size, "starts")
# config 3: screening of 25 000 initial guesses x 57 individuals
idx = fx["train_split_idx"]
m57, t57, y57 = bench._fixture_models(fx, cu, ["train"])
pk57 = cu.pack_models([m57[i] for i in idx], [t57[i] for i in idx], [y57[i] for i in idx])
rng = np.random.default_rng(1)
neural3 = np.stack(cu.initial_parameters(pk57["chain"], 25_000, rng=rng))
cond3 = cu.initial_parameters(57, -2.0, 0.0, 25_000, rng).T
both("config3_screening_57x25000", lambda mk: mk(pk57), lambda p: (p.loss(neural3, cond3),), cond3.size, "starts")
# config 4: profiles, 117 individuals x 10 000 grid points
m, t, y = bench._fixture_models(fx, cu, ["train", "test"])
pk117 = cu.pack_models(m, t, y)
grid = np.linspace(np.full(117, -11.0), np.full(117, 14.0), 10000)
both("config4_profiles_117x10000", lambda mk: mk(pk117), lambda p: p.loss(nn, grid, return_sse=True), grid.size, "starts")
# config 5: population loss + gradient, individuals split, all-reduce inside the library
pk5 = bench.synthetic_population(n5, 1000, bench.simulate_gpu(one))
neural5, cond5 = bench.synthetic_starts(n5, 64, 11, 2000)
both("config5_population_%dx64_loss_grad" % n5, lambda mk: mk(pk5), lambda p: p.loss_grad(neural5, cond5), n5 * 64, "individuals")
print(json.dumps(out))
