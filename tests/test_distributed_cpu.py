"""CPU tier: the N > 1 host logic (shard the individuals, all-reduce the per-start sums) on two gloo
ranks, with the oracle standing in for the per-rank kernel.  The result must equal the single-rank
population loss and gradient."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import conditional_ude_b200 as cu
    from conditional_ude_b200.distributed import shard_bounds, sharded_loss_grad
    from oracle import oracle
    from helpers import train57
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
    models, t, c, nn, betas = train57(fx)
    n = len(models)
    lo, hi = shard_bounds(n, world, rank)
    op = oracle.OraclePopulation(cu.pack_models(models[lo:hi], t, c[lo:hi]))
    rng = np.random.default_rng(0)                                      # same starts on every rank
    neural = nn[None] + 0.05 * rng.standard_normal((3, 37))
    cond = np.tile(betas, (3, 1)) + 0.1 * rng.standard_normal((3, n))
    if True:
        cond[2, 4] = np.nan                                             # a failed trajectory on rank 0's shard

    def local(neural_, cond_):
        r = op.eval(neural_, cond_, grad_mode=0, n_threads=1)
        sums = np.concatenate([r["sse"].sum(axis=1, keepdims=True), r["g_neural"].sum(axis=1)], axis=1)
        return sums, r["g_cond"]

    loss, g, gc = sharded_loss_grad(local, neural, cond[:, lo:hi], n)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=loss, g=g, gc=gc, lo=lo, hi=hi, neural=neural, cond=cond)
    dist.destroy_process_group()


def test_two_rank_population_step_equals_single_rank(tmp_path, fx):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conditional_ude_b200 as cu
    from oracle import oracle
    from helpers import train57
    models, t, c, nn, betas = train57(fx)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    assert (int(r[0]["lo"]), int(r[0]["hi"]), int(r[1]["lo"]), int(r[1]["hi"])) == (0, 28, 28, 57)
    ref = oracle.OraclePopulation(cu.pack_models(models, t, c)).population_loss(r[0]["neural"], r[0]["cond"], with_grad=True)
    for k in range(world):
        assert np.array_equal(np.isinf(r[k]["loss"]), [False, False, True])
        assert np.allclose(r[k]["loss"][:2], ref["loss"][:2], rtol=1e-13)
        assert np.allclose(r[k]["g"], ref["g_neural"], rtol=1e-11, atol=1e-14)
        assert np.all(r[k]["g"][2] == 0)
    gc = np.concatenate([r[0]["gc"], r[1]["gc"]], axis=1)
    assert np.allclose(gc, ref["g_cond"], rtol=1e-13, atol=0)
    assert np.array_equal(r[0]["loss"], r[1]["loss"]) and np.array_equal(r[0]["g"], r[1]["g"])


def _worker_starts(rank, world, port, out_dir):
    """Start-sharded configurations (multi-start training, profiles): no communication on the data path."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import conditional_ude_b200 as cu
    from helpers import train57, OraclePopulationAdapter
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
    models, t, c, nn, betas = train57(fx)
    pop = OraclePopulationAdapter(models[:6], t, c[:6])
    sols = cu.train(pop, t, c[:6], np.random.default_rng(5), initial_guesses=21, selected_initials=3,
                    number_of_iterations_adam=4, number_of_iterations_lbfgs=3, distributed=True)
    nll, nmin, grid = cu.likelihood_profile_population(betas[:6], nn, pop, betas[:6] - 1.0, betas[:6] + 1.0, 0.1, steps=9,
                                                       distributed=True)
    np.savez(os.path.join(out_dir, f"s{rank}.npz"), obj=[s.objective for s in sols], neural=[s.u.neural for s in sols],
             cond=[s.u.conditional for s in sols], nll=nll, nmin=nmin, grid=grid, traj=pop.traj)
    dist.destroy_process_group()


def test_two_rank_start_sharded_train_and_profiles_equal_single_rank(tmp_path, fx):
    world = 2
    mp.spawn(_worker_starts, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conditional_ude_b200 as cu
    from helpers import train57, OraclePopulationAdapter
    models, t, c, nn, betas = train57(fx)
    pop = OraclePopulationAdapter(models[:6], t, c[:6])
    ref = cu.train(pop, t, c[:6], np.random.default_rng(5), initial_guesses=21, selected_initials=3,
                   number_of_iterations_adam=4, number_of_iterations_lbfgs=3)
    rnll, rmin, rgrid = cu.likelihood_profile_population(betas[:6], nn, pop, betas[:6] - 1.0, betas[:6] + 1.0, 0.1, steps=9)
    r = [np.load(tmp_path / f"s{k}.npz") for k in range(world)]
    assert len(ref) == 3
    for k in range(world):
        # every start is optimised independently of the others in its batch, so the split changes nothing
        assert np.array_equal(r[k]["obj"], [s.objective for s in ref])
        assert np.array_equal(r[k]["neural"], [s.u.neural for s in ref]) and np.array_equal(r[k]["cond"], [s.u.conditional for s in ref])
        assert np.array_equal(r[k]["nll"], rnll) and np.array_equal(r[k]["nmin"], rmin) and np.array_equal(r[k]["grid"], rgrid)
    # the work really was split: each rank evaluated about half of the trajectories
    assert max(int(r[0]["traj"]), int(r[1]["traj"])) < 0.75 * pop.traj
