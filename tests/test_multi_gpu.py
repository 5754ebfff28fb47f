"""Multi-GPU behind the C ABI (include/cude_b200.h, csrc/cude_multi.inl).

  * plain-C caller (tests/c_abi/example_multi.c): a single process drives 2 GPUs through cude_mctx_* — results equal one
    GPU's (per-trajectory values bit for bit, per-start sums to 1e-13; STARTS partition bit for bit throughout);
  * process-per-GPU (tests/workers/sharded_rank.py under torchrun): library-side NCCL communicator, collective
    cude_loss_grad_sharded and the device-resident step with cude_allreduce_dev, against one GPU on the whole population.
Both need >= 2 devices and are skipped otherwise (the driver's GPU tier has one; `gpurun --gpus 2` runs them, log in
profiles/).  The single-device variants below run everywhere a GPU is: the same code paths with one rank / one device."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "conditional_ude_b200", "csrc")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_example_multi")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "c_abi", "example_multi.c"),
                           "-o", exe, "-L", LIBDIR, "-lcude_b200", "-lm", "-Wl,-rpath," + LIBDIR])
    return exe


def _n_devices():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_multi_c_program_builds_and_reports_no_device_without_gpu(tmp_path):
    if _n_devices():
        pytest.skip("a GPU is present: covered by the gpu-marked tests")
    r = subprocess.run([_build(tmp_path), "2"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_multi_context_with_one_device_equals_single_context(tmp_path):
    """cude_mctx with one device: worker thread, partition logic and host-matrix strides, no NCCL."""
    r = subprocess.run([_build(tmp_path), "1", "3000", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_two_gpus_one_process_plain_c_equals_one_gpu(tmp_path):
    if _n_devices() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([_build(tmp_path), "2", "20000", "6"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "individuals x2: sse bitwise 1" in r.stdout and "starts x2: sse bitwise 1" in r.stdout


def _torchrun(nproc, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                           "--master-addr", "127.0.0.1", "--master-port", "29617",
                           os.path.join(HERE, "workers", "sharded_rank.py")], capture_output=True, text=True, env=e, timeout=600)


@pytest.mark.gpu
def test_library_communicator_with_one_rank():
    """The process-per-GPU path on one device: NCCL loaded at run time, communicator of size 1 is a no-op all-reduce."""
    r = _torchrun(1, {"CUDE_TEST_N": "5000"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_two_ranks_library_allreduce_equals_one_gpu():
    if _n_devices() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = _torchrun(2)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["n_gpus"] == 2 and out["host_call"]["g_cond_bitwise"] and out["device_call"]["g_cond_bitwise"]


@pytest.mark.gpu
def test_multipopulation_python_mirror(fx):
    """MultiPopulation (the host mirror over cude_mctx) returns what Population returns, on however many GPUs there are."""
    import conditional_ude_b200 as cu
    from helpers import train57
    models, t, c, nn, betas = train57(fx)
    rng = np.random.default_rng(1)
    S = 7
    neural = nn[None] + 0.05 * rng.standard_normal((S, 37))
    cond = np.tile(betas, (S, 1)) + 0.2 * rng.standard_normal((S, 57))
    one = cu.Population(models, t, c, ctx=cu.Context(0))
    l0, gn0, gc0, sse0 = one.loss_grad(neural, cond, return_sse=True)
    mctx = cu.MultiContext(min(2, _n_devices()))
    for shard in ("starts", "individuals"):
        mp = cu.MultiPopulation(models, t, c, mctx=mctx, shard=shard)
        l, gn, gc, sse = mp.loss_grad(neural, cond, return_sse=True)
        assert np.array_equal(sse, sse0) and np.array_equal(gc, gc0)
        assert np.allclose(l, l0, rtol=1e-13) and np.allclose(gn, gn0, rtol=1e-11, atol=1e-14)
        assert np.allclose(mp.loss(nn, cond), one.loss(nn, cond), rtol=1e-13)          # shared network, flat indexing
        assert mctx.stats()["n_traj"] == S * 57
