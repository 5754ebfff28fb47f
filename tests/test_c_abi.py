"""The C ABI from plain C: tests/c_abi/example.c is compiled with gcc against include/cude_b200.h and linked to
libcude_b200.so — no Python, no ctypes on the call path.  CPU tier: it must report "no device" (no CPU fallback);
GPU tier: it must run and agree with the Python mirror on the same inputs."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "conditional_ude_b200", "csrc")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_example")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "c_abi", "example.c"),
                           "-o", exe, "-L", LIBDIR, "-lcude_b200", "-lm", "-Wl,-rpath," + LIBDIR])
    return exe


def test_c_program_builds_and_reports_no_device_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu-marked test")
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_c_program_runs_and_matches_the_python_mirror(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    loss_c = float(r.stdout.split()[1])
    import conditional_ude_b200 as cu
    net = cu.chain(4, 2, "tanh")
    t = np.array([0.0, 30, 60, 90, 120])
    g = np.array([[5.0, 8.5, 9.0, 7.0, 5.5], [6.0, 11.0, 13.5, 12.0, 9.0]])
    y = np.array([[0.5, 1.4, 1.9, 1.7, 1.2], [0.7, 1.2, 1.8, 2.0, 1.9]])
    models = [cu.CPeptideConditionalUDEModel(g[i], t, [35.0, 62.0][i], net, y[i], [False, True][i]) for i in range(2)]
    neural = 0.3 * np.sin(1.0 + 0.7 * np.arange(37))
    l = cu.Population(models, t, y, ctx=cu.Context(0)).loss(neural, np.array([[-1.0, -0.5]]))
    assert abs(l[0] - loss_c) <= 1e-11 * l[0]            # 12 printed digits
