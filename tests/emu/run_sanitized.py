#!/usr/bin/env python
"""Runs the kernel-source emulation tests with the emulation library built under -fsanitize=address,undefined — the memory
check available on this pool (compute-sanitizer is closed):
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tests/emu/run_sanitized.py"""
import os, sys, tempfile, pathlib
os.environ["CUDE_EMU_SANITIZE"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
import test_emu_kernel as T
import test_suppression_kernel as S
import test_emu_train as R
fx = dict(np.load(os.path.join(HERE, "..", "golden", "cpeptide_fixtures.npz")))
sup = dict(np.load(os.path.join(HERE, "..", "golden", "suppression_fixtures.npz")))
n = 0
for mod, arg in ((T, fx), (S, sup), (R, None)):
    for name in sorted(dir(mod)):
        f = getattr(mod, name)
        if not name.startswith("test_") or not callable(f) or "gpu" in name:
            continue
        if any(m.name == "gpu" for m in getattr(f, "pytestmark", [])):
            continue
        import inspect
        params = list(inspect.signature(f).parameters)
        with tempfile.TemporaryDirectory() as d:
            kw = {}
            for p in params:
                kw[p] = arg if p in ("fx", "sup") else pathlib.Path(d)
            f(**kw)
        n += 1
        print("ok", name, flush=True)
print("%d emulation tests passed under ASan/UBSan" % n)
