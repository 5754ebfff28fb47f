#!/usr/bin/env python
"""Extract the golden fixtures of the c-peptide cUDE path from the reference's data files.

Run once in the build container (needs /root/reference, which does NOT exist on the GPU box):

    python tests/golden/make_fixtures.py

The reference stores its data as JLD2 (= HDF5, superblock v2 at base offset 512, contiguous
uncompressed little-endian arrays, column-major).  h5py is not available here, so the arrays
are read at fixed object offsets (every offset is sanity-checked below
against values that are independently known from the CSV sources / the scripts).

Sources (all under /root/reference):
  data/ohashi.jld2           written by c-peptide/00-prepare-data.jl:104-136
  data/fujita.jld2           written by c-peptide/00-prepare-data.jl:181-187
  data/ohashi_csv/*.csv      Ohashi et al. 2018, CC BY 4.0 (data/ohashi_csv/LICENSE)
  source_data/cude_neural_parameters.jld2            c-peptide/02-conditional.jl:44-50
  source_data/cude_covariate_neural_parameters_2.jld2  c-peptide/07-covariate-inclusion.jl:59-65

Output: tests/golden/cpeptide_fixtures.npz (a few tens of kB).
"""
import csv
import os
import sys

import numpy as np

REF = os.environ.get("CUDE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpeptide_fixtures.npz")
BASE = 512  # JLD2 user block


def rd(path, addr, dtype, n):
    with open(os.path.join(REF, path), "rb") as f:
        f.seek(BASE + addr)
        buf = f.read(n * np.dtype(dtype).itemsize)
    return np.frombuffer(buf, dtype=dtype).copy()


def mat(path, addr, rows, cols):
    """Julia `Matrix{Float64}(rows, cols)` is column-major."""
    return rd(path, addr, "<f8", rows * cols).reshape(cols, rows).T.copy()


def main():
    fx = {}
    oh = "data/ohashi.jld2"
    # ---- Ohashi train (82) ----
    fx["ohashi_timepoints"] = rd(oh, 15856, "<f8", 5)
    assert np.array_equal(fx["ohashi_timepoints"], [0.0, 30.0, 60.0, 90.0, 120.0])
    fx["ohashi_train_glucose"] = mat(oh, 6912, 82, 5)
    fx["ohashi_train_cpeptide"] = mat(oh, 10304, 82, 5)
    fx["ohashi_train_subject_numbers"] = rd(oh, 13680, "<i8", 82)
    fx["ohashi_train_ages"] = rd(oh, 15992, "<i8", 82)
    # ---- Ohashi test (35) ----
    fx["ohashi_test_glucose"] = mat(oh, 22248, 35, 5)
    fx["ohashi_test_cpeptide"] = mat(oh, 23760, 35, 5)
    fx["ohashi_test_subject_numbers"] = rd(oh, 25256, "<i8", 35)
    fx["ohashi_test_ages"] = rd(oh, 26296, "<i8", 35)

    # `types` are variable-length strings in the JLD2 global heap: recover them by joining the
    # subject numbers with the CSV (the ages must cross-check exactly).
    info = {}
    with open(os.path.join(REF, "data/ohashi_csv/ohashi_subjectinfo.csv")) as f:
        for row in csv.DictReader(f, delimiter=";"):
            info[int(row["No"])] = (int(row["age"]), row["type"].strip())
    ogtt = {}
    with open(os.path.join(REF, "data/ohashi_csv/ohashi_OGTT.csv")) as f:
        rdr = csv.reader(f, delimiter=";")
        next(rdr)
        for row in rdr:
            if any(c.strip() == "" for c in row):
                continue  # dropmissing, 00-prepare-data.jl:15
            ogtt[int(row[0])] = [float(c) for c in row[1:]]
    for split in ("train", "test"):
        nums = fx[f"ohashi_{split}_subject_numbers"]
        ages = np.array([info[int(s)][0] for s in nums])
        assert np.array_equal(ages, fx[f"ohashi_{split}_ages"]), "age cross-check failed"
        types = [info[int(s)][1] for s in nums]
        fx[f"ohashi_{split}_t2dm"] = np.array([t == "T2DM" for t in types], dtype=np.int32)
        fx[f"ohashi_{split}_types"] = np.array(types)
        # unit conversion cross-check, 00-prepare-data.jl:30-31
        g = np.array([ogtt[int(s)][0:5] for s in nums]) * 0.0551
        c = np.array([ogtt[int(s)][10:15] for s in nums]) * 0.3311
        assert np.allclose(g, fx[f"ohashi_{split}_glucose"], rtol=0, atol=1e-12)
        assert np.allclose(c, fx[f"ohashi_{split}_cpeptide"], rtol=0, atol=1e-12)
    cnt = lambda a, t: int(np.sum(a == t))
    tt = fx["ohashi_train_types"]
    assert (cnt(tt, "T2DM"), cnt(tt, "NGT"), cnt(tt, "IGT")) == (36, 34, 12)

    # ---- Fujita (20 individuals x 14 time points, age 29, non-T2DM: 00-prepare-data.jl:178) ----
    fj = "data/fujita.jld2"
    fx["fujita_timepoints"] = rd(fj, 4848, "<i8", 14).astype(np.float64)
    assert fx["fujita_timepoints"][0] == -10 and fx["fujita_timepoints"][-1] == 240
    fx["fujita_glucose"] = mat(fj, 160, 20, 14)
    fx["fujita_cpeptide"] = mat(fj, 2512, 20, 14)
    fx["fujita_ages"] = rd(fj, 5056, "<i8", 20)
    assert np.all(fx["fujita_ages"] == 29)

    # ---- stored cUDE weights: 25 x 37 weights, 25 x 57 betas, best_model_index = 14 (1-based) ----
    cu = "source_data/cude_neural_parameters.jld2"
    fx["cude_neural"] = np.stack([rd(cu, 5176 + 400 * k, "<f8", 37) for k in range(25)])
    fx["cude_betas"] = np.stack([rd(cu, 15512 + 560 * k, "<f8", 57) for k in range(25)])
    fx["cude_best_model_index"] = np.array(14)  # 1-based, as stored
    assert np.all(np.isfinite(fx["cude_neural"])) and np.all(np.abs(fx["cude_neural"]) < 1e3)
    assert np.all(np.isfinite(fx["cude_betas"])) and np.all(np.abs(fx["cude_betas"]) < 50)

    # ---- covariate cUDE (3 inputs): 24 x 41 weights, 24 x 57 betas, best = 2 ----
    cv = "source_data/cude_covariate_neural_parameters_2.jld2"
    fx["cov_neural"] = np.stack([rd(cv, 5168 + 432 * k, "<f8", 41) for k in range(24)])
    fx["cov_betas"] = np.stack([rd(cv, 15864 + 560 * k, "<f8", 57) for k in range(24)])
    fx["cov_best_model_index"] = np.array(2)
    assert np.all(np.isfinite(fx["cov_neural"])) and np.all(np.abs(fx["cov_neural"]) < 1e3)
    assert np.all(np.isfinite(fx["cov_betas"])) and np.all(np.abs(fx["cov_betas"]) < 50)

    # ---- inferred 57-individual training split inside the 82-row train set:
    # order-preserving alignment of stored beta set 14 to the individuals by |dloss_i/dbeta| ~ 0.
    # tests/test_artifacts.py re-derives the stationarity with the oracle.
    fx["train_split_idx"] = np.array(
        [0, 1, 2, 3, 4, 5, 7, 9, 11, 12, 14, 15, 16, 17, 18, 19, 21, 22, 25, 26, 28, 31, 32, 33,
         35, 36, 37, 39, 40, 42, 44, 45, 46, 47, 49, 50, 51, 52, 53, 55, 56, 57, 58, 59, 60, 61,
         63, 65, 66, 69, 71, 72, 73, 74, 79, 80, 81], dtype=np.int64)
    assert fx["train_split_idx"].size == 57

    # ---- cross-check the fixed offsets above with the structural JLD2 reader (conditional_ude_b200/jld2.py) ----
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
    from conditional_ude_b200 import jld2
    c = jld2.load(os.path.join(REF, cu))
    assert np.array_equal(np.stack(c["parameters"]), fx["cude_neural"]) and np.array_equal(np.stack(c["betas"]), fx["cude_betas"])
    assert c["best_model_index"] == 14 and c["width"] == 4 and c["depth"] == 2
    c = jld2.load(os.path.join(REF, cv))
    assert np.array_equal(np.stack(c["parameters"]), fx["cov_neural"]) and np.array_equal(np.stack(c["betas"]), fx["cov_betas"])
    assert c["best_model_index"] == 2

    # ---- second stored cUDE training run (c-peptide/02-conditional.jl with sigma, `cude_neural_parameters_sigma.jld2`):
    #      25 x 37 weights, 25 x 57 betas, best_model_index = 2 — read with the structural reader ----
    sg = jld2.load(os.path.join(REF, "source_data/cude_neural_parameters_sigma.jld2"))
    fx["cude_sigma_neural"] = np.stack(sg["parameters"])
    fx["cude_sigma_betas"] = np.stack(sg["betas"])
    fx["cude_sigma_best_model_index"] = np.array(int(sg["best_model_index"]))
    assert fx["cude_sigma_neural"].shape == (25, 37) and fx["cude_sigma_betas"].shape == (25, 57) and sg["width"] == 4

    np.savez_compressed(OUT, **fx)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(fx), "arrays")

    # ---- suppression example (suppression/suppression.jl:76-91): data tensors + trained networks per lambda ----
    sup = {}
    for lam in ("0.0", "0.01", "1.0"):
        d = jld2.load(os.path.join(REF, f"suppression/results/lambda={lam}.jld2"))
        assert d["group_data"].shape == (3, 8, 37) and d["validation_data"].shape == (3, 8, 30)
        assert len(d["neural_parameters"]) == 25 and d["neural_parameters"][0].shape == (67,)
        key = lam.replace(".", "p")
        sup[f"neural_{key}"] = np.stack(d["neural_parameters"])
        sup[f"losses_{key}"] = d["losses"]
        sup[f"losses_valid_{key}"] = d["losses_valid"]
        sup[f"lambda_{key}"] = np.array(d["λ"])
        if lam == "0.0":
            sup["group_data"], sup["validation_data"] = d["group_data"], d["validation_data"]
            sup["validation_data_nonoise"] = d["validation_data_nonoise"]
            sup["gt_sup_param"], sup["gt_validation_param"] = d["gt_sup_param"], d["gt_validation_param"]
        else:
            assert np.array_equal(d["group_data"], sup["group_data"])       # same data for every lambda
    sup["timepoints"] = np.linspace(0.0, 30.0, 8)                            # range(0, stop=30, length=8)
    out2 = os.path.join(os.path.dirname(OUT), "suppression_fixtures.npz")
    np.savez_compressed(out2, **sup)
    print("wrote", out2, os.path.getsize(out2), "bytes;", len(sup), "arrays")


if __name__ == "__main__":
    sys.exit(main())
