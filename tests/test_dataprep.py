"""CSV ingest (reference c-peptide/00-prepare-data.jl): synthetic files in the reference's formats, and — when the
reference checkout is present (this container, not the GPU box) — the real CSV files against the arrays extracted from
the reference's own data/ohashi.jld2 and data/fujita.jld2 (tests/golden/cpeptide_fixtures.npz)."""
import os

import numpy as np
import pytest

from conditional_ude_b200 import dataprep

HERE = os.path.dirname(os.path.abspath(__file__))
CSV = os.path.join(HERE, "golden", "csv")
REF = "/root/reference/data"


def test_ohashi_format_units_and_dropmissing():
    d = dataprep.prepare_ohashi(os.path.join(CSV, "ogtt_sample.csv"), os.path.join(CSV, "subjectinfo_sample.csv"))
    assert list(d["subject_numbers"]) == [1, 3]                       # subject 2 has a missing value: dropped
    assert np.allclose(d["glucose"][0], np.array([100, 150, 200, 160, 120]) * 0.0551)
    assert np.allclose(d["cpeptide"][1], np.array([1.5, 4.0, 6.0, 7.0, 6.5]) * 0.3311)
    assert list(d["types"]) == ["NGT", "T2DM"] and list(d["ages"]) == [34, 61]
    assert np.array_equal(d["timepoints"], [0, 30, 60, 90, 120])
    s = dataprep.split_like_reference(d, [3])
    assert s["glucose"].shape == (1, 5) and s["ages"][0] == 61


def test_fujita_format():
    d = dataprep.prepare_fujita(os.path.join(CSV, "fujita_sample.csv"))
    assert np.array_equal(d["timepoints"], [-10, 0, 10]) and d["glucose"].shape == (2, 3)
    assert np.allclose(d["cpeptide"][1], np.array([0.9, 1.0, 2.2]) * 0.3311) and list(d["ages"]) == [29, 29]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_real_csv_files_reproduce_the_reference_jld2_arrays(fx):
    d = dataprep.prepare_ohashi(os.path.join(REF, "ohashi_csv", "ohashi_OGTT.csv"), os.path.join(REF, "ohashi_csv", "ohashi_subjectinfo.csv"))
    assert len(d["ages"]) == 117                                       # 121 rows, 4 with missing values
    for part in ("train", "test"):
        s = dataprep.split_like_reference(d, fx[f"ohashi_{part}_subject_numbers"])
        assert np.allclose(s["glucose"], fx[f"ohashi_{part}_glucose"], rtol=1e-13, atol=0)
        assert np.allclose(s["cpeptide"], fx[f"ohashi_{part}_cpeptide"], rtol=1e-13, atol=0)
        assert np.array_equal(s["ages"], fx[f"ohashi_{part}_ages"]) and np.array_equal(s["types"], fx[f"ohashi_{part}_types"])
    f = dataprep.prepare_fujita(os.path.join(REF, "fujita_csv", "fujita_ogtt.csv"))
    assert np.array_equal(f["timepoints"], fx["fujita_timepoints"])
    assert np.allclose(f["glucose"], fx["fujita_glucose"], rtol=1e-13) and np.allclose(f["cpeptide"], fx["fujita_cpeptide"], rtol=1e-13)
