"""CSV ingest: the data side of the path (reference c-peptide/00-prepare-data.jl).

`prepare_ohashi(ogtt_csv, subjectinfo_csv)` and `prepare_fujita(ogtt_csv)` read the reference's raw CSV files
(`data/ohashi_csv/ohashi_OGTT.csv` + `ohashi_subjectinfo.csv`, `data/fujita_csv/fujita_ogtt.csv`) and return the arrays
the model constructors take — glucose in mmol/L (mg/dL x 0.0551) and c-peptide in nmol/L (ng/mL x 0.3311),
00-prepare-data.jl:30-31, :178-179 — so that a population can be built without the intermediate JLD2 file:

    d = prepare_ohashi("data/ohashi_csv/ohashi_OGTT.csv", "data/ohashi_csv/ohashi_subjectinfo.csv")
    models = [CPeptideConditionalUDEModel(d["glucose"][i], d["timepoints"], d["ages"][i], chain(4, 2, "tanh"),
                                          d["cpeptide"][i], d["types"][i] == "T2DM") for i in range(len(d["ages"]))]

The train/test split of the reference (`stratified_split` with `StableRNG(270523)`) depends on Julia's RNG stream and is
not reproduced here; `split_like_reference(d, subject_numbers)` selects rows by the subject numbers stored in
`data/ohashi.jld2` instead.
"""
import csv

import numpy as np

GLUCOSE_MGDL_TO_MMOLL = 0.0551      # 00-prepare-data.jl:30
CPEPTIDE_NGML_TO_NMOLL = 0.3311     # 00-prepare-data.jl:31

__all__ = ["prepare_ohashi", "prepare_fujita", "split_like_reference"]


def _rows(path, delimiter):
    with open(path, newline="", encoding="utf-8-sig") as f:
        r = list(csv.reader(f, delimiter=delimiter))
    return [h.strip().strip('"') for h in r[0]], [row for row in r[1:] if any(c.strip() for c in row)]


def prepare_ohashi(ogtt_csv, subjectinfo_csv):
    """00-prepare-data.jl:14-31: rows of the OGTT file with any missing value are dropped (`dropmissing`), the subject
    information is joined on `No`; glucose = columns 2-6 (O-PG), c-peptide = columns 12-16 (O-CPR), time points
    0, 30, 60, 90, 120 min.  Returns a dict of arrays (n subjects)."""
    hdr, rows = _rows(ogtt_csv, ";")
    keep = [row for row in rows if len(row) == len(hdr) and all(c.strip() != "" for c in row)]
    data = np.array([[float(c.replace(",", ".")) for c in row] for row in keep], dtype=np.float64)
    subject_numbers = data[:, 0].astype(np.int64)
    ihdr, irows = _rows(subjectinfo_csv, ";")
    col = {name: k for k, name in enumerate(ihdr)}
    info = {int(row[col["No"]]): row for row in irows}
    missing = [s for s in subject_numbers if s not in info]
    if missing:
        raise ValueError(f"subjects without subject information: {missing}")
    pick = lambda name, conv: np.array([conv(info[s][col[name]]) for s in subject_numbers])
    return dict(
        subject_numbers=subject_numbers,
        timepoints=np.array([0.0, 30.0, 60.0, 90.0, 120.0]),
        glucose=data[:, 1:6] * GLUCOSE_MGDL_TO_MMOLL,
        cpeptide=data[:, 11:16] * CPEPTIDE_NGML_TO_NMOLL,
        types=pick("type", str),
        ages=pick("age", lambda x: int(float(x))),
        body_weights=pick("BW", float),
        bmis=pick("BMI", float),
    )


def prepare_fujita(ogtt_csv):
    """00-prepare-data.jl:172-180: rows `Glucose` / `C-peptide` of the long-format table, columns 3..end-1 are the time
    points (minutes, header), every subject is 29 years old."""
    hdr, rows = _rows(ogtt_csv, ",")
    tcols = list(range(2, len(hdr) - 1))
    timepoints = np.array([float(hdr[k]) for k in tcols])
    sel = lambda mol: np.array([[float(row[k]) for k in tcols] for row in rows if row[0] == mol], dtype=np.float64)
    glucose = sel("Glucose") * GLUCOSE_MGDL_TO_MMOLL
    cpeptide = sel("C-peptide") * CPEPTIDE_NGML_TO_NMOLL
    return dict(timepoints=timepoints, glucose=glucose, cpeptide=cpeptide, ages=np.full(glucose.shape[0], 29, dtype=np.int64),
                types=np.array(["NGT"] * glucose.shape[0]))


def split_like_reference(d, subject_numbers):
    """Rows of a `prepare_ohashi` result for the given subject numbers, in that order (e.g. the `subject_numbers` of the
    train / test groups stored in data/ohashi.jld2)."""
    pos = {int(s): k for k, s in enumerate(d["subject_numbers"])}
    idx = np.array([pos[int(s)] for s in subject_numbers], dtype=np.int64)
    return {k: (v if k == "timepoints" else v[idx]) for k, v in d.items()}
