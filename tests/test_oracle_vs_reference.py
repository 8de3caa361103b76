"""Build-container only: runs the UNMODIFIED reference (under /root/reference, through oracle/ref_harness.py) in a
fresh interpreter and checks the travelling oracle against it on a case that is not in tests/golden.
Skipped on the GPU box, where the reference tree does not exist."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

SCRIPT = r"""
import os, sys, json
os.environ["PYTORCH_JIT"] = "0"
sys.path.insert(0, %(root)r)
import numpy as np
from oracle import ref_harness as rh
from oracle import hypad_oracle as ho
rng = np.random.default_rng(99)
T = 700
t = np.arange(T)
s = np.sin(2 * np.pi * t / 61.0) + 0.2 * rng.standard_normal(T)
s[400:404] += 5
ts = 1285027200 + 21600 * t
res = {}
for hyp, comb in ((True, "uncertainty"), (False, "mult")):
    cap = rh.run_univariate(s, ts, hyp, comb, "dtw")
    sig = ho.minmax_scale(s)
    W = ho.rolling_window_sequences(sig[:, None], cap["true_index"], 100)[0][:, :, 0]
    out = ho.univariate_scores(W, cap["weights"], hyp, comb, "dtw", index=cap["true_index"], batch=64)
    rel = np.abs(out["final"] - cap["final_scores"]) / np.abs(cap["final_scores"])
    res["hyp" if hyp else "eucl"] = {
        "critic_equal": bool(np.array_equal(out["critic"], cap["critic"])),
        "kmax_equal": bool(np.array_equal(out["kmax"], rh.kde_argmax_reference(cap["critic"], 100))),
        "final_max_rel": float(np.nanmax(rel)),
        "intervals_equal": bool(out["intervals"].shape == cap["intervals"].shape and
                                np.array_equal(out["intervals"][:, :2], cap["intervals"][:, :2])),
    }
print("RESULT" + json.dumps(res))
"""


@pytest.mark.reference
def test_oracle_matches_live_reference():
    if not os.path.exists("/root/reference/anomaly_detection.py"):
        pytest.skip("reference tree not present (GPU box)")
    p = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT")][-1][6:])
    for k, r in res.items():
        assert r["critic_equal"] and r["kmax_equal"] and r["intervals_equal"], (k, r)
        assert r["final_max_rel"] < 2e-3, (k, r)  # one acosh quantisation step at most (hyperbolic); 1e-6 typical


PRE_SCRIPT = r"""
import os, sys, json
os.environ["PYTORCH_JIT"] = "0"
sys.path.insert(0, %(root)r)
import numpy as np, pandas as pd
from oracle import ref_harness as rh
from oracle import hypad_oracle as ho
rh.bootstrap()
from utils.dataloader import SignalDataset
from sklearn.impute import SimpleImputer
from sklearn.preprocessing import MinMaxScaler
from tests_preprocess_cases import cases
res = {}
for name, (ts, vals, interval) in cases().items():
    df = pd.DataFrame({"timestamp": ts, "value": vals})
    X, index = SignalDataset.time_segments_aggregate(None, df, interval=interval, time_column="timestamp")
    agg, idx = ho.time_segments_aggregate(ts, vals, interval)
    Xs = MinMaxScaler(feature_range=(-1, 1)).fit_transform(SimpleImputer().fit_transform(X))
    mine, _ = ho.preprocess_signal(ts, vals, interval)
    res[name] = {"agg_equal": bool(np.array_equal(X[:, 0], agg, equal_nan=True)), "index_equal": bool(np.array_equal(index, idx)),
                 "scaled_max_abs": float(np.abs(Xs[:, 0] - mine).max()), "n": int(len(agg)), "nan": int(np.isnan(agg).sum())}
print("RESULT" + json.dumps(res))
"""


@pytest.mark.reference
def test_oracle_preprocessing_matches_reference_dataloader():
    """aggregate / impute / MinMax of utils/dataloader.py:83-137, oracle against the reference's own
    time_segments_aggregate and the sklearn transformers it calls, on regular, irregular, gapped and NaN-carrying inputs."""
    if not os.path.exists("/root/reference/anomaly_detection.py"):
        pytest.skip("reference tree not present (GPU box)")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests"))
    p = subprocess.run([sys.executable, "-c", PRE_SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT")][-1][6:])
    assert len(res) >= 4
    for k, r in res.items():
        assert r["agg_equal"] and r["index_equal"], (k, r)
        assert r["scaled_max_abs"] <= 1e-15, (k, r)
