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
