"""Generates the golden vectors of the SURVEY.md 8(f) rows by running the UNMODIFIED reference on the inputs of
tests/tests_preprocess_cases.py:

  tests/golden/preprocess.npz   SignalDataset.time_segments_aggregate (utils/dataloader.py:99-137), then the sklearn
                                SimpleImputer / MinMaxScaler the dataset calls (:86-89); `yahoo/*`: yahoo_preprocess (:41-58)
                                followed by the same chain at interval=1
  tests/golden/pairwise.npz     hyperspace/poincare_distance.py: poincare_distance, pairwise_distances, square_norm
  tests/golden/metrics.json     utils/anomaly_detection_utils.py: contextual_confusion_matrix(weighted=False) (:606-655) and the
                                lines compute_metrics prints (:241-254) on seeded interval lists

Run in the build container only:   python oracle/make_golden_next.py
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

os.environ["PYTORCH_JIT"] = "0"
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import pandas as pd
import torch

from oracle import ref_harness as rh


def main():
    rh.bootstrap()
    from sklearn.impute import SimpleImputer
    from sklearn.preprocessing import MinMaxScaler
    from tests_preprocess_cases import cases, pairwise_cases, yahoo_cases
    from utils.dataloader import SignalDataset, yahoo_preprocess

    def chain(df, interval):
        X, index = SignalDataset.time_segments_aggregate(None, df, interval=interval, time_column="timestamp")
        Xs = MinMaxScaler(feature_range=(-1, 1)).fit_transform(SimpleImputer().fit_transform(X))
        return np.asarray(X[:, 0], dtype=np.float64), np.asarray(index), np.asarray(Xs[:, 0], dtype=np.float64)

    g = {}
    for name, (ts, vals, interval) in cases().items():
        g[name + "/agg"], g[name + "/index"], g[name + "/scaled"] = chain(pd.DataFrame({"timestamp": ts, "value": vals}), interval)
        print(name, g[name + "/agg"].shape, int(np.isnan(g[name + "/agg"]).sum()), "NaN segments")
    for name, (vals, flag) in yahoo_cases().items():
        df = yahoo_preprocess(pd.DataFrame({"timestamp": np.arange(1, len(vals) + 1), "value": vals, "is_anomaly": flag}))
        g["yahoo/" + name + "/detrended"] = df["value"].values.astype(np.float64)
        g["yahoo/" + name + "/timestamp"] = df["timestamp"].values.astype(np.float64)
        _, g["yahoo/" + name + "/index"], g["yahoo/" + name + "/scaled"] = chain(df, 1)
        print("yahoo", name, len(vals))
    # save_known_anomalies (:14-33) writes <csv path minus .csv>_known_anomalies.csv; keep the rows it wrote
    import tempfile

    from utils.dataloader import save_known_anomalies

    for name, (vals, flag) in yahoo_cases().items():
        for col in ("is_anomaly", "anomaly"):
            with tempfile.TemporaryDirectory() as d:
                path = os.path.join(d, "sig.csv")
                df = pd.DataFrame({"timestamp": 1000.0 + np.arange(len(vals)), "value": vals, col: flag})
                try:
                    save_known_anomalies(df, path)
                    runs = pd.read_csv(path[:-4] + "_known_anomalies.csv")[["start", "end"]].values.astype(np.float64)
                except ValueError:  # no labelled run: the reference fails renaming the columns of an empty frame
                    runs = np.empty((0, 2))
            g["yahoo/" + name + "/known_" + col] = runs
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess.npz"), **g)

    from hyperspace.poincare_distance import pairwise_distances, poincare_distance, square_norm

    g = {}
    for name, (p, q) in pairwise_cases().items():
        tp, tq = torch.from_numpy(p), torch.from_numpy(q)
        g[name + "/poincare"] = poincare_distance(tp, tq).numpy()
        g[name + "/sqdist"] = pairwise_distances(tp, tq).numpy()
        g[name + "/sqdist_self"] = pairwise_distances(tp).numpy()
        g[name + "/square_norm"] = square_norm(tp).numpy()
        print("pairwise", name, g[name + "/poincare"].shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pairwise.npz"), **g)

    import contextlib
    import io
    import json

    import utils.anomaly_detection_utils as adu
    from tests_preprocess_cases import interval_cases

    out = []
    for expected, observed in interval_cases():
        e = pd.DataFrame(expected, columns=["start", "end"])
        o = pd.DataFrame([list(x) + [1.0] for x in observed], columns=["start", "end", "score"])
        rec = {"expected": expected, "observed": observed, "counts": list(adu.contextual_confusion_matrix(e, o, weighted=False))}
        rec["counts_lists"] = list(adu.contextual_confusion_matrix([tuple(x) for x in expected], [tuple(x) for x in observed], weighted=False))
        buf = io.StringIO()
        try:
            with contextlib.redirect_stdout(buf):
                adu.compute_metrics(e, o)
            rec["printed"] = buf.getvalue()
        except ZeroDivisionError:
            rec["printed"] = None
        rec["counts"] = [None if c is None else int(c) for c in rec["counts"]]
        rec["counts_lists"] = [None if c is None else int(c) for c in rec["counts_lists"]]
        out.append(rec)
    with open(os.path.join(ROOT, "tests", "golden", "metrics.json"), "w") as f:
        json.dump(out, f)
    print("metrics", len(out), "cases,", sum(r["printed"] is None for r in out), "with an empty denominator")


if __name__ == "__main__":
    main()
