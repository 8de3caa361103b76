"""Generates tests/golden/preprocess.npz by running the UNMODIFIED reference preprocessing (utils/dataloader.py:83-137:
SignalDataset.time_segments_aggregate, then the sklearn SimpleImputer / MinMaxScaler it calls) on the inputs of
tests/tests_preprocess_cases.py.

Run in the build container only:   python oracle/make_golden_preprocess.py
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

from oracle import ref_harness as rh


def main():
    rh.bootstrap()
    from sklearn.impute import SimpleImputer
    from sklearn.preprocessing import MinMaxScaler
    from tests_preprocess_cases import cases
    from utils.dataloader import SignalDataset

    g = {}
    for name, (ts, vals, interval) in cases().items():
        df = pd.DataFrame({"timestamp": ts, "value": vals})
        X, index = SignalDataset.time_segments_aggregate(None, df, interval=interval, time_column="timestamp")
        Xs = MinMaxScaler(feature_range=(-1, 1)).fit_transform(SimpleImputer().fit_transform(X))
        g[name + "/agg"] = np.asarray(X[:, 0], dtype=np.float64)
        g[name + "/index"] = np.asarray(index)
        g[name + "/scaled"] = np.asarray(Xs[:, 0], dtype=np.float64)
        print(name, X.shape, int(np.isnan(X).sum()), "NaN segments")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess.npz"), **g)


if __name__ == "__main__":
    main()
