"""Inputs of the preprocessing tests (utils/dataloader.py:83-137): shared by the oracle-vs-reference test (build container)
and the GPU-vs-oracle test."""
import numpy as np


def cases():
    rng = np.random.default_rng(7)
    out = {}
    # regular: one row per interval (the bundled NASA CSVs)
    T = 3000
    ts = 1285027200 + 21600 * np.arange(T)
    out["regular"] = (ts, np.sin(np.arange(T) / 17.0) + 0.1 * rng.standard_normal(T), 21600)
    # several rows per segment, unsorted input, jittered integer timestamps
    T = 5000
    ts = 1000 + np.sort(rng.choice(5000 * 40, T, replace=False))  # distinct: the reference sorts with pandas' unstable
    perm = rng.permutation(T)                                      # quicksort, ties would make its summation order arbitrary
    out["irregular"] = (ts[perm], (np.cos(ts / 900.0) * 3 + rng.standard_normal(T))[perm], 300)
    # gaps (empty segments -> NaN -> imputed) and NaN values
    T = 2000
    ts = 50 + 10 * np.arange(T)
    keep = np.ones(T, bool)
    keep[300:340] = False
    keep[1200:1203] = False
    v = rng.standard_normal(T).cumsum()
    v[500:505] = np.nan
    out["gaps_nan"] = (ts[keep], v[keep], 20)
    # float timestamps (the YAHOO path replaces them by datetime.timestamp floats), interval 1
    T = 1500
    ts = 1353715200.0 + np.arange(T, dtype=np.float64)
    out["float_ts"] = (ts, rng.uniform(-5, 5, T), 1)
    # a constant signal (MinMax range 0) and a single long segment
    out["constant"] = (np.arange(100) * 5, np.full(100, 2.5), 5)
    out["one_segment"] = (np.arange(50), rng.standard_normal(50), 1000)
    return out
