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


def yahoo_cases():
    """(values, is_anomaly) of YAHOO-shaped signals: trend + seasonality + noise, a few labelled runs."""
    rng = np.random.default_rng(19)
    out = {}
    for name, T, slope in (("a1_like", 1420, 0.8), ("long", 20000, -0.003), ("tiny", 2, 1.0)):
        t = np.arange(T)
        v = 100.0 + slope * t + 20.0 * np.sin(t / 24.0 * 2 * np.pi) + 3.0 * rng.standard_normal(T)
        flag = np.zeros(T, dtype=np.int64)
        if T > 100:
            flag[T // 3:T // 3 + 4] = 1
            flag[T - 20:T - 18] = 1
            v[flag == 1] += 80.0
        out[name] = (v, flag)
    return out


def pairwise_cases():
    """(pred, gt) fp32 point sets inside the unit ball; shapes off the 64-tile grid, D off the 16-feature stage."""
    rng = np.random.default_rng(23)
    out = {}
    for name, n, m, d, r in (("square", 128, 128, 100, 0.6), ("ragged", 77, 201, 123, 0.9), ("thin", 1, 65, 7, 0.5), ("near_origin", 33, 40, 20, 1e-4)):
        p = rng.standard_normal((n, d))
        g = rng.standard_normal((m, d))
        p = p / np.linalg.norm(p, axis=1, keepdims=True) * rng.uniform(0.05, 1.0, (n, 1)) * r
        g = g / np.linalg.norm(g, axis=1, keepdims=True) * rng.uniform(0.05, 1.0, (m, 1)) * r
        out[name] = (p.astype(np.float32), g.astype(np.float32))
    return out


def interval_cases():
    """(expected, observed) lists of closed [start, end] timestamp intervals: random, touching, nested, duplicated, empty."""
    rng = np.random.default_rng(31)
    out = [([[10, 20]], [[21, 30]]),              # adjacent after padding: (10, 21) vs (21, 31) do not overlap
           ([[10, 20]], [[20, 30]]),              # share one timestamp
           ([[10, 50]], [[20, 25], [30, 35], [20, 25]]),  # nested, duplicated observed
           ([[10, 20], [40, 50]], [[0, 100]]),    # one observed covers both
           ([], [[1, 2]]), ([[1, 2]], []),
           ([[5, 5]], [[5, 5]])]
    for _ in range(25):
        ne, no = rng.integers(0, 7), rng.integers(0, 9)
        def draw(k):
            s = np.sort(rng.integers(1_285_000_000, 1_285_000_000 + 21600 * 400, k))
            return [[int(a), int(a + rng.integers(0, 21600 * 30))] for a in s]
        out.append((draw(ne), draw(no)))
    return out
