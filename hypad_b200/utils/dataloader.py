"""Drop-in for utils/dataloader.py of the reference: preprocessing and window construction.

`SignalDataset` keeps the reference's constructor and attributes (utils/dataloader.py:61-97, 224-232).  The CSV is read on
the host (pandas); interval aggregation, mean imputation and MinMax scaling to (-1, 1) (:83-89, :99-137 -- a Python loop
with one pandas slice per segment in the reference) run on the device (`hypad_segments_aggregate`, `hypad_impute_minmax`),
and the scaled signal stays there as `.signal` for the fused scoring path, which never materialises windows.
`rolling_window_sequences` keeps the reference's signature (:139-150) and return values; the window matrix is produced by
the coalesced sm_100a gather kernel (`hypad_window_gather`) for callers that want the materialised array.
The YAHOO branch (:64-79, `yahoo_preprocess` :41-58) detrends on the device too (`hypad_detrend_linear`) and replaces the
timestamps by the reference's one-per-second index starting 2012-11-24 (local time, like `datetime.timestamp`).
"""
import numpy as np
import torch

from .. import _native
from .. import scoring as _sc
from .._native import check, ptr


def segment_starts(first, last, interval):
    """The segment starts of the reference's loop (`while start_ts <= max_ts: ...; start_ts = end_ts`, :127-135): repeated
    addition, so that non-integer intervals accumulate the same rounding."""
    n = int(np.floor((last - first) / interval)) + 2
    starts = np.cumsum(np.concatenate(([first], np.full(n, interval, dtype=np.result_type(first, interval)))))  # sequential adds
    return starts[starts <= last]


def preprocess_signal(timestamps, values, interval=21600, feature_range=(-1.0, 1.0), device=None):
    """utils/dataloader.py:83-89 on the device: time_segments_aggregate(mean) -> SimpleImputer() -> MinMaxScaler(-1, 1).

    timestamps, values: 1-D arrays (any order).  Returns (X, index): the scaled signal as a float64 device tensor (K,) and the
    segment starts (K,) as a numpy array with the timestamps' dtype."""
    ts = np.asarray(timestamps)
    order = np.argsort(ts, kind="stable")  # the reference's sort_values is pandas' unstable quicksort: ties are arbitrary there
    ts_sorted = ts[order]
    index = segment_starts(ts_sorted[0], ts_sorted[-1], interval)
    dev = _sc.cuda_device(device)
    d_ts = torch.from_numpy(np.ascontiguousarray(ts_sorted, dtype=np.float64)).to(dev)
    d_v = torch.from_numpy(np.ascontiguousarray(np.asarray(values, dtype=np.float64)[order])).to(dev)
    d_start = torch.from_numpy(np.ascontiguousarray(index, dtype=np.float64)).to(dev)
    K = index.shape[0]
    agg = torch.empty(K, dtype=torch.float64, device=dev)
    out = torch.empty(K, dtype=torch.float64, device=dev)
    c = _native.default_context(dev)
    with torch.cuda.device(dev):
        check(c.lib.hypad_segments_aggregate(ptr(d_ts), ptr(d_v), d_ts.shape[0], ptr(d_start), float(interval), K, ptr(agg), c.stream()))
        check(c.lib.hypad_impute_minmax(c.handle, ptr(agg), K, float(feature_range[0]), float(feature_range[1]), ptr(out), c.stream()))
    return out, index


def detrend_signal(values, device=None):
    """scipy.signal.detrend(values) (type="linear") on the device, utils/dataloader.py:36-38.  Returns a float64 device tensor."""
    dev = _sc.cuda_device(device)
    v = torch.from_numpy(np.array(values, dtype=np.float64, order="C")).to(dev)  # a copy: pandas hands out read-only views
    if v.dim() != 1 or v.shape[0] < 1:
        raise ValueError("hypad_b200: detrend_signal expects a non-empty 1-D signal")
    out = torch.empty_like(v)
    c = _native.default_context(dev)
    with torch.cuda.device(dev):
        check(c.lib.hypad_detrend_linear(c.handle, ptr(v), v.shape[0], ptr(out), c.stream()))
    return out


def yahoo_index(n):
    """The synthetic timestamps of the YAHOO branch (:44-48, :67-75): one per second from 2012-11-24 00:00:00 to 2012-11-30
    00:00:00 local time (518 401 of them), the first n.  Like the reference, a longer signal is an error (pandas refuses the
    shorter column)."""
    from datetime import datetime, timedelta

    first, last = datetime(2012, 11, 24), datetime(2012, 11, 30)
    total = int((last - first).total_seconds()) + 1
    if n > total:
        raise ValueError("Length of values (%d) does not match length of index (%d)" % (total, n))
    b, e = first.timestamp(), last.timestamp()
    if e - b == total - 1:  # no clock change inside the range: consecutive seconds
        return b + np.arange(n, dtype=np.float64)
    return np.array([(first + timedelta(seconds=i)).timestamp() for i in range(n)])


def known_anomaly_runs(df):
    """The (start, end) timestamp pairs `save_known_anomalies` collects (:14-33): one per run of is_anomaly == 1, last run
    first (the reference prepends).  Files with an `anomaly` column instead (:19-21) are sorted by timestamp first."""
    if "is_anomaly" not in df.columns:
        df = df[["timestamp", "value", "anomaly"]].copy().sort_values(by=["timestamp"])
        df.columns = ["timestamp", "value", "is_anomaly"]
    flag = (df["is_anomaly"].values == 1).astype(np.int8)
    ts = df["timestamp"].values
    edges = np.flatnonzero(np.diff(np.concatenate(([0], flag, [0]))))
    starts, ends = edges[0::2], edges[1::2] - 1
    return df, np.stack([ts[starts], ts[ends]], axis=1)[::-1] if len(starts) else np.empty((0, 2))


def yahoo_preprocess(df, device=None):
    """utils/dataloader.py:41-58: detrended values and the synthetic per-second timestamps; returns df[["timestamp", "value"]]."""
    df = df.copy()
    df["value"] = detrend_signal(df["value"].values, device).cpu().numpy()
    df["timestamp"] = yahoo_index(len(df))
    if "is_anomaly" not in df.columns:
        df = df[["timestamp", "value", "anomaly"]].copy().sort_values(by=["timestamp"])
    return df[["timestamp", "value"]]


class SignalDataset(torch.utils.data.Dataset):
    """utils/dataloader.py:61-97, 224-232.  `.X` (N, window, 1), `.y`, `.X_index`, `.y_index`, `.index` as in the reference
    (host arrays, built on first use); `.signal` is the scaled signal on the device -- what `WindowScorer.score(sliding=True)`
    consumes."""

    def __init__(self, path, interval=21600, windows_size=100, test=False, yahoo=None):
        import pandas as pd

        self.signal_df = pd.read_csv(path)
        if yahoo:
            self.signal_df["value"] = detrend_signal(self.signal_df["value"].values).cpu().numpy()
            self.signal_df["timestamp"] = yahoo_index(len(self.signal_df))
            self.signal_df, runs = known_anomaly_runs(self.signal_df)  # :77, the side file the evaluation reads later
            pd.DataFrame(runs, columns=["start", "end"]).to_csv(path[:-4] + "_known_anomalies.csv")
            self.signal_df = self.signal_df[["timestamp", "value"]]
        self.interval = interval
        self.windows_size = windows_size
        self.test = test
        self.signal, self.index = preprocess_signal(self.signal_df["timestamp"].values, self.signal_df["value"].values, interval)
        self._windows = None

    def _build(self):
        if self._windows is None:
            X = self.signal.cpu().numpy().reshape(-1, 1)
            self._windows = rolling_window_sequences(X, self.index, window_size=self.windows_size, target_size=1, step_size=1,
                                                     target_column=0)
        return self._windows

    X = property(lambda self: self._build()[0])
    y = property(lambda self: self._build()[1])
    X_index = property(lambda self: self._build()[2])
    y_index = property(lambda self: self._build()[3])

    def __len__(self):
        return max(0, self.signal.shape[0] - self.windows_size)

    def __getitem__(self, idx):
        x = torch.from_numpy(self.X[idx])
        if self.test:
            return x, self.index, self.y, self.y_index, self.X_index
        return x


def rolling_window_sequences(X, index, window_size, target_size, step_size, target_column, offset=0, drop=None, drop_windows=False):
    """utils/dataloader.py:139-222 for step_size=1, target_size=1, offset=0 and no dropping (the only call, :90-97).

    X (T, 1) float64 -> (out_X (N, window, 1), out_y (N, 1), X_index (N,), y_index (N,)) with N = T - window_size."""
    if step_size != 1 or target_size != 1 or offset != 0 or drop_windows:
        raise NotImplementedError("hypad_b200: only rolling_window_sequences(step_size=1, target_size=1, offset=0, "
                                  "drop_windows=False) is on the scoring path (utils/dataloader.py:90-97)")
    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 2 or X.shape[1] != 1:
        raise NotImplementedError("hypad_b200: univariate (T, 1) input expected")
    index = np.asarray(index)
    n = len(X) - window_size - target_size - offset + 1
    if n <= 0:
        return np.asarray([]), np.asarray([]), np.asarray([]), np.asarray([])
    dev = _sc.cuda_device()
    W = _sc.window_gather(torch.from_numpy(X[:, 0]).to(dev), window_size)[:n]
    out_X = W.cpu().numpy().reshape(n, window_size, 1)
    out_y = X[window_size:window_size + n, target_column].reshape(n, 1)
    return out_X, out_y, index[:n], index[window_size:window_size + n]
