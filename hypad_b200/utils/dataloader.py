"""Drop-in for the window construction of utils/dataloader.py of the reference.

`rolling_window_sequences` keeps the reference's signature (utils/dataloader.py:139-150) and return values; the window
matrix is produced by the coalesced sm_100a gather kernel (`hypad_window_gather`).  The fused scoring pipeline never
needs it -- windows are overlapping views of the signal -- so this exists for callers that want the materialised
array.  CSV reading, interval aggregation, imputation and MinMax scaling (utils/dataloader.py:61-137) are the step
before the path and stay in pandas/sklearn on the host.
"""
import numpy as np
import torch

from .. import scoring as _sc


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
