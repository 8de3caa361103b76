"""pyts.metrics.dtw restated (pyts==0.12.0, not in the reference tree, not installable here).

Call site: utils/anomaly_detection_utils.py:853 `dtw(true_data, pred_data)` with pyts defaults
dist='square', method='classic', return_cost=False, return_accumulated=False, return_path=False.
Published algorithm (pyts/metrics/dtw.py, `cost_matrix`, `accumulated_cost_matrix`, `_dtw_classic`):
    cost[i, j] = (x[i] - y[j]) ** 2
    acc[0, 0] = cost[0, 0]; acc[0, j] = acc[0, j-1] + cost[0, j]; acc[i, 0] = acc[i-1, 0] + cost[i, 0]
    acc[i, j] = cost[i, j] + min(acc[i-1, j-1], acc[i-1, j], acc[i, j-1])
    dtw = sqrt(acc[-1, -1])                       (float64 throughout)
The reference has no test pinning this: PARITY UNPINNED at this boundary."""
import numpy as np


def dtw(x, y, dist="square", method="classic", **kw):
    if dist != "square" or method != "classic":
        raise NotImplementedError("only the defaults used by the reference are restated")
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    n, m = x.size, y.size
    cost = (x[:, None] - y[None, :]) ** 2
    acc = np.empty((n, m))
    acc[0] = np.cumsum(cost[0])
    acc[:, 0] = np.cumsum(cost[:, 0])
    for i in range(1, n):
        for j in range(1, m):
            acc[i, j] = cost[i, j] + min(acc[i - 1, j - 1], acc[i - 1, j], acc[i, j - 1])
    return float(np.sqrt(acc[-1, -1]))
