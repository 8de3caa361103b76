"""Device timings of the SURVEY.md 8(f) kernels: preprocessing (aggregate / impute+MinMax / detrend) at 1M rows and the pairwise
Poincare distance at 8192 x 8192 x 100.  CUDA events on the context's stream, after warm-up; prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hypad_b200 import _native
from hypad_b200._native import check, ptr
from hypad_b200.hyperspace.poincare_distance import poincare_distance
from hypad_b200.utils.dataloader import segment_starts


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda:0")
    c = _native.default_context(dev)
    res = {}
    rng = np.random.default_rng(0)
    for rows_per_seg in (1, 4, 64):
        K = 1_000_000
        n = K * rows_per_seg
        ts = torch.arange(n, dtype=torch.float64, device=dev)
        v = torch.from_numpy(rng.standard_normal(n)).to(dev)
        starts = torch.from_numpy(segment_starts(0.0, float(n - 1), float(rows_per_seg))).to(dev)
        K = starts.shape[0]
        agg = torch.empty(K, dtype=torch.float64, device=dev)
        out = torch.empty(K, dtype=torch.float64, device=dev)
        t_agg = timed(lambda: check(c.lib.hypad_segments_aggregate(ptr(ts), ptr(v), n, ptr(starts), float(rows_per_seg), K, ptr(agg), c.stream())))
        t_mm = timed(lambda: check(c.lib.hypad_impute_minmax(c.handle, ptr(agg), K, -1.0, 1.0, ptr(out), c.stream())))
        t_dt = timed(lambda: check(c.lib.hypad_detrend_linear(c.handle, ptr(agg), K, ptr(out), c.stream())))
        res["rows_per_segment_%d" % rows_per_seg] = {
            "segments": K, "rows": n, "aggregate_ms": t_agg, "aggregate_GBps": (16.0 * n + 16.0 * K) / t_agg / 1e6,
            "impute_minmax_ms": t_mm, "impute_minmax_GBps": 24.0 * K / t_mm / 1e6,  # two reads + one write of the column
            "detrend_ms": t_dt, "detrend_GBps": 24.0 * K / t_dt / 1e6}
    for n, d in ((8192, 100), (16384, 20)):
        p = (torch.rand(n, d, device=dev) - 0.5) * 0.15
        q = (torch.rand(n, d, device=dev) - 0.5) * 0.15
        t = timed(lambda: poincare_distance(p, q), reps=10)
        res["pairwise_%dx%dx%d" % (n, n, d)] = {"ms": t, "out_GBps": 4.0 * n * n / t / 1e6, "TFLOPs": 2.0 * n * n * d / t / 1e9,
                                                "pairs_per_s": n * n / t * 1e3}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
