"""Throughput of the BASELINE.json configurations other than the bench workload, on ONE B200 (the bench line is config 3):

  cfg1  HypAD univariate, A-1 shape (T = 8640), hyperbolic, uncertainty
  cfg2  TadGAN Euclidean, same signal, rec_error = dtw, combination = mult
  cfg4  HypAD multivariate, S = C = 123, N = 2^20 rows, hyperbolic, mult      (one GPU's worth of rows)
  cfg5  bulk sweep: 493 signals with a length multiset shaped like the bundled data, one model per signal (seed 1000 + i)
        -- plus the shared-weights variant

Random-init weights (torch.manual_seed), synthetic data; wall-clock around whole `score` / `run` calls after warm-up, with a
synchronize on both sides (these configurations are launch- and host-bound, so wall clock is the honest measure); prints one
JSON line.  Not a bench arm: bench.py measures config 3."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hypad_b200.models.tadgan import CriticX, Decoder, Encoder
from hypad_b200.scoring import WindowScorer
from hypad_b200.sweep import SignalSweep

DEV = torch.device("cuda", 0)
T0, DT = 1285027200, 21600


def model(seed, S, hyperbolic):
    torch.manual_seed(seed)
    enc, dec, cx = Encoder(S, 20), Decoder(S, 20, hyperbolic), CriticX(S, 20)
    return WindowScorer(enc.eval().to(DEV), dec.eval().to(DEV), cx.eval().to(DEV))


def config1_signal(T, seed=None):
    t = np.arange(T)
    s = np.sin(2 * np.pi * t / 50.0)
    if seed is not None:
        s = s + 0.05 * np.random.default_rng(seed).standard_normal(T)
    if T > 20:
        s[T // 2:T // 2 + 5] += 3
    return 2 * (s - s.min()) / (s.max() - s.min()) - 1


def wall(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    res = {}
    T = 8640
    x = torch.from_numpy(config1_signal(T)).to(DEV)
    idx = T0 + DT * np.arange(T)
    for name, hyp, comb in (("cfg1_hyperbolic_uncertainty", True, "uncertainty"), ("cfg2_euclidean_dtw_mult", False, "mult")):
        sc = model(0, 100, hyp)
        out = sc.score(x, sliding=True, combination=comb, rec_error_type="dtw", index=idx)
        t = wall(lambda: sc.score(x, sliding=True, combination=comb, rec_error_type="dtw", index=idx), 20)
        res[name] = {"windows": T - 100, "ms": t * 1e3, "windows_per_s": (T - 100) / t, "intervals": int(len(out["intervals"]))}
    # Euclidean at the bench length: the median / DTW / materialised-reconstruction kernels at size
    T = 1_000_000
    xl = torch.from_numpy(config1_signal(T)).to(DEV)
    sc = model(0, 100, False)
    t = wall(lambda: sc.score(xl, sliding=True, combination="mult", rec_error_type="dtw", index=None), 5)
    res["euclidean_1M_timesteps"] = {"windows": T - 100, "ms": t * 1e3, "windows_per_s": (T - 100) / t}
    del xl
    # cfg4
    N, S = 1 << 20, 123
    g = torch.Generator(device=DEV).manual_seed(4)
    rows = torch.rand(N, S, dtype=torch.float32, device=DEV, generator=g) * 2 - 1
    rows[N // 3:N // 3 + 300] *= 0.2
    sc = model(0, S, True)
    t = wall(lambda: sc.score(rows, sliding=False, combination="mult", multivariate=True, index=None), 5)
    res["cfg4_multivariate_s123"] = {"rows": N, "ms": t * 1e3, "windows_per_s": N / t}
    del rows
    # cfg5
    rng = np.random.default_rng(5)
    lengths = np.concatenate([rng.integers(2000, 8700, 80), rng.integers(1100, 22700, 46), rng.integers(1420, 1700, 367)]).tolist()
    signals = [config1_signal(t, seed=i) for i, t in enumerate(lengths)]
    indices = [T0 + DT * np.arange(t) for t in lengths]
    total = sum(t - 100 for t in lengths)
    t_build = time.perf_counter()
    scorers = [model(1000 + i, 100, True) for i in range(len(lengths))]
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    for label, sw in (("cfg5_sweep_one_model_per_signal", SignalSweep(lambda i: scorers[i])),
                      ("cfg5_sweep_one_model_per_signal_4_streams", SignalSweep(lambda i: scorers[i], streams=4)),
                      ("cfg5_sweep_one_model_per_signal_8_streams", SignalSweep(lambda i: scorers[i], streams=8)),
                      ("cfg5_sweep_shared_weights", SignalSweep(scorers[0]))):
        out = sw.run(signals, indices)
        t = wall(lambda: sw.run(signals, indices), 3, warm=1)
        res[label] = {"signals": len(lengths), "windows": total, "ms": t * 1e3, "windows_per_s": total / t,
                      "signals_with_intervals": int(sum(len(v) > 0 for v in out.values()))}
    res["cfg5_model_construction_s"] = t_build
    print(json.dumps(res))


if __name__ == "__main__":
    main()
