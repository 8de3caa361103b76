"""Cost of the O(T) finish (what every rank repeats on the gathered arrays) as a function of the total length."""
import os, sys, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hypad_b200 import scoring
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
for T in (1_000_000, 2_000_000, 4_000_000, 8_000_000):
    n = T - 100
    kmax = torch.randn(T - 1, dtype=torch.float64, device=dev, generator=g) * 1e-3
    rec = torch.rand(n, dtype=torch.float32, device=dev, generator=g) * 1e-2 + 0.01
    rec[::50000] += 0.5
    unorm = torch.rand(n, dtype=torch.float32, device=dev, generator=g) * 0.1 + 0.2
    index = np.arange(T, dtype=np.int64)
    def step(parts):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        cs = scoring.critic_zscore_smooth(kmax, math.trunc(n * 0.01))
        ev[1].record()
        final = scoring.combine("uncertainty", cs[:n], rec, unorm, n=n)
        ev[2].record()
        iv = scoring.find_anomaly_intervals(final, index, 0.33, 0.1, anomaly_padding=50, ddof=1)
        ev[3].record()
        torch.cuda.synchronize()
        for i in range(3): parts[i] += ev[i].elapsed_time(ev[i + 1])
        return iv
    for _ in range(3): step([0, 0, 0])
    parts = [0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    for _ in range(10): iv = step(parts)
    wall = (time.perf_counter() - t0) / 10 * 1e3
    print("T=%8d: critic_zscore_smooth %.3f ms  combine %.3f ms  find_anomalies %.3f ms  | wall %.3f ms  (%d intervals)" % (T, parts[0] / 10, parts[1] / 10, parts[2] / 10, wall, len(iv)))
