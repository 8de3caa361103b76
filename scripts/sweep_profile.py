"""Where the host time of a by-signal sweep goes: cProfile over SignalSweep.run on 200 short signals with one shared scorer,
plus wall-clock of the two phases.  The sweep of BASELINE config 5 is bound by host calls (0.4 ms per signal), not by kernels."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hypad_b200 import _native
from hypad_b200.models.tadgan import CriticX, Decoder, Encoder
from hypad_b200.scoring import WindowScorer
from hypad_b200.sweep import SignalSweep

DEV = torch.device("cuda", 0)
torch.manual_seed(0)
sc = WindowScorer(Encoder(100, 20).eval().to(DEV), Decoder(100, 20, True).eval().to(DEV), CriticX(100, 20).eval().to(DEV))
rng = np.random.default_rng(5)
lengths = rng.integers(1420, 1700, 200).tolist()
signals, indices = [], []
for i, T in enumerate(lengths):
    t = np.arange(T)
    s = np.sin(2 * np.pi * t / 50.0) + 0.05 * rng.standard_normal(T)
    s[T // 2:T // 2 + 5] += 3
    signals.append(2 * (s - s.min()) / (s.max() - s.min()) - 1)
    indices.append(1285027200 + 21600 * t)
sw = SignalSweep(sc)
sw.run(signals, indices)
torch.cuda.synchronize()
lib = _native.load_library()
l0 = lib.hypad_launch_count()
t0 = time.perf_counter()
sw.run(signals, indices)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("sweep of %d signals: %.1f ms, %.3f ms per signal, %d launches per signal" % (len(signals), dt * 1e3, dt * 1e3 / len(signals), (lib.hypad_launch_count() - l0) // len(signals)))
# phase 1 only (enqueue, no host synchronisation)
dev_signals = [torch.from_numpy(s).to(DEV) for s in signals]
torch.cuda.synchronize()
t0 = time.perf_counter()
outs = [sc.score(x, sliding=True, combination="uncertainty", index=None, poll=False) for x in dev_signals]
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_dev = time.perf_counter() - t0
print("phase 1 (score, no index, no poll): host enqueue %.3f ms per signal, device done after %.3f ms per signal" % (t_enq * 1e3 / len(signals), t_dev * 1e3 / len(signals)))
pr = cProfile.Profile()
pr.enable()
sw.run(signals, indices)
torch.cuda.synchronize()
pr.disable()
buf = io.StringIO()
pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(45)
print(buf.getvalue()[:9000])
