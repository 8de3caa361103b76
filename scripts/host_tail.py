"""Where the step's non-kernel time goes: host-side timing of WindowScorer.score on the bench workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_signal
from hypad_b200 import scoring
from hypad_b200.models.tadgan import Encoder, Decoder, CriticX
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mods = [Encoder(100, 20), Decoder(100, 20, True), CriticX(100, 20)]
for m in mods: m.to(dev).eval()
sc = scoring.WindowScorer(*mods)
T = 1000000
sig = torch.from_numpy(make_signal(T)).to(dev)
index = np.arange(T, dtype=np.int64)
for _ in range(3): sc.score(sig, True, "uncertainty", index=index)
torch.cuda.synchronize()
orig_tw, orig_ifr, orig_poll = scoring.threshold_windows, scoring.intervals_from_runs, sc.poll_error
marks = {}
def tw(*a, **k):
    marks["tw_call"] = time.perf_counter(); r = orig_tw(*a, **k); marks["tw_ret"] = time.perf_counter(); return r
def ifr(*a, **k):
    r = orig_ifr(*a, **k); marks["ifr_ret"] = time.perf_counter(); return r
scoring.threshold_windows, scoring.intervals_from_runs = tw, ifr
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = sc.score(sig, True, "uncertainty", index=index)
    t1 = time.perf_counter()
    print("total %.3f ms | host enqueue until threshold call %.3f | threshold_windows (launch + D2H sync) %.3f | intervals_from_runs %.3f | rest (index map, poll) %.3f"
          % (1e3 * (t1 - t0), 1e3 * (marks["tw_call"] - t0), 1e3 * (marks["tw_ret"] - marks["tw_call"]), 1e3 * (marks["ifr_ret"] - marks["tw_ret"]), 1e3 * (t1 - marks["ifr_ret"])))
