"""Debug: critic-mode intervals on cfg1 (GPU) -- dumps final + thresholding stats to gpurun_out/."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from conftest import build_modules, full_signal, golden
from hypad_b200.scoring import WindowScorer, analysis_windows, threshold_windows
from hypad_b200 import scoring

dev = torch.device("cuda", 0)
enc, dec, cx, _ = build_modules("weights_hyp_s100.npz", 100, True, dev)
sc = WindowScorer(enc, dec, cx)
b = golden("cfg1_hyp_uncertainty.npz")
sig = torch.from_numpy(full_signal(b)).to(dev)
res = {}
for comb in ("critic", "critic_uncertainty", "mult"):
    out = sc.score(sig, True, comb, index=b["index"])
    final = out["final"]
    n = final.numel()
    ddof = 0 if comb.startswith("critic") else 1
    w, s, c = analysis_windows(n, None, 0.33, None, 0.1)
    stats, runs, nr = threshold_windows(final, w, s, c, ddof, 50)
    stats_e, runs_e, nr_e = threshold_windows(final, w, s, c, ddof, 50, exhaustive=True)
    print(comb, "windows", w, s, c, "n_runs", nr.tolist(), nr_e.tolist())
    print(" stats equal exhaustive:", np.array_equal(stats, stats_e), "runs equal:", all(np.array_equal(runs[k][:nr[k]], runs_e[k][:nr_e[k]]) for k in range(c)))
    print(out["intervals"])
    res[comb + "_final"] = final.cpu().numpy(); res[comb + "_stats"] = stats; res[comb + "_runs"] = runs; res[comb + "_nr"] = nr
    res[comb + "_intervals"] = out["intervals"]
np.savez("gpurun_out/debug_critic.npz", **res)
