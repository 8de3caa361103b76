import sys, os, time, math
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
from hypad_b200 import scoring
from hypad_b200.models.tadgan import CriticX, Decoder, Encoder
S = 123
dev = torch.device("cuda", 0)
torch.manual_seed(0)
enc, dec, cx = Encoder(S, 20), Decoder(S, 20, True), CriticX(S, 20)
for m in (enc, dec, cx): m.to(dev).eval()
sc = scoring.WindowScorer(enc, dec, cx)
for n in (1 << 20, 1 << 22):
    g = torch.Generator(device=dev).manual_seed(4)
    rows = torch.rand(n, S, dtype=torch.float32, device=dev, generator=g) * 2 - 1
    index = 1353715200.0 + np.arange(n)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fw = sc.forward(rows, False); torch.cuda.synchronize(); t1 = time.perf_counter()
        rec = scoring.zscore_clip(fw["rec"]); cs, kmax = sc.critic_scores(fw["critic"], n)
        final = scoring.combine("mult", cs[:n], rec, fw["unorm"], n=n); torch.cuda.synchronize(); t2 = time.perf_counter()
        wsize, step, count = scoring.analysis_windows(n, None, 0.2, None, 0.1)
        stats, runs, nr = scoring.threshold_windows(final, wsize, step, count, 0, 200); t3 = time.perf_counter()
        merged = scoring.intervals_from_runs(stats, runs, nr, step, 0.1); t4 = time.perf_counter()
    print("rows %d: forward %.2f ms, kde+finish %.2f ms, threshold kernels+copy %.2f ms (windows %d, max runs %d, total runs %d), host tail %.2f ms (%d intervals)"
          % (n, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), count, nr.max(), nr.sum(), 1e3 * (t4 - t3), len(merged)))
