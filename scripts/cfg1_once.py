"""BASELINE config 1 (T = 8640, hyperbolic, uncertainty, intervals) scored a few times: the target of the ncu launch list that
shows where the 34 launches of a short signal spend their device time."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hypad_b200.models.tadgan import CriticX, Decoder, Encoder
from hypad_b200.scoring import WindowScorer

dev = torch.device("cuda", 0)
torch.manual_seed(0)
sc = WindowScorer(Encoder(100, 20).eval().to(dev), Decoder(100, 20, True).eval().to(dev), CriticX(100, 20).eval().to(dev))
T = 8640
t = np.arange(T)
s = np.sin(2 * np.pi * t / 50.0)
s[T // 2:T // 2 + 5] += 3
x = torch.from_numpy(2 * (s - s.min()) / (s.max() - s.min()) - 1).to(dev)
idx = 1285027200 + 21600 * t
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    out = sc.score(x, sliding=True, combination="uncertainty", index=idx)
torch.cuda.synchronize()
print(len(out["intervals"]), "intervals")
