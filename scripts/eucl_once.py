"""Euclidean path (TadGAN, rec_error dtw, combination mult) at the bench length, a few times: target of ncu launch lists / captures."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import make_signal
from hypad_b200.models.tadgan import CriticX, Decoder, Encoder
from hypad_b200.scoring import WindowScorer

dev = torch.device("cuda", 0)
torch.manual_seed(0)
sc = WindowScorer(Encoder(100, 20).eval().to(dev), Decoder(100, 20, False).eval().to(dev), CriticX(100, 20).eval().to(dev))
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
x = torch.from_numpy(make_signal(T)).to(dev)
idx = np.arange(T)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    out = sc.score(x, sliding=True, combination="mult", rec_error_type="dtw", index=idx)
torch.cuda.synchronize()
print(len(out["intervals"]), "intervals")
