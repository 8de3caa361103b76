"""BASELINE config 5 under torchrun (or alone): ~500 NAB / NASA / YAHOO-shaped synthetic signals, one random-init HypAD model per
signal, sharded BY SIGNAL over the ranks (hypad_b200.sweep.SignalSweep: no halo, no collective on the data path, one
all_gather_object of the interval lists).  Every signal carries one injected burst, labelled, so the sweep also reports the
reference's overlap-segment confusion counts / F1 per signal and in total.  Wall clock around whole sweeps (launch- and
host-bound work), max over ranks; rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hypad_b200.models.tadgan import CriticX, Decoder, Encoder
from hypad_b200.scoring import WindowScorer
from hypad_b200.sweep import SignalSweep

T0, DT = 1285027200, 21600


def signal(T, seed):
    t = np.arange(T)
    s = np.sin(2 * np.pi * t / 50.0) + 0.05 * np.random.default_rng(seed).standard_normal(T)
    s[T // 2:T // 2 + 5] += 3
    return 2 * (s - s.min()) / (s.max() - s.min()) - 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--streams", type=int, default=8)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(5)
    lengths = np.concatenate([rng.integers(2000, 8700, 80), rng.integers(1100, 22700, 46), rng.integers(1420, 1700, 367)]).tolist()
    signals = [signal(t, i) for i, t in enumerate(lengths)]
    indices = [T0 + DT * np.arange(t) for t in lengths]
    known = [[(T0 + DT * (t // 2), T0 + DT * (t // 2 + 4))] for t in lengths]
    total = sum(t - 100 for t in lengths)
    cache = {}

    def scorer(i):  # one model per signal, built on the rank that scores it
        if i not in cache:
            torch.manual_seed(1000 + i)
            enc, dec, cx = Encoder(100, 20), Decoder(100, 20, True), CriticX(100, 20)
            cache[i] = WindowScorer(enc.eval().to(dev), dec.eval().to(dev), cx.eval().to(dev))
        return cache[i]

    sw = SignalSweep(scorer, streams=args.streams)
    out, ev = sw.run(signals, indices, known_anomalies=known)  # builds and packs the models of this rank
    times = []
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, ev = sw.run(signals, indices, known_anomalies=known)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": "cfg5: %d signals (lengths 1100..22700), one random-init HypAD model per signal, sharded by signal" % len(lengths),
                          "n_gpus": world, "signals": len(lengths), "windows": total, "ms": float(t.item()) * 1e3,
                          "windows_per_s": total / float(t.item()), "streams_per_gpu": args.streams,
                          "signals_with_intervals": int(sum(len(v) > 0 for v in out.values())),
                          "evaluation": {k: ev[k] for k in ("tp", "fp", "fn", "precision", "recall", "f1")}}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
