"""Under torchrun: CUDA-event time of every phase of ShardedScorer.score_hyperbolic on the bench workload (1M timesteps per GPU),
max over ranks.  Shows where a multi-GPU step spends what the single-GPU step does not."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from bench import S, make_signal
from hypad_b200 import scoring
from hypad_b200.distributed import ShardedScorer
from hypad_b200.models.tadgan import CriticX, Decoder, Encoder


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    enc, dec, cx = Encoder(S, 20), Decoder(S, 20, True), CriticX(S, 20)
    for m in (enc, dec, cx):
        m.to(dev).eval()
    sc = scoring.WindowScorer(enc, dec, cx)
    sh = ShardedScorer(sc)
    T = 1000000 * world
    n = T - S
    sig = make_signal(T)
    index = np.arange(T)
    first, count, h0, lo, hi = sh.plan(n)
    x = torch.from_numpy(sig[lo:hi].copy()).to(dev)
    names = ["forward", "kde", "critic_staged", "combine", "find_anomalies_sharded"]
    acc = {k: 0.0 for k in names}
    reps = 8
    for it in range(reps + 3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        fw = sc.forward(x, True)
        ev[1].record()
        ranges = sh.position_ranges(n)
        t0, tc = ranges[rank]
        kmax = scoring.kde_argmax_overlap(fw["critic"], S, n_windows=n, critic_offset=h0, t0=t0, t_count=tc)
        ev[2].record()
        cs = scoring.critic_scores_staged(kmax, ranges, n + S - 1, math.trunc(n * 0.01), sh.comm)
        ev[3].record()
        lead = first - h0
        final = scoring.combine("uncertainty", cs[:count], fw["rec"][lead:], fw["unorm"][lead:], n=count)
        ev[4].record()
        iv = sh.find_anomaly_intervals_sharded(final, sh.ranges(n), n, index, 0.33, 0.1, anomaly_padding=50, ddof=1)
        ev[5].record()
        torch.cuda.synchronize()
        if it >= 3:
            for i, k in enumerate(names):
                acc[k] += ev[i].elapsed_time(ev[i + 1]) / reps
    t = torch.tensor([acc[k] for k in names], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "phase_ms_max_over_ranks": dict(zip(names, [round(v, 4) for v in t.tolist()])),
                          "total_ms": round(float(t.sum()), 4), "intervals": int(len(iv))}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
