"""BASELINE config 4 under torchrun: HypAD multivariate rows (S = C = 123, synthetic), sharded by row range over the GPUs of
one node (`ShardedScorer.score_multivariate`), weak scaling at 2^20 rows per GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/cfg4_multivariate_sharded.py [--rows-per-gpu R] [--steps K] [--warmup W]

Same timing discipline as bench.py (W warm-up steps, K timed steps each bracketed by CUDA events on the launching stream, a 512 MiB
L2 flush between steps, barrier + synchronize around the region, max over ranks); rank 0 prints one JSON line.  Not the bench
arm -- bench.py measures config 3 -- but the same metric on the multivariate shape."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hypad_b200.distributed import ShardedScorer
from hypad_b200.models.tadgan import CriticX, Decoder, Encoder
from hypad_b200.scoring import WindowScorer

S = 123


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)  # NCCL's banner goes to stderr; stdout carries the JSON line only
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.manual_seed(0)
    enc, dec, cx = Encoder(S, 20), Decoder(S, 20, True), CriticX(S, 20)
    scorer = WindowScorer(enc.eval().to(dev), dec.eval().to(dev), cx.eval().to(dev))
    n_rows = args.rows_per_gpu * world
    index = 1353715200.0 + np.arange(n_rows)
    if world > 1:
        sh = ShardedScorer(scorer)
        first, count, h0, lo, hi = sh.plan_rows(n_rows)
    else:
        sh, lo, hi = None, 0, n_rows
    # this rank's rows (with its halo); synthetic U(-1, 1) with one damped burst -- timing only, parity is tests' business
    g = torch.Generator(device=dev).manual_seed(4)
    rows = torch.rand(hi - lo, S, dtype=torch.float32, device=dev, generator=g) * 2 - 1
    if lo <= n_rows // 3 < hi - 300:
        rows[n_rows // 3 - lo:n_rows // 3 - lo + 300] *= 0.2
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def step():
        if sh is not None:
            return sh.score_multivariate(rows, n_rows, "mult", index=index)
        return scorer.score(rows, sliding=False, combination="mult", index=index, multivariate=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step()
    barrier()
    total = 0.0
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = step()
        b.record()
        b.synchronize()
        total += a.elapsed_time(b)
    barrier()
    t = torch.tensor([total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item()) / args.steps
        print(json.dumps({"metric": "windows_scored_per_sec", "config": "cfg4: HypAD multivariate, S=C=123, %d rows per GPU, hyperbolic, mult, "
                          "rows sharded by contiguous range" % args.rows_per_gpu, "n_gpus": world, "rows": n_rows, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms, "value": n_rows / ms * 1e3, "unit": "windows/s", "scaling": "weak",
                          "intervals": int(len(out["intervals"])), "l2": "flushed between timed steps (512 MiB device write)"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
