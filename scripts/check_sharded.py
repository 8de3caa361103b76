"""Run under torchrun (one process per GPU): the window-sharded scoring of a golden case must equal the single-GPU run
bit for bit (same kernels, same per-window arithmetic; only the gather is added).  Rank 0 prints the verdict."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from conftest import build_modules, full_signal, golden
from hypad_b200.distributed import ShardedScorer
from hypad_b200.scoring import WindowScorer


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    from hypad_b200.distributed import TorchComm

    comm = TorchComm()
    if rank == 0:
        print("stage exchanges: %s" % ("NVLink peer memory (hypad_peer_exchange)" if comm.peer is not None else "NCCL all-gather"))
    if comm.peer is not None:  # the exchange itself: every rank's record arrives intact, in rank order, over many rounds
        for rnd in range(40):
            rec = torch.full((1024 + 8 * (rnd % 5),), float(rank * 1000 + rnd), dtype=torch.float64, device=dev)
            g = comm.all_gather(rec)
            want = torch.tensor([r * 1000.0 + rnd for r in range(world)], dtype=torch.float64, device=dev)
            good = bool((g == want[:, None]).all().item()) and g.shape == (world, rec.numel())
            ok = ok and good
        comm.peer.poll()
        t = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
        if rank == 0:
            print("peer exchange, 40 rounds of varying size: %s" % ok)
    for case in ("cfg1_hyp_uncertainty.npz", "noisy1500_hyp_uncertainty.npz", "edge_n300_hyp.npz", "long200k"):
        enc, dec, cx, _ = build_modules("weights_hyp_s100.npz", 100, True, dev)
        scorer = WindowScorer(enc, dec, cx)
        if case == "long200k":  # the staged statistics at a size where halos and quantile ranks cross rank boundaries
            from conftest import long_signal

            sig = long_signal(200000)
            g = {"index": np.arange(200000)}
        else:
            g = golden(case)
            sig = full_signal(g)
        n = sig.shape[0] - 100
        sh = ShardedScorer(scorer)
        first, count, h0, lo, hi = sh.plan(n)
        local_slice = torch.from_numpy(sig[lo:hi].copy()).to(dev)
        out = sh.gather_full(sh.score_hyperbolic(local_slice, n, "uncertainty", index=g["index"]), n)
        ref = scorer.score(torch.from_numpy(sig).to(dev), True, "uncertainty", index=g["index"])
        same = all(torch.equal(out[k], ref[k]) for k in ("final", "kmax", "rec", "unorm", "critic_scores"))
        same = same and torch.equal(out["final_local"], ref["final"][out["first"]:out["first"] + out["count"]])
        same = same and out["intervals"].shape == ref["intervals"].shape and np.array_equal(out["intervals"], ref["intervals"])
        t = torch.tensor([int(same)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("%s: world=%d sharded == single-GPU: %s" % (case, world, bool(t.item())))
        ok = ok and bool(t.item())
    # BASELINE config 4: multivariate rows (S = C = 123) sharded by row range
    enc, dec, cx, _ = build_modules("weights_hyp_s123.npz", 123, True, dev)
    scorer = WindowScorer(enc, dec, cx)
    rows = np.random.default_rng(33).uniform(-1, 1, (4000, 123))
    rows[1900:1910] *= 3
    index = 1353715200.0 + np.arange(4000)
    sh = ShardedScorer(scorer)
    first, count, h0, lo, hi = sh.plan_rows(4000)
    out = sh.gather_full(sh.score_multivariate(torch.from_numpy(rows[lo:hi].copy()).to(dev), 4000, "mult", index=index), 4000)
    ref = scorer.score(torch.from_numpy(rows).to(dev), False, "mult", index=index, multivariate=True)
    same = all(torch.equal(out[k], ref[k]) for k in ("final", "kmax", "rec")) and np.array_equal(out["intervals"], ref["intervals"])
    t = torch.tensor([int(same)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multivariate rows S=123: world=%d sharded == single-GPU: %s" % (world, bool(t.item())))
    ok = ok and bool(t.item())
    # the per-timestep Euclidean path (BASELINE config 2's model) sharded by window range
    enc, dec, cx, _ = build_modules("weights_eucl_s100.npz", 100, False, dev)
    scorer = WindowScorer(enc, dec, cx)
    g = golden("cfg2_eucl_dtw_mult.npz")
    sig = full_signal(g)
    n = sig.shape[0] - 100
    sh = ShardedScorer(scorer)
    t0, cnt, w_lo, w_hi, lo, hi = sh.plan_euclidean(n)
    out = sh.score_euclidean(torch.from_numpy(sig[lo:hi].copy()).to(dev), n, "mult", "dtw", index=g["index"])
    ref = scorer.score(torch.from_numpy(sig).to(dev), True, "mult", "dtw", index=g["index"])
    same = torch.equal(out["final"], ref["final"]) and torch.equal(out["final_local"], ref["final"][t0:t0 + cnt])
    same = same and torch.equal(out["rec_local"], ref["rec"][t0:t0 + cnt]) and np.array_equal(out["intervals"], ref["intervals"])
    t = torch.tensor([int(same)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("euclidean dtw mult (cfg2 signal): world=%d sharded == single-GPU: %s" % (world, bool(t.item())))
    ok = ok and bool(t.item())
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_CHECK", "OK" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
