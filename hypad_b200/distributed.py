"""Sharding of one signal's windows across GPUs: one process per GPU, a single gather at the end.

The scoring path is embarrassingly parallel per window; the only coupling is the overlap aggregation, where
timestep i needs the critic value of windows i-S+1 .. i.  Rank r therefore
  * owns the contiguous window range [first, first+count) and the timesteps with the same indices (the last
    rank also owns the S-1 trailing timesteps),
  * recomputes the S-1 windows to the left of its range (the "halo": 0.01 % extra work at 1M windows/GPU),
  * runs the fused network + KDE arg-max on its range with no communication,
and the per-timestep / per-window arrays (kmax, rec, unorm, 12 B per timestep as fp32) are gathered in one NCCL all-gather
(NVLink 5 / NVSwitch).  The elementwise O(T) finish (quantile band, z-score, smoothing, combine)
is then run redundantly on every rank; the analysis windows of find_anomalies -- each a third of the signal, twenty-odd
of them, the one part of the finish whose cost grows with the total length -- are dealt out to the ranks and their few
runs gathered (a few KB).  There is no collective inside a kernel.
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import scoring
from ._native import HypadError


def shard_ranges(n_windows, world_size):
    """Balanced contiguous window ranges: [(first, count)] * world_size (counts differ by at most one)."""
    base, extra = divmod(n_windows, world_size)
    out, first = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


def timestep_range(first, count, n_windows, S, is_last):
    """Timesteps a rank aggregates: its window indices, plus the S-1 trailing ones on the last rank."""
    return first, count + (S - 1 if is_last else 0)


def halo_first(first, S):
    return max(0, first - (S - 1))


def _all_gather_padded(local, width, group=None, async_op=False):
    """One all_gather_into_tensor of `local` zero-padded to `width` elements.  Returns (out (world, width), work)."""
    world = dist.get_world_size(group)
    if local.shape[0] == width:
        buf = local.contiguous()
    else:
        buf = local.new_zeros(width)
        buf[: local.shape[0]] = local
    out = local.new_empty(world * width)  # flat: gloo accepts only the concatenated 1-D layout
    work = dist.all_gather_into_tensor(out, buf, group=group, async_op=async_op)
    return out.view(world, width), work


def _concat_rows(out, sizes):
    """Rows of a padded gather result cut to their true lengths and concatenated (a view when nothing was padded)."""
    if all(s == out.shape[1] for s in sizes):
        return out.reshape(-1)
    return torch.cat([out[r, :s] for r, s in enumerate(sizes)])


def gather_concat(local, sizes, group=None):
    """All-gather variable-length 1-D tensors (sizes known on every rank) into one concatenated tensor.

    Works on CUDA tensors with NCCL and on CPU tensors with gloo (used by the world_size-2 CPU tests)."""
    out, _ = _all_gather_padded(local, max(sizes), group)
    return _concat_rows(out, sizes)


class ShardedScorer:
    """WindowScorer over torch.distributed.  Each rank holds only its slice of the signal (`local_slice`)."""

    def __init__(self, scorer, group=None, rank=None, world=None):
        """rank / world default to the process group's; passing both makes an object that plans and packs for that rank
        without touching torch.distributed (tests replay all ranks of a sharded run on one GPU with it)."""
        self.scorer = scorer
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world

    def plan(self, n_windows):
        """(first, count, h0, sample_lo, sample_hi): owned windows, first halo window and the sample range
        [sample_lo, sample_hi) of the signal this rank needs resident (sliding windows)."""
        S = self.scorer.S
        first, count = shard_ranges(n_windows, self.world)[self.rank]
        h0 = halo_first(first, S)
        # window w reads samples [w, w+S); as in the reference (utils/dataloader.py:200) a signal of L samples
        # yields L-S windows, so the slice carries one sample beyond the last window.
        return first, count, h0, h0, first + count + S

    def plan_rows(self, n_rows):
        """Multivariate rows (one row = one window, BASELINE config 4): (first, count, h0, row_lo, row_hi) -- the owned rows, the
        first halo row and the row range [row_lo, row_hi) this rank needs resident.  The S-1 halo rows in front are recomputed
        locally: the critic overlap aggregation of position i reads the critics of rows i-S+1 .. i."""
        first, count = shard_ranges(n_rows, self.world)[self.rank]
        h0 = halo_first(first, self.scorer.S)
        return first, count, h0, h0, first + count

    def pack_local(self, fw, n_windows):
        """This rank's contribution to the gather, from its forward results `fw` (windows h0 .. first+count-1): one fp32 buffer
        [kmax | rec | unorm], each part `width` long.  kmax is one of the fp32 critic values widened to float64, so it travels as
        fp32 without loss; rec and unorm are fp32 anyway.  12 B per position."""
        S = self.scorer.S
        ranges = shard_ranges(n_windows, self.world)
        first, count = ranges[self.rank]
        h0 = halo_first(first, S)
        lead = first - h0
        t0, tc = timestep_range(first, count, n_windows, S, self.rank == self.world - 1)
        kmax_local = scoring.kde_argmax_overlap(fw["critic"], S, n_windows=n_windows, critic_offset=h0, t0=t0, t_count=tc)
        width = self.gather_width(n_windows)
        pack = fw["rec"].new_zeros(3 * width)
        pack[:tc] = kmax_local.float()
        pack[width:width + count] = fw["rec"][lead:]
        pack[2 * width:2 * width + count] = fw["unorm"][lead:]
        return pack

    def gather_width(self, n_windows):
        ranges = shard_ranges(n_windows, self.world)
        return max(timestep_range(f, c, n_windows, self.scorer.S, r == self.world - 1)[1] for r, (f, c) in enumerate(ranges))

    def unpack_gathered(self, flat, n_windows):
        """flat: the world's packs back to back -> (kmax float64 (n+S-1,), rec fp32 (n,), unorm fp32 (n,))."""
        S = self.scorer.S
        ranges = shard_ranges(n_windows, self.world)
        counts = [c for _, c in ranges]
        tcounts = [timestep_range(f, c, n_windows, S, r == self.world - 1)[1] for r, (f, c) in enumerate(ranges)]
        parts = flat.view(self.world, 3, self.gather_width(n_windows))
        return _concat_rows(parts[:, 0, :], tcounts).double(), _concat_rows(parts[:, 1, :], counts), _concat_rows(parts[:, 2, :], counts)

    def finish(self, kmax, rec, unorm, n_windows, combination, multivariate=False):
        """The O(T) finish every rank repeats on the gathered arrays.  multivariate: rec goes through zscore / clip(0) + 1 over
        ALL rows first (utils/anomaly_detection_utils.py:177-178) -- a global statistic, hence after the gather."""
        if multivariate:
            rec = scoring.zscore_clip(rec)
        cs = scoring.critic_zscore_smooth(kmax, math.trunc(n_windows * 0.01))
        final = scoring.combine(combination, cs[:n_windows], rec, unorm, n=n_windows)
        return {"final": final, "kmax": kmax, "rec": rec, "unorm": unorm, "critic_scores": cs[:n_windows]}

    def _gather(self, pack):
        flat = pack.new_empty(self.world * pack.numel())
        dist.all_gather_into_tensor(flat, pack, group=self.group)
        return flat

    def score_hyperbolic(self, local_slice, n_windows, combination="uncertainty", index=None):
        """local_slice: samples [sample_lo, sample_hi) of the scaled signal (see plan()), on this rank's GPU.
        Returns the full-length result on every rank."""
        sc = self.scorer
        first, count, h0, lo, hi = self.plan(n_windows)
        if local_slice.numel() != hi - lo:
            raise ValueError("local slice has %d samples, plan() asks for %d" % (local_slice.numel(), hi - lo))
        fw = sc.forward(local_slice, True)  # windows h0 .. first+count-1
        # One collective for the three per-position arrays
        kmax, rec, unorm = self.unpack_gathered(self._gather(self.pack_local(fw, n_windows)), n_windows)
        out = self.finish(kmax, rec, unorm, n_windows, combination)
        if index is not None:
            out["intervals"] = self.find_anomaly_intervals(out["final"], index, 0.33, 0.1, anomaly_padding=50, ddof=1)
        sc.poll_error()  # after the collectives, so that every rank still reaches them
        return out

    def score_multivariate(self, local_rows, n_rows, combination="mult", index=None):
        """BASELINE config 4: (N, C) rows sharded by contiguous row range.  local_rows: rows [row_lo, row_hi) of plan_rows() on
        this rank's GPU, shape (row_hi - row_lo, C).  Hyperbolic models (the Euclidean multivariate score is a float64 row norm;
        score it unsharded with WindowScorer.score).  Returns the full-length result on every rank; equals
        `WindowScorer.score(rows, sliding=False, multivariate=True)` bit for bit."""
        sc = self.scorer
        if not sc.hyperbolic:
            raise HypadError("hypad_b200: ShardedScorer.score_multivariate needs a hyperbolic model")
        first, count, h0, lo, hi = self.plan_rows(n_rows)
        if local_rows.dim() != 2 or local_rows.shape[0] != hi - lo:
            raise ValueError("local rows have shape %s, plan_rows() asks for %d rows" % (tuple(local_rows.shape), hi - lo))
        fw = sc.forward(local_rows, False)  # rows h0 .. first+count-1
        kmax, rec, unorm = self.unpack_gathered(self._gather(self.pack_local(fw, n_rows)), n_rows)
        out = self.finish(kmax, rec, unorm, n_rows, combination, multivariate=True)
        if index is not None:
            out["intervals"] = self.find_anomaly_intervals(out["final"], index, 0.2, 0.1, anomaly_padding=200, ddof=0)
        sc.poll_error()
        return out

    MAX_RUNS = 64  # per analysis window in the gathered buffer; more (never seen) -> every rank redoes all windows

    def find_anomaly_intervals(self, final, index, window_size_portion, window_step_size_portion, min_percent=0.1,
                               anomaly_padding=50, ddof=0):
        """scoring.find_anomaly_intervals with the analysis windows dealt out to the ranks: rank r thresholds windows
        [r*per, (r+1)*per) of the (identical, gathered) score array, the packed per-window results are all-gathered and the
        host tail (prune, score, merge) runs on every rank.  The kernels work on the whole array with a first-window index, so
        the block sums and the shift sample are the single-GPU ones and the result is bitwise the single-GPU one."""
        n = final.numel()
        wsize, step, count = scoring.analysis_windows(n, None, window_size_portion, None, window_step_size_portion)
        per = -(-count // self.world)
        k0 = min(self.rank * per, count)
        kc = max(0, min(per, count - k0))
        R = self.MAX_RUNS
        blen = scoring.threshold_buffer_len(per, R)
        local = torch.zeros(blen, dtype=torch.float64, device=final.device)
        if kc > 0:
            sub = scoring.threshold_windows_launch(final, wsize, step, kc, ddof, anomaly_padding, R, first_window=k0)
            # re-pack the kc-window buffer into the fixed per-window layout of `per` windows
            s_loc, r_loc, n_loc = scoring.threshold_buffer_len(kc, R), kc * 4, kc * R * 3
            local[: kc * 4] = sub[:r_loc]
            local[per * 4: per * 4 + n_loc] = sub[r_loc:r_loc + n_loc]
            local[per * 4 + per * R * 3: per * 4 + per * R * 3 + (kc + 1) // 2] = sub[r_loc + n_loc:s_loc]
        parts = [torch.empty_like(local) for _ in range(self.world)]
        dist.all_gather(parts, local, group=self.group)
        host = torch.stack(parts).cpu().numpy()
        stats, runs, n_runs = [], [], []
        for r in range(self.world):
            rk0 = min(r * per, count)
            rkc = max(0, min(per, count - rk0))
            st, ru, nr = scoring.threshold_windows_parse(host[r], per, R)
            stats.append(st[:rkc])
            runs.append(ru[:rkc])
            n_runs.append(nr[:rkc])
        stats, runs, n_runs = np.concatenate(stats), np.concatenate(runs), np.concatenate(n_runs)
        if n_runs.max(initial=0) > R:  # same decision on every rank: the gathered counts are identical
            return scoring.find_anomaly_intervals(final, index, window_size_portion, window_step_size_portion,
                                                  min_percent=min_percent, anomaly_padding=anomaly_padding, ddof=ddof)
        merged = scoring.intervals_from_runs(stats, runs, n_runs, step, min_percent)
        return scoring.intervals_to_index(merged, index)
