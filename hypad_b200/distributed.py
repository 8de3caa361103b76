"""Sharding of one signal's windows across GPUs: one process per GPU, small exchanges between the stages of the finish.

The scoring path is embarrassingly parallel per window; the couplings are
  * the overlap aggregation, where timestep i needs the critic value of windows i-S+1 .. i: rank r owns the contiguous window
    range [first, first+count) and the timesteps with the same indices (the last rank also the S-1 trailing ones) and
    recomputes the S-1 windows to the left of its range (the "halo": 0.01 % extra work at 1M windows/GPU);
  * the global statistics of the finish -- the 25 % / 75 % quantiles, band mean and std of the KDE selections
    (utils/anomaly_detection_utils.py:319-325), the z-score moments of the multivariate reconstruction error (:177) -- taken
    in stages (csrc/critic_stats.cu): every rank reduces its slice to a small record (a digit histogram of the radix select,
    a few (hi, lo) partial sums), the records are all-gathered (NCCL over NVLink 5 / NVSwitch, a few KB) and every rank folds
    them in rank order.  Three histogram exchanges and one for the sums;
  * the smoothing (:326-331), a centred rolling mean of 1 % of the signal: the first and last window/2 selections of every
    rank ride on the first histogram exchange and give the neighbours their halo;
  * find_anomalies (:1363-1472), whose analysis windows each span a third of the signal: the per-window final scores
    (8 B per window) are all-gathered once, the twenty-odd analysis windows dealt out to the ranks and their few runs gathered
    (a few KB).
Every rank keeps and returns ITS slice of the per-window results (`final_local`, ...); nothing O(total length) is computed
twice.  The partial sums are carried as unevaluated double pairs (csrc/dd.cuh), so the statistics do not depend on how many
ranks took part: a sharded run equals the single-GPU run bit for bit (tests/test_gpu_parity.py, scripts/check_sharded.py).
There is no collective inside a kernel.
"""
import math
import os

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


BLOCK = 1024  # positions per block of the thresholding summaries (csrc/finish.cu: TW_TILE)


def shard_ranges_aligned(n_windows, world_size, block=BLOCK):
    """Contiguous ranges that start at multiples of `block`: the blocks are dealt out evenly, in order (counts differ by at most
    one block; the last range ends with the array).  What the sharded find_anomalies needs: a block never straddles two ranks."""
    nb = -(-n_windows // block)
    base, extra = divmod(nb, world_size)
    out, b = [], 0
    for r in range(world_size):
        cnt_b = base + (1 if r < extra else 0)
        first = min(b * block, n_windows)
        end = min((b + cnt_b) * block, n_windows)
        out.append((first, end - first))
        b += cnt_b
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


def _pad_to(local, width):
    if local.shape[0] == width:
        return local.contiguous()
    buf = local.new_zeros(width)
    buf[: local.shape[0]] = local
    return buf


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


class PeerExchange:
    """All-gather of small records through NVLink peer memory (csrc/peer.cu): one kernel launch per exchange, no host-side
    collective.  Owns a symmetric buffer (torch.distributed._symmetric_memory) of `slots` data areas + flag rows; exchanges are
    numbered and rotate through the slots.  Raises at construction when symmetric memory is unavailable (the caller then stays
    with NCCL)."""

    def __init__(self, group=None, slots=8, slot_bytes=1 << 20):
        import torch.distributed._symmetric_memory as symm

        from . import _native

        self.lib = _native.load_library()
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.slots, self.slot_bytes = slots, slot_bytes * self.world  # room for `world` records of slot_bytes each
        self.flag_off = self.slots * self.slot_bytes
        total = self.flag_off + self.slots * self.world * 8
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm.empty(total, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group.group_name)
        self.bases = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=dev)
        self.error = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        self.handle.barrier()  # every rank's flags are zero before the first exchange
        self.counter = 0

    def fits(self, nbytes):
        return nbytes % 8 == 0 and 0 < nbytes * self.world <= self.slot_bytes

    def all_gather(self, flat_u8):
        """flat_u8: contiguous uint8 CUDA tensor (a multiple of 8 bytes, the same size on every rank).  Returns a (world, nbytes)
        uint8 view of the local buffer, valid until `slots` further exchanges have been issued."""
        from ._native import check

        nbytes = flat_u8.numel()
        slot, seq = self.counter % self.slots, self.counter // self.slots + 1
        self.counter += 1
        data_off = slot * self.slot_bytes
        flag_off = self.flag_off + slot * self.world * 8
        check(self.lib.hypad_peer_exchange(flat_u8.data_ptr(), nbytes, self.bases.data_ptr(), self.rank, self.world, data_off, flag_off,
                                           seq, self.error.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return self.buf[data_off: data_off + self.world * nbytes].view(self.world, nbytes)

    def poll(self):
        if int(self.error.item()):
            raise HypadError("hypad_b200: a peer's record did not arrive within the exchange time-out")


class TorchComm:
    """The stage exchange over torch.distributed: small records through NVLink peer memory when the node offers symmetric memory
    (PeerExchange: one kernel launch per exchange), anything else -- and everything when it does not -- through
    all_gather_into_tensor."""

    def __init__(self, group=None, peer=True):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.peer = None
        if peer and self.world > 1 and dist.get_backend(group) == "nccl" and os.environ.get("HYPAD_PEER_EXCHANGE", "1") != "0":
            # every rank must take the same path: the outcome of the set-up is agreed on before it is used
            ok = torch.ones(1, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
            px = None
            try:
                px = PeerExchange(group)
            except Exception:  # noqa: BLE001 -- no symmetric memory on this node / build: NCCL carries the exchanges
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            self.peer = px if int(ok.item()) else None

    def all_gather(self, buf):
        flat = buf.reshape(-1)
        if self.peer is not None and flat.is_cuda and flat.is_contiguous():
            nbytes = flat.numel() * flat.element_size()
            if self.peer.fits(nbytes):  # a function of the size alone: every rank takes the same path
                if flat.data_ptr() % 16:
                    flat = flat.clone()  # a view into the middle of an array: the copy kernel wants an aligned source
                return self.peer.all_gather(flat.view(torch.uint8)).view(flat.dtype).view(self.world, -1)
        out = flat.new_empty(self.world * flat.numel())
        dist.all_gather_into_tensor(out, flat, group=self.group)
        return out.view(self.world, -1)


class ShardedScorer:
    """WindowScorer over torch.distributed.  Each rank holds only its slice of the signal (`local_slice`)."""

    def __init__(self, scorer, group=None, rank=None, world=None, comm=None):
        """rank / world default to the process group's; passing both makes an object that plans and packs for that rank
        without touching torch.distributed; `comm` (rank, world, all_gather) then carries the stage exchanges -- tests replay all
        ranks of a sharded run on one GPU, one thread per rank, with it."""
        self.scorer = scorer
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.comm = comm if comm is not None else (TorchComm(group) if rank is None else None)

    def ranges(self, n):
        """[(first, count)] of every rank.  Long arrays are dealt out in whole blocks of 1024 positions (the sharded find_anomalies
        works on block summaries); short ones are balanced to the window and thresholded from one gathered copy."""
        if n >= self.world * 8 * BLOCK:
            return shard_ranges_aligned(n, self.world)
        return shard_ranges(n, self.world)

    def plan(self, n_windows):
        """(first, count, h0, sample_lo, sample_hi): owned windows, first halo window and the sample range
        [sample_lo, sample_hi) of the signal this rank needs resident (sliding windows)."""
        S = self.scorer.S
        first, count = self.ranges(n_windows)[self.rank]
        h0 = halo_first(first, S)
        # window w reads samples [w, w+S); as in the reference (utils/dataloader.py:200) a signal of L samples
        # yields L-S windows, so the slice carries one sample beyond the last window.
        return first, count, h0, h0, first + count + S

    def plan_rows(self, n_rows):
        """Multivariate rows (one row = one window, BASELINE config 4): (first, count, h0, row_lo, row_hi) -- the owned rows, the
        first halo row and the row range [row_lo, row_hi) this rank needs resident.  The S-1 halo rows in front are recomputed
        locally: the critic overlap aggregation of position i reads the critics of rows i-S+1 .. i."""
        first, count = self.ranges(n_rows)[self.rank]
        h0 = halo_first(first, self.scorer.S)
        return first, count, h0, h0, first + count

    def position_ranges(self, n_windows):
        """[(first position, count)] of every rank in the kmax array (n_windows + S - 1 positions)."""
        S = self.scorer.S
        return [timestep_range(f, c, n_windows, S, r == self.world - 1) for r, (f, c) in enumerate(self.ranges(n_windows))]

    def _gather_windows(self, local, n_windows):
        """Per-window array of the own windows -> the full array on every rank (one all-gather)."""
        counts = [c for _, c in self.ranges(n_windows)]
        g = self.comm.all_gather(_pad_to(local, max(counts)))
        return _concat_rows(g, counts)

    def _score(self, fw, n, combination, index, multivariate, portion, padding, ddof, out_host=None, stats_f32=False):
        sc, S = self.scorer, self.scorer.S
        first, count = self.ranges(n)[self.rank]
        lead = first - halo_first(first, S)
        ranges = self.position_ranges(n)
        t0, tc = ranges[self.rank]
        kmax = scoring.kde_argmax_overlap(fw["critic"], S, n_windows=n, critic_offset=halo_first(first, S), t0=t0, t_count=tc)
        rec, unorm = fw["rec"][lead:], fw["unorm"][lead:]
        if multivariate:  # utils/anomaly_detection_utils.py:177-178: a statistic of ALL rows
            rec = scoring.zscore_clip_staged(rec, n, self.comm)
        cs = scoring.critic_scores_staged(kmax, ranges, n + S - 1, math.trunc(n * 0.01), self.comm)
        final = scoring.combine(combination, cs[:count], rec, unorm, n=count)
        out = {"first": first, "count": count, "final_local": final, "kmax_local": kmax, "rec_local": rec, "unorm_local": unorm,
               "critic_scores_local": cs[:count]}
        if out_host is not None:  # this rank's scores travel to its host while the interval extraction still runs
            scoring.download_async(sc, final, out_host)
        if index is not None:
            iv = self.find_anomaly_intervals_sharded(final, self.ranges(n), n, index, portion, 0.1, anomaly_padding=padding, ddof=ddof,
                                                     stats_f32=stats_f32)
            if iv is None:  # short array: one gathered copy, the analysis windows dealt out
                out["final"] = self._gather_windows(final, n)
                iv = self.find_anomaly_intervals(out["final"], index, portion, 0.1, anomaly_padding=padding, ddof=ddof, stats_f32=stats_f32)
            out["intervals"] = iv
        if out_host is not None:
            sc._down_stream.synchronize()
        sc.poll_error()  # after the collectives, so that every rank still reaches them
        if getattr(self.comm, "peer", None) is not None:
            self.comm.peer.poll()
        return out

    def gather_full(self, out, n):
        """The full-length arrays of a sharded result on every rank (tests, callers that want them): final, kmax, rec, unorm."""
        res = dict(out)
        if "final" not in res:
            res["final"] = self._gather_windows(out["final_local"], n)
        res["rec"] = self._gather_windows(out["rec_local"], n)
        res["unorm"] = self._gather_windows(out["unorm_local"], n)
        res["critic_scores"] = self._gather_windows(out["critic_scores_local"], n)
        pc = [c for _, c in self.position_ranges(n)]
        res["kmax"] = _concat_rows(self.comm.all_gather(_pad_to(out["kmax_local"], max(pc))), pc)
        return res

    def score_hyperbolic(self, local_slice, n_windows, combination="uncertainty", index=None, out_host=None):
        """local_slice: samples [sample_lo, sample_hi) of the scaled signal (see plan()), on this rank's GPU -- or in pinned host
        memory: then it is uploaded chunk by chunk under the network (WindowScorer.forward_from_host).  out_host: pinned float64
        host tensor that receives `final_local`.
        Returns this rank's slice of the per-window results (`final_local`, `kmax_local`, `rec_local`, `unorm_local`,
        `critic_scores_local`, with `first` / `count`); when `index` is given also the gathered `final` and the intervals, the
        same on every rank."""
        first, count, h0, lo, hi = self.plan(n_windows)
        if local_slice.numel() != hi - lo:
            raise ValueError("local slice has %d samples, plan() asks for %d" % (local_slice.numel(), hi - lo))
        ddof, stats_f32 = scoring.univariate_hyperbolic_semantics(combination)  # what find_anomalies is handed
        if local_slice.is_cuda:
            fw = self.scorer.forward(local_slice, True)  # windows h0 .. first+count-1
        else:
            fw = self.scorer.forward_from_host(local_slice)
        return self._score(fw, n_windows, combination, index, False, 0.33, 50, ddof, out_host=out_host, stats_f32=stats_f32)

    def score_multivariate(self, local_rows, n_rows, combination="mult", index=None):
        """BASELINE config 4: (N, C) rows sharded by contiguous row range.  local_rows: rows [row_lo, row_hi) of plan_rows() on
        this rank's GPU, shape (row_hi - row_lo, C).  Hyperbolic models (the Euclidean multivariate score is a float64 row norm;
        score it unsharded with WindowScorer.score).  Same result layout as score_hyperbolic; equals
        `WindowScorer.score(rows, sliding=False, multivariate=True)` bit for bit."""
        sc = self.scorer
        if not sc.hyperbolic:
            raise HypadError("hypad_b200: ShardedScorer.score_multivariate needs a hyperbolic model")
        first, count, h0, lo, hi = self.plan_rows(n_rows)
        if local_rows.dim() != 2 or local_rows.shape[0] != hi - lo:
            raise ValueError("local rows have shape %s, plan_rows() asks for %d rows" % (tuple(local_rows.shape), hi - lo))
        fw = sc.forward(local_rows, False)  # rows h0 .. first+count-1
        return self._score(fw, n_rows, combination, index, True, 0.2, 200, 0)

    # Positions either side that the windowed reconstruction errors reach (score_window 10): the area error looks 5 either way; the
    # DTW error of position i is the distance of the windows over [i - 10, i] (the reference stores window c of the zero-padded
    # arrays at position c + 5, utils/anomaly_detection_utils.py:834-861) and the last 6 positions of the array are zero -- an end
    # effect that must stay outside the owned positions wherever the slice end is not the signal's end.
    ERR_HALO = 10

    def plan_euclidean(self, n_windows):
        """(t0, tc, w_lo, w_hi, sample_lo, sample_hi) for the Euclidean per-timestep path: the owned positions [t0, t0 + tc) of the
        n_windows + S - 1 timesteps, the windows [w_lo, w_hi) whose reconstructions this rank computes -- every window under a
        position it needs the prediction of, i.e. the owned positions and ERR_HALO more on either side -- and the samples
        [sample_lo, sample_hi) of the signal to keep resident (window w reads [w, w + S))."""
        S, H = self.scorer.S, self.ERR_HALO
        t0, tc = self.position_ranges(n_windows)[self.rank]
        w_lo = max(0, t0 - H - (S - 1))
        w_hi = min(n_windows, t0 + tc + H)
        return t0, tc, w_lo, w_hi, w_lo, w_hi + S

    def score_euclidean(self, local_slice, n_windows, combination="mult", rec_error_type="dtw", index=None, lambda_rec=0.5):
        """The per-timestep Euclidean path (score_anomalies, utils/anomaly_detection_utils.py:407-576) of a signal whose windows
        are sharded over the ranks.  local_slice: the samples [sample_lo, sample_hi) of plan_euclidean() on this rank's GPU.
        Halos: S - 1 windows to the left for the overlap aggregations (KDE arg-max of the critics, median of the
        reconstructions), ERR_HALO positions either side for the windowed reconstruction error; the smoothing halo and the
        global z-score / critic statistics travel in small all-gathers (scoring.*_staged).  Returns this rank's positions
        (`final_local`, `rec_local`, `critic_scores_local`, `kmax_local`, `pred_local`, `t0`, `count`) and, with `index`, the
        gathered `final` and the intervals.  Bitwise WindowScorer.score(signal, True, combination, rec_error_type)."""
        sc, S, H = self.scorer, self.scorer.S, self.ERR_HALO
        if sc.hyperbolic:
            raise HypadError("hypad_b200: ShardedScorer.score_euclidean needs a Euclidean model")
        mode = {"mult": "mult", "sum": "euclidean_sum", "rec": "rec", "critic": "critic"}.get(combination)
        if mode is None:
            raise ValueError('Unknown combination specified {}, use "mult", "sum", or "rec" instead.'.format(combination))
        n = n_windows
        n_pos = n + S - 1
        t0, tc, w_lo, w_hi, lo, hi = self.plan_euclidean(n)
        if local_slice.numel() != hi - lo:
            raise ValueError("local slice has %d samples, plan_euclidean() asks for %d" % (local_slice.numel(), hi - lo))
        ranges = self.position_ranges(n)
        x = local_slice.reshape(-1)
        fw = sc.forward(x, True, keep=("eucl",), count=w_hi - w_lo)  # windows w_lo .. w_hi-1
        kmax = scoring.kde_argmax_overlap(fw["critic"], S, n_windows=n, critic_offset=w_lo, t0=t0, t_count=tc)
        cs = scoring.critic_scores_staged(kmax, ranges, n_pos, math.trunc(n * 0.01), self.comm)
        # prediction and truth on the positions [e_lo, e_hi): the owned ones and the error halo
        e_lo, e_hi = max(0, t0 - H), min(n_pos, t0 + tc + H)
        pred_all = scoring.median_overlap(fw["eucl"])  # local position k <-> global position w_lo + k; complete diagonals only
        pred = pred_all[e_lo - w_lo: e_hi - w_lo]      # inside [w_lo + S - 1, w_hi) or at a global end, which is what is read
        true = x[e_lo - lo: e_hi - lo].double()        # true[t] = signal[t] (:908-910 on sliding windows)
        kind = rec_error_type.lower()
        if kind == "dtw":
            err = scoring.dtw_error(true, pred, 10)
        elif kind == "point":
            err = scoring.point_error(true, pred)
        elif kind == "area":
            err = scoring.area_error(true, pred, 10)
        else:
            raise ValueError("unknown rec_error_type %r" % (rec_error_type,))
        err = err[t0 - e_lo: t0 - e_lo + tc]  # the slice ends' padding only reaches the halo positions
        err = scoring.rolling_mean_staged(err, ranges, n_pos, math.trunc(n * 0.01), self.comm)
        rec = scoring.zscore_clip_staged(err, n_pos, self.comm)
        final = scoring.combine(mode, cs, rec, None, n=tc, lambda_rec=lambda_rec)
        out = {"t0": t0, "count": tc, "final_local": final, "rec_local": rec, "critic_scores_local": cs, "kmax_local": kmax,
               "pred_local": pred[t0 - e_lo: t0 - e_lo + tc], "errors_local": err}
        if index is not None:
            pc = [c for _, c in ranges]
            out["final"] = _concat_rows(self.comm.all_gather(_pad_to(final, max(pc))), pc)
            out["intervals"] = self.find_anomaly_intervals(out["final"], index, 0.33, 0.1, anomaly_padding=50, ddof=0)
        sc.poll_error()
        return out

    def find_anomaly_intervals_sharded(self, final_local, ranges, n, index, window_size_portion, window_step_size_portion,
                                       min_percent=0.1, anomaly_padding=50, ddof=0, stats_f32=False):
        """find_anomalies (fixed threshold) on scores sharded over the ranks by block-aligned ranges, without gathering them:
        every rank reduces its scores to a record (block summaries, the analysis windows' edge elements in its range, its
        boundary values for the neighbours' dilation halo), the records are all-gathered (a few hundred KB), every rank computes
        all windows' statistics from them and the run fragments of its own positions, the fragments are all-gathered and joined
        on the host (csrc/finish.cu, hypad_tw_shard_*).  Same arithmetic on the same values as the single-GPU path: bitwise its
        intervals.  Returns None when the array is too short for this path (ranges not block-aligned, windows under two blocks)."""
        wsize, step, count = scoring.analysis_windows(n, None, window_size_portion, None, window_step_size_portion)
        first, cnt = ranges[self.rank]
        aligned = all(f % BLOCK == 0 and (c % BLOCK == 0 or f + c == n) for f, c in ranges)
        if not aligned or wsize < 2 * BLOCK or anomaly_padding + 1 > min(c for _, c in ranges) or self.world > 64:
            return None
        final_local = final_local.double().contiguous()
        dev = final_local.device
        ctx = scoring._ctx(final_local)  # the stream's own workspace context, like the other finish steps
        lib, h = ctx.lib, ctx.handle
        bpr = max(-(-c // BLOCK) for _, c in ranges)
        hp = anomaly_padding + 1
        rec_len = lib.hypad_tw_shard_record_doubles(bpr, count, anomaly_padding)
        blk = np.asarray([f // BLOCK for f, _ in ranges] + [-(-n // BLOCK)], dtype=np.int64)
        flags = ddof | (scoring._native.STATS_F32 if stats_f32 else 0)
        with torch.cuda.device(dev):
            record = torch.empty(rec_len, dtype=torch.float64, device=dev)
            scoring.check(lib.hypad_tw_shard_pack(h, scoring.ptr(final_local), first, cnt, n, wsize, step, count, anomaly_padding, bpr,
                                                  scoring.ptr(record), ctx.stream()))
            g = self.comm.all_gather(record).clone()  # kept across the fragment exchanges below
            strips = g[:, rec_len - 2 * hp:].contiguous().view(self.world, 2, hp)
            lo, hi = max(first - hp, 0), min(first + cnt + hp, n)
            left, right = scoring._halo_from_strips(strips, ranges, self.rank, first - lo, hi - (first + cnt))
            ext = torch.cat(left + [final_local] + right) if (left or right) else final_local
            while True:
                R = self.max_runs
                per = 8 + 3 * R
                frag = torch.empty(count * per, dtype=torch.float64, device=dev)
                scoring.check(lib.hypad_tw_shard_runs(h, scoring.ptr(g), self.world, self.rank, blk.ctypes.data, bpr, scoring.ptr(ext), lo,
                                                      ext.shape[0], n, wsize, step, count, flags, anomaly_padding, R, scoring.ptr(frag),
                                                      ctx.stream()))
                host = self.comm.all_gather(frag).cpu().numpy()
                cap = self.world * R
                stats = np.empty((count, 4), dtype=np.float64)
                runs = np.empty((count, cap, 3), dtype=np.float64)
                n_runs = np.zeros(count, dtype=np.int32)
                need, over = scoring.ctypes.c_int64(0), scoring.ctypes.c_int(0)
                scoring.check(lib.hypad_tw_shard_merge(host.ctypes.data, self.world, count, R, stats.ctypes.data, runs.ctypes.data,
                                                       n_runs.ctypes.data, cap, scoring.ctypes.byref(need), scoring.ctypes.byref(over)))
                if not over.value:
                    break
                self.max_runs = int(need.value) * 3 // 2 + 16  # the same decision on every rank: the gathered records are identical
        merged = scoring.intervals_from_runs(stats, runs, n_runs, step, min_percent, f32=stats_f32)
        return scoring.intervals_to_index(merged, index)

    max_runs = 64  # room per analysis window in the gathered buffer; grown (on every rank alike) when a window holds more runs

    def find_anomaly_intervals(self, final, index, window_size_portion, window_step_size_portion, min_percent=0.1,
                               anomaly_padding=50, ddof=0, stats_f32=False):
        """scoring.find_anomaly_intervals with the analysis windows dealt out to the ranks: rank r thresholds windows
        [r*per, (r+1)*per) of the (identical, gathered) score array, the packed per-window results are all-gathered and the
        host tail (prune, score, merge) runs on every rank.  The kernels work on the whole array with a first-window index, so
        the block sums are the single-GPU ones and the result is bitwise the single-GPU one."""
        n = final.numel()
        wsize, step, count = scoring.analysis_windows(n, None, window_size_portion, None, window_step_size_portion)
        per = -(-count // self.world)
        k0 = min(self.rank * per, count)
        kc = max(0, min(per, count - k0))
        while True:
            R = self.max_runs
            blen = scoring.threshold_buffer_len(per, R)
            local = torch.zeros(blen, dtype=torch.float64, device=final.device)
            if kc > 0:
                sub = scoring.threshold_windows_launch(final, wsize, step, kc, ddof | (scoring._native.STATS_F32 if stats_f32 else 0),
                                                       anomaly_padding, R, first_window=k0)
                # re-pack the kc-window buffer into the fixed per-window layout of `per` windows
                s_loc, r_loc, n_loc = scoring.threshold_buffer_len(kc, R), kc * 4, kc * R * 3
                local[: kc * 4] = sub[:r_loc]
                local[per * 4: per * 4 + n_loc] = sub[r_loc:r_loc + n_loc]
                local[per * 4 + per * R * 3: per * 4 + per * R * 3 + (kc + 1) // 2] = sub[r_loc + n_loc:s_loc]
            host = self.comm.all_gather(local).cpu().numpy()
            stats, runs, n_runs = [], [], []
            for r in range(self.world):
                rk0 = min(r * per, count)
                rkc = max(0, min(per, count - rk0))
                st, ru, nr = scoring.threshold_windows_parse(host[r], per, R)
                stats.append(st[:rkc])
                runs.append(ru[:rkc])
                n_runs.append(nr[:rkc])
            stats, runs, n_runs = np.concatenate(stats), np.concatenate(runs), np.concatenate(n_runs)
            most = int(n_runs.max(initial=0))
            if most <= R:
                break
            self.max_runs = most * 3 // 2 + 16  # the same decision on every rank: the gathered counts are identical
        merged = scoring.intervals_from_runs(stats, runs, n_runs, step, min_percent, f32=stats_f32)
        return scoring.intervals_to_index(merged, index)
