"""Bulk sweep: many signals, one TadGAN / HypAD model each, sharded BY SIGNAL across the GPUs of a node (SURVEY.md 8e-ii,
BASELINE config 5).

The reference scores one signal per process invocation (`python anomaly_detection.py -c cfg.yaml`, one trained model per signal:
train.py:430-437 puts the signal name into the model path); a sweep over the 493 bundled NASA / NAB / YAHOO signals is a shell
loop.  Here every rank takes the signals `assign_signals` deals it (greedy longest-first on the window count: no halo, no
cross-GPU dependency), enqueues the whole scoring pipeline of all of them on its stream without synchronising in between,
extracts the anomaly intervals once everything is queued, and the per-signal interval lists (a few rows each) are exchanged
with one `all_gather_object` at the end.  No collective on the data path.
"""
import numpy as np
import torch

from . import _native
from . import scoring as _sc
from ._native import HypadError


def _n_samples(signal):
    """Samples of one univariate signal given as (T,), (T, 1) or (1, T) array, tensor or list."""
    return int(np.prod(np.shape(signal), dtype=np.int64))


def assign_signals(n_windows, world_size):
    """Greedy longest-processing-time assignment: signals in descending window count (ties: lower signal id first), each to the
    rank with the least work so far (ties: lower rank).  Returns one list of signal ids per rank, in scoring order."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(n_windows)), key=lambda i: (-int(n_windows[i]), i))
    load = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(n_windows[i])
    return out


class SignalSweep:
    """Scores a list of univariate signals, each with its own scorer (or one shared scorer), on this rank's share.

    scorers: a WindowScorer, or a callable `signal_id -> WindowScorer` (models are built / loaded lazily on the rank that needs
    them).  With torch.distributed initialised the signals are dealt out over the ranks of `group`; otherwise this process
    scores all of them."""

    def __init__(self, scorers, group=None, window=100, streams=4):
        """streams > 1 deals a rank's signals out over that many CUDA streams so that the small kernels of short signals overlap
        on the device (a short signal's network launch occupies a dozen of the 148 SMs).  A scorer's packed weights and workspace
        serve one stream at a time: with one shared scorer every stream gets its own copy of the packed model (same modules,
        packed once per stream)."""
        self.shared = not callable(scorers)
        self.n_streams = max(1, int(streams))
        if self.shared:
            clones = {0: scorers}

            def per_lane(_i, lane=0, base=scorers):
                if lane not in clones:
                    clones[lane] = _sc.WindowScorer(base.encoder, base.decoder, base.critic_x, own_context=True)
                return clones[lane]

            self.scorers = per_lane
        else:
            self.scorers = lambda i, lane=0, f=scorers: f(i)
        self.group = group
        self.window = window
        self._streams = None
        self._host_buf = None
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.rank = torch.distributed.get_rank(group)
            self.world = torch.distributed.get_world_size(group)
        else:
            self.rank, self.world = 0, 1

    def plan(self, lengths):
        """Signal ids per rank; signals too short to hold one window are scored by nobody (empty result)."""
        n = [max(0, int(t) - self.window) for t in lengths]
        plan = assign_signals(n, self.world)
        return [[i for i in ids if n[i] > 0] for ids in plan]

    def score_local(self, signals, indices, ids, combination="uncertainty", rec_error_type="dtw", keep_scores=False):
        """Scores the signals `ids`; returns {id: {"intervals": (K,3) array[, "final": device tensor]}}.  Two phases so that the
        device never waits for the host: (1) enqueue the pipeline of every signal -- scores, the thresholding kernels of
        find_anomalies and the copy of their packed result to pinned host memory -- without a synchronisation, (2) one
        synchronisation, then the host bookkeeping (prune, score, merge of the few runs) per signal."""
        queued = []
        used = {}
        if self.n_streams > 1 and self._streams is None:
            self._streams = [torch.cuda.Stream() for _ in range(self.n_streams)]
        lanes = self._streams if self.n_streams > 1 else [torch.cuda.current_stream()]
        # one upload for all host-resident signals of this rank (a pageable copy per signal would stall the host once each)
        resident = {}
        host_ids = [i for i in ids if not (isinstance(signals[i], torch.Tensor) and signals[i].is_cuda)]
        if host_ids:
            flat = np.concatenate([np.asarray(signals[i], dtype=np.float64).reshape(-1) for i in host_ids])
            big = torch.from_numpy(flat).pin_memory().to(_sc.cuda_device(), non_blocking=True)
            off = 0
            for i in host_ids:
                n = _n_samples(signals[i])
                resident[i] = big[off:off + n]
                off += n
        # one device buffer for the packed thresholding results of all signals and one pinned host buffer to receive it: a single
        # device-to-host copy per rank (a pinned allocation and a small copy per signal cost more than the kernels)
        MAX_RUNS = 64
        S = self.window
        slot = {}
        total = 0
        for i in ids:
            T = _n_samples(signals[i])
            # hyperbolic models score T-S windows, Euclidean ones T-1 timesteps: room for the larger layout
            need = max(_sc.threshold_buffer_len(_sc.analysis_windows(n, None, 0.33, None, 0.1)[2], MAX_RUNS) for n in (T - S, T - 1))
            slot[i] = (total, need)
            total += need
        dev = _sc.cuda_device()
        dev_buf = torch.empty(max(total, 1), dtype=torch.float64, device=dev)
        if self._host_buf is None or self._host_buf.numel() < total:
            self._host_buf = torch.empty(max(total, 1), dtype=torch.float64, pin_memory=True)
        if self.n_streams > 1:
            start = torch.cuda.Event()
            start.record()
            for st in lanes:
                st.wait_event(start)  # the side streams start after whatever the caller queued
        for k, i in enumerate(ids):
            lane = lanes[k % len(lanes)]
            with torch.cuda.stream(lane):
                sc = self.scorers(i, k % len(lanes))  # inside the lane: a scorer built on demand packs its weights on the stream that uses them
                first_use = id(sc) not in used
                used[id(sc)] = sc
                x = resident[i] if i in resident else _sc._as_dev(signals[i], torch.float64, sc.device).reshape(-1)
                off, room = slot[i]
                f32 = False
                if sc.hyperbolic:
                    # the whole path of the signal, find_anomalies' device part included, as one library call
                    ddof, f32 = _sc.univariate_hyperbolic_semantics(combination)  # what find_anomalies is handed
                    wsize, step, count = _sc.analysis_windows(x.numel() - S, None, 0.33, None, 0.1)
                    used_len = _sc.threshold_buffer_len(count, MAX_RUNS)
                    flags = ddof | (_native.STATS_F32 if f32 else 0)
                    out = sc.score_chain(x, combination, (wsize, step, count, flags, 50, MAX_RUNS, dev_buf[off:off + used_len]),
                                         check_weights=first_use)
                    final = out["final"]
                else:
                    out = sc.score(x, sliding=True, combination=combination, rec_error_type=rec_error_type, index=None, poll=False)
                    final = out["final"]
                    # find_anomalies' device part queued right behind the scores: no host synchronisation per signal
                    ddof = 0  # an ndarray on the Euclidean path
                    wsize, step, count = _sc.analysis_windows(final.numel(), None, 0.33, None, 0.1)
                    used_len = _sc.threshold_buffer_len(count, MAX_RUNS)
                    if used_len > room:
                        raise HypadError("hypad_b200: signal %d yields %d scores, neither T-window nor T-1" % (i, final.numel()))
                    _sc.threshold_windows_launch(final, wsize, step, count, ddof, 50, MAX_RUNS, out=dev_buf[off:off + used_len])
            queued.append((i, final, lane, ddof | (_native.STATS_F32 if f32 else 0), (wsize, step, count), off, used_len, f32))
        cur = torch.cuda.current_stream()
        if self.n_streams > 1:
            for st in lanes:
                cur.wait_stream(st)
        host_buf = self._host_buf[:max(total, 1)]
        host_buf.copy_(dev_buf, non_blocking=True)
        cur.synchronize()
        host_np = host_buf.numpy()
        res = {}
        for i, final, lane, ddof, (wsize, step, count), off, used_len, f32 in queued:
            stats, runs, nr = _sc.threshold_windows_parse(host_np[off:off + used_len], count, MAX_RUNS)
            if nr.max(initial=0) > MAX_RUNS:  # more runs in one analysis window than the buffer holds (never seen): the one-by-one path
                stats, runs, nr = _sc.threshold_windows(final, wsize, step, count, ddof, 50, max_runs=int(nr.max()) + 16)
            merged = _sc.intervals_from_runs(stats, runs, nr, step, 0.1, f32=f32)
            res[i] = {"intervals": _sc.intervals_to_index(merged, np.asarray(indices[i]))}
            if keep_scores:
                res[i]["final"] = final
        for sc in used.values():
            sc.poll_error()
        return res

    def run(self, signals, indices, combination="uncertainty", rec_error_type="dtw", known_anomalies=None):
        """Every rank returns the full {signal id: (K,3) intervals} map (empty (0,3) array for signals without a window).

        known_anomalies: optional list (one entry per signal) of [(start, end), ...] labelled anomalies in index units.  Then the
        return value is (intervals map, evaluation): per signal the overlap-segment confusion counts the reference records for it
        (utils/anomaly_detection_utils.py:96-105, :579-599 -- `[tn, fp, fn, tp]`, tn is None) and, summed over the sweep, the
        precision / recall / F1 its compute_metrics prints (:241-254)."""
        if len(signals) != len(indices):
            raise HypadError("hypad_b200: %d signals but %d index arrays" % (len(signals), len(indices)))
        plan = self.plan([_n_samples(s) for s in signals])
        local = {i: r["intervals"] for i, r in self.score_local(signals, indices, plan[self.rank], combination, rec_error_type).items()}
        if self.world > 1:
            parts = [None] * self.world
            torch.distributed.all_gather_object(parts, local, group=self.group)
        else:
            parts = [local]
        merged = {}
        for part in parts:
            merged.update(part)
        intervals = {i: merged.get(i, np.empty((0, 3))) for i in range(len(signals))}
        if known_anomalies is None:
            return intervals
        return intervals, evaluate_sweep(intervals, known_anomalies)


def evaluate_sweep(intervals, known_anomalies):
    """Per-signal overlap-segment confusion counts (contextual_confusion_matrix(..., weighted=False), the only variant the
    reference's callers use) and the sweep totals with the reference's metric formulas.  Host work on a handful of intervals."""
    from .utils.anomaly_detection_utils import _overlap_segment, _pad

    per_signal = {}
    fp = fn = tp = 0
    for i, iv in intervals.items():
        expected = [(float(a), float(b)) for a, b in known_anomalies[i]]
        observed = [(float(r[0]), float(r[1])) for r in np.asarray(iv).reshape(-1, 3)]
        _tn, f_p, f_n, t_p = _overlap_segment(_pad(expected), _pad(observed))
        per_signal[i] = [None, f_p, f_n, t_p]
        fp, fn, tp = fp + f_p, fn + f_n, tp + t_p
    precision = tp / (tp + fp) if tp + fp else float("nan")
    recall = tp / (tp + fn) if tp + fn else float("nan")
    f1 = 2 * precision * recall / (precision + recall) if precision + recall > 0 else float("nan")
    return {"per_signal": per_signal, "fp": fp, "fn": fn, "tp": tp, "precision": precision, "recall": recall, "f1": f1}
