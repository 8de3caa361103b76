"""Bulk sweep: many signals, one TadGAN / HypAD model each, sharded BY SIGNAL across the GPUs of a node (SURVEY.md 8e-ii,
BASELINE config 5).

The reference scores one signal per process invocation (`python anomaly_detection.py -c cfg.yaml`, one trained model per signal:
train.py:430-437 puts the signal name into the model path); a sweep over the 493 bundled NASA / NAB / YAHOO signals is a shell
loop.  Here every rank takes the signals `assign_signals` deals it (greedy longest-first on the window count: no halo, no
cross-GPU dependency), enqueues the whole scoring pipeline of all of them on its stream without synchronising in between,
extracts the anomaly intervals once everything is queued, and the per-signal interval lists (a few rows each) are exchanged
with one `all_gather_object` at the end.  No collective on the data path.
"""
import ctypes

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

    def __init__(self, scorers, group=None, window=100, streams=8):
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

    _ITEM = np.dtype([("ctx", "<u8"), ("x", "<u8"), ("n", "<i8"), ("tw_window", "<i8"), ("tw_step", "<i8"), ("tw_count", "<i8"),
                      ("critic", "<u8"), ("rec", "<u8"), ("unorm", "<u8"), ("kmax", "<u8"), ("cs", "<u8"), ("final", "<u8"), ("tw", "<u8")])

    def _enqueue_batched(self, scorers, ids, lanes, resident, slot, dev_buf, combination, max_runs, S, used):
        """All signals of this rank in ONE library call (hypad_score_signals_hyperbolic): what is left per signal on the host is
        a row of a table.  Hyperbolic models with device-resident signals only; returns None otherwise (the per-signal loop
        then runs).  A scorer's weights are compared with its packed copy the first time this sweep sees it."""
        if not all(sc.hyperbolic for sc in scorers) or any(i not in resident for i in ids):
            return None
        ddof, f32 = _sc.univariate_hyperbolic_semantics(combination)
        need_c = combination in _sc._NEEDS_CRITIC
        if not hasattr(self, "_checked"):
            self._checked = set()
        for sc in scorers:
            if id(sc) not in self._checked:
                sc.net.ensure(sc.encoder, sc.decoder, sc.critic_x)
                self._checked.add(id(sc))
            used[id(sc)] = sc
        m = len(ids)
        n = np.asarray([resident[i].numel() - S for i in ids], dtype=np.int64)
        npos = n + S - 1
        # analysis windows of find_anomalies (scoring.analysis_windows, vectorised): 33 % windows, 10 % steps
        wsize = np.ceil(n * 0.33).astype(np.int64)
        step = np.ceil(wsize * 0.1).astype(np.int64)
        count = 1 + -(-np.maximum(n - wsize, 0) // step)
        # one device buffer for every signal's arrays: final | kmax | critic_scores (float64), critic | rec | unorm (float32)
        n64 = n + (2 * npos if need_c else 0)
        nbytes = (n64 * 8 + 3 * n * 4 + 15) // 16 * 16
        offs = np.concatenate([[0], np.cumsum(nbytes)[:-1]])
        dev = dev_buf.device
        out_buf = torch.empty(int(nbytes.sum()), dtype=torch.uint8, device=dev)
        base = out_buf.data_ptr() + offs
        items = np.zeros(m, dtype=self._ITEM)
        items["ctx"] = [sc.net.ctx.handle.value for sc in scorers]
        items["x"] = [resident[i].data_ptr() for i in ids]
        items["n"], items["tw_window"], items["tw_step"], items["tw_count"] = n, wsize, step, count
        items["final"] = base
        if need_c:
            items["kmax"], items["cs"] = base + 8 * n, base + 8 * (n + npos)
        f32_base = base + 8 * n64
        items["critic"], items["rec"], items["unorm"] = f32_base, f32_base + 4 * n, f32_base + 8 * n
        slot_off = np.asarray([slot[i][0] for i in ids], dtype=np.int64)
        items["tw"] = dev_buf.data_ptr() + 8 * slot_off
        used_len = count * 4 + count * max_runs * 3 + (count + 1) // 2
        if (used_len > np.asarray([slot[i][1] for i in ids])).any():
            raise HypadError("hypad_b200: a signal's thresholding result does not fit its slot")
        import ctypes

        streams = (ctypes.c_void_p * len(lanes))(*[st.cuda_stream for st in lanes])
        lib = scorers[0].net.ctx.lib
        flags = ddof | (_native.STATS_F32 if f32 else 0)
        with torch.cuda.device(dev):
            _native.check(lib.hypad_score_signals_hyperbolic(items.ctypes.data, m, 1, _native.COMBINE_MODES[combination], flags, 50, max_runs,
                                                             streams, len(lanes)))
        out_buf.record_stream(lanes[0])
        queued = []
        for k, i in enumerate(ids):
            final = out_buf[int(offs[k]): int(offs[k]) + 8 * int(n[k])].view(torch.float64)
            queued.append((i, final, lanes[k % len(lanes)], flags, (int(wsize[k]), int(step[k]), int(count[k])), int(slot_off[k]),
                           int(used_len[k]), f32))
        self._keep = out_buf  # alive until the next sweep: the lanes may still be writing when this method returns
        return queued

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
            parts = [np.asarray(signals[i], dtype=np.float64).reshape(-1) for i in host_ids]
            total_samples = sum(p.shape[0] for p in parts)
            if getattr(self, "_stage", None) is None or self._stage.numel() < total_samples:
                self._stage = torch.empty(max(total_samples, 1), dtype=torch.float64, pin_memory=True)  # pinned once, reused
            np.concatenate(parts, out=self._stage.numpy()[:total_samples])
            big = self._stage[:total_samples].to(_sc.cuda_device(), non_blocking=True)
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
        scorers_now = [self.scorers(i, k % len(lanes)) for k, i in enumerate(ids)]  # built / packed on the caller's stream
        if self.n_streams > 1:
            start = torch.cuda.Event()
            start.record()
            for st in lanes:
                st.wait_event(start)  # the side streams start after whatever the caller queued (uploads, weight packing)
        batched = self._enqueue_batched(scorers_now, ids, lanes, resident, slot, dev_buf, combination, MAX_RUNS, S, used) if ids else None
        if batched is not None:
            queued = batched
            ids = []  # everything is queued
        for k, i in enumerate(ids):
            lane = lanes[k % len(lanes)]
            with torch.cuda.stream(lane):
                sc = scorers_now[k]
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
        # the host tails (prune, score, merge) of all signals in one library call per statistics flavour
        m = len(queued)
        lib = _native.load_library()
        n_out = np.zeros(max(m, 1), dtype=np.int64)
        where = {}  # queue position -> (triples array, first row)
        for flavour in sorted({bool(q[7]) for q in queued}):
            sel = [k for k, q in enumerate(queued) if bool(q[7]) == flavour]
            offs = np.asarray([queued[k][5] for k in sel], dtype=np.int64)
            counts = np.asarray([queued[k][4][2] for k in sel], dtype=np.int64)
            steps = np.asarray([queued[k][4][1] for k in sel], dtype=np.int64)
            part = np.zeros(len(sel), dtype=np.int64)
            cap = 64 * len(sel)
            while True:
                out = np.empty((cap, 3), dtype=np.float64)
                tot = ctypes.c_int64(0)
                _native.check(lib.hypad_sweep_intervals(host_np.ctypes.data, len(sel), offs.ctypes.data, counts.ctypes.data, steps.ctypes.data,
                                                        MAX_RUNS, 0.1, int(flavour), out.ctypes.data, cap, part.ctypes.data, ctypes.byref(tot)))
                if tot.value <= cap:
                    break
                cap = int(tot.value) + 16
            row = 0
            for j, k in enumerate(sel):
                n_out[k] = part[j]
                where[k] = (out, row)
                row += max(int(part[j]), 0)
        for k, (i, final, lane, ddof, (wsize, step, count), off, used_len, f32) in enumerate(queued):
            if n_out[k] == -2:
                raise ZeroDivisionError("Weights sum to zero, can't be normalized")
            if n_out[k] == -1:  # more runs in one analysis window than the buffer holds (never seen): the one-by-one path
                stats, runs, nr = _sc.threshold_windows(final, wsize, step, count, ddof, 50)
                merged = np.asarray(_sc.intervals_from_runs(stats, runs, nr, step, 0.1, f32=f32), dtype=np.float64).reshape(-1, 3)
            else:
                arr, row = where[k]
                merged = arr[row:row + int(n_out[k])]
            idx = np.asarray(indices[i])
            iv = np.empty((merged.shape[0], 3), dtype=np.float64)
            if merged.shape[0]:
                iv[:, 0] = idx[merged[:, 0].astype(np.int64)]
                iv[:, 1] = idx[merged[:, 1].astype(np.int64)]
                iv[:, 2] = merged[:, 2]
            res[i] = {"intervals": iv}
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
