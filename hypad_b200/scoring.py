"""Device-side scoring steps and the fused window-scoring pipeline.

Thin Python over the C-ABI (include/hypad_b200.h): every function takes/returns CUDA tensors, allocates its
output with torch (plumbing) and enqueues hand-written sm_100a kernels on the current torch stream.
`WindowScorer` chains them into what anomaly_detection.py:67-155 + utils/anomaly_detection_utils.py:21-94 of the
reference compute for one signal, without ever materialising the (N, S) window matrix for univariate signals.
"""
import ctypes
import math

import numpy as np
import torch

from . import _native, _weights
from ._native import HypadError, check, ptr


def _ctx(t):
    return _native.default_context(t.device)


def _as_dev(a, dtype, device):
    """numpy / CPU tensor / CUDA tensor -> contiguous CUDA tensor of `dtype` on `device`."""
    if isinstance(a, torch.Tensor):
        t = a.detach()
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    return t.to(device=device, dtype=dtype).contiguous()


def cuda_device(device=None):
    if device is not None:
        device = torch.device(device)
        if device.type != "cuda":
            raise HypadError("hypad_b200: device %s is not a CUDA device (no CPU implementation exists)" % device)
        return device if device.index is not None else torch.device("cuda", torch.cuda.current_device())
    if not torch.cuda.is_available():
        raise HypadError("hypad_b200: no CUDA device available; the scoring path has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


# ---------------------------------------------------------------------------------------------------------
# single steps
# ---------------------------------------------------------------------------------------------------------


def window_gather(X, window, out_dtype=torch.float64):
    """rolling_window_sequences(window_size=window, target_size=1, step_size=1): (T,) f64 -> (T-window, window)."""
    X = _native.require_cuda(X, "X").reshape(-1)
    if X.dtype != torch.float64:
        X = X.double()
    n = X.shape[0] - window
    if n <= 0:
        return torch.empty((0, window), dtype=out_dtype, device=X.device)
    out = torch.empty((n, window), dtype=out_dtype, device=X.device)
    c = _ctx(X)
    with torch.cuda.device(X.device):
        check(c.lib.hypad_window_gather(ptr(X), n, window, ptr(out), int(out_dtype == torch.float64), c.stream()))
    return out


def poincare_rowdist(recons, truth):
    recons = _native.require_cuda(recons, "recons").float().contiguous()
    truth = _native.require_cuda(truth, "truth").float().contiguous()
    n, S = recons.shape
    out = torch.empty(n, dtype=torch.float32, device=recons.device)
    c = _ctx(recons)
    with torch.cuda.device(recons.device):
        check(c.lib.hypad_poincare_rowdist(ptr(recons), ptr(truth), n, S, ptr(out), c.stream()))
    return out


def rownorm(x):
    x = _native.require_cuda(x, "x").float().contiguous()
    n, S = x.shape
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    c = _ctx(x)
    with torch.cuda.device(x.device):
        check(c.lib.hypad_rownorm(ptr(x), n, S, ptr(out), c.stream()))
    return out


def rowdiff_norm(truth, recons):
    """np.linalg.norm(truth - recons, axis=1) in float64 (utils/anomaly_detection_utils.py:157)."""
    truth = _native.require_cuda(truth, "truth")
    if truth.dtype not in (torch.float32, torch.float64):
        truth = truth.double()
    truth = truth.reshape(truth.shape[0], -1).contiguous()
    recons = _native.require_cuda(recons, "recons").float().contiguous()
    n, S = recons.shape
    if tuple(truth.shape) != (n, S):
        raise HypadError("hypad_b200: rowdiff_norm shapes differ: %s vs %s" % (tuple(truth.shape), (n, S)))
    out = torch.empty(n, dtype=torch.float64, device=recons.device)
    c = _ctx(recons)
    with torch.cuda.device(recons.device):
        check(c.lib.hypad_rowdiff_norm(ptr(truth), int(truth.dtype == torch.float64), ptr(recons), n, S, ptr(out), c.stream()))
    return out


def kde_argmax_overlap(critic, S, n_windows=None, critic_offset=0, t0=0, t_count=None, exhaustive=False):
    """critic (fp32, one value per window) -> kmax (float64, one value per timestep in [t0, t0+t_count))."""
    critic = _native.require_cuda(critic, "critic").reshape(-1)
    if critic.dtype != torch.float32:
        critic = critic.float()
    critic = critic.contiguous()
    n_windows = critic.shape[0] if n_windows is None else n_windows
    total = n_windows + S - 1
    t_count = total - t0 if t_count is None else t_count
    out = torch.empty(t_count, dtype=torch.float64, device=critic.device)
    c = _ctx(critic)
    fn = c.lib.hypad_kde_argmax_overlap_exhaustive if exhaustive else c.lib.hypad_kde_argmax_overlap
    with torch.cuda.device(critic.device):
        check(fn(ptr(critic), critic_offset, critic.shape[0], n_windows, S, t0, t_count, ptr(out), c.stream()))
    return out


def critic_zscore_smooth(kmax, smooth_window):
    kmax = _native.require_cuda(kmax, "kmax").double().contiguous()
    out = torch.empty_like(kmax)
    c = _ctx(kmax)
    with torch.cuda.device(kmax.device):
        check(c.lib.hypad_critic_zscore_smooth(c.handle, ptr(kmax), kmax.shape[0], int(smooth_window), ptr(out), c.stream()))
    return out


class LocalComm:
    """The exchange of a world of one: a stage's record is handed straight to the next stage."""
    rank, world = 0, 1

    def all_gather(self, buf):
        return buf.view(1, -1)


def _record_buffer(dev, *parts):
    """One uint8 device buffer holding the named parts back to back (8-byte aligned): [(name, dtype, numel)] ->
    (buffer, {name: view}, {name: (byte offset, byte length)})."""
    off, spans = 0, {}
    for name, dtype, numel in parts:
        nbytes = numel * torch.empty(0, dtype=dtype).element_size()
        spans[name] = (off, nbytes, dtype)
        off += (nbytes + 7) // 8 * 8
    buf = torch.empty(max(off, 8), dtype=torch.uint8, device=dev)
    views = {name: buf[o:o + n].view(dt) for name, (o, n, dt) in spans.items()}
    return buf, views, spans


def _halo_from_strips(strips, ranges, rank, need_left, need_right):
    """strips: (world, 2, H) -- every rank's first and last min(len, H) values, right-/left-aligned as stored by
    critic_scores_staged; ranges: [(p0, len)] per rank.  Returns (left, right): the `need_left` values before this rank's
    positions and the `need_right` values after them, walking outwards over as many ranks as it takes."""
    H = strips.shape[2]
    left, right = [], []
    got, r = 0, rank - 1
    while got < need_left and r >= 0:
        ln = min(ranges[r][1], H)
        take = min(ln, need_left - got)
        if take:
            left.insert(0, strips[r, 1, H - take:])  # the rank's last `take` values (its strip is right-aligned)
        got += take
        if ranges[r][1] > H and got < need_left:
            raise HypadError("hypad_b200: smoothing halo reaches beyond a neighbour's strip")
        r -= 1
    got, r = 0, rank + 1
    while got < need_right and r < len(ranges):
        ln = min(ranges[r][1], H)
        take = min(ln, need_right - got)
        if take:
            right.append(strips[r, 0, :take])  # the rank's first `take` values (left-aligned)
        got += take
        if ranges[r][1] > H and got < need_right:
            raise HypadError("hypad_b200: smoothing halo reaches beyond a neighbour's strip")
        r += 1
    return left, right


def critic_scores_staged(kmax_local, ranges, n_total, smooth_window, comm=None, keys_f32=True):
    """_compute_critic_score (:307-333) for this rank's positions of a kmax array sharded by contiguous range.

    kmax_local: float64 device tensor, the positions ranges[comm.rank] = (p0, len) of the n_total-long array.
    Stages (critic_stats.cu) with one small all-gather between them: digit histograms of the radix select (3 passes on
    fp32-representable values) with the smoothing halo riding on the first, then the partial sums of mean / std.
    Returns the smoothed z-scores of the own positions (len,)."""
    comm = comm or LocalComm()
    kmax_local = _native.require_cuda(kmax_local, "kmax").double().contiguous()
    dev = kmax_local.device
    c = _ctx(kmax_local)
    lib, h, st = c.lib, c.handle, c.stream
    if comm.world == 1:  # the same stages chained inside the library: one call instead of a dozen
        out = torch.empty_like(kmax_local)
        with torch.cuda.device(dev):
            check(lib.hypad_critic_scores(h, ptr(kmax_local), kmax_local.shape[0], int(smooth_window), int(keys_f32), ptr(out), st()))
        return out
    p0, ln = ranges[comm.rank]
    if kmax_local.shape[0] != ln:
        raise HypadError("hypad_b200: kmax slice has %d positions, the plan says %d" % (kmax_local.shape[0], ln))
    w = int(smooth_window)
    back, fwd = (w // 2, (w - 1) // 2) if w > 0 else (0, 0)
    H = back if comm.world > 1 else 0
    HW = 8192  # HYPAD_SELECT_HIST_WORDS
    with torch.cuda.device(dev):
        check(lib.hypad_stats_select_begin(h, int(n_total), int(keys_f32), st()))
        strips = None
        for p in range(lib.hypad_stats_select_passes(int(keys_f32))):
            parts = [("hist", torch.int32, HW)]
            if p == 0 and H:
                parts.append(("strips", torch.float64, 2 * H))
            buf, v, spans = _record_buffer(dev, *parts)
            check(lib.hypad_stats_select_hist(h, ptr(kmax_local), ln, p, ptr(v["hist"]), st()))
            if p == 0 and H:
                k = min(ln, H)
                v["strips"].zero_()
                if k:
                    v["strips"][:k] = kmax_local[:k]             # first values, left-aligned
                    v["strips"][2 * H - k:] = kmax_local[ln - k:]  # last values, right-aligned
            g = comm.all_gather(buf)
            if p == 0 and H:
                o, n, _ = spans["strips"]
                strips = g[:, o:o + n].contiguous().view(torch.float64).view(comm.world, 2, H)
            if comm.world > 1:
                hists = g[:, :HW * 4].contiguous()
            else:
                hists = g
            check(lib.hypad_stats_select_pick(h, ptr(hists), comm.world, p, st()))
        rec = torch.empty(8, dtype=torch.float64, device=dev)
        check(lib.hypad_stats_moments_partial(h, ptr(kmax_local), 0, ln, 1, ptr(rec), st()))
        recs = comm.all_gather(rec.view(torch.uint8)).contiguous()
        check(lib.hypad_stats_moments_final(h, ptr(recs), comm.world, int(n_total), 1, 0, st()))
        # the slice with its smoothing halo
        lo, hi = max(p0 - back, 0), min(p0 + ln + fwd, n_total)
        if comm.world > 1 and ln and (lo < p0 or hi > p0 + ln):
            left, right = _halo_from_strips(strips, ranges, comm.rank, p0 - lo, hi - (p0 + ln))
            ext = torch.cat(left + [kmax_local] + right)
        else:
            ext, lo, hi = kmax_local, p0, p0 + ln
        out = torch.empty(ln, dtype=torch.float64, device=dev)
        if ln:
            if ext.shape[0] != hi - lo:
                raise HypadError("hypad_b200: assembled %d positions of kmax, expected %d" % (ext.shape[0], hi - lo))
            check(lib.hypad_critic_smooth_shard(h, ptr(ext), ext.shape[0], lo, int(n_total), p0, ln, w, ptr(out), st()))
    return out


def rolling_mean_staged(x_local, ranges, n_total, window, comm=None, min_periods=None):
    """rolling_mean_centered of an array sharded by contiguous range (ranges[comm.rank] = (p0, len) of n_total positions): the
    first and last window/2 values of every rank are exchanged in one all-gather and give the neighbours their halo."""
    comm = comm or LocalComm()
    x_local = _native.require_cuda(x_local, "x").double().contiguous()
    dev = x_local.device
    c = _ctx(x_local)
    p0, ln = ranges[comm.rank]
    w = int(window)
    mp = w // 2 if min_periods is None else int(min_periods)
    back, fwd = (w // 2, (w - 1) // 2) if w > 0 else (0, 0)
    out = torch.empty(ln, dtype=torch.float64, device=dev)
    if ln == 0 and comm.world == 1:
        return out
    ext, lo, hi = x_local, p0, p0 + ln
    if comm.world > 1 and back > 0:
        H = back
        strips = torch.zeros(2 * H, dtype=torch.float64, device=dev)
        k = min(ln, H)
        if k:
            strips[:k] = x_local[:k]
            strips[2 * H - k:] = x_local[ln - k:]
        g = comm.all_gather(strips).view(comm.world, 2, H)
        lo, hi = max(p0 - back, 0), min(p0 + ln + fwd, n_total)
        if ln:
            left, right = _halo_from_strips(g, ranges, comm.rank, p0 - lo, hi - (p0 + ln))
            ext = torch.cat(left + [x_local] + right)
    if ln:
        with torch.cuda.device(dev):
            check(c.lib.hypad_rolling_mean_shard(c.handle, ptr(ext), ext.shape[0], lo, int(n_total), p0, ln, w, mp, ptr(out), c.stream()))
    return out


def zscore_clip_staged(x_local, n_total, comm=None):
    """stats.zscore + clip(0) + 1 (:177-178, :523-524) of an array sharded over the ranks: mean / std of ALL rows from the
    ranks' partial sums, applied to the local rows."""
    comm = comm or LocalComm()
    x_local = _native.require_cuda(x_local, "x")
    if x_local.dtype not in (torch.float32, torch.float64):
        x_local = x_local.double()
    x_local = x_local.contiguous()
    c = _ctx(x_local)
    out = torch.empty(x_local.shape, dtype=torch.float64, device=x_local.device)
    f32 = int(x_local.dtype == torch.float32)
    with torch.cuda.device(x_local.device):
        rec = torch.empty(8, dtype=torch.float64, device=x_local.device)
        check(c.lib.hypad_stats_moments_partial(c.handle, ptr(x_local), f32, x_local.numel(), 0, ptr(rec), c.stream()))
        recs = comm.all_gather(rec.view(torch.uint8)).contiguous()
        check(c.lib.hypad_stats_moments_final(c.handle, ptr(recs), comm.world, int(n_total), 0, 0, c.stream()))
        check(c.lib.hypad_zscore_clip_apply(c.handle, ptr(x_local), f32, x_local.numel(), ptr(out), c.stream()))
    return out


def rolling_mean_centered(x, window, min_periods=None):
    x = _native.require_cuda(x, "x").double().contiguous()
    out = torch.empty_like(x)
    mp = window // 2 if min_periods is None else min_periods
    c = _ctx(x)
    with torch.cuda.device(x.device):
        check(c.lib.hypad_rolling_mean_centered(c.handle, ptr(x), x.shape[0], int(window), int(mp), ptr(out), c.stream()))
    return out


def zscore_clip(x):
    x = _native.require_cuda(x, "x")
    if x.dtype not in (torch.float32, torch.float64):
        x = x.double()
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.float64, device=x.device)
    c = _ctx(x)
    with torch.cuda.device(x.device):
        check(c.lib.hypad_zscore_clip(c.handle, ptr(x), int(x.dtype == torch.float32), x.numel(), ptr(out), c.stream()))
    return out


def combine(mode, critic_scores=None, rec=None, unorm=None, n=None, lambda_rec=0.5):
    ref = next(t for t in (critic_scores, rec, unorm) if t is not None)
    dev = ref.device
    n = ref.shape[0] if n is None else n
    if critic_scores is not None:
        critic_scores = critic_scores.double().contiguous()
    rec32 = 0
    if rec is not None:
        if rec.dtype == torch.float32:
            rec32 = 1
        elif rec.dtype != torch.float64:
            rec = rec.double()
        rec = rec.contiguous()
    if unorm is not None:
        unorm = unorm.float().contiguous()
    out = torch.empty(n, dtype=torch.float64, device=dev)
    c = _native.default_context(dev)
    with torch.cuda.device(dev):
        check(c.lib.hypad_combine_scores(_native.COMBINE_MODES[mode], ptr(critic_scores), ptr(rec), rec32, ptr(unorm),
                                         float(lambda_rec), n, ptr(out), c.stream()))
    return out


def median_overlap(y_hat):
    y_hat = _native.require_cuda(y_hat, "y_hat").float().contiguous()
    n, S = y_hat.shape
    out = torch.empty(n + S - 1, dtype=torch.float32, device=y_hat.device)
    c = _ctx(y_hat)
    with torch.cuda.device(y_hat.device):
        check(c.lib.hypad_median_overlap(ptr(y_hat), n, S, ptr(out), c.stream()))
    return out


def true_from_signal(x, n, row_stride, S):
    x = _native.require_cuda(x, "x")
    if x.dtype not in (torch.float32, torch.float64):
        x = x.double()
    x = x.contiguous()
    out = torch.empty(n + S - 1, dtype=torch.float64, device=x.device)
    c = _ctx(x)
    with torch.cuda.device(x.device):
        check(c.lib.hypad_true_from_signal(ptr(x), int(x.dtype == torch.float64), n, row_stride, S, ptr(out), c.stream()))
    return out


def _err_inputs(y, y_hat):
    y = _native.require_cuda(y, "y").reshape(-1).double().contiguous()
    y_hat = _native.require_cuda(y_hat, "y_hat").reshape(-1)
    if y_hat.dtype not in (torch.float32, torch.float64):
        y_hat = y_hat.double()
    y_hat = y_hat.contiguous()
    if y.shape != y_hat.shape:
        raise HypadError("hypad_b200: y and y_hat differ in length (%d vs %d)" % (y.shape[0], y_hat.shape[0]))
    return y, y_hat, torch.empty_like(y)


def dtw_error(y, y_hat, score_window=10):
    y, y_hat, out = _err_inputs(y, y_hat)
    c = _ctx(y)
    with torch.cuda.device(y.device):
        check(c.lib.hypad_dtw_error(ptr(y), ptr(y_hat), int(y_hat.dtype == torch.float32), y.shape[0], int(score_window),
                                    ptr(out), c.stream()))
    return out


def point_error(y, y_hat):
    y, y_hat, out = _err_inputs(y, y_hat)
    c = _ctx(y)
    with torch.cuda.device(y.device):
        check(c.lib.hypad_point_error(ptr(y), ptr(y_hat), int(y_hat.dtype == torch.float32), y.shape[0], ptr(out), c.stream()))
    return out


def area_error(y, y_hat, score_window=10):
    y, y_hat, out = _err_inputs(y, y_hat)
    c = _ctx(y)
    with torch.cuda.device(y.device):
        check(c.lib.hypad_area_error(ptr(y), ptr(y_hat), int(y_hat.dtype == torch.float32), y.shape[0], int(score_window),
                                     ptr(out), c.stream()))
    return out


# ---------------------------------------------------------------------------------------------------------
# find_anomalies: device statistics / run extraction + host bookkeeping on a handful of runs
# ---------------------------------------------------------------------------------------------------------


def analysis_windows(n, window_size=None, window_size_portion=None, window_step_size=None, window_step_size_portion=None):
    """Window size, step and count of utils/anomaly_detection_utils.py:1423-1457."""
    window_size = window_size or n
    if window_size_portion:
        window_size = int(np.ceil(n * window_size_portion))
    step = window_step_size or window_size
    if window_step_size_portion:
        step = int(np.ceil(window_size * window_step_size_portion))
    count = 1
    while (count - 1) * step + window_size < n:
        count += 1
    return int(window_size), int(step), count


def threshold_buffer_len(count, max_runs):
    """float64 slots of the packed threshold result: stats (count,4) | runs (count,max_runs,3) | n_runs int32 (count,)."""
    return count * 4 + count * max_runs * 3 + (count + 1) // 2


def threshold_windows_launch(errors, window_size, step, count, ddof, anomaly_padding, max_runs, out=None, exhaustive=False,
                             first_window=0):
    """Enqueues the per-window statistics / run kernels for windows [first_window, first_window + count); returns the packed
    device buffer (no synchronisation)."""
    errors = _native.require_cuda(errors, "errors").reshape(-1).double().contiguous()
    dev = errors.device
    c = _ctx(errors)
    n_stats, n_runs_f = count * 4, count * max_runs * 3
    buf = torch.empty(threshold_buffer_len(count, max_runs), dtype=torch.float64, device=dev) if out is None else out
    base = buf.data_ptr()
    with torch.cuda.device(dev):
        if first_window:
            check(c.lib.hypad_threshold_windows_range(c.handle, ptr(errors), errors.shape[0], window_size, step, int(first_window), count,
                                                      int(ddof), int(anomaly_padding), base, base + 8 * n_stats,
                                                      base + 8 * (n_stats + n_runs_f), max_runs, c.stream()))
        else:
            fn = c.lib.hypad_threshold_windows_exhaustive if exhaustive else c.lib.hypad_threshold_windows
            check(fn(c.handle, ptr(errors), errors.shape[0], window_size, step, count, int(ddof), int(anomaly_padding), base,
                     base + 8 * n_stats, base + 8 * (n_stats + n_runs_f), max_runs, c.stream()))
    return buf


def threshold_windows_parse(host, count, max_runs):
    """(stats (count,4), runs (count,max_runs,3), n_runs (count,)) views of a packed buffer brought to the host."""
    n_stats, n_runs_f = count * 4, count * max_runs * 3
    nr = host[n_stats + n_runs_f: n_stats + n_runs_f + (count + 1) // 2].view(np.int32)[:count]
    return host[:n_stats].reshape(count, 4), host[n_stats:n_stats + n_runs_f].reshape(count, max_runs, 3), nr


_RUNS_HINT = {"max": 64}  # the largest run count of a window seen so far: noisy inputs then get room on the first try


def threshold_windows(errors, window_size, step, count, ddof, anomaly_padding, max_runs=None, exhaustive=False):
    """Per analysis window: (mean, std, threshold, max_below) and the padded above-threshold runs.
    One device buffer holds the three outputs so that a single device-to-host copy (one synchronisation) brings them back."""
    if max_runs is None:
        max_runs = _RUNS_HINT["max"]
    while True:
        host = threshold_windows_launch(errors, window_size, step, count, ddof, anomaly_padding, max_runs,
                                        exhaustive=exhaustive).cpu().numpy()
        stats, runs, nr = threshold_windows_parse(host, count, max_runs)
        if nr.max(initial=0) <= max_runs:
            return stats, runs, nr
        max_runs = int(nr.max()) * 3 // 2 + 16
        _RUNS_HINT["max"] = max(_RUNS_HINT["max"], min(max_runs, 4096))


def _ieee_div(a, b):
    """a / b with numpy's float64 semantics for a zero divisor (the reference divides ndarrays, :1226)."""
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return float("nan")
        return math.copysign(float("inf"), a) * math.copysign(1.0, b)


def _np_sum(vals):
    """np.add.reduce of a 1-D float64 array, in numpy's order: a plain loop below 8 elements, else eight interleaved
    accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and the tail added one by one (numpy's pairwise_sum for
    n <= 128, its unrolled block size)."""
    n = len(vals)
    if n < 8:
        res = 0.0
        for v in vals:
            res += v
        return res
    if n > 128:
        return float(np.add.reduce(np.asarray(vals, dtype=np.float64)))
    r = list(vals[:8])
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] += vals[i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    for v in vals[i:]:
        res += v
    return res


def _np_average(values, weights):
    """np.average(values, weights=weights) (:1297): (v*w).sum() / w.sum(), bit for bit, without numpy's per-call cost.
    Like numpy it refuses a zero weight sum (merged runs of one position each: end - start == 0)."""
    scl = _np_sum(weights)
    if scl == 0.0:
        raise ZeroDivisionError("Weights sum to zero, can't be normalized")
    return _ieee_div(_np_sum([v * w for v, w in zip(values, weights)]), scl)


def _f32(x):
    return float(np.float32(x))


def intervals_from_runs(stats, runs, n_runs, step, min_percent, f32=False):
    """Host tail of find_anomalies on the few runs per window: prune (:1203-1237), score (:1240-1269), merge (:1272-1313), in
    the library's host code (csrc/host_tail.cu; noise can produce tens of thousands of runs, see there).
    f32: the scores came as a float32 torch tensor, so `(max - threshold) / (mean + std)` is single-precision arithmetic.
    Returns [[start, end, score], ...]."""
    stats = np.ascontiguousarray(stats, dtype=np.float64).reshape(-1, 4)
    count = stats.shape[0]
    runs = np.ascontiguousarray(runs, dtype=np.float64).reshape(count, -1, 3)
    n_runs = np.ascontiguousarray(n_runs, dtype=np.int32).reshape(-1)
    max_runs = runs.shape[1]
    lib = _native.load_library()
    cap = int(n_runs.sum()) + count + 1
    out = np.empty((cap, 3), dtype=np.float64)
    n_out = ctypes.c_int64(0)
    rc = lib.hypad_intervals_from_runs(stats.ctypes.data, runs.ctypes.data, n_runs.ctypes.data, count, max(max_runs, 1), int(step),
                                       float(min_percent), int(bool(f32)), out.ctypes.data, cap, ctypes.byref(n_out))
    if rc == -4:
        raise ZeroDivisionError("Weights sum to zero, can't be normalized")
    check(rc)
    return out[: n_out.value].tolist()


def intervals_from_runs_py(stats, runs, n_runs, step, min_percent, f32=False):
    """The same tail in plain Python floats (the cross-check of the library's host code in tests/test_host_logic.py)."""
    stats, n_runs = np.asarray(stats).tolist(), np.asarray(n_runs).tolist()
    sequences = []
    for k in range(len(stats)):
        mean, std, thr, max_below = stats[k]
        rk = runs[k][: int(n_runs[k])].tolist()
        rows = [(max_below, -1.0, -1.0)] + [(r[2], r[0], r[1]) for r in rk]
        # descending by max error, stable, NaN last (pandas sort_values(ascending=False), :1224)
        rows.sort(key=lambda t: (t[0] != t[0], -t[0] if t[0] == t[0] else 0.0))
        last = -1  # the last position whose drop to the next maximum is not "too small" (increase < min_percent is False)
        for i in range(len(rows) - 1):
            if not (_ieee_div(rows[i][0] - rows[i + 1][0], rows[i][0]) < min_percent):
                last = i
        denom = _f32(mean + std) if f32 else mean + std
        shift = k * step
        for m, s, e in rows[: last + 1]:
            score = _f32(_ieee_div(_f32(m - thr), denom)) if f32 else _ieee_div(m - thr, denom)
            sequences.append([s + shift, e + shift, score])
    if not sequences:
        return []
    sequences.sort(key=lambda s: s[0])
    merged = [sequences[0]]
    score, weights = [sequences[0][2]], [sequences[0][1] - sequences[0][0]]
    grouped = False
    for seq in sequences[1:]:
        prev = merged[-1]
        if seq[0] <= prev[1] + 1:
            score.append(seq[2])
            weights.append(seq[1] - seq[0])
            merged[-1] = [prev[0], max(prev[1], seq[1]), None]  # the weighted mean is taken once, when the group closes
            grouped = True
        else:
            if grouped:
                merged[-1][2] = _np_average(score, weights)
            score, weights, grouped = [seq[2]], [seq[1] - seq[0]], False
            merged.append(seq)
    if grouped:
        merged[-1][2] = _np_average(score, weights)
    return merged


def intervals_to_index(merged, index):
    idx = index.detach().cpu().numpy() if isinstance(index, torch.Tensor) else np.asarray(index)
    out = [[float(idx[int(s)]), float(idx[int(e)]), float(sc)] for s, e, sc in merged]
    return np.asarray(out, dtype=np.float64).reshape(-1, 3)


def find_anomaly_intervals(errors, index, window_size_portion=None, window_step_size_portion=None, window_size=None,
                           window_step_size=None, min_percent=0.1, anomaly_padding=50, ddof=0, stats_f32=False):
    """find_anomalies(..., fixed_threshold=True) on a device array; returns (K,3) float64 [index[start], index[end], score].
    stats_f32: the reference would be handed a float32 torch tensor (statistics, threshold and scores in single precision)."""
    n = errors.numel()
    wsize, step, count = analysis_windows(n, window_size, window_size_portion, window_step_size, window_step_size_portion)
    stats, runs, n_runs = threshold_windows(errors, wsize, step, count, ddof | (_native.STATS_F32 if stats_f32 else 0), anomaly_padding)
    merged = intervals_from_runs(stats, runs, n_runs, step, min_percent, f32=stats_f32)
    return intervals_to_index(merged, index)


def univariate_hyperbolic_semantics(combination):
    """What reaches find_anomalies on the reference's univariate hyperbolic path (utils/anomaly_detection_utils.py:54-94), where
    the critic scores are a float64 ndarray, the reconstruction scores a float32 torch tensor and the norms a float32 ndarray:
      mult, uncertainty        float64 torch tensor  -> unbiased std (ddof 1)              (:340, :343)
      critic, critic_uncertainty  float64 ndarray    -> ddof 0                             (:345, :349)
      rec, rec_uncertainty     float32 torch tensor  -> ddof 1, single-precision statistics (:356, :360)
      sum, sum_uncertainty     `ndarray + tensor` raises TypeError in the reference (:338, :352-355); so does this path.
    Returns (ddof, stats_f32)."""
    if combination in ("mult", "uncertainty"):
        return 1, False
    if combination in ("critic", "critic_uncertainty"):
        return 0, False
    if combination in ("rec", "rec_uncertainty"):
        return 1, True
    if combination in ("sum", "sum_uncertainty"):
        raise TypeError("combination %r adds a float64 ndarray and a float32 torch tensor on the univariate hyperbolic path; the "
                        "reference raises here too (utils/anomaly_detection_utils.py:338, :352-355: 'Concatenation operation is "
                        "not implemented for NumPy arrays'); it is defined for multivariate scoring, where both are ndarrays"
                        % (combination,))
    raise ValueError("unknown combination %r" % (combination,))


# ---------------------------------------------------------------------------------------------------------
# the fused pipeline
# ---------------------------------------------------------------------------------------------------------

HYPERBOLIC_COMBINATIONS = ("mult", "uncertainty", "sum", "sum_uncertainty", "critic", "critic_uncertainty", "rec", "rec_uncertainty")
_NEEDS_CRITIC = ("mult", "uncertainty", "sum", "sum_uncertainty", "critic", "critic_uncertainty")


def download_async(owner, src, out_host):
    """src (device) -> out_host (pinned) on owner's side stream, ordered after the work queued so far on the current stream."""
    if getattr(owner, "_down_stream", None) is None:
        owner._down_stream = torch.cuda.Stream(device=src.device)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(src.device))
    owner._down_stream.wait_event(ev)
    with torch.cuda.stream(owner._down_stream):
        out_host[: src.numel()].copy_(src, non_blocking=True)
    src.record_stream(owner._down_stream)


class WindowScorer:
    """Scores every window of a signal with a (random-init or trained) TadGAN / HypAD model on one B200.

    encoder, decoder, critic_x: hypad_b200.models.tadgan modules (or reference modules with the same attribute
    names) whose parameters live on a CUDA device.
    """

    def __init__(self, encoder, decoder, critic_x, own_context=False):
        """own_context: pack the model into a context of this scorer's own instead of the one cached for the module triple -- a
        context (packed weights, workspace, statistics state) serves one stream at a time, so scorers of the same modules that
        run side by side on different streams (sweep.SignalSweep) each need theirs."""
        for m in (encoder, decoder, critic_x):
            if m.training:
                raise HypadError("hypad_b200: call .eval() on the modules first (scoring is eval-mode only)")
        self.encoder, self.decoder, self.critic_x = encoder, decoder, critic_x
        if own_context:
            self.net = _weights.PackedNet(next(encoder.parameters()).device).ensure(encoder, decoder, critic_x)
        else:
            self.net = _weights.packed_net(encoder, decoder, critic_x)
        self.device = self.net.device
        self.S = self.net.S
        self.hyperbolic = self.net.hyperbolic

    # -- network -------------------------------------------------------------------------------------------
    def _input(self, x, sliding):
        x = _native.require_cuda(x, "signal" if sliding else "windows")
        if x.dtype not in (torch.float32, torch.float64):
            x = x.double()
        if sliding:
            x = x.reshape(-1).contiguous()
            return x, x.shape[0] - self.S, 1
        if x.dim() < 2 or x.shape[0] == 0:
            raise HypadError("hypad_b200: no windows to score (windows must be a non-empty (N, %d) array, got shape %s)"
                             % (self.S, tuple(x.shape)))
        x = x.reshape(x.shape[0], -1).contiguous()
        if x.shape[1] != self.S:
            raise HypadError("hypad_b200: windows have %d samples, the model expects %d" % (x.shape[1], self.S))
        return x, x.shape[0], self.S

    def forward(self, x, sliding, keep=(), first=0, count=None, ffma=False, into=None, into_offset=0):
        """Runs the fused network over windows [first, first+count).  Returns dict of device tensors:
        critic (n,) always; rec, unorm (n,) when hyperbolic; plus any of z/eucl/hyper/hyper_x named in `keep`.
        ffma=True runs the FFMA cross-check kernel instead of the tensor-core product path.
        into: a result dict of an earlier call sized for more windows; this call writes its rows from `into_offset` on
        (forward_from_host scores a signal chunk by chunk into one set of arrays)."""
        self.net.ensure(self.encoder, self.decoder, self.critic_x)
        x, n_all, stride = self._input(x, sliding)
        count = n_all - first if count is None else count
        if count <= 0:
            raise HypadError("hypad_b200: no windows to score (signal shorter than the window?)")
        dev, S = x.device, self.S
        if into is None:
            res = self._alloc_results(count, keep, dev)
            views = res
        else:
            res = into
            views = {k: v[into_offset:into_offset + count] for k, v in into.items() if not k.startswith("_")}
        out = _native.hypad_forward_out()
        for name, t in views.items():
            setattr(out, name, t.data_ptr())
        stages = _native.STAGE_ENCODER | _native.STAGE_DECODER | _native.STAGE_CRITIC
        if self.hyperbolic:
            stages |= _native.STAGE_MOBIUS_X
        base = x.data_ptr() + first * stride * x.element_size()
        fn = self.net.ctx.lib.hypad_forward_ffma if ffma else self.net.ctx.lib.hypad_forward
        with torch.cuda.device(dev):
            check(fn(self.net.ctx.handle, base, int(x.dtype == torch.float64), count, stride, None, stages, out,
                     self.net.ctx.stream()))
        if into is None:
            res["_x"], res["_n"], res["_stride"] = x, n_all, stride
        return res

    def _alloc_results(self, count, keep, dev):
        S = self.S
        res = {"critic": torch.empty(count, dtype=torch.float32, device=dev)}
        if self.hyperbolic:
            res["rec"] = torch.empty(count, dtype=torch.float32, device=dev)
            res["unorm"] = torch.empty(count, dtype=torch.float32, device=dev)
        widths = {"z": self.net.latent, "eucl": S, "hyper": S, "hyper_x": S}
        for name in keep:
            if name in ("hyper", "hyper_x") and not self.hyperbolic:
                continue
            res[name] = torch.empty((count, widths[name]), dtype=torch.float32, device=dev)
        return res

    UPLOAD_CHUNKS = (0.08, 0.25, 0.67)  # fractions of the windows per upload / launch: a short first chunk starts the GPU early

    def forward_from_host(self, host_signal, keep=()):
        """forward() of a sliding-window signal that lives in (pinned) host memory: the signal is uploaded in a few chunks on a
        copy stream and the fused kernel is launched per chunk as its samples arrive, so that all but the first upload is
        hidden behind the network.  Same results as forward(host_signal.to(device), True): every window is computed by the
        same code from the same samples."""
        if host_signal.is_cuda:
            return self.forward(host_signal, True, keep)
        dev, S = self.device, self.S
        T = host_signal.numel()
        n = T - S
        if n <= 0:
            raise HypadError("hypad_b200: no windows to score (signal shorter than the window?)")
        host_signal = host_signal.reshape(-1)
        if host_signal.dtype not in (torch.float32, torch.float64):
            host_signal = host_signal.double()
        x = torch.empty(T, dtype=host_signal.dtype, device=dev)
        if n < 200000:  # short signals: one copy, one launch
            x.copy_(host_signal, non_blocking=True)
            return self.forward(x, True, keep)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        self._copy_stream.wait_stream(main)  # x was allocated on the main stream
        bounds, acc = [0], 0.0
        for f in self.UPLOAD_CHUNKS[:-1]:
            acc += f
            bounds.append(int(n * acc) // 64 * 64)  # whole 64-window tiles per launch
        bounds.append(n)
        res = self._alloc_results(n, keep, dev)
        up = 0
        for a, b in zip(bounds[:-1], bounds[1:]):
            hi = T if b == n else b + S  # samples the windows [a, b) read: [a, b + S - 1]; the last chunk takes the tail
            with torch.cuda.stream(self._copy_stream):
                x[up:hi].copy_(host_signal[up:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            up = hi
            main.wait_event(ev)
            self.forward(x, True, keep, first=a, count=b - a, into=res, into_offset=a)
        res["_x"], res["_n"], res["_stride"] = x, n, 1
        return res

    def poll_error(self):
        """Synchronises and raises if the forward kernel reported an error (a bounded barrier wait timed out; in strict mode an
        operand outside the range of the scaled fp16 split, |x| < 63 -- include/hypad_b200.h, hypad_forward).  A call that left the
        range was redone by the FFMA kernel on the device: valid results, warned about once per scorer and counted in
        `range_fallbacks`."""
        check(self.net.ctx.lib.hypad_ctx_poll_error(self.net.ctx.handle))
        n = self.range_fallbacks
        if n and not getattr(self, "_warned_fallback", False):
            import warnings

            self._warned_fallback = True
            warnings.warn("hypad_b200: an operand left the tensor-core kernel's range (|x| < 63, linear activations < 255); the "
                          "call was served by the FFMA kernel instead (valid results, about 5x slower)", RuntimeWarning, stacklevel=2)

    @property
    def range_fallbacks(self):
        return int(self.net.ctx.lib.hypad_ctx_range_fallbacks(self.net.ctx.handle))

    def set_strict_range(self, strict=True):
        check(self.net.ctx.lib.hypad_ctx_set_strict_range(self.net.ctx.handle, int(bool(strict))))

    # -- scoring -------------------------------------------------------------------------------------------
    def score_chain(self, x, combination, tw=None, check_weights=True):
        """The univariate hyperbolic path of one device-resident signal as ONE library call (hypad_score_signal_hyperbolic):
        network, KDE aggregation, critic scores, combination and -- tw = (window, step, count, ddof flags, padding, max_runs,
        packed buffer) -- the device part of find_anomalies, queued back to back without returning to Python in between.  All
        results are views of one device allocation.  Same kernels, same results as the step-by-step calls.
        check_weights=False skips the comparison of the modules' parameters with the packed copy (a sweep checks each scorer
        once per run: the walk over 43 parameters costs as much host time as the launches of a short signal)."""
        if check_weights:
            self.net.ensure(self.encoder, self.decoder, self.critic_x)
        x, n, _ = self._input(x, True)
        if n <= 0:
            raise HypadError("hypad_b200: no windows to score (signal shorter than the window?)")
        S, dev = self.S, x.device
        npos = n + S - 1
        need_c = combination in _NEEDS_CRITIC
        # one allocation: final f64 (n) | kmax f64 (npos) | critic_scores f64 (npos) | critic, rec, unorm f32 (n each)
        n64 = n + (2 * npos if need_c else 0)
        buf = torch.empty(n64 * 8 + 3 * n * 4, dtype=torch.uint8, device=dev)
        d64 = buf[: n64 * 8].view(torch.float64)
        f32 = buf[n64 * 8:].view(torch.float32)
        out = {"final": d64[:n], "critic": f32[:n], "rec": f32[n:2 * n], "unorm": f32[2 * n:3 * n]}
        so = _native.hypad_signal_out()
        base = buf.data_ptr()
        so.final = base
        if need_c:
            out["kmax"], out["critic_scores_full"] = d64[n:n + npos], d64[n + npos:]
            out["critic_scores"] = out["critic_scores_full"][:n]
            so.kmax, so.critic_scores = base + 8 * n, base + 8 * (n + npos)
        else:
            out["critic_scores"] = None
        so.critic, so.rec, so.unorm = base + 8 * n64, base + 8 * n64 + 4 * n, base + 8 * n64 + 8 * n
        w = s_ = c = dd = pad = mr = 0
        if tw is not None:
            w, s_, c, dd, pad, mr, twbuf = tw
            so.tw = twbuf.data_ptr()
        ctx = self.net.ctx
        with torch.cuda.device(dev):
            check(ctx.lib.hypad_score_signal_hyperbolic(ctx.handle, ptr(x), int(x.dtype == torch.float64), n, _native.COMBINE_MODES[combination],
                                                        w, s_, c, dd, pad, mr, ctypes.byref(so), ctx.stream()))
        return out

    def score_chain_euclidean(self, x, combination, rec_error_type="dtw", lambda_rec=0.5, tw=None, check_weights=True):
        """The per-timestep Euclidean path of one device-resident signal as ONE library call (hypad_score_signal_euclidean);
        same layout of the result as score(): critic, eucl, kmax, critic_scores, true, pred, errors, rec, final."""
        if check_weights:
            self.net.ensure(self.encoder, self.decoder, self.critic_x)
        x, n, _ = self._input(x, True)
        if n <= 0:
            raise HypadError("hypad_b200: no windows to score (signal shorter than the window?)")
        mode = {"mult": "mult", "sum": "euclidean_sum", "rec": "rec", "critic": "critic"}.get(combination)
        if mode is None:
            raise ValueError('Unknown combination specified {}, use "mult", "sum", or "rec" instead.'.format(combination))
        kind = {"dtw": 0, "point": 1, "area": 2}.get(rec_error_type.lower())
        if kind is None:
            raise ValueError("unknown rec_error_type %r" % (rec_error_type,))
        S, dev = self.S, x.device
        npos = n + S - 1
        # one allocation: six float64 arrays of npos | eucl f32 (n, S) | critic f32 (n) | pred f32 (npos)
        buf = torch.empty(6 * npos * 8 + (n * S + n + npos) * 4, dtype=torch.uint8, device=dev)
        d64 = buf[: 6 * npos * 8].view(torch.float64).view(6, npos)
        f32 = buf[6 * npos * 8:].view(torch.float32)
        out = {"final": d64[0], "kmax": d64[1], "critic_scores": d64[2], "true": d64[3], "errors": d64[4], "rec": d64[5],
               "eucl": f32[: n * S].view(n, S), "critic": f32[n * S: n * S + n], "pred": f32[n * S + n:]}
        so = _native.hypad_signal_eucl_out()
        for name, key in (("final", "final"), ("kmax", "kmax"), ("critic_scores", "critic_scores"), ("truth", "true"), ("errors", "errors"),
                          ("rec", "rec"), ("eucl", "eucl"), ("critic", "critic"), ("pred", "pred")):
            setattr(so, name, out[key].data_ptr())
        w = s_ = c = dd = pad = mr = 0
        if tw is not None:
            w, s_, c, dd, pad, mr, twbuf = tw
            so.tw = twbuf.data_ptr()
        ctx = self.net.ctx
        with torch.cuda.device(dev):
            check(ctx.lib.hypad_score_signal_euclidean(ctx.handle, ptr(x), int(x.dtype == torch.float64), n, _native.COMBINE_MODES[mode], kind,
                                                       float(lambda_rec), w, s_, c, dd, pad, mr, ctypes.byref(so), ctx.stream()))
        return out

    def critic_scores(self, critic, n_windows):
        """final_critic_scores (:365-404): KDE arg-max overlap aggregation + quantile-band z-score + smoothing."""
        kmax = kde_argmax_overlap(critic, self.S)
        total = kmax.shape[0]
        # kmax holds fp32 critic values widened to float64: 32-bit keys, three select passes
        return critic_scores_staged(kmax, [(0, total)], total, math.trunc(n_windows * 0.01)), kmax

    def score(self, x, sliding=True, combination="uncertainty", rec_error_type="dtw", index=None, keep=(), multivariate=False,
              lambda_rec=0.5, poll=True, out_host=None):
        """Per-position anomaly scores (+ intervals when `index` is given) for one signal.

        sliding=True : x is the scaled signal (T,), windows are x[n:n+S], n in [0, T-S)   (univariate configs)
        sliding=False: x is (N, S) materialised windows / multivariate rows.
        Hyperbolic models return one score per window (N,), Euclidean ones one per timestep (N+S-1,), like the reference.
        poll=False leaves out the final synchronising error poll: a caller that enqueues many signals (sweep.SignalSweep) calls
        `poll_error()` once after the last one.
        x may be a pinned host tensor when sliding (uploaded chunk by chunk under the network, forward_from_host); out_host, a
        pinned float64 host tensor, receives the final scores on a side stream while the interval extraction still runs (the
        call returns after both are done).
        """
        keep = tuple(keep)
        stats_f32 = False
        if self.hyperbolic and not multivariate:
            univariate_hyperbolic_semantics(combination)  # unknown / undefined combinations fail before any work is queued
        if (self.hyperbolic and sliding and not multivariate and not keep and isinstance(x, torch.Tensor) and x.is_cuda
                and x.numel() - self.S < 200000):
            # short device-resident signals are launch- and host-bound: the whole path as one library call
            ddof, stats_f32 = univariate_hyperbolic_semantics(combination)
            tw = None
            if index is not None:
                n = x.numel() - self.S
                wsize, step, count = analysis_windows(n, None, 0.33, None, 0.1)
                mr = _RUNS_HINT["max"]
                twbuf = torch.empty(threshold_buffer_len(count, mr), dtype=torch.float64, device=x.device)
                tw = (wsize, step, count, ddof | (_native.STATS_F32 if stats_f32 else 0), 50, mr, twbuf)
            out = self.score_chain(x, combination, tw)
            if out_host is not None:
                download_async(self, out["final"], out_host)
            if index is not None:
                stats, runs, nr = threshold_windows_parse(twbuf.cpu().numpy(), count, mr)
                if nr.max(initial=0) > mr:  # more runs than the buffer holds: the growing one-by-one path
                    stats, runs, nr = threshold_windows(out["final"], wsize, step, count, tw[3], 50)
                out["intervals"] = intervals_to_index(intervals_from_runs(stats, runs, nr, step, 0.1, f32=stats_f32), index)
            if out_host is not None:
                self._down_stream.synchronize()
            if poll:
                self.poll_error()
            return out
        if (not self.hyperbolic and sliding and not multivariate and set(keep) <= {"eucl"} and isinstance(x, torch.Tensor) and x.is_cuda
                and x.numel() - self.S < 200000):
            tw = None
            if index is not None:
                npos = x.numel() - 1
                wsize, step, count = analysis_windows(npos, None, 0.33, None, 0.1)
                mr = _RUNS_HINT["max"]
                twbuf = torch.empty(threshold_buffer_len(count, mr), dtype=torch.float64, device=x.device)
                tw = (wsize, step, count, 0, 50, mr, twbuf)  # an ndarray on the Euclidean path: ddof 0
            out = self.score_chain_euclidean(x, combination, rec_error_type, lambda_rec, tw)
            if out_host is not None:
                download_async(self, out["final"], out_host)
            if index is not None:
                stats, runs, nr = threshold_windows_parse(twbuf.cpu().numpy(), count, mr)
                if nr.max(initial=0) > mr:
                    stats, runs, nr = threshold_windows(out["final"], wsize, step, count, 0, 50)
                out["intervals"] = intervals_to_index(intervals_from_runs(stats, runs, nr, step, 0.1), index)
            if out_host is not None:
                self._down_stream.synchronize()
            if poll:
                self.poll_error()
            return out
        need = keep if self.hyperbolic else tuple(set(keep) | {"eucl"})
        if sliding and isinstance(x, torch.Tensor) and not x.is_cuda:
            fw = self.forward_from_host(x, need)
        else:
            fw = self.forward(x, sliding, need)
        n, S = fw["critic"].shape[0], self.S
        out = {k: v for k, v in fw.items() if not k.startswith("_")}
        if self.hyperbolic:
            if combination not in HYPERBOLIC_COMBINATIONS:
                raise ValueError("unknown combination %r" % (combination,))
            rec = fw["rec"]
            if multivariate:
                rec = zscore_clip(rec)  # utils/anomaly_detection_utils.py:177-178
                out["rec"] = rec
            cs = None
            if combination in _NEEDS_CRITIC:
                cs_full, kmax = self.critic_scores(fw["critic"], n)
                out["kmax"], out["critic_scores_full"] = kmax, cs_full
                cs = cs_full[:n]
            out["critic_scores"] = cs
            ddof, stats_f32 = (0, False) if multivariate else univariate_hyperbolic_semantics(combination)
            final = combine(combination, cs, rec, fw["unorm"], n=n)
        elif multivariate:
            # utils/anomaly_detection_utils.py:153-213, Euclidean branch: one score per row
            if sliding:
                raise ValueError("multivariate scoring takes (N, C) rows (sliding=False)")
            rec = zscore_clip(rowdiff_norm(fw["_x"], fw["eucl"]))  # :157-161
            out["rec"] = rec
            cs = None
            if combination in _NEEDS_CRITIC:
                cs_full, kmax = self.critic_scores(fw["critic"], n)
                out["kmax"], out["critic_scores_full"] = kmax, cs_full
                cs = cs_full[:n]
            out["critic_scores"] = cs
            unorm = rownorm(fw["eucl"]) if "uncertainty" in combination else None
            final = combine(combination, cs, rec, unorm, n=n)
            ddof = 0
        else:
            mode = {"mult": "mult", "sum": "euclidean_sum", "rec": "rec", "critic": "critic"}.get(combination)
            if mode is None:
                raise ValueError('Unknown combination specified {}, use "mult", "sum", or "rec" instead.'.format(combination))
            cs, kmax = self.critic_scores(fw["critic"], n)
            true = true_from_signal(fw["_x"], n, fw["_stride"], S)
            pred = median_overlap(fw["eucl"])
            kind = rec_error_type.lower()
            if kind == "dtw":
                errors = dtw_error(true, pred, 10)
            elif kind == "point":
                errors = point_error(true, pred)
            elif kind == "area":
                errors = area_error(true, pred, 10)
            else:
                raise ValueError("unknown rec_error_type %r" % (rec_error_type,))
            errors = rolling_mean_centered(errors, math.trunc(n * 0.01))
            rec = zscore_clip(errors)
            final = combine(mode, cs, rec, None, n=n + S - 1, lambda_rec=lambda_rec)
            out.update(kmax=kmax, critic_scores=cs, rec=rec, pred=pred, true=true, errors=errors)
            ddof = 0
        out["final"] = final
        if out_host is not None:
            download_async(self, final, out_host)
        if index is not None:
            if multivariate:
                out["intervals"] = find_anomaly_intervals(final, index, 0.2, 0.1, anomaly_padding=200, ddof=ddof)
            else:
                out["intervals"] = find_anomaly_intervals(final, index, 0.33, 0.1, anomaly_padding=50, ddof=ddof,
                                                          stats_f32=stats_f32)
        # loud failure: pipeline protocol error, or an operand outside the tensor-core path's range.  Last, so that the
        # synchronisation it implies does not stall the launches above.
        if out_host is not None:
            self._down_stream.synchronize()
        if poll:
            self.poll_error()
        return out
