"""Host-side logic of the product (window bookkeeping of find_anomalies, shard planning), CPU only."""
import numpy as np
import pytest

from conftest import golden
from oracle import hypad_oracle as ho


def numpy_threshold_windows(errors, wsize, step, count, ddof, pad):
    """What csrc/finish.cu threshold_windows_kernel computes, in numpy, to drive the host tail without a GPU."""
    stats = np.zeros((count, 4))
    runs, n_runs = [], []
    for k in range(count):
        w = errors[k * step:k * step + wsize]
        mean, std = w.mean(), w.std(ddof=ddof)
        thr = mean + 4 * std
        seqs, max_below = ho._find_sequences(w, thr, pad)
        stats[k] = (mean, std, thr, max_below)
        runs.append([(s, e, w[s:e + 1].max()) for s, e in seqs])
        n_runs.append(len(seqs))
    width = max(n_runs + [1])
    arr = np.zeros((count, width, 3))
    for k, r in enumerate(runs):
        for i, (s, e, m) in enumerate(r):
            arr[k, i] = (s, e, m)
    return stats, arr, np.asarray(n_runs)


@pytest.mark.parametrize("ddof,portion,pad", [(0, 0.33, 50), (1, 0.33, 50), (0, 0.2, 200)])
def test_interval_bookkeeping_matches_oracle(ddof, portion, pad):
    from hypad_b200 import scoring

    p = golden("pieces.npz")
    e, idx = p["fa_errors"], p["fa_index"]
    wsize, step, count = scoring.analysis_windows(len(e), None, portion, None, 0.1)
    stats, runs, n_runs = numpy_threshold_windows(e, wsize, step, count, ddof, pad)
    merged = scoring.intervals_from_runs(stats, runs, n_runs, step, 0.1)
    mine = np.asarray([[idx[int(s)], idx[int(t)], sc] for s, t, sc in merged], dtype=np.float64).reshape(-1, 3)
    want = ho.find_anomalies(e, idx, portion, 0.1, anomaly_padding=pad, ddof=ddof)
    assert mine.shape == want.shape and len(want) > 0
    assert np.array_equal(mine[:, :2], want[:, :2])
    np.testing.assert_allclose(mine[:, 2], want[:, 2], rtol=1e-12)


def test_analysis_window_count_matches_reference_loop():
    from hypad_b200 import scoring

    for n in (1, 7, 99, 100, 8540, 8639, 999900):
        wsize, step, count = scoring.analysis_windows(n, None, 0.33, None, 0.1)
        # the reference's loop (utils/anomaly_detection_utils.py:1431-1457)
        start = end = 0
        k = 0
        while end < n:
            end = start + wsize
            k += 1
            start += step
        assert count == k, n
        assert (count - 1) * step < n


def test_window_subset_of_a_suffix_equals_the_same_windows_of_the_full_array():
    """What ShardedScorer.find_anomaly_intervals relies on: analysis window k of errors[k0*step:] is window k0+k of errors
    (same elements, same clamp at the end), so dealing window ranges out to ranks and concatenating reproduces the
    single-rank result."""
    from hypad_b200 import scoring

    p = golden("pieces.npz")
    e = p["fa_errors"]
    wsize, step, count = scoring.analysis_windows(len(e), None, 0.33, None, 0.1)
    full = numpy_threshold_windows(e, wsize, step, count, 1, 50)
    for world in (2, 3, 8):
        per = -(-count // world)
        stats, nruns, runs = [], [], []
        for r in range(world):
            k0 = min(r * per, count)
            kc = max(0, min(per, count - k0))
            if kc == 0:
                continue
            st, ru, nr = numpy_threshold_windows(e[k0 * step:], wsize, step, kc, 1, 50)
            stats.append(st)
            nruns.append(nr)
            runs.extend(ru[k, :nr[k]] for k in range(kc))
        assert np.array_equal(np.concatenate(stats), full[0])
        assert np.array_equal(np.concatenate(nruns), full[2])
        for k, rk in enumerate(runs):
            assert np.array_equal(rk, full[1][k, :full[2][k]])


def test_numpy_order_sum_and_average_are_bitwise_numpy():
    """The host tail averages interval scores with np.average in the reference (:1297); the product's pure-Python version
    reproduces numpy's summation order (8 interleaved accumulators) exactly."""
    from hypad_b200.scoring import _np_average, _np_sum

    rng = np.random.default_rng(5)
    for n in list(range(1, 140)) + [200, 257]:
        v = rng.standard_normal(n) * 10 ** rng.uniform(-3, 3, n)
        w = rng.uniform(1, 500, n)
        assert _np_sum(list(v)) == float(np.add.reduce(v)), n
        assert _np_average(list(v), list(w)) == float(np.average(v, weights=w)), n


def test_evaluation_tail_matches_reference_golden(capsys):
    """contextual_confusion_matrix(weighted=False) / compute_metrics (utils/anomaly_detection_utils.py:241-254,
    :579-655) against the reference's own outputs on seeded interval lists (tests/golden/metrics.json, oracle/make_golden_next.py)."""
    import json
    import os

    import pandas as pd
    from conftest import GOLDEN
    from hypad_b200.utils import anomaly_detection_utils as adu

    cases = json.load(open(os.path.join(GOLDEN, "metrics.json")))
    assert len(cases) >= 30
    for rec in cases:
        e = pd.DataFrame(rec["expected"], columns=["start", "end"])
        o = pd.DataFrame([list(x) + [1.0] for x in rec["observed"]], columns=["start", "end", "score"])
        assert list(adu.contextual_confusion_matrix(e, o, weighted=False)) == rec["counts"], rec
        as_lists = adu.contextual_confusion_matrix([tuple(x) for x in rec["expected"]], [tuple(x) for x in rec["observed"]], weighted=False)
        assert list(as_lists) == rec["counts_lists"], rec
        capsys.readouterr()
        if rec["printed"] is None:
            with pytest.raises(ZeroDivisionError):
                adu.compute_metrics(e, o)
        else:
            m = adu.compute_metrics(e, o)
            assert capsys.readouterr().out == rec["printed"]
            assert 0.0 < m["f1"] <= 1.0
        # the driver's tail: zeros whenever the metrics raise (the reference swallows the exception, :107-110)
        class P:
            save_result = False
        want = [0, 0, 0, 0] if rec["printed"] is None else rec["counts"]
        assert adu._evaluate_and_record(np.asarray([list(x) + [1.0] for x in rec["observed"]]).reshape(-1, 3), e, None, P, "s") == want
    with pytest.raises(NotImplementedError):
        adu.contextual_confusion_matrix([], [], weighted=True)


def test_preprocessing_host_helpers_match_reference_golden():
    """The host side of utils/dataloader.py's preprocessing: segment starts (:127-135, repeated addition), the YAHOO per-second
    index (:44-48) and the known-anomaly runs of save_known_anomalies (:14-33), against the reference's outputs."""
    import pandas as pd
    from hypad_b200.utils import dataloader as dl
    from tests_preprocess_cases import cases, yahoo_cases

    g = golden("preprocess.npz")
    for name, (ts, _vals, interval) in cases().items():
        s = np.sort(np.asarray(ts))
        starts = dl.segment_starts(s[0], s[-1], interval)
        assert np.array_equal(starts, g[name + "/index"]) and starts.dtype == g[name + "/index"].dtype, name
    for name, (vals, flag) in yahoo_cases().items():
        idx = dl.yahoo_index(len(vals))
        want = g["yahoo/" + name + "/timestamp"]
        assert np.array_equal(idx - idx[0], want - want[0]) and idx.dtype == np.float64
        for col in ("is_anomaly", "anomaly"):
            df = pd.DataFrame({"timestamp": 1000.0 + np.arange(len(vals)), "value": vals, col: flag})
            _, runs = dl.known_anomaly_runs(df)
            assert np.array_equal(np.asarray(runs, dtype=np.float64).reshape(-1, 2), g["yahoo/%s/known_%s" % (name, col)].reshape(-1, 2))
    with pytest.raises(ValueError):
        dl.yahoo_index(518402)
    assert dl.yahoo_index(518401).shape == (518401,)


def test_segment_starts_equal_the_repeated_addition_loop():
    """utils/dataloader.py:127-135: `while start_ts <= max_ts: ...; start_ts = end_ts` with end_ts = start_ts + interval.  For
    float timestamps / intervals the k-th start is the k-fold rounded sum, not first + k * interval; segment_starts must give the
    former."""
    from hypad_b200.utils.dataloader import segment_starts

    rng = np.random.default_rng(41)
    cases = [(0, 0, 1), (5, 5, 3), (0, 99, 1), (0, 100, 7), (1285027200, 1285027200 + 21600 * 999, 21600)]
    for _ in range(40):
        first = float(rng.uniform(-1e3, 1.4e9))
        span = float(rng.uniform(0, 5e4))
        cases.append((first, first + span, float(rng.uniform(0.05, 977.3))))
        cases.append((np.float64(first), np.float64(first + span), int(rng.integers(1, 500))))
    for first, last, interval in cases:
        want, s = [], first
        while s <= last:
            want.append(s)
            s = s + interval
        got = segment_starts(first, last, interval)
        assert len(got) == len(want) and all(a == b for a, b in zip(got.tolist(), want)), (first, last, interval)


def test_overlap_segment_counts_equal_the_nested_loops():
    """_overlap_segment as one sign table against the counting rule spelled out with loops (utils/anomaly_detection_utils.py:
    579-599: an expected sequence is a hit when any observed one overlaps it; an observed one is a false positive when it overlaps
    none), on random closed intervals with duplicates, nesting, shared end points and float end points."""
    from hypad_b200.utils import anomaly_detection_utils as adu

    def naive(expected, observed):
        hit_obs = [False] * len(observed)
        tp = fn = 0
        for e in expected:
            found = False
            for j, o in enumerate(observed):
                if (e[0] - o[1]) * (e[1] - o[0]) < 0:
                    found = True
                    hit_obs[j] = True
            tp += found
            fn += not found
        return None, hit_obs.count(False), fn, tp

    rng = np.random.default_rng(43)
    for trial in range(300):
        ne, no = int(rng.integers(0, 8)), int(rng.integers(0, 10))
        scale = 50 if trial % 2 else 1_400_000_000
        def draw(k):
            out = []
            for _ in range(k):
                a = int(rng.integers(0, 60)) + scale
                out.append((a, a + int(rng.integers(0, 12))))
            return out
        expected, observed = draw(ne), draw(no)
        if observed and trial % 3 == 0:
            observed.append(observed[0])  # a duplicate
        if trial % 5 == 0:
            expected = [(float(a) + 0.5, float(b) + 0.5) for a, b in expected]
        pe, po = adu._pad(expected), adu._pad(observed)
        assert adu._overlap_segment(pe, po) == naive(pe, po), (expected, observed)


def test_merge_tail_degenerate_inputs_follow_numpy_and_pandas():
    """merged runs of one position each have weights (end - start) that sum to zero -- np.average raises
    ZeroDivisionError (utils/anomaly_detection_utils.py:1297; recorded from the reference in tests/golden/pieces_r2.npz) and so
    must the host tail; a NaN run maximum sorts LAST in the reference's `sort_values(ascending=False)` (:1224)."""
    from hypad_b200 import scoring

    assert str(golden("pieces_r2.npz")["merge_zero_weights"]).startswith("ZeroDivisionError")
    # two single-position runs next to each other in one window: merged group, both weights 0
    stats = np.asarray([[1.0, 0.5, 3.0, 2.0]])
    runs = np.zeros((1, 2, 3))
    runs[0, 0] = (5, 5, 9.0)
    runs[0, 1] = (6, 6, 8.0)
    with pytest.raises(ZeroDivisionError):
        scoring.intervals_from_runs(stats, runs, np.asarray([2]), 10, 0.1)
    with pytest.raises(ZeroDivisionError):
        ho._merge_sequences([(5, 5, 1.0), (6, 6, 2.0)])
    # NaN maximum: pandas puts it last, so the prune sees [9, 2 (max_below), nan]; (2 - nan) / 2 < 0.1 is False, hence the last
    # "not too small" position is 1 and the reference keeps the 9-run AND the max_below pseudo row (start = stop = -1, :1192)
    import pandas as pd

    order = pd.DataFrame({"max_error": [2.0, 9.0, np.nan]}).sort_values("max_error", ascending=False)["max_error"].tolist()
    assert order[:2] == [9.0, 2.0] and np.isnan(order[2])
    runs[0, 0] = (5, 9, 9.0)
    runs[0, 1] = (40, 44, np.nan)
    got = scoring.intervals_from_runs(stats, runs, np.asarray([2]), 10, 0.1)
    assert [g[:2] for g in got] == [[-1.0, -1.0], [5.0, 9.0]]
    np.testing.assert_allclose([g[2] for g in got], [(2.0 - 3.0) / 1.5, (9.0 - 3.0) / 1.5])


def test_fp32_statistics_interval_scores_follow_the_oracle():
    """find_anomalies on a float32 torch tensor (combination rec / rec_uncertainty): mean, std, threshold and the interval
    scores are single-precision; host tail with f32=True against the oracle's torch-based restatement and the reference's own
    result (tests/golden/pieces_r2.npz)."""
    from hypad_b200 import scoring

    p = golden("pieces_r2.npz")
    e32, idx = p["fa32_errors"], p["fa32_index"]
    e = e32.astype(np.float64)
    wsize, step, count = scoring.analysis_windows(len(e), None, 0.33, None, 0.1)
    stats, runs, n_runs = numpy_threshold_windows(e, wsize, step, count, 1, 50)
    # what the kernel does with HYPAD_STATS_F32: round mean / std, form the threshold in fp32
    stats[:, 0] = stats[:, 0].astype(np.float32)
    stats[:, 1] = stats[:, 1].astype(np.float32)
    stats[:, 2] = (stats[:, 0].astype(np.float32) + np.float32(4) * stats[:, 1].astype(np.float32)).astype(np.float32)
    merged = scoring.intervals_from_runs(stats, runs, n_runs, step, 0.1, f32=True)
    mine = np.asarray([[idx[int(s)], idx[int(t)], sc] for s, t, sc in merged], dtype=np.float64).reshape(-1, 3)
    assert np.array_equal(mine[:, :2], p["fa32_uni"][:, :2])
    np.testing.assert_allclose(mine[:, 2], p["fa32_uni"][:, 2], rtol=2e-6)


def test_library_host_tail_equals_the_python_restatement():
    """hypad_intervals_from_runs (csrc/host_tail.cu, what the product path calls) against the plain-Python tail, bit for bit, on
    random windows: ties, NaN maxima, merged groups of more than 128 members (numpy's pairwise summation recursion), float32."""
    from hypad_b200 import scoring

    rng = np.random.default_rng(5)
    for trial in range(120):
        count, R, step = int(rng.integers(1, 12)), int(rng.integers(1, 40)), int(rng.integers(1, 50))
        if trial % 10 == 0:
            count, R, step = 30, 300, 3  # dense overlapping runs: groups of hundreds of members
        stats = np.c_[rng.normal(1, .1, count), rng.uniform(0.1, 1, count), rng.normal(3, .2, count), rng.normal(2.5, .3, count)]
        nr = rng.integers(0, R + 1, count).astype(np.int32)
        runs = np.zeros((count, R, 3))
        for k in range(count):
            st = np.sort(rng.integers(0, 500, nr[k]))
            runs[k, :nr[k], 0] = st
            runs[k, :nr[k], 1] = st + rng.integers(1, 30, nr[k])
            runs[k, :nr[k], 2] = np.round(rng.normal(3.5, .5, nr[k]), 1)  # rounded: equal maxima do occur
            if rng.random() < 0.1 and nr[k]:
                runs[k, 0, 2] = np.nan
        for f32 in (False, True):
            a = scoring.intervals_from_runs(stats, runs, nr, step, 0.1, f32=f32)
            b = scoring.intervals_from_runs_py(stats, runs, nr, step, 0.1, f32=f32)
            assert len(a) == len(b), (trial, len(a), len(b))
            for x, y in zip(a, b):
                assert all((p == q) or (p != p and q != q) for p, q in zip(x, map(float, y))), (trial, x, y)


def test_sweep_evaluation_counts_follow_the_reference_rule():
    """SignalSweep.run(..., known_anomalies=...): per signal the reference's overlap-segment confusion counts (tp = labelled
    anomalies hit by at least one detection, fn = the others, fp = detections that hit nothing; closed intervals padded by one,
    utils/anomaly_detection_utils.py:579-603), and the totals with its precision / recall / F1 formulas (:241-254)."""
    from hypad_b200.sweep import evaluate_sweep

    rng = np.random.default_rng(9)
    intervals, known = {}, []
    want_fp = want_fn = want_tp = 0
    for i in range(40):
        obs = sorted((int(a), int(a) + int(w)) for a, w in zip(rng.integers(0, 1000, rng.integers(0, 5)), rng.integers(0, 30, 5)))
        exp = sorted((int(a), int(a) + int(w)) for a, w in zip(rng.integers(0, 1000, rng.integers(0, 4)), rng.integers(0, 30, 4)))
        intervals[i] = np.asarray([[a, b, 1.0] for a, b in obs], dtype=np.float64).reshape(-1, 3)
        known.append(exp)
        # the nested-loop rule of the reference on half-open intervals
        e2, o2 = [(a, b + 1) for a, b in exp], [(a, b + 1) for a, b in obs]
        hit_e = [any((e[0] - o[1]) * (e[1] - o[0]) < 0 for o in o2) for e in e2]
        hit_o = [any((e[0] - o[1]) * (e[1] - o[0]) < 0 for e in e2) for o in o2]
        tp, fn, fp = sum(hit_e), len(e2) - sum(hit_e), len(o2) - sum(hit_o)
        want_fp, want_fn, want_tp = want_fp + fp, want_fn + fn, want_tp + tp
        got = evaluate_sweep({i: intervals[i]}, {i: exp})["per_signal"][i]
        assert got == [None, fp, fn, tp], (i, got, (fp, fn, tp))
    ev = evaluate_sweep(intervals, known)
    assert (ev["fp"], ev["fn"], ev["tp"]) == (want_fp, want_fn, want_tp)
    p, r = want_tp / (want_tp + want_fp), want_tp / (want_tp + want_fn)
    assert ev["precision"] == p and ev["recall"] == r and ev["f1"] == 2 * p * r / (p + r)


def test_shard_fragment_merge_joins_runs_across_ranks():
    """hypad_tw_shard_merge (host code): the run fragments of a sharded find_anomalies -- every rank's sorted starts / ends of ITS
    positions, the maximum per own start, the maximum in front of its first start (`lead`, a run begun on an earlier rank) and the
    maximum outside runs -- joined into whole runs.  Reference: the runs of the unsharded array, cut into rank ranges here."""
    import ctypes

    from hypad_b200 import _native

    lib = _native.load_library()

    def key(v):  # the order-preserving integer the kernels carry maxima in
        b = np.float64(v).view(np.uint64)
        return np.uint64(~b) if (int(b) >> 63) else np.uint64(int(b) | (1 << 63))

    rng = np.random.default_rng(17)
    for trial in range(60):
        world, R, count = int(rng.integers(1, 6)), 8, int(rng.integers(1, 4))
        n = 400
        bounds = np.sort(rng.choice(np.arange(1, n), world - 1, replace=False)).tolist() if world > 1 else []
        edges = [0] + bounds + [n]
        per = 8 + 3 * R
        rec = np.zeros((world, count, per))
        want_runs, want_below = [], []
        for k in range(count):
            x = rng.uniform(0, 1, n)
            flagged = np.zeros(n, bool)
            for _ in range(int(rng.integers(0, 4))):
                a = int(rng.integers(0, n - 5))
                flagged[a:a + int(rng.integers(1, 60))] = True
            x[flagged] += 2.0
            # runs = maximal stretches of `flagged` (the dilation is the kernels' business; here a run is a flagged stretch)
            d = np.diff(np.concatenate([[0], flagged.astype(int), [0]]))
            starts, ends = np.flatnonzero(d == 1), np.flatnonzero(d == -1) - 1
            want_runs.append([(float(s), float(e), float(x[s:e + 1].max())) for s, e in zip(starts, ends)])
            want_below.append(float(x[~flagged].max()) if (~flagged).any() else 0.0)
            for r in range(world):
                lo, hi = edges[r], edges[r + 1]
                o = rec[r, k]
                st = [s for s in starts if lo <= s < hi]
                en = [e for e in ends if lo <= e < hi]
                o[0], o[1] = len(st), len(en)
                first = st[0] if st else hi
                lead = x[lo:first][flagged[lo:first]]
                o[2:3].view(np.uint64)[0] = key(lead.max()) if lead.size else 0
                below = x[lo:hi][~flagged[lo:hi]]
                o[3:4].view(np.uint64)[0] = key(below.max()) if below.size else 0
                o[4], o[5], o[6] = 1.0 + k, 0.25, 2.0 + k
                for j, s0 in enumerate(st):
                    nxt = st[j + 1] if j + 1 < len(st) else hi
                    own = x[s0:nxt][flagged[s0:nxt]]
                    o[8 + j] = s0
                    o[8 + R + j] = own.max() if own.size else -np.inf
                for j, e0 in enumerate(en):
                    o[8 + 2 * R + j] = e0
        cap = world * R
        stats, runs = np.empty((count, 4)), np.empty((count, cap, 3))
        n_runs = np.zeros(count, np.int32)
        need, over = ctypes.c_int64(0), ctypes.c_int(0)
        _native.check(lib.hypad_tw_shard_merge(rec.ctypes.data, world, count, R, stats.ctypes.data, runs.ctypes.data, n_runs.ctypes.data, cap,
                                               ctypes.byref(need), ctypes.byref(over)))
        assert not over.value
        for k in range(count):
            assert n_runs[k] == len(want_runs[k]), (trial, k)
            assert [tuple(r) for r in runs[k, :n_runs[k]].tolist()] == want_runs[k], (trial, k)
            assert stats[k].tolist() == [1.0 + k, 0.25, 2.0 + k, want_below[k]]
    # a rank whose fragment list did not fit: reported, with the room that is needed
    rec = np.zeros((2, 1, 8 + 3 * 4))
    rec[1, 0, 0] = rec[1, 0, 1] = 9
    _native.check(lib.hypad_tw_shard_merge(rec.ctypes.data, 2, 1, 4, stats.ctypes.data, runs.ctypes.data, n_runs.ctypes.data, 8, ctypes.byref(need),
                                           ctypes.byref(over)))
    assert over.value == 1 and need.value == 9


def test_sweep_host_tails_equal_the_per_signal_tail():
    """hypad_sweep_intervals (the host tails of a whole sweep in one call) against hypad_intervals_from_runs per signal, on packed
    thresholding results laid out the way the device writes them (stats | runs | n_runs as int32 pairs); a signal one of whose
    windows overflows its run room is reported (-1), not mis-parsed."""
    import ctypes

    from hypad_b200 import _native, scoring

    lib = _native.load_library()
    rng = np.random.default_rng(23)
    R = 16
    bufs, metas = [], []
    off = 0
    for item in range(25):
        count, step = int(rng.integers(1, 9)), int(rng.integers(1, 40))
        stats = np.c_[rng.normal(1, .1, count), rng.uniform(0.1, 1, count), rng.normal(3, .2, count), rng.normal(2.5, .3, count)]
        nr = rng.integers(0, R + 1, count).astype(np.int32)
        if item == 7:
            nr[0] = R + 3  # overflow
        runs = np.zeros((count, R, 3))
        for k in range(count):
            m = min(int(nr[k]), R)
            st = np.sort(rng.integers(0, 300, m))
            runs[k, :m, 0], runs[k, :m, 1], runs[k, :m, 2] = st, st + rng.integers(1, 20, m), np.round(rng.normal(3.5, .5, m), 1)
        tail = np.zeros((count + 1) // 2, dtype=np.float64)
        tail.view(np.int32)[:count] = nr
        bufs.append(np.concatenate([stats.ravel(), runs.ravel(), tail]))
        metas.append((off, count, step, stats, runs, nr))
        off += bufs[-1].shape[0]
    host = np.concatenate(bufs)
    offs = np.asarray([m[0] for m in metas], dtype=np.int64)
    counts = np.asarray([m[1] for m in metas], dtype=np.int64)
    steps = np.asarray([m[2] for m in metas], dtype=np.int64)
    for f32 in (0, 1):
        n_out = np.zeros(len(metas), dtype=np.int64)
        out = np.empty((4096, 3))
        tot = ctypes.c_int64(0)
        _native.check(lib.hypad_sweep_intervals(host.ctypes.data, len(metas), offs.ctypes.data, counts.ctypes.data, steps.ctypes.data, R, 0.1, f32,
                                                out.ctypes.data, 4096, n_out.ctypes.data, ctypes.byref(tot)))
        pos = 0
        for k, (_o, count, step, stats, runs, nr) in enumerate(metas):
            if k == 7:
                assert n_out[k] == -1
                continue
            want = scoring.intervals_from_runs(stats, runs, nr, step, 0.1, f32=bool(f32))
            got = out[pos:pos + n_out[k]].tolist()
            pos += int(n_out[k])
            assert len(got) == len(want) and all((a == b) or (a != a and b != b) for x, y in zip(got, want) for a, b in zip(x, y)), k
        assert pos == tot.value
