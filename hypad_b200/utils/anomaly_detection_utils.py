"""Drop-in for the scoring half of utils/anomaly_detection_utils.py of the reference.

Same function names, argument lists and defaults as the reference (file:line cited per function); the numeric
work runs in hand-written sm_100a kernels through the C-ABI (hypad_b200/scoring.py).  Inputs may be numpy arrays,
CPU tensors or CUDA tensors -- they are moved to the current CUDA device -- and the return types follow the
reference (numpy arrays; a float64 torch tensor where the reference's mixed numpy/torch arithmetic produces one).
A CUDA device is required: nothing here computes on the CPU except the bookkeeping on the handful of anomalous
runs found per analysis window and the evaluation of the detected intervals against the known ones
(`contextual_confusion_matrix`, `compute_metrics`; SURVEY.md 8f rank 4 -- a few intervals per signal).

Not provided (out of scope): plotting, dynamic threshold
search (`_find_threshold`, scipy fmin), `prune_false_positive`, `detect_anomaly`, `regression_errors`, `find_scores`.
"""
import math
import os
import pickle

import numpy as np
import pandas as pd
import torch

from .. import scoring as _sc
from .._native import HypadError


def _dev():
    return _sc.cuda_device()


def _np(t):
    return t.detach().cpu().numpy()


def _critic_dev(critic_score, dev):
    if isinstance(critic_score, torch.Tensor):
        return critic_score.detach().to(dev, torch.float32).reshape(-1).contiguous()
    return torch.as_tensor(np.asarray(critic_score, dtype=np.float32).reshape(-1), device=dev)


# ---- critic scores ------------------------------------------------------------------------------------------


def _compute_critic_score(critics, smooth_window):
    """utils/anomaly_detection_utils.py:307-333."""
    k = _sc._as_dev(critics, torch.float64, _dev()).reshape(-1)
    return _np(_sc.critic_zscore_smooth(k, int(smooth_window)))


def final_critic_scores(critic_score, true_signal):
    """utils/anomaly_detection_utils.py:365-404: (N,) critic values + (N, S[,1]) signal -> (N+S-1,) float64."""
    dev = _dev()
    n, S = true_signal.shape[0], true_signal.shape[1]
    critic = _critic_dev(critic_score, dev)
    if critic.shape[0] != n:
        raise HypadError("final_critic_scores: %d critic values for %d windows" % (critic.shape[0], n))
    kmax = _sc.kde_argmax_overlap(critic, S)
    return _np(_sc.critic_zscore_smooth(kmax, math.trunc(n * 0.01)))


def compute_critic_scores(rec_scores, critic_score, true_signal, params, path):
    """utils/anomaly_detection_utils.py:225-238 (same pickle cache file, same truncation to len(rec_scores))."""
    if params.load and os.path.exists(path + "critic_scores.pickle"):
        with open(path + "critic_scores.pickle", "rb") as handle:
            critic_scores = pickle.load(handle)
    else:
        critic_scores = final_critic_scores(critic_score, true_signal)
        with open(path + "critic_scores.pickle", "wb") as handle:
            pickle.dump(critic_scores, handle, protocol=pickle.HIGHEST_PROTOCOL)
    return critic_scores[: rec_scores.shape[0]]


def combine_scores(combination, critic_scores=[], rec_scores=[], recons_signal=[]):
    """utils/anomaly_detection_utils.py:336-362, including what the reference's mixed numpy / torch arithmetic makes of the
    operand types (critic_scores float64 ndarray, recons_signal float32 ndarray, rec_scores a float32 torch tensor on the
    univariate path and a float64 ndarray on the multivariate one):
      mult, uncertainty     -> float64 torch tensor when rec_scores is a tensor, else float64 ndarray
      critic                -> critic_scores itself;  critic_uncertainty -> float64 ndarray
      rec                   -> rec_scores itself;     rec_uncertainty -> float32 tensor (fp32 product) / float64 ndarray
      sum, sum_uncertainty  -> TypeError when rec_scores is a tensor (`ndarray + Tensor`, :338 / :352-355), else float64 ndarray."""
    if combination not in _sc.HYPERBOLIC_COMBINATIONS:
        raise UnboundLocalError("local variable 'final_scores' referenced before assignment")  # what the reference raises
    want_tensor = isinstance(rec_scores, torch.Tensor)
    if want_tensor and combination in ("sum", "sum_uncertainty"):
        raise TypeError("Concatenation operation is not implemented for NumPy arrays, use np.concatenate() instead. Please do "
                        "not rely on this error; it may not be given on all Python implementations.")
    if combination == "critic":
        return critic_scores
    if combination == "rec":
        return rec_scores
    dev = _dev()
    c = _sc._as_dev(critic_scores, torch.float64, dev) if combination in _sc._NEEDS_CRITIC else None
    r = None
    if combination not in ("critic", "critic_uncertainty"):
        if isinstance(rec_scores, torch.Tensor):
            r = rec_scores.detach().to(dev)
            r = r if r.dtype in (torch.float32, torch.float64) else r.double()
        else:
            a = np.asarray(rec_scores)
            r = torch.as_tensor(a if a.dtype in (np.float32, np.float64) else a.astype(np.float64), device=dev)
    u = None
    if combination.endswith("uncertainty"):
        u = _sc.rownorm(_sc._as_dev(recons_signal, torch.float32, dev))
        n_u = r.shape[0] if r is not None else c.shape[0]
        u = u[:n_u]
    out = _sc.combine(combination, c, r, u)
    if combination == "critic_uncertainty":
        return _np(out)
    if combination == "rec_uncertainty" and want_tensor and rec_scores.dtype == torch.float32:
        return out.float().cpu()  # the kernel formed the product in fp32; the widened copy narrows back without loss
    return out.cpu() if want_tensor else _np(out)


# ---- reconstruction errors (Euclidean path) -----------------------------------------------------------------------


def _point_wise_error(y, y_hat):
    """utils/anomaly_detection_utils.py:761-777."""
    dev = _dev()
    return _np(_sc.point_error(_sc._as_dev(y, torch.float64, dev), _sc._as_dev(y_hat, torch.float64, dev)))


def _area_error(y, y_hat, score_window=10):
    """utils/anomaly_detection_utils.py:780-812 (returns a pandas Series like the reference)."""
    dev = _dev()
    return pd.Series(_np(_sc.area_error(_sc._as_dev(y, torch.float64, dev), _sc._as_dev(y_hat, torch.float64, dev), score_window)))


def _dtw_error(y, y_hat, score_window=10):
    """utils/anomaly_detection_utils.py:815-863 (returns a list like the reference)."""
    dev = _dev()
    return _np(_sc.dtw_error(_sc._as_dev(y, torch.float64, dev), _sc._as_dev(y_hat, torch.float64, dev), score_window)).tolist()


def _true_from_windows(y):
    y = np.asarray(y)
    y2 = y.reshape(y.shape[0], -1)
    return np.concatenate([y2[:, 0], y2[-1, 1:]]).astype(np.float64)


def reconstruction_errors(y, y_hat, step_size=1, score_window=10, smoothing_window=0.01, smooth=True, rec_error_type="point"):
    """utils/anomaly_detection_utils.py:866-962.  Returns (errors float64 (N+S-1,), predictions_vs).

    `predictions_vs` -- min / 25 / 50 / 75 / max of every anti-diagonal, which no caller of the reference uses
    (:925-935) -- is reduced to the medians, shape (N+S-1, 1, 1)."""
    if step_size != 1:
        raise NotImplementedError("hypad_b200: step_size != 1 (the reference hard-codes 1, :461)")
    if isinstance(smoothing_window, float):
        smoothing_window = min(math.trunc(len(y) * smoothing_window), 200)
    dev = _dev()
    true = _sc._as_dev(_true_from_windows(y), torch.float64, dev)
    pred = _sc.median_overlap(_sc._as_dev(y_hat, torch.float32, dev))
    kind = rec_error_type.lower()
    if kind == "point":
        errors = _sc.point_error(true, pred)
    elif kind == "area":
        errors = _sc.area_error(true, pred, score_window)
    elif kind == "dtw":
        errors = _sc.dtw_error(true, pred, score_window)
    else:
        raise UnboundLocalError("local variable 'errors' referenced before assignment")
    if smooth:
        errors = _sc.rolling_mean_centered(errors, int(smoothing_window))
    return _np(errors), _np(pred).reshape(-1, 1, 1)


def score_anomalies(y, y_hat, critic, index, score_window=10, critic_smooth_window=None, error_smooth_window=None, smooth=True,
                    rec_error_type="point", comb="mult", lambda_rec=0.5, path=None, samples_num="0"):
    """utils/anomaly_detection_utils.py:407-576.  Returns (final_scores, true_index, true, predictions).

    With `path` set the reference also computes and pickles the point, area and dtw scores (:516-528) and reuses
    `critic_scores.pickle` when it exists (:470, :512-514); both behaviours are kept."""
    y = np.asarray(y)
    critic_smooth_window = critic_smooth_window or math.trunc(y.shape[0] * 0.01)
    error_smooth_window = error_smooth_window or math.trunc(y.shape[0] * 0.01)
    dev = _dev()
    true = _true_from_windows(y)
    y_hat_d = _sc._as_dev(y_hat, torch.float32, dev)
    S = y_hat_d.shape[1]
    if (not path) or (path and not os.path.exists(path + "critic_scores.pickle")):
        kmax = _sc.kde_argmax_overlap(_critic_dev(critic, dev), S)
        critic_scores = _np(_sc.critic_zscore_smooth(kmax, int(critic_smooth_window)))
        if path:
            with open(path + "critic_scores.pickle", "wb") as handle:
                pickle.dump(critic_scores, handle, protocol=pickle.HIGHEST_PROTOCOL)
    else:
        with open(path + "critic_scores.pickle", "rb") as handle:
            critic_scores = pickle.load(handle)

    true_d = _sc._as_dev(true, torch.float64, dev)
    pred_d = _sc.median_overlap(y_hat_d)

    def rec_scores_for(kind):
        if kind == "point":
            e = _sc.point_error(true_d, pred_d)
        elif kind == "area":
            e = _sc.area_error(true_d, pred_d, score_window)
        elif kind == "dtw":
            e = _sc.dtw_error(true_d, pred_d, score_window)
        else:
            raise UnboundLocalError("local variable 'errors' referenced before assignment")
        if smooth:
            e = _sc.rolling_mean_centered(e, int(error_smooth_window))
        return _np(_sc.zscore_clip(e))

    for ret in ["point", "area", "dtw"]:
        if path and not os.path.exists(path + ret + ".pickle"):
            with open(path + ret + ".pickle", "wb") as handle:
                pickle.dump(rec_scores_for(ret), handle, protocol=pickle.HIGHEST_PROTOCOL)
    if (not path) or (path and not os.path.exists(path + rec_error_type + ".pickle")):
        rec_scores = rec_scores_for(rec_error_type.lower())
        predictions = _np(pred_d).reshape(-1, 1, 1)
        if path:
            with open(path + rec_error_type + ".pickle", "wb") as handle:
                pickle.dump(rec_scores, handle, protocol=pickle.HIGHEST_PROTOCOL)
    else:
        with open(path + rec_error_type + ".pickle", "rb") as handle:
            rec_scores = pickle.load(handle)
            predictions = []
    if comb == "mult":
        final_scores = np.multiply(critic_scores, rec_scores)
    elif comb == "sum":
        final_scores = (1 - lambda_rec) * (critic_scores - 1) + lambda_rec * (rec_scores - 1)
    elif comb == "rec":
        final_scores = rec_scores
    elif comb == "critic":
        final_scores = critic_scores
    else:
        raise ValueError('Unknown combination specified {}, use "mult", "sum", or "rec" instead.'.format(comb))
    return final_scores, index, [[t] for t in true], predictions


# ---- thresholding ---------------------------------------------------------------------------------------------


def find_anomalies(errors, index, z_range=(0, 10), window_size=None, window_size_portion=None, window_step_size=None,
                   window_step_size_portion=None, min_percent=0.1, anomaly_padding=50, lower_threshold=False, fixed_threshold=None):
    """utils/anomaly_detection_utils.py:1363-1472 with fixed_threshold=True (what both callers pass, :88-94, :206-213).

    The reference calls `errors.std()` on whatever it receives: a torch tensor (unbiased, ddof=1) on the univariate
    hyperbolic path and an ndarray (ddof=0) elsewhere; the same rule is applied here.
    Per-window mean/std/threshold, padded runs, run maxima and max_below are computed on the device."""
    if not fixed_threshold:
        raise NotImplementedError("hypad_b200: dynamic thresholding (_find_threshold, scipy fmin) is out of scope; "
                                  "pass fixed_threshold=True as the reference's callers do")
    if lower_threshold:
        raise NotImplementedError("hypad_b200: lower_threshold=True is not used by the reference's callers")
    ddof = 1 if isinstance(errors, torch.Tensor) else 0
    f32 = isinstance(errors, torch.Tensor) and errors.dtype == torch.float32  # fp32 tensor: fp32 mean / std / threshold / scores
    e = _sc._as_dev(errors, torch.float64, _dev()).reshape(-1)
    return _sc.find_anomaly_intervals(e, index, window_size_portion, window_step_size_portion, window_size, window_step_size,
                                      min_percent, anomaly_padding, ddof, stats_f32=f32)


# ---- orchestration ----------------------------------------------------------------------------------------------


def _hyperbolic_rec_scores(recons_signal, true_signal, signal_shape, dev):
    """utils/anomaly_detection_utils.py:58-66 / :167-175."""
    r = _sc._as_dev(recons_signal, torch.float32, dev).reshape(-1, signal_shape)
    t = _sc._as_dev(true_signal, torch.float32, dev).reshape(-1, signal_shape)
    return _sc.poincare_rowdist(r, t)


def _pad(lst):
    """utils/anomaly_detection_utils.py:602-603: closed intervals -> half-open."""
    return [(part[0], part[1] + 1) for part in lst]


def _overlap(expected, observed):
    """utils/anomaly_detection_utils.py:301-304."""
    return (expected[0] - observed[1]) * (expected[1] - observed[0]) < 0


def _overlap_segment(expected, observed, start=None, end=None):
    """utils/anomaly_detection_utils.py:579-599: tp = expected sequences hit by at least one observed one, fn = the others, fp =
    observed sequences that hit nothing.  Returns (None, fp, fn, tp) like the reference (no true negatives for segments).
    One (expected x observed) sign table instead of the nested loops; evaluation over a handful of intervals -- host work."""
    ne, no = len(expected), len(observed)
    if ne == 0 or no == 0:
        return None, no, ne, 0
    e = np.asarray([(x[0], x[1]) for x in expected], dtype=np.float64)
    o = np.asarray([(x[0], x[1]) for x in observed], dtype=np.float64)
    hit = np.sign(e[:, None, 0] - o[None, :, 1]) * np.sign(e[:, None, 1] - o[None, :, 0]) < 0
    tp = int(hit.any(axis=1).sum())
    return None, int(no - hit.any(axis=0).sum()), ne - tp, tp


def contextual_confusion_matrix(expected, observed, data=None, start=None, end=None, weighted=True):
    """utils/anomaly_detection_utils.py:606-655.  `weighted=True` calls `_weighted_segment` / `_contextual_partition`, which the
    reference never defines (NameError there); only the overlap-segment algorithm its callers use (:100-105, :245) exists."""
    if weighted:
        raise NotImplementedError("hypad_b200: contextual_confusion_matrix(weighted=True) relies on _weighted_segment, which the "
                                  "reference does not define either; its callers pass weighted=False")
    if data is not None:
        start = data["timestamp"].min()
        end = data["timestamp"].max()
    if not isinstance(expected, list):
        expected = list(expected[["start", "end"]].itertuples(index=False))
    if not isinstance(observed, list):
        observed = list(observed[["start", "end"]].itertuples(index=False))
    return _overlap_segment(_pad(expected), _pad(observed), start, end)


def compute_metrics(known_anomalies, pred_anomalies):
    """utils/anomaly_detection_utils.py:241-254: prints precision / recall / F1 / gmean of the overlap-segment counts; like the
    reference it raises ZeroDivisionError when a denominator is empty (its callers swallow that).  Also returns the numbers."""
    tn, fp, fn, tp = contextual_confusion_matrix(known_anomalies, pred_anomalies, weighted=False)
    precision = tp / (tp + fp)
    recall = tp / (tp + fn)
    F1 = 2 * (precision * recall) / (precision + recall)
    gmean = np.sqrt(precision * recall)
    print("precision: {}, recall: {}".format(precision, recall))
    print("f1_score: {}, gmean: {}".format(F1, gmean))
    return {"precision": precision, "recall": recall, "f1": F1, "gmean": gmean}


def _evaluate_and_record(intervals, known_anomalies, df, params, signal):
    """The tail of the reference's univariate driver (:96-125): confusion counts against the known anomalies -- [0, 0, 0, 0]
    whenever anything in the block raises, e.g. an empty denominator in compute_metrics -- and the optional results CSV."""
    try:
        pred_anomalies = pd.DataFrame(intervals, columns=["start", "end", "score"])
        out = list(contextual_confusion_matrix(known_anomalies, pred_anomalies, data=df, weighted=False))
        compute_metrics(known_anomalies, pred_anomalies)
    except Exception:
        out = [0, 0, 0, 0]
    if getattr(params, "save_result", False):
        file_place = "./results/{}".format(params.filename)
        res = pd.read_csv(file_place) if os.path.isfile(file_place) else pd.DataFrame(columns=["signal", "tn", "fp", "fn", "tp"])
        if params.signal not in list(res["signal"]):
            res.loc[len(res)] = [signal] + out
            res.to_csv(file_place, index=False)
    return out


def univariate_anomaly_detection(recons_signal, true_signal, params, combination, critic_score, path, read_path,
                                 rec_error_type="euclidean", true_index=None, known_anomalies=None, signal=None, signal_shape=None):
    """utils/anomaly_detection_utils.py:21-126: scores -> find_anomalies -> `path + "anomalies.csv"` -> confusion counts against
    `known_anomalies` (and the results CSV when `params.save_result`).  Returns the (K,3) interval array (the reference returns
    None); the confusion counts of the last call are kept in `univariate_anomaly_detection.last_counts`."""
    dev = _dev()
    if not params.hyperbolic:
        final_scores, true_index, _true, _pred = score_anomalies(true_signal, recons_signal, critic_score, true_index,
                                                                 rec_error_type=rec_error_type, comb=combination, path=path)
        final = np.asarray(final_scores).reshape(-1)
    else:
        rec = _hyperbolic_rec_scores(recons_signal, true_signal, params.signal_shape, dev)
        critic_scores = []
        if combination in _sc._NEEDS_CRITIC:
            critic_scores = compute_critic_scores(rec, critic_score, np.asarray(true_signal), params, path)
        final = combine_scores(combination, critic_scores, rec.cpu(), recons_signal).reshape(-1)  # types as in the reference
    intervals = find_anomalies(final, true_index, window_size_portion=0.33, window_step_size_portion=0.1, fixed_threshold=True)
    if len(intervals):
        # :96-98 sits in a try block: with no interval the reference's find_anomalies returns a shape-(0,) array, the DataFrame
        # constructor raises, the except swallows it and no anomalies.csv is written
        pd.DataFrame(intervals, columns=["start", "end", "score"]).to_csv(path + "anomalies.csv")
    univariate_anomaly_detection.last_counts = None
    if known_anomalies is not None or getattr(params, "save_result", False):
        df = None
        if read_path and os.path.isfile(read_path):
            df = pd.read_csv(read_path)  # only its timestamp range is read (:640-642), so the YAHOO detrend of :35-36 is not needed
        univariate_anomaly_detection.last_counts = _evaluate_and_record(intervals, known_anomalies, df, params, signal)
    return intervals


def multivariate_anomaly_detection(recons_signal, true_signal, params, combination, critic_score, path, x_index=None):
    """utils/anomaly_detection_utils.py:129-222 without the ground-truth loading (:143-151), plotting and metrics.

    x_index defaults to the reference's synthetic per-second timestamps starting 2012-11-24 (:133-137)."""
    dev = _dev()
    recons = np.asarray(recons_signal)
    n = recons.shape[0]
    if x_index is None:
        from datetime import datetime

        x_index = datetime.timestamp(datetime(2012, 11, 24)) + np.arange(n, dtype=np.float64)
    torch.save(x_index, path + "x_index.pt")
    if not params.hyperbolic:
        # :157-161: np.linalg.norm(true_signal - recons_signal, axis=1) -> zscore -> clip(0) + 1
        truth = _sc._as_dev(np.asarray(true_signal).reshape(n, -1), torch.float64, dev)
        rec = _np(_sc.zscore_clip(_sc.rowdiff_norm(truth, _sc._as_dev(recons.reshape(n, -1), torch.float32, dev))))
    else:
        rec = _np(_sc.zscore_clip(_hyperbolic_rec_scores(recons, true_signal, params.signal_shape, dev)))
    critic_scores = []
    if combination in _sc._NEEDS_CRITIC:
        critic_scores = compute_critic_scores(rec, critic_score, np.asarray(true_signal), params, path)
    final_scores = combine_scores(combination, critic_scores, rec, recons)
    torch.save(x_index, path + "true_index.pt")
    intervals = find_anomalies(final_scores, x_index, window_size_portion=0.2, window_step_size_portion=0.1,
                               fixed_threshold=True, anomaly_padding=200)
    pd.DataFrame(intervals, columns=["start", "end", "score"]).to_csv(path + "pred_anomalies.csv")
    return intervals
