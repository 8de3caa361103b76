"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) through oracle/ref_harness.py.

Run in the build container only:   python oracle/make_golden.py
The reference has no tests or golden vectors of its own, so these files are the pin: every
array in them was produced by the reference's own code (test_tadgan -> univariate_anomaly_detection, or the
named reference function called directly) under this container's torch/numpy/scipy/pandas.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

os.environ["PYTORCH_JIT"] = "0"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np

from oracle import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
T0, DT = 1285027200, 21600


def config1_signal(T=8640):
    """sine + one spike burst (MinMax scaling is done by the reference's SignalDataset)."""
    t = np.arange(T)
    s = np.sin(2 * np.pi * t / 50.0)
    s[T // 2:T // 2 + 5] += 3
    return s, T0 + DT * t


def noisy_signal(T, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(T)
    s = np.sin(2 * np.pi * t / 37.0) + 0.3 * rng.standard_normal(T) + 0.002 * t
    k = int(T * 0.6)
    s[k:k + 4] += 4
    return s, T0 + DT * t


def subset_rows(n):
    idx = sorted(set(list(range(min(n, 64))) + list(range(max(0, n - 30), n)) + [n // 2, n // 3]))
    return np.asarray([i for i in idx if 0 <= i < n], dtype=np.int64)


def save_weights(name, w):
    np.savez(os.path.join(OUT, name), **w)


def run_case(name, values, ts, hyperbolic, combination, rec_error="dtw", full_rows=False, extra_spy=True):
    import utils.anomaly_detection_utils as adu

    spy = {}
    orig_dtw = adu._dtw_error

    def dtw_spy(y, y_hat, score_window=10):
        out = orig_dtw(y, y_hat, score_window)
        spy["true"], spy["pred"], spy["dtw_raw"] = np.asarray(y).copy(), np.asarray(y_hat).copy(), np.asarray(out, dtype=np.float64)
        return out

    adu._dtw_error = dtw_spy
    try:
        cap = rh.run_univariate(values, ts, hyperbolic, combination, rec_error)
    finally:
        adu._dtw_error = orig_dtw
    N = cap["critic"].shape[0]
    S = cap["recons_signal"].shape[1]
    rows = np.arange(N) if full_rows else subset_rows(N)
    g = {
        "signal_raw": np.asarray(values, dtype=np.float64),
        "timestamps": np.asarray(ts, dtype=np.int64),
        "signal": np.concatenate([cap["signal"], [np.nan]]),  # scaled X[0:T-1]; the last sample X[T-1] is in no window
        "index": cap["true_index"].astype(np.int64),
        "critic": cap["critic"],
        "kmax": rh.kde_argmax_reference(cap["critic"], S),
        "critic_scores": cap["critic_scores"],
        "final": cap["final_scores"],
        "intervals": cap["intervals"],
        "rows": rows,
        "recons_rows": cap["recons_signal"][rows],
        "z_head": cap["z_head"],
        "hyperbolic": np.asarray(hyperbolic),
    }
    if hyperbolic:
        g["rec"] = cap["rec_scores"]
        g["unorm"] = np.linalg.norm(cap["recons_signal"], axis=1)
        g["eucl_rows"] = cap["eucl_recons"][rows]
        g["hyper_x_rows"] = cap["real_hyper"][rows]
    else:
        for k in ("rec_point", "rec_area", "rec_dtw"):
            g[k] = cap[k]
        g.update(spy)
    np.savez_compressed(os.path.join(OUT, name), **g)
    print(name, "N=%d" % N, "intervals=%d" % len(cap["intervals"]), "final type", cap["final_scores_type"])
    return cap


def main():
    os.makedirs(OUT, exist_ok=True)
    rh.bootstrap()
    # random-init weights of SURVEY.md 8(d): manual_seed(0); Encoder, Decoder, CriticX
    for hyp, name in ((True, "weights_hyp_s100.npz"), (False, "weights_eucl_s100.npz")):
        save_weights(name, rh.state_dicts(*rh.build_modules(100, hyp, 0)))
    save_weights("weights_hyp_s123.npz", rh.state_dicts(*rh.build_modules(123, True, 0)))

    s, ts = config1_signal()
    run_case("cfg1_hyp_uncertainty.npz", s, ts, True, "uncertainty")
    run_case("cfg2_eucl_dtw_mult.npz", s, ts, False, "mult")
    s, ts = noisy_signal(1500, 1)
    run_case("noisy1500_hyp_uncertainty.npz", s, ts, True, "uncertainty", full_rows=True)
    run_case("noisy1500_hyp_mult.npz", s, ts, True, "mult")
    run_case("noisy1500_eucl_dtw_mult.npz", s, ts, False, "mult", full_rows=True)
    # edge cases: last batch of one window (N=65, anomaly_detection.py:76-88), fewer windows than the window length
    s, ts = noisy_signal(165, 2)
    run_case("edge_n65_hyp.npz", s, ts, True, "uncertainty", full_rows=True)
    run_case("edge_n65_eucl.npz", s, ts, False, "mult", full_rows=True)
    s, ts = noisy_signal(400, 3)
    run_case("edge_n300_hyp.npz", s, ts, True, "uncertainty", full_rows=True)
    # real data shipped with the reference: NASA A-1 test split (data/A-1-test.csv)
    import pandas as pd

    df = pd.read_csv(os.path.join(rh.REF_ROOT, "data", "A-1-test.csv"))
    run_case("a1test_hyp_uncertainty.npz", df["value"].values, df["timestamp"].values, True, "uncertainty")

    # stand-alone pieces of the reference ------------------------------------------------------------------
    import torch
    import utils.anomaly_detection_utils as adu
    from hyperspace.hyrnn_nets import mobius_linear

    rng = np.random.default_rng(7)
    pieces = {}
    # find_anomalies on ndarray (ddof 0) and tensor (ddof 1) inputs, univariate and multivariate parameters
    e = np.abs(rng.standard_normal(5000)) + 1
    e[1200:1210] += 9
    e[3300:3303] += 14
    e[4990:] += 11
    idx = T0 + DT * np.arange(5100)
    pieces["fa_errors"], pieces["fa_index"] = e, idx
    pieces["fa_np_uni"] = np.asarray(adu.find_anomalies(e, idx, window_size_portion=0.33, window_step_size_portion=0.1, fixed_threshold=True), dtype=np.float64)
    pieces["fa_t_uni"] = np.asarray(adu.find_anomalies(torch.from_numpy(e), idx, window_size_portion=0.33, window_step_size_portion=0.1, fixed_threshold=True), dtype=np.float64)
    pieces["fa_np_multi"] = np.asarray(adu.find_anomalies(e, idx, window_size_portion=0.2, window_step_size_portion=0.1, fixed_threshold=True, anomaly_padding=200), dtype=np.float64)
    # _compute_critic_score and rolling means with odd / even / tiny windows
    k = rng.standard_normal(777) * 0.01 - 0.2
    pieces["ccs_in"] = k
    for w in (1, 2, 7, 8, 77):
        pieces["ccs_w%d" % w] = adu._compute_critic_score(k, w)
    # mobius_linear with the path's flags, non-square, with and without hyperbolic bias, near the ball boundary
    x = torch.from_numpy(rng.standard_normal((70, 51)).astype(np.float32))
    W = torch.from_numpy((rng.standard_normal((33, 51)) * 0.05).astype(np.float32))
    b = torch.from_numpy((rng.standard_normal(33) * 0.01).astype(np.float32))
    pieces["ml_x"], pieces["ml_W"], pieces["ml_b"] = x.numpy(), W.numpy(), b.numpy()
    pieces["ml_hb"] = mobius_linear(x, W, b, hyperbolic_input=False, hyperbolic_bias=True).numpy()
    pieces["ml_eb"] = mobius_linear(x, W, b, hyperbolic_input=False, hyperbolic_bias=False).numpy()
    pieces["ml_nb"] = mobius_linear(x, W, None, hyperbolic_input=False).numpy()
    pieces["ml_big"] = mobius_linear(x * 40, W, b, hyperbolic_input=False, hyperbolic_bias=True).numpy()  # projected rows
    # reconstruction_errors of the reference for the three error types on a small random problem
    yw = rng.standard_normal((140, 100, 1))
    yh = (yw[:, :, 0] + 0.1 * rng.standard_normal((140, 100))).astype(np.float32)
    pieces["re_y"], pieces["re_yhat"] = yw, yh
    for kind in ("point", "area", "dtw"):
        err, _ = adu.reconstruction_errors(yw, yh, 1, 10, 3, True, kind)
        pieces["re_" + kind] = np.asarray(err, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "pieces.npz"), **pieces)
    print("pieces ok")


if __name__ == "__main__":
    main()
