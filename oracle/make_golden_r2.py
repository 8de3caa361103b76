"""Round-2 golden vectors, produced by the UNMODIFIED reference (/root/reference) through oracle/ref_harness.py.

Run in the build container only:   python oracle/make_golden_r2.py [combos] [long] [dropin] [trained]
TEST INFRASTRUCTURE ONLY.

  combos   tests/golden/combos_noisy1500.npz -- every `combination` the reference accepts: the eight hyperbolic ones
           (utils/anomaly_detection_utils.py:336-362) and the Euclidean mult / sum / rec / critic (:554-570) with
           rec_error dtw / point / area, each through test_tadgan -> univariate_anomaly_detection -> find_anomalies.
  long     tests/golden/cfg3_long300k.npz -- BASELINE config 3's signal generator at T = 300,100 (300,000 windows =
           2,344 tiles of the tensor-core kernel, ~8 per CTA slot), hyperbolic / uncertainty, through the reference's
           own dataset, modules, batch loop (64 windows per batch), scipy KDE loop and find_anomalies.
  dropin   tests/golden/dropin_noisy1500/ -- the files test_tadgan writes (anomaly_detection.py:116-131,
           utils/anomaly_detection_utils.py:97-98, :234-235) for the noisy1500 signal, plus whole-module pickles of the
           reference's own classes as train.py:381-385 writes them.
  trained  tests/golden/trained_regime.npz -- the reference's modules with scaled weights (Mobius rows on the 0.996
           ball, gates saturating, critic activations near the tensor path's range limit) on 700 windows.
"""
import os
import sys

os.environ["PYTORCH_JIT"] = "0"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import argparse
import contextlib
import pickle
import shutil
import tempfile
import time

import numpy as np

from oracle import ref_harness as rh
from oracle.make_golden import T0, DT, noisy_signal

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

HYP_COMBOS = ("mult", "uncertainty", "sum", "sum_uncertainty", "critic", "critic_uncertainty", "rec", "rec_uncertainty")
EUCL_COMBOS = (("mult", "dtw"), ("sum", "dtw"), ("rec", "dtw"), ("critic", "dtw"), ("mult", "point"), ("sum", "area"))


def long_signal_raw(T, seed=0):
    """bench.py:make_signal before its MinMax step (the reference's SignalDataset scales): sine + a 5-sample burst every
    50,000 steps."""
    t = np.arange(T, dtype=np.float64)
    s = np.sin(2 * np.pi * t / 50.0)
    rng = np.random.default_rng(seed)
    for k in range(25000, T, 50000):
        s[k:k + 5] += rng.uniform(2, 4)
    return s, T0 + DT * np.arange(T, dtype=np.int64)


class BatchLoader:
    """What test_tadgan iterates over (anomaly_detection.py:67): (sample, index, y, y_index, x_index) batches of 64 windows.
    torch's default collate would stack 64 copies of the (T,) index per batch (utils/dataloader.py:227-232) -- 150 MB per batch
    at this length -- of which the reference reads `index[0]` only (:133); this loader hands it one row."""

    def __init__(self, ds, batch_size=64):
        self.ds, self.batch_size = ds, batch_size

    def __iter__(self):
        import torch

        X = torch.from_numpy(np.asarray(self.ds.X))
        index = torch.from_numpy(np.asarray(self.ds.index))[None]
        for s in range(0, X.shape[0], self.batch_size):
            yield X[s:s + self.batch_size], index, None, None, None


def make_combos(name="combos_noisy1500.npz", signal=None):
    import utils.anomaly_detection_utils as adu  # noqa: F401  (reference)

    s, ts = signal if signal is not None else noisy_signal(1500, 1)
    g = {}
    for comb in HYP_COMBOS:
        try:
            cap = rh.run_univariate(s, ts, True, comb)
        except Exception as e:  # the reference itself fails for this combination: the error is the golden
            g["hyp/%s/error" % comb] = np.asarray("%s: %s" % (type(e).__name__, e))
            print("hyp", comb, "REFERENCE RAISES", type(e).__name__, str(e)[:100])
            continue
        g["hyp/%s/final" % comb] = np.asarray(cap["final_scores"], dtype=np.float64)
        g["hyp/%s/final_is_f32" % comb] = np.asarray(cap["final_scores"].dtype == np.float32)
        g["hyp/%s/intervals" % comb] = cap["intervals"]
        print("hyp", comb, cap["final_scores_type"], cap["final_scores"].dtype, len(cap["intervals"]))
    for comb, rec_error in EUCL_COMBOS:
        cap = rh.run_univariate(s, ts, False, comb, rec_error)
        g["eucl/%s_%s/final" % (comb, rec_error)] = np.asarray(cap["final_scores"], dtype=np.float64)
        g["eucl/%s_%s/intervals" % (comb, rec_error)] = cap["intervals"]
        print("eucl", comb, rec_error, cap["final_scores_type"], len(cap["intervals"]))
    np.savez_compressed(os.path.join(OUT, name), **g)


def make_long(T=300100):
    import torch

    import anomaly_detection as ad
    import utils.anomaly_detection_utils as adu

    values, ts = long_signal_raw(T)
    cap = {}
    t_start = time.time()
    with tempfile.TemporaryDirectory() as work:
        ds, csv_path = rh.make_dataset(values, ts, work, DT)
        print("dataset %.0f s" % (time.time() - t_start), ds.X.shape)
        enc, dec, cx = rh.build_modules(100, True, 0)
        params = argparse.Namespace(dataset="MSL", signal="signal", hyperbolic=True, signal_shape=100, rec_error="dtw",
                                    combination="uncertainty", load=False, save_result=False, filename="", interval=DT)
        out_path = os.path.join(work, "out")
        os.makedirs(out_path)
        orig_find, orig_comb = adu.find_anomalies, adu.combine_scores

        def find_spy(errors, index, *a, **k):
            cap["final"] = np.asarray(errors.detach().numpy() if hasattr(errors, "detach") else errors).copy()
            res = orig_find(errors, index, *a, **k)
            cap["intervals"] = np.asarray(res, dtype=np.float64).reshape(-1, 3)
            return res

        def comb_spy(combination, critic_scores=[], rec_scores=[], recons_signal=[]):
            cap["rec"] = np.asarray(rec_scores.detach().numpy()).copy()
            cap["unorm"] = np.linalg.norm(recons_signal, axis=1)
            return orig_comb(combination, critic_scores, rec_scores, recons_signal)

        adu.find_anomalies, adu.combine_scores = find_spy, comb_spy
        try:
            with torch.no_grad(), contextlib.redirect_stdout(open(os.devnull, "w")):
                ad.test_tadgan(BatchLoader(ds), enc, dec, cx, read_path=csv_path, signal="signal", path=out_path, signal_shape=100,
                               params=params)
        finally:
            adu.find_anomalies, adu.combine_scores = orig_find, orig_comb
        print("reference done %.0f s" % (time.time() - t_start))
        critic = np.asarray(torch.load(out_path + "/critic_score.pt", weights_only=False), dtype=np.float32)
        with open(out_path + "/critic_scores.pickle", "rb") as fh:
            critic_scores = np.asarray(pickle.load(fh))
        signal = np.concatenate([ds.X[:, 0, 0], ds.X[-1, 1:, 0]])
    # the reference never exposes critic_kde_max; it is recovered exactly from its critic_scores by inverting nothing: the
    # literal loop is re-run on its critics (ref_harness.kde_argmax_reference) on three stretches, the rest by the vectorised
    # oracle, and the reference's own critic_scores pin the whole array through _compute_critic_score
    from oracle import hypad_oracle as ho

    kmax = ho.kde_argmax_overlap(critic, 100)
    cs = ho.compute_critic_score(kmax, int(critic.shape[0] * 0.01))
    assert np.allclose(cs, critic_scores, rtol=1e-12, atol=0, equal_nan=True), "oracle kmax does not reproduce the reference's critic_scores"
    assert np.array_equal(kmax.astype(np.float32).astype(np.float64), kmax)
    np.savez_compressed(os.path.join(OUT, "cfg3_long300k.npz"),
                        T=np.asarray(T), signal=np.concatenate([signal, [np.nan]]),  # scaled X[0:T-1] as the reference's dataset made it
                        
                        critic=critic, kmax=kmax.astype(np.float32), rec=cap["rec"].astype(np.float32),
                        unorm=cap["unorm"].astype(np.float32), final=cap["final"], intervals=cap["intervals"])
    print("long ok: N=%d intervals=%d, %.0f s" % (critic.shape[0], len(cap["intervals"]), time.time() - t_start))


def geoopt050_reduce():
    """Gives the shim's ManifoldParameter the pickle layout of geoopt==0.5.0 (geoopt/tensor.py): the rebuild function
    `geoopt.tensor._rebuild_manifold_parameter(*tensor_rebuild_args, cls, manifold, requires_grad)`, where the leading arguments
    are those of torch._utils._rebuild_tensor_v2.  geoopt itself is absent (no network): this is an emulation of its published
    __reduce_ex__, so that the stub in hypad_b200/compat is tested against the layout real checkpoints carry."""
    import types

    import geoopt
    import torch

    mod = types.ModuleType("geoopt.tensor")

    def _rebuild_manifold_parameter(*args):
        tensor = torch._utils._rebuild_tensor_v2(*args[:-3])
        return args[-3](tensor, manifold=args[-2], requires_grad=args[-1])

    _rebuild_manifold_parameter.__module__ = "geoopt.tensor"
    _rebuild_manifold_parameter.__qualname__ = "_rebuild_manifold_parameter"
    mod._rebuild_manifold_parameter = _rebuild_manifold_parameter
    geoopt.ManifoldParameter.__module__ = "geoopt.tensor"
    geoopt.ManifoldParameter.__qualname__ = "ManifoldParameter"
    mod.ManifoldParameter = geoopt.ManifoldParameter
    geoopt.tensor = mod
    sys.modules["geoopt.tensor"] = mod

    def reduce_ex(self, proto):
        build, args = torch.Tensor._reduce_ex_internal(self.data, proto)
        assert build is torch._utils._rebuild_tensor_v2
        return _rebuild_manifold_parameter, tuple(args) + (self.__class__, self.manifold, self.requires_grad)

    geoopt.ManifoldParameter.__reduce_ex__ = reduce_ex
    # the manifold object: geoopt.manifolds.stereographic.manifold.PoincareBall, an nn.Module holding the curvature
    man = types.ModuleType("geoopt.manifolds.stereographic.manifold")

    class PoincareBall(torch.nn.Module):
        def __init__(self, c=1.0):
            super().__init__()
            # geoopt 0.5.0: Stereographic holds the curvature as a (non-learnable) Parameter `k`; PoincareBall(c) sets k = -c
            self.k = torch.nn.Parameter(-torch.as_tensor(c, dtype=torch.get_default_dtype()), requires_grad=False)

        @property
        def c(self):
            return -self.k

    PoincareBall.__module__ = "geoopt.manifolds.stereographic.manifold"
    PoincareBall.__qualname__ = "PoincareBall"
    man.PoincareBall = PoincareBall
    sys.modules["geoopt.manifolds.stereographic.manifold"] = man
    geoopt.PoincareBall = PoincareBall


def make_dropin():
    import torch

    rh.bootstrap()
    geoopt050_reduce()
    s, ts = noisy_signal(1500, 1)  # the noisy1500 case: one detected interval, so anomalies.csv exists (none is written without)
    dst = os.path.join(OUT, "dropin_noisy1500")
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    enc, dec, cx = rh.build_modules(100, True, 0)
    for name, m in (("encoder", enc), ("decoder", dec), ("critic_x", cx)):
        torch.save(m, os.path.join(dst, name + ".pt"))  # train.py:381-385 -- whole-module pickles
    from torch.utils.data import DataLoader

    import anomaly_detection as ad

    with tempfile.TemporaryDirectory() as work:
        ds, csv_path = rh.make_dataset(s, ts, work, DT)
        shutil.copy(csv_path, os.path.join(dst, "signal.csv"))
        loader = DataLoader(ds, batch_size=64, drop_last=False, shuffle=False, num_workers=0)
        params = argparse.Namespace(dataset="MSL", signal="signal", hyperbolic=True, signal_shape=100, rec_error="dtw",
                                    combination="uncertainty", load=False, save_result=False, filename="", interval=DT)
        with torch.no_grad(), contextlib.redirect_stdout(open(os.devnull, "w")):
            ad.test_tadgan(loader, enc, dec, cx, read_path=csv_path, signal="signal", path=dst, signal_shape=100, params=params)
    # gt_signal.pt (the float64 window matrix, 1.1 MB) and true_index.pt (a view whose pickle drags the 64-row collated index
    # along, 0.8 MB) are functions of signal.csv alone; their content is checked against it, the files are not committed
    idx = torch.load(os.path.join(dst, "true_index.pt"), weights_only=False)
    gt = torch.load(os.path.join(dst, "gt_signal.pt"), weights_only=False)
    assert np.array_equal(np.asarray(idx), np.asarray(ds.index)) and np.array_equal(gt, ds.X)
    os.remove(os.path.join(dst, "true_index.pt"))
    os.remove(os.path.join(dst, "gt_signal.pt"))
    print("dropin artefacts:", sorted(os.listdir(dst)))


def trained_regime_weights(sd):
    """Scales the random-init weights into the regime a trained model can reach: Mobius rows pushed onto the 0.996 ball
    (projection branch: 222 of 700 reconstructed rows and 526 of 700 real rows are projected, the others are not), LSTM
    pre-activations x4 (gates saturate), dense2 x6, CriticX hidden activations up to 204 (the tensor path's operand limit for
    linear activations is 255)."""
    sd = {k: v.copy() for k, v in sd.items()}
    sd["decoder.hyperbolic_linear.weight"] *= 1200.0
    sd["decoder.dense2.weight"] *= 6.0
    sd["decoder.dense2.bias"] *= 6.0
    for k in sd:
        if k.startswith("decoder.lstm.weight_ih") or k.startswith("encoder.lstm.weight_ih"):
            sd[k] *= 4.0
    for i in (1, 2, 3, 4):
        sd["critic_x.dense%d.weight" % i] *= 8.4
    return sd


def make_trained():
    import torch

    rh.bootstrap()
    import utils.anomaly_detection_utils as adu  # noqa: F401

    enc, dec, cx = rh.build_modules(100, True, 0)
    sd = trained_regime_weights(rh.state_dicts(enc, dec, cx))
    for pre, m in (("encoder.", enc), ("decoder.", dec), ("critic_x.", cx)):
        m.load_state_dict({k[len(pre):]: torch.from_numpy(v) for k, v in sd.items() if k.startswith(pre)})
    s, ts = noisy_signal(800, 5)
    cap = rh.run_univariate(s, ts, True, "uncertainty", modules=(enc, dec, cx))
    hyper_norm = np.linalg.norm(cap["recons_signal"], axis=1)
    real_norm = np.linalg.norm(cap["real_hyper"], axis=1)
    print("trained regime: |hyper| max %.4f (projected rows: %d of %d), |hyper_x| max %.4f (projected %d), eucl max %.3f, critic range %.1f..%.1f"
          % (hyper_norm.max(), int((hyper_norm > 0.9959).sum()), len(hyper_norm), real_norm.max(), int((real_norm > 0.9959).sum()),
             np.abs(cap["eucl_recons"]).max(), cap["critic"].min(), cap["critic"].max()))
    np.savez_compressed(os.path.join(OUT, "trained_regime.npz"),
                        signal_raw=np.asarray(s, dtype=np.float64), timestamps=np.asarray(ts, dtype=np.int64),
                        signal=np.concatenate([cap["signal"], [np.nan]]), index=cap["true_index"].astype(np.int64),
                        critic=cap["critic"], hyper=cap["recons_signal"], eucl=cap["eucl_recons"], hyper_x=cap["real_hyper"],
                        rec=cap["rec_scores"], final=cap["final_scores"], intervals=cap["intervals"],
                        critic_scores=cap["critic_scores"], kmax=rh.kde_argmax_reference(cap["critic"], 100),
                        **{"w/" + k: v for k, v in sd.items()})


def make_pieces():
    """Reference functions called directly: find_anomalies on a float32 torch tensor (single-precision statistics, what the
    "rec" / "rec_uncertainty" combinations hand it) and combine_scores on ndarray operands (the multivariate path, where all
    eight combinations are defined)."""
    import torch
    import utils.anomaly_detection_utils as adu

    rng = np.random.default_rng(17)
    p = {}
    e = (np.abs(rng.standard_normal(6000)) * 0.2 + 1).astype(np.float32)
    e[1500:1506] += 5
    e[3900:3902] += 9
    e[5990:] += 6
    idx = T0 + DT * np.arange(6100)
    p["fa32_errors"], p["fa32_index"] = e, idx
    p["fa32_uni"] = np.asarray(adu.find_anomalies(torch.from_numpy(e), idx, window_size_portion=0.33, window_step_size_portion=0.1,
                                                  fixed_threshold=True), dtype=np.float64)
    n, S = 900, 123
    cs = 1 + np.abs(rng.standard_normal(n + S - 1))
    rec = np.clip(rng.standard_normal(n), 0, None) + 1
    recons = (rng.standard_normal((n, S)) * 0.05).astype(np.float32)
    p["mc_critic_scores"], p["mc_rec"], p["mc_recons"] = cs, rec, recons
    for comb in HYP_COMBOS:
        p["mc_" + comb] = np.asarray(adu.combine_scores(comb, cs[:n], rec, recons), dtype=np.float64)
    # a merged group whose weights (end - start) sum to zero: np.average raises ZeroDivisionError (:1297)
    try:
        adu._merge_sequences([(5, 5, 1.0), (6, 6, 2.0)])
        p["merge_zero_weights"] = np.asarray("no error")
    except Exception as ex:
        p["merge_zero_weights"] = np.asarray("%s: %s" % (type(ex).__name__, ex))
    np.savez_compressed(os.path.join(OUT, "pieces_r2.npz"), **p)
    print("pieces_r2 ok:", p["fa32_uni"].shape, str(p["merge_zero_weights"]))


def main():
    what = sys.argv[1:] or ["combos", "pieces", "dropin", "trained", "long"]
    os.makedirs(OUT, exist_ok=True)
    rh.bootstrap()
    if "combos" in what:
        make_combos()
        from oracle.make_golden import config1_signal

        make_combos("combos_cfg1.npz", config1_signal())  # BASELINE config 1 / 2 signal: several detected intervals
    if "pieces" in what:
        make_pieces()
    if "trained" in what:
        make_trained()
    if "long" in what:
        make_long()
    if "dropin" in what:  # last: it re-wires the shim's pickling
        make_dropin()


if __name__ == "__main__":
    main()
