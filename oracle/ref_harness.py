"""Runs the UNMODIFIED reference (aleflabo/HypAD under /root/reference) on CPU, with shims.

TEST INFRASTRUCTURE ONLY.  Nothing under hypad_b200/ may import this file; it exists to
(1) pin oracle/hypad_oracle.py (the travelling restatement) against the reference's own code and
(2) generate the committed golden vectors in tests/golden/ (see oracle/make_golden.py).
It only works in the build container, where /root/reference exists.

What is shimmed, and why (nothing in the reference's arithmetic is touched):
  * geoopt / pyts / matplotlib  -> oracle/shims (not installable here; geoopt's math is the
    reference's own math_.py executed in place; pyts.metrics.dtw is a restatement: PARITY UNPINNED)
  * torch.Tensor.cuda / nn.Module.cuda -> identity   (anomaly_detection.py:68-110 hard-codes .cuda())
  * scipy.integrate.trapz -> numpy.trapezoid          (removed from scipy>=1.14; reached at
    utils/anomaly_detection_utils.py:802 when `path` is set)
  * utils.data.load_anomalies -> empty frame          (would download from S3; only feeds the
    swallowed metrics block at utils/anomaly_detection_utils.py:96-110)
  * PYTORCH_JIT=0                                      (TorchScript in torch 2.11 rejects math_.py:1315)
"""
import os
import sys

os.environ.setdefault("PYTORCH_JIT", "0")
if "torch" in sys.modules and os.environ.get("PYTORCH_JIT") != "0":
    raise ImportError("oracle.ref_harness must be imported before torch (needs PYTORCH_JIT=0)")

import argparse
import contextlib
import pickle
import tempfile

import numpy as np

REF_ROOT = os.environ.get("HYPAD_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_BOOTSTRAPPED = False


def available():
    return os.path.exists(os.path.join(REF_ROOT, "anomaly_detection.py"))


def bootstrap():
    """Make `import anomaly_detection`, `models.tadgan`, `utils.*`, `hyperspace.*` resolve to the reference."""
    global _BOOTSTRAPPED
    if _BOOTSTRAPPED:
        return
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    for name in ("models", "utils", "hyperspace", "anomaly_detection", "geoopt", "pyts"):
        if name in sys.modules:
            raise RuntimeError("module %r already imported: the harness needs a fresh interpreter" % name)
    sys.path.insert(0, _SHIMS)
    sys.path.insert(0, REF_ROOT)
    import scipy.integrate
    import torch

    if not hasattr(scipy.integrate, "trapz"):
        scipy.integrate.trapz = np.trapezoid
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import pandas as pd
    import utils.data as od

    od.load_anomalies = lambda signal, edges=False: pd.DataFrame({"start": [], "end": []})
    _BOOTSTRAPPED = True


def build_modules(signal_shape=100, hyperbolic=True, seed=0, latent=20):
    """Weights protocol of SURVEY.md 8(d): manual_seed, then Encoder, Decoder, CriticX on CPU, eval()."""
    bootstrap()
    import torch
    from models.tadgan import CriticX, Decoder, Encoder

    torch.manual_seed(seed)
    enc = Encoder(signal_shape, latent)
    dec = Decoder(signal_shape, latent, hyperbolic)
    cx = CriticX(signal_shape, latent)
    return enc.eval(), dec.eval(), cx.eval()


def state_dicts(enc, dec, cx):
    out = {}
    for pre, m in (("encoder.", enc), ("decoder.", dec), ("critic_x.", cx)):
        for k, v in m.state_dict().items():
            out[pre + k] = v.detach().cpu().numpy().copy()
    return out


def make_dataset(values, timestamps, workdir, interval=21600, name="signal"):
    """CSV -> the reference's own SignalDataset (aggregate, impute, MinMax(-1,1), rolling windows)."""
    bootstrap()
    import pandas as pd
    from utils.dataloader import SignalDataset

    path = os.path.join(workdir, name + ".csv")
    pd.DataFrame({"timestamp": np.asarray(timestamps), "value": np.asarray(values)}).to_csv(path, index=False)
    return SignalDataset(path=path, interval=interval, test=True), path


def run_univariate(values, timestamps, hyperbolic, combination, rec_error="dtw", seed=0, interval=21600,
                   batch_size=64, num_workers=0, modules=None, keep_rows=None):
    """test_tadgan -> univariate_anomaly_detection of the reference, capturing every intermediate."""
    bootstrap()
    import torch
    from torch.utils.data import DataLoader

    import anomaly_detection as ad
    import utils.anomaly_detection_utils as adu

    cap = {}
    with tempfile.TemporaryDirectory() as work:
        ds, csv_path = make_dataset(values, timestamps, work, interval)
        S = ds.X.shape[1]
        enc, dec, cx = modules or build_modules(S, hyperbolic, seed)
        loader = DataLoader(ds, batch_size=batch_size, drop_last=False, shuffle=False, num_workers=num_workers)
        params = argparse.Namespace(dataset="MSL", signal="signal", hyperbolic=hyperbolic, signal_shape=S,
                                    rec_error=rec_error, combination=combination, load=False,
                                    save_result=False, filename="", interval=interval)
        out_path = os.path.join(work, "out")
        os.makedirs(out_path)

        orig_find, orig_comb = adu.find_anomalies, adu.combine_scores

        def find_spy(errors, index, *a, **k):
            cap["final_scores_type"] = type(errors).__module__ + "." + type(errors).__name__
            cap["final_scores"] = np.asarray(errors.detach().numpy() if hasattr(errors, "detach") else errors).copy()
            res = orig_find(errors, index, *a, **k)
            cap["intervals"] = np.asarray(res, dtype=np.float64).reshape(-1, 3) if len(res) else np.zeros((0, 3))
            return res

        def comb_spy(combination, critic_scores=[], rec_scores=[], recons_signal=[]):
            cap["critic_scores_trunc"] = np.asarray(critic_scores).copy()
            cap["rec_scores"] = np.asarray(rec_scores.detach().numpy() if hasattr(rec_scores, "detach") else rec_scores).copy()
            return orig_comb(combination, critic_scores, rec_scores, recons_signal)

        adu.find_anomalies, adu.combine_scores = find_spy, comb_spy
        try:
            with torch.no_grad(), contextlib.redirect_stdout(open(os.devnull, "w")):
                ad.test_tadgan(loader, enc, dec, cx, read_path=csv_path, signal="signal", path=out_path,
                               signal_shape=S, params=params)
        finally:
            adu.find_anomalies, adu.combine_scores = orig_find, orig_comb

        p = out_path + "/"
        cap["X"] = ds.X.reshape(ds.X.shape[0], S)[:, 0].copy()  # first sample of each window
        cap["signal"] = np.concatenate([ds.X[:, 0, 0], ds.X[-1, 1:, 0]])  # scaled signal X[0:T-1]
        cap["index"] = np.asarray(ds.index).copy()
        cap["recons_signal"] = np.asarray(torch.load(p + "recons_signal.pt", weights_only=False))
        cap["critic"] = np.asarray(torch.load(p + "critic_score.pt", weights_only=False), dtype=np.float32)
        cap["true_index"] = np.asarray(torch.load(p + "true_index.pt", weights_only=False))
        if hyperbolic:
            cap["eucl_recons"] = np.asarray(torch.load(p + "eucl_recons.pt", weights_only=False))
            cap["real_hyper"] = np.asarray(torch.load(p + "real_hyper.pt", weights_only=False))
        if os.path.exists(p + "critic_scores.pickle"):  # absent for combination in (rec, rec_uncertainty): :68-82
            with open(p + "critic_scores.pickle", "rb") as fh:
                cap["critic_scores"] = np.asarray(pickle.load(fh))
        for ret in ("point", "area", "dtw"):
            if os.path.exists(p + ret + ".pickle"):
                with open(p + ret + ".pickle", "rb") as fh:
                    cap["rec_" + ret] = np.asarray(pickle.load(fh))
        if os.path.exists(p + "anomalies.csv"):
            import pandas as pd

            cap["anomalies_csv"] = pd.read_csv(p + "anomalies.csv").values[:, 1:].astype(np.float64)
        cap["weights"] = state_dicts(enc, dec, cx)
        with torch.no_grad():
            x = torch.from_numpy(ds.X[: (keep_rows or 64)])
            cap["z_head"] = enc(x.float()).numpy()[0]
    return cap


def kde_argmax_reference(critic, S):
    """critic_kde_max of utils/anomaly_detection_utils.py:372-400, literally (before _compute_critic_score)."""
    bootstrap()
    from scipy import stats

    critic_extended = list()
    for c in critic:
        critic_extended.extend(np.repeat(c, S).tolist())
    critic_extended = np.asarray(critic_extended).reshape((-1, S))
    out = []
    num_errors = S + (len(critic) - 1)
    for i in range(num_errors):
        inter = [critic_extended[i - j, j] for j in range(max(0, i - num_errors + S), min(i + 1, S))]
        if len(inter) > 1:
            d = np.asarray(inter)
            try:
                out.append(d[np.argmax(stats.gaussian_kde(d)(inter))])
            except np.linalg.LinAlgError:
                out.append(np.median(d))
        else:
            out.append(np.median(np.asarray(inter)))
    return np.asarray(out)
