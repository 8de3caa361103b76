"""Drop-in for `test_tadgan` of anomaly_detection.py:20-155 of the reference.

Same signature.  Instead of looping over DataLoader batches of 64 windows with four blocking device-to-host copies
per batch (anomaly_detection.py:67-113), all windows of the dataset go through the fused sm_100a pipeline in one
call; the same artefacts are written to `path` (recons_signal.pt, gt_signal.pt, critic_score.pt, true_index.pt,
eucl_recons.pt, real_hyper.pt, critic_scores.pickle, anomalies.csv; anomaly_detection.py:116-131,
utils/anomaly_detection_utils.py:97-98, :234-235).  Ground-truth loading (`utils.data.load_anomalies`: S3 / bundled label
files, anomaly_detection.py:136-150) is not reproduced; with the labels in hand the evaluation itself is available as
`utils.anomaly_detection_utils.contextual_confusion_matrix` / `compute_metrics`, or through the `known_anomalies` argument of
`univariate_anomaly_detection`.
"""
import pickle

import numpy as np
import pandas as pd
import torch

from .scoring import WindowScorer, cuda_device


def _dataset_windows(test_loader):
    ds = getattr(test_loader, "dataset", test_loader)
    X = np.asarray(ds.X)
    index = np.asarray(getattr(ds, "index", np.arange(X.shape[0] + X.shape[1])))
    return X, index


def test_tadgan(test_loader, encoder, decoder, critic_x, read_path="", signal="", path="", signal_shape=100, params=[]):
    path += "/"
    dev = cuda_device()
    for m in (encoder, decoder, critic_x):
        m.to(dev).eval()
    X, index = _dataset_windows(test_loader)
    n = X.shape[0]
    windows = torch.from_numpy(np.ascontiguousarray(X.reshape(n, -1))).to(dev)
    scorer = WindowScorer(encoder, decoder, critic_x)
    multivariate = params.signal == "multivariate"
    keep = ("eucl", "hyper", "hyper_x") if decoder.hyperbolic else ("eucl",)
    out = scorer.score(windows, False, params.combination, params.rec_error, index=None if multivariate else index, keep=keep,
                       multivariate=multivariate)
    recons = (out["hyper"] if decoder.hyperbolic else out["eucl"]).cpu().numpy()
    torch.save(recons, path + "recons_signal.pt")
    torch.save(X, path + "gt_signal.pt")
    torch.save([np.float32(v) for v in out["critic"].cpu().numpy()], path + "critic_score.pt")
    torch.save(torch.from_numpy(index), path + "true_index.pt")
    if decoder.hyperbolic:
        torch.save(out["eucl"].cpu().numpy(), path + "eucl_recons.pt")
        torch.save(out["hyper_x"].cpu().numpy(), path + "real_hyper.pt")
    if out.get("critic_scores_full") is not None or out.get("critic_scores") is not None:
        cs = out.get("critic_scores_full", out.get("critic_scores"))
        with open(path + "critic_scores.pickle", "wb") as handle:
            pickle.dump(cs.cpu().numpy(), handle, protocol=pickle.HIGHEST_PROTOCOL)
    if multivariate:
        from .utils.anomaly_detection_utils import find_anomalies

        x_index = 1353715200.0 + np.arange(n, dtype=np.float64)
        intervals = find_anomalies(out["final"].cpu().numpy(), x_index, window_size_portion=0.2, window_step_size_portion=0.1,
                                   fixed_threshold=True, anomaly_padding=200)
        pd.DataFrame(intervals, columns=["start", "end", "score"]).to_csv(path + "pred_anomalies.csv")
    else:
        intervals = out["intervals"]
        pd.DataFrame(intervals, columns=["start", "end", "score"]).to_csv(path + "anomalies.csv")
    return {"final_scores": out["final"].cpu().numpy(), "intervals": intervals}
