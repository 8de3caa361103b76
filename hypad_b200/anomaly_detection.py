"""Drop-in for `test_tadgan` of anomaly_detection.py:20-155 of the reference.

Same signature.  Instead of looping over DataLoader batches of 64 windows with four blocking device-to-host copies
per batch (anomaly_detection.py:67-113), all windows of the dataset go through the fused sm_100a pipeline in one
call; the same artefacts are written to `path` (recons_signal.pt, gt_signal.pt, critic_score.pt, true_index.pt,
eucl_recons.pt, real_hyper.pt, critic_scores.pickle, anomalies.csv; anomaly_detection.py:116-131,
utils/anomaly_detection_utils.py:97-98, :234-235), in the reference's formats: numpy arrays / a list of numpy float32
scalars / the index as handed over, pickled by torch.save.  `params.load` is honoured like the reference does (:51-60, and
critic_scores.pickle at utils/anomaly_detection_utils.py:229-231): cached tensors are read back and only the scoring
tail runs.  Like the reference, anomalies.csv only appears when at least one interval was found (its DataFrame
constructor raises on the empty result inside a try block, utils/anomaly_detection_utils.py:96-110).
Ground-truth loading (`utils.data.load_anomalies`: S3 / bundled label files, anomaly_detection.py:32-37) is not
reproduced; with the labels in hand the evaluation itself is available as
`utils.anomaly_detection_utils.contextual_confusion_matrix` / `compute_metrics`, or through the `known_anomalies`
argument of `univariate_anomaly_detection`.
"""
import os

import numpy as np
import torch

from .scoring import WindowScorer, cuda_device
from .utils import anomaly_detection_utils as adu


def _dataset_windows(test_loader):
    ds = getattr(test_loader, "dataset", test_loader)
    X = np.asarray(ds.X)
    index = getattr(ds, "index", None)
    if index is None:
        index = np.arange(X.shape[0] + X.shape[1])
    return X, index


def _index_tensor(index):
    """`index[0]` of the collated batch (anomaly_detection.py:123, :133): a tensor for numeric indices, else as given."""
    if isinstance(index, torch.Tensor):
        return index
    a = np.asarray(index)
    return torch.from_numpy(a) if a.dtype.kind in "iufb" else index


def test_tadgan(test_loader, encoder, decoder, critic_x, read_path="", signal="", path="", signal_shape=100, params=[]):
    path += "/"
    dev = cuda_device()
    for m in (encoder, decoder, critic_x):
        m.to(dev).eval()
    multivariate = params.signal == "multivariate"
    if getattr(params, "load", False) and os.path.exists(path + "critic_score.pt") and os.path.exists(path + "recons_signal.pt"):
        # :51-60, as the reference does it -- including that with a hyperbolic model the cached `gt_signal.pt` holds the raw
        # windows, not their Mobius images (real_hyper.pt), so the reference's cached path scores against the raw windows
        recons_signal = torch.load(path + "recons_signal.pt", weights_only=False)
        true_signal = torch.load(path + "gt_signal.pt", weights_only=False)
        critic_score = torch.load(path + "critic_score.pt", weights_only=False)
        true_index = torch.load(path + "true_index.pt", weights_only=False)
    else:
        X, index = _dataset_windows(test_loader)
        n = X.shape[0]
        windows = torch.from_numpy(np.ascontiguousarray(X.reshape(n, -1))).to(dev)
        scorer = WindowScorer(encoder, decoder, critic_x)
        keep = ("eucl", "hyper", "hyper_x") if decoder.hyperbolic else ("eucl",)
        fw = scorer.forward(windows, False, keep)
        scorer.poll_error()
        recons_signal = (fw["hyper"] if decoder.hyperbolic else fw["eucl"]).cpu().numpy()
        critic_score = [np.float32(v) for v in fw["critic"].cpu().numpy()]
        torch.save(recons_signal, path + "recons_signal.pt")
        torch.save(X, path + "gt_signal.pt")
        torch.save(critic_score, path + "critic_score.pt")
        true_index = _index_tensor(index)
        torch.save(true_index, path + "true_index.pt")
        true_signal = X
        if decoder.hyperbolic:
            true_signal = fw["hyper_x"].cpu().numpy()
            torch.save(fw["eucl"].cpu().numpy(), path + "eucl_recons.pt")
            torch.save(true_signal, path + "real_hyper.pt")
    if multivariate:
        return adu.multivariate_anomaly_detection(recons_signal, true_signal, params, params.combination, critic_score, path)
    return adu.univariate_anomaly_detection(recons_signal, true_signal, params, params.combination, critic_score, path, read_path,
                                            params.rec_error, true_index, None, signal, signal_shape)
