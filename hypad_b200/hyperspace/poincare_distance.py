"""Drop-in for hyperspace/poincare_distance.py of the reference: `poincare_distance`, `square_norm`, `pairwise_distances`.

Same names and arguments as hyperspace/poincare_distance.py:5-48; fp32 like the reference.  Each function is one C-ABI call
into csrc/pairwise.cu (row norms + a tiled N x M x D contraction with the clamp / acosh epilogue fused); there is no CPU path.
Off the executed scoring path: the reference's only caller is hyperspace/losses.py:154.
"""
import torch

from .. import _native
from .._native import check, ptr


def _rows(t, name):
    _native.require_cuda(t, name)
    if t.dim() != 2:
        raise ValueError("hypad_b200: %s must be a 2-D tensor (rows, D), got shape %s" % (name, tuple(t.shape)))
    t = t.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


def poincare_distance(pred, gt):
    """Pair-wise Poincare distance between the rows of `pred` (N_pred, D) and `gt` (N_gt, D) -> (N_pred, N_gt).
    hyperspace/poincare_distance.py:5-16."""
    p, g = _rows(pred, "pred"), _rows(gt, "gt")
    if p.shape[1] != g.shape[1]:
        raise ValueError("hypad_b200: pred and gt differ in their last dimension (%d vs %d)" % (p.shape[1], g.shape[1]))
    out = torch.empty((p.shape[0], g.shape[0]), dtype=torch.float32, device=p.device)
    c = _native.default_context(p.device)
    with torch.cuda.device(p.device):
        check(c.lib.hypad_poincare_distance_pairwise(c.handle, ptr(p), p.shape[0], ptr(g), g.shape[0], p.shape[1], ptr(out), c.stream()))
    return out


def square_norm(x):
    """clamp(|x|^2, min=1e-5) over the last dimension.  hyperspace/poincare_distance.py:19-25."""
    _native.require_cuda(x, "x")
    xr = _rows(x.reshape(-1, x.shape[-1]), "x")
    out = torch.empty(xr.shape[0], dtype=torch.float32, device=xr.device)
    c = _native.default_context(xr.device)
    with torch.cuda.device(xr.device):
        check(c.lib.hypad_square_norm(ptr(xr), xr.shape[0], xr.shape[1], ptr(out), c.stream()))
    return out.reshape(x.shape[:-1])


def pairwise_distances(x, y=None):
    """dist[i, j] = clamp(|x_i - y_j|^2, 1e-7, inf) as |x_i|^2 + |y_j|^2 - 2 <x_i, y_j>; y=None means y=x.
    hyperspace/poincare_distance.py:28-48."""
    xr = _rows(x, "x")
    yr = xr if y is None else _rows(y, "y")
    if xr.shape[1] != yr.shape[1]:
        raise ValueError("hypad_b200: x and y differ in their last dimension (%d vs %d)" % (xr.shape[1], yr.shape[1]))
    out = torch.empty((xr.shape[0], yr.shape[0]), dtype=torch.float32, device=xr.device)
    c = _native.default_context(xr.device)
    with torch.cuda.device(xr.device):
        check(c.lib.hypad_pairwise_sqdist(c.handle, ptr(xr), xr.shape[0], ptr(yr), yr.shape[0], xr.shape[1], ptr(out), c.stream()))
    return out
