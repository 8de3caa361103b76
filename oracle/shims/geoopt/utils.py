"""The five helpers math_.py:26 imports from geoopt.utils (geoopt==0.5.0), restated from the
published package because they are not in the reference tree.  Only `sabs` is reached on the hot
path (math_.py:226 tan_k, :347 _project) where, for k=-1 in fp32, it is numerically a no-op."""
from typing import List

import torch


def sign(x):
    return torch.sign(x.sign() + 0.5)


def sabs(x, eps: float = 1e-15):
    return x.abs().add_(eps)


def clamp_abs(x, eps: float = 1e-15):
    s = sign(x)
    return s * sabs(x, eps=eps)


def list_range(end: int) -> List[int]:
    return list(range(end))


def drop_dims(tensor: torch.Tensor, dims: List[int]):
    seen = 0
    for d in dims:
        tensor = tensor.squeeze(d - seen)
        seen += 1
    return tensor
