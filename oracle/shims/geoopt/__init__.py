"""Minimal stand-in for geoopt==0.5.0 (environment.yml:91 of the reference), TEST INFRASTRUCTURE ONLY.

The reference imports geoopt at hyperspace/hyrnn_nets.py:5-6 and uses
  geoopt.PoincareBall(c=...)            hyrnn_nets.py:168
  geoopt.ManifoldParameter(p, manifold) hyrnn_nets.py:169
  geoopt.manifolds.stereographic.math   hyrnn_nets.py:6  (expmap0 / mobius_add / project)
geoopt is not installed in this image and cannot be installed (no network).  The arithmetic
lives in the reference tree as /root/reference/math_.py (a vendored copy of geoopt's
stereographic math that nothing imports); the submodule `manifolds.stereographic.math` of
this shim executes THAT FILE IN PLACE (it is never copied into this repository).  Only the
five tiny helpers of geoopt/utils.py, absent from the reference tree, are restated here.
"""
import torch


class PoincareBall:
    """Carrier for the curvature only; the reference reads `.c` in extra_repr (hyrnn_nets.py:204)."""

    def __init__(self, c=1.0):
        self.c = torch.as_tensor(c)

    def __repr__(self):
        return "PoincareBall(c={})".format(self.c)


class ManifoldParameter(torch.nn.Parameter):
    """nn.Parameter tagged with a manifold (geoopt.tensor.ManifoldParameter in 0.5.0)."""

    def __new__(cls, data=None, manifold=None, requires_grad=True):
        if data is None:
            data = torch.empty(0)
        inst = torch.nn.Parameter._make_subclass(cls, data.data if isinstance(data, torch.nn.Parameter) else data, requires_grad)
        inst.manifold = manifold
        return inst

    def __reduce_ex__(self, proto):
        return _rebuild_manifold_parameter, (self.data, self.manifold, self.requires_grad)


def _rebuild_manifold_parameter(data, manifold, requires_grad):
    return ManifoldParameter(data, manifold=manifold, requires_grad=requires_grad)


from . import utils  # noqa: E402,F401
from . import manifolds  # noqa: E402,F401
from . import optim  # noqa: E402,F401
