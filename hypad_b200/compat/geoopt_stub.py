"""Just enough of the `geoopt` namespace to unpickle reference checkpoints without geoopt installed.

The reference saves whole modules (`torch.save(decoder, ...)`, train.py:381-385); a hyperbolic Decoder's pickle names
`geoopt.tensor.ManifoldParameter` (the MobiusLinear bias, hyperspace/hyrnn_nets.py:169) and
`geoopt.manifolds.stereographic.manifold.PoincareBall` (`self.ball`).  Neither carries arithmetic that the scoring path
needs: the bias is used as a plain tensor and the curvature is fixed at -1."""
import sys
import types

import torch


class ManifoldParameter(torch.nn.Parameter):
    def __new__(cls, data=None, manifold=None, requires_grad=True):
        if data is None:
            data = torch.empty(0)
        inst = torch.nn.Parameter._make_subclass(cls, data.data if isinstance(data, torch.nn.Parameter) else data, requires_grad)
        inst.manifold = manifold
        return inst

    def __reduce_ex__(self, proto):
        return _rebuild, (self.data, getattr(self, "manifold", None), self.requires_grad)


class ManifoldTensor(torch.Tensor):
    pass


def _rebuild(data, manifold, requires_grad):
    return ManifoldParameter(data, manifold=manifold, requires_grad=requires_grad)


class _Manifold(torch.nn.Module):
    def __init__(self, c=1.0, **kw):
        super().__init__()
        self.c = torch.as_tensor(c)


class PoincareBall(_Manifold):
    pass


class Stereographic(_Manifold):
    pass


def install():
    """Registers the stub as `geoopt` (and the sub-module paths pickles refer to) unless the real package imports."""
    try:
        import geoopt  # noqa: F401

        return False
    except Exception:
        pass
    root = types.ModuleType("geoopt")
    root.__doc__ = __doc__
    names = {"geoopt": root}
    for sub in ("tensor", "manifolds", "manifolds.stereographic", "manifolds.stereographic.manifold", "manifolds.base"):
        names["geoopt." + sub] = types.ModuleType("geoopt." + sub)
    for mod in names.values():
        mod.ManifoldParameter, mod.ManifoldTensor = ManifoldParameter, ManifoldTensor
        mod.PoincareBall, mod.Stereographic = PoincareBall, Stereographic
        mod._rebuild_manifold_parameter = _rebuild
    root.tensor = names["geoopt.tensor"]
    root.manifolds = names["geoopt.manifolds"]
    names["geoopt.manifolds"].stereographic = names["geoopt.manifolds.stereographic"]
    names["geoopt.manifolds.stereographic"].manifold = names["geoopt.manifolds.stereographic.manifold"]
    sys.modules.update(names)
    return True
