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


def _rebuild(*args):
    """geoopt.tensor._rebuild_manifold_parameter.  geoopt 0.5.0 pickles a ManifoldParameter as
    `_rebuild_manifold_parameter(*tensor_rebuild_args, cls, manifold, requires_grad)`, the leading arguments being those of
    `torch._utils._rebuild_tensor_v2` (storage, offset, size, stride, requires_grad, hooks[, metadata]); this stub's own
    pickles (and those of the oracle's shim) carry `(data, manifold, requires_grad)`.  Both layouts are accepted; the class named
    in the pickle is replaced by the stub's, the manifold object is kept as an attribute and otherwise ignored."""
    if len(args) == 3 and isinstance(args[0], torch.Tensor):
        data, manifold, requires_grad = args
    else:
        if len(args) < 7:
            raise TypeError("geoopt stub: unknown ManifoldParameter pickle layout (%d arguments)" % len(args))
        from torch._utils import _rebuild_tensor_v2

        data = _rebuild_tensor_v2(*args[:-3])
        manifold, requires_grad = args[-2], args[-1]
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
