"""Makes the reference's import names resolve to this package:

    import hypad_b200.dropin; hypad_b200.dropin.install()
    from models.tadgan import Encoder, Decoder, CriticX            # -> hypad_b200.models.tadgan
    from hyperspace.hyrnn_nets import MobiusLinear                 # -> hypad_b200.hyperspace.hyrnn_nets
    from utils.anomaly_detection_utils import univariate_anomaly_detection
    from anomaly_detection import test_tadgan

After install(), `torch.load("encoder.pt", weights_only=False)` of a checkpoint written by the reference's train.py
(whole-module pickles of models.tadgan.*; anomaly_detection.py:214-227) yields modules whose forward runs the sm_100a
kernels.  Run it from a directory that is NOT the reference checkout (its own `models/` would shadow nothing -- the
aliases are placed in sys.modules first -- but relative data paths of the reference obviously do not exist here).
"""
import importlib
import sys

_ALIASES = {
    "models": "hypad_b200.models",
    "models.tadgan": "hypad_b200.models.tadgan",
    "hyperspace": "hypad_b200.hyperspace",
    "hyperspace.hyrnn_nets": "hypad_b200.hyperspace.hyrnn_nets",
    "hyperspace.poincare_distance": "hypad_b200.hyperspace.poincare_distance",
    "utils": "hypad_b200.utils",
    "utils.anomaly_detection_utils": "hypad_b200.utils.anomaly_detection_utils",
    "utils.dataloader": "hypad_b200.utils.dataloader",
    "anomaly_detection": "hypad_b200.anomaly_detection",
}


def install(force=False):
    """Registers the aliases; refuses to replace already-imported foreign modules unless force=True."""
    from .compat import geoopt_stub

    geoopt_stub.install()
    for alias, target in _ALIASES.items():
        mod = importlib.import_module(target)
        cur = sys.modules.get(alias)
        if cur is not None and cur is not mod and not force:
            raise RuntimeError("hypad_b200.dropin: module %r is already imported from %s; call install(force=True) "
                               "before importing the reference's modules" % (alias, getattr(cur, "__file__", "?")))
        sys.modules[alias] = mod
    return sorted(_ALIASES)


def uninstall():
    for alias, target in _ALIASES.items():
        if sys.modules.get(alias) is sys.modules.get(target):
            sys.modules.pop(alias, None)
