"""Mounts the reference's own math_.py (vendored geoopt stereographic math) as
geoopt.manifolds.stereographic.math by executing the file where it lies under the reference
tree.  Needs PYTORCH_JIT=0: torch 2.11's TorchScript rejects math_.py:1315 (off-path)."""
import importlib.util
import os
import sys

_REF = os.environ.get("HYPAD_REFERENCE_ROOT", "/root/reference")
_src = os.path.join(_REF, "math_.py")
if not os.path.exists(_src):
    raise ImportError("geoopt shim needs the reference tree (math_.py) at %s" % _REF)
if os.environ.get("PYTORCH_JIT", "1") != "0":
    raise ImportError("geoopt shim must be imported with PYTORCH_JIT=0 set before `import torch`")
_spec = importlib.util.spec_from_file_location(__name__ + ".math", _src)
math = importlib.util.module_from_spec(_spec)
sys.modules[__name__ + ".math"] = math
_spec.loader.exec_module(math)
