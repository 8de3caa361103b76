"""Drop-in for hyperspace/hyrnn_nets.py of the reference: `mobius_linear` and `MobiusLinear`.

Same names, arguments and defaults as hyperspace/hyrnn_nets.py:13-35 and :154-207.  The forward is one
fused sm_100a kernel (GEMM + expmap0 + mobius_add + project with row reductions; csrc/forward.cu)
reached through the C-ABI call `hypad_mobius_linear`.  Only the configuration that exists on the scoring
path is implemented -- hyperbolic_input=False, nonlin=None, k=-1 (models/tadgan.py:43-52) -- and there is
no CPU path; anything else raises instead of silently computing something different.  The Mobius GRU,
MobiusDist2Hyperplane and mobius_matvec of the reference file are dead code there and are
not provided.

geoopt is not needed: the reference only uses it for the arithmetic (now in the kernel) and to tag the bias
as a ManifoldParameter for its Riemannian optimiser (training, out of scope).
"""
import math

import torch
import torch.nn

from .. import _native


def _expmap0_cpu(u):
    """expmap0 at k=-1 (math_.py:1132-1136) -- used once, at construction, for the bias initialisation."""
    n = u.norm(dim=-1, p=2, keepdim=True).clamp_min(1e-15)
    return n.clamp(-15, 15).tanh() * (u / n)


def mobius_linear(input, weight, bias=None, hyperbolic_input=True, hyperbolic_bias=True, nonlin=None, k=-1.0):
    """project(mobius_add(expmap0(input @ weight.T), bias)) -- hyperspace/hyrnn_nets.py:13-35."""
    if hyperbolic_input:
        raise NotImplementedError("hypad_b200: mobius_linear(hyperbolic_input=True) (Mobius matvec) is not on the "
                                  "HypAD scoring path (models/tadgan.py:43-52 uses hyperbolic_input=False)")
    if nonlin is not None:
        raise NotImplementedError("hypad_b200: mobius_linear(nonlin=...) is not on the HypAD scoring path")
    if float(k) != -1.0:
        raise NotImplementedError("hypad_b200: only curvature k=-1 is implemented (the reference never changes it)")
    _native.require_cuda(input, "input")
    _native.require_cuda(weight, "weight")
    out_f, in_f = weight.shape
    x = input.reshape(-1, in_f)
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.float().contiguous()
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    b = None
    if bias is not None:
        b = _native.require_cuda(bias, "bias").detach()
        if b.dtype != torch.float32 or not b.is_contiguous():
            b = b.float().contiguous()
    out = torch.empty((x.shape[0], out_f), dtype=torch.float32, device=x.device)
    ctx = _native.default_context(x.device)
    with torch.cuda.device(x.device):
        _native.check(ctx.lib.hypad_mobius_linear(ctx.handle, _native.ptr(x), x.shape[0], in_f, out_f, _native.ptr(w),
                                                  _native.ptr(b), int(bool(hyperbolic_bias)), _native.ptr(out), ctx.stream()))
    return out.reshape(*input.shape[:-1], out_f)


class MobiusLinear(torch.nn.Linear):
    """hyperspace/hyrnn_nets.py:154-207.  Parameters: weight (out,in), bias (out,) stored on the Poincare ball."""

    def __init__(self, *args, hyperbolic_input=True, hyperbolic_bias=True, nonlin=None, k=-1.0, fp64_hyper=True, **kwargs):
        k = torch.tensor(k)
        super().__init__(*args, **kwargs)
        # Same RNG consumption as the reference constructor (:166-179): Linear init, bias.normal_(), weight.normal_().
        if self.bias is not None and hyperbolic_bias:
            with torch.no_grad():
                self.bias.set_(_expmap0_cpu(self.bias.normal_() / 400))
        with torch.no_grad():
            std = 1 / math.sqrt(2 * self.weight.shape[0] * self.weight.shape[1]) / 100
            self.weight.normal_(std=std)
        self.hyperbolic_bias = hyperbolic_bias
        self.hyperbolic_input = hyperbolic_input
        self.nonlin = nonlin
        self.k = k
        self.fp64_hyper = fp64_hyper

    def forward(self, input):
        if self.fp64_hyper:
            raise NotImplementedError("hypad_b200: MobiusLinear(fp64_hyper=True) is not on the HypAD scoring path "
                                      "(models/tadgan.py:51 builds it with fp64_hyper=False)")
        return mobius_linear(input, weight=self.weight, bias=self.bias, hyperbolic_input=self.hyperbolic_input,
                             nonlin=self.nonlin, hyperbolic_bias=self.hyperbolic_bias, k=self.k)

    def extra_repr(self):
        return "{}, c=1.0, hyperbolic_input={}, hyperbolic_bias={}".format(super().extra_repr(), self.hyperbolic_input,
                                                                           self.hyperbolic_bias)
