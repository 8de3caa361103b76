"""Drop-in for models/tadgan.py of the reference: Encoder, Decoder, CriticX (accelerated) and CriticZ (kept).

Constructor arguments, attribute names, sub-module names and the order in which sub-modules are created
follow models/tadgan.py:10-132 so that (a) state dicts and whole-module pickles are interchangeable and
(b) `torch.manual_seed(s)` followed by Encoder, Decoder, CriticX construction yields the same random-init
weights as the reference.  The nn.LSTM / nn.Linear children only hold parameters: `forward` runs the fused
sm_100a kernel through the C-ABI (`hypad_forward`, csrc/forward.cu).  Inference only (eval mode, no autograd,
dropout off) and CUDA only -- there is no CPU implementation.
"""
import torch
import torch.nn as nn

from .. import _native, _weights
from .._native import HypadError
from ..hyperspace.hyrnn_nets import MobiusLinear


def _rows(x, width, what):
    """`x.view(1, -1, width)` semantics: any shape whose element count is a multiple of `width` -> (N, width)."""
    _native.require_cuda(x, what)
    if x.numel() % width:
        raise RuntimeError("shape '[1, -1, %d]' is invalid for input of size %d" % (width, x.numel()))
    x = x.detach().reshape(-1, width)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.float()
    return x.contiguous()


def _no_training(module):
    if module.training:
        raise HypadError("hypad_b200: %s is in training mode; the B200 path implements eval-mode scoring only "
                         "(anomaly_detection.py:46-48 calls .eval() on all three modules)" % type(module).__name__)


def _forward(net, x, z_in, stages, n, **outs):
    out = _native.hypad_forward_out()
    for name, t in outs.items():
        setattr(out, name, t.data_ptr() if t is not None else None)
    xp, is64, stride = None, 0, 1
    if x is not None:
        xp, is64, stride = _native.ptr(x), int(x.dtype == torch.float64), x.shape[1]
    with torch.cuda.device(net.device):
        _native.check(net.ctx.lib.hypad_forward(net.ctx.handle, xp, is64, n, stride, _native.ptr(z_in), stages, out,
                                                net.ctx.stream()))


def poll_error(*modules):
    """Synchronises and raises if a kernel launched by one of these modules' forward() reported an error (include/hypad_b200.h,
    hypad_ctx_poll_error).  forward() itself never synchronises -- the reference's `.cpu()` after every batch
    (anomaly_detection.py:90-95) is where a caller naturally places this.  An operand outside the tensor-core kernel's range is
    not an error: that call was redone by the FFMA kernel on the device."""
    for m in modules:
        kw = {"encoder": m} if isinstance(m, Encoder) else {"decoder": m} if isinstance(m, Decoder) else {"critic_x": m}
        net = _weights.packed_net(**kw)
        _native.check(net.ctx.lib.hypad_ctx_poll_error(net.ctx.handle))


class Encoder(nn.Module):
    """BiLSTM(signal_shape -> 2x50, sequence length 1) + Linear(100 -> latent): models/tadgan.py:10-27."""

    def __init__(self, signal_shape=100, latent_space_dim=20, hyperbolic=False):
        super().__init__()
        self.signal_shape = signal_shape
        self.latent_space_dim = latent_space_dim
        self.lstm = nn.LSTM(input_size=signal_shape, hidden_size=50, num_layers=1, bidirectional=True)
        self.dense = nn.Linear(in_features=100, out_features=latent_space_dim)

    def forward(self, x):
        _no_training(self)
        x = _rows(x, self.signal_shape, "Encoder input")
        net = _weights.packed_net(encoder=self)
        z = torch.empty((x.shape[0], self.latent_space_dim), dtype=torch.float32, device=x.device)
        _forward(net, x, None, _native.STAGE_ENCODER, x.shape[0], z=z)
        return z.view(1, -1, self.latent_space_dim)


class Decoder(nn.Module):
    """Linear(latent -> 50) + 2-layer BiLSTM(50 -> 2x64) + Linear(128 -> S) + tanh [+ MobiusLinear(S -> S)]:
    models/tadgan.py:30-67.  Returns `eucl` or, when hyperbolic, `(hyper, eucl)`, both (1, N, S)."""

    def __init__(self, signal_shape=100, latent_space_dim=20, hyperbolic=False):
        super().__init__()
        self.signal_shape = signal_shape
        self.latent_space_dim = latent_space_dim
        self.dense1 = nn.Linear(in_features=latent_space_dim, out_features=50)
        self.lstm = nn.LSTM(input_size=50, hidden_size=64, num_layers=2, dropout=0.2, bidirectional=True)
        self.dense2 = nn.Linear(in_features=128, out_features=signal_shape)
        self.tanh = nn.Tanh()
        self.hyperbolic = hyperbolic
        if hyperbolic:
            self.hyperbolic_linear = MobiusLinear(signal_shape, signal_shape, hyperbolic_input=False, hyperbolic_bias=True,
                                                  nonlin=None, fp64_hyper=False)

    def forward(self, x):
        _no_training(self)
        z = _rows(x, self.latent_space_dim, "Decoder input").float()
        net = _weights.packed_net(decoder=self)
        n, S = z.shape[0], self.signal_shape
        eucl = torch.empty((n, S), dtype=torch.float32, device=z.device)
        hyper = torch.empty((n, S), dtype=torch.float32, device=z.device) if self.hyperbolic else None
        _forward(net, None, z, _native.STAGE_DECODER, n, eucl=eucl, hyper=hyper)
        if self.hyperbolic:
            return hyper.view(1, -1, S), eucl.view(1, -1, S)
        return eucl.view(1, -1, S)


class CriticX(nn.Module):
    """4 x (Linear + LeakyReLU(0.2) + Dropout) + Linear(latent -> 1): models/tadgan.py:70-106; output (1, N, 1)."""

    def __init__(self, signal_shape=10, latent_space_dim=20):
        super().__init__()
        self.signal_shape = signal_shape
        self.latent_space_dim = latent_space_dim
        self.dropout = nn.Dropout(p=0.25)
        self.leakyrelu = nn.LeakyReLU(0.2)
        self.dense1 = nn.Linear(in_features=signal_shape, out_features=latent_space_dim)
        self.dense2 = nn.Linear(in_features=latent_space_dim, out_features=latent_space_dim)
        self.dense3 = nn.Linear(in_features=latent_space_dim, out_features=latent_space_dim)
        self.dense4 = nn.Linear(in_features=latent_space_dim, out_features=latent_space_dim)
        self.dense5 = nn.Linear(in_features=latent_space_dim, out_features=1)

    def forward(self, x):
        _no_training(self)
        x = _rows(x, self.signal_shape, "CriticX input")
        net = _weights.packed_net(critic_x=self)
        critic = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
        _forward(net, x, None, _native.STAGE_CRITIC, x.shape[0], critic=critic)
        return critic.view(1, -1, 1)


class CriticZ(nn.Module):
    """models/tadgan.py:109-132.  Training-only critic on the latent space: kept for checkpoint compatibility,
    plain PyTorch, not part of the scoring path."""

    def __init__(self, latent_space_dim=20):
        super().__init__()
        self.latent_space_dim = latent_space_dim
        self.dense1 = nn.Linear(in_features=latent_space_dim, out_features=latent_space_dim)
        self.dense2 = nn.Linear(in_features=latent_space_dim, out_features=latent_space_dim)
        self.dense3 = nn.Linear(in_features=latent_space_dim, out_features=1)
        self.dropout = nn.Dropout(p=0.2)
        self.leakyrelu = nn.LeakyReLU(0.2)

    def forward(self, x):
        x = self.dropout(self.leakyrelu(self.dense1(x)))
        x = self.dropout(self.leakyrelu(self.dense2(x)))
        return self.dense3(x)
