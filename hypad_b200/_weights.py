"""Collects the parameters of Encoder / Decoder / CriticX modules into a packed device context.

The modules keep ordinary nn.LSTM / nn.Linear children as parameter holders (state-dict and pickle
compatibility with the reference, models/tadgan.py:10-106); this file hands their raw fp32 device
pointers to hypad_pack_weights and re-packs when a parameter changes (data_ptr or in-place version).
"""
import weakref

import torch

from . import _native
from ._native import HypadError

ENC_HIDDEN, DEC_HIDDEN, DEC_DENSE1 = 50, 64, 50  # fixed by models/tadgan.py:15-20, :35-38


def _f32(t, device, keep):
    t = t.detach()
    if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32).contiguous()
    keep.append(t)
    return t


def _expect(t, shape, name):
    if tuple(t.shape) != tuple(shape):
        raise HypadError("hypad_b200: %s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    return t


class PackedNet:
    """hypad_ctx + the identity of the parameters it was packed from."""

    def __init__(self, device):
        self.ctx = _native.Context(device)
        self.device = self.ctx.device
        self.key = None
        self.S = self.latent = self.critic_dim = None
        self.hyperbolic = False

    def _param_slots(self, mods):
        """[(owner module, parameter name)] in sorted-name order per top-level module, cached: walking the module trees on every
        forward call cost more host time than the launch itself on short signals.  The cache is valid while every module of the
        walk still has the same children (a replaced sub-module or top-level module rebuilds it); a replaced Parameter object
        is picked up because the slot is read through its owner each time."""
        shape = tuple(id(m) for m in mods)
        cached = getattr(self, "_slots", None)
        if cached is not None and cached[0] == shape and all(tuple(map(id, mod._modules.values())) == kids for mod, kids in cached[1]):
            return cached[2]
        walk, slots = [], []
        for m in mods:
            if m is None:
                continue
            named = dict(m.named_modules())
            walk.extend((mod, tuple(map(id, mod._modules.values()))) for mod in named.values())
            for full, _p in sorted(m.named_parameters()):
                owner, _, leaf = full.rpartition(".")
                slots.append((named[owner], leaf))
        self._slots = (shape, walk, slots)
        return slots

    def ensure(self, encoder, decoder, critic_x):
        params = [owner._parameters[name] for owner, name in self._param_slots((encoder, decoder, critic_x))]
        # identity of the packed weights: storage address and in-place version of every parameter (a dtype / device change
        # replaces the storage); 43 parameters, checked on every forward call -- kept to two attribute reads each
        key = tuple([(p.data_ptr(), p._version) for p in params])
        if key == self.key:
            return self
        dev = self.device
        keep = []
        S = latent = cdim = None
        if encoder is not None:
            S, latent = int(encoder.signal_shape), int(encoder.latent_space_dim)
        if decoder is not None:
            S = int(decoder.signal_shape) if S is None else S
            latent = int(decoder.latent_space_dim) if latent is None else latent
            if int(decoder.signal_shape) != S or int(decoder.latent_space_dim) != latent:
                raise HypadError("hypad_b200: encoder and decoder disagree on signal_shape / latent_space_dim")
        if critic_x is not None:
            cdim = int(critic_x.latent_space_dim)
            if S is None:
                S = int(critic_x.signal_shape)
            elif int(critic_x.signal_shape) != S:
                raise HypadError("hypad_b200: CriticX.signal_shape %d != %d" % (critic_x.signal_shape, S))
        latent = 20 if latent is None else latent
        cdim = 20 if cdim is None else cdim
        hyperbolic = bool(decoder is not None and getattr(decoder, "hyperbolic", False))

        def zeros(*shape):
            t = torch.zeros(*shape, dtype=torch.float32, device=dev)
            keep.append(t)
            return t

        w = _native.hypad_weights()
        w.signal_shape, w.latent_dim, w.hyperbolic, w.critic_dim = S, latent, int(hyperbolic), cdim
        H = ENC_HIDDEN
        for d, suf in enumerate(("", "_reverse")):
            if encoder is not None:
                wi = _expect(_f32(getattr(encoder.lstm, "weight_ih_l0" + suf), dev, keep), (4 * H, S), "encoder.lstm.weight_ih_l0" + suf)
                bi = _expect(_f32(getattr(encoder.lstm, "bias_ih_l0" + suf), dev, keep), (4 * H,), "encoder.lstm.bias_ih_l0" + suf)
                bh = _expect(_f32(getattr(encoder.lstm, "bias_hh_l0" + suf), dev, keep), (4 * H,), "encoder.lstm.bias_hh_l0" + suf)
            else:
                wi, bi, bh = zeros(4 * H, S), zeros(4 * H), zeros(4 * H)
            w.enc_w_ih[d], w.enc_b_ih[d], w.enc_b_hh[d] = wi.data_ptr(), bi.data_ptr(), bh.data_ptr()
        if encoder is not None:
            ew = _expect(_f32(encoder.dense.weight, dev, keep), (latent, 2 * H), "encoder.dense.weight")
            eb = _expect(_f32(encoder.dense.bias, dev, keep), (latent,), "encoder.dense.bias")
        else:
            ew, eb = zeros(latent, 2 * H), zeros(latent)
        w.enc_dense_w, w.enc_dense_b = ew.data_ptr(), eb.data_ptr()

        HD = DEC_HIDDEN
        if decoder is not None:
            d1w = _expect(_f32(decoder.dense1.weight, dev, keep), (DEC_DENSE1, latent), "decoder.dense1.weight")
            d1b = _expect(_f32(decoder.dense1.bias, dev, keep), (DEC_DENSE1,), "decoder.dense1.bias")
            d2w = _expect(_f32(decoder.dense2.weight, dev, keep), (S, 2 * HD), "decoder.dense2.weight")
            d2b = _expect(_f32(decoder.dense2.bias, dev, keep), (S,), "decoder.dense2.bias")
        else:
            d1w, d1b, d2w, d2b = zeros(DEC_DENSE1, latent), zeros(DEC_DENSE1), zeros(S, 2 * HD), zeros(S)
        w.dec_dense1_w, w.dec_dense1_b, w.dec_dense2_w, w.dec_dense2_b = d1w.data_ptr(), d1b.data_ptr(), d2w.data_ptr(), d2b.data_ptr()
        for layer, kin in ((0, DEC_DENSE1), (1, 2 * HD)):
            for d, suf in enumerate(("", "_reverse")):
                if decoder is not None:
                    n = "l%d%s" % (layer, suf)
                    wi = _expect(_f32(getattr(decoder.lstm, "weight_ih_" + n), dev, keep), (4 * HD, kin), "decoder.lstm.weight_ih_" + n)
                    bi = _expect(_f32(getattr(decoder.lstm, "bias_ih_" + n), dev, keep), (4 * HD,), "decoder.lstm.bias_ih_" + n)
                    bh = _expect(_f32(getattr(decoder.lstm, "bias_hh_" + n), dev, keep), (4 * HD,), "decoder.lstm.bias_hh_" + n)
                else:
                    wi, bi, bh = zeros(4 * HD, kin), zeros(4 * HD), zeros(4 * HD)
                w.dec_w_ih[layer][d], w.dec_b_ih[layer][d], w.dec_b_hh[layer][d] = wi.data_ptr(), bi.data_ptr(), bh.data_ptr()
        if hyperbolic:
            hl = decoder.hyperbolic_linear
            if not (hl.hyperbolic_bias and not hl.hyperbolic_input and hl.nonlin is None and float(hl.k) == -1.0):
                raise HypadError("hypad_b200: only MobiusLinear(hyperbolic_input=False, hyperbolic_bias=True, nonlin=None, k=-1) "
                                 "is on the accelerated path (models/tadgan.py:43-52)")
            mw = _expect(_f32(hl.weight, dev, keep), (S, S), "decoder.hyperbolic_linear.weight")
            mb = _expect(_f32(hl.bias, dev, keep), (S,), "decoder.hyperbolic_linear.bias")
            w.mobius_w, w.mobius_b = mw.data_ptr(), mb.data_ptr()
        else:
            w.mobius_w = w.mobius_b = None
        dims = [(cdim, S), (cdim, cdim), (cdim, cdim), (cdim, cdim), (1, cdim)]
        for i, (o, k) in enumerate(dims):
            if critic_x is not None:
                lin = getattr(critic_x, "dense%d" % (i + 1))
                cw = _expect(_f32(lin.weight, dev, keep), (o, k), "critic_x.dense%d.weight" % (i + 1))
                cb = _expect(_f32(lin.bias, dev, keep), (o,), "critic_x.dense%d.bias" % (i + 1))
            else:
                cw, cb = zeros(o, k), zeros(o)
            w.critic_w[i], w.critic_b[i] = cw.data_ptr(), cb.data_ptr()
        self.ctx.pack(w, keep)
        self.key = key
        self.S, self.latent, self.critic_dim, self.hyperbolic = S, latent, cdim, hyperbolic
        return self


_CACHE = {}


def packed_net(encoder=None, decoder=None, critic_x=None, device=None):
    """Cached PackedNet for this combination of live modules (weakly keyed by module identity)."""
    mods = tuple(m for m in (encoder, decoder, critic_x))
    first = next(m for m in mods if m is not None)
    if device is None:
        device = next(first.parameters()).device
    device = torch.device(device)
    if device.type != "cuda":
        raise HypadError("hypad_b200: module parameters are on %s; move the modules to a CUDA device "
                         "(there is no CPU implementation of the scoring path)" % device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = tuple(id(m) if m is not None else 0 for m in mods) + (device.index,)
    entry = _CACHE.get(key)
    if entry is not None and all((r() is m) if m is not None else r is None for r, m in zip(entry[1], mods)):
        net = entry[0]
    else:
        net = PackedNet(device)
        refs = tuple(weakref.ref(m, lambda _r, k=key: _CACHE.pop(k, None)) if m is not None else None for m in mods)
        _CACHE[key] = (net, refs)
    return net.ensure(encoder, decoder, critic_x)
