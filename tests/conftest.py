import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def weights(name):
    import torch

    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name)).items()}


def build_modules(weights_name, S, hyperbolic, device=None):
    """hypad_b200 modules loaded with the golden random-init weights (torch.manual_seed(0) protocol)."""
    from hypad_b200.models.tadgan import CriticX, Decoder, Encoder

    w = weights(weights_name)
    enc, dec, cx = Encoder(S, 20), Decoder(S, 20, hyperbolic), CriticX(S, 20)
    for pre, m in (("encoder.", enc), ("decoder.", dec), ("critic_x.", cx)):
        m.load_state_dict({k[len(pre):]: v for k, v in w.items() if k.startswith(pre)})
        m.eval()
        if device is not None:
            m.to(device)
    return enc, dec, cx, w


def full_signal(g):
    """Scaled signal X[0:T] of a golden case: the reference keeps X[0:T-1] in its windows; the last sample (in no
    window) is re-derived with the same MinMax map."""
    from oracle import hypad_oracle as ho

    x = ho.minmax_scale(g["signal_raw"]) if "signal_raw" in g else None
    sig = g["signal"].copy()
    if x is not None and x.shape[0] == sig.shape[0]:
        assert np.allclose(x[:-1], sig[:-1], rtol=0, atol=1e-12)
        sig[-1] = x[-1]
    return sig


def long_signal(T, seed=0):
    """BASELINE config 3's generator (bench.py:make_signal): sine + a 5-sample burst every 50,000 steps, MinMax to [-1, 1] the
    way sklearn does it (utils/dataloader.py:88-89); float64 (T,)."""
    from oracle import hypad_oracle as ho

    t = np.arange(T, dtype=np.float64)
    s = np.sin(2 * np.pi * t / 50.0)
    rng = np.random.default_rng(seed)
    for k in range(25000, T, 50000):
        s[k:k + 5] += rng.uniform(2, 4)
    return ho.minmax_scale(s)


def long_golden_signal(g):
    """Scaled signal of tests/golden/cfg3_long300k.npz as the reference's dataset produced it (its CSV round trip moves some
    samples by an ulp against long_signal()); the last sample, which is in no window, comes from the generator."""
    sig = g["signal"].copy()
    gen = long_signal(sig.shape[0])
    assert np.abs(gen[:-1] - sig[:-1]).max() <= 4e-16
    sig[-1] = gen[-1]
    return sig


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
