"""Measurement, on the B200, of what a tcgen05 (TF32, TMEM accumulator) contraction does to the numbers:
layout/descriptor correctness on exactly representable inputs, then the error of 1xTF32 / 3xTF32 / 6-term split
products against an fp64 product, next to the error of the fp32 FFMA order the fused kernel uses."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def probe(A, B, pieces, terms, device):
    from hypad_b200 import _native

    lib = _native.load_library()
    a = torch.from_numpy(np.ascontiguousarray(A, dtype=np.float32)).to(device)
    b = torch.from_numpy(np.ascontiguousarray(B, dtype=np.float32)).to(device)
    d = torch.full((128, B.shape[0]), float("nan"), dtype=torch.float32, device=device)
    _native.check(lib.hypad_tc_probe_gemm(_native.ptr(a), _native.ptr(b), _native.ptr(d), A.shape[1], B.shape[0], pieces, terms,
                                          _native.stream_ptr(device)))
    torch.cuda.synchronize()
    return d.cpu().numpy()


def ffma_order(A, B):
    """fp32 product in ascending-k fused multiply-add order (what forward_kernel does)."""
    acc = np.zeros((A.shape[0], B.shape[0]), np.float32)
    for k in range(A.shape[1]):
        acc = (A[:, k:k + 1].astype(np.float64) * B[None, :, k].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    return acc


@pytest.mark.parametrize("K,N", [(8, 16), (16, 64), (104, 192), (128, 128), (56, 192), (24, 64)])
def test_tcgen05_layout_exact_on_small_integers(K, N, cuda_device):
    rng = np.random.default_rng(K * 1000 + N)
    A = rng.integers(-8, 9, (128, K)).astype(np.float32)
    B = rng.integers(-8, 9, (N, K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    got = probe(A, B, 1, 1, cuda_device)
    assert np.array_equal(got.astype(np.float64), want)


def test_tcgen05_split_product_accuracy(cuda_device, capsys):
    rng = np.random.default_rng(1)
    rows = []
    for K, N in ((104, 64), (128, 64), (56, 192), (64, 64)):
        A = rng.uniform(-1, 1, (128, K)).astype(np.float32)
        B = (rng.standard_normal((N, K)) * 0.1).astype(np.float32)
        exact = A.astype(np.float64) @ B.astype(np.float64).T
        scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T  # sum |a||b|
        variants = {"ffma_fp32": ffma_order(A, B), "tf32x1": probe(A, B, 1, 1, cuda_device), "tf32x3": probe(A, B, 2, 3, cuda_device)}
        if K * (128 + N) * 4 * 3 <= 200 * 1024:
            variants["tf32x6"] = probe(A, B, 3, 6, cuda_device)
        for name, got in variants.items():
            err = (got.astype(np.float64) - exact) / scale
            rows.append((K, N, name, np.abs(err).max(), np.sqrt((err ** 2).mean()), err.mean()))
    with capsys.disabled():
        print("\n  K   N  variant      max|err|/sum|a||b|   rms        mean (bias)")
        for r in rows:
            print("%4d %4d  %-10s  %.3e           %.3e  %+.3e" % r)
    by = {(r[0], r[1], r[2]): r for r in rows}
    for K, N in ((104, 64), (128, 64), (56, 192)):
        assert by[(K, N, "tf32x1")][3] < 2e-3          # plain TF32: ~2^-11 per operand
        assert by[(K, N, "tf32x3")][3] < 2e-6          # compensated product: fp32 class
