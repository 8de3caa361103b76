"""Device functions that are pure arithmetic, compiled for the host (g++) and checked on the CPU: the summation order of
`np_pairwise_sum` in csrc/preprocess.cu decides whether the segment means are bit-identical to pandas' (numpy's add.reduce:
plain loop below 8 elements, 8 interleaved accumulators up to 128, halving to a multiple of 8 beyond)."""
import ctypes
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def host_pairwise(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    src = open(os.path.join(ROOT, "hypad_b200", "csrc", "preprocess.cu")).read()
    a = src.index("__device__ __forceinline__ double np_sum_leaf")
    b = src.index("__global__ void segments_aggregate_kernel")
    body = re.sub(r"__device__|__forceinline__|__restrict__", "", src[a:b])
    d = tmp_path_factory.mktemp("hostk")
    cpp = d / "pw.cpp"
    cpp.write_text('#include <cstdint>\n' + body + '\nextern "C" double pw(const double* v, int64_t n) { return np_pairwise_sum(v, n); }\n')
    so = d / "pw.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(cpp)])
    lib = ctypes.CDLL(str(so))
    lib.pw.restype = ctypes.c_double
    lib.pw.argtypes = [ctypes.c_void_p, ctypes.c_int64]
    return lib


def test_pairwise_sum_walk_is_numpys_order(host_pairwise):
    rng = np.random.default_rng(0)
    sizes = list(range(1, 300)) + [511, 512, 513, 1000, 1023, 1024, 1025, 4097, 65537, 200001, 1 << 20, (1 << 22) + 13]
    for n in sizes:
        v = np.ascontiguousarray(rng.standard_normal(n) * 10.0 ** rng.integers(-3, 6, n))
        assert host_pairwise.pw(v.ctypes.data, n) == np.add.reduce(v), n
    # NaNs count as zeros (pandas' nanmean zeroes them before numpy sums the column)
    v = rng.standard_normal(5000)
    v[rng.integers(0, 5000, 300)] = np.nan
    assert host_pairwise.pw(v.ctypes.data, 5000) == np.add.reduce(np.where(np.isnan(v), 0.0, v))
