"""ctypes binding of libhypad_b200.so (include/hypad_b200.h).

PyTorch is used for device memory and streams only: tensors are passed as raw `data_ptr()`s, the
current torch CUDA stream as `cudaStream_t`.  There is no CPU fallback: if the library is missing
or no CUDA device is present every compute call raises.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhypad_b200.so")

STAGE_ENCODER, STAGE_DECODER, STAGE_MOBIUS_X, STAGE_CRITIC = 1, 2, 4, 8
STAGE_ALL = 15
ABI_VERSION = 2
STATS_F32 = 16  # HYPAD_STATS_F32: OR-ed into the ddof argument of the thresholding entry points

COMBINE_MODES = {"mult": 0, "uncertainty": 1, "sum": 2, "critic": 3, "critic_uncertainty": 4, "sum_uncertainty": 5,
                 "rec": 6, "rec_uncertainty": 7, "euclidean_sum": 8}

c_f32p = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_vp = ctypes.c_void_p


class HypadError(RuntimeError):
    pass


class hypad_weights(ctypes.Structure):
    _fields_ = [
        ("signal_shape", ctypes.c_int32), ("latent_dim", ctypes.c_int32), ("hyperbolic", ctypes.c_int32),
        ("critic_dim", ctypes.c_int32),
        ("enc_w_ih", _vp * 2), ("enc_b_ih", _vp * 2), ("enc_b_hh", _vp * 2),
        ("enc_dense_w", _vp), ("enc_dense_b", _vp),
        ("dec_dense1_w", _vp), ("dec_dense1_b", _vp),
        ("dec_w_ih", (_vp * 2) * 2), ("dec_b_ih", (_vp * 2) * 2), ("dec_b_hh", (_vp * 2) * 2),
        ("dec_dense2_w", _vp), ("dec_dense2_b", _vp),
        ("mobius_w", _vp), ("mobius_b", _vp),
        ("critic_w", _vp * 5), ("critic_b", _vp * 5),
    ]


class hypad_forward_out(ctypes.Structure):
    _fields_ = [("z", _vp), ("eucl", _vp), ("hyper", _vp), ("hyper_x", _vp), ("critic", _vp), ("rec", _vp), ("unorm", _vp)]


class hypad_signal_out(ctypes.Structure):
    _fields_ = [("critic", _vp), ("rec", _vp), ("unorm", _vp), ("kmax", _vp), ("critic_scores", _vp), ("final", _vp), ("tw", _vp)]


class hypad_signal_eucl_out(ctypes.Structure):
    _fields_ = [("critic", _vp), ("eucl", _vp), ("kmax", _vp), ("critic_scores", _vp), ("truth", _vp), ("pred", _vp), ("errors", _vp),
                ("rec", _vp), ("final", _vp), ("tw", _vp)]


# name -> (restype, argtypes); mirrors include/hypad_b200.h one to one
_SIGNATURES = {
    "hypad_abi_version": (_int, []),
    "hypad_last_error": (ctypes.c_char_p, []),
    "hypad_launch_count": (ctypes.c_int64, []),
    "hypad_ctx_create": (_int, [ctypes.POINTER(_vp), _int]),
    "hypad_ctx_destroy": (_int, [_vp]),
    "hypad_pack_weights": (_int, [_vp, ctypes.POINTER(hypad_weights), _vp]),
    "hypad_forward": (_int, [_vp, _vp, _int, _i64, _i64, _vp, _int, ctypes.POINTER(hypad_forward_out), _vp]),
    "hypad_forward_ffma": (_int, [_vp, _vp, _int, _i64, _i64, _vp, _int, ctypes.POINTER(hypad_forward_out), _vp]),
    "hypad_ctx_poll_error": (_int, [_vp]),
    "hypad_ctx_set_strict_range": (_int, [_vp, _int]),
    "hypad_ctx_range_fallbacks": (ctypes.c_int64, [_vp]),
    "hypad_segments_aggregate": (_int, [_vp, _vp, _i64, _vp, ctypes.c_double, _i64, _vp, _vp]),
    "hypad_impute_minmax": (_int, [_vp, _vp, _i64, ctypes.c_double, ctypes.c_double, _vp, _vp]),
    "hypad_detrend_linear": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "hypad_poincare_distance_pairwise": (_int, [_vp, _vp, _i64, _vp, _i64, _int, _vp, _vp]),
    "hypad_pairwise_sqdist": (_int, [_vp, _vp, _i64, _vp, _i64, _int, _vp, _vp]),
    "hypad_square_norm": (_int, [_vp, _i64, _int, _vp, _vp]),
    "hypad_rowdiff_norm": (_int, [_vp, _int, _vp, _i64, _int, _vp, _vp]),
    "hypad_forward_debug_cycles": (_int, [_vp, _int, _vp]),
    "hypad_mobius_linear": (_int, [_vp, _vp, _i64, _int, _int, _vp, _vp, _int, _vp, _vp]),
    "hypad_poincare_rowdist": (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    "hypad_rownorm": (_int, [_vp, _i64, _int, _vp, _vp]),
    "hypad_window_gather": (_int, [_vp, _i64, _int, _vp, _int, _vp]),
    "hypad_kde_argmax_overlap": (_int, [_vp, _i64, _i64, _i64, _int, _i64, _i64, _vp, _vp]),
    "hypad_kde_argmax_overlap_exhaustive": (_int, [_vp, _i64, _i64, _i64, _int, _i64, _i64, _vp, _vp]),
    "hypad_critic_zscore_smooth": (_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "hypad_rolling_mean_centered": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "hypad_zscore_clip": (_int, [_vp, _vp, _int, _i64, _vp, _vp]),
    "hypad_combine_scores": (_int, [_int, _vp, _vp, _int, _vp, ctypes.c_double, _i64, _vp, _vp]),
    "hypad_median_overlap": (_int, [_vp, _i64, _int, _vp, _vp]),
    "hypad_true_from_signal": (_int, [_vp, _int, _i64, _i64, _int, _vp, _vp]),
    "hypad_dtw_error": (_int, [_vp, _vp, _int, _i64, _int, _vp, _vp]),
    "hypad_point_error": (_int, [_vp, _vp, _int, _i64, _vp, _vp]),
    "hypad_area_error": (_int, [_vp, _vp, _int, _i64, _int, _vp, _vp]),
    "hypad_threshold_windows": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _vp, _vp, _vp, _int, _vp]),
    "hypad_threshold_windows_range": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _int, _int, _vp, _vp, _vp, _int, _vp]),
    "hypad_threshold_windows_exhaustive": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _vp, _vp, _vp, _int, _vp]),
    "hypad_tc_probe_gemm": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _vp]),
    "hypad_tc_probe_bench": (_int, [_int, _int, _int, _vp, _vp]),
    "hypad_score_signal_hyperbolic": (_int, [_vp, _vp, _int, _i64, _int, _i64, _i64, _i64, _int, _int, _int,
                                             ctypes.POINTER(hypad_signal_out), _vp]),
    "hypad_score_signals_hyperbolic": (_int, [_vp, _i64, _int, _int, _int, _int, _int, _vp, _int]),
    "hypad_score_signal_euclidean": (_int, [_vp, _vp, _int, _i64, _int, _int, ctypes.c_double, _i64, _i64, _i64, _int, _int, _int,
                                            ctypes.POINTER(hypad_signal_eucl_out), _vp]),
    "hypad_critic_small_max": (_int, []),
    "hypad_critic_combine_small": (_int, [_vp, _vp, _i64, _i64, _int, _int, _vp, _vp, _i64, _vp, _vp, _vp]),
    "hypad_critic_scores": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _vp]),
    "hypad_stats_select_passes": (_int, [_int]),
    "hypad_stats_select_begin": (_int, [_vp, _i64, _int, _vp]),
    "hypad_stats_select_hist": (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    "hypad_stats_select_pick": (_int, [_vp, _vp, _int, _int, _vp]),
    "hypad_stats_moments_partial": (_int, [_vp, _vp, _int, _i64, _int, _vp, _vp]),
    "hypad_stats_moments_final": (_int, [_vp, _vp, _int, _i64, _int, _int, _vp]),
    "hypad_stats_read": (_int, [_vp, _vp, _vp]),
    "hypad_critic_smooth_shard": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    "hypad_rolling_mean_shard": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    "hypad_zscore_clip_apply": (_int, [_vp, _vp, _int, _i64, _vp, _vp]),
    "hypad_intervals_from_runs": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, ctypes.c_double, _int, _vp, _i64, ctypes.POINTER(_i64)]),
    "hypad_tw_shard_record_doubles": (ctypes.c_size_t, [_i64, _i64, _int]),
    "hypad_tw_shard_pack": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _int, _i64, _vp, _vp]),
    "hypad_tw_shard_runs": (_int, [_vp, _vp, _int, _int, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _vp, _vp]),
    "hypad_tw_shard_merge": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _i64, ctypes.POINTER(_i64), ctypes.POINTER(_int)]),
    "hypad_peer_exchange": (_int, [_vp, _i64, _vp, _int, _int, _i64, _i64, ctypes.c_uint64, _vp, _vp]),
    "hypad_sweep_intervals": (_int, [_vp, _i64, _vp, _vp, _vp, _int, ctypes.c_double, _int, _vp, _i64, _vp, ctypes.POINTER(_i64)]),
    "hypad_peak_probe": (_int, [_int, _int, _int, _vp, ctypes.POINTER(ctypes.c_longlong), _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lock = threading.Lock()


def load_library():
    """dlopen the in-tree library; raises HypadError (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise HypadError(
                    "hypad_b200: %s is missing -- build it with `python -m hypad_b200.build` "
                    "(there is no CPU or PyTorch fallback)" % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the header and the library disagree
                fn.restype = res
                fn.argtypes = args
            if lib.hypad_abi_version() != ABI_VERSION:
                raise HypadError("hypad_b200: ABI version mismatch")
            _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load_library().hypad_last_error()
        raise HypadError("hypad_b200 error %d: %s" % (rc, msg.decode(errors="replace") if msg else "?"))


def require_cuda(t, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise HypadError("hypad_b200: %s must be a CUDA tensor (the scoring path has no CPU implementation)" % name)
    return t


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Context:
    """Owns one hypad_ctx (packed weights + workspace) on one device."""

    def __init__(self, device):
        self.lib = load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise HypadError("hypad_b200: a CUDA device is required, got %s" % self.device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.empty(1, device=self.device)  # make sure the primary context exists
            h = ctypes.c_void_p()
            check(self.lib.hypad_ctx_create(ctypes.byref(h), idx))
        self.handle = h
        self._keepalive = None

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.hypad_ctx_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stream(self):
        return stream_ptr(self.device)

    def pack(self, w, keepalive):
        with torch.cuda.device(self.device):
            check(self.lib.hypad_pack_weights(self.handle, ctypes.byref(w), self.stream()))
        self._keepalive = keepalive


_default_ctx = {}


def default_context(device):
    """A context (workspace only) for the stateless entry points: one per device AND per current stream, because calls on one
    context must be issued on one stream at a time (the workspace is shared) -- pipelines running side by side on different
    streams (sweep.SignalSweep) each get their own."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    ctx = _default_ctx.get(key)
    if ctx is None:
        ctx = _default_ctx[key] = Context(torch.device("cuda", idx))
    return ctx
