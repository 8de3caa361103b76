"""Pipe-rate peaks measured on this box ("the builder must micro-benchmark"): FP32 FFMA, FP64 DFMA, MUFU.EX2
(fp32 and packed f16x2) and SHFL, through hypad_peak_probe (csrc/peaks.cu), CUDA-event timed, best of 5 after a warm-up.
Writes profiles/peaks.json (read by bench.py for the KDE kernel's roofline).  Run on the B200: python scripts/measure_peaks.py"""
import ctypes
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypad_b200 import _native  # noqa: E402

KINDS = [("fp32_ffma", 0, 2.0, "TFLOP/s (2 flop per FFMA)"), ("fp64_dfma", 1, 2.0, "TFLOP/s (2 flop per DFMA)"),
         ("mufu_ex2_f32", 2, 1.0, "T ex2/s"), ("mufu_ex2_f16x2", 3, 2.0, "T ex2/s (two per instruction)"),
         ("shfl_idx_b32", 4, 1.0, "T lane-shuffles/s")]


def main():
    lib = _native.load_library()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {"gpu": torch.cuda.get_device_name(0), "sms": torch.cuda.get_device_properties(0).multi_processor_count,
           "how": "hypad_peak_probe: 2 CTAs x 1024 threads per SM, 8 independent chains per thread, CUDA events, best of 5"}
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True)
        out["sm_max_mhz"] = float(q.stdout.split()[0])
    except Exception:
        out["sm_max_mhz"] = None
    for name, kind, per, unit in KINDS:
        iters = 4096 if kind != 1 else 1024
        n = ctypes.c_longlong(0)
        best = None
        for rep in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _native.check(lib.hypad_peak_probe(kind, iters, 2, ctypes.c_void_p(sink.data_ptr()), ctypes.byref(n), stream))
            b.record()
            b.synchronize()
            ms = a.elapsed_time(b)
            if rep and (best is None or ms < best):
                best = ms
        rate = n.value * per / (best * 1e-3) / 1e12
        out[name] = {"value": rate, "unit": unit, "ms": best, "thread_instructions": n.value}
        if out.get("sm_max_mhz"):
            out[name]["per_sm_per_clk"] = rate * 1e12 / per / (out["sms"] * out["sm_max_mhz"] * 1e6)
        print(name, out[name])
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "peaks.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "peaks.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
