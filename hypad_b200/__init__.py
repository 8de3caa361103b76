"""hypad_b200 -- B200-native (sm_100a) implementation of HypAD's windowed anomaly-scoring hot path.

Layout
  csrc/                      hand-written CUDA kernels + the C-ABI (include/hypad_b200.h) -> libhypad_b200.so
  _native.py, _weights.py    ctypes binding, weight packing
  scoring.py                 device-side steps and the fused WindowScorer pipeline
  models/, hyperspace/, utils/, anomaly_detection.py
                             host-side mirror of the reference's modules: same names and signatures as
                             models/tadgan.py, hyperspace/hyrnn_nets.py, utils/anomaly_detection_utils.py,
                             utils/dataloader.py, anomaly_detection.py of aleflabo/HypAD
  distributed.py             window sharding across GPUs: one process per GPU, the finish sharded, small stage exchanges through
                             NVLink peer memory (or NCCL)
  sweep.py                   many signals, one model each, sharded by signal
  dropin.py                  registers the mirror under the reference's import names

There is no CPU fallback anywhere in this package.
"""
from ._native import HypadError, LIB_PATH, load_library  # noqa: F401

__version__ = "0.1.0"
