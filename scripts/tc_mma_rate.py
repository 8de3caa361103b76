import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hypad_b200 import _native
torch.zeros(1, device="cuda")
lib = _native.load_library()
buf = (ctypes.c_longlong * 2)()
for mode in (0, 3, 4):
    for N in (64, 128):
        for reps in (1020,):
            _native.check(lib.hypad_tc_probe_bench(N, reps, mode, buf, None))
            print("mode %d N %3d reps %4d: %7.1f cyc/MMA retired, %6.1f cyc/MMA issue" % (mode, N, reps, buf[0] / reps, buf[1] / reps))
