"""Prints the per-role cycle breakdown of forward_tc_kernel (CTA 0) on the bench workload."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_signal
from hypad_b200 import _native, scoring
from hypad_b200.models.tadgan import Encoder, Decoder, CriticX
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mods = [Encoder(100, 20), Decoder(100, 20, True), CriticX(100, 20)]
for m in mods: m.to(dev).eval()
sc = scoring.WindowScorer(*mods)
sig = torch.from_numpy(make_signal(1000000)).to(dev)
lib = _native.load_library()
for _ in range(3): sc.forward(sig, True)
_native.check(lib.hypad_forward_debug_cycles(sc.net.ctx.handle, 1, None))
sc.forward(sig, True); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 48)()
_native.check(lib.hypad_forward_debug_cycles(sc.net.ctx.handle, 0, buf))
v = list(buf)
tot = max(v[0], 1)
print("tiles %d  epilogue total %.0f cyc/tile | wait acc %.1f%% | xload+handover %.1f%% | work %.1f%%" % (v[6], tot / max(v[6], 1), 100 * v[1] / tot, 100 * v[5] / tot, 100 * (tot - v[1] - v[5]) / tot))
print("MMA warp: wait operands %.1f%%  wait weights %.1f%% of epilogue total;  producer wait slots %.1f%%" % (100 * v[2] / tot, 100 * v[3] / tot, 100 * v[4] / tot))
names = "ENC_GI ENC_O Z D0 L0_GI L0_O L1_GI L1_O D2 MR MX C1 C2 C3 C4".split()
nt = max(v[6], 1)
for p, nm in enumerate(names):
    print("%-6s wait acc %7.0f  work %7.0f cyc/tile" % (nm, v[8 + p] / nt, v[24 + p] / nt))
print("Mobius row phase (both calls, cyc/tile): " + "  ".join("%s %d" % (n, v[40 + i] / nt) for i, n in enumerate(["load+sum1", "reduce1", "scalars(tanh)", "expmap+sums", "reduce2", "mobius_add", "reduce3+proj", "store"])))
print("Mobius row phase (both calls, cyc/tile): " + "  ".join("%s %d" % (n, v[40 + i] / nt) for i, n in enumerate(["tmem load", "row sums", "reduce", "scalars(tanh)", "mobius_add+sum", "reduce+proj", "store"])))
