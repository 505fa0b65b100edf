"""One matcher call against the oracle, everything from the command line (fuzz-failure triage; a crashed CUDA context is sticky, so
one process per case): python tests/fuzz/repro_case.py W H B kind "minD,D,bs,P1,P2,d12,cap,uq,sw,sr,mode" [knob=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import orc
from stereo_depth_ruler_b200 import capi, synth
W, H, B, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
args = tuple(int(v) for v in sys.argv[5].split(","))
ctx = capi.Context(0)
pairs = [synth.stereo_pair(W, H, args[1], 7 + b) if kind == "tex" else synth.noise_pair(W, H, 7 + b, 256 if kind == "noise" else 8) for b in range(B)]
m = capi.Sgbm(ctx, capi.make_params(*args))
for kv in sys.argv[6:]:
    k, v = kv.split("=")
    m.set_option(k, int(v))
try:
    if B == 1:
        got = [m.compute(*pairs[0])]
    else:
        out = m.compute_batch(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]), np.empty((B, H, W), np.int16))
        got = list(out)
    ctx.synchronize()
except Exception as e:
    print("EXC", str(e)[:160])
    sys.exit(2)
bad = [i for i, (l, r) in enumerate(pairs) if not np.array_equal(orc.sgbm_compute(l, r, orc.make_params(*args)), got[i])]
print("OK" if not bad else "MISMATCH frames %s" % bad)
