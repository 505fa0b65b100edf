"""Small driver for ncu: a few batched SGBM computes on device-resident synthetic pairs (no other kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stereo_depth_ruler_b200 import capi, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = capi.Context(0)
pairs = [synth.stereo_pair(1280, 720, 128, s % 2) for s in range(2)]
L = torch.from_numpy(np.stack([pairs[i % 2][0] for i in range(B)])).cuda()
R = torch.from_numpy(np.stack([pairs[i % 2][1] for i in range(B)])).cuda()
O = torch.empty((B, 720, 1280), dtype=torch.int16, device="cuda")
m = capi.Sgbm(ctx, capi.make_params(mode=mode))
# profile the kernels the multi-handle benchmark runs (one handle alone would pick k_path_hh); extra knobs: name=value args
m.set_option("shared_device", 1)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    m.set_option(k, int(v))
m.set_profiling(True)
for r in range(reps):
    m.compute_batch(L, R, O)
    ctx.synchronize()
print({k: round(v, 4) for k, v in m.stage_times()})
