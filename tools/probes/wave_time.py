import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from stereo_depth_ruler_b200 import capi, synth
ctx = capi.Context(0)
W, H, args = 1280, 720, (0, 128, 5, 600, 2400, 1, 63, 12, 200, 2, 0)
B = int(os.environ.get("B", 8))
pairs = [synth.stereo_pair(W, H, 128, s) for s in range(2)]
L = torch.from_numpy(np.stack([pairs[i % 2][0] for i in range(B)])).cuda(); R = torch.from_numpy(np.stack([pairs[i % 2][1] for i in range(B)])).cuda()
O = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
ref = None
for setting in sys.argv[1:]:
    m = capi.Sgbm(ctx, capi.make_params(*args)); m.set_option("shared_device", 1); m.set_profiling(True)
    for kv in setting.split():
        k, v = kv.split("="); m.set_option(k, int(v))
    acc = {}
    for rep in range(6):
        m.compute_batch(L, R, O); ctx.synchronize()
        if rep >= 2:
            for k, t in m.stage_times(): acc[k] = acc.get(k, 0) + t / 4
    o = O.cpu().numpy()
    same = "ref" if ref is None else ("== ref" if np.array_equal(o, ref) else "differs")
    ref = o if ref is None else ref
    print("%-28s total %.3f ms (%s)" % (setting, sum(acc.values()), same), {k: round(v, 3) for k, v in acc.items() if k not in ("lrcheck", "median3", "speckle", "prefilter")}, flush=True)
    m.close()
