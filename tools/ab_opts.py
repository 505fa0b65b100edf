"""A/B of kernel knobs on the SGBM core alone (device-resident, one handle): per-stage ms for each setting.
usage: python tools/ab_opts.py W H D mode batch "opt=val opt=val" "..." """
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stereo_depth_ruler_b200 import capi, synth
W, H, D, mode, B = (int(v) for v in sys.argv[1:6])
ctx = capi.Context(0)
pairs = [synth.stereo_pair(W, H, D, s) for s in range(min(B, 2))]
L = torch.from_numpy(np.stack([pairs[i % len(pairs)][0] for i in range(B)])).cuda()
R = torch.from_numpy(np.stack([pairs[i % len(pairs)][1] for i in range(B)])).cuda()
O = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
ref = None
for setting in sys.argv[6:] or [""]:
    m = capi.Sgbm(ctx, capi.make_params(0, D, 5, 600, 2400, 1, 63, 12, 200, 2, mode))
    for kv in setting.split():
        k, v = kv.split("=")
        m.set_option(k, int(v))
    m.set_profiling(True)
    acc = {}
    reps = 5
    for rep in range(reps + 2):
        m.compute_batch(L, R, O); ctx.synchronize()
        if rep < 2:
            continue
        for k, t in m.stage_times():
            acc[k] = acc.get(k, 0.0) + t / reps
    out = O.cpu().numpy()
    same = "ref" if ref is None else ("== ref" if np.array_equal(out, ref) else "DIFFERS")
    ref = out if ref is None else ref
    tot = sum(acc.values())
    print("%-44s total %7.3f ms (%s) | %s" % (setting or "(default)", tot, same, " ".join("%s=%.3f" % (k, v) for k, v in acc.items())), flush=True)
    m.close()
