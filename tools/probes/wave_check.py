"""Wavefront sweeps (option wave=1 / 2) against the default path: small ragged shapes, then the headline batch with timings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from stereo_depth_ruler_b200 import capi, synth
ctx = capi.Context(0)
modes = [int(a) for a in sys.argv[1:]] or [2, 1]
def run(W, H, args, wave, batch=1, seed=5):
    pairs = [synth.stereo_pair(W, H, args[1], seed + i) for i in range(min(batch, 2))]
    L = np.stack([pairs[i % len(pairs)][0] for i in range(batch)]); R = np.stack([pairs[i % len(pairs)][1] for i in range(batch)])
    m = capi.Sgbm(ctx, capi.make_params(*args))
    m.set_option("wave", wave)
    out = m.compute_batch(L, R, np.empty((batch, H, W), np.int16))
    raw = m.debug_read(2, (H, W)) if batch == 1 else None
    Sv = m.debug_read(1, (H, W - args[1] - max(args[0], 0) + 0, args[1])) if batch == 1 and args[0] == 0 else None
    m.close()
    return out, raw, Sv
shapes = [(200, 24, (0, 64, 5, 600, 2400, 1, 63, 12, 0, 0, 0)), (161, 9, (0, 64, 3, 72, 288, 1, 63, 12, 0, 0, 0)),
          (300, 40, (0, 128, 5, 600, 2400, 1, 63, 12, 50, 2, 0)), (257, 31, (0, 128, 3, 72, 288, 2, 63, 5, 0, 0, 0)),
          (224, 17, (0, 128, 5, 600, 2400, 1, 63, 12, 0, 0, 0)), (400, 50, (-16, 64, 5, 600, 2400, 1, 63, 12, 0, 0, 0)),
          (1280, 720, (0, 128, 5, 600, 2400, 1, 63, 12, 200, 2, 0))]
for W, H, args in shapes:
    ref, rraw, rS = run(W, H, args, 0)
    for wave in modes:
        out, raw, Sv = run(W, H, args, wave)
        bad = np.argwhere(raw != rraw)
        msg = "OK" if len(bad) == 0 and np.array_equal(out, ref) else "DIFF raw %d px, first %s" % (len(bad), bad[:4].tolist())
        if wave == 1 and rS is not None:
            msg += " | S equal" if np.array_equal(Sv, rS) else " | S differs at %s" % np.argwhere(Sv != rS)[:3].tolist()
        print(W, H, args[:3], "wave=%d" % wave, msg, flush=True)
# headline batch
W, H, args = 1280, 720, (0, 128, 5, 600, 2400, 1, 63, 12, 200, 2, 0)
B = 8
pairs = [synth.stereo_pair(W, H, 128, s) for s in range(2)]
L = torch.from_numpy(np.stack([pairs[i % 2][0] for i in range(B)])).cuda(); R = torch.from_numpy(np.stack([pairs[i % 2][1] for i in range(B)])).cuda()
O = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
ref = None
for wave in [0] + modes:
    m = capi.Sgbm(ctx, capi.make_params(*args)); m.set_option("shared_device", 1); m.set_option("wave", wave); m.set_profiling(True)
    acc = {}
    for rep in range(6):
        m.compute_batch(L, R, O); ctx.synchronize()
        if rep >= 2:
            for k, t in m.stage_times(): acc[k] = acc.get(k, 0) + t / 4
    o = O.cpu().numpy()
    same = "ref" if ref is None else ("== ref" if np.array_equal(o, ref) else "DIFFERS")
    ref = o if ref is None else ref
    print("batch 8 wave=%d total %.3f ms (%s)" % (wave, sum(acc.values()), same), {k: round(v, 3) for k, v in acc.items()}, flush=True)
    m.close()
