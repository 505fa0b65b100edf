"""Per-kernel times of other modes / sizes (device-resident, batched)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stereo_depth_ruler_b200 import capi, synth
ctx = capi.Context(0)
for (W, H, D, mode, B) in ((1280, 720, 128, 0, 8), (1280, 720, 128, 0, 1), (1280, 720, 80, 2, 8), (1280, 720, 128, 2, 8), (1280, 720, 128, 1, 8), (1920, 1080, 256, 1, 2), (1920, 1080, 256, 1, 4), (3840, 2160, 256, 2, 1), (640, 360, 80, 2, 16)):
    l, r = synth.stereo_pair(W, H, D, 1)
    L = torch.from_numpy(np.stack([l] * B)).cuda(); R = torch.from_numpy(np.stack([r] * B)).cuda()
    O = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
    m = capi.Sgbm(ctx, capi.make_params(0, D, 5, 600, 2400, 1, 63, 12, 200, 2, mode))
    m.set_profiling(True)
    for kv in sys.argv[1:]:            # kernel-selection knobs: name=value (sdr_sgbm_set_option)
        k, v = kv.split("=")
        m.set_option(k, int(v))
    for _ in range(3):
        m.compute_batch(L, R, O); ctx.synchronize()
    st = m.stage_times()
    tot = sum(t for _, t in st)
    print(f"{W}x{H}x{D} mode{mode} B={B}: total {tot:.3f} ms = {tot/B:.3f} ms/frame ({B/tot*1e3:.0f} fps, {W*H*D*B/tot/1e6:.1f} Gpix.disp/s)", {k: round(v, 3) for k, v in st})
    m.close(); del L, R, O
