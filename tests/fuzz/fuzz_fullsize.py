"""Full-size randomised test of the matcher against cv2.StereoSGBM itself: the launch shapes only big frames reach (all SMs in
one cooperative launch, many strips, several bands and waves), random tuples / modes / batch sizes at 640x360 ... 1920x1080.
usage: python tests/fuzz/fuzz_fullsize.py [cases] [seed]. Test infrastructure (needs cv2)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cv2
import numpy as np
import torch
from stereo_depth_ruler_b200 import capi, synth

cv2.setNumThreads(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ctx = capi.Context(0)
fails = 0
t0 = time.time()
for case in range(N):
    W, H = [(640, 360), (1280, 720), (1280, 720), (960, 540), (1920, 1080), (1000, 700), (1281, 721)][int(rng.integers(0, 7))]
    mode = int(rng.integers(0, 3))
    D = 16 * int(rng.choice([2, 4, 5, 6, 8, 8, 8, 12, 16]))
    bs = int(rng.choice([3, 5, 5, 7, 11]))
    minD = int(rng.choice([0, 0, 0, -16, 8]))
    ft = 63
    P1, P2 = 8 * bs * bs, 32 * bs * bs
    if rng.random() < 0.5 and (2 * ft + 63) * bs * bs + 3 * P2 + 3 * P1 <= 32767:
        P1, P2 = 3 * P1, 3 * P2
    if (2 * ft + 63) * bs * bs + P2 + P1 > 32767:
        continue
    args = (minD, D, bs, P1, P2, int(rng.choice([1, 1, 2, -1])), 63, int(rng.choice([0, 10, 12, 15])), int(rng.choice([0, 100, 200])), 2, mode)
    B = int(rng.choice([1, 1, 2, 4, 8])) if W * H <= 1280 * 720 else int(rng.choice([1, 2]))
    pairs = [synth.stereo_pair(W, H, min(D, 160), int(rng.integers(0, 1000))) for _ in range(B)]
    knobs = [{}, {}, {"shared_device": 1}, {"lr_split": 1}, {"cf_fpl": 3}, {"path_hh": 1}][int(rng.integers(0, 6))]
    tag = "case %d %dx%d B %d args %s knobs %s" % (case, W, H, B, args, knobs)
    try:
        m = capi.Sgbm(ctx, capi.make_params(*args))
        for k, v in knobs.items():
            m.set_option(k, v)
        L = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda()
        R = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
        O = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
        torch.cuda.synchronize()
        m.compute_batch(L, R, O)
        ctx.synchronize()
        out = O.cpu().numpy()
        m.close()
    except Exception as e:
        print("EXC", tag, str(e)[:200])
        fails += 1
        continue
    ref = cv2.StereoSGBM_create(*args[:10], mode=args[10])
    for i, (l, r) in enumerate(pairs):
        want = ref.compute(l, r)
        if not np.array_equal(want, out[i]):
            print("MISMATCH", tag, "frame", i, int((want != out[i]).sum()), "px")
            fails += 1
            break
print("fuzz_fullsize: %d cases, seed %d, %d failures, %.0f s" % (N, seed, fails, time.time() - t0))
sys.exit(1 if fails else 0)
