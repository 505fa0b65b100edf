"""Randomised differential test of the stripe split of MODE_SGBM_3WAY (BASELINE config 5) against the oracle: random shapes and
tuples; (a) sdr_sgbm_match_stripes over a random partition of the four stripes + one sdr_sgbm_postfilter, host and device-resident
buffers; (b) the native split sdr_sgbm_split_* with a random number of parts (all on device 0 when the box has one GPU, spread over
the GPUs otherwise). usage: python tests/fuzz/fuzz_split.py [cases] [seed]. Test infrastructure."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from oracle import orc
from stereo_depth_ruler_b200 import capi, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ctx = capi.Context(0)
ngpu = torch.cuda.device_count()
fails = 0
t0 = time.time()
for case in range(N):
    D = 16 * int(rng.choice([1, 2, 4, 5, 8, 12, 16]))
    bs = int(rng.choice([3, 5, 5, 7, 9]))
    minD = int(rng.choice([0, 0, -8, 16]))
    W = int(rng.integers(D + max(minD, 0) + 2 * bs + 3, D + max(minD, 0) + 2 * bs + 400))
    H = int(rng.integers(1, 260))
    P1, P2 = 8 * bs * bs, 32 * bs * bs
    args = (minD, D, bs, P1, P2, int(rng.choice([1, 2, -1])), 63, int(rng.choice([0, 10, 12])), int(rng.choice([0, 50, 200])), 2, 2)
    l, r = synth.stereo_pair(W, H, D, int(rng.integers(0, 1000)))
    ref = orc.sgbm_compute(l, r, orc.make_params(*args))
    tag = "case %d %dx%d args %s" % (case, W, H, args)
    try:
        m = capi.Sgbm(ctx, capi.make_params(*args))
        ns = len(capi.stripe_rows(m.params, W, H))
        cuts = sorted(set([0, ns] + [int(x) for x in rng.integers(0, ns + 1, size=int(rng.integers(0, 3)))]))
        dev = bool(rng.integers(0, 2))
        if dev:
            lt, rt = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
            raw = torch.full((H, W), -1234, dtype=torch.int16, device="cuda")
            torch.cuda.synchronize()
        else:
            lt, rt, raw = l, r, np.full((H, W), -1234, np.int16)
        order = list(range(len(cuts) - 1))
        rng.shuffle(order)
        for j in order:
            if cuts[j + 1] > cuts[j]:
                m.match_stripes(lt, rt, cuts[j], cuts[j + 1] - cuts[j], raw)
        out = m.postfilter(raw)
        ctx.synchronize()
        got = out.cpu().numpy() if dev else out
        m.close()
        if not np.array_equal(got, ref):
            print("MISMATCH stripes", tag, "cuts", cuts, "device" if dev else "host", int((got != ref).sum()), "px")
            fails += 1
        nparts = int(rng.integers(1, 7))
        devices = [int(i % ngpu) for i in range(nparts)]
        sp = capi.SgbmSplit(devices, capi.make_params(*args))
        got2 = sp.compute(l, r)
        if dev:
            o3 = sp.compute(lt, rt)
            sp.synchronize()
            if not np.array_equal(o3.cpu().numpy(), ref):
                print("MISMATCH split(device)", tag, "parts", devices)
                fails += 1
        sp.close()
        if not np.array_equal(got2, ref):
            print("MISMATCH split", tag, "parts", devices, int((got2 != ref).sum()), "px")
            fails += 1
    except Exception as e:
        print("EXC", tag, str(e)[:200])
        fails += 1
print("fuzz_split: %d cases, seed %d, %d GPUs, %d failures, %.0f s" % (N, seed, ngpu, fails, time.time() - t0))
sys.exit(1 if fails else 0)
