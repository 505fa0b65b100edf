"""Randomised differential test of the CUDA matcher against the C oracle: random shapes, StereoSGBM tuples, modes, batch sizes and
kernel-selection knobs (results must never depend on the knobs). usage: python tests/fuzz/fuzz_sgbm.py [cases] [seed] [max rows, default 70] [max extra width, default 260]
Prints one line per failure and a summary; exit status 1 on any mismatch. Test infrastructure (uses oracle/)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import orc
from stereo_depth_ruler_b200 import capi, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
HMAX = int(sys.argv[3]) if len(sys.argv) > 3 else 70
WEXTRA = int(sys.argv[4]) if len(sys.argv) > 4 else 260
rng = np.random.default_rng(seed)
ctx = capi.Context(0)
KNOBS = [{}, {"cf_cols": 64}, {"cf_cols": 32}, {"cf_cols": 16}, {"cf_cols": 8}, {"path_hh": 1}, {"path_hh": 0}, {"lr_split": 1}, {"fused_sweep": 0},
         {"cost_fusion": 0}, {"shared_device": 1}, {"h_lanes": 16}, {"h_ws": 0}, {"v3_cols": 16}, {"v1_cm": 1}, {"v1_cm": 2}, {"wave": 1}]
fails = 0
t0 = time.time()
stats = {}
for case in range(N):
    mode = int(rng.integers(0, 3))
    D = 16 * int(rng.choice([1, 2, 3, 4, 5, 6, 8, 8, 8, 10, 12, 16]))
    bs = int(rng.choice([1, 3, 5, 5, 5, 7, 9, 11]))
    minD = int(rng.choice([0, 0, 0, 4, -8, -16, 16]))
    W = int(rng.integers(D + max(minD, 0) + 2 * bs + 3, D + max(minD, 0) + 2 * bs + WEXTRA))
    H = int(rng.integers(1, HMAX))
    cap = int(rng.choice([15, 31, 45, 63, 63, 63]))
    ft = max(cap, 15) | 1
    P1 = int(rng.choice([0, 8 * bs * bs, 8 * 3 * bs * bs, int(rng.integers(1, 400))]))
    P2 = int(rng.choice([0, 32 * bs * bs, 32 * 3 * bs * bs, P1 + int(rng.integers(1, 2000))]))
    P1e, P2e = (P1 if P1 > 0 else 2), 0
    P2e = max(P2 if P2 > 0 else 5, P1e + 1)
    if (2 * ft + 63) * bs * bs + P2e + P1e > 32767:
        continue
    d12 = int(rng.choice([-1, 0, 1, 1, 2, 5]))
    uq = int(rng.choice([0, 0, 5, 10, 12, 15, 30]))
    sw = int(rng.choice([0, 0, 50, 100, 200]))
    sr = int(rng.choice([1, 2, 4])) if sw else 0
    args = (minD, D, bs, P1, P2, d12, cap, uq, sw, sr, mode)
    kind = rng.choice(["tex", "noise", "low"])
    B = int(rng.choice([1, 1, 2, 3, 8]))
    pairs = []
    for b in range(B):
        if kind == "tex":
            pairs.append(synth.stereo_pair(W, H, D, int(rng.integers(0, 1000))))
        else:
            pairs.append(synth.noise_pair(W, H, int(rng.integers(0, 1000)), 256 if kind == "noise" else 8))
    knobs = KNOBS[int(rng.integers(0, len(KNOBS)))]
    try:
        m = capi.Sgbm(ctx, capi.make_params(*args))
        for k, v in knobs.items():
            m.set_option(k, v)
        if B == 1:
            got = [m.compute(*pairs[0])]
        else:
            out = m.compute_batch(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]), np.empty((B, H, W), np.int16))
            got = [out[i] for i in range(B)]
        m.close()
    except Exception as e:   # a refusal must be a clean error, never a wrong map
        print("case %d args %s W %d H %d B %d knobs %s: exception %r" % (case, args, W, H, B, knobs, e))
        fails += 1
        continue
    par = orc.make_params(*args)
    for i, (l, r) in enumerate(pairs):
        ref = orc.sgbm_compute(l, r, par)
        if not np.array_equal(ref, got[i]):
            fails += 1
            print("MISMATCH case %d frame %d args %s W %d H %d B %d kind %s knobs %s: %d px differ" % (case, i, args, W, H, B, kind, knobs,
                                                                                                   int((ref != got[i]).sum())))
            break
    stats[(mode, tuple(knobs.items()))] = stats.get((mode, tuple(knobs.items())), 0) + 1
print("fuzz: %d cases, seed %d, %d failures, %.0f s; (mode, knob) combinations covered: %d" % (N, seed, fails, time.time() - t0, len(stats)))
sys.exit(1 if fails else 0)
