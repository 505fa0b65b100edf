"""Randomised pin of the ORACLE itself against cv2 (CPU only, no GPU): the same random matcher tuples and stage inputs the GPU
fuzzers use, oracle/sdr_oracle.c against cv2.StereoSGBM / remap / cvtColor / medianBlur / filterSpeckles / reprojectImageTo3D /
initUndistortRectifyMap / resize / cvtColor(YUV). usage: python tests/fuzz/fuzz_oracle_vs_cv2.py [cases] [seed]."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cv2
import numpy as np
from oracle import orc
from stereo_depth_ruler_b200 import synth

cv2.setNumThreads(0)   # MODE_SGBM_3WAY runs its four stripes as one when OpenCV has a single thread (parallel_for_ hands the body the whole range)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
fails, runs = {}, {}


def check(name, ok, info=""):
    runs[name] = runs.get(name, 0) + 1
    if not ok:
        fails[name] = fails.get(name, 0) + 1
        print("MISMATCH %s %s" % (name, info))


t0 = time.time()
for case in range(N):
    # ---- matcher ----
    mode = int(rng.integers(0, 3))
    D = 16 * int(rng.choice([1, 2, 3, 4, 5, 6, 8, 8, 10, 12, 16]))
    bs = int(rng.choice([1, 3, 5, 5, 7, 9, 11]))
    minD = int(rng.choice([0, 0, 0, 4, -8, -16, 16]))
    W = int(rng.integers(D + max(minD, 0) + 2 * bs + 3, D + max(minD, 0) + 2 * bs + 200))
    H = int(rng.integers(1, 60))
    if mode == 2 and rng.random() < 0.3:
        H = int(rng.integers(1, 4 * (bs // 2 + 1) + 2))      # tiny frames: the shifted stripes of computeDisparity3WaySGBM
    cap = int(rng.choice([15, 31, 45, 63, 63]))
    ft = max(cap, 15) | 1
    P1 = int(rng.choice([0, 8 * bs * bs, 8 * 3 * bs * bs, int(rng.integers(1, 400))]))
    P2 = int(rng.choice([0, 32 * bs * bs, 32 * 3 * bs * bs, P1 + int(rng.integers(1, 2000))]))
    P1e = P1 if P1 > 0 else 2
    P2e = max(P2 if P2 > 0 else 5, P1e + 1)
    if (2 * ft + 63) * bs * bs + P2e + P1e <= 32767:
        args = (minD, D, bs, P1, P2, int(rng.choice([-1, 0, 1, 2, 5])), cap, int(rng.choice([0, 0, 5, 10, 15, 30])),
                int(rng.choice([0, 50, 200])), int(rng.choice([1, 2, 4])), mode)
        kind = rng.choice(["tex", "noise", "low"])
        l, r = synth.stereo_pair(W, H, D, int(rng.integers(0, 1000))) if kind == "tex" else synth.noise_pair(W, H, int(rng.integers(0, 1000)), 256 if kind == "noise" else 8)
        try:
            ref = cv2.StereoSGBM_create(*args[:10], mode=args[10]).compute(l, r)
            got = orc.sgbm_compute(l, r, orc.make_params(*args))
            check("sgbm", np.array_equal(ref, got), "case %d %dx%d %s %s: %d px" % (case, W, H, args, kind, int((ref != got).sum())))
        except cv2.error:
            runs["sgbm_cv2_refused"] = runs.get("sgbm_cv2_refused", 0) + 1
    # ---- stages ----
    H2, W2 = int(rng.integers(1, 80)), int(rng.integers(2, 260))
    sh, sw = int(rng.integers(1, 70)), int(rng.integers(1, 180))
    ch = int(rng.choice([1, 3]))
    src = rng.integers(0, 256, size=(sh, sw) if ch == 1 else (sh, sw, 3)).astype(np.uint8)
    spread = int(rng.choice([0, 2, 40, 3000]))
    m1 = np.stack([rng.integers(-spread - 2, sw + spread + 2, size=(H2, W2)), rng.integers(-spread - 2, sh + spread + 2, size=(H2, W2))], -1).astype(np.int16)
    m2 = rng.integers(0, 1024, size=(H2, W2)).astype(np.uint16)
    check("remap", np.array_equal(orc.remap(src, m1, m2), cv2.remap(src, m1, m2, cv2.INTER_LINEAR)), "case %d" % case)
    bgr = rng.integers(0, 256, size=(H2, W2, 3)).astype(np.uint8)
    check("bgr2gray", np.array_equal(orc.bgr2gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)), "case %d" % case)
    d16 = rng.integers(-16, 2048, size=(H2, W2)).astype(np.int16) if rng.random() < 0.5 else (np.cumsum(rng.integers(-3, 4, size=(H2, W2)), axis=1) + 400).astype(np.int16)
    d16[rng.random((H2, W2)) < 0.15] = -16
    check("median3", np.array_equal(orc.median3(d16), cv2.medianBlur(d16, 3)), "case %d %dx%d" % (case, W2, H2))
    ms, md = int(rng.choice([0, 1, 10, 200, 100000])), int(rng.choice([0, 1, 2, 16, 32, 1000]))
    ref = d16.copy()
    cv2.filterSpeckles(ref, -16, ms, md)
    check("speckle", np.array_equal(orc.filter_speckles(d16, -16, ms, md), ref), "case %d max_size %d max_diff %d" % (case, ms, md))
    Q = np.eye(4)
    Q[0, 3], Q[1, 3], Q[2, 2], Q[2, 3], Q[3, 2], Q[3, 3] = -rng.uniform(0, W2), -rng.uniform(0, H2), 0.0, rng.uniform(100, 2000), rng.choice([-1, 1]) / rng.uniform(10, 200), rng.choice([0.0, rng.uniform(-1, 1)])
    if rng.random() < 0.3:
        Q += rng.normal(0, 0.01, size=(4, 4))
    f32 = d16.astype(np.float32) / 16
    check("disp_to_float", np.array_equal(orc.disp_to_float(d16).view(np.uint32), f32.view(np.uint32)), "case %d" % case)
    for hm in (False, True):
        a = orc.reproject_to_3d(f32, Q, hm)
        b = cv2.reprojectImageTo3D(f32, Q, handleMissingValues=hm)
        check("reproject", np.array_equal(a.view(np.uint32), b.view(np.uint32)), "case %d hm %d" % (case, hm))
    He, We = max(2, H2 - H2 % 2), max(2, W2 - W2 % 2)
    nv12 = rng.integers(0, 256, size=(He * 3 // 2, We)).astype(np.uint8)
    check("nv12", np.array_equal(orc.ingest_to_bgr(1, nv12), cv2.cvtColor(nv12, cv2.COLOR_YUV2BGR_NV12)), "case %d" % case)
    yuyv = rng.integers(0, 256, size=(He, We, 2)).astype(np.uint8)
    check("yuyv", np.array_equal(orc.ingest_to_bgr(2, yuyv), cv2.cvtColor(yuyv, cv2.COLOR_YUV2BGR_YUY2)), "case %d" % case)
    # ---- rectification maps on a perturbed calibration ----
    if case % 5 == 0:
        Wc, Hc = int(rng.integers(40, 400)), int(rng.integers(30, 300))
        K = np.array([[rng.uniform(0.5, 1.5) * Wc, 0, Wc / 2 + rng.uniform(-10, 10)], [0, rng.uniform(0.5, 1.5) * Wc, Hc / 2 + rng.uniform(-10, 10)], [0, 0, 1]])
        dist = np.array([[rng.normal(0, 0.1), rng.normal(0, 0.05), rng.normal(0, 0.001), rng.normal(0, 0.001), rng.normal(0, 0.01)]])
        rv = rng.normal(0, 0.02, size=3)
        R, _ = cv2.Rodrigues(rv)
        P = np.hstack([K * rng.uniform(0.9, 1.1), np.array([[rng.uniform(-100, 0)], [0], [0]])])
        P[2, 2] = 1.0
        a1, a2 = orc.init_rectify_map(K, dist, R, P, (Wc, Hc))
        b1, b2 = cv2.initUndistortRectifyMap(K, dist, R, P, (Wc, Hc), cv2.CV_16SC2)
        # beyond the int16 range OpenCV is not consistent with itself (its SIMD body saturates map1, the scalar tail of a row
        # wraps): pinned for source coordinates inside (-32768, 32767) only; the restatement and the GPU saturate
        if np.abs(a1.astype(np.int32)).max() < 32767:
            check("rectify_map", np.array_equal(a1, b1) and np.array_equal(a2, b2), "case %d %dx%d: %d / %d differ" % (case, Wc, Hc, int((a1 != b1).sum()), int((a2 != b2).sum())))
print("fuzz_oracle_vs_cv2: %d cases, seed %d, %.0f s; runs %s; failures %s" % (N, seed, time.time() - t0, runs, fails or "none"))
sys.exit(1 if fails else 0)
