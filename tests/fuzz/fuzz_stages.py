"""Randomised differential test of the stages around the matcher against the C oracle: remap (random CV_16SC2 maps incl. far
out-of-range), BGR2GRAY, half-size INTER_AREA (vs cv2 when importable), median 3x3, filterSpeckles, reprojectImageTo3D (random Q,
both handleMissingValues), VoxelGrid (random clouds: NaN / inf, duplicates, negative coordinates, points on voxel faces, leaf sizes
down to the int32-guard passthrough), NV12 / YUYV ingest. usage: python tests/fuzz/fuzz_stages.py [cases] [seed]. Test infrastructure."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import orc
from stereo_depth_ruler_b200 import capi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ctx = capi.Context(0)
try:
    import cv2
except Exception:
    cv2 = None
fails = {}
runs = {}


def check(name, ok, info=""):
    runs[name] = runs.get(name, 0) + 1
    if not ok:
        fails[name] = fails.get(name, 0) + 1
        print("MISMATCH %s %s" % (name, info))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32 if a.dtype == np.float32 else a.dtype)


t0 = time.time()
for case in range(N):
    H, W = int(rng.integers(1, 90)), int(rng.integers(2, 300))
    # ---- remap with random fixed-point maps ----
    sh, sw = int(rng.integers(1, 80)), int(rng.integers(1, 200))
    ch = int(rng.choice([1, 3]))
    src = rng.integers(0, 256, size=(sh, sw) if ch == 1 else (sh, sw, 3)).astype(np.uint8)
    spread = int(rng.choice([0, 2, 40, 3000]))
    m1 = np.stack([rng.integers(-spread - 2, sw + spread + 2, size=(H, W)), rng.integers(-spread - 2, sh + spread + 2, size=(H, W))], -1).astype(np.int16)
    m2 = rng.integers(0, 1024, size=(H, W)).astype(np.uint16)
    got = capi.remap_with_maps(ctx, src, m1, m2)
    check("remap", np.array_equal(got, orc.remap(src, m1, m2)), "case %d src %s dst %dx%d spread %d" % (case, src.shape, W, H, spread))
    # ---- gray, half-size resize ----
    bgr = rng.integers(0, 256, size=(H, W, 3)).astype(np.uint8)
    check("bgr2gray", np.array_equal(capi.bgr2gray(ctx, bgr), orc.bgr2gray(bgr)), "case %d" % case)
    if cv2 is not None and H >= 2:
        He, We = H - H % 2, W - W % 2
        g = bgr[:He, :We, 0].copy()
        check("resize_half", np.array_equal(capi.resize_half_area(ctx, g), cv2.resize(g, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)),
              "case %d %dx%d" % (case, We, He))
    # ---- median + speckle on a disparity-like image ----
    kind = int(rng.integers(0, 3))
    if kind == 0:
        d16 = rng.integers(-16, 2048, size=(H, W)).astype(np.int16)
    elif kind == 1:
        d16 = (np.cumsum(rng.integers(-3, 4, size=(H, W)), axis=1) + 400).astype(np.int16)
        d16[rng.random((H, W)) < 0.2] = -16
    else:
        d16 = np.full((H, W), int(rng.integers(0, 500)), np.int16)
        d16[rng.random((H, W)) < 0.5] = -16
    check("median3", np.array_equal(capi.median3_16s(ctx, d16), orc.median3(d16)), "case %d kind %d" % (case, kind))
    ms, md = int(rng.choice([0, 1, 10, 200, 100000])), int(rng.choice([0, 1, 2, 16, 32, 1000]))
    check("speckle", np.array_equal(capi.filter_speckles(ctx, d16, -16, ms, md), orc.filter_speckles(d16, -16, ms, md)),
          "case %d kind %d max_size %d max_diff %d %dx%d" % (case, kind, ms, md, W, H))
    # ---- reproject ----
    Q = np.eye(4)
    f, bl = float(rng.uniform(100, 2000)), float(rng.uniform(10, 200))
    Q[0, 3], Q[1, 3], Q[2, 2], Q[2, 3], Q[3, 2], Q[3, 3] = -rng.uniform(0, W), -rng.uniform(0, H), 0.0, f, rng.choice([-1, 1]) / bl, rng.choice([0.0, rng.uniform(-1, 1)])
    if rng.random() < 0.3:
        Q += rng.normal(0, 0.01, size=(4, 4))
    for hm in (False, True):
        got = capi.reproject_to_3d(ctx, d16, Q, hm)
        ref = orc.reproject_to_3d(orc.disp_to_float(d16), Q, hm)
        check("reproject", np.array_equal(bits(got), bits(ref)), "case %d hm %d" % (case, hm))
    # ---- voxel grid on random clouds ----
    npts = H * W
    mode = int(rng.integers(0, 4))
    if mode == 0:
        pts = rng.normal(0, 1.0, size=(H, W, 3)).astype(np.float32)
    elif mode == 1:
        pts = (rng.integers(-40, 40, size=(H, W, 3)) * np.float32(0.05)).astype(np.float32)      # on voxel faces, many duplicates
    elif mode == 2:
        pts = rng.uniform(-5000, 5000, size=(H, W, 3)).astype(np.float32)
    else:
        pts = orc.reproject_to_3d(orc.disp_to_float(np.abs(d16)), Q, True)
    bad = rng.random((H, W)) < 0.1
    pts[bad] = rng.choice([np.nan, np.inf, -np.inf])
    leaf = float(rng.choice([0.05, 0.1, 0.5, 5.0, 0.001, 1e-5]))
    col = bgr if rng.random() < 0.7 else None
    cl = capi.Cloud(ctx)
    got = cl.voxelgrid(pts, col, capi.voxel_options(leaf))
    cl.close()
    with np.errstate(all="ignore"):
        ref = orc.voxelgrid(pts, col, leaf)
    if ref is None:
        check("voxel", got["passthrough"], "case %d expected passthrough leaf %g" % (case, leaf))
    else:
        ok = (not got["passthrough"]) and np.array_equal(got["idx"], ref["idx"]) and np.array_equal(got["cnt"], ref["cnt"]) and \
            np.array_equal(bits(got["xyz"]), bits(ref["xyz"])) and (col is None or np.array_equal(got["rgb"], ref["rgb"]))
        check("voxel", ok, "case %d mode %d leaf %g n %d nvox %d" % (case, mode, leaf, npts, len(ref["idx"])))
    # ---- ingest ----
    He, We = max(2, H - H % 2), max(2, W - W % 2)
    nv12 = rng.integers(0, 256, size=(He * 3 // 2, We)).astype(np.uint8)
    check("nv12", np.array_equal(capi.ingest_to_bgr(ctx, capi.PIX_NV12, nv12), orc.ingest_to_bgr(1, nv12)), "case %d" % case)
    yuyv = rng.integers(0, 256, size=(He, We, 2)).astype(np.uint8)
    check("yuyv", np.array_equal(capi.ingest_to_bgr(ctx, capi.PIX_YUYV, yuyv), orc.ingest_to_bgr(2, yuyv)), "case %d" % case)
    # ---- rectification maps of a random (sometimes extreme) calibration ----
    if case % 3 == 0:
        Wc, Hc = int(rng.integers(40, 300)), int(rng.integers(30, 200))
        c = capi.Calibration()
        c.image_width, c.image_height = Wc, Hc
        sc = float(rng.choice([1.0, 1.0, 3.0]))
        mats = {}
        for side, (Kn, dn, Rn, Pn) in enumerate((("K_left", "dist_left", "R1", "P1"), ("K_right", "dist_right", "R2", "P2"))):
            K = np.array([[rng.uniform(0.5, 1.5) * Wc, 0, Wc / 2 + rng.uniform(-10, 10)], [0, rng.uniform(0.5, 1.5) * Wc, Hc / 2 + rng.uniform(-10, 10)], [0, 0, 1]])
            dist = np.array([rng.normal(0, 0.1) * sc, rng.normal(0, 0.05) * sc, rng.normal(0, 0.001), rng.normal(0, 0.001), rng.normal(0, 0.01) * sc])
            rv = rng.normal(0, 0.02, size=3)
            th = np.linalg.norm(rv); kx = rv / th
            Kx = np.array([[0, -kx[2], kx[1]], [kx[2], 0, -kx[0]], [-kx[1], kx[0], 0]])
            R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
            P = np.hstack([K * rng.uniform(0.9, 1.1), np.array([[rng.uniform(-100, 0)], [0], [0]])]); P[2, 2] = 1.0
            for name, m in ((Kn, K), (dn, dist), (Rn, R), (Pn, P)):
                getattr(c, name)[:] = list(np.asarray(m, np.float64).ravel())
            mats[side] = (K, dist.reshape(1, 5), R, P)
        c.present = 0xFFFFFFFF
        rect = capi.Rectifier(ctx, c)
        for eye in (0, 1):
            g1, g2 = rect.maps(eye)
            o1, o2 = orc.init_rectify_map(*mats[eye], (Wc, Hc))
            check("rectify_map", np.array_equal(g1, o1) and np.array_equal(g2, o2), "case %d eye %d %dx%d sc %g" % (case, eye, Wc, Hc, sc))
        rect.close()
print("fuzz_stages: %d cases, seed %d, %.0f s; runs %s; failures %s" % (N, seed, time.time() - t0, runs, fails or "none"))
sys.exit(1 if fails else 0)
