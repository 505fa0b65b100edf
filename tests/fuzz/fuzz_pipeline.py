"""Randomised differential test of the whole path per frame (sdr_pipeline_run: SBS frame -> rectify -> matcher -> reproject ->
VoxelGrid) against the oracle chain, on calibrations scaled down from the ZED 2 one (random sizes), random matcher tuples, batch
sizes, BGR / NV12 / YUYV input and leaf sizes. usage: python tests/fuzz/fuzz_pipeline.py [cases] [seed]. Test infrastructure."""
import copy
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import orc
from stereo_depth_ruler_b200 import capi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ctx = capi.Context(0)
CAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "golden", "zed2_calibration.yaml")
fails = 0
t0 = time.time()


def scaled_calib(W, H):
    c = capi.load_calibration(CAL)
    sx, sy = 1280.0 / W, 720.0 / H
    c.image_width, c.image_height = W, H
    for name in ("K_left", "K_right"):
        a = getattr(c, name)
        a[0] /= sx; a[2] /= sx; a[4] /= sy; a[5] /= sy
    for name in ("P1", "P2"):
        a = getattr(c, name)
        a[0] /= sx; a[2] /= sx; a[3] /= sx; a[5] /= sy; a[6] /= sy
    q = c.Q
    q[3] /= sx; q[7] /= sy; q[11] /= sx      # -cx, -cy, f (Q as cv::stereoRectify lays it out)
    return c


def oracle_chain(c, sbs, args, leaf, zmax, drop):
    W, H = c.image_width, c.image_height
    rect = []
    for eye, (K, d, R, P) in enumerate((("K_left", "dist_left", "R1", "P1"), ("K_right", "dist_right", "R2", "P2"))):
        m1, m2 = orc.init_rectify_map(c.mat(K), c.mat(d), c.mat(R), c.mat(P), (W, H))
        rect.append(orc.remap(sbs[:, eye * W:(eye + 1) * W], m1, m2))
    d16 = orc.sgbm_compute(orc.bgr2gray(rect[0]), orc.bgr2gray(rect[1]), orc.make_params(*args))
    pts = orc.reproject_to_3d(orc.disp_to_float(d16), c.mat("Q"), True)
    dm = np.zeros(d16.shape, bool)
    if drop:
        dm |= d16 < args[0] * 16
    if zmax > 0:
        dm |= ~((pts[..., 2] > 0) & (pts[..., 2] <= zmax))
    with np.errstate(all="ignore"):
        cloud = orc.voxelgrid(pts, rect[0], leaf, drop_mask=dm)
    return rect[0], d16, pts, cloud


def bgr_to_nv12_like(rng, H, W2):
    return rng.integers(0, 256, size=(H * 3 // 2, W2)).astype(np.uint8)


for case in range(N):
    W = int(rng.integers(40, 200)) * 2
    H = int(rng.integers(10, 60)) * 2
    D = 16 * int(rng.choice([1, 2, 4, 5, 8]))
    if W <= D + 20:
        continue
    mode = int(rng.integers(0, 3))
    bs = int(rng.choice([3, 5, 7]))
    args = (0, D, bs, 8 * 3 * bs * bs, 32 * 3 * bs * bs, int(rng.choice([1, 2, -1])), 63, int(rng.choice([0, 10, 12])), int(rng.choice([0, 100, 200])), 2, mode)
    B = int(rng.choice([1, 2, 3]))
    fmt = int(rng.choice([capi.PIX_BGR, capi.PIX_BGR, capi.PIX_NV12, capi.PIX_YUYV]))
    leaf = float(rng.choice([5.0, 20.0, 1.0]))
    zmax = float(rng.choice([0.0, 5000.0, 20000.0]))
    drop = bool(rng.integers(0, 2))
    c = scaled_calib(W, H)
    # smooth random texture shifted between the eyes so the matcher finds structure
    base = rng.integers(0, 256, size=(H, W + D, 3)).astype(np.uint8)
    base = ((base.astype(np.int32) + np.roll(base, 1, 1) + np.roll(base, 1, 0) + np.roll(base, 2, 1)) // 4).astype(np.uint8)
    frames_bgr, frames_in = [], []
    for b in range(B):
        sh = int(rng.integers(0, D))
        sbs = np.ascontiguousarray(np.concatenate([np.roll(base, b * 3, 0)[:, sh:sh + W], np.roll(base, b * 3, 0)[:, :W]], axis=1))
        if fmt == capi.PIX_BGR:
            frames_in.append(sbs); frames_bgr.append(sbs)
        elif fmt == capi.PIX_NV12:
            raw = rng.integers(0, 256, size=(H * 3 // 2, 2 * W)).astype(np.uint8)
            raw[:H] = sbs[..., 1]
            frames_in.append(raw); frames_bgr.append(orc.ingest_to_bgr(1, raw))
        else:
            raw = rng.integers(96, 160, size=(H, 2 * W, 2)).astype(np.uint8)
            raw[..., 0] = sbs[..., 1]
            frames_in.append(raw); frames_bgr.append(orc.ingest_to_bgr(2, raw))
    tag = "case %d %dx%d D %d mode %d bs %d B %d fmt %d leaf %g zmax %g drop %d args %s" % (case, W, H, D, mode, bs, B, fmt, leaf, zmax, drop, args)
    try:
        p = capi.Pipeline(ctx, c, capi.make_params(*args), capi.voxel_options(leaf, drop_invalid=drop, z_max=zmax), handle_missing=True, batch=B, want_points=True)
        if fmt != capi.PIX_BGR:
            p.set_input_format(fmt)
        disp = [np.empty((H, W), np.int16) for _ in range(B)]
        pts = [np.empty((H, W, 3), np.float32) for _ in range(B)]
        lrect = [np.empty((H, W, 3), np.uint8) for _ in range(B)]
        p.run(frames_in, disp=disp, points=pts, left_rect=lrect)
        clouds = [p.cloud_to_host(i) for i in range(B)]
        p.close()
    except Exception as e:
        print("EXC", tag, str(e)[:200])
        fails += 1
        continue
    for i in range(B):
        rl, d16, pref, cref = oracle_chain(c, frames_bgr[i], args, leaf, zmax, drop)
        what = None
        if not np.array_equal(lrect[i], rl): what = "left_rect"
        elif not np.array_equal(disp[i], d16): what = "disparity (%d px)" % int((disp[i] != d16).sum())
        elif not np.array_equal(pts[i].view(np.uint32), pref.view(np.uint32)): what = "points"
        else:
            g = clouds[i]
            if cref is None:
                if not g["passthrough"]: what = "cloud: expected passthrough"
            elif g["passthrough"] or not (np.array_equal(g["idx"], cref["idx"]) and np.array_equal(g["cnt"], cref["cnt"]) and
                                          np.array_equal(g["xyz"].view(np.uint32), cref["xyz"].view(np.uint32)) and np.array_equal(g["rgb"], cref["rgb"])):
                what = "cloud (%d vs %d voxels)" % (len(g["xyz"]), len(cref["idx"]))
        if what:
            print("MISMATCH", tag, "frame", i, what)
            fails += 1
            break
print("fuzz_pipeline: %d cases, seed %d, %d failures, %.0f s" % (N, seed, fails, time.time() - t0))
sys.exit(1 if fails else 0)
