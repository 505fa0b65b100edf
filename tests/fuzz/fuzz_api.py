"""Randomised test of the C ABI's buffer handling and error behaviour: (a) pitched images (rows padded by a random number of bytes,
ROI views of larger arrays, as cv::Mat ROIs of an SBS frame are) in host and device memory, for the matcher, the post-filters, the
reprojection and the rectifier, against the same call on packed copies; (b) invalid arguments (bad tuples, mismatched sizes and
types, too-small images) must come back as an error status -- never a crash or a wrong map -- and leave the handle and the context
usable. usage: python tests/fuzz/fuzz_api.py [cases] [seed]. Test infrastructure."""
import ctypes as C
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
fails = 0
t0 = time.time()


def fail(msg):
    global fails
    fails += 1
    print("FAIL", msg)


def pitched(a, dev):
    """A view with padded rows (and a random offset inside a larger buffer) holding a's data; numpy or torch cuda."""
    H, W = a.shape[:2]
    rest = a.shape[2:]
    pw, ph, ox, oy = int(rng.integers(0, 40)), int(rng.integers(0, 3)), int(rng.integers(0, 9)), int(rng.integers(0, 3))
    big = np.zeros((H + ph + oy, W + pw + ox) + rest, a.dtype)
    big[oy:oy + H, ox:ox + W] = a
    if dev:
        t = torch.from_numpy(big).cuda()
        return t[oy:oy + H, ox:ox + W]
    return big[oy:oy + H, ox:ox + W]


def tonp(x):
    return x.cpu().numpy() if capi._is_torch(x) else np.ascontiguousarray(x)


for case in range(N):
    mode = int(rng.integers(0, 3))
    D = 16 * int(rng.choice([1, 2, 4, 8]))
    bs = int(rng.choice([3, 5, 7]))
    W, H = int(rng.integers(D + 30, D + 260)), int(rng.integers(2, 80))
    args = (0, D, bs, 8 * bs * bs, 32 * bs * bs, 1, 63, 12, int(rng.choice([0, 100])), 2, mode)
    l, r = synth.stereo_pair(W, H, D, int(rng.integers(0, 1000)))
    ref = orc.sgbm_compute(l, r, orc.make_params(*args))
    tag = "case %d %dx%d %s" % (case, W, H, args)
    dev = bool(rng.integers(0, 2))
    try:
        # ---- (a) pitched buffers ----
        m = capi.Sgbm(ctx, capi.make_params(*args))
        lp, rp = pitched(l, dev), pitched(r, dev)
        op = pitched(np.full((H, W), 77, np.int16), dev)
        if dev:
            torch.cuda.synchronize()
        m.compute(lp, rp, op)
        ctx.synchronize()
        if not np.array_equal(tonp(op), ref):
            fail(tag + " pitched compute (%s)" % ("device" if dev else "host"))
        med = capi.median3_16s(ctx, np.ascontiguousarray(ref))
        sp = pitched(ref, False).copy() if False else pitched(ref, False)
        # post-filters on a pitched host view (in place for the speckle filter)
        view = pitched(ref, False)
        ii, _ = capi.as_image(view, 1)
        ctx.check(capi.lib().sdr_filter_speckles(ctx._h, C.byref(ii), -16, 50, 2))
        if not np.array_equal(np.ascontiguousarray(view), orc.filter_speckles(ref, -16, 50, 2)):
            fail(tag + " pitched filter_speckles")
        if not np.array_equal(med, orc.median3(ref)):
            fail(tag + " median3")
        Q = np.eye(4); Q[2, 3] = 700.0; Q[3, 2] = 1 / 120.0; Q[0, 3] = -W / 2; Q[1, 3] = -H / 2; Q[2, 2] = 0
        pv = pitched(ref, dev)
        if dev:
            torch.cuda.synchronize()
        pts = capi.reproject_to_3d(ctx, pv, Q, True)
        ctx.synchronize()
        if not np.array_equal(tonp(pts).view(np.uint32), orc.reproject_to_3d(orc.disp_to_float(ref), Q, True).view(np.uint32)):
            fail(tag + " pitched reproject")
        # ---- (b) invalid arguments: error status, handle still good ----
        bad_calls = 0
        for kind in rng.permutation(8)[:4]:
            try:
                if kind == 0:
                    m.compute(l, r[:, :W - 1].copy())                      # size mismatch
                elif kind == 1:
                    m.compute(l, r, np.empty((H + 1, W), np.int16))        # output of another size
                elif kind == 2:
                    capi.Sgbm(ctx, capi.make_params(0, D + 8, bs, 10, 20, 1, 63, 12, 0, 0, mode))     # D not a multiple of 16
                elif kind == 3:
                    capi.Sgbm(ctx, capi.make_params(0, D, bs + 1, 10, 20, 1, 63, 12, 0, 0, mode))     # even window
                elif kind == 4:
                    m.compute(l[:, :D].copy(), r[:, :D].copy())            # narrower than the disparity range
                elif kind == 5:
                    li, _ = capi.as_image(l.astype(np.int16), 1); ri, _ = capi.as_image(r, 1); oi, _ = capi.as_image(np.empty((H, W), np.int16), 1)
                    ctx.check(capi.lib().sdr_sgbm_compute(m._h, C.byref(li), C.byref(ri), C.byref(oi)))   # wrong element type
                elif kind == 6:
                    capi.Sgbm(ctx, capi.make_params(0, 512, bs, 10, 20, 1, 63, 12, 0, 0, mode))       # D beyond the supported range
                else:
                    capi.reproject_to_3d(ctx, ref, np.eye(4)[:3], True) if False else capi.remap_with_maps(
                        ctx, l, np.zeros((H, W, 2), np.int16), np.zeros((H + 1, W), np.uint16))       # map sizes disagree
                fail(tag + " invalid call kind %d was accepted" % int(kind))
            except capi.SdrError:
                bad_calls += 1
            except AssertionError:
                bad_calls += 1                                             # refused by the Python binding before the ABI
        again = m.compute(l, r)
        if not np.array_equal(again, ref):
            fail(tag + " handle unusable after %d refused calls" % bad_calls)
        m.close()
    except Exception as e:
        fail(tag + " exception %r" % (e,))
print("fuzz_api: %d cases, seed %d, %d failures, %.0f s" % (N, seed, fails, time.time() - t0))
sys.exit(1 if fails else 0)
