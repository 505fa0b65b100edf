"""Small end-to-end invocations for compute-sanitizer (memcheck): every kernel family once on tiny shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stereo_depth_ruler_b200 import capi, synth

ctx = capi.Context(0)
calib = capi.load_calibration(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "zed2_calibration.yaml"))
for mode, (W, H, D, bs, minD) in ((0, (150, 40, 32, 5, 0)), (1, (131, 33, 48, 3, -8)), (2, (170, 45, 80, 7, 0)), (0, (90, 19, 16, 11, 4))):
    l, r = synth.stereo_pair(W, H, D, mode)
    m = capi.Sgbm(ctx, capi.make_params(minD, D, bs, 8 * bs * bs, 32 * bs * bs, 1, 63, 12, 50, 2, mode))
    d = m.compute(l, r)
    m.compute_batch(np.stack([l, l]), np.stack([r, r]), np.empty((2, H, W), np.int16))
    pts = capi.reproject_to_3d(ctx, d, calib.mat("Q"), True)
    cl = capi.Cloud(ctx)
    cl.depth_cloud(d, synth.colorize(l), calib.mat("Q"), capi.voxel_options(5.0, drop_invalid=True, z_max=5000.0, min_disparity=minD), True)
    cl.voxelgrid(pts, None, capi.voxel_options(0.005))
    cl.close(); m.close()
    print("mode", mode, "ok", int((d >= minD * 16).sum()))
# rectifier on a small synthetic calibration (scaled intrinsics) + pipeline
c = capi.load_calibration(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "zed2_calibration.yaml"))
s = 8.0
c.image_width, c.image_height = 160, 90
for name in ("K_left", "K_right"):
    a = getattr(c, name)
    for i in (0, 2, 4, 5):
        a[i] /= s
for name in ("P1", "P2"):
    a = getattr(c, name)
    for i in (0, 2, 3, 5, 6):
        a[i] /= s
rect = capi.Rectifier(ctx, c)
rng = np.random.default_rng(0)
sbs = rng.integers(0, 255, size=(90, 320, 3)).astype(np.uint8)
rect.rectify(sbs[:, :160], sbs[:, 160:], want_gray=True)
rect.maps(1)
p = capi.Pipeline(ctx, c, capi.make_params(0, 32, 5, 600, 2400, 1, 63, 12, 50, 2, 0), capi.voxel_options(5.0, True, 5000.0), True, batch=2, want_points=True)
p.run([sbs, sbs], disp=[np.empty((90, 160), np.int16) for _ in range(2)], points=[np.empty((90, 160, 3), np.float32) for _ in range(2)])
p.cloud_to_host(1)
p.close(); rect.close()
x = rng.integers(-16, 500, size=(33, 47)).astype(np.int16)
capi.median3_16s(ctx, x); capi.filter_speckles(ctx, x, -16, 20, 16); capi.bgr2gray(ctx, sbs)
print("sanitize run complete, launches", ctx.kernel_launches)
