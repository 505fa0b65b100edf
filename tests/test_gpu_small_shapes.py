"""GPU: the whole chain on a small, scaled calibration (odd sizes, every kernel family once) against the oracle chain.
compute-sanitizer is closed on this GPU pool, so out-of-range accesses are hunted with small ragged shapes instead."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scaled_calibration(calib_path, W, H, s):
    from stereo_depth_ruler_b200 import capi
    c = capi.load_calibration(calib_path)
    c.image_width, c.image_height = W, H
    for name in ("K_left", "K_right"):
        a = getattr(c, name)
        for i in (0, 2, 4, 5):
            a[i] /= s
    for name in ("P1", "P2"):
        a = getattr(c, name)
        for i in (0, 2, 3, 5, 6):
            a[i] /= s
    return c


@pytest.mark.parametrize("W,H,D,mode", [(163, 91, 32, 0), (200, 57, 48, 1), (177, 83, 80, 2), (131, 20, 16, 0)])
def test_small_pipeline_matches_oracle_chain(ctx, orc, calib_path, W, H, D, mode):
    from stereo_depth_ruler_b200 import capi
    calib = _scaled_calibration(calib_path, W, H, 1280.0 / W)
    rng = np.random.default_rng(W + H)
    sbs = [rng.integers(0, 255, size=(H, 2 * W, 3)).astype(np.uint8) for _ in range(3)]
    # make the right half a shifted copy of the left so that some disparities are valid
    for f in sbs:
        f[:, W:] = np.roll(f[:, :W], -7, axis=1)
    args = (0, D, 5, 600, 2400, 1, 63, 12, 30, 2, mode)
    p = capi.Pipeline(ctx, calib, capi.make_params(*args), capi.voxel_options(5.0, drop_invalid=True, z_max=5000.0), handle_missing=True,
                      batch=3, want_points=True)
    disp = [np.empty((H, W), np.int16) for _ in sbs]
    pts = [np.empty((H, W, 3), np.float32) for _ in sbs]
    lrect = [np.empty((H, W, 3), np.uint8) for _ in sbs]
    p.run(sbs, disp=disp, points=pts, left_rect=lrect)
    for i, f in enumerate(sbs):
        rect = []
        for eye, (K, d, R, P) in enumerate((("K_left", "dist_left", "R1", "P1"), ("K_right", "dist_right", "R2", "P2"))):
            m1, m2 = orc.init_rectify_map(calib.mat(K), calib.mat(d), calib.mat(R), calib.mat(P), (W, H))
            rect.append(orc.remap(f[:, eye * W:(eye + 1) * W], m1, m2))
        assert np.array_equal(lrect[i], rect[0])
        d16 = orc.sgbm_compute(orc.bgr2gray(rect[0]), orc.bgr2gray(rect[1]), orc.make_params(*args))
        assert np.array_equal(disp[i], d16)
        pref = orc.reproject_to_3d(orc.disp_to_float(d16), calib.mat("Q"), True)
        assert np.array_equal(pts[i].view(np.uint32), pref.view(np.uint32))
        drop = (d16 < 0) | ~((pref[..., 2] > 0) & (pref[..., 2] <= 5000.0))
        cref = orc.voxelgrid(pref, rect[0], 5.0, drop_mask=drop)
        got = p.cloud_to_host(i)
        if cref is None:
            assert got["passthrough"]
        else:
            assert np.array_equal(got["idx"], cref["idx"]) and np.array_equal(got["cnt"], cref["cnt"])
            assert np.allclose(got["xyz"], cref["xyz"], rtol=1e-4, atol=0) and np.array_equal(got["rgb"], cref["rgb"])
    p.close()
