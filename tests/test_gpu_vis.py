"""GPU: the reference's live path around the matcher (stereo_disparity.cpp): INTER_AREA half resize, the right matcher of
cv::ximgproc::createRightMatcher, show_disparityMap and show_depthMap -- each against cv2 4.13.0 itself, bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")

REF_TUPLE = (0, 80, 5, 8 * 5 * 5 * 3, 32 * 5 * 5 * 3, 1, 63, 12, 200, 2, 2)      # stereo_disparity.cpp:5-6 (MODE_SGBM_3WAY)


def test_resize_half_area_equals_cv2(ctx):
    from stereo_depth_ruler_b200 import capi, synth
    l, _ = synth.stereo_pair(1280, 720, 128, 2)
    assert np.array_equal(capi.resize_half_area(ctx, l), cv2.resize(l, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA))
    bgr = synth.colorize(l[:200, :300], 3)
    assert np.array_equal(capi.resize_half_area(ctx, bgr), cv2.resize(bgr, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA))
    rng = np.random.default_rng(0)
    noise = rng.integers(0, 256, (66, 130), dtype=np.uint8)          # every rounding case of (a+b+c+d+2)>>2
    assert np.array_equal(capi.resize_half_area(ctx, noise), cv2.resize(noise, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA))
    with pytest.raises(capi.SdrError):                                 # odd sizes leave OpenCV's exact-factor path: refused
        capi.resize_half_area(ctx, noise[:65])


def test_live_path_left_and_right_matcher_equal_cv2(ctx):
    """StereoDisparity::computeDisparity up to (not including) the WLS filter: BGR2GRAY, half-size INTER_AREA, the matcher on
    (left, right) and createRightMatcher's matcher on (right, left), 640x360x80 MODE_SGBM_3WAY."""
    from stereo_depth_ruler_b200 import capi, synth
    l, r = synth.stereo_pair(1280, 720, 160, 5)      # disparities up to ~140 px -> ~70 px at half size
    lb, rb = synth.colorize(l, 1), synth.colorize(r, 2)
    lg, rg = capi.bgr2gray(ctx, lb), capi.bgr2gray(ctx, rb)
    assert np.array_equal(lg, cv2.cvtColor(lb, cv2.COLOR_BGR2GRAY))
    ls, rs = capi.resize_half_area(ctx, lg), capi.resize_half_area(ctx, rg)
    lp = capi.make_params(*REF_TUPLE)
    rp = capi.right_matcher_params(lp)
    assert (rp.minDisparity, rp.numDisparities, rp.blockSize, rp.P1, rp.P2, rp.disp12MaxDiff, rp.preFilterCap, rp.uniquenessRatio,
            rp.speckleWindowSize, rp.speckleRange, rp.mode) == (-79, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, 2)
    ml, mr = capi.Sgbm(ctx, lp), capi.Sgbm(ctx, rp)
    dl, dr = ml.compute(ls, rs), mr.compute(rs, ls)
    cl = cv2.StereoSGBM_create(*REF_TUPLE[:10], mode=cv2.STEREO_SGBM_MODE_SGBM_3WAY)
    cr = cv2.StereoSGBM_create(-79, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, cv2.STEREO_SGBM_MODE_SGBM_3WAY)
    assert np.array_equal(dl, cl.compute(ls, rs))
    ref_r = cr.compute(rs, ls)
    assert np.array_equal(dr, ref_r)
    assert (ref_r[ref_r != -80 * 16] <= 0).all() and (ref_r != -80 * 16).mean() > 0.5      # right-view disparities are negative
    # the other modes of the right matcher as well (createRightMatcher copies the mode)
    for mode, cvm in ((0, cv2.STEREO_SGBM_MODE_SGBM), (1, cv2.STEREO_SGBM_MODE_HH)):
        p = capi.right_matcher_params(capi.make_params(*(REF_TUPLE[:10] + (mode,))))
        m = capi.Sgbm(ctx, p)
        ref = cv2.StereoSGBM_create(-79, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, cvm).compute(rs[:120], ls[:120])
        assert np.array_equal(m.compute(rs[:120], ls[:120]), ref), mode
        m.close()
    ml.close(); mr.close()


def _cv_convert_u8(x, a, b):
    """Mat::convertTo(CV_8U, a, b) on float32 through cv2 (convertScaleAbs is the same kernel plus an abs)."""
    v = x.astype(np.float64) * float(np.float32(a)) + float(np.float32(b))
    out = cv2.convertScaleAbs(x, alpha=a, beta=b)
    out[v < 0] = 0
    out[~np.isfinite(v) | (np.abs(v) >= 2.0 ** 31)] = 0       # cvRound's INT_MIN -> saturate_cast -> 0
    return out


def _ref_show_disparity(disp, num_disp, prev):
    """stereo_disparity.cpp:42-73 with cv2."""
    masked = disp.copy()
    masked[~(disp > 0)] = 0
    norm01 = masked * np.float32(np.float32(1.0) / np.float32(max(1, num_disp)))
    show = _cv_convert_u8(cv2.pow(norm01, 0.6), 255.0, 0.0)
    if prev is not None and prev.shape == show.shape:
        al = np.float32(0.63)
        show = cv2.addWeighted(prev, float(al), show, float(np.float32(1) - al), 0)
    return show


def test_show_disparity_map_every_cv16s_value(ctx):
    """Every disparity a CV_16S / 16 map can hold (plus NaN / inf / negatives) over three frames: the temporal blend state
    lives in the handle like prev_vis in the reference's object."""
    from stereo_depth_ruler_b200 import capi
    d16 = np.arange(-64, 32768 - 64, dtype=np.int32).reshape(128, 256)
    frames = [(d16 / 16.0).astype(np.float32), (d16[::-1] / 16.0).astype(np.float32), (np.roll(d16, 77, 1) / 16.0).astype(np.float32)]
    frames[1][0, :4] = [np.nan, np.inf, -np.inf, -0.0]
    vis = capi.Vis(ctx)
    prev = None
    for i, f in enumerate(frames):
        got = vis.disparity(f, 80)
        prev = _ref_show_disparity(f, 80, prev)
        assert np.array_equal(got, prev), "frame %d: %d pixels differ" % (i, int((got != prev).sum()))
    other = vis.disparity(frames[0][:64], 80)                        # size change: prev_vis is not blended in
    assert np.array_equal(other, _ref_show_disparity(frames[0][:64], 80, None))
    vis.close()


def test_show_depth_map_equals_cv2(ctx):
    """stereo_disparity.cpp:83-124 on a reprojected depth image (valid range EMA, convertTo, COLORMAP_TURBO, blend), three
    frames, 3-channel points and 1-channel depth, incl. inf / NaN / the 10000 fill value."""
    from stereo_depth_ruler_b200 import capi, synth
    calib = capi.load_calibration(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "zed2_calibration.yaml"))
    Q = calib.mat("Q")
    vis = capi.Vis(ctx)
    zs, ze = 1000.0, 2000.0
    prev = None
    for i in range(3):
        l, r = synth.stereo_pair(640, 200, 64, 20 + i)
        d16 = cv2.StereoSGBM_create(0, 64, 5, 600, 2400, 1, 63, 12, 100, 2).compute(l, r)
        if i == 1:
            d16[:5] = 0                                                  # d = 0 -> inf
        pts = cv2.reprojectImageTo3D(d16.astype(np.float32) / 16, Q, handleMissingValues=(i == 2))
        if i == 2:
            pts[10, :7, 2] = np.nan
        src = pts if i != 1 else np.ascontiguousarray(pts[..., 2])
        got = vis.depth(src)
        z = pts[..., 2]
        valid = ((z > 0) & (z < 10000) & (z == z)).astype(np.uint8)
        zmin, zmax, _, _ = cv2.minMaxLoc(z, valid)
        if not zmax > zmin:
            zmin, zmax = 1000.0, 2000.0
        zs, ze = 0.9 * zs + 0.1 * zmin, 0.9 * ze + 0.1 * zmax
        zs = max(0.0, min(zs, 10000.0)); ze = max(zs + 1.0, min(ze, 10000.0))
        assert vis.depth_range() == (zs, ze)
        g = _cv_convert_u8(np.ascontiguousarray(z), 255.0 / (ze - zs), -255.0 * zs / (ze - zs))
        col = cv2.applyColorMap(g, cv2.COLORMAP_TURBO)
        if prev is not None:
            al = np.float32(0.63)
            col = cv2.addWeighted(prev, float(al), col, float(np.float32(1) - al), 0)
        prev = col
        assert np.array_equal(got, col), "frame %d: %d values differ" % (i, int((got != col).sum()))
    vis.close()
