"""CPU: the C oracle against cv2 itself (when importable) on a wider parameter sweep than the fixtures."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from stereo_depth_ruler_b200 import synth  # noqa: E402

CASES = []
for mode in (0, 1, 2):
    CASES += [
        (96, 48, (0, 16, 1, 8, 32, 1, 0, 10, 0, 0, mode), "tex"),
        (130, 40, (0, 32, 11, 2904, 11616, 1, 31, 1, 20, 1, mode), "tex"),
        (100, 33, (0, 64, 5, 600, 2400, 2, 63, 30, 200, 2, mode), "noise"),
        (90, 17, (4, 16, 3, 0, 0, 0, 15, -1, 10, 4, mode), "low"),
        (67, 20, (0, 64, 5, 600, 2400, 1, 63, 12, 0, 0, mode), "tex"),   # narrowest width cv2 accepts: W1 = 3
    ]


@pytest.mark.parametrize("W,H,args,kind", CASES)
def test_sgbm_sweep(orc, W, H, args, kind):
    cv2.setNumThreads(1)
    if kind == "tex":
        l, r = synth.stereo_pair(W, H, args[1], 3)
    else:
        l, r = synth.noise_pair(W, H, 5, 256 if kind == "noise" else 8)
    ref = cv2.StereoSGBM_create(*args).compute(l, r)
    got = orc.sgbm_compute(l, r, orc.make_params(*args))
    assert np.array_equal(ref, got)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_tied_minimum_without_uniqueness(orc, mode):
    """uniquenessRatio 0 (the matcher createRightMatcher builds, stereo_disparity.cpp:10) exposes how cv2 resolves a tied
    minimum of S: first minimum in MODE_SGBM / MODE_HH, the 8-lane rule of computeDisparity3WaySGBM in MODE_SGBM_3WAY
    (oracle/sdr_oracle.c wta_row). The half-size seed-2 frame has ties 8 and 16 disparities apart (rows 240..300)."""
    l, r = synth.stereo_pair(1280, 720, 128, 2)
    ls = cv2.resize(l, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)[240:300]
    rs = cv2.resize(r, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)[240:300]
    for a, b, minD, D in ((rs, ls, -79, 80), (ls, rs, 0, 64)):
        args = (minD, D, 5, 600, 2400, 1000000, 63, 0, 0, 0, mode)
        assert np.array_equal(cv2.StereoSGBM_create(*args).compute(a, b), orc.sgbm_compute(a, b, orc.make_params(*args)))


def test_reproject_bits(orc, golden):
    rng = np.random.default_rng(3)
    d = rng.integers(-16, 1600, size=(30, 40)).astype(np.int16)
    d[0, 0] = 0
    df = d.astype(np.float32) / 16
    for hm in (False, True):
        a = cv2.reprojectImageTo3D(df, golden["Q"], handleMissingValues=hm)
        b = orc.reproject_to_3d(df, golden["Q"], hm)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_speckles_and_median(orc):
    rng = np.random.default_rng(4)
    for size, diff in ((200, 32), (30, 16), (5, 0)):
        y = (rng.integers(0, 30, size=(70, 90)) * 8).astype(np.int16)
        y[rng.random(y.shape) < 0.25] = -16
        z = y.copy()
        cv2.filterSpeckles(z, -16, size, diff)
        assert np.array_equal(z, orc.filter_speckles(y, -16, size, diff))
    x = rng.integers(-300, 300, size=(9, 5)).astype(np.int16)
    assert np.array_equal(cv2.medianBlur(x, 3), orc.median3(x))


@pytest.mark.parametrize("mode,W", [(0, 66), (1, 66), (2, 64)])
def test_too_small_images_are_refused_like_cv2(orc, mode, W):
    args = (0, 64, 5, 600, 2400, 1, 63, 12, 0, 0, mode)
    l, r = synth.noise_pair(W, 20, 1)
    with pytest.raises(cv2.error):
        cv2.StereoSGBM_create(*args).compute(l, r)
    with pytest.raises(ValueError):
        orc.sgbm_compute(l, r, orc.make_params(*args))
