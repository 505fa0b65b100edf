"""CPU: the C oracle against the golden vectors generated from cv2 4.13.0 (tests/golden/make_golden.py).
This is what pins the oracle: every stage bit-exact."""
import hashlib

import numpy as np
import pytest


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_sgbm_cases(golden, orc):
    for name in golden["sgbm_cases"]:
        args = [int(v) for v in golden[name + "_args"]]
        got = orc.sgbm_compute(golden[name + "_left"], golden[name + "_right"], orc.make_params(*args))
        assert np.array_equal(got, golden[name + "_disp"]), name


def test_small_stages(golden, orc):
    assert np.array_equal(orc.bgr2gray(golden["gray_in"]), golden["gray_out"])
    assert np.array_equal(orc.median3(golden["median_in"]), golden["median_out"])
    assert np.array_equal(orc.filter_speckles(golden["speckle_in"], -16, 200, 32), golden["speckle_out"])
    d16 = golden["reproj_disp16"]
    df = orc.disp_to_float(d16)
    assert np.array_equal(df, d16.astype(np.float32) / 16)
    for hm, key in ((False, "reproj_false"), (True, "reproj_true")):
        got = orc.reproject_to_3d(df, golden["Q"], hm)
        assert np.array_equal(got.view(np.uint32), golden[key]), key


def test_rectify_maps_and_remap(golden, orc, calib_path):
    from stereo_depth_ruler_b200 import capi, synth
    c = capi.load_calibration(calib_path)
    gl, gr = synth.stereo_pair(1280, 720, 128, 7)
    sbs = np.concatenate([synth.colorize(gl, 1), synth.colorize(gr, 2)], axis=1)
    assert sha(sbs) == str(golden["sbs_sha"])
    for eye, (K, d, R, P) in enumerate((("K_left", "dist_left", "R1", "P1"), ("K_right", "dist_right", "R2", "P2"))):
        m1, m2 = orc.init_rectify_map(c.mat(K), c.mat(d), c.mat(R), c.mat(P), (1280, 720))
        assert sha(m1) == str(golden["map1_sha_%d" % eye]) and sha(m2) == str(golden["map2_sha_%d" % eye])
        assert np.array_equal(m1[[0, 359, 719]], golden["map1_rows_%d" % eye])
        rect = orc.remap(sbs[:, eye * 1280:(eye + 1) * 1280], m1, m2)
        assert sha(rect) == str(golden["rect_sha_%d" % eye])
        assert np.array_equal(rect[:48, :64], golden["rect_crop_%d" % eye])
        assert sha(orc.bgr2gray(rect)) == str(golden["rect_gray_sha_%d" % eye])


@pytest.mark.slow
def test_headline_config_full_size(golden, orc):
    """1280x720x128 MODE_SGBM, the configuration BASELINE.json's metric is quoted on (a few seconds of CPU)."""
    from stereo_depth_ruler_b200 import synth
    l, r = synth.stereo_pair(1280, 720, 128, 0)
    assert sha(l) + sha(r) == str(golden["full_c1_inputs_sha"])
    got = orc.sgbm_compute(l, r, orc.make_params())
    assert sha(got) == str(golden["full_c1_sha"])
    assert np.array_equal(got[300:364, 600:728], golden["full_c1_crop"])
