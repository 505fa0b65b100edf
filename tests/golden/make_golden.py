"""Generates the golden fixtures in this directory from the REAL reference arithmetic:
OpenCV (cv2 4.13.0, importable in the build container) driven exactly like the reference drives it
(StereoSGBM::create argument order of stereo_disparity.cpp:5-6 / pcd_write.cpp:102-108, CV_16SC2
rectify maps of stereo_rectifier.cpp:7-11, reprojectImageTo3D of pcd_write.cpp:116) and the
reference's own calibration file config/stereo.yaml. Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference, so the calibration is re-emitted here with the repo's own writer.
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from stereo_depth_ruler_b200 import capi, synth  # noqa: E402

REF_YAML = "/root/reference/config/stereo.yaml"
cv2.setNumThreads(1)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    calib = capi.load_calibration(REF_YAML)
    capi.save_calibration(os.path.join(HERE, "zed2_calibration.yaml"), calib)
    fs = cv2.FileStorage(REF_YAML, cv2.FILE_STORAGE_READ)
    g = lambda k: fs.getNode(k).mat()
    out = {}

    # ---- SGBM cases (inputs from synth; stored so the fixture is self-contained) ----
    cases = []
    def add(name, W, H, args, kind="tex", seed=0):
        D = args[1]
        if kind == "tex":
            l, r = synth.stereo_pair(W, H, D, seed)
        else:
            l, r = synth.noise_pair(W, H, seed, 256 if kind == "noise" else 8)
        disp = cv2.StereoSGBM_create(*args).compute(l, r)
        out[name + "_left"] = l; out[name + "_right"] = r; out[name + "_disp"] = disp
        out[name + "_args"] = np.array(args, np.int32)
        cases.append(name)
    for mode in (0, 1, 2):
        add("sgbm_m%d_ref" % mode, 160, 96, (0, 32, 5, 600, 2400, 1, 63, 12, 200, 2, mode))       # reference tuple, small D
        add("sgbm_m%d_d80" % mode, 200, 64, (0, 80, 5, 600, 2400, 1, 63, 12, 200, 2, mode), seed=1)  # the reference's D=80
        add("sgbm_m%d_bs3" % mode, 150, 75, (0, 48, 3, 216, 864, 2, 31, 5, 50, 1, mode), seed=2)
        add("sgbm_m%d_noise" % mode, 120, 50, (0, 16, 7, 1176, 4704, -1, 5, 0, 0, 0, mode), kind="noise", seed=3)
        add("sgbm_m%d_low" % mode, 120, 50, (0, 32, 9, 1944, 7776, 1, 63, 15, 100, 2, mode), kind="low", seed=4)
        add("sgbm_m%d_neg" % mode, 140, 60, (-8, 32, 5, 600, 2400, 1, 63, 12, 200, 2, mode), seed=5)
    out["sgbm_cases"] = np.array(cases)

    # ---- headline config: 1280x720x128 MODE_SGBM (hash + crops; full map is 1.8 MB) ----
    l, r = synth.stereo_pair(1280, 720, 128, 0)
    full = cv2.StereoSGBM_create(0, 128, 5, 600, 2400, 1, 63, 12, 200, 2, 0).compute(l, r)
    out["full_c1_sha"] = np.array(sha(full)); out["full_c1_inputs_sha"] = np.array(sha(l) + sha(r))
    out["full_c1_crop"] = full[300:364, 600:728].copy()
    out["full_c1_valid"] = np.array(int((full >= 0).sum()))
    l3, r3 = synth.stereo_pair(1280, 720, 80, 1)
    full3 = cv2.StereoSGBM_create(0, 80, 5, 600, 2400, 1, 63, 12, 200, 2, 2).compute(l3, r3)     # the reference's own tuple
    out["full_3way_sha"] = np.array(sha(full3))

    # ---- rectification on the reference calibration ----
    for eye, (K, d, R, P) in enumerate((("cameraMatrixLeft", "distCoeffsLeft", "R1", "P1"),
                                         ("cameraMatrixRight", "distCoeffsRight", "R2", "P2"))):
        m1, m2 = cv2.initUndistortRectifyMap(g(K), g(d), g(R), g(P), (1280, 720), cv2.CV_16SC2)
        out["map1_sha_%d" % eye] = np.array(sha(m1)); out["map2_sha_%d" % eye] = np.array(sha(m2))
        out["map1_rows_%d" % eye] = m1[[0, 359, 719]].copy(); out["map2_rows_%d" % eye] = m2[[0, 359, 719]].copy()
    gl, gr = synth.stereo_pair(1280, 720, 128, 7)
    sbs = np.concatenate([synth.colorize(gl, 1), synth.colorize(gr, 2)], axis=1)
    for eye in range(2):
        K, d, R, P = (("cameraMatrixLeft", "distCoeffsLeft", "R1", "P1"), ("cameraMatrixRight", "distCoeffsRight", "R2", "P2"))[eye]
        m1, m2 = cv2.initUndistortRectifyMap(g(K), g(d), g(R), g(P), (1280, 720), cv2.CV_16SC2)
        view = sbs[:, eye * 1280:(eye + 1) * 1280]          # non-contiguous ROI, as stereo_displayer.cpp:155-156
        rect = cv2.remap(view, m1, m2, cv2.INTER_LINEAR)
        out["rect_sha_%d" % eye] = np.array(sha(rect))
        out["rect_crop_%d" % eye] = rect[:48, :64].copy()     # includes the border region
        out["rect_gray_sha_%d" % eye] = np.array(sha(cv2.cvtColor(rect, cv2.COLOR_BGR2GRAY)))
    out["sbs_sha"] = np.array(sha(sbs))

    # ---- small stages ----
    rng = np.random.default_rng(11)
    bgr = rng.integers(0, 256, size=(37, 53, 3)).astype(np.uint8)
    out["gray_in"] = bgr; out["gray_out"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    x = rng.integers(-16, 2000, size=(41, 67)).astype(np.int16)
    out["median_in"] = x; out["median_out"] = cv2.medianBlur(x, 3)
    y = (rng.integers(0, 40, size=(60, 80)) * 20).astype(np.int16); y[rng.random(y.shape) < 0.2] = -16
    z = y.copy(); cv2.filterSpeckles(z, -16, 200, 32)
    out["speckle_in"] = y; out["speckle_out"] = z
    Q = g("Q")
    d16 = out["sgbm_m0_ref_disp"]
    df = d16.astype(np.float32) / 16
    out["reproj_disp16"] = d16
    out["reproj_false"] = cv2.reprojectImageTo3D(df, Q, handleMissingValues=False).view(np.uint32)
    out["reproj_true"] = cv2.reprojectImageTo3D(df, Q, handleMissingValues=True).view(np.uint32)
    out["Q"] = Q
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote golden.npz:", os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes,", len(out), "arrays; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
