"""GPU: the C++ facade (include/sdr/stereo_vision.hpp) used the way the reference's callers use its classes -- compiled
with plain g++, run on the GPU, every dumped buffer compared with cv2 / the oracle."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CALIB = os.path.join(ROOT, "tests", "golden", "zed2_calibration.yaml")


def test_cpp_facade_runs_the_path_on_the_gpu(tmp_path, orc):
    cv2 = pytest.importorskip("cv2")
    from oracle import calib as ocalib
    from stereo_depth_ruler_b200 import synth
    W, H = 1280, 720
    l, r = synth.stereo_pair(W, H, 128, 3)
    sbs = np.ascontiguousarray(np.concatenate([synth.colorize(l, 3), synth.colorize(r, 4)], axis=1))
    (tmp_path / "sbs.raw").write_bytes(sbs.tobytes())
    exe = tmp_path / "facade_demo"
    pkg = os.path.join(ROOT, "stereo_depth_ruler_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "facade_demo.cpp"),
                           "-o", str(exe), "-L", pkg, "-lsdr_b200", "-Wl,-rpath," + pkg])
    p = subprocess.run([str(exe), CALIB, str(tmp_path / "sbs.raw"), str(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr + p.stdout
    assert "half=640x360" in p.stdout and "wls=0" in p.stdout and "matcher_numdisp=80" in p.stdout, p.stdout

    def load(name, dtype, shape):
        return np.fromfile(str(tmp_path / name), dtype).reshape(shape)
    c = ocalib.load(CALIB)
    maps = [cv2.initUndistortRectifyMap(c[K], c[d], c[R], c[P], (W, H), cv2.CV_16SC2)
            for K, d, R, P in (("K_left", "dist_left", "R1", "P1"), ("K_right", "dist_right", "R2", "P2"))]
    lr = cv2.remap(sbs[:, :W], maps[0][0], maps[0][1], cv2.INTER_LINEAR)
    rr = cv2.remap(sbs[:, W:], maps[1][0], maps[1][1], cv2.INTER_LINEAR)
    assert np.array_equal(load("left_rect.u8", np.uint8, (H, W, 3)), lr)
    assert np.array_equal(load("right_rect.u8", np.uint8, (H, W, 3)), rr)
    lg, rg = cv2.cvtColor(lr, cv2.COLOR_BGR2GRAY), cv2.cvtColor(rr, cv2.COLOR_BGR2GRAY)
    # live path (stereo_disparity.cpp:19-28, :34) up to the WLS filter
    ls = cv2.resize(lg, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
    rs = cv2.resize(rg, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
    m = cv2.StereoSGBM_create(0, 80, 5, 600, 2400, 1, 63, 12, 200, 2, cv2.STEREO_SGBM_MODE_SGBM_3WAY)
    mr = cv2.StereoSGBM_create(-79, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, cv2.STEREO_SGBM_MODE_SGBM_3WAY)
    dh = m.compute(ls, rs).astype(np.float32) / 16
    assert np.array_equal(load("disp_half.f32", np.float32, (H // 2, W // 2)), dh)
    assert np.array_equal(load("disp_right_half.s16", np.int16, (H // 2, W // 2)), mr.compute(rs, ls))
    from test_gpu_vis import _ref_show_disparity
    assert np.array_equal(load("vis_disp.u8", np.uint8, (H // 2, W // 2)), _ref_show_disparity(dh, 80, None))
    # pcd_write path (pcd_write.cpp:88-130)
    d = m.compute(lg, rg).astype(np.float32) / 16
    assert np.array_equal(load("disp.f32", np.float32, (H, W)), d)
    pts = cv2.reprojectImageTo3D(d, c["Q"], handleMissingValues=True)
    assert np.array_equal(load("depth.f32", np.float32, (H, W, 3)).view(np.uint32), pts.view(np.uint32))
    drop = ~((pts[..., 2] > 0) & (pts[..., 2] <= 5000.0))
    ref = orc.voxelgrid(pts, lr, 5.0, drop_mask=drop)
    xyz = np.fromfile(str(tmp_path / "cloud_xyz.f32"), np.float32).reshape(-1, 3)
    rgb = np.fromfile(str(tmp_path / "cloud_rgb.u8"), np.uint8).reshape(-1, 3)
    assert len(xyz) == len(ref["idx"]) and np.allclose(xyz, ref["xyz"], rtol=1e-4, atol=0) and np.array_equal(rgb, ref["rgb"])
    assert (tmp_path / "vis_depth.u8").stat().st_size == H * W * 3
