"""VoxelGrid has no importable reference (PCL is absent: parity UNPINNED). This removes the single-author risk at least:
the C oracle, an independent numpy restatement written from SURVEY A.8 alone, and (with a GPU) the CUDA kernels must
agree bit for bit on adversarial clouds and on the full-size scene."""
import os

import numpy as np
import pytest

from voxelgrid_numpy import adversarial_clouds, voxelgrid_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = adversarial_clouds()


def _same(a, b, what):
    if a is None or b is None:
        assert a is None and b is None, "%s: only one side tripped the int32 guard" % what
        return
    assert np.array_equal(a["idx"], b["idx"]), what + ": occupied voxel set"
    assert np.array_equal(a["cnt"], b["cnt"]), what + ": points per voxel"
    assert np.array_equal(a["xyz"].view(np.uint32), b["xyz"].view(np.uint32)), what + ": centroids (bit patterns)"
    assert np.array_equal(a["rgb"], b["rgb"]), what + ": colours"
    assert np.array_equal(a["minb"], b["minb"]) and np.array_equal(a["divb"], b["divb"]), what + ": grid"


@pytest.mark.parametrize("name,pts,col,leaf", CASES, ids=[c[0] for c in CASES])
def test_oracle_equals_independent_numpy_restatement(orc, name, pts, col, leaf):
    _same(orc.voxelgrid(pts.reshape(-1, 1, 3), col.reshape(-1, 1, 3), leaf), voxelgrid_np(pts, col, leaf), name)


def test_oracle_equals_numpy_on_a_depth_scene(orc):
    """A reprojected SGBM scene (what the pipeline feeds the filter), with the drop mask of measurement mode."""
    cv2 = pytest.importorskip("cv2")
    from oracle import calib
    from stereo_depth_ruler_b200 import synth
    l, r = synth.stereo_pair(640, 240, 64, 8)
    d16 = cv2.StereoSGBM_create(0, 64, 5, 600, 2400, 1, 63, 12, 100, 2).compute(l, r)
    pts = cv2.reprojectImageTo3D(d16.astype(np.float32) / 16, calib.load(os.path.join(ROOT, "tests", "golden", "zed2_calibration.yaml"))["Q"], True)
    bgr = synth.colorize(l, 8)
    drop = (d16 < 0) | ~((pts[..., 2] > 0) & (pts[..., 2] <= 5000))
    a, b = orc.voxelgrid(pts, bgr, 5.0, drop_mask=drop), voxelgrid_np(pts, bgr, 5.0, drop)
    assert len(a["idx"]) > 10000
    _same(a, b, "depth scene")
    _same(orc.voxelgrid(pts, bgr, 0.005), voxelgrid_np(pts, bgr, 0.005), "literal leaf 0.005 (guard)")


@pytest.mark.gpu
@pytest.mark.parametrize("name,pts,col,leaf", CASES, ids=[c[0] for c in CASES])
def test_gpu_equals_independent_numpy_restatement(ctx, name, pts, col, leaf):
    from stereo_depth_ruler_b200 import capi
    cloud = capi.Cloud(ctx)
    got = cloud.voxelgrid(pts.reshape(-1, 1, 3), col.reshape(-1, 1, 3), capi.voxel_options(leaf))
    ref = voxelgrid_np(pts, col, leaf)
    if ref is None:
        assert got["passthrough"]
    else:
        assert not got["passthrough"]
        _same(got | dict(minb=got["minb"].astype(np.int32), divb=got["divb"].astype(np.int32)), ref, name)
    cloud.close()


@pytest.mark.gpu
def test_gpu_equals_numpy_on_the_full_size_scene(ctx):
    """1280x720x128 headline scene through the fused reproject + voxel path: ~265 k voxels, identical to the numpy restatement."""
    cv2 = pytest.importorskip("cv2")
    from stereo_depth_ruler_b200 import capi, synth
    calib = capi.load_calibration(os.path.join(ROOT, "tests", "golden", "zed2_calibration.yaml"))
    l, r = synth.stereo_pair(1280, 720, 128, 0)
    d16 = capi.Sgbm(ctx, capi.make_params()).compute(l, r)
    bgr = synth.colorize(l, 0)
    pts = cv2.reprojectImageTo3D(d16.astype(np.float32) / 16, calib.mat("Q"), handleMissingValues=True)
    drop = (d16 < 0) | ~((pts[..., 2] > 0) & (pts[..., 2] <= 5000))
    ref = voxelgrid_np(pts, bgr, 5.0, drop)
    cloud = capi.Cloud(ctx)
    got = cloud.depth_cloud(d16, bgr, calib.mat("Q"), capi.voxel_options(5.0, drop_invalid=True, z_max=5000.0), handle_missing=True)
    assert 200000 < len(ref["idx"]) < 350000
    _same(got | dict(minb=got["minb"].astype(np.int32), divb=got["divb"].astype(np.int32)), ref, "full-size scene")
    cloud.close()
