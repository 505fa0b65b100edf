"""GPU parity for rectification, grayscale, reprojection and the voxel grid, through the C ABI."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def sbs(golden):
    from stereo_depth_ruler_b200 import synth
    gl, gr = synth.stereo_pair(1280, 720, 128, 7)
    s = np.concatenate([synth.colorize(gl, 1), synth.colorize(gr, 2)], axis=1)
    assert sha(s) == str(golden["sbs_sha"])
    return s


def test_rectify_maps_match_cv2(ctx, golden, calib_path):
    """The fp64 map evaluation inside the kernel reproduces cv::initUndistortRectifyMap(CV_16SC2) on the reference's
    calibration, both eyes (sha256 of the full maps + three full rows)."""
    from stereo_depth_ruler_b200 import capi
    r = capi.Rectifier(ctx, capi.load_calibration(calib_path))
    for eye in range(2):
        m1, m2 = r.maps(eye)
        assert np.array_equal(m1[[0, 359, 719]], golden["map1_rows_%d" % eye])
        assert np.array_equal(m2[[0, 359, 719]], golden["map2_rows_%d" % eye])
        assert sha(m1) == str(golden["map1_sha_%d" % eye]) and sha(m2) == str(golden["map2_sha_%d" % eye])
    r.close()


def test_rectify_sbs_views(ctx, golden, calib_path, sbs, orc):
    """rectify() on the two non-contiguous ROI views of a 2560x720 SBS frame (stereo_displayer.cpp:155-159), BGR,
    with the fused BGR2GRAY outputs."""
    from stereo_depth_ruler_b200 import capi
    r = capi.Rectifier(ctx, capi.load_calibration(calib_path))
    ld, rd, lg, rg = r.rectify(sbs[:, :1280], sbs[:, 1280:], want_gray=True)
    for eye, (d, g) in enumerate(((ld, lg), (rd, rg))):
        assert np.array_equal(d[:48, :64], golden["rect_crop_%d" % eye])
        assert sha(d) == str(golden["rect_sha_%d" % eye])
        assert sha(g) == str(golden["rect_gray_sha_%d" % eye])
    # gray input, device-resident tensors
    import torch
    gl = torch.from_numpy(np.ascontiguousarray(sbs[:, :1280, 1])).cuda()
    gr = torch.from_numpy(np.ascontiguousarray(sbs[:, 1280:, 1])).cuda()
    a, b = r.rectify(gl, gr)
    m1, m2 = r.maps(0)
    assert np.array_equal(a.cpu().numpy(), orc.remap(np.ascontiguousarray(sbs[:, :1280, 1]), m1, m2))
    r.close()


def test_remap_with_caller_maps_border_heavy(ctx, orc, golden, calib_path, sbs):
    from stereo_depth_ruler_b200 import capi
    r = capi.Rectifier(ctx, capi.load_calibration(calib_path))
    m1, m2 = r.maps(1)
    ms = m1.copy(); ms[..., 0] -= 100; ms[..., 1] += 30        # shifted map: many taps outside the image
    src = np.ascontiguousarray(sbs[:, 1280:])
    assert np.array_equal(capi.remap_with_maps(ctx, src, ms, m2), orc.remap(src, ms, m2))
    small = np.ascontiguousarray(src[:101, :77])               # odd sizes, source smaller than the map range
    assert np.array_equal(capi.remap_with_maps(ctx, small, ms[:50, :60], m2[:50, :60]), orc.remap(small, ms[:50, :60], m2[:50, :60]))
    g = np.ascontiguousarray(src[..., 0])
    assert np.array_equal(capi.remap_with_maps(ctx, g, ms, m2), orc.remap(g, ms, m2))
    r.close()


def test_bgr2gray(ctx, golden):
    from stereo_depth_ruler_b200 import capi
    assert np.array_equal(capi.bgr2gray(ctx, golden["gray_in"]), golden["gray_out"])


def test_reproject_bit_patterns(ctx, golden, orc):
    """reprojectImageTo3D incl. OpenCV's float-then-reciprocal rounding and the inf/NaN positions; CV_16S input has
    the convertTo(CV_32F, 1/16) fused."""
    from stereo_depth_ruler_b200 import capi
    d16 = golden["reproj_disp16"]
    Q = golden["Q"]
    for hm, key in ((False, "reproj_false"), (True, "reproj_true")):
        got = capi.reproject_to_3d(ctx, d16, Q, hm)
        assert np.array_equal(got.view(np.uint32), golden[key]), key
        got32 = capi.reproject_to_3d(ctx, d16.astype(np.float32) / 16, Q, hm)
        assert np.array_equal(got32.view(np.uint32), golden[key]), key
    # tolerance the north star states for floating point: 1e-4 relative (we are in fact bit-exact)
    ref = orc.reproject_to_3d(orc.disp_to_float(d16), Q, False)
    got = capi.reproject_to_3d(ctx, d16, Q, False)
    fin = np.isfinite(ref)
    assert np.array_equal(fin, np.isfinite(got))
    assert np.allclose(got[fin], ref[fin], rtol=1e-4, atol=0)
    d = d16.copy(); d[d < 0] = 37                              # no invalid pixel: the smallest valid disparity gets Z=10000
    ref = orc.reproject_to_3d(orc.disp_to_float(d), Q, True)
    assert np.array_equal(capi.reproject_to_3d(ctx, d, Q, True).view(np.uint32), ref.view(np.uint32))


def _scene(golden, W=320, H=180, D=64, seed=3):
    from stereo_depth_ruler_b200 import synth
    import numpy as np
    l, r = synth.stereo_pair(W, H, D, seed)
    return l, r, synth.colorize(l, seed)


def _check_cloud(got, ref, rtol=1e-4):
    assert not got["passthrough"]
    assert np.array_equal(got["idx"], ref["idx"])            # identical occupied-voxel set, ascending idx (PCL's order)
    assert np.array_equal(got["cnt"], ref["cnt"])
    assert np.array_equal(got["minb"], ref["minb"]) and np.array_equal(got["divb"], ref["divb"])
    assert np.allclose(got["xyz"], ref["xyz"], rtol=rtol, atol=0)   # tolerance stated by the north star
    assert np.array_equal(got["rgb"], ref["rgb"])


def test_voxelgrid_measurement_mode(ctx, orc, golden):
    """Mode B of SURVEY 8(d): invalid disparities dropped, 5 mm leaf in the cloud's mm units."""
    from stereo_depth_ruler_b200 import capi
    l, r, bgr = _scene(golden)
    args = (0, 64, 5, 600, 2400, 1, 63, 12, 200, 2, 0)
    d16 = orc.sgbm_compute(l, r, orc.make_params(*args))
    Q = golden["Q"]
    pts = orc.reproject_to_3d(orc.disp_to_float(d16), Q, True)
    drop = (d16 < 0)
    ref = orc.voxelgrid(pts, bgr, 5.0, drop_mask=drop)
    assert ref is not None and len(ref["idx"]) > 1000
    cloud = capi.Cloud(ctx)
    opt = capi.voxel_options(5.0, drop_invalid=True, min_disparity=0)
    _check_cloud(cloud.voxelgrid(pts, bgr, opt, disp16=d16), ref)
    # fused reproject + voxelise straight from the CV_16S map
    _check_cloud(cloud.depth_cloud(d16, bgr, Q, opt, handle_missing=True), ref)
    # z_max clip
    zdrop = drop | ~((pts[..., 2] > 0) & (pts[..., 2] <= 2500.0))
    ref2 = orc.voxelgrid(pts, bgr, (5.0, 7.0, 11.0), drop_mask=zdrop)
    opt2 = capi.voxel_options((5.0, 7.0, 11.0), drop_invalid=True, z_max=2500.0)
    _check_cloud(cloud.depth_cloud(d16, bgr, Q, opt2, handle_missing=True), ref2)
    # no colour image: colours are zero (pcd_write.cpp:25 hasColor == false)
    ref3 = orc.voxelgrid(pts, None, 5.0, drop_mask=drop)
    _check_cloud(cloud.voxelgrid(pts, None, opt, disp16=d16), ref3)
    cloud.close()


def test_voxelgrid_literal_reference_semantics_passthrough(ctx, orc, golden):
    """Mode A: every finite point incl. the Z=10000 fill; with the reference's literal leaf 0.005 PCL's int32 guard
    trips and the cloud is passed through unchanged (status SDR_EVOXEL_OVERFLOW)."""
    from stereo_depth_ruler_b200 import capi
    l, r, bgr = _scene(golden, 160, 96, 32)
    d16 = orc.sgbm_compute(l, r, orc.make_params(0, 32, 5, 600, 2400, 1, 63, 12, 200, 2, 2))
    Q = golden["Q"]
    pts = orc.reproject_to_3d(orc.disp_to_float(d16), Q, True)
    assert orc.voxelgrid(pts, bgr, 0.005) is None
    cloud = capi.Cloud(ctx)
    got = cloud.voxelgrid(pts, bgr, capi.voxel_options(0.005))
    assert got["passthrough"] and got["xyz"].shape == (160 * 96, 3)
    fin = np.isfinite(pts).all(axis=2).ravel()
    assert np.array_equal(np.isfinite(got["xyz"]).all(axis=1), fin)
    assert np.array_equal(got["xyz"][fin], pts.reshape(-1, 3)[fin])
    assert np.array_equal(got["rgb"][fin], bgr.reshape(-1, 3)[fin][:, ::-1])
    # a leaf large enough for the literal cloud: voxelised, identical to the restatement
    ref = orc.voxelgrid(pts, bgr, 50.0)
    assert ref is not None
    _check_cloud(cloud.voxelgrid(pts, bgr, capi.voxel_options(50.0)), ref)
    # empty cloud (every point non-finite)
    nanpts = np.full((8, 8, 3), np.nan, np.float32)
    got = cloud.voxelgrid(nanpts, None, capi.voxel_options(1.0))
    assert not got["passthrough"] and got["xyz"].shape[0] == 0
    cloud.close()


def test_voxelgrid_full_size_properties(ctx, orc, golden):
    """Headline size: counts sum to the number of valid points, idx strictly ascending, centroids inside their voxel."""
    from stereo_depth_ruler_b200 import capi, synth
    l, r = synth.stereo_pair(1280, 720, 128, 0)
    bgr = synth.colorize(l, 0)
    m = capi.Sgbm(ctx, capi.make_params())
    d16 = m.compute(l, r)
    m.close()
    cloud = capi.Cloud(ctx)
    got = cloud.depth_cloud(d16, bgr, golden["Q"], capi.voxel_options(5.0, drop_invalid=True), handle_missing=True)
    assert not got["passthrough"]
    assert int(got["cnt"].sum()) == int((d16 >= 0).sum())
    assert np.all(np.diff(got["idx"].astype(np.int64)) > 0)
    ijk = np.floor(got["xyz"] * np.float32(1 / 5.0)).astype(np.int64) - got["minb"]
    idx = ijk[:, 0] + ijk[:, 1] * got["divb"][0] + ijk[:, 2] * got["divb"][0] * got["divb"][1]
    assert (idx == got["idx"]).mean() > 0.999      # centroid of points of one voxel lies in that voxel (up to fp32 edge cases)
    ref = orc.voxelgrid(orc.reproject_to_3d(orc.disp_to_float(d16), golden["Q"], True), bgr, 5.0, drop_mask=(d16 < 0))
    _check_cloud(got, ref)
    cloud.close()


def test_ruler_and_coverage_consumers(ctx, orc, golden):
    """the mouse ruler (stereo_displayer.cpp:41-47) and depth_coverage (:105-118) on the device."""
    import torch
    from stereo_depth_ruler_b200 import capi
    d16 = golden["reproj_disp16"]
    pts = capi.reproject_to_3d(ctx, d16, golden["Q"], False)
    H, W = d16.shape
    assert capi.depth_coverage(ctx, pts) == orc.depth_coverage(pts)
    assert capi.depth_coverage(ctx, pts)[0] > 0
    rng = np.random.default_rng(2)
    dev = torch.from_numpy(pts).cuda()
    for _ in range(20):
        (x1, x2), (y1, y2) = rng.integers(0, W, 2), rng.integers(0, H, 2)
        ref = orc.measure_distance(pts, (x1, y1), (x2, y2))
        for src in (pts, dev):
            got = capi.measure_distance(ctx, src, (x1, y1), (x2, y2))
            assert (np.isnan(ref) and np.isnan(got)) or got == ref
    assert capi.depth_coverage(ctx, dev) == orc.depth_coverage(pts)
    with pytest.raises(capi.SdrError):
        capi.measure_distance(ctx, pts, (W, 0), (0, 0))      # the reference ignores clicks outside the image


def test_pcd_binary_writer(ctx, orc, golden, tmp_path):
    """savePCDFileBinary of the filtered cloud (header restated from PCL: parity unpinned)."""
    from stereo_depth_ruler_b200 import capi
    l, r, bgr = _scene(golden, 160, 96, 32)
    d16 = orc.sgbm_compute(l, r, orc.make_params(0, 32, 5, 600, 2400, 1, 63, 12, 200, 2, 0))
    cloud = capi.Cloud(ctx)
    # D=32 at this baseline puts the scene beyond 5.7 m: a 50 mm leaf and a 60 m clip keep the int32 guard quiet
    got = cloud.depth_cloud(d16, bgr, golden["Q"], capi.voxel_options(50.0, drop_invalid=True, z_max=60000.0), handle_missing=True)
    assert not got["passthrough"] and len(got["idx"]) > 100
    path = tmp_path / "frame_00100.pcd"
    cloud.save_pcd_binary(path)
    raw = path.read_bytes()
    head, body = raw.split(b"DATA binary\n", 1)
    n = len(got["idx"])
    assert b"FIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\n" in head
    assert ("WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\n" % (n, n)).encode() in head
    rec = np.frombuffer(body, dtype=np.dtype([("xyz", "<f4", 3), ("rgb", "<u4")]))
    assert len(rec) == n and np.array_equal(rec["xyz"], got["xyz"])
    packed = (0xFF << 24) | (got["rgb"][:, 0].astype(np.uint32) << 16) | (got["rgb"][:, 1].astype(np.uint32) << 8) | got["rgb"][:, 2]
    assert np.array_equal(rec["rgb"], packed)
    empty = capi.Cloud(ctx)
    empty.voxelgrid(np.full((4, 4, 3), np.nan, np.float32), None, capi.voxel_options(1.0))
    with pytest.raises(capi.SdrError):
        empty.save_pcd_binary(tmp_path / "none.pcd")
    cloud.close(); empty.close()


def test_voxelgrid_absurd_leaves_random(ctx, orc):
    """Leaf sizes so small that PCL's int64 index guard wraps around and passes (the reference's literal leaf 0.005 on a millimetre
    cloud is of this kind), or trips: the GPU must take the same branch as the restatement and, when the guard passes, produce the same
    wrapped int32 indices -- including coordinates whose float -> int conversion leaves the int range (x86: INT_MIN). Found by
    tests/fuzz/fuzz_stages.py."""
    from stereo_depth_ruler_b200 import capi
    rng = np.random.default_rng(11)
    passed = tripped = 0
    for case in range(45):
        n = int(rng.integers(50, 3000))
        kind = case % 3
        if kind == 0:
            pts = rng.uniform(-5000, 5000, size=(n, 3))
        elif kind == 1:
            pts = np.stack([rng.uniform(-3000, 3000, n), rng.uniform(-2000, 2000, n), rng.choice([10000.0, 500.0, 2500.0], n) + rng.uniform(0, 50, n)], -1)
        else:
            pts = rng.uniform(-3e5, 3e5, size=(n, 3))            # coordinate / leaf beyond the int range
        pts = pts.astype(np.float32).reshape(1, n, 3)
        leaf = float(rng.choice([1e-5, 1e-3, 3e-4]))
        with np.errstate(all="ignore"):
            ref = orc.voxelgrid(pts, None, leaf)
        cl = capi.Cloud(ctx)
        got = cl.voxelgrid(pts, None, capi.voxel_options(leaf))
        cl.close()
        if ref is None:
            tripped += 1
            assert got["passthrough"], (case, kind, leaf)
            continue
        passed += 1
        assert not got["passthrough"], (case, kind, leaf)
        assert np.array_equal(got["idx"], ref["idx"]) and np.array_equal(got["cnt"], ref["cnt"]), (case, kind, leaf)
        assert np.array_equal(got["xyz"].view(np.uint32), ref["xyz"].view(np.uint32)), (case, kind, leaf)
    assert passed >= 10 and tripped >= 10, (passed, tripped)
