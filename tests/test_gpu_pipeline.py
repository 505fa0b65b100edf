"""GPU: the whole path per frame (SBS frame -> rectify -> SGBM -> reproject -> VoxelGrid) against the oracle chain,
and the host-side interface mirror."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

Z_MAX = 5000.0   # mm; a stray sub-pixel disparity would otherwise blow the bounding box past PCL's int32 guard


def _oracle_chain(orc, calib, sbs, args, leaf):
    W = calib.image_width
    rect = []
    for eye, (K, d, R, P) in enumerate((("K_left", "dist_left", "R1", "P1"), ("K_right", "dist_right", "R2", "P2"))):
        m1, m2 = orc.init_rectify_map(calib.mat(K), calib.mat(d), calib.mat(R), calib.mat(P), (W, calib.image_height))
        rect.append(orc.remap(sbs[:, eye * W:(eye + 1) * W], m1, m2))
    d16 = orc.sgbm_compute(orc.bgr2gray(rect[0]), orc.bgr2gray(rect[1]), orc.make_params(*args))
    pts = orc.reproject_to_3d(orc.disp_to_float(d16), calib.mat("Q"), True)
    drop = (d16 < args[0] * 16) | ~((pts[..., 2] > 0) & (pts[..., 2] <= Z_MAX))
    cloud = orc.voxelgrid(pts, rect[0], leaf, drop_mask=drop)
    return rect[0], d16, pts, cloud


def _raw_sbs(seed):
    """A raw (unrectified-looking) SBS frame: textured pair, colourised. The oracle chain and the CUDA path both
    rectify the same bytes, so no inverse warp is needed for parity."""
    from stereo_depth_ruler_b200 import synth
    l, r = synth.stereo_pair(1280, 720, 128, seed)
    return np.concatenate([synth.colorize(l, seed), synth.colorize(r, seed + 1)], axis=1)


def test_pipeline_batch_matches_oracle_chain(ctx, orc, calib_path):
    from stereo_depth_ruler_b200 import capi
    calib = capi.load_calibration(calib_path)
    args = (0, 128, 5, 600, 2400, 1, 63, 12, 200, 2, 0)
    p = capi.Pipeline(ctx, calib, capi.make_params(*args), capi.voxel_options(5.0, drop_invalid=True, z_max=Z_MAX), handle_missing=True, batch=2)
    frames = [_raw_sbs(21), _raw_sbs(23)]
    disp = [np.empty((720, 1280), np.int16) for _ in frames]
    pts = [np.empty((720, 1280, 3), np.float32) for _ in frames]
    lrect = [np.empty((720, 1280, 3), np.uint8) for _ in frames]
    rc = p.run(frames, disp=disp, points=pts, left_rect=lrect)
    assert rc == capi.SDR_OK
    for i, f in enumerate(frames[:1]):      # one full-size oracle chain is ~10 s of CPU
        rl, d16, pref, cref = _oracle_chain(orc, calib, f, args, 5.0)
        assert np.array_equal(lrect[i], rl)
        assert np.array_equal(disp[i], d16)
        assert np.array_equal(pts[i].view(np.uint32), pref.view(np.uint32))
        got = p.cloud_to_host(i)
        assert np.array_equal(got["idx"], cref["idx"]) and np.array_equal(got["cnt"], cref["cnt"])
        assert np.allclose(got["xyz"], cref["xyz"], rtol=1e-4, atol=0) and np.array_equal(got["rgb"], cref["rgb"])
    # frames are independent: the second slot equals a batch-of-one run of the same frame
    p1 = capi.Pipeline(ctx, calib, capi.make_params(*args), capi.voxel_options(5.0, drop_invalid=True, z_max=Z_MAX), handle_missing=True, batch=1)
    d1 = [np.empty((720, 1280), np.int16)]
    p1.run(frames[1:], disp=d1)
    assert np.array_equal(d1[0], disp[1])
    assert int(p1.views[0].n) == int(p.views[1].n)
    with pytest.raises(capi.SdrError):
        p1.run(frames)                      # more frames than the batch the pipeline was sized for
    with pytest.raises(capi.SdrError):
        p1.run([frames[0][:, :2000]])       # not a 2*W SBS frame
    p.close(); p1.close()


def test_interface_mirror_reads_like_the_reference(ctx, orc, calib_path):
    """StereoConfiguration / StereoRectifier / StereoDisparity / VoxelGrid used the way stereo_displayer.cpp:125-162 and
    pcd_write.cpp:82-130 use the reference's classes."""
    from stereo_depth_ruler_b200 import StereoConfiguration, StereoRectifier, StereoDisparity, VoxelGrid, capi
    config = StereoConfiguration()
    assert config.loadFromFile(calib_path)
    rectifier = StereoRectifier(config, ctx)
    disparity_computer = StereoDisparity(config.Q, ctx)       # reference tuple: D=80, MODE_SGBM_3WAY
    assert disparity_computer.get_matcher().getNumDisparities() == 80
    frame = _raw_sbs(31)
    left_raw, right_raw = frame[:, :frame.shape[1] // 2], frame[:, frame.shape[1] // 2:]
    left_rect, right_rect = rectifier.rectify(left_raw, right_raw)
    disp_float = disparity_computer.computeDisparity(left_rect, right_rect)
    depth_map = disparity_computer.computeDepth(disp_float)
    assert disp_float.dtype == np.float32 and depth_map.shape == (720, 1280, 3)
    args = (0, 80, 5, 600, 2400, 1, 63, 12, 200, 2, 2)
    calib = capi.load_calibration(calib_path)
    rl, d16, _, _ = _oracle_chain(orc, calib, frame, args, 5.0)
    assert np.array_equal(left_rect, rl)
    assert np.array_equal(disp_float, d16.astype(np.float32) / 16)
    ref_pts = orc.reproject_to_3d(orc.disp_to_float(d16), config.Q, False)
    assert np.array_equal(depth_map.view(np.uint32), ref_pts.view(np.uint32))
    assert rectifier.rectify(np.empty((0, 0, 3), np.uint8), right_raw) is None     # empty input: message + return
    vg = VoxelGrid(ctx)
    vg.setLeafSize(0.005, 0.005, 0.005)                       # the reference's literal leaf (pcd_write.cpp:123) on a mm cloud
    pts_true = disparity_computer.computeDepth(disp_float, handleMissingValues=True)
    out = vg.filter(pts_true, left_rect)
    ref = orc.voxelgrid(pts_true, left_rect, 0.005)
    if ref is None:     # PCL's int32 guard trips: the cloud is passed through unchanged
        assert out["passthrough"] and out["xyz"].shape[0] == 720 * 1280
    else:               # the int64 product wrapped below INT32_MAX: PCL voxelises with wrapped indices, and so do we
        assert not out["passthrough"]
        assert np.array_equal(out["idx"], ref["idx"]) and np.array_equal(out["cnt"], ref["cnt"])
        assert np.allclose(out["xyz"], ref["xyz"], rtol=1e-4, atol=0)


def test_stream_runner_shards_frames_over_devices(ctx, orc, calib_path):
    """sdr_stream_run (BASELINE config 4): worker threads per (device, handle) pull frame batches from a shared cursor;
    every frame is processed exactly once and equals the single-pipeline result, whatever device took it."""
    import threading
    import torch
    from stereo_depth_ruler_b200 import capi
    calib = capi.load_calibration(calib_path)
    args = (0, 128, 5, 600, 2400, 1, 63, 12, 200, 2, 0)
    vox = capi.voxel_options(5.0, drop_invalid=True, z_max=Z_MAX)
    frames = [_raw_sbs(41), _raw_sbs(43), _raw_sbs(47)]
    # expected results from one ordinary pipeline
    p = capi.Pipeline(ctx, calib, capi.make_params(*args), vox, handle_missing=True, batch=1)
    expect = []
    for f in frames:
        d = [np.empty((720, 1280), np.int16)]
        p.run([f], disp=d)
        expect.append((d[0].copy(), int(p.views[0].n)))
    p.close()
    n_dev = torch.cuda.device_count()
    devices = list(range(min(n_dev, 4)))
    seen, lock = {}, threading.Lock()

    def on_result(i, dev, disp, n_cloud):
        ok = bool(np.array_equal(disp, expect[i % 3][0])) and n_cloud == expect[i % 3][1]
        with lock:
            seen.setdefault(i, []).append((dev, ok))

    n_frames = 22
    st = capi.stream_run(devices, 2, calib, capi.make_params(*args), vox, frames, n_frames=n_frames, batch=3, on_result=on_result)
    assert st.frames == n_frames and sorted(seen) == list(range(n_frames))
    assert all(len(v) == 1 and v[0][1] for v in seen.values())           # exactly once, bit-identical
    assert sum(st.per_device[:len(devices)]) == n_frames and st.frames_per_s > 0
    if len(devices) > 1:
        assert len({v[0][0] for v in seen.values()}) > 1                   # more than one GPU took frames


def test_device_image_on_another_gpu_is_a_peer_access_or_an_error(ctx, calib_path):
    """A device-resident frame pool that lives on GPU 0 read by a worker on GPU 1 (found by tests/fuzz/fuzz_stream.py on a 2-GPU
    box: the kernels used the pointer in place and faulted): peer access is enabled on first use, or the call is refused."""
    import torch
    from stereo_depth_ruler_b200 import capi
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    calib = capi.load_calibration(calib_path)
    args = (0, 64, 5, 600, 2400, 1, 63, 12, 100, 2, 0)
    vox = capi.voxel_options(5.0, drop_invalid=True, z_max=Z_MAX)
    frames = np.stack([_raw_sbs(51), _raw_sbs(53)])
    pool0 = torch.from_numpy(frames).to("cuda:0")
    torch.cuda.synchronize()
    s = capi.Stream([0, 1], 1, calib, capi.make_params(*args), vox, handle_missing=True, batch=1, download_clouds=False)
    st = s.process_pool(8, pool0)
    assert st.frames == 8 and st.per_device[0] + st.per_device[1] == 8
    s.close()
    # one handle on GPU 1 reading GPU 0's memory directly: same map as from host memory
    c1 = capi.Context(1)
    p = capi.Pipeline(c1, calib, capi.make_params(*args), vox, handle_missing=True, batch=1)
    d_host, d_peer = [np.empty((720, 1280), np.int16)], [np.empty((720, 1280), np.int16)]
    p.run([frames[0]], disp=d_host)
    p.run([pool0[0]], disp=d_peer)
    assert np.array_equal(d_host[0], d_peer[0])
    p.close(); c1.close()
