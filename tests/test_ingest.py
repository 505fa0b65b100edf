"""Frame ingest: NV12 / YUYV -> BGR, the conversion between a hardware decoder or the UVC camera and
the BGR side-by-side frame of stereo_displayer.cpp:145-156 / pcd_write.cpp:59-84. CPU: the oracle against the cv2 fixtures
(tests/golden/make_ingest_golden.py) and against cv2 itself; GPU: sdr_ingest_to_bgr and the pipeline fed with NV12 / YUYV."""
import hashlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NV12, YUYV = 1, 2


@pytest.fixture(scope="module")
def ig():
    return np.load(os.path.join(ROOT, "tests", "golden", "ingest_golden.npz"))


def all_triples_sha(convert):
    """sha256 of convert(NV12 picture) over all 2^24 (Y, U, V) triples, laid out as make_ingest_golden.all_triples_bgr."""
    h = hashlib.sha256()
    Y = np.repeat(np.repeat(np.arange(256, dtype=np.uint8), 2)[None, :], 512, 0)
    for u in range(256):
        uv = np.empty((256, 512), np.uint8)
        uv[:, 0::2] = u
        uv[:, 1::2] = np.arange(256, dtype=np.uint8)[:, None]
        h.update(np.ascontiguousarray(convert(np.concatenate([Y, uv], 0))).tobytes())
    return h.hexdigest()


def bgr_to_nv12(bgr):
    """A BGR picture as a decoder would deliver it (BT.601 limited range, 2x2 chroma average)."""
    import cv2
    H, W = bgr.shape[:2]
    i420 = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV_I420)
    out = np.empty((H * 3 // 2, W), np.uint8)
    out[:H] = i420[:H]
    u = i420[H:H + H // 4].reshape(H // 2, W // 2)
    v = i420[H + H // 4:].reshape(H // 2, W // 2)
    out[H:, 0::2] = u
    out[H:, 1::2] = v
    return out


def nv12_to_yuyv(nv, H):
    W = nv.shape[1]
    yu = np.empty((H, W, 2), np.uint8)
    yu[..., 0] = nv[:H]
    yu[:, 0::2, 1] = np.repeat(nv[H:, 0::2], 2, 0)
    yu[:, 1::2, 1] = np.repeat(nv[H:, 1::2], 2, 0)
    return yu


# ---- CPU: the oracle is pinned ------------------------------------------------------------------------------------------
def test_oracle_equals_cv2_fixtures(orc, ig):
    assert np.array_equal(orc.ingest_to_bgr(NV12, ig["nv12"]), ig["nv12_bgr"])
    assert np.array_equal(orc.ingest_to_bgr(YUYV, ig["yuyv"]), ig["yuyv_bgr"])
    for s, ref in zip(ig["sweep"], ig["sweep_bgr"]):
        assert np.array_equal(orc.ingest_to_bgr(NV12, s), ref)


@pytest.mark.slow
def test_oracle_equals_cv2_on_every_yuv_triple(orc, ig):
    assert all_triples_sha(lambda nv: orc.ingest_to_bgr(NV12, nv)) == str(ig["all_triples_sha256"])


def test_oracle_equals_cv2_directly(orc):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    nv = rng.integers(0, 256, (30, 52), dtype=np.uint8)
    assert np.array_equal(orc.ingest_to_bgr(NV12, nv), cv2.cvtColor(nv, cv2.COLOR_YUV2BGR_NV12))
    yu = rng.integers(0, 256, (7, 18, 2), dtype=np.uint8)                         # odd height is fine for 4:2:2
    assert np.array_equal(orc.ingest_to_bgr(YUYV, yu), cv2.cvtColor(yu, cv2.COLOR_YUV2BGR_YUY2))
    with pytest.raises(ValueError):                                                # cv2 asserts on odd widths too
        orc.ingest_to_bgr(YUYV, yu[:, :17])


# ---- GPU ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_ingest_equals_cv2_fixtures_and_every_triple(ctx, ig):
    from stereo_depth_ruler_b200 import capi
    assert np.array_equal(capi.ingest_to_bgr(ctx, capi.PIX_NV12, ig["nv12"]), ig["nv12_bgr"])
    assert np.array_equal(capi.ingest_to_bgr(ctx, capi.PIX_YUYV, ig["yuyv"]), ig["yuyv_bgr"])
    for s, ref in zip(ig["sweep"], ig["sweep_bgr"]):
        assert np.array_equal(capi.ingest_to_bgr(ctx, capi.PIX_NV12, s), ref)
    assert all_triples_sha(lambda nv: capi.ingest_to_bgr(ctx, capi.PIX_NV12, nv)) == str(ig["all_triples_sha256"])


@pytest.mark.gpu
def test_gpu_ingest_ragged_views_and_device_tensors(ctx, orc):
    """Widths that are not a multiple of 4, ROI views with odd byte offsets (the generic kernel), device-resident tensors,
    a full 2560x720 SBS frame; error behaviour."""
    import torch
    from stereo_depth_ruler_b200 import capi
    rng = np.random.default_rng(5)
    for W, H in ((6, 4), (34, 10), (130, 66)):
        nv = rng.integers(0, 256, (H * 3 // 2, W), dtype=np.uint8)
        assert np.array_equal(capi.ingest_to_bgr(ctx, capi.PIX_NV12, nv), orc.ingest_to_bgr(NV12, nv))
        yu = rng.integers(0, 256, (H + 1, W, 2), dtype=np.uint8)
        assert np.array_equal(capi.ingest_to_bgr(ctx, capi.PIX_YUYV, yu), orc.ingest_to_bgr(YUYV, yu))
    big = rng.integers(0, 256, (40, 70, 2), dtype=np.uint8)
    view = big[3:23, 5:45]                                                         # pitched, starts at an odd byte offset
    assert np.array_equal(capi.ingest_to_bgr(ctx, capi.PIX_YUYV, view), orc.ingest_to_bgr(YUYV, np.ascontiguousarray(view)))
    sbs = rng.integers(0, 256, (1080, 2560), dtype=np.uint8)                       # NV12 SBS frame of the ZED 2 (2560x720)
    t = torch.from_numpy(sbs).cuda()
    out = capi.ingest_to_bgr(ctx, capi.PIX_NV12, t)
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), orc.ingest_to_bgr(NV12, sbs))
    with pytest.raises(capi.SdrError):
        capi.ingest_to_bgr(ctx, capi.PIX_NV12, sbs[:, :2559])                      # odd width
    with pytest.raises(capi.SdrError):
        capi.ingest_to_bgr(ctx, 7, sbs)                                            # unknown format


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", [NV12, YUYV])
def test_pipeline_fed_with_decoder_frames_equals_cv2_chain(ctx, calib_path, fmt):
    """sdr_pipeline_run on NV12 / YUYV SBS frames == cv2.cvtColor -> the same pipeline on the BGR frames, and the CV_16S map
    == cv2 remap + StereoSGBM on that BGR frame (the full reference chain starting at the decoder's output)."""
    cv2 = pytest.importorskip("cv2")
    from stereo_depth_ruler_b200 import capi, synth
    W, H, D = 1280, 720, 64
    args = (0, D, 5, 600, 2400, 1, 63, 12, 200, 2, 0)
    calib = capi.load_calibration(calib_path)
    frames, bgrs = [], []
    for s in range(2):
        l, r = synth.stereo_pair(W, H, D, 20 + s)
        sbs = np.concatenate([synth.colorize(l, s), synth.colorize(r, s + 1)], axis=1)
        nv = bgr_to_nv12(sbs)
        f = nv if fmt == NV12 else nv12_to_yuyv(nv, H)
        frames.append(f)
        bgrs.append(cv2.cvtColor(f, cv2.COLOR_YUV2BGR_NV12 if fmt == NV12 else cv2.COLOR_YUV2BGR_YUY2))
    vox = capi.voxel_options(5.0, drop_invalid=True, z_max=5000.0)
    pa = capi.Pipeline(ctx, calib, capi.make_params(*args), vox, handle_missing=True, batch=2)
    pb = capi.Pipeline(ctx, calib, capi.make_params(*args), vox, handle_missing=True, batch=2)
    pa.set_input_format(fmt)
    da, db = np.empty((2, H, W), np.int16), np.empty((2, H, W), np.int16)
    la, lb = np.empty((2, H, W, 3), np.uint8), np.empty((2, H, W, 3), np.uint8)
    pa.run(frames, disp=list(da), left_rect=list(la))
    pb.run(bgrs, disp=list(db), left_rect=list(lb))
    assert np.array_equal(da, db) and np.array_equal(la, lb)
    assert [int(pa.views[i].n) for i in range(2)] == [int(pb.views[i].n) for i in range(2)]
    ca, cb = pa.cloud_to_host(1), pb.cloud_to_host(1)
    assert np.array_equal(ca["idx"], cb["idx"]) and np.array_equal(ca["rgb"], cb["rgb"])
    # the reference chain on frame 0 with cv2 itself
    rect = capi.Rectifier(ctx, calib)
    m1l, m2l = rect.maps(0)
    m1r, m2r = rect.maps(1)
    rect.close()
    gl = cv2.cvtColor(cv2.remap(bgrs[0][:, :W], m1l, m2l, cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
    gr = cv2.cvtColor(cv2.remap(bgrs[0][:, W:], m1r, m2r, cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
    assert np.array_equal(da[0], cv2.StereoSGBM_create(*args[:10], mode=0).compute(gl, gr))
    with pytest.raises(capi.SdrError):                                             # not the rows a frame of this format has
        pa.run([f[:H - 2] for f in frames])
    pa.close(); pb.close()


@pytest.mark.gpu
def test_stream_runner_with_nv12_pool(ctx, calib_path):
    """The native frame-parallel runner over a pool of NV12 frames: same number of frames, H2D bytes = 1.5 B per pixel."""
    from stereo_depth_ruler_b200 import capi, synth
    W, H, D = 1280, 720, 64
    calib = capi.load_calibration(calib_path)
    l, r = synth.stereo_pair(W, H, D, 31)
    sbs = np.concatenate([synth.colorize(l, 0), synth.colorize(r, 1)], axis=1)
    pool = np.stack([bgr_to_nv12(sbs), bgr_to_nv12(np.roll(sbs, 5, axis=0))])
    st = capi.Stream([0], 2, calib, capi.make_params(0, D, 5, 600, 2400, 1, 63, 12, 200, 2, 0),
                     capi.voxel_options(5.0, drop_invalid=True, z_max=5000.0), handle_missing=True, batch=2, download_clouds=True)
    st.set_input_format(capi.PIX_NV12)
    stats = st.process_pool(6, pool, mem=capi.MEM_HOST)
    st.close()
    assert stats.frames == 6 and stats.h2d_bytes == 6 * (H * 3 // 2) * 2 * W
