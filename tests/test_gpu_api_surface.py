"""GPU: the rest of the C-ABI surface -- caller streams, per-kernel timers, pitched device buffers, handle reuse across
sizes, launch counting -- always checked against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_caller_stream_timers_and_launch_count(orc):
    import torch
    from stereo_depth_ruler_b200 import capi, synth
    c = capi.Context(0)
    s = torch.cuda.Stream()
    c.set_stream(s.cuda_stream)
    l, r = synth.stereo_pair(320, 120, 64, 2)
    args = (0, 64, 5, 600, 2400, 1, 63, 12, 100, 2, 0)
    m = capi.Sgbm(c, capi.make_params(*args))
    m.set_profiling(True)
    L, R = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    O = torch.empty((120, 320), dtype=torch.int16, device="cuda")
    n0 = c.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        m.compute(L, R, O)
        e1.record(s)
    s.synchronize()                                   # the library ran on the caller's stream: torch events bracket it
    assert np.array_equal(O.cpu().numpy(), orc.sgbm_compute(l, r, orc.make_params(*args)))
    st = dict(m.stage_times())
    # cost volume + top-down sweep run as one kernel where the configuration is covered, else as two
    assert ("cost_path_v_fused" in st) or ("cost" in st and "path_v_fused_down" in st), st
    # one matcher alive on the device: both horizontal paths in one launch; several: two launches
    assert ("path_h_both_wta" in st) or ("path_h_fwd" in st and "path_h_bwd_wta" in st), st
    for k in ("prefilter", "lrcheck", "median3", "speckle"):
        assert k in st and st[k] > 0, k
    assert c.kernel_launches - n0 >= 6
    assert e0.elapsed_time(e1) >= 0.9 * sum(st.values())
    m.close(); c.close()


def test_pitched_device_buffers_and_handle_reuse(ctx, orc):
    """Row-padded (non-contiguous) device inputs and outputs; one handle reused for growing and shrinking sizes."""
    import torch
    from stereo_depth_ruler_b200 import capi, synth
    args = (0, 48, 5, 600, 2400, 1, 63, 12, 50, 2, 1)
    m = capi.Sgbm(ctx, capi.make_params(*args))
    for (W, H) in ((200, 60), (333, 97), (150, 31)):
        l, r = synth.stereo_pair(W, H, 48, W)
        Lb = torch.zeros((H, W + 37), dtype=torch.uint8, device="cuda"); Lb[:, :W] = torch.from_numpy(l).cuda()
        Rb = torch.zeros((H, W + 5), dtype=torch.uint8, device="cuda"); Rb[:, :W] = torch.from_numpy(r).cuda()
        Ob = torch.full((H, W + 11), 7777, dtype=torch.int16, device="cuda")
        m.compute(Lb[:, :W], Rb[:, :W], Ob[:, :W])
        ctx.synchronize()
        got = Ob.cpu().numpy()
        assert np.array_equal(got[:, :W], orc.sgbm_compute(l, r, orc.make_params(*args))), (W, H)
        assert (got[:, W:] == 7777).all()              # nothing written outside the view
    m.close()


def test_reproject_and_gray_on_pitched_views(ctx, orc, golden):
    import torch
    from stereo_depth_ruler_b200 import capi
    d16 = golden["reproj_disp16"]
    H, W = d16.shape
    Db = torch.zeros((H, W + 3), dtype=torch.int16, device="cuda"); Db[:, :W] = torch.from_numpy(d16).cuda()
    pts = capi.reproject_to_3d(ctx, Db[:, :W], golden["Q"], True)
    ctx.synchronize()
    assert np.array_equal(pts.cpu().numpy().view(np.uint32), golden["reproj_true"])
    f32 = capi.disp_to_float(ctx, Db[:, :W])
    ctx.synchronize()
    assert np.array_equal(f32.cpu().numpy(), d16.astype(np.float32) / 16)
    bgr = golden["gray_in"]
    Bb = torch.zeros((bgr.shape[0], bgr.shape[1] + 2, 3), dtype=torch.uint8, device="cuda"); Bb[:, :bgr.shape[1]] = torch.from_numpy(bgr).cuda()
    g = capi.bgr2gray(ctx, Bb[:, :bgr.shape[1]])
    ctx.synchronize()
    assert np.array_equal(g.cpu().numpy(), golden["gray_out"])


def test_large_batch_spans_several_cooperative_launches(ctx, orc):
    """More frames than one cooperative sweep launch can hold: the batch is split internally, results unchanged."""
    import torch
    from stereo_depth_ruler_b200 import capi, synth
    W, H, D, B = 1280, 40, 128, 20                    # 36 strips per frame: at most 8 frames per launch
    args = (0, D, 5, 600, 2400, 1, 63, 12, 0, 0, 0)
    pairs = [synth.stereo_pair(W, H, D, s) for s in range(3)]
    L = torch.from_numpy(np.stack([pairs[i % 3][0] for i in range(B)])).cuda()
    R = torch.from_numpy(np.stack([pairs[i % 3][1] for i in range(B)])).cuda()
    O = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
    m = capi.Sgbm(ctx, capi.make_params(*args))
    m.compute_batch(L, R, O)
    ctx.synchronize()
    out = O.cpu().numpy()
    refs = [orc.sgbm_compute(p[0], p[1], orc.make_params(*args)) for p in pairs]
    for i in range(B):
        assert np.array_equal(out[i], refs[i % 3]), i
    m.close()
