"""GPU parity: the CUDA SGBM path (through the C ABI) against the C oracle and the cv2 golden vectors.
Bit-exact CV_16S is the bar (integer work)."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_golden_vectors(ctx, golden):
    from stereo_depth_ruler_b200 import capi
    for name in golden["sgbm_cases"]:
        args = [int(v) for v in golden[name + "_args"]]
        m = capi.Sgbm(ctx, capi.make_params(*args))
        got = m.compute(golden[name + "_left"], golden[name + "_right"])
        assert np.array_equal(got, golden[name + "_disp"]), name
        m.close()


SWEEP = []
for mode in (0, 1, 2):
    SWEEP += [
        (96, 48, (0, 16, 1, 8, 32, 1, 0, 10, 0, 0, mode), "tex"),
        (130, 40, (0, 32, 11, 2904, 11616, 1, 31, 1, 20, 1, mode), "tex"),
        (100, 33, (0, 64, 5, 600, 2400, 2, 63, 30, 200, 2, mode), "noise"),
        (90, 17, (4, 16, 3, 0, 0, 0, 15, -1, 10, 4, mode), "low"),
        (67, 20, (0, 64, 5, 600, 2400, 1, 63, 12, 0, 0, mode), "tex"),
        (333, 61, (0, 112, 7, 1176, 4704, 1, 63, 12, 100, 2, mode), "tex"),
        (300, 37, (-16, 48, 5, 600, 2400, 1, 63, 12, 200, 2, mode), "tex"),
        (420, 50, (0, 256, 5, 600, 2400, 1, 63, 12, 200, 2, mode), "tex"),
        (131, 3, (0, 16, 5, 600, 2400, 1, 63, 12, 0, 0, mode), "noise"),   # fewer rows than the window
        (50, 1, (0, 16, 3, 216, 864, 1, 63, 0, 0, 0, mode), "noise"),      # single row
        # every D / 64 of the latency-mode chain kernel (k_path_hh_ws: D / 32 values per lane; D = 192 is the odd register count
        # that tests/fuzz/fuzz_sgbm.py found broken) and D that are no multiple of 32 or 64
        (400, 32, (0, 192, 5, 200, 800, 1, 63, 12, 0, 0, mode), "tex"),
        (289, 32, (-16, 192, 5, 200, 800, 5, 63, 0, 0, 0, mode), "tex"),
        (300, 21, (0, 96, 5, 600, 2400, 1, 63, 12, 50, 2, mode), "tex"),
        (330, 18, (0, 160, 3, 72, 288, 1, 63, 10, 0, 0, mode), "tex"),
        (380, 12, (0, 240, 5, 600, 2400, 1, 63, 12, 0, 0, mode), "noise"),
        # saturated cost sums (large window and P2): a pixel whose S is SHRT_MAX at every disparity keeps cv2's bestDisp = -1
        # (MODE_SGBM / MODE_HH: invalid) and never enters disp2 (found by tests/fuzz/fuzz_sgbm.py)
        (196, 16, (16, 128, 9, 0, 7776, 1, 63, 0, 100, 4, mode), "tex"),
        (196, 16, (0, 128, 9, 0, 7776, 1, 63, 0, 0, 0, mode), "noise"),
        # the 64-column cost tile does not fit shared memory (233 KB): 32-column tiles
        (409, 16, (4, 256, 11, 101, 2012, 1, 63, 15, 200, 1, mode), "tex"),
    ]


@pytest.mark.parametrize("W,H,args,kind", SWEEP)
def test_against_oracle(ctx, orc, W, H, args, kind):
    from stereo_depth_ruler_b200 import capi, synth
    if kind == "tex":
        l, r = synth.stereo_pair(W, H, args[1], 3)
    else:
        l, r = synth.noise_pair(W, H, 5, 256 if kind == "noise" else 8)
    ref = orc.sgbm_compute(l, r, orc.make_params(*args))
    m = capi.Sgbm(ctx, capi.make_params(*args))
    got = m.compute(l, r)
    assert np.array_equal(ref, got)
    # raw matcher output (before median / speckle) as well
    assert np.array_equal(m.debug_read(2, (H, W)), orc.sgbm_match(l, r, orc.make_params(*args)))
    m.close()


def test_cost_and_aggregate_volumes(ctx, orc):
    """Stage-level parity: C volume and the 4-path aggregate S of MODE_SGBM."""
    from stereo_depth_ruler_b200 import capi, synth
    W, H, D = 200, 60, 48
    l, r = synth.stereo_pair(W, H, D, 9)
    po = orc.make_params(0, D, 5, 600, 2400, 1, 63, 12, 0, 0, 0)
    raw, Cv, Sv, Sp = orc.sgbm_match(l, r, po, want_volumes=True)
    m = capi.Sgbm(ctx, capi.make_params(0, D, 5, 600, 2400, 1, 63, 12, 0, 0, 0))
    m.set_option("path_hh", 0)   # the separate L->R kernel leaves the 4-path aggregate in S
    m.compute(l, r)
    W1 = W - D
    assert np.array_equal(m.debug_read(0, (H, W1, D)).astype(np.int32) + 2400, Cv)
    assert np.array_equal(m.debug_read(1, (H, W1, D)), Sp)
    m.close()


def test_headline_config_bit_exact_vs_cv2_golden(ctx, golden):
    """1280x720x128 MODE_SGBM with the reference tuple: the map cv2 4.13.0 produced (sha256 + crop)."""
    from stereo_depth_ruler_b200 import capi, synth
    l, r = synth.stereo_pair(1280, 720, 128, 0)
    assert sha(l) + sha(r) == str(golden["full_c1_inputs_sha"])
    m = capi.Sgbm(ctx, capi.make_params())
    got = m.compute(l, r)
    assert np.array_equal(got[300:364, 600:728], golden["full_c1_crop"])
    assert int((got >= 0).sum()) == int(golden["full_c1_valid"])
    assert sha(got) == str(golden["full_c1_sha"])
    # idempotence of the handle: a second call on the same handle gives the same bytes
    assert sha(m.compute(l, r)) == str(golden["full_c1_sha"])
    m.close()


def test_reference_tuple_3way_full_size(ctx, golden):
    """The reference's own parameter tuple (D=80, MODE_SGBM_3WAY) at 1280x720 vs the cv2 sha."""
    from stereo_depth_ruler_b200 import capi, synth
    l, r = synth.stereo_pair(1280, 720, 80, 1)
    m = capi.Sgbm(ctx, capi.make_params(0, 80, 5, 600, 2400, 1, 63, 12, 200, 2, capi.MODE_SGBM_3WAY))
    assert sha(m.compute(l, r)) == str(golden["full_3way_sha"])
    m.close()


def test_batch_equals_single_and_device_resident(ctx, orc):
    """Frames share no state: a batch gives the per-frame results; torch CUDA tensors are used in place."""
    import torch
    from stereo_depth_ruler_b200 import capi, synth
    W, H, D, B = 256, 64, 64, 5
    pairs = [synth.stereo_pair(W, H, D, s) for s in range(B)]
    args = (0, D, 5, 600, 2400, 1, 63, 12, 100, 2, 0)
    m = capi.Sgbm(ctx, capi.make_params(*args))
    L = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda()
    R = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
    O = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
    m.compute_batch(L, R, O)
    ctx.synchronize()
    out = O.cpu().numpy()
    for i in range(B):
        assert np.array_equal(out[i], orc.sgbm_compute(pairs[i][0], pairs[i][1], orc.make_params(*args))), i
    # host batch path, ragged strides (row-padded inputs)
    padL = np.zeros((H, W + 13), np.uint8); padL[:, :W] = pairs[0][0]
    padR = np.zeros((H, W + 13), np.uint8); padR[:, :W] = pairs[0][1]
    got = m.compute(padL[:, :W], padR[:, :W])
    assert np.array_equal(got, out[0])
    m.close()


def test_error_behaviour(ctx):
    from stereo_depth_ruler_b200 import capi, synth
    for bad in (dict(numDisparities=100), dict(numDisparities=0), dict(blockSize=4), dict(mode=3), dict(numDisparities=512)):
        with pytest.raises(capi.SdrError) as e:
            capi.Sgbm(ctx, capi.make_params(**bad))
        assert e.value.status == capi.SDR_EINVAL
    m = capi.Sgbm(ctx, capi.make_params(numDisparities=64))
    l, r = synth.noise_pair(66, 20, 1)
    with pytest.raises(capi.SdrError) as e:      # cv2 raises for this input too
        m.compute(l, r)
    assert e.value.status == capi.SDR_EINVAL and "too small" in str(e.value)
    with pytest.raises(capi.SdrError) as e:
        m.compute(np.zeros((20, 100), np.uint8), np.zeros((21, 100), np.uint8))
    assert e.value.status == capi.SDR_ESIZE
    m.close()


def test_post_filters_standalone(ctx, orc, golden):
    from stereo_depth_ruler_b200 import capi
    assert np.array_equal(capi.median3_16s(ctx, golden["median_in"]), golden["median_out"])
    assert np.array_equal(capi.filter_speckles(ctx, golden["speckle_in"], -16, 200, 32), golden["speckle_out"])
    rng = np.random.default_rng(5)
    for size, diff, shape in ((200, 32, (300, 400)), (30, 16, (97, 131)), (5, 0, (64, 64)), (100000, 1000, (50, 70))):
        y = (rng.integers(0, 30, size=shape) * 8).astype(np.int16)
        y[rng.random(shape) < 0.25] = -16
        assert np.array_equal(capi.filter_speckles(ctx, y, -16, size, diff), orc.filter_speckles(y, -16, size, diff))
    x = rng.integers(-3000, 3000, size=(1, 9)).astype(np.int16)
    assert np.array_equal(capi.median3_16s(ctx, x), orc.median3(x))
    # speckle filtering is idempotent
    y1 = capi.filter_speckles(ctx, y, -16, 50, 16)
    assert np.array_equal(capi.filter_speckles(ctx, y1, -16, 50, 16), y1)


FUSED = [(300, 70, 128, 5, 0), (290, 45, 128, 3, 0), (333, 40, 128, 5, 1), (197, 50, 64, 3, 0), (161, 37, 64, 5, 1),
         (161, 2, 128, 5, 0), (164, 9, 64, 5, 0), (420, 44, 256, 5, 0), (389, 30, 256, 3, 1), (300, 3, 256, 5, 0)]


@pytest.mark.parametrize("W,H,D,bs,mode", FUSED)
def test_fused_cost_sweep_kernel(ctx, orc, W, H, D, bs, mode):
    """The configurations the fused cost + top-down sweep kernel covers (D = 64 / 128 / 256, blockSize 3 / 5, MODE_SGBM and the
    first pass of MODE_HH), ragged strip widths and fewer rows than the box window: bit-exact against the oracle incl.
    the C and S volumes, identical to the two-kernel path, and the stage timer shows which path ran."""
    from stereo_depth_ruler_b200 import capi, synth
    args = (0, D, bs, 8 * bs * bs, 32 * bs * bs, 1, 63, 12, 50, 2, mode)
    B = 3
    pairs = [synth.stereo_pair(W, H, D, 20 + s) for s in range(B)]
    po = orc.make_params(*args)
    refs = [orc.sgbm_compute(l, r, po) for l, r in pairs]
    outs = {}
    for fused in (True, False):
        m = capi.Sgbm(ctx, capi.make_params(*args))
        m.set_option("path_hh", 0)   # the separate L->R kernel leaves the 4-path aggregate in S
        m.set_option("cost_fusion", int(fused))
        m.set_profiling(True)
        got = m.compute_batch(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]), np.empty((B, H, W), np.int16))
        names = [k for k, _ in m.stage_times()]
        assert ("cost_path_v_fused" in names) == fused, names
        for i in range(B):
            assert np.array_equal(got[i], refs[i]), (fused, i)
        # the debug view reads frame 0 of the last call: run the first pair alone
        m.compute(*pairs[0])
        W1 = W - D
        outs[fused] = (m.debug_read(0, (H, W1, D)).copy(), m.debug_read(1, (H, W1, D)).copy())
        m.close()
    assert np.array_equal(outs[True][0], outs[False][0]) and np.array_equal(outs[True][1], outs[False][1])
    if mode == 0:
        raw, Cv, Sv, Sp = orc.sgbm_match(pairs[0][0], pairs[0][1], po, want_volumes=True)
        assert np.array_equal(outs[True][0].astype(np.int32) + args[4], Cv)
        assert np.array_equal(outs[True][1], Sp)


@pytest.mark.parametrize("cols,lanes", [(16, 8), (32, 8), (0, 16)])
@pytest.mark.parametrize("W,H,D,mode", [(210, 40, 80, 0), (331, 35, 256, 1), (180, 21, 64, 1), (150, 2, 48, 0), (300, 26, 256, 0)])
def test_sweep_kernel_strip_widths(ctx, orc, W, H, D, mode, cols, lanes):
    """k_path_v3 (the sweep without the cost fusion: D other than 64 / 128, and the bottom-up pass of MODE_HH) with 16- and
    32-column strips and, at D = 256, a column's costs over 8 or 16 lanes; ragged last strips: bit-exact against the oracle."""
    from stereo_depth_ruler_b200 import capi, synth
    args = (0, D, 5, 600, 2400, 1, 63, 12, 50, 2, mode)
    l, r = synth.stereo_pair(W, H, D, 9)
    m = capi.Sgbm(ctx, capi.make_params(*args))
    m.set_option("cost_fusion", 0)       # the sweep kernel on its own (D = 256 would otherwise take the fused cost + sweep kernel)
    m.set_option("v3_cols", cols)
    m.set_option("v3_lanes", lanes)
    m.set_profiling(True)
    got = m.compute(l, r)
    names = [k for k, _ in m.stage_times()]
    assert any(n.startswith("path_v_fused") for n in names), names
    assert np.array_equal(got, orc.sgbm_compute(l, r, orc.make_params(*args)))
    m.close()


@pytest.mark.parametrize("W,H,args", [(300, 41, (0, 128, 5, 600, 2400, 1, 63, 12, 50, 2, 0)), (203, 33, (0, 64, 3, 72, 288, 1, 63, 12, 0, 0, 1)),
                                      (190, 57, (0, 80, 5, 600, 2400, 1, 63, 12, 200, 2, 2)), (140, 9, (3, 48, 3, 72, 288, 2, 31, 5, 20, 1, 0)),
                                      (333, 27, (0, 256, 5, 600, 2400, 1, 63, 12, 200, 2, 1))])
def test_forward_pass_with_its_own_volume(ctx, orc, W, H, args):
    """What several matchers on one device run (option shared_device): the L->R pass writes its path costs to a volume of its
    own and the R->L + WTA kernel adds C-independent S and that volume (lr_split): bit-exact against the oracle and against the
    read-modify-write form, batch of 3."""
    from stereo_depth_ruler_b200 import capi, synth
    B = 3
    pairs = [synth.stereo_pair(W, H, args[1], 60 + s) for s in range(B)]
    refs = [orc.sgbm_compute(l, r, orc.make_params(*args)) for l, r in pairs]
    for split in (1, 0):
        m = capi.Sgbm(ctx, capi.make_params(*args))
        m.set_option("shared_device", 1)
        m.set_option("lr_split", split)
        m.set_profiling(True)
        got = m.compute_batch(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]), np.empty((B, H, W), np.int16))
        names = [k for k, _ in m.stage_times()]
        assert "path_h_fwd" in names and "path_h_bwd_wta" in names, names
        for i in range(B):
            assert np.array_equal(got[i], refs[i]), (split, i)
        m.close()


FUSED_3WAY = [(300, 70, 80, 5, 1), (333, 64, 128, 5, 2), (197, 50, 64, 3, 4), (161, 9, 80, 3, 0), (420, 38, 256, 5, 0), (229, 3, 64, 5, 2),
              (165, 61, 128, 3, 1)]


@pytest.mark.parametrize("W,H,D,bs,cm", FUSED_3WAY)
def test_fused_cost_vertical_kernel_3way(ctx, orc, W, H, D, bs, cm):
    """MODE_SGBM_3WAY (the reference's own mode): cost volume + the top-down path of every stripe in one kernel
    (k_cost_path_v1; D = 64 / 80 / 128 / 256, blockSize 3 / 5, 1 / 2 / 4 columns per consumer group, ragged strips, stripes
    shorter than the box window): bit-exact against the oracle, C and S volumes identical to k_cost + k_path_v."""
    from stereo_depth_ruler_b200 import capi, synth
    args = (0, D, bs, 8 * bs * bs, 32 * bs * bs, 1, 63, 12, 50, 2, 2)
    B = 2
    pairs = [synth.stereo_pair(W, H, D, 40 + s) for s in range(B)]
    po = orc.make_params(*args)
    refs = [orc.sgbm_compute(l, r, po) for l, r in pairs]
    outs = {}
    for fused in (True, False):
        m = capi.Sgbm(ctx, capi.make_params(*args))
        m.set_option("path_hh", 0)
        m.set_option("cost_fusion", int(fused))
        m.set_option("v1_cm", cm)
        m.set_profiling(True)
        got = m.compute_batch(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]), np.empty((B, H, W), np.int16))
        names = [k for k, _ in m.stage_times()]
        assert ("cost_path_v1_fused" in names) == fused and ("path_v" in names) == (not fused), names
        for i in range(B):
            assert np.array_equal(got[i], refs[i]), (fused, i)
        m.compute(*pairs[0])
        crows = sum(e - s0 for s0, _, e in capi.stripe_rows(capi.make_params(*args), W, H))
        outs[fused] = (m.debug_read(0, (crows, W - D, D)).copy(), m.debug_read(1, (H, W - D, D)).copy())
        m.close()
    assert np.array_equal(outs[True][0], outs[False][0]) and np.array_equal(outs[True][1], outs[False][1])


@pytest.mark.parametrize("W,H,args", [(300, 41, (0, 128, 5, 600, 2400, 1, 63, 12, 50, 2, 0)), (203, 33, (0, 64, 3, 72, 288, 1, 63, 12, 0, 0, 1)),
                                      (177, 50, (0, 80, 5, 600, 2400, 1, 63, 12, 200, 2, 2)), (97, 9, (-16, 48, 5, 600, 2400, 2, 63, 5, 0, 0, 0)),
                                      (67, 3, (0, 64, 5, 600, 2400, 1, 63, 12, 0, 0, 2)), (68, 5, (0, 64, 5, 600, 2400, 1, 63, 12, 0, 0, 0))])
def test_two_direction_horizontal_kernel(ctx, orc, W, H, args):
    """Option path_hh=1: both horizontal paths of a row in one launch, meeting in the middle (odd and even W1,
    W1 = 3 and 4, every mode): same CV_16S map and raw matcher output as the oracle."""
    from stereo_depth_ruler_b200 import capi, synth
    l, r = synth.stereo_pair(W, H, args[1], 31)
    po = orc.make_params(*args)
    m = capi.Sgbm(ctx, capi.make_params(*args)).set_option("path_hh", 1)
    m.set_profiling(True)
    got = m.compute(l, r)
    assert "path_h_both_wta" in [k for k, _ in m.stage_times()]
    assert np.array_equal(got, orc.sgbm_compute(l, r, po))
    assert np.array_equal(m.debug_read(2, (H, W)), orc.sgbm_match(l, r, po))
    B = 3
    pairs = [synth.stereo_pair(W, H, args[1], 40 + s) for s in range(B)]
    out = m.compute_batch(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]), np.empty((B, H, W), np.int16))
    for i in range(B):
        assert np.array_equal(out[i], orc.sgbm_compute(pairs[i][0], pairs[i][1], po)), i
    m.close()


@pytest.mark.parametrize("W,H,args,lanes", [(300, 41, (0, 128, 5, 600, 2400, 1, 63, 12, 50, 2, 0), 32), (300, 41, (0, 128, 5, 600, 2400, 1, 63, 12, 50, 2, 0), 16),
                                            (203, 33, (0, 64, 3, 72, 288, 1, 63, 12, 0, 0, 1), 32), (177, 50, (0, 96, 5, 600, 2400, 1, 63, 12, 200, 2, 2), 16),
                                            (420, 37, (0, 256, 5, 600, 2400, 1, 63, 12, 100, 2, 2), 32), (97, 9, (-16, 64, 5, 600, 2400, 2, 63, 5, 0, 0, 0), 32),
                                            (69, 4, (0, 64, 5, 600, 2400, 1, 63, 12, 0, 0, 0), 16), (330, 21, (0, 80, 5, 600, 2400, 1, 63, 12, 0, 0, 2), 32)])
def test_two_direction_kernel_wide_lanes(ctx, orc, W, H, args, lanes):
    """Option h_lanes = 16 / 32: a row's D costs over 16 or 32 lanes (what one frame or one stripe gets automatically; D = 80
    cannot and falls back to 8): same maps as the oracle, single frame and batch."""
    from stereo_depth_ruler_b200 import capi, synth
    l, r = synth.stereo_pair(W, H, args[1], 61)
    po = orc.make_params(*args)
    m = capi.Sgbm(ctx, capi.make_params(*args)).set_option("path_hh", 1).set_option("h_lanes", lanes)
    got = m.compute(l, r)
    assert np.array_equal(got, orc.sgbm_compute(l, r, po))
    assert np.array_equal(m.debug_read(2, (H, W)), orc.sgbm_match(l, r, po))
    pairs = [synth.stereo_pair(W, H, args[1], 70 + s) for s in range(2)]
    out = m.compute_batch(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]), np.empty((2, H, W), np.int16))
    for i in range(2):
        assert np.array_equal(out[i], orc.sgbm_compute(pairs[i][0], pairs[i][1], po)), i
    m.close()


@pytest.mark.parametrize("W,H,args", [(300, 41, (0, 128, 5, 600, 2400, 1, 63, 12, 50, 2, 0)), (177, 50, (0, 80, 5, 600, 2400, 1, 63, 12, 200, 2, 2)),
                                      (97, 9, (-16, 48, 5, 600, 2400, 2, 63, 5, 0, 0, 0)), (203, 33, (0, 64, 3, 72, 288, 1, 63, 12, 0, 0, 1)),
                                      (340, 19, (0, 256, 5, 600, 2400, 1, 63, 12, 0, 0, 0))])
def test_backward_wta_kernel(ctx, orc, W, H, args):
    """The separate R->L + WTA kernel (what runs with several matchers per device): raw matcher output equals the oracle."""
    from stereo_depth_ruler_b200 import capi, synth
    l, r = synth.stereo_pair(W, H, args[1], 33)
    po = orc.make_params(*args)
    m = capi.Sgbm(ctx, capi.make_params(*args)).set_option("path_hh", 0)
    m.set_profiling(True)
    got = m.compute(l, r)
    assert "path_h_bwd_wta" in [k for k, _ in m.stage_times()]
    assert np.array_equal(m.debug_read(2, (H, W)), orc.sgbm_match(l, r, po))
    assert np.array_equal(got, orc.sgbm_compute(l, r, po))
    m.close()


@pytest.mark.parametrize("opts", [{}, {"path_hh": 0}, {"path_hh": 1, "h_lanes": 16}, {"path_hh": 1, "h_lanes": 32}])
def test_three_way_without_uniqueness_takes_cv2s_tied_minimum(ctx, orc, opts):
    """MODE_SGBM_3WAY with uniquenessRatio 0 (the matcher createRightMatcher builds, stereo_disparity.cpp:10): where the
    minimum of S is tied 8 or more disparities apart cv2's 8-lane search does not return the first minimum. Half-size frames
    with such ties (seed 2 has them: the first-minimum rule differs from cv2 in 10 pixels) against cv2 itself."""
    cv2 = pytest.importorskip("cv2")
    from stereo_depth_ruler_b200 import capi, synth
    l, r = synth.stereo_pair(1280, 720, 128, 2)
    ls = cv2.resize(l, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
    rs = cv2.resize(r, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
    for a, b, args in ((rs, ls, (-79, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, 2)), (ls, rs, (0, 64, 5, 600, 2400, 1000000, 63, 0, 0, 0, 2))):
        m = capi.Sgbm(ctx, capi.make_params(*args))
        for k, v in opts.items():
            m.set_option(k, v)
        got = m.compute(a, b)
        ref = cv2.StereoSGBM_create(*args).compute(a, b)
        assert np.array_equal(got, ref), (args, int((got != ref).sum()))
        assert np.array_equal(got, orc.sgbm_compute(a, b, orc.make_params(*args)))
        m.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_per_direction_fallback_kernels(ctx, orc, mode):
    """Option fused_sweep=0: the one-launch-per-direction path kernels (the fallback when a cooperative grid does not
    fit on the device) give the same map."""
    from stereo_depth_ruler_b200 import capi, synth
    W, H, args = 230, 37, (0, 64, 5, 600, 2400, 1, 63, 12, 30, 2, mode)
    l, r = synth.stereo_pair(W, H, args[1], 17)
    m = capi.Sgbm(ctx, capi.make_params(*args)).set_option("fused_sweep", 0)
    m.set_profiling(True)
    got = m.compute(l, r)
    names = [k for k, _ in m.stage_times()]
    assert "path_v0" in names and "cost_path_v_fused" not in names, names
    assert np.array_equal(got, orc.sgbm_compute(l, r, orc.make_params(*args)))
    m.close()


@pytest.mark.parametrize("cols,D", [(64, 128), (16, 128), (16, 64), (32, 128), (8, 128), (8, 64)])
@pytest.mark.parametrize("W,H,bs", [(300, 40, 5), (333, 23, 3), (257, 31, 5)])
def test_fused_kernel_strip_widths_small(ctx, orc, W, H, bs, cols, D):
    """Option cf_cols: the fused kernel with 64-column strips (one CTA per SM; what a batch of 8 at 1280 wide runs), 32-column
    strips and 16-column strips (what a single frame runs: twice the CTAs) on ragged widths (last strip partial, a single
    strip + 1 column), against the oracle."""
    from stereo_depth_ruler_b200 import capi, synth
    args = (0, D, bs, 8 * bs * bs, 32 * bs * bs, 1, 63, 12, 50, 2, 0)
    pairs = [synth.stereo_pair(W, H, D, 50 + s) for s in range(2)]
    m = capi.Sgbm(ctx, capi.make_params(*args)).set_option("cf_cols", cols)
    m.set_profiling(True)
    out = m.compute_batch(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]), np.empty((2, H, W), np.int16))
    assert "cost_path_v_fused" in [k for k, _ in m.stage_times()]
    for i, (l, r) in enumerate(pairs):
        assert np.array_equal(out[i], orc.sgbm_compute(l, r, orc.make_params(*args))), i
    m.close()


def test_headline_batch_of_8_equals_single_frames(ctx, golden):
    """The benchmark's shape: 8 frames of 1280x720x128 in one call (fused kernel with 64-column strips, several handles'
    worth of work per launch) give, frame by frame, the maps cv2 itself computes for the same pairs and the maps of
    one-frame calls (32-column strips); frame 0 is also the committed cv2 golden map."""
    import torch
    from stereo_depth_ruler_b200 import capi, synth
    B = 8
    pairs = [synth.stereo_pair(1280, 720, 128, s) for s in range(B)]
    m = capi.Sgbm(ctx, capi.make_params())
    L = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda()
    R = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
    O = torch.empty((B, 720, 1280), dtype=torch.int16, device="cuda")
    m.compute_batch(L, R, O)
    ctx.synchronize()
    out = O.cpu().numpy()
    assert sha(out[0]) == str(golden["full_c1_sha"])
    import cv2
    ref = cv2.StereoSGBM_create(0, 128, 5, 600, 2400, 1, 63, 12, 200, 2, cv2.STEREO_SGBM_MODE_SGBM)
    for i in range(B):
        assert np.array_equal(out[i], ref.compute(*pairs[i])), "frame %d differs from cv2" % i
        assert np.array_equal(out[i], m.compute(*pairs[i])), i
    m.close()


@pytest.mark.parametrize("wave", [1, 2])
@pytest.mark.parametrize("W,H,args", [(200, 24, (0, 64, 5, 600, 2400, 1, 63, 12, 0, 0, 0)), (161, 9, (0, 64, 3, 72, 288, 1, 63, 12, 0, 0, 0)),
                                      (300, 40, (0, 128, 5, 600, 2400, 1, 63, 12, 50, 2, 0)), (257, 31, (0, 128, 3, 72, 288, 2, 63, 5, 0, 0, 0)),
                                      (400, 50, (-16, 64, 5, 600, 2400, 1, 63, 12, 0, 0, 0))])
def test_wavefront_sweeps(ctx, orc, W, H, args, wave):
    """Option wave: MODE_SGBM as two strip sweeps whose strips run one behind the other (k_wave_fwd: cost + -> \\ |,
    k_wave_bwd: / <- + WTA; wave = 2: the forward sweep followed by the per-direction kernels). Ragged strips (W1 not a
    multiple of 32), fewer rows than strips, D = 64 / 128, blockSize 3 / 5, a batch of two: raw matcher output == oracle."""
    from stereo_depth_ruler_b200 import capi, synth
    l, r = synth.stereo_pair(W, H, args[1], 21)
    po = orc.make_params(*args)
    m = capi.Sgbm(ctx, capi.make_params(*args)).set_option("wave", wave)
    m.set_profiling(True)
    got = m.compute(l, r)
    assert "wave_fwd" in [k for k, _ in m.stage_times()]
    assert np.array_equal(m.debug_read(2, (H, W)), orc.sgbm_match(l, r, po))
    assert np.array_equal(got, orc.sgbm_compute(l, r, po))
    pairs = [synth.stereo_pair(W, H, args[1], 80 + s) for s in range(2)]
    out = m.compute_batch(np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs]), np.empty((2, H, W), np.int16))
    for i in range(2):
        assert np.array_equal(out[i], orc.sgbm_compute(pairs[i][0], pairs[i][1], po)), i
    m.close()


def test_wavefront_sweeps_headline_equals_cv2(ctx, golden):
    """The headline shape through the wavefront sweeps: bit-exact against the cv2 fixture (sha256 of the CV_16S map)."""
    from stereo_depth_ruler_b200 import capi, synth
    l, r = synth.stereo_pair(1280, 720, 128, 0)
    m = capi.Sgbm(ctx, capi.make_params()).set_option("wave", 1)
    assert sha(m.compute(l, r)) == str(golden["full_c1_sha"])
    m.close()
