"""GPU: row-strip split of one frame (BASELINE config 5). Stripes are matched separately, gathered, post-filtered once;
the result must equal the single-call result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stripe_geometry_matches_opencv_rule():
    from stereo_depth_ruler_b200 import capi
    p = capi.make_params(0, 256, 5, 600, 2400, 1, 63, 12, 200, 2, capi.MODE_SGBM_3WAY)
    rows = capi.stripe_rows(p, 3840, 2160)
    # stripe i reads source rows [max(i*540-57, 0), (i+1)*540) and writes rows [i*540, (i+1)*540)
    assert rows == [(0, 0, 540), (483, 540, 1080), (1023, 1080, 1620), (1563, 1620, 2160)]
    assert capi.stripe_rows(capi.make_params(mode=capi.MODE_SGBM), 1280, 720) == [(0, 0, 720)]


def test_stripes_on_separate_handles_equal_single_call(ctx, orc):
    """Four handles (stand-ins for four GPUs) match one stripe each into one raw map."""
    import torch
    from stereo_depth_ruler_b200 import capi, synth
    W, H, D = 640, 203, 64          # 203 rows: uneven last stripe
    args = (0, D, 5, 600, 2400, 1, 63, 12, 100, 2, capi.MODE_SGBM_3WAY)
    l, r = synth.stereo_pair(W, H, D, 4)
    lt, rt = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    whole = capi.Sgbm(ctx, capi.make_params(*args))
    ref = whole.compute(l, r)
    assert np.array_equal(ref, orc.sgbm_compute(l, r, orc.make_params(*args)))
    raw = torch.full((H, W), -1234, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()          # torch's stream and the handles' own streams are unordered
    stripes = capi.stripe_rows(whole.params, W, H)
    assert len(stripes) == 4
    for i in range(4):
        h = capi.Sgbm(ctx, capi.make_params(*args))
        h.match_stripes(lt, rt, i, 1, raw)
        ctx.synchronize()
        got = raw.cpu().numpy()
        assert (got[stripes[i][2]:] == -1234).all()          # rows of later stripes untouched
        h.close()
    assert np.array_equal(raw.cpu().numpy(), whole.debug_read(2, (H, W)))
    out = whole.postfilter(raw)
    ctx.synchronize()                 # device outputs are asynchronous on the handle's stream
    assert np.array_equal(out.cpu().numpy(), ref)
    # two stripes per call (the 2-GPU split)
    raw2 = torch.empty_like(raw)
    whole.match_stripes(lt, rt, 0, 2, raw2); whole.match_stripes(lt, rt, 2, 2, raw2)
    out2 = whole.postfilter(raw2)
    ctx.synchronize()
    assert np.array_equal(out2.cpu().numpy(), ref)
    with pytest.raises(capi.SdrError):
        whole.match_stripes(lt, rt, 4, 1, raw2)
    whole.close()


@pytest.mark.parametrize("devices", [[0, 0, 0, 0], [0, 0], [0, 0, 0], [0] * 6])
def test_native_split_parts_on_one_device(ctx, orc, devices):
    """sdr_sgbm_split_* with every part on device 0 (the code path of the multi-GPU split: per-part contexts and streams,
    partial row uploads, strips pushed into part 0's map, events, one post-filter): host and device-resident inputs,
    uneven stripes, more parts than stripes; equals the single call and the oracle bit for bit."""
    import torch
    from stereo_depth_ruler_b200 import capi, synth
    W, H, D = 640, 203, 64
    args = (0, D, 5, 600, 2400, 1, 63, 12, 100, 2, capi.MODE_SGBM_3WAY)
    l, r = synth.stereo_pair(W, H, D, 4)
    ref = orc.sgbm_compute(l, r, orc.make_params(*args))
    sp = capi.SgbmSplit(devices, capi.make_params(*args))
    assert np.array_equal(sp.compute(l, r), ref)                       # host frames in, host map out
    st = sp.stats()
    assert st["n_stripes"] == 4 and sum(st["stripes"]) == 4 and st["ms"] > 0
    assert all(rows < H for rows in st["input_rows"][:4] if len(devices) >= 4)   # nobody needs the whole frame
    lt, rt = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    torch.cuda.synchronize()
    out = sp.compute(lt, rt)                                            # device-resident: asynchronous
    sp.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    l2, r2 = synth.stereo_pair(W, H, D, 5)                              # handles are reusable
    assert np.array_equal(sp.compute(l2, r2), orc.sgbm_compute(l2, r2, orc.make_params(*args)))
    sp.close()
    # a single-stripe mode degenerates to part 0 doing everything
    args0 = (0, D, 5, 600, 2400, 1, 63, 12, 100, 2, capi.MODE_SGBM)
    sp0 = capi.SgbmSplit(devices, capi.make_params(*args0))
    assert np.array_equal(sp0.compute(l, r), orc.sgbm_compute(l, r, orc.make_params(*args0)))
    assert sp0.stats()["stripes"][0] == 1 and sum(sp0.stats()["stripes"]) == 1
    sp0.close()


def test_native_split_across_gpus_4k():
    """config 5 (3840x2160, D=256, MODE_SGBM_3WAY) split natively over 2 and 4 GPUs of the box (needs >= 2): bit-exact against
    cv2 itself, host and device-resident inputs; the stats name the halo / gather bytes that crossed NVLink."""
    import cv2
    import torch
    from stereo_depth_ruler_b200 import capi, synth
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    W, H, D = 3840, 2160, 256
    args = (0, D, 5, 600, 2400, 1, 63, 12, 200, 2, capi.MODE_SGBM_3WAY)
    l, r = synth.stereo_pair(W, H, D, 6)
    ref = cv2.StereoSGBM_create(*args[:10], mode=cv2.STEREO_SGBM_MODE_SGBM_3WAY).compute(l, r)
    for world in [w for w in (2, 4) if w <= n]:
        sp = capi.SgbmSplit(list(range(world)), capi.make_params(*args))
        assert np.array_equal(sp.compute(l, r), ref), "host input, %d GPUs" % world
        torch.cuda.set_device(0)
        lt, rt = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
        torch.cuda.synchronize()
        out = sp.compute(lt, rt)
        sp.synchronize()
        assert np.array_equal(out.cpu().numpy(), ref), "device input, %d GPUs" % world
        st = sp.stats()
        print("native split over %d GPUs: %s" % (world, st))
        assert all(st["peer_access"]) and all(b > 0 for b in st["gather_bytes"][1:]) and all(b > 0 for b in st["input_bytes"][1:])
        sp.close()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:  # noqa: BLE001 -- report instead of leaving the parent waiting on the queue
        if rank == 0:
            q.put("rank 0 failed: %r" % (e,))
        raise


def _worker_body(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from stereo_depth_ruler_b200 import capi, synth, strip_split
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    W, H, D = 3840, 2160, 256
    args = (0, D, 5, 600, 2400, 1, 63, 12, 200, 2, capi.MODE_SGBM_3WAY)
    l, r = synth.stereo_pair(W, H, D, 6)
    ctx = capi.Context(rank)
    m = capi.Sgbm(ctx, capi.make_params(*args))
    lt, rt = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    out = strip_split.compute(m, lt, rt, rank, world, dist)
    if rank == 0:
        ref = m.compute(l, r)
        q.put(bool(np.array_equal(out.cpu().numpy(), ref)))
    dist.barrier()
    dist.destroy_process_group()


def test_strip_split_across_gpus_4k():
    """config 5 (3840x2160, D=256, 3-way) split over every GPU of the box (needs >= 2)."""
    import torch
    import torch.multiprocessing as mp
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = _free_port()
    procs = [mpctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    assert ok is True, ok
