"""CPU (gloo, world_size 2): the frame-parallel sharding and the max-over-ranks timing reduction bench.py uses for
N > 1. Frames share no state, so the N > 1 path is a partition of the frame stream plus one MAX all-reduce."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out):
    from stereo_depth_ruler_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.frames_for_rank(n_frames, rank, world)
    # every rank "processes" its frames: checksum = sum of frame ids; elapsed differs per rank
    local = torch.tensor([float(sum(mine)), float(len(mine))], dtype=torch.float64)
    dist.all_reduce(local, op=dist.ReduceOp.SUM)
    ms = sharding.max_over_ranks(10.0 + 5.0 * rank, dist, device="cpu")
    agg = sharding.aggregate_throughput(frames_per_rank=len(mine), world=world, ms=ms)
    if rank == 0:
        out.put((local.tolist(), ms, agg, mine))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_partition_is_exact_cover():
    from stereo_depth_ruler_b200 import sharding
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            parts = [sharding.frames_for_rank(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            for r, p in enumerate(parts):
                assert all(i % world == r for i in p)      # frame i -> rank i mod N


def test_two_rank_gloo_roundtrip():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 101, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (total, count), ms, agg, mine = res
    assert total == sum(range(101)) and count == 101        # no frame lost or duplicated across ranks
    assert ms == 15.0                                        # MAX over ranks, not rank 0's own clock
    assert mine == list(range(0, 101, 2))
    assert agg == pytest.approx(2 * 51 / 0.015)


# ---- row-strip split of one frame (BASELINE config 5): the host-side plan, no GPU ----
def test_stripe_plan_covers_every_stripe_once():
    from stereo_depth_ruler_b200 import capi, strip_split
    for n_stripes in (1, 4):
        for world in (1, 2, 3, 4, 8):
            plan = strip_split.assign_stripes(n_stripes, world)
            assert len(plan) == world
            covered = [s for first, cnt in plan for s in range(first, first + cnt)]
            assert covered == list(range(n_stripes))                       # contiguous blocks, in order, nothing twice
            assert max(c for _, c in plan) - min(c for _, c in plan) <= 1
    # the geometry the plan is applied to needs no device either (sdr_sgbm_stripe_rows): OpenCV's rule at 4K
    p = capi.make_params(0, 256, 5, 600, 2400, 1, 63, 12, 200, 2, capi.MODE_SGBM_3WAY)
    assert capi.stripe_rows(p, 3840, 2160) == [(0, 0, 540), (483, 540, 1080), (1023, 1080, 1620), (1563, 1620, 2160)]


def _strip_worker(rank, world, port, q):
    """Every rank 'matches' its stripes (fills their output rows with a rank-dependent pattern), rank 0 gathers the raw
    strips the way strip_split.compute does (send / recv of the row blocks as bytes) and checks the assembled map."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from stereo_depth_ruler_b200 import capi, strip_split
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W = 64, 48
    stripes = capi.stripe_rows(capi.make_params(0, 16, 5, 600, 2400, 1, 63, 12, 200, 2, capi.MODE_SGBM_3WAY), W, H)
    plan = strip_split.assign_stripes(len(stripes), world)
    raw = torch.full((H, W), -1, dtype=torch.int16)
    first, cnt = plan[rank]
    for s in range(first, first + cnt):
        raw[stripes[s][1]:stripes[s][2]] = 100 * rank + s
    for r in range(1, world):
        f, c = plan[r]
        if not c:
            continue
        y0, y1 = stripes[f][1], stripes[f + c - 1][2]
        if rank == r:
            dist.send(raw[y0:y1].view(torch.uint8), dst=0)
        elif rank == 0:
            dist.recv(raw[y0:y1].view(torch.uint8), src=r)
    if rank == 0:
        owner = {s: r for r, (f, c) in enumerate(plan) for s in range(f, f + c)}
        want = np.full((H, W), -1, np.int16)
        for s, (_, d0, e) in enumerate(stripes):
            want[d0:e] = 100 * owner[s] + s
        q.put(bool(np.array_equal(raw.numpy(), want)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_strip_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_strip_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
