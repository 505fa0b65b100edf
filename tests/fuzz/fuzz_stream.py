"""Randomised test of the native frame-parallel runner (sdr_stream_*: worker threads per (device, handle) pulling batches of
frame indices from a shared cursor) against one ordinary pipeline handle: random frame counts (not multiples of the batch), pool
sizes (wrap-around), handles, batch sizes, input formats; every frame exactly once, every map and cloud equal to the pipeline's.
usage: python tests/fuzz/fuzz_stream.py [cases] [seed]. Test infrastructure."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from stereo_depth_ruler_b200 import capi

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ctx = capi.Context(0)
CAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "golden", "zed2_calibration.yaml")
ngpu = torch.cuda.device_count()
fails = 0
t0 = time.time()


def scaled_calib(W, H):
    c = capi.load_calibration(CAL)
    sx, sy = 1280.0 / W, 720.0 / H
    c.image_width, c.image_height = W, H
    for name in ("K_left", "K_right"):
        a = getattr(c, name)
        a[0] /= sx; a[2] /= sx; a[4] /= sy; a[5] /= sy
    for name in ("P1", "P2"):
        a = getattr(c, name)
        a[0] /= sx; a[2] /= sx; a[3] /= sx; a[5] /= sy; a[6] /= sy
    q = c.Q
    q[3] /= sx; q[7] /= sy; q[11] /= sx
    return c


for case in range(N):
    W, H = int(rng.integers(60, 160)) * 2, int(rng.integers(10, 40)) * 2
    D = 16 * int(rng.choice([1, 2, 4]))
    mode = int(rng.integers(0, 3))
    args = (0, D, 5, 600, 2400, 1, 63, 12, int(rng.choice([0, 100])), 2, mode)
    B = int(rng.choice([1, 2, 3, 4]))
    handles = int(rng.choice([1, 2, 3]))
    P = int(rng.integers(1, 7))
    n_frames = int(rng.integers(1, 23))
    clouds = bool(rng.integers(0, 2))
    c = scaled_calib(W, H)
    vox = capi.voxel_options(float(rng.choice([5.0, 20.0])), drop_invalid=True, z_max=5000.0)
    base = rng.integers(0, 256, size=(H, W + D, 3)).astype(np.uint8)
    base = ((base.astype(np.int32) + np.roll(base, 1, 1) + np.roll(base, 1, 0)) // 3).astype(np.uint8)
    pool = []
    for i in range(P):
        sh = int(rng.integers(0, D))
        b2 = np.roll(base, i * 5, 0)
        pool.append(np.ascontiguousarray(np.concatenate([b2[:, sh:sh + W], b2[:, :W]], axis=1)))
    tag = "case %d %dx%d D %d mode %d B %d handles %d pool %d frames %d clouds %d" % (case, W, H, D, mode, B, handles, P, n_frames, clouds)
    try:
        p = capi.Pipeline(ctx, c, capi.make_params(*args), vox, handle_missing=True, batch=1)
        expect = []
        for f in pool:
            d = [np.empty((H, W), np.int16)]
            p.run([f], disp=d)
            cl = p.cloud_to_host(0)
            expect.append((d[0].copy(), cl["xyz"].copy(), cl["rgb"].copy()))
        p.close()
        devices = [int(i % ngpu) for i in range(min(ngpu, 2))]
        s = capi.Stream(devices, handles, c, capi.make_params(*args), vox, handle_missing=True, batch=B, download_clouds=clouds)
        seen, bad, lock = {}, [], threading.Lock()

        def on_result(i, dev, disp, ncloud):
            with lock:
                seen[i] = seen.get(i, 0) + 1
                e = expect[i % P]
                if not np.array_equal(disp, e[0]):
                    bad.append("frame %d map" % i)
                if ncloud != len(e[1]):
                    bad.append("frame %d cloud size %d vs %d" % (i, ncloud, len(e[1])))

        def on_cloud(i, dev, xyz, rgb):
            with lock:
                e = expect[i % P]
                if not (np.array_equal(xyz.view(np.uint32), e[1].view(np.uint32)) and np.array_equal(rgb, e[2])):
                    bad.append("frame %d cloud" % i)

        for rep in range(2):          # the runner is persistent: a second call on the same workers
            seen.clear()
            st = s.process(n_frames, pool, on_result=on_result, on_cloud=on_cloud if clouds else None)
            if st.frames != n_frames or sorted(seen) != list(range(n_frames)) or any(v != 1 for v in seen.values()):
                bad.append("rep %d: frames %d, seen %s" % (rep, st.frames, sorted(seen.items())[:6]))
        # the pool entry point (what bench.py uses): pinned host pool and device-resident pool
        arr = np.stack(pool)
        st = s.process_pool(n_frames, arr)
        if st.frames != n_frames:
            bad.append("process_pool(host) frames %d" % st.frames)
        st = s.process_pool(n_frames, torch.from_numpy(arr).cuda())
        if st.frames != n_frames:
            bad.append("process_pool(device) frames %d" % st.frames)
        s.close()
        if bad:
            fails += 1
            print("MISMATCH", tag, bad[:4])
    except Exception as e:
        fails += 1
        print("EXC", tag, repr(e)[:200])
print("fuzz_stream: %d cases, seed %d, %d GPUs, %d failures, %.0f s" % (N, seed, ngpu, fails, time.time() - t0))
sys.exit(1 if fails else 0)
