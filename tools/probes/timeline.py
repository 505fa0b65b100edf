"""Who runs when: absolute stage / kernel timelines of several pipeline handles working side by side on one GPU
(sdr_pipeline_stage_timeline). python tools/probes/timeline.py [handles] [iterations] [name=value sgbm options ...]"""
import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from stereo_depth_ruler_b200 import capi

NH = int(sys.argv[1]) if len(sys.argv) > 1 else 4
IT = int(sys.argv[2]) if len(sys.argv) > 2 else 6
opts = [kv.split("=") for kv in sys.argv[3:]]
calib = capi.load_calibration(bench.CALIB)
frames = torch.from_numpy(np.stack(bench.make_frames(16))).cuda()
base = torch.cuda.Event(enable_timing=True)
pipes, ctxs = [], []
for h in range(NH):
    c = capi.Context(0)
    p = capi.Pipeline(c, calib, capi.make_params(*bench.SGBM_ARGS), capi.voxel_options(5.0, drop_invalid=True, z_max=5000.0), handle_missing=True, batch=8)
    if NH > 1:
        p.set_sgbm_option("shared_device", 1)
    for k, v in opts:
        p.set_sgbm_option(k, int(v))
    p.run([frames[i] for i in range(8)])      # allocations
    p.set_profiling(True)
    pipes.append(p); ctxs.append(c)
torch.cuda.synchronize()
base.record()
torch.cuda.synchronize()
rows = []
lock = threading.Lock()


def worker(h):
    for it in range(IT):
        pipes[h].run([frames[(it * 8 + i) % 16] for i in range(8)])
        a = pipes[h].stage_timeline(base.cuda_event, 0)
        b = pipes[h].stage_timeline(base.cuda_event, 1)
        with lock:
            for tl, lvl in ((a, "P"), (b, "K")):
                for (n0, t0), (n1, t1) in zip(tl[:-1], tl[1:]):
                    rows.append((t0, t1, h, it, lvl, n1))


th = [threading.Thread(target=worker, args=(h,)) for h in range(NH)]
[t.start() for t in th]; [t.join() for t in th]
tmax = max(r[1] for r in rows)
print("handles %d, %d iterations each: %.2f ms in all = %.3f ms per batch of 8 = %.0f frames/s" % (NH, IT, tmax, tmax / (NH * IT), NH * IT * 8 / tmax * 1e3))
# steady-state window: the middle iterations; kernel level
lo = min(r[0] for r in rows if r[3] == IT // 2 and r[4] == "K")
hi = max(r[1] for r in rows if r[3] == IT // 2 and r[4] == "K")
print("kernel-level timeline of the window [%.2f, %.2f] ms (start, end, duration, handle, iteration, kernel):" % (lo, hi))
for t0, t1, h, it, lvl, n in sorted(r for r in rows if r[4] == "K" and r[1] > lo and r[0] < hi):
    print("  %8.3f %8.3f  %6.3f  h%d it%d  %s" % (t0, t1, t1 - t0, h, it, n))
print("pipeline-level stages in the same window:")
for t0, t1, h, it, lvl, n in sorted(r for r in rows if r[4] == "P" and r[1] > lo and r[0] < hi):
    print("  %8.3f %8.3f  %6.3f  h%d it%d  %s" % (t0, t1, t1 - t0, h, it, n))
# per-kernel mean duration under contention
import collections
acc = collections.defaultdict(list)
for t0, t1, h, it, lvl, n in rows:
    if it >= 1:
        acc[(lvl, n)].append(t1 - t0)
print("mean duration under contention:", {k[1]: round(sum(v) / len(v), 3) for k, v in sorted(acc.items())})
