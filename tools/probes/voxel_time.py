"""Per-stage times of the pipeline (rectify / sgbm / voxelgrid) for one handle, batch 8: python tools/probes/voxel_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from stereo_depth_ruler_b200 import capi
ctx = capi.Context(0)
calib = capi.load_calibration(bench.CALIB)
p = capi.Pipeline(ctx, calib, capi.make_params(*bench.SGBM_ARGS), capi.voxel_options(5.0, drop_invalid=True, z_max=5000.0), handle_missing=True, batch=8)
frames = torch.from_numpy(np.stack(bench.make_frames(8))).cuda()
p.set_profiling(True)
acc = {}
for rep in range(6):
    p.run([frames[i] for i in range(8)])
    if rep:
        for k, v in p.stage_times():
            acc[k] = acc.get(k, 0) + v / 5
print({k: round(v, 4) for k, v in acc.items()}, [int(p.views[i].n) for i in range(8)])
