import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, cv2
from stereo_depth_ruler_b200 import capi, synth
from oracle import orc
ctx = capi.Context(0)
W, H = 1280, 720
for seed in (3, 2):
    l, r = synth.stereo_pair(W, H, 128, seed)
    ls = cv2.resize(l, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
    rs = cv2.resize(r, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
    for mode in (2, 0):
        args = (-79, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, mode)
        ref = cv2.StereoSGBM_create(*args).compute(rs, ls)
        o = orc.sgbm_compute(rs, ls, orc.make_params(*args))
        for opts in ({}, {"path_hh": 0}, {"path_hh": 1, "h_lanes": 8}, {"path_hh": 1, "h_lanes": 16}, {"path_hh": 1, "h_lanes": 32}, {"path_hh": 0, "h_ws": 0}):
            m = capi.Sgbm(ctx, capi.make_params(*args))
            for k, v in opts.items(): m.set_option(k, v)
            g = m.compute(rs, ls)
            bad = np.argwhere(g != ref)
            print(seed, mode, opts, "gpu!=cv2:", len(bad), "oracle!=cv2:", int((o != ref).sum()), bad[:5].tolist(), [(int(g[y, x]), int(ref[y, x])) for y, x in bad[:5]], flush=True)
            m.close()
