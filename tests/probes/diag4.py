import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests/probes")
import numpy as np, cv2
from stereo_depth_ruler_b200 import synth
from oracle import orc
import importlib.util
spec = importlib.util.spec_from_file_location("d3", "/root/repo/tests/probes/diag3.py")
src = open("/root/repo/tests/probes/diag3.py").read().split("tot = {}")[0]
exec(src)
tot = {}
for seed in range(1, 7):
    l, r = synth.stereo_pair(1280, 720, 128, seed)
    ls = cv2.resize(l, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
    rs = cv2.resize(r, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
    for mode in (0, 1):
        for (a, b, minD, D) in ((rs, ls, -79, 80), (ls, rs, 0, 64)):
            A = (minD, D, 5, 600, 2400, 1000000, 63, 0, 0, 0, mode)
            ref = cv2.StereoSGBM_create(*A).compute(a, b)
            raw, _, Sv, _ = orc.sgbm_match(a, b, orc.make_params(*A), want_volumes=True)
            S = Sv.astype(np.int32); mn = S.min(axis=2, keepdims=True); nt = int(((S == mn).sum(axis=2) > 1).sum())
            for rule in ("first", (8, True)):
                o = orc.median3(wta(Sv, minD, a.shape[1], rule))
                tot[(mode, str(rule))] = tot.get((mode, str(rule)), 0) + int((o != ref).sum())
            print(seed, mode, minD, D, "tied pixels", nt, tot, flush=True)
