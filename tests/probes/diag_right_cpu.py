import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, cv2
from stereo_depth_ruler_b200 import synth
from oracle import orc
W, H = 1280, 720
l, r = synth.stereo_pair(W, H, 128, 2)
ls = cv2.resize(l, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
rs = cv2.resize(r, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
def run(args, a, b):
    ref = cv2.StereoSGBM_create(*args).compute(a, b)
    o = orc.sgbm_compute(a, b, orc.make_params(*args))
    bad = np.argwhere(o != ref)
    print(args, "diff", len(bad), bad[:12].tolist(), [(int(o[y,x]), int(ref[y,x])) for y,x in bad[:12]])
    return ref, o
run((-79, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, 2), rs, ls)
run((-79, 80, 5, 600, 2400, 1000000, 63, 10, 0, 0, 2), rs, ls)
run((-79, 80, 5, 600, 2400, 1, 63, 0, 0, 0, 2), rs, ls)
run((-80, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, 2), rs, ls)
run((-64, 64, 5, 600, 2400, 1000000, 63, 0, 0, 0, 2), rs, ls)
run((-79, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, 2), rs[240:300], ls[240:300])
run((-79, 80, 5, 60, 240, 1000000, 63, 0, 0, 0, 2), rs, ls)
run((-79, 80, 3, 600, 2400, 1000000, 63, 0, 0, 0, 2), rs, ls)
print("--- local")
A = (-79, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, 2)
a, b = rs[255:275], ls[255:275]
run(A, a, b)
run(A, a // 2, b // 2)
run((-79, 80, 5, 600, 2400, 1000000, 31, 0, 0, 0, 2), a, b)
run(A, a[:, 200:400].copy(), b[:, 200:400].copy())
run((0, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, 2), b, a)
run((-79, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, 0), a, b)
run((-79, 80, 5, 600, 2400, 1000000, 63, 1, 0, 0, 2), a, b)
print(a[8:16, 275:295]); print(b[8:16, 200:300].max(), a[8:16, 275:295].max())
