import sys
sys.path.insert(0, "/root/repo")
import numpy as np, cv2
from stereo_depth_ruler_b200 import synth
from oracle import orc
l, r = synth.stereo_pair(1280, 720, 128, 2)
ls = cv2.resize(l, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
rs = cv2.resize(r, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
A = (-79, 80, 5, 600, 2400, 1000000, 63, 0, 0, 0, 2)
a, b = rs[240:300], ls[240:300]
ref = cv2.StereoSGBM_create(*A).compute(a, b)
out = orc.sgbm_match(a, b, orc.make_params(*A), want_volumes=True)
print(type(out), [getattr(o, 'shape', None) for o in out] if isinstance(out, tuple) else out.keys())
raw, Cv, Sv, Sp = out
print("S filled:", int(np.abs(Sv).sum()), int(np.abs(Sp).sum()), int(np.abs(Cv).sum()))
med = orc.median3(raw)
print("median(raw)==ref except:", np.argwhere(med != ref).tolist())
np.set_printoptions(linewidth=200)
print(raw[44:49, 277:284]); print(ref[44:49, 277:284]); print(med[44:49, 277:284])
print(raw[25:30, 284:290]); print(ref[25:30, 284:290])
for (y, x) in ((46, 280), (26, 287), (27, 287), (27, 286)):
    xs = x - 1   # minX1 = 1
    s = Sv[y, xs].astype(int)
    m = s.min(); w = np.argwhere(s == m).ravel().tolist()
    print((y, x), "raw", raw[y, x], "minS", m, "at", w, "S near:", s[max(w[0]-2,0):w[-1]+3].tolist())
