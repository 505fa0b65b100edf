import sys
sys.path.insert(0, "/root/repo")
import numpy as np, cv2
from stereo_depth_ruler_b200 import synth
from oracle import orc

def wta(S, minD, W, rule):
    H, W1, D = S.shape
    S = S.astype(np.int32)
    minS = S.min(axis=2)
    ism = S == minS[..., None]
    idx = np.arange(D)
    if rule == "first":
        best = np.where(ism, idx, 10**6).min(axis=2)
    else:
        V, last = rule
        # per lane: lane min, position (first or last occurrence), then among lanes with global min the smallest position
        Sl = S.reshape(H, W1, D // V, V)
        lmin = Sl.min(axis=2)                                # [H,W1,V]
        il = idx.reshape(D // V, V)
        eq = Sl == lmin[:, :, None, :]
        pos = np.where(eq, il, -1).max(axis=2) if last else np.where(eq, il, 10**6).min(axis=2)
        best = np.where(lmin == minS[..., None], pos, 10**6).min(axis=2)
    d = best.copy()
    inner = (best > 0) & (best < D - 1)
    bm = np.clip(best - 1, 0, D - 1); bp = np.clip(best + 1, 0, D - 1)
    Sm = np.take_along_axis(S, bm[..., None], 2)[..., 0]; Sp = np.take_along_axis(S, bp[..., None], 2)[..., 0]
    den = np.maximum(Sm + Sp - 2 * minS, 1)
    num = (Sm - Sp) * 16 + den
    q = np.trunc(num / (2.0 * den)).astype(np.int64)   # C division truncates toward zero
    sub = np.where(inner, best * 16 + q, best * 16)
    out = np.full((H, W), (minD - 1) * 16, np.int16)
    minX1 = max(minD + D, 0)
    out[:, minX1:minX1 + W1] = (sub + minD * 16).astype(np.int16)
    return out

tot = {}
for seed in range(1, 9):
    l, r = synth.stereo_pair(1280, 720, 128, seed)
    ls = cv2.resize(l, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
    rs = cv2.resize(r, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
    for (a, b, minD, D) in ((rs, ls, -79, 80), (ls, rs, 0, 64), (ls[:, ::-1].copy(), rs[:, ::-1].copy(), -40, 96)):
        A = (minD, D, 5, 600, 2400, 1000000, 63, 0, 0, 0, 2)
        ref = cv2.StereoSGBM_create(*A).compute(a, b)
        raw, _, Sv, _ = orc.sgbm_match(a, b, orc.make_params(*A), want_volumes=True)
        for rule in ("first", (8, True), (16, True), (8, False)):
            o = orc.median3(wta(Sv, minD, a.shape[1], rule))
            n = int((o != ref).sum())
            tot[str(rule)] = tot.get(str(rule), 0) + n
        print(seed, minD, D, {k: v for k, v in tot.items()}, flush=True)
