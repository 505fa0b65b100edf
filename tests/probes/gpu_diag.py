"""GPU-side stage-by-stage comparison of the CUDA SGBM path against the C oracle.
Debug tool (not a test): python tests/probes/gpu_diag.py  -> gpurun_out/diag.txt"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import orc
from stereo_depth_ruler_b200 import capi, synth

out_dir = "gpurun_out"; os.makedirs(out_dir, exist_ok=True)
log = open(os.path.join(out_dir, "diag.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s); log.write(s + "\n"); log.flush()

ctx = capi.Context(0)
P(capi.lib().sdr_version().decode())

def first_bad(a, b, n=5):
    idx = np.argwhere(a != b)[:n]
    return [(tuple(int(v) for v in i), int(a[tuple(i)]), int(b[tuple(i)])) for i in idx]

def case(W, H, D, mode, bs=5, uq=12, d12=1, cap=63, sw=200, sr=2, minD=0, seed=0, kind="tex"):
    if kind == "tex": l, r = synth.stereo_pair(W, H, D, seed)
    else: l, r = synth.noise_pair(W, H, seed, 256 if kind == "noise" else 8)
    P1, P2 = 8 * 3 * bs * bs, 32 * 3 * bs * bs
    args = (minD, D, bs, P1, P2, d12, cap, uq, sw, sr, mode)
    po = orc.make_params(*args); pg = capi.make_params(*args)
    m = capi.Sgbm(ctx, pg)
    t = time.time(); got = m.compute(l, r); tg = time.time() - t
    ref = orc.sgbm_compute(l, r, po)
    bad = int((ref != got).sum())
    msg = f"{W}x{H}x{D} mode{mode} bs{bs} uq{uq} d12{d12} cap{cap} sw{sw} minD{minD} {kind}: final mism {bad}/{ref.size} gpu {tg*1e3:.1f}ms"
    if bad or True:
        raw_g = m.debug_read(2, (H, W))
        if mode != 2:
            raw_o, Cv, Sv, Sp = orc.sgbm_match(l, r, po, want_volumes=True)
            W1 = Cv.shape[1]
            if W1 > 0:
                Cg = m.debug_read(0, (H, W1, D)); Pp2 = max(P2, P1 + 1)
                cb = int((Cg.astype(np.int32) + Pp2 != Cv).sum())
                msg += f" | C mism {cb}"
                if cb: msg += f" first {first_bad(Cg.astype(np.int32) + Pp2, Cv.astype(np.int32))}"
                if mode == 0:
                    Sg = m.debug_read(1, (H, W1, D)); sb = int((Sg != Sp).sum())
                    msg += f" | Spre mism {sb}"
                    if sb: msg += f" first {first_bad(Sg, Sp)}"
        else:
            raw_o = orc.sgbm_match(l, r, po)
        rb = int((raw_g != raw_o).sum())
        msg += f" | raw mism {rb}"
        if rb: msg += f" first {first_bad(raw_g, raw_o)}"
    P(msg)
    m.close()
    return bad

tot = 0
for mode in (0, 1, 2):
    tot += case(160, 96, 32, mode)
    tot += case(200, 123, 48, mode, bs=3, uq=5, d12=2, cap=31)
    tot += case(140, 70, 16, mode, bs=7, uq=0, d12=-1, cap=5, kind="noise")
    tot += case(140, 64, 32, mode, bs=9, uq=15, kind="low")
    tot += case(180, 90, 64, mode, bs=1, uq=30, sw=0)
    tot += case(180, 90, 32, mode, bs=5, minD=-8)
    tot += case(180, 90, 32, mode, bs=5, minD=4)
    tot += case(300, 100, 80, mode)
tot += case(640, 360, 128, 0)
tot += case(1280, 720, 128, 0)
tot += case(1280, 720, 128, 0)   # warm second run for a timing hint
P("TOTAL", tot)
# standalone post filters
rng = np.random.default_rng(0)
x = rng.integers(-16, 2000, size=(97, 131)).astype(np.int16)
P("median mism", int((capi.median3_16s(ctx, x) != orc.median3(x)).sum()))
for (ms_, md) in ((200, 32), (50, 16), (10, 1)):
    y = (rng.integers(0, 40, size=(120, 160)) * ms_ // 10).astype(np.int16); y[rng.random(y.shape) < 0.2] = -16
    P("speckle mism", int((capi.filter_speckles(ctx, y, -16, ms_, md) != orc.filter_speckles(y, -16, ms_, md)).sum()))
P("launches", ctx.kernel_launches)
