"""Synthetic textured stereo pairs: numpy only, seeded, byte-identical for
every consumer (cv2, the C oracle and the CUDA path all get the same arrays)."""
import numpy as np


def _bilinear_upsample(a, H, W):
    h, w = a.shape
    ys = np.linspace(0, h - 1, H); xs = np.linspace(0, w - 1, W)
    y0 = np.minimum(ys.astype(np.int64), h - 2); x0 = np.minimum(xs.astype(np.int64), w - 2)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    a00 = a[y0][:, x0]; a01 = a[y0][:, x0 + 1]; a10 = a[y0 + 1][:, x0]; a11 = a[y0 + 1][:, x0 + 1]
    return a00 * (1 - fy) * (1 - fx) + a01 * (1 - fy) * fx + a10 * fy * (1 - fx) + a11 * fy * fx


def texture(H, W, rng):
    t = np.zeros((H, W), np.float64)
    for s in (1, 2, 4, 8, 16):
        n = rng.random((H // s + 2, W // s + 2))
        t += s * _bilinear_upsample(n, H, W)
    t -= t.min(); t /= max(t.max(), 1e-12)
    return t * 255.0


def disparity_field(H, W, D):
    y = np.arange(H, dtype=np.float64)[:, None]; x = np.arange(W)[None, :]
    scale = D / 128.0
    d = (20 + 60 * y / H + 20 * ((x // 160) % 2)) * scale
    return np.clip(d, 1, D - 18)


def stereo_pair(W=1280, H=720, D=128, seed=0):
    """Returns (left_gray, right_gray) uint8 [H][W]. left(x) = right(x - d(x,y))."""
    rng = np.random.default_rng(seed)
    pad = 2 * D
    T = texture(H, W + pad, rng)
    right = T[:, :W]
    d = disparity_field(H, W, D)
    xs = np.arange(W, dtype=np.float64)[None, :] - d
    xs = np.abs(xs)  # reflect at the left border
    x0 = np.floor(xs).astype(np.int64); fx = xs - x0
    x0 = np.clip(x0, 0, W + pad - 2)
    rows = np.arange(H)[:, None]
    left = T[rows, x0] * (1 - fx) + T[rows, x0 + 1] * fx
    to8 = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)
    return to8(left), to8(right)


def colorize(gray, seed=0):
    """BGR = clip(gray * (0.9, 1.0, 1.1) + U{-3..3})."""
    rng = np.random.default_rng(seed + 7919)
    g = gray.astype(np.float64)[..., None] * np.array([0.9, 1.0, 1.1])
    g += rng.integers(-3, 4, size=g.shape)
    return np.clip(np.rint(g), 0, 255).astype(np.uint8)


def noise_pair(W, H, seed=0, amplitude=256):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, amplitude, size=(H, W)).astype(np.uint8)
    b = rng.integers(0, amplitude, size=(H, W)).astype(np.uint8)
    return a, b


def bgr_to_nv12(bgr):
    """A BGR picture as a hardware decoder would hand it over: NV12, BT.601 limited range, 2x2 chroma average.
    [H][W][3] -> [H * 3 / 2][W] (Y plane, then interleaved U, V). Synthetic-input helper: integer BT.601 approximation."""
    H, W = bgr.shape[:2]
    b = bgr[..., 0].astype(np.int32); g = bgr[..., 1].astype(np.int32); r = bgr[..., 2].astype(np.int32)
    y = ((66 * r + 129 * g + 25 * b + 128) >> 8) + 16
    u = ((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128
    v = ((112 * r - 94 * g - 18 * b + 128) >> 8) + 128
    avg = lambda c: (c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2] + 2) >> 2
    out = np.empty((H * 3 // 2, W), np.uint8)
    out[:H] = np.clip(y, 0, 255)
    out[H:, 0::2] = np.clip(avg(u), 0, 255)
    out[H:, 1::2] = np.clip(avg(v), 0, 255)
    return out
