"""ctypes loader for the C oracle (oracle/sdr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py. The product package
(stereo_depth_ruler_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libsdr_oracle.so")

MODE_SGBM, MODE_HH, MODE_SGBM_3WAY = 0, 1, 2


class SgbmParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
        "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")]


def _src_digest():
    import hashlib
    h = hashlib.sha256()
    for f in ("sdr_oracle.c", "Makefile"):
        with open(os.path.join(_HERE, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False):
    """Builds libsdr_oracle.so when it is missing or was built from other sources (content hash in a sidecar file: file
    times mean nothing after the tree has been copied to another machine). Safe against concurrent callers -- the worker
    processes of the CPU baseline all come through here: one builder at a time (flock), the library is written under a
    temporary name and moved into place, so nobody ever maps a half-written file."""
    import fcntl
    stamp = _LIB + ".sha256"
    want = _src_digest()

    def fresh():
        try:
            with open(stamp) as fh:
                return os.path.exists(_LIB) and fh.read().strip() == want
        except OSError:
            return False
    if not force and fresh():
        return _LIB
    with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if force or not fresh():
            tmp = "%s.tmp.%d" % (_LIB, os.getpid())
            subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libsdr_oracle.so", "OUT=" + tmp])
            os.replace(tmp, _LIB)
            with open(stamp + ".tmp", "w") as fh:
                fh.write(want)
            os.replace(stamp + ".tmp", stamp)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_voxelgrid.restype = C.c_long
        _lib.orc_measure_distance.restype = C.c_double
        _lib.orc_depth_coverage.restype = C.c_double
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def make_params(minDisparity=0, numDisparities=128, blockSize=5, P1=600, P2=2400, disp12MaxDiff=1,
                preFilterCap=63, uniquenessRatio=12, speckleWindowSize=200, speckleRange=2, mode=MODE_SGBM):
    return SgbmParams(minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap,
                      uniquenessRatio, speckleWindowSize, speckleRange, mode)


def init_rectify_map(K, dist, R, P, size):
    W, H = size
    K = np.ascontiguousarray(K, np.float64); dist = np.ascontiguousarray(dist, np.float64).ravel()
    R = np.ascontiguousarray(R, np.float64); P = np.ascontiguousarray(P, np.float64)
    m1 = np.empty((H, W, 2), np.int16); m2 = np.empty((H, W), np.uint16)
    rc = lib().orc_init_rectify_map(_p(K), _p(dist), _p(R), _p(P), W, H, _p(m1), _p(m2))
    assert rc == 0
    return m1, m2


def remap(src, m1, m2):
    assert src.dtype == np.uint8 and src.strides[-1] == 1
    ch = 1 if src.ndim == 2 else src.shape[2]
    if src.ndim == 3:
        assert src.strides[1] == ch
    H, W = m2.shape
    dst = np.empty((H, W) + (() if src.ndim == 2 else (ch,)), np.uint8)
    lib().orc_remap_bilinear(_p(src), C.c_size_t(src.strides[0]), src.shape[1], src.shape[0], ch,
                             _p(np.ascontiguousarray(m1)), _p(np.ascontiguousarray(m2)), W, H,
                             _p(dst), C.c_size_t(dst.strides[0]))
    return dst


def bgr2gray(bgr):
    assert bgr.dtype == np.uint8 and bgr.ndim == 3 and bgr.shape[2] == 3 and bgr.strides[1] == 3
    H, W = bgr.shape[:2]
    g = np.empty((H, W), np.uint8)
    lib().orc_bgr2gray(_p(bgr), C.c_size_t(bgr.strides[0]), W, H, _p(g), C.c_size_t(W))
    return g


def sgbm_match(left, right, params, want_volumes=False):
    """Raw matcher output (before median/speckle). Optionally the C (+P2), final S and pre-WTA S volumes."""
    left = np.ascontiguousarray(left); right = np.ascontiguousarray(right)
    H, W = left.shape
    disp = np.empty((H, W), np.int16)
    Cv = Sv = Sp = None
    if want_volumes:
        D = params.numDisparities
        minX1 = max(params.minDisparity + D, 0); maxX1 = W + min(params.minDisparity, 0)
        W1 = max(maxX1 - minX1, 0)
        Cv = np.zeros((H, W1, D), np.int16); Sv = np.zeros((H, W1, D), np.int16); Sp = np.zeros((H, W1, D), np.int16)
    rc = lib().orc_sgbm_match(_p(left), C.c_size_t(W), _p(right), C.c_size_t(W), W, H, C.byref(params),
                              _p(disp), _p(Cv), _p(Sv), _p(Sp))
    if rc:
        raise ValueError("orc_sgbm_match rc=%d" % rc)
    return (disp, Cv, Sv, Sp) if want_volumes else disp


def sgbm_compute(left, right, params):
    left = np.ascontiguousarray(left); right = np.ascontiguousarray(right)
    H, W = left.shape
    disp = np.empty((H, W), np.int16)
    rc = lib().orc_sgbm_compute(_p(left), C.c_size_t(W), _p(right), C.c_size_t(W), W, H, C.byref(params), _p(disp))
    if rc:
        raise ValueError("orc_sgbm_compute rc=%d" % rc)
    return disp


def median3(src):
    src = np.ascontiguousarray(src, np.int16)
    H, W = src.shape
    dst = np.empty_like(src)
    lib().orc_median3_16s(_p(src), C.c_size_t(W), W, H, _p(dst), C.c_size_t(W))
    return dst


def filter_speckles(img, new_val, max_size, max_diff):
    out = np.ascontiguousarray(img, np.int16).copy()
    H, W = out.shape
    rc = lib().orc_filter_speckles(_p(out), C.c_size_t(W), W, H, int(new_val), int(max_size), int(max_diff))
    assert rc == 0
    return out


def disp_to_float(disp):
    disp = np.ascontiguousarray(disp, np.int16)
    out = np.empty(disp.shape, np.float32)
    lib().orc_disp_to_float(_p(disp), disp.size, _p(out))
    return out


def reproject_to_3d(disp_f32, Q, handle_missing=False):
    disp_f32 = np.ascontiguousarray(disp_f32, np.float32)
    Q = np.ascontiguousarray(Q, np.float64)
    H, W = disp_f32.shape
    out = np.empty((H, W, 3), np.float32)
    lib().orc_reproject_to_3d(_p(disp_f32), W, H, _p(Q), int(bool(handle_missing)), _p(out))
    return out


VOXEL_OVERFLOW = -3


def voxelgrid(points, bgr, leaf, drop_mask=None):
    """Returns dict(xyz, rgb, idx, cnt, minb, divb) or None when PCL's int32 guard trips (passthrough)."""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    n = pts.shape[0]
    col = None if bgr is None else np.ascontiguousarray(bgr, np.uint8).reshape(-1, 3)
    dm = None if drop_mask is None else np.ascontiguousarray(drop_mask, np.uint8).ravel()
    leaf = (leaf,) * 3 if np.isscalar(leaf) else tuple(leaf)
    xyz = np.empty((n, 3), np.float32); rgb = np.empty((n, 3), np.uint8)
    idx = np.empty(n, np.uint32); cnt = np.empty(n, np.uint32); mb = np.zeros(6, np.int32)
    nv = lib().orc_voxelgrid(_p(pts), _p(col), _p(dm), C.c_long(n), C.c_float(leaf[0]), C.c_float(leaf[1]),
                             C.c_float(leaf[2]), _p(xyz), _p(rgb), _p(idx), _p(cnt), _p(mb))
    if nv == VOXEL_OVERFLOW:
        return None
    assert nv >= 0
    return dict(xyz=xyz[:nv].copy(), rgb=rgb[:nv].copy(), idx=idx[:nv].copy(), cnt=cnt[:nv].copy(),
                minb=mb[:3].copy(), divb=mb[3:].copy())


def measure_distance(points, p1, p2):
    pts = np.ascontiguousarray(points, np.float32)
    return lib().orc_measure_distance(_p(pts), pts.shape[1], int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1]))


def depth_coverage(points):
    pts = np.ascontiguousarray(points, np.float32)
    n = C.c_long()
    pct = lib().orc_depth_coverage(_p(pts), pts.shape[1], pts.shape[0], C.byref(n))
    return n.value, pct


def ingest_to_bgr(fmt, src):
    """cv2.cvtColor(src, COLOR_YUV2BGR_NV12 (fmt 1) | COLOR_YUV2BGR_YUY2 (fmt 2)) restated."""
    src = np.ascontiguousarray(src, np.uint8)
    H = src.shape[0] * 2 // 3 if fmt == 1 else src.shape[0]
    W = src.shape[1]
    out = np.empty((H, W, 3), np.uint8)
    rc = lib().orc_ingest_to_bgr(int(fmt), _p(src), C.c_long(src.strides[0]), W, H, _p(out), C.c_long(out.strides[0]))
    if rc:
        raise ValueError("ingest: bad format or odd size")
    return out
