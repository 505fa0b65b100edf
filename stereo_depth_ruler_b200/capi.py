"""ctypes binding of the C ABI in include/sdr_c.h (libsdr_b200.so).

There is no CPU fallback: if the CUDA library is missing or no B200 is visible every
compute entry point raises. Images may be numpy arrays (host memory) or torch CUDA tensors
(device memory, used in place).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdr_b200.so")

SDR_OK, SDR_EINVAL, SDR_ESIZE, SDR_ECUDA, SDR_ENOMEM, SDR_EIO, SDR_EVOXEL_OVERFLOW = range(7)
MEM_HOST, MEM_PINNED, MEM_DEVICE = 0, 1, 2
MODE_SGBM, MODE_HH, MODE_SGBM_3WAY = 0, 1, 2

_STATUS = {0: "SDR_OK", 1: "SDR_EINVAL", 2: "SDR_ESIZE", 3: "SDR_ECUDA", 4: "SDR_ENOMEM", 5: "SDR_EIO",
           6: "SDR_EVOXEL_OVERFLOW"}


class SdrError(RuntimeError):
    def __init__(self, status, msg=""):
        self.status = status
        super().__init__("%s: %s" % (_STATUS.get(status, status), msg))


class Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("channels", C.c_int),
                ("elem_size", C.c_int), ("step_bytes", C.c_size_t), ("mem", C.c_int)]


class Calibration(C.Structure):
    _fields_ = [("image_width", C.c_int), ("image_height", C.c_int),
                ("K_left", C.c_double * 9), ("dist_left", C.c_double * 5),
                ("K_right", C.c_double * 9), ("dist_right", C.c_double * 5),
                ("R", C.c_double * 9), ("T", C.c_double * 3), ("E", C.c_double * 9), ("F", C.c_double * 9),
                ("R1", C.c_double * 9), ("R2", C.c_double * 9), ("P1", C.c_double * 12), ("P2", C.c_double * 12),
                ("Q", C.c_double * 16), ("present", C.c_uint32)]

    def mat(self, name):
        shape = {"K_left": (3, 3), "K_right": (3, 3), "dist_left": (1, 5), "dist_right": (1, 5), "R": (3, 3),
                 "T": (3, 1), "E": (3, 3), "F": (3, 3), "R1": (3, 3), "R2": (3, 3), "P1": (3, 4), "P2": (3, 4),
                 "Q": (4, 4)}[name]
        return np.array(getattr(self, name), np.float64).reshape(shape)


class SgbmParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
        "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")]


class VoxelOptions(C.Structure):
    _fields_ = [("leaf", C.c_float * 3), ("drop_invalid", C.c_int), ("z_max", C.c_float), ("min_disparity", C.c_int)]


class CloudView(C.Structure):
    _fields_ = [("n", C.c_int64), ("xyz", C.c_void_p), ("rgb", C.c_void_p), ("idx", C.c_void_p), ("cnt", C.c_void_p),
                ("min_b", C.c_int32 * 3), ("div_b", C.c_int32 * 3), ("passthrough", C.c_int)]


class PipelineOptions(C.Structure):
    _fields_ = [("sgbm", SgbmParams), ("voxel", VoxelOptions), ("handle_missing", C.c_int), ("batch", C.c_int),
                ("want_points", C.c_int)]


# every symbol include/sdr_c.h declares (tests check the library exports all of them)
SYMBOLS = [
    "sdr_version", "sdr_device_count", "sdr_create", "sdr_destroy", "sdr_set_stream", "sdr_get_stream",
    "sdr_synchronize", "sdr_last_error", "sdr_kernel_launches",
    "sdr_load_calibration_yaml", "sdr_save_calibration_yaml", "sdr_calibration_is_valid",
    "sdr_rectifier_create", "sdr_rectifier_destroy", "sdr_rectifier_get_maps", "sdr_rectify",
    "sdr_remap_with_maps", "sdr_bgr2gray",
    "sdr_sgbm_create", "sdr_sgbm_destroy", "sdr_sgbm_compute", "sdr_sgbm_compute_batch",
    "sdr_sgbm_set_profiling", "sdr_sgbm_stage_times", "sdr_sgbm_debug_read", "sdr_sgbm_set_option",
    "sdr_pipeline_set_sgbm_option", "sdr_resize_half_area", "sdr_sgbm_right_matcher_params", "sdr_vis_create", "sdr_vis_destroy",
    "sdr_vis_reset", "sdr_vis_disparity", "sdr_vis_depth", "sdr_vis_depth_range", "sdr_sgbm_split_create", "sdr_sgbm_split_destroy", "sdr_sgbm_split_compute",
    "sdr_sgbm_split_synchronize", "sdr_sgbm_split_last_stats", "sdr_sgbm_split_last_error",
    "sdr_sgbm_stripe_rows", "sdr_sgbm_match_stripes", "sdr_sgbm_postfilter",
    "sdr_median3_16s", "sdr_filter_speckles", "sdr_disp_to_float", "sdr_reproject_to_3d",
    "sdr_cloud_create", "sdr_cloud_destroy", "sdr_voxelgrid", "sdr_depth_cloud", "sdr_cloud_download",
    "sdr_pipeline_create", "sdr_pipeline_destroy", "sdr_pipeline_run", "sdr_pipeline_stage_times",
    "sdr_pipeline_cloud_download", "sdr_pipeline_clouds_download", "sdr_pipeline_clouds_download_packed",
    "sdr_pipeline_set_profiling",
    "sdr_measure_distance", "sdr_depth_coverage", "sdr_cloud_save_pcd_binary", "sdr_stream_run",
    "sdr_stream_create", "sdr_stream_destroy", "sdr_stream_process", "sdr_stream_process_pool", "sdr_bind_host_to_device",
    "sdr_ingest_to_bgr", "sdr_pipeline_set_input_format", "sdr_stream_set_input_format", "sdr_stream_set_sgbm_option", "sdr_pipeline_stage_timeline",
]

_lib = None


def lib():
    """Loads libsdr_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdrError(SDR_ECUDA, "%s is missing: run `python -m stereo_depth_ruler_b200.build` "
                                      "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.sdr_version.restype = C.c_char_p
        L.sdr_last_error.restype = C.c_char_p
        L.sdr_last_error.argtypes = [C.c_void_p]
        L.sdr_get_stream.restype = C.c_void_p
        L.sdr_get_stream.argtypes = [C.c_void_p]
        L.sdr_kernel_launches.restype = C.c_uint64
        L.sdr_kernel_launches.argtypes = [C.c_void_p]
        L.sdr_destroy.argtypes = [C.c_void_p]
        L.sdr_destroy.restype = None
        for n in ("sdr_rectifier_destroy", "sdr_sgbm_destroy", "sdr_cloud_destroy", "sdr_pipeline_destroy", "sdr_stream_destroy", "sdr_sgbm_split_destroy", "sdr_vis_destroy"):
            if hasattr(L, n):
                getattr(L, n).argtypes = [C.c_void_p]
                getattr(L, n).restype = None
        _lib = L
    return _lib


def make_params(minDisparity=0, numDisparities=128, blockSize=5, P1=600, P2=2400, disp12MaxDiff=1,
                preFilterCap=63, uniquenessRatio=12, speckleWindowSize=200, speckleRange=2, mode=MODE_SGBM):
    """Same positional order as cv::StereoSGBM::create (stereo_disparity.cpp:5-6)."""
    return SgbmParams(minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap,
                      uniquenessRatio, speckleWindowSize, speckleRange, mode)


def stripe_rows(params, width, height):
    """[(src_start, dst_start, row_end)] of the independent stripes StereoSGBM splits this image into (4 for 3-way)."""
    n = lib().sdr_sgbm_stripe_rows(C.byref(params), int(width), int(height), -1, None, None, None)
    out = []
    for i in range(n):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().sdr_sgbm_stripe_rows(C.byref(params), int(width), int(height), i, C.byref(a), C.byref(b), C.byref(c))
        out.append((a.value, b.value, c.value))
    return out


def _is_torch(a):
    return type(a).__module__.startswith("torch")


_NP_ELEM = {np.dtype(np.uint8): 1, np.dtype(np.int16): 2, np.dtype(np.uint16): 2, np.dtype(np.float32): 4}


def as_image(a, channels=None):
    """numpy array / torch tensor [H][W] or [H][W][C] -> (Image, keepalive)."""
    if isinstance(a, Image):
        return a, a
    if _is_torch(a):
        assert a.is_cuda, "torch tensors must live on the GPU (host data: pass numpy)"
        es = a.element_size()
        shape, strides = tuple(a.shape), tuple(s * es for s in a.stride())
        ptr, mem = a.data_ptr(), MEM_DEVICE
    else:
        assert isinstance(a, np.ndarray) and a.dtype in _NP_ELEM, "unsupported array %r" % (type(a),)
        es = a.dtype.itemsize
        shape, strides = a.shape, a.strides
        ptr, mem = a.ctypes.data, MEM_HOST
    if len(shape) == 2:
        H, W = shape
        ch = 1
        assert strides[1] == es
    else:
        H, W, ch = shape
        assert strides[2] == es and strides[1] == es * ch, "pixels must be packed"
    if channels is not None:
        assert ch == channels, "expected %d channels, got %d" % (channels, ch)
    return Image(ptr, H, W, ch, es, strides[0], mem), a


class Context:
    """One device, one stream (sdr_create / sdr_destroy)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().sdr_create(int(device), C.byref(self._h))
        if rc != SDR_OK:
            raise SdrError(rc, "sdr_create(device=%d) failed: no usable sm_100 CUDA device (no CPU fallback)" % device)
        self.device = device

    def check(self, rc, ok=(SDR_OK,)):
        if rc not in ok:
            raise SdrError(rc, (lib().sdr_last_error(self._h) or b"").decode())
        return rc

    def set_stream(self, cuda_stream_ptr):
        self.check(lib().sdr_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self.check(lib().sdr_synchronize(self._h))

    @property
    def kernel_launches(self):
        return int(lib().sdr_kernel_launches(self._h))

    def close(self):
        if self._h:
            lib().sdr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_calibration(path):
    c = Calibration()
    rc = lib().sdr_load_calibration_yaml(str(path).encode(), C.byref(c))
    if rc != SDR_OK:
        raise SdrError(rc, "could not load calibration %s" % path)
    return c


def save_calibration(path, calib):
    rc = lib().sdr_save_calibration_yaml(str(path).encode(), C.byref(calib))
    if rc != SDR_OK:
        raise SdrError(rc, "could not save calibration %s" % path)


class Sgbm:
    """cv::StereoSGBM stand-in (sdr_sgbm_*)."""

    def __init__(self, ctx, params):
        self.ctx = ctx
        self.params = params
        self._h = C.c_void_p()
        ctx.check(lib().sdr_sgbm_create(ctx._h, C.byref(params), C.byref(self._h)))

    def compute(self, left, right, out=None):
        li, k1 = as_image(left, 1)
        ri, k2 = as_image(right, 1)
        if out is None:
            if _is_torch(left):
                import torch
                out = torch.empty(tuple(left.shape), dtype=torch.int16, device=left.device)
            else:
                out = np.empty(left.shape, np.int16)
        oi, k3 = as_image(out, 1)
        self.ctx.check(lib().sdr_sgbm_compute(self._h, C.byref(li), C.byref(ri), C.byref(oi)))
        return out

    def compute_batch(self, lefts, rights, outs):
        """lefts/rights/outs: [B][H][W] arrays (numpy or torch cuda)."""
        n = len(lefts)
        L = (Image * n)(*[as_image(lefts[i], 1)[0] for i in range(n)])
        R = (Image * n)(*[as_image(rights[i], 1)[0] for i in range(n)])
        O = (Image * n)(*[as_image(outs[i], 1)[0] for i in range(n)])
        self.ctx.check(lib().sdr_sgbm_compute_batch(self._h, n, L, R, O))
        return outs

    def match_stripes(self, left, right, first, count, raw):
        """Raw matcher output of stripes [first, first+count) into the rows they own of `raw` (int16, full size)."""
        li, _ = as_image(left, 1); ri, _ = as_image(right, 1); oi, _ = as_image(raw, 1)
        self.ctx.check(lib().sdr_sgbm_match_stripes(self._h, C.byref(li), C.byref(ri), int(first), int(count), C.byref(oi)))
        return raw

    def postfilter(self, raw, out=None):
        if out is None:
            out = _empty_like_kind(raw, tuple(raw.shape), np.int16, "int16")
        ri, _ = as_image(raw, 1); oi, _ = as_image(out, 1)
        self.ctx.check(lib().sdr_sgbm_postfilter(self._h, C.byref(ri), C.byref(oi)))
        return out

    def set_option(self, name, value):
        """Kernel-selection knob (sdr_sgbm_set_option); results never depend on it."""
        self.ctx.check(lib().sdr_sgbm_set_option(self._h, name.encode(), int(value)))
        return self

    def set_profiling(self, on=True):
        lib().sdr_sgbm_set_profiling(self._h, int(on))

    def stage_times(self):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = lib().sdr_sgbm_stage_times(self._h, 64, names, ms)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def debug_read(self, what, shape):
        out = np.empty(shape, np.int16)
        self.ctx.check(lib().sdr_sgbm_debug_read(self._h, int(what), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes)))
        return out

    def close(self):
        if self._h:
            lib().sdr_sgbm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SplitStats(C.Structure):
    _fields_ = [("n_parts", C.c_int), ("n_stripes", C.c_int), ("device", C.c_int * 16), ("stripes", C.c_int * 16),
                ("input_rows", C.c_int * 16), ("peer_access", C.c_int * 16), ("input_bytes", C.c_int64 * 16),
                ("gather_bytes", C.c_int64 * 16), ("ms", C.c_float)]


class SgbmSplit:
    """One oversize frame split by stripes over several GPUs in one process (sdr_sgbm_split_*; BASELINE config 5)."""

    def __init__(self, devices, params):
        self.params = params
        self._h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        rc = lib().sdr_sgbm_split_create(devs, len(devices), C.byref(params), C.byref(self._h))
        if rc != SDR_OK:
            raise SdrError(rc, "sdr_sgbm_split_create(%r) failed" % (list(devices),))
        lib().sdr_sgbm_split_last_error.restype = C.c_char_p
        lib().sdr_sgbm_split_last_error.argtypes = [C.c_void_p]

    def compute(self, left, right, out=None):
        li, _ = as_image(left, 1); ri, _ = as_image(right, 1)
        if out is None:
            out = _empty_like_kind(left, tuple(left.shape), np.int16, "int16")
        oi, _ = as_image(out, 1)
        rc = lib().sdr_sgbm_split_compute(self._h, C.byref(li), C.byref(ri), C.byref(oi))
        if rc != SDR_OK:
            raise SdrError(rc, (lib().sdr_sgbm_split_last_error(self._h) or b"").decode())
        return out

    def synchronize(self):
        rc = lib().sdr_sgbm_split_synchronize(self._h)
        if rc != SDR_OK:
            raise SdrError(rc, "sdr_sgbm_split_synchronize failed")

    def stats(self):
        st = SplitStats()
        rc = lib().sdr_sgbm_split_last_stats(self._h, C.byref(st))
        if rc != SDR_OK:
            raise SdrError(rc, "sdr_sgbm_split_last_stats failed")
        n = st.n_parts
        return dict(ms=float(st.ms), n_parts=n, n_stripes=st.n_stripes, device=list(st.device[:n]), stripes=list(st.stripes[:n]),
                    input_rows=list(st.input_rows[:n]), peer_access=list(st.peer_access[:n]), input_bytes=list(st.input_bytes[:n]),
                    gather_bytes=list(st.gather_bytes[:n]))

    def close(self):
        if self._h:
            lib().sdr_sgbm_split_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def median3_16s(ctx, src):
    out = np.empty_like(src)
    si, _ = as_image(src, 1)
    oi, _ = as_image(out, 1)
    ctx.check(lib().sdr_median3_16s(ctx._h, C.byref(si), C.byref(oi)))
    return out


def filter_speckles(ctx, img, new_val, max_size, max_diff):
    out = np.ascontiguousarray(img, np.int16).copy()
    ii, _ = as_image(out, 1)
    ctx.check(lib().sdr_filter_speckles(ctx._h, C.byref(ii), int(new_val), int(max_size), int(max_diff)))
    return out


def _empty_like_kind(ref, shape, np_dtype, torch_dtype_name):
    if _is_torch(ref):
        import torch
        return torch.empty(shape, dtype=getattr(torch, torch_dtype_name), device=ref.device)
    return np.empty(shape, np_dtype)


class Rectifier:
    """StereoRectifier stand-in (sdr_rectifier_*): fused initUndistortRectifyMap + remap (+ BGR2GRAY)."""

    def __init__(self, ctx, calib):
        self.ctx = ctx
        self.calib = calib
        self.size = (calib.image_width, calib.image_height)
        self._h = C.c_void_p()
        ctx.check(lib().sdr_rectifier_create(ctx._h, C.byref(calib), C.byref(self._h)))

    def maps(self, eye):
        W, H = self.size
        m1 = np.empty((H, W, 2), np.int16)
        m2 = np.empty((H, W), np.uint16)
        self.ctx.check(lib().sdr_rectifier_get_maps(self._h, int(eye), m1.ctypes.data_as(C.c_void_p), m2.ctypes.data_as(C.c_void_p)))
        return m1, m2

    def rectify(self, left_src, right_src, want_gray=False):
        W, H = self.size
        ch = 1 if len(left_src.shape) == 2 else left_src.shape[2]
        shp = (H, W) if ch == 1 else (H, W, ch)
        ld = _empty_like_kind(left_src, shp, np.uint8, "uint8")
        rd = _empty_like_kind(left_src, shp, np.uint8, "uint8")
        li, _ = as_image(left_src); ri, _ = as_image(right_src)
        ldi, _ = as_image(ld); rdi, _ = as_image(rd)
        if want_gray:
            lg = _empty_like_kind(left_src, (H, W), np.uint8, "uint8")
            rg = _empty_like_kind(left_src, (H, W), np.uint8, "uint8")
            lgi, _ = as_image(lg); rgi, _ = as_image(rg)
            self.ctx.check(lib().sdr_rectify(self._h, C.byref(li), C.byref(ri), C.byref(ldi), C.byref(rdi), C.byref(lgi), C.byref(rgi)))
            return ld, rd, lg, rg
        self.ctx.check(lib().sdr_rectify(self._h, C.byref(li), C.byref(ri), C.byref(ldi), C.byref(rdi), None, None))
        return ld, rd

    def close(self):
        if self._h:
            lib().sdr_rectifier_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def remap_with_maps(ctx, src, map1, map2):
    H, W = map2.shape
    if tuple(map1.shape) != (H, W, 2):      # the ABI takes bare pointers sized by the destination: check here
        raise SdrError(SDR_ESIZE, "remap: map1 must be [H][W][2] and map2 [H][W] of the same H x W")
    ch = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty((H, W) if ch == 1 else (H, W, ch), np.uint8)
    si, _ = as_image(src); di, _ = as_image(dst)
    m1 = np.ascontiguousarray(map1, np.int16); m2 = np.ascontiguousarray(map2, np.uint16)
    ctx.check(lib().sdr_remap_with_maps(ctx._h, C.byref(si), m1.ctypes.data_as(C.c_void_p), m2.ctypes.data_as(C.c_void_p), C.byref(di)))
    return dst


def bgr2gray(ctx, bgr):
    out = _empty_like_kind(bgr, tuple(bgr.shape[:2]), np.uint8, "uint8")
    si, _ = as_image(bgr, 3); di, _ = as_image(out, 1)
    ctx.check(lib().sdr_bgr2gray(ctx._h, C.byref(si), C.byref(di)))
    return out


def resize_half_area(ctx, src):
    """cv::resize(src, Size(), 0.5, 0.5, INTER_AREA) for even sizes (stereo_disparity.cpp:23-24)."""
    shp = (src.shape[0] // 2, src.shape[1] // 2) + tuple(src.shape[2:])
    out = _empty_like_kind(src, shp, np.uint8, "uint8")
    si, _ = as_image(src); di, _ = as_image(out)
    ctx.check(lib().sdr_resize_half_area(ctx._h, C.byref(si), C.byref(di)))
    return out


PIX_BGR, PIX_NV12, PIX_YUYV = 0, 1, 2
_PIX_CHANNELS = {PIX_BGR: 3, PIX_NV12: 1, PIX_YUYV: 2}


def ingest_to_bgr(ctx, fmt, src):
    """cv::cvtColor(src, COLOR_YUV2BGR_NV12 | COLOR_YUV2BGR_YUY2) (sdr_ingest_to_bgr): the colour conversion between a
    hardware decoder / UVC camera and the BGR frame of stereo_displayer.cpp:145-156. src: NV12 [H*3/2][W] or YUYV [H][W][2]."""
    H = src.shape[0] * 2 // 3 if fmt == PIX_NV12 else src.shape[0]
    out = _empty_like_kind(src, (H, src.shape[1], 3), np.uint8, "uint8")
    si, _ = as_image(src, _PIX_CHANNELS.get(fmt)); di, _ = as_image(out, 3)
    ctx.check(lib().sdr_ingest_to_bgr(ctx._h, int(fmt), C.byref(si), C.byref(di)))
    return out


def right_matcher_params(left_params):
    """cv::ximgproc::createRightMatcher (stereo_disparity.cpp:10) as parameters."""
    r = SgbmParams()
    rc = lib().sdr_sgbm_right_matcher_params(C.byref(left_params), C.byref(r))
    if rc != SDR_OK:
        raise SdrError(rc, "sdr_sgbm_right_matcher_params failed")
    return r


class Vis:
    """show_disparityMap / show_depthMap with the reference's frame-to-frame state (sdr_vis_*)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._h = C.c_void_p()
        ctx.check(lib().sdr_vis_create(ctx._h, C.byref(self._h)))

    def disparity(self, disp32f, num_disparities):
        out = _empty_like_kind(disp32f, tuple(disp32f.shape), np.uint8, "uint8")
        di, _ = as_image(disp32f, 1); oi, _ = as_image(out, 1)
        self.ctx.check(lib().sdr_vis_disparity(self._h, C.byref(di), int(num_disparities), C.byref(oi)))
        return out

    def depth(self, depth32f):
        out = _empty_like_kind(depth32f, tuple(depth32f.shape[:2]) + (3,), np.uint8, "uint8")
        di, _ = as_image(depth32f); oi, _ = as_image(out, 3)
        self.ctx.check(lib().sdr_vis_depth(self._h, C.byref(di), C.byref(oi)))
        return out

    def depth_range(self):
        a, b = C.c_double(), C.c_double()
        lib().sdr_vis_depth_range(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def reset(self):
        lib().sdr_vis_reset(self._h)

    def close(self):
        if self._h:
            lib().sdr_vis_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def disp_to_float(ctx, disp16):
    """convertTo(CV_32F, 1/16) on the device."""
    out = _empty_like_kind(disp16, tuple(disp16.shape), np.float32, "float32")
    di, _ = as_image(disp16, 1); oi, _ = as_image(out, 1)
    ctx.check(lib().sdr_disp_to_float(ctx._h, C.byref(di), C.byref(oi)))
    return out


def reproject_to_3d(ctx, disp, Q, handle_missing=False):
    """disp: int16 (CV_16S, /16 fused) or float32. Returns float32 [H][W][3]."""
    out = _empty_like_kind(disp, tuple(disp.shape) + (3,), np.float32, "float32")
    di, _ = as_image(disp, 1); oi, _ = as_image(out, 3)
    q = (C.c_double * 16)(*np.asarray(Q, np.float64).ravel())
    ctx.check(lib().sdr_reproject_to_3d(ctx._h, C.byref(di), q, int(bool(handle_missing)), C.byref(oi)))
    return out


def measure_distance(ctx, points, p1, p2):
    """Distance between two reprojected pixels p = (x, y), as the reference's mouse ruler (stereo_displayer.cpp:41-47)."""
    pi, _ = as_image(points, 3)
    out = C.c_double()
    ctx.check(lib().sdr_measure_distance(ctx._h, C.byref(pi), int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1]), C.byref(out)))
    return out.value


def depth_coverage(ctx, points, first_col=80, z_min=0.0, z_max=12000.0):
    """(count, percent) of StereoDisplayer::depth_coverage (stereo_displayer.cpp:105-118)."""
    pi, _ = as_image(points, 3)
    n, pct = C.c_int64(), C.c_double()
    ctx.check(lib().sdr_depth_coverage(ctx._h, C.byref(pi), int(first_col), C.c_float(z_min), C.c_float(z_max), C.byref(n), C.byref(pct)))
    return n.value, pct.value


def voxel_options(leaf, drop_invalid=False, z_max=0.0, min_disparity=0):
    leaf = (leaf,) * 3 if np.isscalar(leaf) else tuple(leaf)
    return VoxelOptions((C.c_float * 3)(*leaf), int(bool(drop_invalid)), float(z_max), int(min_disparity))


class Cloud:
    """convertCVMatToPCL + pcl::VoxelGrid stand-in (sdr_cloud_*)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._h = C.c_void_p()
        ctx.check(lib().sdr_cloud_create(ctx._h, C.byref(self._h)))
        self.view = CloudView()

    def _download(self):
        v = self.view
        n = int(v.n)
        xyz = np.empty((n, 3), np.float32); rgb = np.empty((n, 3), np.uint8)
        idx = np.empty(n, np.uint32); cnt = np.empty(n, np.uint32)
        self.ctx.check(lib().sdr_cloud_download(self._h, xyz.ctypes.data_as(C.c_void_p), rgb.ctypes.data_as(C.c_void_p),
                                                idx.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p)))
        res = dict(xyz=xyz, rgb=rgb, passthrough=bool(v.passthrough), minb=np.array(v.min_b[:]), divb=np.array(v.div_b[:]))
        if not v.passthrough:
            res.update(idx=idx, cnt=cnt)
        return res

    def voxelgrid(self, points, bgr, opt, disp16=None, download=True):
        pi, _ = as_image(points, 3)
        bi = as_image(bgr, 3)[0] if bgr is not None else None
        di = as_image(disp16, 1)[0] if disp16 is not None else None
        rc = lib().sdr_voxelgrid(self._h, C.byref(pi), C.byref(bi) if bi else None, C.byref(di) if di else None,
                                 C.byref(opt), C.byref(self.view))
        self.ctx.check(rc, ok=(SDR_OK, SDR_EVOXEL_OVERFLOW))
        return self._download() if download else self.view

    def save_pcd_binary(self, path, frame=0):
        self.ctx.check(lib().sdr_cloud_save_pcd_binary(self._h, int(frame), str(path).encode()))

    def depth_cloud(self, disp16, bgr, Q, opt, handle_missing=False, download=True):
        di, _ = as_image(disp16, 1)
        bi = as_image(bgr, 3)[0] if bgr is not None else None
        q = (C.c_double * 16)(*np.asarray(Q, np.float64).ravel())
        rc = lib().sdr_depth_cloud(self._h, C.byref(di), C.byref(bi) if bi else None, q, int(bool(handle_missing)),
                                   C.byref(opt), C.byref(self.view))
        self.ctx.check(rc, ok=(SDR_OK, SDR_EVOXEL_OVERFLOW))
        return self._download() if download else self.view

    def close(self):
        if self._h:
            lib().sdr_cloud_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pipeline:
    """SBS frame -> rectify -> SGBM -> reproject -> VoxelGrid, device resident (sdr_pipeline_*)."""

    def __init__(self, ctx, calib, sgbm_params, voxel_opt, handle_missing=False, batch=1, want_points=False):
        self.ctx = ctx
        self.calib = calib
        self.batch = batch
        self.opt = PipelineOptions(sgbm_params, voxel_opt, int(bool(handle_missing)), int(batch), int(bool(want_points)))
        self._h = C.c_void_p()
        ctx.check(lib().sdr_pipeline_create(ctx._h, C.byref(calib), C.byref(self.opt), C.byref(self._h)))
        self.views = (CloudView * batch)()

    def set_input_format(self, fmt):
        """SBS frames arrive as PIX_BGR (default), PIX_NV12 or PIX_YUYV and are converted on the device."""
        self.ctx.check(lib().sdr_pipeline_set_input_format(self._h, int(fmt)))
        self.input_format = int(fmt)

    def run(self, sbs_frames, disp=None, points=None, left_rect=None, clouds=True):
        n = len(sbs_frames)
        S = (Image * n)(*[as_image(f, _PIX_CHANNELS[getattr(self, "input_format", PIX_BGR)])[0] for f in sbs_frames])
        D = (Image * n)(*[as_image(d, 1)[0] for d in disp]) if disp is not None else None
        Pn = (Image * n)(*[as_image(p, 3)[0] for p in points]) if points is not None else None
        Lr = (Image * n)(*[as_image(p, 3)[0] for p in left_rect]) if left_rect is not None else None
        rc = lib().sdr_pipeline_run(self._h, n, S, D, Pn, Lr, self.views if clouds else None)
        self.ctx.check(rc, ok=(SDR_OK, SDR_EVOXEL_OVERFLOW))
        return rc

    def cloud_to_host(self, i):
        """Copies cloud i of the last run to numpy (device -> host)."""
        v = self.views[i]
        n = int(v.n)
        xyz = np.empty((n, 3), np.float32); rgb = np.empty((n, 3), np.uint8)
        idx = np.empty(n, np.uint32); cnt = np.empty(n, np.uint32)
        self.ctx.check(lib().sdr_pipeline_cloud_download(self._h, int(i), xyz.ctypes.data_as(C.c_void_p), rgb.ctypes.data_as(C.c_void_p),
                                                         idx.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p)))
        out = dict(xyz=xyz, rgb=rgb, passthrough=bool(v.passthrough), minb=np.array(v.min_b[:]), divb=np.array(v.div_b[:]))
        if not v.passthrough:
            out.update(idx=idx, cnt=cnt)
        return out

    def cloud_to_host_into(self, i, xyz, rgb, idx, cnt):
        """Copies cloud i of the last run into caller-provided (ideally pinned) numpy buffers; returns the bytes copied."""
        v = self.views[i]
        self.ctx.check(lib().sdr_pipeline_cloud_download(self._h, int(i), xyz.ctypes.data_as(C.c_void_p), rgb.ctypes.data_as(C.c_void_p),
                                                         idx.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p)))
        n = int(v.n)
        return n * (12 + 3) + (0 if v.passthrough else n * 8)

    def clouds_to_host_into(self, n, xyz, rgb, idx, cnt):
        """All n clouds of the last run into caller buffers [n][...] (ideally pinned) with ONE host wait; returns bytes copied."""
        P = C.c_void_p * n
        ptrs = [None if a is None else P(*[a[i].ctypes.data for i in range(n)]) for a in (xyz, rgb, idx, cnt)]
        self.ctx.check(lib().sdr_pipeline_clouds_download(self._h, int(n), *ptrs))
        per_pt = (12 if xyz is not None else 0) + (3 if rgb is not None else 0)
        per_vox = (4 if idx is not None else 0) + (4 if cnt is not None else 0)
        nb = 0
        for i in range(n):
            v = self.views[i]
            nb += int(v.n) * per_pt + (0 if v.passthrough else int(v.n) * per_vox)
        return nb

    def set_sgbm_option(self, name, value):
        self.ctx.check(lib().sdr_pipeline_set_sgbm_option(self._h, name.encode(), int(value)))
        return self

    def clouds_to_host_packed(self, n, xyz, rgb, idx=None, cnt=None):
        """All n clouds of the last run, packed back to back, into caller buffers (ideally pinned; flat arrays big enough for
        the sum of the cloud sizes): one device->host copy per array and one wait. Returns (offsets[n+1], bytes copied)."""
        off = (C.c_int64 * (n + 1))()
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)   # noqa: E731
        self.ctx.check(lib().sdr_pipeline_clouds_download_packed(self._h, ptr(xyz), ptr(rgb), ptr(idx), ptr(cnt), off))
        total = int(off[n])
        per = (12 if xyz is not None else 0) + (3 if rgb is not None else 0) + (4 if idx is not None else 0) + (4 if cnt is not None else 0)
        return np.array(off[:], np.int64), total * per

    def stage_timeline(self, base_event, which=0):
        """[(name, ms since base_event)]: "begin" and the end of every stage of the last run (which=1: the matcher's kernels)."""
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = lib().sdr_pipeline_stage_timeline(self._h, C.c_void_p(int(base_event)), int(which), 64, names, ms)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def set_profiling(self, on=True):
        lib().sdr_pipeline_set_profiling(self._h, int(on))

    def stage_times(self):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = lib().sdr_pipeline_stage_times(self._h, 64, names, ms)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def close(self):
        if self._h:
            lib().sdr_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StreamStats(C.Structure):
    _fields_ = [("frames", C.c_int64), ("seconds", C.c_double), ("frames_per_s", C.c_double), ("per_device", C.c_int64 * 16),
                ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


def bind_host_to_device(device):
    """Bind this thread (and threads / pinned allocations made later) to the GPU's NUMA node; -1 if the topology is hidden."""
    return int(lib().sdr_bind_host_to_device(int(device)))


class Stream:
    """Persistent frame-parallel runner (sdr_stream_*): native worker threads, one per (device, handle)."""

    def __init__(self, devices, handles_per_device, calib, sgbm_params, voxel_opt, handle_missing=True, batch=8, download_clouds=True):
        self.opt = PipelineOptions(sgbm_params, voxel_opt, int(bool(handle_missing)), int(batch), 0)
        self._h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        rc = lib().sdr_stream_create(devs, len(devices), int(handles_per_device), C.byref(calib), C.byref(self.opt),
                                     int(bool(download_clouds)), C.byref(self._h))
        if rc != SDR_OK:
            raise SdrError(rc, "sdr_stream_create failed")

    def set_sgbm_option(self, name, value):
        rc = lib().sdr_stream_set_sgbm_option(self._h, name.encode(), int(value))
        if rc != SDR_OK:
            raise SdrError(rc, "sdr_stream_set_sgbm_option failed")

    def set_input_format(self, fmt):
        rc = lib().sdr_stream_set_input_format(self._h, int(fmt))
        if rc != SDR_OK:
            raise SdrError(rc, "sdr_stream_set_input_format failed")

    def process_pool(self, n_frames, pool, mem=None):
        """pool: [P][H][2W][3] uint8 (or [P][H*3/2][2W] NV12 / [P][H][2W][2] YUYV after set_input_format), numpy (host /
        pinned) or torch cuda tensor; frame i of the stream = pool[i mod P]."""
        if _is_torch(pool):
            ptr, stride, step = pool.data_ptr(), pool.stride(0), pool.stride(1)
            mem = MEM_DEVICE if mem is None else mem
        else:
            ptr, stride, step = pool.ctypes.data, pool.strides[0], pool.strides[1]
            mem = MEM_PINNED if mem is None else mem
        st = StreamStats()
        rc = lib().sdr_stream_process_pool(self._h, C.c_int64(n_frames), C.c_void_p(ptr), C.c_int64(len(pool)), C.c_size_t(stride),
                                           C.c_size_t(step), int(mem), C.byref(st))
        if rc != SDR_OK:
            raise SdrError(rc, "sdr_stream_process_pool failed")
        return st

    def process(self, n_frames, frames, on_result=None, on_cloud=None):
        """frames: list of SBS numpy arrays (cycled). on_result(index, device, disp, n_cloud_points) / on_cloud(index, device,
        xyz, rgb) are called from worker threads with views that are valid during the call."""
        imgs = [as_image(f, 3)[0] for f in frames]

        def src(_u, i, out):
            out[0] = imgs[i % len(imgs)]
            return SDR_OK

        def snk(_u, i, dev, disp, cloud):
            d = disp[0]
            arr = np.ctypeslib.as_array(C.cast(d.data, C.POINTER(C.c_int16)), shape=(d.rows, d.cols))
            on_result(int(i), int(dev), arr, int(cloud[0].n))

        def csnk(_u, i, dev, n, xyz, rgb):
            n = int(n)
            x = np.ctypeslib.as_array(xyz, shape=(n, 3)) if n else np.empty((0, 3), np.float32)
            c = np.ctypeslib.as_array(rgb, shape=(n, 3)) if n else np.empty((0, 3), np.uint8)
            on_cloud(int(i), int(dev), x, c)

        cb_src = FRAME_SOURCE(src)
        cb_snk = FRAME_SINK(snk) if on_result is not None else C.cast(None, FRAME_SINK)
        cb_c = CLOUD_SINK(csnk) if on_cloud is not None else C.cast(None, CLOUD_SINK)
        st = StreamStats()
        rc = lib().sdr_stream_process(self._h, C.c_int64(n_frames), cb_src, cb_snk, cb_c, None, C.byref(st))
        if rc != SDR_OK:
            raise SdrError(rc, "sdr_stream_process failed")
        return st

    def close(self):
        if self._h:
            lib().sdr_stream_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


FRAME_SOURCE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.POINTER(Image))
FRAME_SINK = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int, C.POINTER(Image), C.POINTER(CloudView))
CLOUD_SINK = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_uint8))


def stream_run(devices, handles_per_device, calib, sgbm_params, voxel_opt, frames, n_frames=None, handle_missing=True, batch=4,
               on_result=None):
    """Frame-parallel run over several GPUs in one process (sdr_stream_run). frames: list of SBS numpy arrays (cycled when
    n_frames exceeds its length). on_result(index, device, disp_numpy_view, n_cloud_points) is called from worker threads."""
    n_frames = len(frames) if n_frames is None else n_frames
    imgs = [as_image(f, 3)[0] for f in frames]

    def src(_u, i, out):
        out[0] = imgs[i % len(imgs)]
        return SDR_OK

    def snk(_u, i, dev, disp, cloud):
        if on_result is not None:
            d = disp[0]
            arr = np.ctypeslib.as_array(C.cast(d.data, C.POINTER(C.c_int16)), shape=(d.rows, d.cols))
            on_result(int(i), int(dev), arr, int(cloud[0].n))

    opt = PipelineOptions(sgbm_params, voxel_opt, int(bool(handle_missing)), int(batch), 0)
    devs = (C.c_int * len(devices))(*devices)
    st = StreamStats()
    cb_src, cb_snk = FRAME_SOURCE(src), FRAME_SINK(snk)
    rc = lib().sdr_stream_run(devs, len(devices), int(handles_per_device), C.byref(calib), C.byref(opt), C.c_int64(n_frames), cb_src, cb_snk,
                              None, C.byref(st))
    if rc != SDR_OK:
        raise SdrError(rc, "sdr_stream_run failed")
    return st
