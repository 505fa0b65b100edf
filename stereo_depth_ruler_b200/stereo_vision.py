"""Host-side mirror of the reference's stereo_vision / point_cloud interface over the C ABI.

Same class names, method names, argument meaning and error behaviour as
  StereoConfiguration  stereo_vision/include/stereo_configuration.hpp:6-16
  StereoRectifier      stereo_vision/include/stereo_rectifier.hpp:7-19
  StereoDisparity      stereo_vision/include/stereo_disparity.hpp:8-24
  convertCVMatToPCL + pcl::VoxelGrid usage   point_cloud/src/pcd_write.cpp:17-51, :123-130
so the parity tests read like tests of the reference. Images are numpy arrays (host) or torch CUDA
tensors (device resident). Everything computes on the GPU through libsdr_b200.so; there is no CPU path.
The C++ facade with the same names is include/sdr/stereo_vision.hpp.
"""
import sys

import numpy as np

from . import capi

_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = capi.Context(0)
    return _default_ctx


class StereoConfiguration:
    """POD of the calibration matrices + imageSize; YAML through the library's own OpenCV-dialect reader."""

    _FIELDS = {"cameraMatrixLeft": "K_left", "distCoeffsLeft": "dist_left", "cameraMatrixRight": "K_right",
               "distCoeffsRight": "dist_right", "R": "R", "T": "T", "E": "E", "F": "F", "R1": "R1", "R2": "R2",
               "P1": "P1", "P2": "P2", "Q": "Q"}

    def __init__(self):
        for k in self._FIELDS:
            setattr(self, k, None)
        self.imageSize = (0, 0)
        self._c = None

    def loadFromFile(self, filename):
        """stereo_configuration.cpp:4-46: False (and a message on stderr) when the file cannot be opened,
        the image size is invalid or an essential matrix is missing."""
        try:
            c = capi.load_calibration(filename)
        except capi.SdrError:
            sys.stderr.write("Error: Failed to load stereo configuration from %s\n" % filename)
            return False
        self._c = c
        self.imageSize = (c.image_width, c.image_height)
        for k, f in self._FIELDS.items():
            setattr(self, k, c.mat(f))
        return True

    def saveToFile(self, filename):
        try:
            capi.save_calibration(filename, self.c_struct())
        except capi.SdrError:
            return False
        return True

    def isValid(self):
        """stereo_configuration.cpp:77-80."""
        return all(getattr(self, k) is not None and np.size(getattr(self, k)) > 0
                   for k in ("cameraMatrixLeft", "cameraMatrixRight", "R", "T", "Q"))

    def c_struct(self):
        c = capi.Calibration()
        c.image_width, c.image_height = int(self.imageSize[0]), int(self.imageSize[1])
        present = 0
        for bit, (k, f) in enumerate(self._FIELDS.items()):
            m = getattr(self, k)
            if m is None:
                continue
            flat = np.asarray(m, np.float64).ravel()
            arr = getattr(c, f)
            for i in range(len(arr)):
                arr[i] = float(flat[i])
            present |= 1 << bit
        c.present = present
        return c


class StereoRectifier:
    """stereo_rectifier.cpp:6-14 (maps are evaluated inside the kernel, nothing is precomputed) and :17-41."""

    def __init__(self, config, ctx=None):
        self.config = config
        self.ctx = ctx or default_context()
        self._r = capi.Rectifier(self.ctx, config.c_struct())

    def rectify(self, left_src, right_src, want_gray=False):
        if left_src is None or right_src is None or left_src.size == 0 or right_src.size == 0:
            sys.stderr.write("Error: Input images are empty\n")  # stereo_rectifier.cpp:18-21
            return None
        return self._r.rectify(left_src, right_src, want_gray=want_gray)

    def maps(self, eye):
        return self._r.maps(eye)


class StereoSGBM:
    """cv::StereoSGBM subset used by the reference (create / compute / getters)."""

    MODE_SGBM, MODE_HH, MODE_SGBM_3WAY = 0, 1, 2

    def __init__(self, ctx, params):
        self._m = capi.Sgbm(ctx, params)
        self.params = params

    @staticmethod
    def create(minDisparity=0, numDisparities=16, blockSize=3, P1=0, P2=0, disp12MaxDiff=0, preFilterCap=0,
               uniquenessRatio=0, speckleWindowSize=0, speckleRange=0, mode=0, ctx=None):
        return StereoSGBM(ctx or default_context(),
                          capi.make_params(minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap,
                                           uniquenessRatio, speckleWindowSize, speckleRange, mode))

    def compute(self, left, right, out=None):
        return self._m.compute(left, right, out)

    def getNumDisparities(self):
        return self.params.numDisparities

    def getMinDisparity(self):
        return self.params.minDisparity


class StereoDisparity:
    """stereo_disparity.cpp:4-14 (matcher creation) / :17-39 / :76-80.

    computeDisparity follows the reference's un-filtered path (pcd_write.cpp:88-112: BGR2GRAY ->
    StereoSGBM::compute -> convertTo(CV_32F, 1/16)). The half-resolution right-matcher + WLS filter of
    stereo_disparity.cpp:22-36 needs opencv_contrib ximgproc, for which no oracle exists here
    (SURVEY.md 8(f) rank 1): it is out of scope and not emulated.
    """

    def __init__(self, Q_matrix, ctx=None, params=None):
        self.ctx = ctx or default_context()
        self.Q = np.asarray(Q_matrix, np.float64).reshape(4, 4)
        # the reference's tuple (stereo_disparity.cpp:5-6)
        self.matcher = StereoSGBM(self.ctx, params or capi.make_params(0, 80, 5, 8 * 5 * 5 * 3, 32 * 5 * 5 * 3, 1, 63, 12, 200, 2,
                                                                      capi.MODE_SGBM_3WAY))

    def computeDisparity16(self, left, right):
        if left.ndim == 3:
            left = capi.bgr2gray(self.ctx, left)
        if right.ndim == 3:
            right = capi.bgr2gray(self.ctx, right)
        return self.matcher.compute(left, right)

    def computeDisparity(self, left, right):
        return capi.disp_to_float(self.ctx, self.computeDisparity16(left, right))   # convertTo(CV_32F, 1/16) on the device

    def computeDepth(self, disparity, handleMissingValues=False):
        """disparity: CV_32F (as the reference passes) or CV_16S (the /16 is then fused)."""
        return capi.reproject_to_3d(self.ctx, disparity, self.Q, handleMissingValues)

    def get_matcher(self):
        return self.matcher


class VoxelGrid:
    """pcl::VoxelGrid<PointXYZRGB> as used at pcd_write.cpp:123-130 (setLeafSize + filter), fed straight from the
    OpenCV point image and the colour image (convertCVMatToPCL, pcd_write.cpp:17-51)."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self._c = capi.Cloud(self.ctx)
        self.leaf = (0.005, 0.005, 0.005)
        self.drop_invalid = False
        self.z_max = 0.0
        self.min_disparity = 0

    def setLeafSize(self, lx, ly, lz):
        self.leaf = (lx, ly, lz)

    def filter(self, pointCloud_CV, colorImage_CV=None, disp16=None):
        opt = capi.voxel_options(self.leaf, self.drop_invalid, self.z_max, self.min_disparity)
        return self._c.voxelgrid(pointCloud_CV, colorImage_CV, opt, disp16=disp16)

    def filter_from_disparity(self, disp16, colorImage_CV, Q, handleMissingValues=False):
        opt = capi.voxel_options(self.leaf, self.drop_invalid, self.z_max, self.min_disparity)
        return self._c.depth_cloud(disp16, colorImage_CV, Q, opt, handleMissingValues)
