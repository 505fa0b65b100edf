"""B200-native stereo depth engine: drop-in for the rectify -> SGBM -> reproject -> VoxelGrid hot path of
Amar-Aliaga/Stereo_Depth_Ruler. CUDA (sm_100a) behind the C ABI of include/sdr_c.h; no CPU fallback."""
from . import capi, synth  # noqa: F401
from .stereo_vision import StereoConfiguration, StereoDisparity, StereoRectifier, StereoSGBM, VoxelGrid  # noqa: F401

__version__ = "0.1.0"
