"""Calibration reader for the CPU reference legs (TEST / BASELINE INFRASTRUCTURE, never imported by the product).

Reads config/stereo.yaml-style files the way StereoConfiguration::loadFromFile does
(stereo_vision/src/stereo_configuration.cpp:4-46 of the reference): cv::FileStorage. When cv2 is missing a minimal
parser of the same OpenCV YAML 1.0 dialect takes over. Nothing here touches libsdr_b200.so, so the reference arm of
bench.py never maps the product library.
"""
import re

import numpy as np

_MATS = ("cameraMatrixLeft", "distCoeffsLeft", "cameraMatrixRight", "distCoeffsRight", "R", "T", "E", "F", "R1", "R2", "P1", "P2", "Q")
_SHORT = {"cameraMatrixLeft": "K_left", "distCoeffsLeft": "dist_left", "cameraMatrixRight": "K_right", "distCoeffsRight": "dist_right"}


def load(path):
    """-> dict: image_width, image_height and the matrices under the names the tests use (K_left, dist_left, ..., Q)."""
    out = {}
    try:
        import cv2
        fs = cv2.FileStorage(str(path), cv2.FILE_STORAGE_READ)
        if not fs.isOpened():
            raise IOError("cannot open %s" % path)
        out["image_width"] = int(fs.getNode("imageWidth").real())
        out["image_height"] = int(fs.getNode("imageHeight").real())
        for k in _MATS:
            node = fs.getNode(k)
            if not node.empty():
                out[_SHORT.get(k, k)] = np.asarray(node.mat(), np.float64)
        fs.release()
    except ImportError:
        text = open(path).read()
        out["image_width"] = int(float(re.search(r"^imageWidth:\s*(\S+)", text, re.M).group(1)))
        out["image_height"] = int(float(re.search(r"^imageHeight:\s*(\S+)", text, re.M).group(1)))
        for m in re.finditer(r"^(\w+): !!opencv-matrix\s+rows: (\d+)\s+cols: (\d+)\s+dt: \w+\s+data: \[([^\]]*)\]", text, re.M):
            k, rows, cols = m.group(1), int(m.group(2)), int(m.group(3))
            vals = [float(v) for v in m.group(4).replace("\n", " ").split(",") if v.strip()]
            out[_SHORT.get(k, k)] = np.array(vals, np.float64).reshape(rows, cols)
    if out.get("image_width", 0) <= 0 or out.get("image_height", 0) <= 0:
        raise IOError("invalid image size in %s" % path)          # stereo_configuration.cpp:15-19
    return out
