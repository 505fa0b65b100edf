"""Golden vectors for the frame ingest (sdr_ingest_to_bgr): cv2 4.13.0's own cvtColor on NV12 / YUYV pictures.

    python tests/golden/make_ingest_golden.py        (run in the build container; writes tests/golden/ingest_golden.npz)

Contents: random NV12 and YUYV pictures with cv2's BGR, a structured sweep that visits every Y with a lattice of (U, V)
(the clamps at 0 / 255 and the Y < 16 floor included), and the sha256 of cv2's answer for ALL 2^24 (Y, U, V) triples.
"""
import hashlib
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def all_triples_bgr():
    """cv2 on every (Y, U, V): one NV12 picture per U, 512 x 512 (a pixel pair per Y across, a row pair per V down)."""
    h = hashlib.sha256()
    Y = np.repeat(np.repeat(np.arange(256, dtype=np.uint8), 2)[None, :], 512, 0)          # [512][512]
    for u in range(256):
        uv = np.empty((256, 512), np.uint8)
        uv[:, 0::2] = u
        uv[:, 1::2] = np.arange(256, dtype=np.uint8)[:, None]
        h.update(cv2.cvtColor(np.concatenate([Y, uv], 0), cv2.COLOR_YUV2BGR_NV12).tobytes())
    return h.hexdigest()


def main():
    rng = np.random.default_rng(7)
    nv = rng.integers(0, 256, (36, 40), dtype=np.uint8)                 # 40 x 24 picture
    yu = rng.integers(0, 256, (24, 40, 2), dtype=np.uint8)
    lat = np.array([0, 1, 15, 16, 17, 64, 127, 128, 129, 200, 239, 240, 254, 255], np.uint8)
    Yl = np.repeat(np.arange(256, dtype=np.uint8), 2)[None, :].repeat(2 * len(lat), 0)   # [28][512]
    uvl = np.empty((len(lat), 512), np.uint8)
    sweeps = []
    for u in lat:
        uvl[:, 0::2] = u
        uvl[:, 1::2] = lat[:, None]
        sweeps.append(np.concatenate([Yl, uvl], 0))
    sweep = np.stack(sweeps)                                               # [14][42][512] NV12
    np.savez_compressed(
        os.path.join(HERE, "ingest_golden.npz"),
        nv12=nv, nv12_bgr=cv2.cvtColor(nv, cv2.COLOR_YUV2BGR_NV12),
        yuyv=yu, yuyv_bgr=cv2.cvtColor(yu, cv2.COLOR_YUV2BGR_YUY2),
        sweep=sweep, sweep_bgr=np.stack([cv2.cvtColor(s, cv2.COLOR_YUV2BGR_NV12) for s in sweep]),
        all_triples_sha256=np.array(all_triples_bgr()), cv2_version=np.array(cv2.__version__))
    print("wrote ingest_golden.npz with cv2", cv2.__version__)


if __name__ == "__main__":
    main()
