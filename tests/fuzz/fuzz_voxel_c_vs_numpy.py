"""CPU: the two independent restatements of pcl::VoxelGrid (oracle/sdr_oracle.c and tests/voxelgrid_numpy.py) against each other on
random clouds -- NaN / inf, duplicates, points on voxel faces, negative coordinates, leaves down to the int32-guard passthrough and
the regime where the guard wraps and coordinates leave the int range. usage: python tests/fuzz/fuzz_voxel_c_vs_numpy.py [cases] [seed]"""
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orc
from voxelgrid_numpy import voxelgrid_np

warnings.filterwarnings("ignore")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
fails = passth = 0
t0 = time.time()
for case in range(N):
    n = int(rng.integers(1, 4000))
    mode = int(rng.integers(0, 5))
    if mode == 0:
        pts = rng.normal(0, 1.0, size=(n, 3))
    elif mode == 1:
        pts = rng.integers(-40, 40, size=(n, 3)) * 0.05
    elif mode == 2:
        pts = rng.uniform(-5000, 5000, size=(n, 3))
    elif mode == 3:
        pts = np.stack([rng.uniform(-3000, 3000, n), rng.uniform(-2000, 2000, n), rng.choice([10000.0, 500.0, 2500.0], n) + rng.uniform(0, 50, n)], -1)
    else:
        pts = rng.uniform(-3e5, 3e5, size=(n, 3))
    pts = pts.astype(np.float32)
    pts[rng.random(n) < 0.1] = rng.choice([np.nan, np.inf, -np.inf])
    leaf = float(rng.choice([0.05, 0.1, 0.5, 5.0, 20.0, 0.5, 5.0, 0.001, 3e-4, 1e-5]))
    col = rng.integers(0, 256, size=(n, 3)).astype(np.uint8) if rng.random() < 0.7 else None
    drop = rng.random(n) < 0.2 if rng.random() < 0.5 else None
    with np.errstate(all="ignore"):
        a = orc.voxelgrid(pts.reshape(1, n, 3), None if col is None else col.reshape(1, n, 3), leaf, drop_mask=drop)
        b = voxelgrid_np(pts, col, leaf, drop_mask=drop)
    if a is None or b is None:
        passth += 1
        if (a is None) != (b is None):
            fails += 1
            print("MISMATCH case %d mode %d leaf %g: passthrough C %s numpy %s" % (case, mode, leaf, a is None, b is None))
        continue
    ok = np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["cnt"], b["cnt"]) and np.array_equal(a["xyz"].view(np.uint32), b["xyz"].view(np.uint32)) \
        and (col is None or np.array_equal(a["rgb"], b["rgb"])) and np.array_equal(a["minb"], b["minb"]) and np.array_equal(a["divb"], b["divb"])
    if not ok:
        fails += 1
        print("MISMATCH case %d mode %d leaf %g n %d: %d vs %d voxels" % (case, mode, leaf, n, len(a["idx"]), len(b["idx"])))
print("fuzz_voxel_c_vs_numpy: %d cases, seed %d, %d passthrough, %d failures, %.0f s" % (N, seed, passth, fails, time.time() - t0))
sys.exit(1 if fails else 0)
