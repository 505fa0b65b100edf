"""CPU: a short run of the randomised pin of the oracle against cv2 (tests/fuzz/fuzz_oracle_vs_cv2.py; the long runs are recorded
under profiles/fuzz/). Skipped when cv2 is not importable."""
import os
import subprocess
import sys

import pytest

pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_equals_cv2_on_random_inputs():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz", "fuzz_oracle_vs_cv2.py"), "120", "201"], capture_output=True, text=True,
                       timeout=600)
    tail = "\\n".join((r.stdout + r.stderr).strip().splitlines()[-10:])
    assert r.returncode == 0 and "failures none" in r.stdout, tail


def test_voxelgrid_restatements_agree_on_random_clouds():
    """The C and the numpy restatement of pcl::VoxelGrid (written independently) on random clouds, absurd leaves included."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz", "fuzz_voxel_c_vs_numpy.py"), "150", "202"], capture_output=True, text=True,
                       timeout=600)
    tail = "\n".join((r.stdout + r.stderr).strip().splitlines()[-10:])
    assert r.returncode == 0 and " 0 failures" in r.stdout, tail
