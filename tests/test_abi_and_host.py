"""CPU: the C-ABI library loads and exports every symbol include/sdr_c.h declares; host-side logic that needs
no GPU (YAML reader / writer, parameter plumbing, the interface mirror's error behaviour). No compute calls."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "sdr_c.h")).read()
    return sorted(set(re.findall(r"SDR_API\s+[\w\s\*]+?\b(sdr_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from stereo_depth_ruler_b200 import capi
    L = capi.lib()
    syms = header_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == syms


def test_no_cpu_fallback_without_device():
    from stereo_depth_ruler_b200 import capi
    if capi.lib().sdr_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SdrError) as e:
        capi.Context(0)
    assert e.value.status == capi.SDR_ECUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "stereo_depth_ruler_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import\s+orc\b|libsdr_oracle|orc_\w+\(", txt, re.M), f
    out = subprocess.run(["nm", "-D", os.path.join(pkg, "libsdr_b200.so")], capture_output=True, text=True).stdout
    assert "orc_" not in out


def test_yaml_roundtrip(tmp_path, calib_path):
    from stereo_depth_ruler_b200 import capi
    c = capi.load_calibration(calib_path)
    assert (c.image_width, c.image_height) == (1280, 720) and c.present == 0x1FFF
    assert c.mat("Q")[3, 2] == pytest.approx(0.00832541998100415)
    p = str(tmp_path / "out.yaml")
    capi.save_calibration(p, c)
    c2 = capi.load_calibration(p)
    for k in ("K_left", "dist_left", "K_right", "dist_right", "R", "T", "E", "F", "R1", "R2", "P1", "P2", "Q"):
        assert np.array_equal(c.mat(k), c2.mat(k)), k


def test_yaml_opencv_dialect_multiline_and_errors(tmp_path):
    from stereo_depth_ruler_b200 import capi, StereoConfiguration
    p = tmp_path / "c.yaml"
    mat = lambda n, r, c, d: "%s: !!opencv-matrix\n   rows: %d\n   cols: %d\n   dt: d\n   data: [ %s ]\n" % (
        n, r, c, ",\n       ".join(", ".join(repr(float(v)) for v in d[i:i + 3]) for i in range(0, len(d), 3)))
    body = "%YAML:1.0\n---\nimageWidth: 64\nimageHeight: 48\n"
    eye = [1, 0, 0, 0, 1, 0, 0, 0, 1]
    for n in ("cameraMatrixLeft", "cameraMatrixRight", "R", "R1", "R2"):
        body += mat(n, 3, 3, eye)
    body += mat("distCoeffsLeft", 1, 5, [0.1, -0.2, 0, 0, 0.3]) + mat("distCoeffsRight", 1, 5, [0] * 5)
    body += mat("T", 3, 1, [-120, 0, 0]) + mat("P1", 3, 4, [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0])
    body += mat("P2", 3, 4, [1, 0, 0, -5, 0, 1, 0, 0, 0, 0, 1, 0]) + mat("Q", 4, 4, list(range(16)))
    p.write_text(body)
    cfg = StereoConfiguration()
    assert cfg.loadFromFile(str(p)) and cfg.isValid() and cfg.imageSize == (64, 48)
    assert cfg.distCoeffsLeft.tolist() == [[0.1, -0.2, 0, 0, 0.3]] and cfg.Q[3, 3] == 15 and cfg.T[0, 0] == -120
    # invalid size (stereo_configuration.cpp:15-19) and missing essential matrix (:36-41) -> False
    (tmp_path / "bad1.yaml").write_text(body.replace("imageWidth: 64", "imageWidth: 0"))
    assert not StereoConfiguration().loadFromFile(str(tmp_path / "bad1.yaml"))
    (tmp_path / "bad2.yaml").write_text(body.replace("Q: !!opencv-matrix", "Qx: !!opencv-matrix"))
    assert not StereoConfiguration().loadFromFile(str(tmp_path / "bad2.yaml"))
    assert not StereoConfiguration().loadFromFile(str(tmp_path / "missing.yaml"))
    with pytest.raises(capi.SdrError):
        capi.load_calibration(str(tmp_path / "missing.yaml"))


def test_param_struct_layout_matches_reference_order():
    from stereo_depth_ruler_b200 import capi
    p = capi.make_params(0, 80, 5, 8 * 5 * 5 * 3, 32 * 5 * 5 * 3, 1, 63, 12, 200, 2, capi.MODE_SGBM_3WAY)  # stereo_disparity.cpp:5-6
    assert [getattr(p, n) for n, _ in p._fields_] == [0, 80, 5, 600, 2400, 1, 63, 12, 200, 2, 2]
    assert ctypes.sizeof(p) == 44 and ctypes.sizeof(capi.Image) == 40


def test_cpp_facade_compiles_and_links(tmp_path):
    """include/sdr/stereo_vision.hpp (the C++ host-side mirror) builds against the C ABI with plain g++."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "sdr/stereo_vision.hpp"\n'
                   'int main(int argc, char** argv) {\n'
                   '  sdr::StereoConfiguration cfg;\n'
                   '  if (argc < 2 || !cfg.loadFromFile(argv[1])) return 2;\n'
                   '  if (!cfg.isValid() || cfg.imageSize.width != 1280) return 3;\n'
                   '  try { sdr::Context ctx(0); } catch (const sdr::Error& e) { return e.status() == SDR_ECUDA ? 0 : 4; }\n'
                   '  return 0; }\n')
    exe = tmp_path / "t"
    pkg = os.path.join(ROOT, "stereo_depth_ruler_b200")
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", pkg, "-lsdr_b200", "-Wl,-rpath," + pkg])
    rc = subprocess.run([str(exe), os.path.join(ROOT, "tests", "golden", "zed2_calibration.yaml")]).returncode
    assert rc == 0
