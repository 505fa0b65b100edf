set -x
for v in il2 il3; do
  cp ab/lib_$v.so stereo_depth_ruler_b200/libsdr_b200.so
  timeout 300 python -m pytest tests/test_gpu_sgbm.py -x -q -k "fused_cost_sweep or headline or strip_widths_small or golden_vectors" 2>&1 | tail -3
done
cp ab/lib_base.so stereo_depth_ruler_b200/libsdr_b200.so
bash tools/ab_libs.sh ab/lib_base.so ab/lib_il2.so ab/lib_il3.so
for v in base il2 il3; do
  cp ab/lib_$v.so stereo_depth_ruler_b200/libsdr_b200.so
  echo "== one frame $v"; timeout 120 python tools/profile_sgbm.py 1 0 3 | tail -1
  echo "== HH $v"; timeout 120 python tools/profile_sgbm.py 8 1 3 | tail -1
done
