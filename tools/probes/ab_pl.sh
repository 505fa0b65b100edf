cp ab/lib_pl32bt4.so stereo_depth_ruler_b200/libsdr_b200.so
timeout 600 python -m pytest tests/test_gpu_sgbm.py tests/test_gpu_small_shapes.py -x -q -m gpu 2>&1 | tail -3
cp ab/lib_base2.so stereo_depth_ruler_b200/libsdr_b200.so
bash tools/ab_libs.sh ab/lib_base2.so ab/lib_pl32.so ab/lib_pl32bt4.so
for v in base2 pl32bt4; do
  cp ab/lib_$v.so stereo_depth_ruler_b200/libsdr_b200.so
  echo "== one frame $v"; timeout 120 python tools/profile_sgbm.py 1 0 3 | tail -1
  echo "== 3way $v"; timeout 120 python tools/profile_sgbm.py 8 2 3 | tail -1
done
