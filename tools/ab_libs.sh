#!/bin/bash
# A/B several builds of the library on one GPU box: tools/ab_libs.sh ab/lib_x.so ab/lib_y.so ... (restores the first at the end)
set -e
KEEP=/tmp/libsdr_keep.so
cp stereo_depth_ruler_b200/libsdr_b200.so $KEEP
for rep in 1 2; do
for l in "$@"; do
  cp $l stereo_depth_ruler_b200/libsdr_b200.so
  echo "== $l"
  timeout 120 python tools/profile_sgbm.py 8 0 3 | tail -1
done
done
cp $KEEP stereo_depth_ruler_b200/libsdr_b200.so
