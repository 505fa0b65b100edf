#!/bin/bash
# ncu --set full capture of the SGBM kernels of one batched compute (after the same command ran clean without ncu).
# usage: tools/profile_sgbm_ncu.sh <tag> [kernel-regex] [option=value ...]   (options: sdr_sgbm_set_option knobs, e.g. wave=1)
set -e
TAG=${1:-r1}
KRE=${2:-"k_cost|k_path"}
shift; shift || true
OPTS="$@"
mkdir -p gpurun_out
python tools/profile_sgbm.py ${B:-8} 0 2 $OPTS > gpurun_out/${TAG}_profile_plain.txt 2>&1
ncu --set full --clock-control none --import-source on -k "regex:${KRE}" --launch-skip 0 --launch-count 8 \
    -o gpurun_out/${TAG}_sgbm -f python tools/profile_sgbm.py ${B:-8} 0 2 $OPTS > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_sgbm.ncu-rep > gpurun_out/${TAG}_ncu_summary.txt 2>&1
