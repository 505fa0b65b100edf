# same-box A/B of library builds through the whole benchmark: bash tools/probes/ab_bench_libs.sh ab/lib_a.so ab/lib_b.so ...
KEEP=/tmp/libsdr_keep.so
cp stereo_depth_ruler_b200/libsdr_b200.so $KEEP
for rep in 1 2; do for l in "$@"; do cp $l stereo_depth_ruler_b200/libsdr_b200.so
python bench.py --steps 20 --no-cpu-baseline --no-parity-gate --no-ingest-leg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%-24s value %.0f (%.3f ms/step) e2e %.0f' % ('$l', d['value'], d['ms_per_step'], d['e2e']['value']))"; done; done
cp $KEEP stereo_depth_ruler_b200/libsdr_b200.so
