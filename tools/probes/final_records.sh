set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; tail -3 gpurun_out/r2f_tests.log
python __graft_entry__.py --smoke > gpurun_out/r2f_smoke.log 2>&1; tail -1 gpurun_out/r2f_smoke.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 600 gpurun_out/r2f_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; tail -c 400 gpurun_out/r2f_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-gate --no-ingest-leg > gpurun_out/r2f_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r2f_launches.csv > gpurun_out/r2f_launches_summary.txt; head -12 gpurun_out/r2f_launches_summary.txt
timeout 200 python tools/probes/voxel_time.py 2>&1 | tail -1
