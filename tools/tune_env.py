"""Run bench.py under a list of kernel-knob settings (sdr_sgbm_set_option names) and print one compact line per setting:
value, e2e and the per-kernel stage times. Usage: python tools/tune_env.py "h_k=2 h_stages=3" "cf_cols=32" ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
extra = os.environ.get("TUNE_BENCH_ARGS", "--steps 30 --warmup 3 --no-cpu-baseline").split()
for setting in sys.argv[1:] or [""]:
    opts = []
    for kv in setting.split():
        opts += ["--sgbm-opt", kv]
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra + opts, capture_output=True, text=True)
    try:
        j = json.loads(p.stdout.strip().splitlines()[-1])
        st = j["roofline"]["stage_ms_per_batch"]
        print("%-40s value %7.1f e2e %7.1f | %s" % (setting or "(default)", j["value"], j["e2e"]["value"],
                                                     " ".join("%s=%.3f" % (k[:12], v) for k, v in st.items())), flush=True)
    except Exception as e:  # noqa: BLE001
        print("%-40s FAILED rc=%d %s\n%s" % (setting, p.returncode, e, p.stderr[-800:]), flush=True)
