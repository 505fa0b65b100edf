"""GPU: short runs of the randomised differential fuzzers (tests/fuzz/fuzz_*.py) -- each in its own process, because a kernel fault
would poison the CUDA context of the test session. The long runs of the same tools are recorded under profiles/fuzz/."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool,args", [("fuzz_sgbm.py", ["80", "101"]), ("fuzz_sgbm.py", ["25", "102", "300", "400"]),
                                       ("fuzz_stages.py", ["40", "103"]), ("fuzz_pipeline.py", ["16", "104"]),
                                       ("fuzz_split.py", ["16", "105"]), ("fuzz_api.py", ["30", "106"]), ("fuzz_stream.py", ["10", "107"])])
def test_fuzzer_finds_nothing(tool, args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz", tool)] + args, capture_output=True, text=True, timeout=900)
    tail = "\n".join((r.stdout + r.stderr).strip().splitlines()[-12:])
    assert r.returncode == 0, tail
    assert "0 failures" in r.stdout or "failures none" in r.stdout, tail
