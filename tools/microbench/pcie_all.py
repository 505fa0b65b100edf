"""Host <-> device copy bandwidth of all GPUs of the box at once (pinned memory, H2D and D2H on separate streams), to tell a
host-feed limit from a library limit in the end-to-end benchmark: python tools/microbench/pcie_all.py [n_gpus] [seconds]"""
import sys
import threading
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
MB = 64
res = {}


def worker(dev, mode, barrier):
    torch.cuda.set_device(dev)
    h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    barrier.wait()
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < secs:
        for _ in range(4):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        reps += 4
    dt = time.perf_counter() - t0
    res[(dev, mode)] = reps * MB / 1024 / dt      # GB/s per direction


for mode in ("h2d", "d2h", "both"):
    for k in sorted({1, min(2, n), min(4, n), n}):
        res.clear()
        bar = threading.Barrier(k)
        th = [threading.Thread(target=worker, args=(d, mode, bar)) for d in range(k)]
        [t.start() for t in th]
        [t.join() for t in th]
        per = [res[(d, mode)] for d in range(k)]
        tot = sum(per) * (2 if mode == "both" else 1)
        print("%-4s %d GPU(s): per GPU %s GB/s per direction, aggregate %.1f GB/s" % (mode, k, " ".join("%.1f" % v for v in per), tot), flush=True)
