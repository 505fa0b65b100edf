#!/usr/bin/env python
"""Benchmark of the stereo hot path (BASELINE.json metric: SGBM frames/s and Mpix*disp/s at 1280x720x128).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference] [--workload c2|c4|c5]

Workload c2 (default; config.workload = BASELINE.json configs[1]): the full pipeline per frame -- remap rectification of
both halves of a 2560x720 SBS frame, StereoSGBM-compatible MODE_SGBM (D=128, reference parameter tuple) with LR check,
median and speckle filter, reprojectImageTo3D and the 5 mm VoxelGrid -- on synthetic textured frames. A step = one pass
over a batch of B frames per GPU. `value` is measured with the frames resident in HBM and `e2e` with pinned HOST frames,
both through the native
frame-parallel runner of the C ABI (sdr_stream_*); e2e has the H2D of the frames and the D2H of the disparity
maps and the clouds inside the timed region. After the timed legs rank 0 pushes its timed frames through cv2 (the
OpenCV calls the reference makes) and compares: `parity_gate`; a mismatch exits non-zero.
Workload c4 (configs[3]): the same pipeline over a stream of --frames (4096) SBS frames from 64 unique pairs, sharded
frame-parallel over the GPUs, both legs through the native runner. Workload c5 (configs[4]): one 3840x2160x256
MODE_SGBM_3WAY frame split by stripes over N GPUs in ONE process (sdr_sgbm_split_*, strong scaling; run without torchrun).
`--impl reference` times the reference's CPU arithmetic (cv2 + C VoxelGrid restatement) on all host cores.
"""
import argparse
import glob
import hashlib
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_emit = None

W, H, D = 1280, 720, 128
SGBM_ARGS = (0, D, 5, 8 * 3 * 5 * 5, 32 * 3 * 5 * 5, 1, 63, 12, 200, 2, 0)   # pcd_write.cpp:102-108 with D=128, MODE_SGBM
LEAF_MM, Z_MAX_MM = 5.0, 5000.0
CALIB = os.path.join(ROOT, "tests", "golden", "zed2_calibration.yaml")
MPIX_DISP_PER_FRAME = W * H * D / 1e6
METRIC = "sgbm_full_pipeline_frames_per_s_1280x720x128"
WORKLOADS = {
    "c2": "configs[1]: full pipeline per frame at 1280x720x128 (remap + MODE_SGBM + LR + median + speckle + reproject + 5 mm VoxelGrid)",
    "c4": "configs[3]: synthetic 2560x720 SBS stream, frame pairs sharded frame-parallel, full pipeline per frame at 1280x720x128",
    "c5": "configs[4]: one 3840x2160 pair, numDisparities=256, MODE_SGBM_3WAY, row-strip split across GPUs (peer copies over NVLink)",
}


def base_config(workload):
    """The part of `config` both arms share (same keys, same values)."""
    return {"workload": WORKLOADS[workload], "sgbm": list(SGBM_ARGS), "leaf_mm": LEAF_MM, "z_max_mm": Z_MAX_MM}


def make_frames(n, base=4):
    """n distinct SBS frames from `base` synthetic textured pairs: rolling the rows of a rectified pair gives
    another valid pair (disparity is horizontal), so the pool is cheap to build yet every frame differs."""
    from stereo_depth_ruler_b200 import synth
    pool = []
    for s in range(min(base, n)):
        l, r = synth.stereo_pair(W, H, D, s)
        pool.append(np.concatenate([synth.colorize(l, s), synth.colorize(r, s + 1)], axis=1))
    frames = []
    for i in range(n):
        f = pool[i % len(pool)]
        k = (i // len(pool)) * 37
        frames.append(np.ascontiguousarray(np.roll(f, k, axis=0)) if k else f)
    return frames


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(p[1])); smax = float(p[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:     # timed region shorter than the sampling period: take every sample we have
            for t, line in self.rows:
                p = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(p[1])); smax = float(p[2])
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ---- ncu numbers come from the committed summaries under profiles/, never from literals in this file ----
_STAGE_KERNEL = [   # stage name of sdr_sgbm_stage_times -> kernel-name pattern in the ncu summary
    ("cost_path_v_fused", r"k_cost_path_v3<"), ("path_h_fwd", r"k_path_h2<\d+, 0>|k_path_h<\d+, \d+, 0"), ("path_h_bwd_wta", r"k_path_h<\d+, \d+, 1|k_path_h2<\d+, 1>"),
    ("path_h_both_wta", r"k_path_hh<"), ("prefilter", r"k_prefilter"), ("cost", r"k_cost<"), ("path_v_fused_down", r"k_path_v3<"),
]


def parse_ncu_summary(path):
    """tools/ncu_summary.py output -> [{'name': kernel, metric: float, ...}] (per launch, as captured)."""
    out, cur = [], None
    for line in open(path):
        if line.startswith("Kernel Name"):
            cur = {"name": line[len("Kernel Name"):].strip()}
            out.append(cur)
            continue
        m = re.match(r"^(\S+)\s+([-+0-9.eE]+)\s*(\S*)", line)
        if cur is not None and m and not line.startswith("#"):
            try:
                v = float(m.group(2))
            except ValueError:
                continue
            unit = m.group(3)
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit, 1.0)
            cur[m.group(1)] = v * scale
    return out


def ncu_numbers():
    """DRAM traffic per launch and SM-side utilisation of the SGBM kernels from the newest profiles/r*_ncu_sgbm_kernels.txt."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_sgbm_kernels.txt")),
                   key=lambda p: [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", os.path.basename(p))])
    if not files:
        return None
    path = files[-1]
    rows = parse_ncu_summary(path)
    res = {"file": os.path.relpath(path, ROOT), "sha256": hashlib.sha256(open(path, "rb").read()).hexdigest()[:16], "stages": {}}
    m = re.search(r"batch[ =](\d+)", open(path).read(400))
    res["frames_per_launch"] = int(m.group(1)) if m else 8
    for stage, pat in _STAGE_KERNEL:
        for r in rows:
            if re.search(pat, r["name"]) and "dram__bytes_read.sum" in r:
                res["stages"][stage] = {
                    "traffic": r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"], "ms": r.get("gpu__time_duration.sum"),
                    "issue_active_pct": r.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "alu_pipe_pct": r.get("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"), "kernel": r["name"][:60]}
                break
    return res


def run_reference(args):
    """The reference's CPU arithmetic on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.cpu_pipeline import CpuPipeline
    cores = os.cpu_count() or 1
    cpu = CpuPipeline(CALIB, SGBM_ARGS, LEAF_MM, Z_MAX_MM, cores=cores)
    frames = make_frames(cores)                       # a step = one frame per core (bounded sample of the workload)
    cpu.prepare(frames)
    for _ in range(max(0, args.warmup - 1)):
        cpu.run()
    total = 0.0
    for _ in range(args.steps):
        dt, _ = cpu.run()
        total += dt
    fps = args.steps * len(frames) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "s16", "data": "synthetic",
        "mpix_disp_per_s": fps * MPIX_DISP_PER_FRAME,
        "config": base_config("c2"), "run": {"frames_per_step_per_gpu": len(frames), "frames_per_submission": 1},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": cpu.kind,
                         "sample": "%d steps x %d frames; %s" % (args.steps, len(frames), cpu.describe())},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    cpu.close()
    _emit(line)


def run_c5(args):
    """One 4K x 256 MODE_SGBM_3WAY frame split over N GPUs in one process (strong scaling)."""
    import torch
    from stereo_depth_ruler_b200 import capi, synth
    if int(os.environ.get("RANK", "0")) != 0:
        return                                                    # one process drives every GPU
    n = args.gpus
    if torch.cuda.device_count() < n:
        raise SystemExit("bench.py --workload c5: %d GPUs asked for, %d visible" % (n, torch.cuda.device_count()))
    Wc, Hc, Dc = 3840, 2160, 256
    sargs = (0, Dc, 5, 600, 2400, 1, 63, 12, 200, 2, capi.MODE_SGBM_3WAY)
    torch.cuda.set_device(0)
    pairs = [synth.stereo_pair(Wc, Hc, Dc, 6 + s) for s in range(2)]
    sp = capi.SgbmSplit(list(range(n)), capi.make_params(*sargs))
    # parity first: the maps cv2 gives for the timed inputs
    import cv2
    ref = cv2.StereoSGBM_create(*sargs[:10], mode=cv2.STEREO_SGBM_MODE_SGBM_3WAY)
    parity = all(np.array_equal(sp.compute(l, r), ref.compute(l, r)) for l, r in pairs)
    if not parity:
        _emit({"metric": "sgbm_3way_4k_split_frames_per_s", "parity_gate": "FAIL", "n_gpus": n})
        raise SystemExit(3)
    dev = [(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()) for l, r in pairs]
    host = [(torch.from_numpy(l).pin_memory().numpy(), torch.from_numpy(r).pin_memory().numpy()) for l, r in pairs]
    out_d = torch.empty((Hc, Wc), dtype=torch.int16, device="cuda")
    out_h = torch.empty((Hc, Wc), dtype=torch.int16).pin_memory().numpy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(max(warmup, 3)):
            fn(i)
        sp.synchronize(); torch.cuda.synchronize()
        tot = 0.0
        for i in range(steps):
            flush.zero_()                                           # L2 flush between timed iterations (256 MB > 126 MB L2)
            torch.cuda.synchronize()
            fn(i)
            tot += sp.stats()["ms"]                                 # CUDA events on devices[0]'s stream around the whole call
        return tot / steps

    sampler = ClockSampler(0)
    t0 = time.perf_counter()
    ms = timed(lambda i: sp.compute(dev[i % 2][0], dev[i % 2][1], out_d), args.steps, args.warmup)
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    st_dev = sp.stats()
    ms_e = timed(lambda i: sp.compute(host[i % 2][0], host[i % 2][1], out_h), args.steps, args.warmup)
    st_host = sp.stats()
    fps, fps_e = 1e3 / ms, 1e3 / ms_e
    mpd = Wc * Hc * Dc / 1e6
    _emit({"metric": "sgbm_3way_4k_split_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": n, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "s16", "data": "synthetic",
           "mpix_disp_per_s": fps * mpd, "parity_gate": "pass",
           "config": {"workload": WORKLOADS["c5"], "sgbm": list(sargs), "l2": "256 MB flush between timed iterations",
                      "split": "stripes -> GPUs in contiguous blocks; device 0 assembles and post-filters"},
           "e2e": {"value": fps_e, "unit": "frames/s", "ms_per_step": ms_e, "h2d_bytes_per_step": int(sum(st_host["input_bytes"])),
                   "d2h_bytes_per_step": Wc * Hc * 2},
           "split": {"device_input": st_dev, "host_input": st_host}, "clocks": clocks, "gpu_launches": None})
    sp.close()


def run_ours(args):
    if args.workload == "c5":
        return run_c5(args)
    import torch
    import torch.distributed as dist
    from stereo_depth_ruler_b200 import capi, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    # host threads and pinned buffers on the GPU's own socket (no-op when the container hides the topology)
    numa_node = capi.bind_host_to_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch                                                # frames per submission (one launch of every kernel)
    FPS = max(args.frames_per_step, B)                            # frames per step per GPU
    FPS -= FPS % B
    c4 = args.workload == "c4"
    sgbm_opts = [kv.split("=") for kv in args.sgbm_opt]

    # ---- CPU legs first (rank 0): baseline timing at N=1, and the parity gate's reference maps at every N ----
    n_pool = max(64 if c4 else args.pool, 2 * FPS)
    n_pool -= n_pool % B                                         # whole batches: consecutive frames are consecutive in memory
    frames_np = make_frames(n_pool, base=64 if c4 else 4)   # c4: 64 unique pairs cycled
    mine = [frames_np[i % n_pool] for i in sharding.frames_for_rank(n_pool * world, rank, world)]
    cpu_baseline, gate_ref = None, None
    n_gate = min(len(mine), args.gate_frames)
    if rank == 0 and (not args.no_cpu_baseline or not args.no_parity_gate):
        from oracle.cpu_pipeline import CpuPipeline
        cores = os.cpu_count() or 1
        cpu = CpuPipeline(CALIB, SGBM_ARGS, LEAF_MM, Z_MAX_MM, cores=cores)
        if world == 1 and not args.no_cpu_baseline:
            cframes = make_frames(cores)
            cpu.prepare(cframes)
            reps, total = 0, 0.0
            while total < 8.0 and reps < 8:               # bounded: about 10-20 s of CPU work in all
                dt, _ = cpu.run()
                total += dt
                reps += 1
            cpu_baseline = {"value": reps * len(cframes) / total, "unit": "frames/s", "cores": cores, "kind": cpu.kind,
                            "sample": "%d passes x %d frames of the same workload; %s" % (reps, len(cframes), cpu.describe())}
        if not args.no_parity_gate:
            gate_ref = cpu.maps_and_counts(mine[:n_gate])   # cv2 on the frames the timed legs run on
        cpu.close()

    ctx = capi.Context(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    calib = capi.load_calibration(CALIB)
    params = capi.make_params(*SGBM_ARGS)
    vox = capi.voxel_options(LEAF_MM, drop_invalid=True, z_max=Z_MAX_MM)
    NSLOT = args.slots

    def new_pipe(c):
        p = capi.Pipeline(c, calib, params, vox, handle_missing=True, batch=B)
        if NSLOT > 1:
            p.set_sgbm_option("shared_device", 1)     # several matchers share this GPU: keep the two horizontal passes separate
        for k, v in sgbm_opts:
            p.set_sgbm_option(k, int(v))
        return p
    pipe = new_pipe(ctx)

    # frame pool larger than L2 (126 MB): POOL x 5.5 MB resident in HBM, a different batch every step;
    # frame i of the stream goes to rank i mod N: each rank builds the frames it owns (weak scaling: B per rank per step)
    dev_pool = torch.from_numpy(np.stack(mine)).cuda()                       # [P][H][2W][3] u8
    host_pool = torch.from_numpy(np.stack(mine)).pin_memory()
    host_np = host_pool.numpy()

    def batch_dev(step):
        s = (step * B) % (n_pool - B + 1)
        return [dev_pool[s + i] for i in range(B)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: frames resident in HBM. Several pipeline handles on their own streams, one host thread each (a handle is
    # one device <-> one stream; different handles may run concurrently): kernels of one batch overlap those of another ----
    slots = []
    for k in range(NSLOT if args.value_runner == "python" else 1):
        if k == 0:
            c_k, st_k, p_k = ctx, stream, pipe
        else:
            c_k = capi.Context(local)
            st_k = torch.cuda.Stream()
            c_k.set_stream(st_k.cuda_stream)
            p_k = new_pipe(c_k)
        slots.append(dict(ctx=c_k, stream=st_k, pipe=p_k))

    def dev_step(slot, s):
        slots[slot]["pipe"].run(batch_dev(s))

    def run_slots(fn, steps, first):
        errs = []

        def worker(slot):
            try:
                for s in range(slot, steps, NSLOT):
                    fn(slot, first + s)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=worker, args=(k,)) for k in range(NSLOT)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]

    def timed_slots(fn, steps, warmup):
        """W warm-up steps, then exactly `steps` steps between CUDA events spanning all streams; max over ranks."""
        run_slots(fn, max(warmup, NSLOT), 0)
        barrier()
        l0 = sum(sl["ctx"].kernel_launches for sl in slots)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for sl in slots[1:]:
            sl["stream"].wait_event(e0)
        run_slots(fn, steps, warmup)
        for sl in slots[1:]:
            ev = torch.cuda.Event()
            ev.record(sl["stream"])
            stream.wait_event(ev)
        e1.record(stream)
        barrier()
        t1 = time.perf_counter()
        ms = sharding.max_over_ranks(e0.elapsed_time(e1), dist if world > 1 else None)
        return ms, sum(sl["ctx"].kernel_launches for sl in slots) - l0, t0, t1

    steps = args.steps if not c4 else max(1, -(-args.frames // (world * FPS)))

    def timed_stream(runner, pool, n_frames, warmup_frames):
        """Native runner: warm-up frames, then exactly n_frames between CUDA events; max over ranks."""
        runner.process_pool(warmup_frames, pool)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        st = runner.process_pool(n_frames, pool)             # returns when every result of the run is in host memory
        e1.record(stream)
        barrier()
        t1 = time.perf_counter()
        ms = sharding.max_over_ranks(e0.elapsed_time(e1), dist if world > 1 else None)
        assert st.frames == n_frames, (st.frames, n_frames)
        return ms, st, t0, t1

    sampler = ClockSampler(local) if rank == 0 else None
    if args.value_runner == "native":
        # the library's own worker threads (one per handle) over the device-resident pool: no host language on the path
        rv = capi.Stream([local], NSLOT, calib, params, vox, handle_missing=True, batch=B, download_clouds=False)
        for k, v in sgbm_opts:
            rv.set_sgbm_option(k, int(v))
        ms, st_v, t0, t1 = timed_stream(rv, dev_pool, steps * FPS, max(args.warmup * FPS, NSLOT * B))
        launches = int(st_v.kernel_launches)
        rv.close()
    else:
        ms, launches, t0, t1 = timed_slots(dev_step, steps * (FPS // B), args.warmup * (FPS // B))
    clocks = sampler.stop(t0, t1) if sampler else None
    fps = sharding.aggregate_throughput(FPS * steps, world, ms)

    # ---- per-kernel times of one step (CUDA events on the launching stream, inside the library) ----
    sg = capi.Sgbm(ctx, params)
    sg.set_profiling(True)
    if NSLOT > 1:
        sg.set_option("shared_device", 1)
    for k, v in sgbm_opts:
        sg.set_option(k, int(v))
    gl = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
    gr = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
    do = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
    rect = capi.Rectifier(ctx, calib)
    with torch.cuda.stream(stream):
        for i in range(B):
            f = batch_dev(0)[i]
            _, _, a, b = rect.rectify(f[:, :W], f[:, W:], want_gray=True)
            gl[i].copy_(a); gr[i].copy_(b)
        acc = {}
        reps = 3
        for rep in range(reps + 1):
            sg.compute_batch(gl, gr, do)
            ctx.synchronize()
            if rep == 0:
                continue                                  # warm-up (allocations)
            for name, t in sg.stage_times():
                acc[name] = acc.get(name, 0.0) + t / reps
    sg.close(); rect.close()
    W1 = W - D
    vol = W1 * H * D * 2                                   # bytes of one int16 volume per frame
    img = W * H
    pixdisp = W1 * H * D                                   # ops are counted over W1*H*D
    # algorithmic bytes per frame of each SGBM kernel (DESIGN.md "Kernels"): distinct bytes it must read + write
    algo = {"prefilter": 2 * img + (16 + 25) * img, "cost": (16 + 25) * img + vol, "path_v_fused_down": 2 * vol,
            "cost_path_v_fused": (16 + 25) * img + 2 * vol, "path_v_fused_up": 3 * vol, "path_v0": 2 * vol, "path_v1": 3 * vol, "path_v2": 3 * vol, "path_h_fwd": 3 * vol,
            "path_h_bwd_wta": 2 * vol + 6 * img, "path_h_both_wta": 5 * vol + 6 * img, "lrcheck": 8 * img, "median3": 4 * img, "speckle": 24 * img}
    # algorithmic int16 ops per pix*disp of each SGBM kernel (BT cost 20, box 4, 9 per path, WTA etc. 8)
    ops = {"cost": 24, "cost_path_v_fused": 24 + 3 * 9, "path_v_fused_down": 3 * 9, "path_v_fused_up": 3 * 9, "path_h_fwd": 9,
           "path_h_bwd_wta": 9 + 8, "path_h_both_wta": 2 * 9 + 8}
    ncu = ncu_numbers()
    dom = max(acc, key=acc.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    # integer-SIMD ceiling (tools/microbench/alu_rates.cu): two pipes of half a warp-instruction per clock and
    # scheduler = 128 lanes per clock per SM, two int16 lanes per 32-bit lane
    alu_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12           # T int16-op/s
    hbm_ach = algo.get(dom, 2 * vol) * B / (acc[dom] * 1e-3) / 1e9
    sgbm_ms = sum(acc.values())
    core_bytes = (2 * vol + 4 * img) * B                   # compulsory 3.69 MB + one int16 volume round trip per frame
    nst = (ncu or {}).get("stages", {}).get(dom)
    nscale = B / float((ncu or {}).get("frames_per_launch", 8))
    hbm_view = {"bound": "hbm", "achieved": hbm_ach, "peak": peak, "unit": "GB/s", "frac": hbm_ach / peak,
                "traffic": nst["traffic"] * nscale if nst else None,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650",
                "algorithmic_bytes_per_launch": algo.get(dom, 2 * vol) * B}
    if dom in ops and nst and (nst.get("alu_pipe_pct") or 0) > 1.5 * 100 * hbm_view["frac"]:
        # the dominant kernel keeps the integer pipes busier than the memory system: report it against the ALU ceiling
        alu_ach = ops[dom] * pixdisp * B / (acc[dom] * 1e-3) / 1e12
        roofline = {"bound": "alu", "kernel": dom, "achieved": alu_ach, "peak": alu_peak, "unit": "Tint16op/s", "frac": alu_ach / alu_peak,
                    "traffic": hbm_view["traffic"], "algorithmic_ops_per_pixdisp": ops[dom],
                    "peak_source": "148 SMs x 128 lanes/clk x 2 int16 per lane x %.0f MHz (two integer pipes of 0.5 warp-instr/clk/scheduler, "
                                   "measured: tools/microbench/alu_rates.cu)" % sm_mhz,
                    "hbm": hbm_view}
    else:
        roofline = dict(hbm_view, kernel=dom)
    roofline.update({
        "traffic_source": ("%s sha256 %s (ncu --set full, per launch at %d frames, scaled to this batch)" % (
            ncu["file"], ncu["sha256"], ncu["frames_per_launch"])) if ncu else None,
        "sm_issue_active_pct": nst["issue_active_pct"] if nst else None, "sm_alu_pipe_pct": nst["alu_pipe_pct"] if nst else None,
        "kernel_ms": acc[dom], "kernel_share_of_sgbm": acc[dom] / sgbm_ms,
        "sgbm_core": {"achieved": core_bytes / (sgbm_ms * 1e-3) / 1e9, "frac": core_bytes / (sgbm_ms * 1e-3) / 1e9 / peak,
                      "ms_per_batch": sgbm_ms, "algorithmic_bytes": core_bytes, "unit": "GB/s"},
        "stage_ms_per_batch": {k: round(v, 4) for k, v in acc.items()},
        "stage_hbm_frac": {k: round(algo[k] * B / (v * 1e-3) / 1e9 / peak, 3) for k, v in acc.items() if k in algo and v > 0},
        "stage_alu_frac": {k: round(ops[k] * pixdisp * B / (v * 1e-3) / 1e12 / alu_peak, 3) for k, v in acc.items() if k in ops and v > 0},
    })

    # ---- per-stage times of the whole pipeline (one extra profiled step) ----
    torch.cuda.synchronize()                               # nothing of the timed legs may still be in flight
    pipe.set_profiling(True)
    with torch.cuda.stream(stream):
        for _ in range(3):                                 # the first profiled step creates the events; report the last one
            pipe.run(batch_dev(1))
            ctx.synchronize()
            pipe_stages = {k: round(v, 4) for k, v in pipe.stage_times()}
    pipe.set_profiling(False)
    roofline["pipeline_stage_ms_per_batch"] = pipe_stages
    for sl in slots[1:]:
        sl["pipe"].close()

    # ---- e2e: pinned host frames in, disparity maps + clouds out, through the native frame-parallel runner of the C ABI
    # (worker threads in the library, one per handle; Python only starts the run) ----
    ESLOT = max(1, args.e2e_slots)      # more handles than the resident leg: the copies of one batch hide behind the kernels of the others
    runner = capi.Stream([local], ESLOT, calib, params, vox, handle_missing=True, batch=B, download_clouds=True)
    for k, v in sgbm_opts:
        runner.set_sgbm_option(k, int(v))
    ms_e, st_e, _, _ = timed_stream(runner, host_np, steps * FPS, max(args.warmup * FPS, ESLOT * B))
    runner.close()
    fps_e = sharding.aggregate_throughput(FPS * steps, world, ms_e)

    # ---- the same end-to-end leg fed with NV12 frames, the format a hardware decoder delivers (sdr_stream_set_input_format):
    # half the host->device bytes, the colour conversion runs on the device in front of the rectification ----
    e2e_nv12 = None
    if not args.no_ingest_leg:
        from stereo_depth_ruler_b200 import synth
        nv_pool = torch.from_numpy(np.stack([synth.bgr_to_nv12(f) for f in mine])).pin_memory()
        runner = capi.Stream([local], ESLOT, calib, params, vox, handle_missing=True, batch=B, download_clouds=True)
        runner.set_input_format(capi.PIX_NV12)
        ms_n, st_n, _, _ = timed_stream(runner, nv_pool.numpy(), steps * FPS, max(args.warmup * FPS, ESLOT * B))
        runner.close()
        e2e_nv12 = {"value": sharding.aggregate_throughput(FPS * steps, world, ms_n), "unit": "frames/s", "ms_per_step": ms_n / steps,
                    "h2d_bytes_per_step": int(st_n.h2d_bytes // steps), "d2h_bytes_per_step": int(st_n.d2h_bytes // steps),
                    "input": "NV12 SBS frames in pinned host memory (sdr_ingest_to_bgr on the device); not the headline e2e: the "
                             "reference arm starts from BGR frames"}
        del nv_pool

    # ---- parity gate: the frames the timed legs ran on, through the same C-ABI call, against cv2 ----
    gate = None
    if rank == 0 and gate_ref is not None:
        bad = []
        dh = np.empty((B, H, W), np.int16)
        for s0 in range(0, n_gate, B):
            nb = min(B, n_gate - s0)
            pipe.run([host_np[s0 + i] for i in range(nb)], disp=[dh[i] for i in range(nb)])
            for i in range(nb):
                ref_map, ref_nvox = gate_ref[s0 + i]
                if not np.array_equal(dh[i], ref_map):
                    bad.append("frame %d: %d CV_16S values differ from cv2" % (s0 + i, int((dh[i] != ref_map).sum())))
                elif int(pipe.views[i].n) != int(ref_nvox):
                    bad.append("frame %d: %d voxels, CPU path %d" % (s0 + i, int(pipe.views[i].n), int(ref_nvox)))
        gate = {"status": "pass" if not bad else "FAIL", "frames": n_gate,
                "what": "CV_16S maps array_equal to cv2 StereoSGBM on the timed frames (after the GPU's own rectification == cv2.remap), equal voxel counts",
                "failures": bad[:5]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s16",
            "data": "synthetic", "mpix_disp_per_s": fps * MPIX_DISP_PER_FRAME,
            "config": base_config(args.workload),
            "run": {"frames_per_step_per_gpu": FPS, "frames_per_submission": B, "frame_pool": n_pool, "frames_timed_per_gpu": FPS * steps,
                    "l2": "inputs larger than L2 (frame pool %d MB, a different batch each step; volumes %d MB per step)" % (
                        n_pool * H * 2 * W * 3 >> 20, 2 * vol * FPS >> 20),
                    "sharding": "frame i -> rank i mod N, no collective", "numa_node": numa_node, "sgbm_options": dict((k, int(v)) for k, v in sgbm_opts),
                    "concurrency": "%d pipeline handles (own stream + host thread each) per GPU alternate batches, so the copies and the "
                                   "ALU-bound kernels of one batch overlap the HBM-bound kernels of another; value: %s; e2e: %d native worker "
                                   "threads of sdr_stream_*" % (NSLOT, "native runner (sdr_stream_process_pool) on a device-resident pool" if args.value_runner == "native" else "Python threads over sdr_pipeline_run", ESLOT)},
            "e2e": {"value": fps_e, "unit": "frames/s", "h2d_bytes_per_step": int(st_e.h2d_bytes // steps), "d2h_bytes_per_step": int(st_e.d2h_bytes // steps),
                    "ms_per_step": ms_e / steps, "api": "sdr_stream_create / sdr_stream_process_pool (include/sdr_c.h)"},
            "e2e_ingest_nv12": e2e_nv12,
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity_gate": gate["status"] if gate else "skipped", "parity": gate,
        }
        _emit(line)
    pipe.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if gate and gate["status"] != "pass":
        raise SystemExit(3)


def main():
    # Libraries (NCCL, torchrun) may print to stdout; the contract is ONE JSON line there. Route fd 1 to stderr for the
    # duration of the run and write the JSON line to the real stdout at the end.
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit

    def _emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="frames per submission (one launch of every kernel)")
    ap.add_argument("--frames-per-step", type=int, default=16, help="frames per step per GPU (a multiple of --batch): 20 steps time 320 frames")
    ap.add_argument("--pool", type=int, default=32, help="distinct frames resident per GPU (pool > L2)")
    ap.add_argument("--slots", type=int, default=4, help="pipeline handles (streams + host threads) per GPU")
    ap.add_argument("--e2e-slots", type=int, default=6, help="pipeline handles per GPU of the end-to-end legs (measured: 4 -> 6 handles +4 %% e2e)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=4096, help="c4: frames of the whole stream")
    ap.add_argument("--gate-frames", type=int, default=16, help="frames of rank 0's timed pool the parity gate checks against cv2")
    ap.add_argument("--sgbm-opt", action="append", default=[], help="name=value kernel knob (sdr_sgbm_set_option); results never change")
    ap.add_argument("--value-runner", default="native", choices=["native", "python"],
                    help="who drives the HBM-resident leg: the library's worker threads (default) or Python threads over sdr_pipeline_run")
    ap.add_argument("--no-ingest-leg", action="store_true", help="skip the extra e2e leg fed with NV12 frames")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-gate", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
