"""Frame-parallel sharding across the GPUs of one box: every frame pair is independent (the
reference's loop body at stereo_displayer.cpp:145-198 carries no state from frame to frame), so frame i goes to rank
i mod N and no data-path collective exists. The only cross-rank operation is the MAX reduction of the device time."""


def frames_for_rank(n_frames, rank, world):
    """Indices of the frames rank `rank` of `world` processes: i with i mod world == rank."""
    return list(range(rank, n_frames, world))


def max_over_ranks(ms, dist, device="cuda"):
    """Device-measured milliseconds -> max over all ranks (every multi-GPU number is timed as the slowest rank)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(ms)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(frames_per_rank, world, ms):
    """Whole-job frames/s: every rank processed frames_per_rank frames in (at most) ms milliseconds."""
    return world * frames_per_rank / (ms / 1e3)
