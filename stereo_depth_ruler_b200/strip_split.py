"""Row-strip split of ONE oversize frame across the GPUs of a box (BASELINE config 5) -- the multi-PROCESS
form (one rank per GPU under torchrun). The native single-process form, with peer copies instead of NCCL, is
sdr_sgbm_split_* / capi.SgbmSplit (csrc/split.cu); this module is the same plan for callers that already run one
process per GPU.

MODE_SGBM_3WAY is four independent stripes by construction (OpenCV recomputes the warm-up rows of each stripe), so each
rank matches its stripes from the input rows it needs, the raw strips travel to rank 0 (NCCL send/recv over NVLink, a
few MB) and the 3x3 median + speckle filter -- the only stages that look across stripe borders -- run once there.
MODE_SGBM / MODE_HH have a single stripe (rows depend on the whole image): rank 0 does everything.
"""
import numpy as np

from . import capi


def assign_stripes(n_stripes, world):
    """Contiguous blocks of stripes per rank: [(first, count)] * world (count may be 0)."""
    base, extra = divmod(n_stripes, world)
    out, first = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


def compute(sgbm, left, right, rank=0, world=1, dist=None):
    """Every rank calls this with the same (device-resident torch) gray images; rank 0 returns the CV_16S map, others None."""
    import torch
    H, W = left.shape
    stripes = capi.stripe_rows(sgbm.params, W, H)
    plan = assign_stripes(len(stripes), world)
    raw = torch.empty((H, W), dtype=torch.int16, device=left.device)
    first, cnt = plan[rank]
    if cnt:
        sgbm.match_stripes(left, right, first, cnt, raw)
        sgbm.ctx.synchronize()
    if world > 1:
        for r in range(1, world):
            f, c = plan[r]
            if not c:
                continue
            y0, y1 = stripes[f][1], stripes[f + c - 1][2]
            # rows y0..y1 are contiguous in `raw`; NCCL has no int16, so the bytes travel as uint8
            if rank == r:
                dist.send(raw[y0:y1].view(torch.uint8), dst=0)
            elif rank == 0:
                dist.recv(raw[y0:y1].view(torch.uint8), src=r)
        # dist.recv only orders torch's current stream; the handle's own (non-blocking) stream knows nothing of it, so
        # the host waits for the received strips before the post-filter is queued there
        torch.cuda.current_stream().synchronize()
    if rank != 0:
        return None
    out = sgbm.postfilter(raw)
    sgbm.ctx.synchronize()            # the handle's stream is asynchronous to torch's streams
    return out
