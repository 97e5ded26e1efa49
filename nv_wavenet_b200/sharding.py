"""Batch sharding across GPUs: utterances never interact, so rank g simply owns a contiguous
slice of the batch -- its own conditioning slice, selector slice and yOut rows.  The only communication is one
broadcast of the packed weight blob at start-up (NCCL over NVLink on GPUs, gloo in the CPU tests) and an optional
gather of the int32 results.  No per-step collective exists on this path.
"""
import numpy as np


def shard_range(batch, rank, world):
    """[lo, hi) utterances of `rank` when `batch` is split as evenly as possible over `world` ranks."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_inputs(Lh, selectors, rank, world):
    """Lh [N][L][B][2R] and selectors [N][B] (reference layouts: batch is an INNER dimension, nv_wavenet.cuh:144,
    singleblock.cuh:232) -> this rank's contiguous copies."""
    lo, hi = shard_range(Lh.shape[2], rank, world)
    return np.ascontiguousarray(Lh[:, :, lo:hi]), np.ascontiguousarray(selectors[:, lo:hi])


def broadcast_weights(w, src=0):
    """Broadcast a dict of fp32 weight arrays from rank `src` with torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist
    out = {}
    for k in sorted(w):
        t = torch.from_numpy(np.ascontiguousarray(w[k], dtype=np.float32))
        dist.broadcast(t, src)
        out[k] = t.numpy()
    return out


def gather_outputs(y_local, batch, world):
    """All ranks' yOut rows [b_local][N] -> the full [batch][N] on every rank (yOut is row-contiguous per utterance)."""
    import torch
    import torch.distributed as dist
    n = y_local.shape[1]
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    pad = max(sizes)
    mine = torch.zeros((pad, n), dtype=torch.int32)
    mine[: y_local.shape[0]] = torch.from_numpy(np.ascontiguousarray(y_local, dtype=np.int32))
    bufs = [torch.zeros((pad, n), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(bufs, mine)
    return np.concatenate([bufs[r][: sizes[r]].numpy() for r in range(world)], axis=0)
