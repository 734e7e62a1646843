"""Multi-GPU plumbing: one process per GPU (torch.distributed), chromosomes sharded over ranks.

gextract / gscreen / per-row quantiles need no collective: every rank owns a disjoint, order-preserving row range.
gsummary / gdist combine their small partial results -- the six IntervalSummary doubles and the uint64 histogram --
exactly where the reference's parent process merges its children (src/GenomeTrackSummary.cpp:210-216,
src/GenomeTrackDistribution.cpp:131-141): here one all-reduce over NCCL (NVLink) instead of a shared-memory merge.
The same code runs on the gloo backend with CPU tensors (the world_size-2 CPU tests).
"""
import numpy as np

from .synth import lpt_shards


def shard_chromosomes(chroms, world):
    """owner[chromid] -> rank; whole chromosomes, longest-processing-time greedy."""
    return lpt_shards(chroms, world)


def _dev():
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def reduce_summary(partial6):
    """All-reduce of IntervalSummary partials {n, nn, sum, min, max, ssq}: sum / min / max as IntervalSummary::merge. One record
    (gsummary) or an array of records of shape (..., 6) -- the per-bin partials of gbins.summary, which every rank holds for the
    same bins -- reduced in the same three collectives."""
    import torch
    import torch.distributed as dist
    p = np.asarray(partial6, dtype=np.float64)
    if p.shape[-1] != 6:
        raise ValueError("summary partials have 6 values per record")
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return p.copy()
    t = torch.from_numpy(p.reshape(-1, 6).copy()).to(_dev())
    sums = t[:, [0, 1, 2, 5]].contiguous()
    mn, mx = t[:, 3].contiguous(), t[:, 4].contiguous()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    s = sums.cpu().numpy()
    out = np.stack([s[:, 0], s[:, 1], s[:, 2], mn.cpu().numpy(), mx.cpu().numpy(), s[:, 3]], axis=1)
    return out.reshape(p.shape)


def reduce_hist(counts):
    """All-reduce (sum) of a uint64 histogram; counts stay exact (int64 on the wire)."""
    import torch
    import torch.distributed as dist
    c = np.asarray(counts, dtype=np.uint64)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return c.copy()
    t = torch.from_numpy(c.astype(np.int64)).to(_dev())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.uint64).reshape(c.shape)
