"""Multi-GPU side of the rollout: trajectories are independent, so the batch is sharded across
ranks with no data-path collective; the only exchange is this one — 8 doubles per rank of rnd
statistics — after which every rank holds the global loss / log-Z numbers.
"""
from __future__ import annotations

import torch


def combine_stats(stats: torch.Tensor, process_group=None) -> torch.Tensor:
    """Merge per-rank `sdes_rnd_stats` vectors (layout: include/sdes_b200.h).

    [0] n_kept, [1] sum, [2] sum of squares, [5] n_total add; [3] max(-rnd) is a max; [4]
    sum exp(-rnd - max_rank) is rescaled to the global max before adding."""
    if process_group is None:
        return stats
    import torch.distributed as dist

    world = dist.get_world_size(process_group)
    if world == 1:
        return stats
    if not stats.is_cuda:  # gloo (CPU tests of the host logic): list all-gather + the host restatement of the merge
        parts = [torch.empty_like(stats) for _ in range(world)]
        dist.all_gather(parts, stats.contiguous(), group=process_group)
        return merge_stats(torch.stack(parts))
    gathered = torch.empty((world, stats.numel()), dtype=stats.dtype, device=stats.device)
    dist.all_gather_into_tensor(gathered, stats.contiguous(), group=process_group)
    # one launch (csrc/sdes_api.cu merge_stats_kernel) instead of ~15 tiny torch kernels serialised on the stream
    import ctypes as C

    from . import _cabi

    out = torch.empty_like(stats)
    with torch.cuda.device(stats.device):
        _cabi.check(_cabi.lib().sdes_merge_stats(gathered.data_ptr(), world, out.data_ptr(),
                                                 C.c_void_p(torch.cuda.current_stream(stats.device).cuda_stream)), "sdes_merge_stats")
    return out


def merge_stats(gathered: torch.Tensor) -> torch.Tensor:
    """(W, 8) -> (8,) ; pure bookkeeping on W*8 doubles."""
    out = gathered.sum(dim=0)
    mx = gathered[:, 3]
    has = gathered[:, 0] > 0
    # ranks that kept nothing report max = -inf and exp-sum 0; NaN propagates like the reference's .max()
    gmx = torch.where(has, mx, torch.full_like(mx, -float("inf"))).max()
    gmx = torch.where(torch.isnan(mx).any(), torch.full_like(gmx, float("nan")), gmx)
    scale = torch.where(has, torch.exp(mx - gmx), torch.zeros_like(mx))
    out[3] = gmx
    out[4] = (gathered[:, 4] * scale).sum()
    n, s1, s2 = out[0], out[1], out[2]
    out[6] = (s2 - s1 * s1 / n) / (n - 1.0)  # the derived slots of the combined statistics
    out[7] = s1 / n
    return out


def shard_range(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of the global trajectory batch owned by `rank`."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per
