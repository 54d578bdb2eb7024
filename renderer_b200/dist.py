"""Multi-GPU frame assembly, host-side mirror: row-cyclic sharding + ONE all-gather + de-interleave.

The device path is renderer_b200.Pipeline (b200r_pipeline_*, csrc/cuda/dist.cu); the functions here restate its sharding
arithmetic on numpy arrays so the host logic can be tested on CPU with gloo (tests/test_cpu_dist.py).

Rank r of P renders rows r, r+P, r+2P, ... (the orbit view is ~90 % background, so contiguous bands would
be badly unbalanced) into a packed buffer of ceil(H/P) rows; one all-gather of the packed rows (NCCL over
NVLink on GPUs; gloo on CPU for the host-logic tests) gives every rank all shards; a de-interleave pass
restores scan order. Nothing else is exchanged (the scene is replicated, the Z-buffer never leaves a GPU).
"""
import numpy as np


def rows_per_shard(height, n_shards):
    return (height + n_shards - 1) // n_shards


def shard_rows(height, n_shards, rank):
    """The screen rows rank `rank` owns."""
    return list(range(rank, height, n_shards))


def pack_shard(rows_img, height, n_shards):
    """Pad a rank's packed rows to rows_per_shard (all-gather needs equal counts)."""
    rps = rows_per_shard(height, n_shards)
    if rows_img.shape[0] == rps:
        return rows_img
    out = np.zeros((rps, rows_img.shape[1]), dtype=rows_img.dtype)
    out[: rows_img.shape[0]] = rows_img
    return out


def deinterleave(gathered, height, n_shards):
    """gathered: (n_shards*rows_per_shard, W) as produced by the all-gather -> (height, W) in scan order.
    Host-side mirror of b200r_deinterleave_device."""
    rps = rows_per_shard(height, n_shards)
    g = np.asarray(gathered).reshape(n_shards, rps, -1)
    out = np.empty((height, g.shape[2]), dtype=g.dtype)
    for s in range(n_shards):
        rows = shard_rows(height, n_shards, s)
        out[rows] = g[s, : len(rows)]
    return out


def all_gather_rows(shard, group=None):
    """torch.distributed all-gather of equally sized packed shards -> stacked tensor (P*rps, W)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * shard.shape[0], shard.shape[1]), dtype=shard.dtype, device=shard.device)
    try:
        dist.all_gather_into_tensor(out, shard.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):
        parts = [torch.empty_like(shard) for _ in range(world)]
        dist.all_gather(parts, shard.contiguous(), group=group)
        out = torch.cat(parts, 0)
    return out


def init_nccl(local_rank):
    """torch.distributed over NCCL (bench.py / tools/dist_check.py plumbing: barriers, timing reductions, handing out the job id).
    The frame pipeline itself is renderer_b200.Pipeline (b200r_pipeline_*, csrc/cuda/dist.cu)."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    except (AttributeError, TypeError):
        dist.init_process_group("nccl", device_id=dev)
    return dist
