"""Multi-GPU sharding of independent instances (BASELINE config C5 / north_star: "batches of
independent instances are sharded across GPUs with no communication").

One process per GPU (torch.distributed); the only collectives are control-plane: a barrier
around the timed region, a MAX over the per-rank device times, and an optional gather of the
per-instance results on rank 0.  The data path (the solves) never communicates.
"""
import numpy as np


def shard_instances(n_instances, rank, world):
    """Static sharding `instance i -> rank i mod world`."""
    return list(range(rank, n_instances, world))


def max_over_ranks(value, dist=None, device=None):
    """Whole-job time = the slowest rank's device time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(local, n_instances, dist=None):
    """local: {instance_id: 1-D float array}; returns the full list on every rank (gloo/NCCL object
    gather — control plane only, results are tiny compared with the solves)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [local[i] for i in range(n_instances)]
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, {int(k): np.asarray(v) for k, v in local.items()})
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged[i] for i in range(n_instances)]
