"""Multi-GPU plumbing: circuit instances are independent given their closed-form inputs
(/root/reference/src/fsm_input_output/mod.rs:32-48), so a job of `n_total` instances is sharded round-robin over the
ranks (one process per GPU) with no data-path collective; the only exchange is the gather of the 4-element
commitments.  torch.distributed is plumbing: NCCL on the GPUs, gloo in the CPU tests."""
import numpy as np


def shard_instances(n_total: int, rank: int, world: int):
    """indices of the instances rank `rank` owns (instance i -> rank i mod world)"""
    return list(range(rank, n_total, world))


def gather_commitments(local_commitments: np.ndarray, n_total: int, rank: int, world: int, device=None):
    """all-gather of the per-instance commitments; returns [n_total, 4] uint64 in instance order on every rank.
    Ranks may own different numbers of instances (n_total need not divide by world): shards are padded to the
    largest one for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    per = (n_total + world - 1) // world
    buf = torch.zeros((per, 4), dtype=torch.int64, device=device)
    mine = torch.from_numpy(np.ascontiguousarray(local_commitments, dtype=np.uint64).view(np.int64))
    buf[: mine.shape[0]] = mine.to(buf.device)
    out = torch.zeros((world * per, 4), dtype=torch.int64, device=device)
    if world > 1:
        dist.all_gather_into_tensor(out, buf)
    else:
        out.copy_(buf)
    out = out.cpu().numpy().view(np.uint64).reshape(world, per, 4)
    result = np.zeros((n_total, 4), dtype=np.uint64)
    for r in range(world):
        idx = shard_instances(n_total, r, world)
        result[idx] = out[r, : len(idx)]
    return result
