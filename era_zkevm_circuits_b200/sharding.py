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
    mine = torch.from_numpy(np.ascontiguousarray(local_commitments, dtype=np.uint64).reshape(-1, 4).view(np.int64))  # (0, 4) for a rank without instances
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


GL_P = 0xFFFFFFFF00000001


def row_range(n_rows: int, rank: int, world: int):
    """contiguous row range [lo, hi) of rank `rank` when one long loop is cut over `world` ranks"""
    per = (n_rows + world - 1) // world
    return min(rank * per, n_rows), min((rank + 1) * per, n_rows)


def distributed_grand_products(local_fn, rank: int, world: int, acc_in=(1, 1, 1, 1), device=None, scale_fn=None):
    """ONE long grand product (utils.rs:81-137: the running lhs / rhs accumulators of a sorter circuit) over `world`
    ranks, each owning a contiguous row range.  The accumulators are running PRODUCTS, so a
    rank's rows only need the product of everything before them: every rank accumulates its own rows from the
    neutral element ONCE, the 4 local totals are exchanged in the path's ONE collective (an all-gather of 4 x u64 per
    rank), and each rank multiplies its 4 accumulator columns by acc_in * (the exclusive prefix of the lower ranks'
    totals) -- exactly the reference's own instance chaining (hidden_fsm_output of part k = hidden_fsm_input of part
    k + 1, ram_permutation/input.rs:52-62) with the chain resolved in one exchange instead of sequentially.  Products in
    the field are exact, so the scaled columns are bit-identical to a seeded second pass.

    local_fn(acc_in [4] uint64) -> (acc_out, acc_final [4] uint64) runs the rank's rows (zkc_accumulate_grand_products
    on its shard); scale_fn(acc_out, seed [4]) multiplies the columns in place (zkc_scale_accumulators; default: numpy
    big-int arithmetic on host arrays).  Returns (acc_out, acc_final of this rank, grand totals [4] over all ranks)."""
    import torch
    import torch.distributed as dist
    ones = np.ones(4, dtype=np.uint64)
    acc_out, local_total = local_fn(ones)
    mine = torch.from_numpy(np.ascontiguousarray(local_total, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        mine = mine.to(device)
    allt = torch.zeros((world, 4), dtype=torch.int64, device=mine.device)
    if world > 1:
        dist.all_gather_into_tensor(allt, mine.reshape(1, 4))
    else:
        allt.copy_(mine.reshape(1, 4))
    totals = allt.cpu().numpy().view(np.uint64)
    seed = [int(a) % GL_P for a in acc_in]
    for r in range(rank):
        seed = [s * int(t) % GL_P for s, t in zip(seed, totals[r])]
    if any(s != 1 for s in seed):
        if scale_fn is not None:
            scale_fn(acc_out, np.array(seed, dtype=np.uint64))
        elif acc_out is not None:
            a = np.asarray(acc_out)
            for c in range(4):
                a[c] = np.array([int(v) * seed[c] % GL_P for v in a[c]], dtype=np.uint64)
    acc_final = np.array([int(t) * s % GL_P for t, s in zip(local_total, seed)], dtype=np.uint64)
    grand = [int(a) % GL_P for a in acc_in]
    for r in range(world):
        grand = [g * int(t) % GL_P for g, t in zip(grand, totals[r])]
    return acc_out, acc_final, np.array(grand, dtype=np.uint64)
