"""N > 1 host logic on CPU: two gloo ranks shard independent ram_permutation instances round-robin, each rank produces
its instances' commitments (with the CPU oracle standing in for the GPU engine) and the all-gather reassembles them in
instance order -- identical to the single-process result."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
N_TOTAL = 5  # deliberately not divisible by the world size


def _commitment_of_instance(i):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    import helpers as H
    import orc as O
    from era_zkevm_circuits_b200 import synthetic
    lib = O.load()
    u, s = synthetic.ram_trace(200, seed=100 + i, n_cells=10, n_nondet=1)
    io, _, _ = H.ram_instance(lib, u, s, 1)
    rc, _, _, com, _ = O.ram_entry_point(lib, io, u, s, 256, want_trace=False)
    assert rc == 0
    return com


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(HERE))
    from era_zkevm_circuits_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_instances(N_TOTAL, rank, world)
    local = np.stack([_commitment_of_instance(i) for i in mine]) if mine else np.zeros((0, 4), dtype=np.uint64)
    allc = sharding.gather_commitments(local, N_TOTAL, rank, world)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), allc)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gather_commitments(tmp_path):
    world = 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    want = np.stack([_commitment_of_instance(i) for i in range(N_TOTAL)])
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npy"))
        assert np.array_equal(got, want)
    assert len({tuple(c) for c in want.tolist()}) == N_TOTAL


def test_shard_assignment():
    sys.path.insert(0, os.path.dirname(HERE))
    from era_zkevm_circuits_b200 import sharding
    for world in (1, 2, 4, 8):
        owned = sorted(i for r in range(world) for i in sharding.shard_instances(64, r, world))
        assert owned == list(range(64))
        assert all(len(sharding.shard_instances(64, r, world)) == 64 // world for r in range(world))


# ---- one long grand product cut over the ranks by row ranges ----------------------------------
GP_ROWS, GP_ENC = 1000, 20


def _gp_inputs():
    rng = np.random.default_rng(44)
    P = 0xFFFFFFFF00000001
    lhs = rng.integers(0, P, size=(GP_ENC, GP_ROWS), dtype=np.uint64)
    rhs = rng.integers(0, P, size=(GP_ENC, GP_ROWS), dtype=np.uint64)
    ch = rng.integers(0, P, size=(2, GP_ENC + 1), dtype=np.uint64)
    flags = (rng.integers(0, 10, size=GP_ROWS) != 0).astype(np.uint8)
    return lhs, rhs, ch, flags


def _gp_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    import orc as O
    from era_zkevm_circuits_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = O.load()
    lhs, rhs, ch, flags = _gp_inputs()
    lo, hi = sharding.row_range(GP_ROWS, rank, world)

    def local(acc_in):  # the CPU oracle stands in for zkc_accumulate_grand_products on this rank's rows
        acc, _, fin = O.accumulate_grand_products(lib, lhs[:, lo:hi], rhs[:, lo:hi], ch, acc_in, flags[lo:hi])
        return acc, fin

    acc, fin, grand = sharding.distributed_grand_products(local, rank, world, acc_in=(3, 5, 7, 11))
    np.save(os.path.join(out_dir, f"gp{rank}.npy"), acc)
    np.save(os.path.join(out_dir, f"gpt{rank}.npy"), grand)
    dist.barrier()
    dist.destroy_process_group()


def test_grand_product_over_row_ranges(tmp_path):
    sys.path.insert(0, HERE)
    import orc as O
    world = 3
    port = 31500 + os.getpid() % 2000
    mp.spawn(_gp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    lhs, rhs, ch, flags = _gp_inputs()
    want, _, fin = O.accumulate_grand_products(O.load(), lhs, rhs, ch, np.array([3, 5, 7, 11], dtype=np.uint64), flags)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), f"gp{r}.npy")) for r in range(world)], axis=1)
    assert np.array_equal(got, want)
    for r in range(world):
        assert np.load(os.path.join(str(tmp_path), f"gpt{r}.npy")).tolist() == fin.tolist()
