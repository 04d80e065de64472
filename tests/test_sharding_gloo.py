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


# ---- ONE storage_validity instance cut by row range over two gloo ranks (the CPU oracle stands in for the engine) -----------------
def _storage_instance():
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    import orc as O
    from era_zkevm_circuits_b200 import StorageDeduplicatorInstanceWitness, abi, synthetic
    lib = O.load()
    n, limit = 1200, 1250
    u, s, ts = synthetic.storage_trace(n, seed=5, n_cells=12)  # ~100 rows per cell: every cut lands inside a cell
    up, ufin = O.log_queue_simulate(lib, u)
    sp, sfin = O.log_queue_simulate(lib, s, ts)
    io = O.storage_closed_form(ufin, sfin, 0, True)
    whole = O.storage_validity_entry_point(lib, io, u, s, ts, limit)
    assert whole[0] == 0
    return lib, StorageDeduplicatorInstanceWitness(io, u, up, s, ts, sp, whole[5]), n, limit, whole


def _oracle_backend(lib):
    from types import SimpleNamespace
    import orc as O
    from era_zkevm_circuits_b200 import sharding

    def run(io, u, up, s, ts, sp, tails, limit, want_trace):
        rc, io2, trace, com, st, _ = O.storage_validity_entry_point(lib, io, u, s, ts, limit, want_trace=want_trace)
        return SimpleNamespace(closed_form_input=io2, trace=trace, status=st, commitment=com)

    def scale(cols, seed):
        for c in range(4):
            cols[c] = np.array([int(v) * int(seed[c]) % sharding.GL_P for v in cols[c]], dtype=np.uint64)

    return run, scale, (lambda e: O.commit_encoding(lib, e))


def _push_offsets(whole_trace, n, world):
    from era_zkevm_circuits_b200 import abi, sharding
    cum = np.concatenate([[0], np.cumsum(whole_trace[abi.ST_COLS["SHOULD_PUSH"]])]).astype(np.int64)
    return [int(cum[sharding.row_range(n, r, world)[0]]) for r in range(world)]


def _storage_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    lib, w, n, limit, whole = _storage_instance()
    from era_zkevm_circuits_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    run, scale, commit = _oracle_backend(lib)
    offs = _push_offsets(whole[2], n, world)
    res, lo, hi, rec = sharding.storage_rows_local(run, w, limit, rank, world, offs)
    allr = torch.zeros((world, len(rec)), dtype=torch.int64)
    dist.all_gather_into_tensor(allr, torch.from_numpy(rec.copy()).reshape(1, -1))  # the path's ONE collective
    com, io, trace, st = sharding.storage_rows_finish(res, rank, world, allr.numpy(), w.closed_form_input, offs, scale, commit)
    np.savez(os.path.join(out_dir, f"st{rank}.npz"), com=com, trace=trace, lo=lo, hi=hi, code=st.code, failed=st.failed_checks,
             fsm=np.frombuffer(bytes(io.hidden_fsm_output), dtype=np.uint8), done=io.completion_flag)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_one_storage_instance_by_rows(tmp_path):
    world = 2
    port = 31500 + os.getpid() % 2000
    mp.spawn(_storage_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    lib, w, n, limit, whole = _storage_instance()
    parts = [np.load(os.path.join(str(tmp_path), f"st{r}.npz")) for r in range(world)]
    for p in parts:
        assert int(p["code"]) == 0 and int(p["failed"]) == 0
        assert np.array_equal(p["com"], whole[3])  # every rank ends with the WHOLE instance's commitment
        assert p["fsm"].tobytes() == bytes(whole[1].hidden_fsm_output) and int(p["done"]) == whole[1].completion_flag
    assert [(int(p["lo"]), int(p["hi"])) for p in parts] == [(0, 600), (600, limit)]
    assert np.array_equal(np.concatenate([p["trace"] for p in parts], axis=1), whole[2])  # scaled accumulator columns included


def test_row_sharded_storage_chained_instance_and_bad_offsets():
    """single process, three virtual ranks: a chained (start_flag = 0) instance cut by rows, and wrong push offsets"""
    lib, w, n, limit, whole = _storage_instance()
    import orc as O
    from era_zkevm_circuits_b200 import StorageDeduplicatorInstanceWitness, abi, sharding
    run, scale, commit = _oracle_backend(lib)
    # second instance of a chain: rows [400, ...) with the FSM state the first instance left
    first = O.storage_validity_entry_point(lib, w.closed_form_input, w.unsorted_queue_witness, w.intermediate_sorted_queue_witness,
                                           w.intermediate_sorted_queue_timestamps, 400)
    nxt = abi.StorageClosedForm.from_buffer_copy(bytes(first[1])); nxt.start_flag = 0
    nxt.hidden_fsm_input = first[1].hidden_fsm_output
    cut = lambda a: a[400:]
    rest = O.storage_validity_entry_point(lib, nxt, cut(w.unsorted_queue_witness), cut(w.intermediate_sorted_queue_witness),
                                          cut(w.intermediate_sorted_queue_timestamps), limit - 400)
    assert rest[0] == 0
    w2 = StorageDeduplicatorInstanceWitness(nxt, cut(w.unsorted_queue_witness), cut(w.unsorted_queue_prev_tails), cut(w.intermediate_sorted_queue_witness),
                                            cut(w.intermediate_sorted_queue_timestamps), cut(w.intermediate_sorted_queue_prev_tails), rest[5])
    world = 3
    offs = _push_offsets(rest[2], n - 400, world)
    locs = [sharding.storage_rows_local(run, w2, limit - 400, r, world, offs) for r in range(world)]
    recs = np.stack([l[3] for l in locs])
    traces = []
    for r in range(world):
        com, io, trace, st = sharding.storage_rows_finish(locs[r][0], r, world, recs, nxt, offs, scale, commit)
        assert st.code == 0 and np.array_equal(com, rest[3]) and bytes(io.hidden_fsm_output) == bytes(rest[1].hidden_fsm_output)
        traces.append(trace)
    assert np.array_equal(np.concatenate(traces, axis=1), rest[2])
    bad = list(offs); bad[2] += 1  # the host's push count before rank 2 is off by one: found when the counts are exchanged
    locs = [sharding.storage_rows_local(run, w2, limit - 400, r, world, bad) for r in range(world)]
    com, io, trace, st = sharding.storage_rows_finish(locs[0][0], 0, world, np.stack([l[3] for l in locs]), nxt, bad, scale, commit)
    assert st.code != 0 and st.failed_checks & abi.ST_CHK["QUEUE_HINT"]


def test_row_sharded_log_sorter_virtual_ranks():
    """ONE log_sorter instance cut by rows (its FSM record is the previous key / item: the replay window is row lo - 1 alone);
    2 and 4 virtual ranks with the oracle as the backend"""
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    from types import SimpleNamespace
    import orc as O
    from era_zkevm_circuits_b200 import EventsDeduplicatorInstanceWitness, abi, sharding, synthetic
    lib = O.load()
    n, limit = 1500, 1530
    u, s = synthetic.events_trace(n, seed=8, rollback_pct=20)
    up, ufin = O.log_queue_simulate(lib, u)
    sp, sfin = O.log_queue_simulate(lib, s)
    io = O.events_closed_form(ufin, sfin, True)
    whole = O.log_sorter_entry_point(lib, io, u, s, limit)
    assert whole[0] == 0
    assert np.array_equal(sharding.events_closed_form_commitment(lambda e: O.commit_encoding(lib, e), whole[1]), whole[3])
    w = EventsDeduplicatorInstanceWitness(io, u, up, s, sp, whole[5])

    def run(io_, u_, up_, s_, sp_, tails, lim, want_trace):
        rc, io2, trace, com, st, _ = O.log_sorter_entry_point(lib, io_, u_, s_, lim, want_trace=want_trace)
        return SimpleNamespace(closed_form_input=io2, trace=trace, status=st, commitment=com)

    _, scale, commit = _oracle_backend(lib)
    cum = np.concatenate([[0], np.cumsum(whole[2][abi.EV_COLS["ADD_TO_QUEUE"]])]).astype(np.int64)
    for world in (2, 4):
        offs = [int(cum[sharding.row_range(n, r, world)[0]]) for r in range(world)]
        locs = [sharding.events_rows_local(run, w, limit, r, world, offs) for r in range(world)]
        recs = np.stack([l[3] for l in locs])
        traces = []
        for r in range(world):
            com, io_g, trace, st = sharding.events_rows_finish(locs[r][0], r, world, recs, io, offs, scale, commit)
            assert st.code == 0 and np.array_equal(com, whole[3])
            assert bytes(io_g.hidden_fsm_output) == bytes(whole[1].hidden_fsm_output) and bytes(io_g.final_queue_state) == bytes(whole[1].final_queue_state)
            traces.append(trace)
        assert np.array_equal(np.concatenate(traces, axis=1), whole[2])


def test_row_sharded_decommit_sorter_virtual_ranks():
    """ONE sort_decommittment_requests instance cut by rows (12-element full-state queues; the replay window is the run of equal code
    hashes that straddles the cut, which owns the carried first-encountered timestamp); 2 and 4 virtual ranks, oracle backend"""
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    from types import SimpleNamespace
    import orc as O
    from era_zkevm_circuits_b200 import CodeDecommittmentsDeduplicatorInstanceWitness, abi, sharding, synthetic
    lib = O.load()
    n, limit = 1500, 1530
    u, s = synthetic.decommit_requests_trace(n, seed=5, n_hashes=14)  # ~100 requests per hash: every cut lands inside a run
    up, ufin = O.decommit_queue_simulate(lib, u)
    sp, sfin = O.decommit_queue_simulate(lib, s)
    io = O.decommit_sorter_closed_form(ufin, sfin, True)
    whole = O.sort_decommittments_entry_point(lib, io, u, s, limit)
    assert whole[0] == 0
    assert np.array_equal(sharding.decommit_sorter_closed_form_commitment(lambda e: O.commit_encoding(lib, e), whole[1]), whole[3])
    w = CodeDecommittmentsDeduplicatorInstanceWitness(io, u, up, s, sp, whole[5])

    def run(io_, u_, up_, s_, sp_, states, lim, want_trace):
        rc, io2, trace, com, st, _ = O.sort_decommittments_entry_point(lib, io_, u_, s_, lim, want_trace=want_trace)
        return SimpleNamespace(closed_form_input=io2, trace=trace, status=st, commitment=com)

    _, scale, commit = _oracle_backend(lib)
    cum = np.concatenate([[0], np.cumsum(whole[2][abi.DQ_COLS["ADD_TO_QUEUE"]])]).astype(np.int64)
    for world in (2, 4):
        offs = [int(cum[sharding.row_range(n, r, world)[0]]) for r in range(world)]
        locs = [sharding.decommit_rows_local(run, w, limit, r, world, offs) for r in range(world)]
        recs = np.stack([l[3] for l in locs])
        traces = []
        for r in range(world):
            com, io_g, trace, st = sharding.decommit_rows_finish(locs[r][0], r, world, recs, io, offs, scale, commit)
            assert st.code == 0 and np.array_equal(com, whole[3])
            assert bytes(io_g.hidden_fsm_output) == bytes(whole[1].hidden_fsm_output) and bytes(io_g.final_queue_state) == bytes(whole[1].final_queue_state)
            traces.append(trace)
        assert np.array_equal(np.concatenate(traces, axis=1), whole[2])


def test_row_sharded_ram_permutation_virtual_ranks():
    """ONE ram_permutation instance cut by rows: no result queue, but a COUNTER in the FSM record (num_nondeterministic_writes): every
    rank counts from 0 and the column gets the lower ranks' counts added after the exchange.  1800 of the 3000 queries are
    non-deterministic bootloader writes, so every cut has writes on both sides; 2 and 4 virtual ranks, oracle backend."""
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    from types import SimpleNamespace
    import helpers as H
    import orc as O
    from era_zkevm_circuits_b200 import RamPermutationCircuitInstanceWitness, abi, sharding, synthetic
    lib = O.load()
    n, limit = 3000, 3100
    u, s = synthetic.ram_trace(n, seed=21, n_cells=50, n_nondet=1800)
    io, up, sp = H.ram_instance(lib, u, s, 1800)
    whole = O.ram_entry_point(lib, io, u, s, limit)
    assert whole[0] == 0 and whole[1].hidden_fsm_output.num_nondeterministic_writes == 1800
    commit = lambda e: O.commit_encoding(lib, e)
    assert np.array_equal(sharding.ram_closed_form_commitment(commit, whole[1]), whole[3])
    w = RamPermutationCircuitInstanceWitness(io, u, up, s, sp)

    def run(io_, u_, up_, s_, sp_, lim, want_trace):
        rc, io2, trace, com, st = O.ram_entry_point(lib, io_, u_, s_, lim, want_trace=want_trace)
        return SimpleNamespace(closed_form_input=io2, trace=trace, status=st, commitment=com)

    _, scale, _ = _oracle_backend(lib)
    for world in (2, 4):
        locs = [sharding.ram_rows_local(run, w, limit, r, world) for r in range(world)]
        recs = np.stack([l[3] for l in locs])
        assert int(recs[:, 11].sum()) == 1800 and np.count_nonzero(recs[:, 11]) >= 2  # counts on both sides of a cut
        traces = []
        for r in range(world):
            com, io_g, trace, st = sharding.ram_rows_finish(locs[r][0], r, world, recs, io, scale, commit)
            assert st.code == 0 and np.array_equal(com, whole[3]) and bytes(io_g.hidden_fsm_output) == bytes(whole[1].hidden_fsm_output)
            traces.append(trace)
        assert np.array_equal(np.concatenate(traces, axis=1), whole[2])
    # a wrong snapshot length is found on the exchanged counts (the per-rank check is masked: only the sum is meaningful)
    io_bad = abi.RamClosedForm.from_buffer_copy(bytes(io)); io_bad.observable_input.non_deterministic_bootloader_memory_snapshot_length = 1799
    w_bad = RamPermutationCircuitInstanceWitness(io_bad, u, up, s, sp)
    locs = [sharding.ram_rows_local(run, w_bad, limit, r, 2) for r in range(2)]
    com, io_g, trace, st = sharding.ram_rows_finish(locs[0][0], 0, 2, np.stack([l[3] for l in locs]), io_bad, scale, commit)
    assert st.code == abi.ZKC_ERR_UNSATISFIED and st.failed_checks == abi.RAM_CHK["NONDET_COUNT"]
