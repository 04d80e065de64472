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


# ---- ONE storage_validity instance cut over the ranks by row range -----------------------------------------
# The loop of sort_and_deduplicate_storage_access_inner (storage_validity_by_grand_product/mod.rs:560-800) carries a small
# state from row to row: the two grand-product accumulators per repetition, the three queue states, the cycle counter, the
# previous key / timestamp and the state of the storage cell under construction (input.rs:37-52: exactly the hidden FSM record
# the reference itself uses to chain circuit instances).  Cut at row `lo`, everything in that record except the accumulators is
# known WITHOUT running rows [0, lo):
#   * queue heads / lengths: the host's queue-state hints (prev_tails[lo]) and `lo` itself;
#   * the result queue: the host's tail hints, indexed by the number of pushes before the cut (push_offsets, a hint like the
#     tails: verified after the exchange against the counts the ranks measured);
#   * previous key / timestamp: the sorted record at lo - 1;
#   * the cell state: a cell's state is reset by its first row, so replaying the rows of the cell that straddles the cut,
#     [cell_start, lo), reproduces it exactly -- a short run through the same entry point;
#   * the accumulators are running products: every rank starts from 1 and the columns are scaled after ONE all-gather.
# Each rank then runs the stock entry point over its rows as a chained instance; the exchange carries {4 products, push
# count, status, the FSM output record} per rank; the fix-up is 8 columns (GP_NEW, GP_ACC) times 4 field elements.

def _np(x):
    """host copy of a numpy array / torch tensor slice"""
    return x.cpu().numpy() if type(x).__module__.startswith("torch") else np.asarray(x)


def _qs4(head, tail, length):
    from . import abi
    q = abi.QueueState4()
    for i in range(4):
        q.head[i] = int(head[i]); q.tail[i] = int(tail[i])
    q.length = int(length)
    return q


def storage_start_state(io):
    """the selection the entry point makes between the observable input and the hidden FSM input (mod.rs:190-395)"""
    f = io.hidden_fsm_input
    start = bool(io.start_flag)
    uq0 = io.unsorted_log_queue_state if start else f.current_unsorted_queue_state
    sq0 = io.intermediate_sorted_queue_state if start else f.current_intermediate_sorted_queue_state
    rq0 = _qs4([0] * 4, [0] * 4, 0) if start else f.current_final_sorted_queue_state
    acc0 = [1, 1, 1, 1] if start else [int(f.lhs_accumulator[0]), int(f.rhs_accumulator[0]), int(f.lhs_accumulator[1]), int(f.rhs_accumulator[1])]
    return uq0, sq0, rq0, acc0, (0 if start else int(f.cycle_idx))


def storage_fsm_encoding(f):
    """77 field elements, StorageDeduplicatorFSMInputOutput's encoding order (input.rs:37-52)"""
    e = [f.lhs_accumulator[0], f.lhs_accumulator[1], f.rhs_accumulator[0], f.rhs_accumulator[1]]
    for q in (f.current_unsorted_queue_state, f.current_intermediate_sorted_queue_state, f.current_final_sorted_queue_state):
        e += list(q.head) + list(q.tail) + [q.length]
    e += [f.cycle_idx] + list(f.previous_packed_key) + list(f.previous_key) + list(f.previous_address) + [f.previous_timestamp]
    e += [f.this_cell_has_explicit_read_and_rollback_depth_zero] + list(f.this_cell_base_value) + list(f.this_cell_current_value) + [f.this_cell_current_depth]
    return np.array([int(x) for x in e], dtype=np.uint64)


def storage_closed_form_commitment(commit_fn, io):
    """ClosedFormInputCompactForm::from_full_form + commit (fsm_input_output/mod.rs:281-326) of a finished storage_validity closed
    form; commit_fn(elements [len] uint64) -> [4] is the variable-length Poseidon2 commitment (Engine.commit_encoding)."""
    qs = lambda q: list(q.head) + list(q.tail) + [q.length]
    obs_in = np.array([io.shard_id_to_process & 0xFF] + qs(io.unsorted_log_queue_state) + qs(io.intermediate_sorted_queue_state), dtype=np.uint64)
    compact = np.zeros(18, dtype=np.uint64)
    compact[0], compact[1] = int(bool(io.start_flag)), int(bool(io.completion_flag))
    compact[2:6] = commit_fn(obs_in)
    if io.completion_flag:
        compact[6:10] = commit_fn(np.array(qs(io.final_sorted_queue_state), dtype=np.uint64))
    else:
        compact[14:18] = commit_fn(storage_fsm_encoding(io.hidden_fsm_output))
    if not io.start_flag:
        compact[10:14] = commit_fn(storage_fsm_encoding(io.hidden_fsm_input))
    return commit_fn(compact)


_ST_LOCAL_WORDS = 16  # int64 words of the fixed part of a rank's exchange record


def storage_rows_local(run_fn, witness, limit, rank, world, push_offsets):
    """Phase 1 on rank `rank`: derive the FSM input at the rank's first row, run the rank's rows through the entry point.
    run_fn(io, unsorted, unsorted_prev_tails, sorted, sorted_ts, sorted_prev_tails, result_tails, limit, want_trace) ->
    SorterResult-like (closed_form_input, trace, status).  Returns (result of the rank's rows, lo, hi, exchange record int64)."""
    import ctypes as C
    from . import abi
    from .storage_validity import ST_CHK_GRAND_PRODUCT
    w = witness
    io0 = w.closed_form_input
    uq0, sq0, rq0, acc0, cycle0 = storage_start_state(io0)
    n_active = min(limit, int(uq0.length), int(sq0.length), len(w.unsorted_queue_witness), len(w.intermediate_sorted_queue_witness))
    if world > 1 and (w.result_queue_tails is None or push_offsets is None or len(push_offsets) != world):
        raise ValueError("a row-sharded run needs the result-queue tail hints and one push offset per rank")
    if world > 1 and n_active < world:
        raise ValueError("fewer active rows than ranks")
    lo, hi = row_range(n_active, rank, world)
    if rank == world - 1:
        hi = limit  # the padding rows after the queues run empty stay with the last rank
    k_lo = int(push_offsets[rank]) if world > 1 else 0
    io = abi.StorageClosedForm.from_buffer_copy(bytes(io0))
    sl = lambda a, a0, a1: None if a is None else a[a0:a1]
    if rank > 0:
        # the cell that straddles the cut: rows [cs, lo) share the packed key (address, key) of row lo - 1
        cs = lo - 1
        win = 256
        key_of = lambda recs: np.concatenate([_np(recs)["address"] if _np(recs).dtype.names else _np(recs)[:, :20].view(np.uint32),
                                              _np(recs)["key"] if _np(recs).dtype.names else _np(recs)[:, 20:52].view(np.uint32)], axis=1)
        while True:
            a0 = max(0, lo - win)
            keys = key_of(w.intermediate_sorted_queue_witness[a0:lo])
            same = (keys == keys[-1]).all(axis=1)
            first_other = np.flatnonzero(~same)
            if len(first_other) or a0 == 0:
                cs = a0 + (int(first_other[-1]) + 1 if len(first_other) else 0)
                break
            win *= 4
        mini = abi.StorageClosedForm.from_buffer_copy(bytes(io0))
        if cs > 0:
            mini.start_flag = 0
            f = mini.hidden_fsm_input
            C.memset(C.byref(f), 0, C.sizeof(f))
            f.lhs_accumulator[0] = f.lhs_accumulator[1] = f.rhs_accumulator[0] = f.rhs_accumulator[1] = 1
            f.current_unsorted_queue_state = _qs4(_np(w.unsorted_queue_prev_tails[cs:cs + 1]).view(np.uint64)[0], uq0.tail, int(uq0.length) - cs)
            f.current_intermediate_sorted_queue_state = _qs4(_np(w.intermediate_sorted_queue_prev_tails[cs:cs + 1]).view(np.uint64)[0], sq0.tail, int(sq0.length) - cs)
            f.cycle_idx = cycle0 + cs
            prev = key_of(w.intermediate_sorted_queue_witness[cs - 1:cs])[0]
            for i in range(5):
                f.previous_address[i] = int(prev[i]); f.previous_packed_key[8 + i] = int(prev[i])
            for i in range(8):
                f.previous_key[i] = int(prev[5 + i]); f.previous_packed_key[i] = int(prev[5 + i])
            f.previous_timestamp = int(_np(w.intermediate_sorted_queue_timestamps[cs - 1:cs]).view(np.uint32)[0])
        m = run_fn(mini, sl(w.unsorted_queue_witness, cs, lo), sl(w.unsorted_queue_prev_tails, cs, lo), sl(w.intermediate_sorted_queue_witness, cs, lo),
                   sl(w.intermediate_sorted_queue_timestamps, cs, lo), sl(w.intermediate_sorted_queue_prev_tails, cs, lo), None, lo - cs, False)
        if m.status.code not in (abi.ZKC_OK, abi.ZKC_ERR_UNSATISFIED):
            raise RuntimeError(f"replay of the straddling cell [{cs}, {lo}) failed: code {m.status.code}")
        io.start_flag = 0
        io.hidden_fsm_input = m.closed_form_input.hidden_fsm_output
        f = io.hidden_fsm_input
        f.lhs_accumulator[0] = f.lhs_accumulator[1] = f.rhs_accumulator[0] = f.rhs_accumulator[1] = 1
        tail = _np(w.result_queue_tails[k_lo - 1:k_lo]).view(np.uint64)[0] if k_lo > 0 else rq0.tail
        f.current_final_sorted_queue_state = _qs4(rq0.head, tail, int(rq0.length) + k_lo)
    res = run_fn(io, sl(w.unsorted_queue_witness, lo, hi), sl(w.unsorted_queue_prev_tails, lo, hi), sl(w.intermediate_sorted_queue_witness, lo, hi),
                 sl(w.intermediate_sorted_queue_timestamps, lo, hi), sl(w.intermediate_sorted_queue_prev_tails, lo, hi),
                 None if w.result_queue_tails is None else w.result_queue_tails[k_lo:], hi - lo, True)
    out = res.closed_form_input.hidden_fsm_output
    st = res.status
    failed = int(st.failed_checks)
    if world > 1:
        failed &= ~ST_CHK_GRAND_PRODUCT  # lhs == rhs holds for the WHOLE loop: re-evaluated on the exchanged products
    code = int(st.code) if (failed or st.code != abi.ZKC_ERR_UNSATISFIED) else 0
    fsm_bytes = np.frombuffer(bytes(out) + bytes(res.closed_form_input.final_sorted_queue_state), dtype=np.uint8)
    pad = (-len(fsm_bytes)) % 8
    rec = np.zeros(_ST_LOCAL_WORDS + (len(fsm_bytes) + pad) // 8, dtype=np.int64)
    loc = np.array([out.lhs_accumulator[0], out.rhs_accumulator[0], out.lhs_accumulator[1], out.rhs_accumulator[1]], dtype=np.uint64)
    f_in = io.hidden_fsm_input
    pushes_in = int(f_in.current_final_sorted_queue_state.length) if not io.start_flag else 0
    rec[0:4] = loc.view(np.int64)
    rec[4] = int(out.current_final_sorted_queue_state.length) - pushes_in  # pushes of this rank's rows (+ the finalisation push on the last)
    rec[5], rec[6], rec[7] = code, failed, int(st.first_bad_row) + lo if st.first_bad_row >= 0 else -1
    rec[8] = int(res.closed_form_input.completion_flag)
    rec[9], rec[10] = lo, hi
    rec[_ST_LOCAL_WORDS:] = np.concatenate([fsm_bytes, np.zeros(pad, np.uint8)]).view(np.int64)
    return res, lo, hi, rec


def storage_rows_finish(res, rank, world, records, io0, push_offsets, scale_fn, commit_fn):
    """Phase 2: `records` [world, words] = every rank's exchange record.  Scales this rank's GP_NEW / GP_ACC columns by the
    product of everything before its rows, verifies the push offsets, assembles the WHOLE instance's closed form, status and
    commitment (identical on every rank).  Returns (commitment, closed form, trace of this rank's rows, status)."""
    import ctypes as C
    from . import abi
    from .storage_validity import ST_CHK_GRAND_PRODUCT
    records = np.asarray(records, dtype=np.int64).reshape(world, -1)
    totals = records[:, 0:4].copy().view(np.uint64)
    seed = [1, 1, 1, 1]  # rank 0 runs from the instance's own accumulators: its totals carry them
    for r in range(rank):
        seed = [s * int(t) % GL_P for s, t in zip(seed, totals[r])]
    K = abi.ST_COLS
    if world > 1 and rank > 0 and res.trace is not None and any(s != 1 for s in seed):
        scale_fn(res.trace[K["GP_NEW"]:K["GP_NEW"] + 4], np.array(seed, dtype=np.uint64))
        scale_fn(res.trace[K["GP_ACC"]:K["GP_ACC"] + 4], np.array(seed, dtype=np.uint64))
    io = abi.StorageClosedForm.from_buffer_copy(bytes(io0))
    last = records[world - 1]
    fsm_len = C.sizeof(abi.StorageFsm)
    tail_bytes = last[_ST_LOCAL_WORDS:].view(np.uint8)
    io.hidden_fsm_output = abi.StorageFsm.from_buffer_copy(tail_bytes[:fsm_len].tobytes())
    io.final_sorted_queue_state = abi.QueueState4.from_buffer_copy(tail_bytes[fsm_len:fsm_len + C.sizeof(abi.QueueState4)].tobytes())
    io.completion_flag = int(last[8])
    st = abi.Status()
    st.first_bad_row = -1
    if world > 1:
        grand = [1, 1, 1, 1]
        for r in range(world):
            grand = [g * int(t) % GL_P for g, t in zip(grand, totals[r])]
        out = io.hidden_fsm_output
        out.lhs_accumulator[0], out.rhs_accumulator[0], out.lhs_accumulator[1], out.rhs_accumulator[1] = grand
        failed, code, first_bad = 0, 0, -1
        for r in range(world):
            failed |= int(records[r, 6])
            if records[r, 5] and not code:
                code = int(records[r, 5])
            if records[r, 7] >= 0 and first_bad < 0:
                first_bad = int(records[r, 7])
            expect = int(push_offsets[r + 1]) if r + 1 < world else None
            if expect is not None and int(push_offsets[r]) + int(records[r, 4]) != expect:
                failed |= abi.ST_CHK["QUEUE_HINT"]
                code = code or abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT
        if io.completion_flag and (grand[0] != grand[1] or grand[2] != grand[3]):
            failed |= ST_CHK_GRAND_PRODUCT
        if failed and not code:
            code = abi.ZKC_ERR_UNSATISFIED
        st.code, st.failed_checks, st.first_bad_row = code, failed, first_bad
        commitment = storage_closed_form_commitment(commit_fn, io)
    else:
        st = res.status
        commitment = res.commitment
    return commitment, io, res.trace, st


def storage_validity_row_sharded(engine, witness, limit, rank, world, push_offsets=None, device=None):
    """sort_and_deduplicate_storage_access_entry_point of ONE instance over `world` ranks (one process per GPU, torch.distributed
    initialised by the caller): phase 1, the single all-gather, phase 2.  Every rank holds (or can slice) the whole witness;
    push_offsets[r] = result-queue pushes in the rows before rank r's first row (row_range over the active rows)."""
    import torch
    import torch.distributed as dist
    from .log_sorter import SorterResult
    from .storage_validity import StorageDeduplicatorInstanceWitness, sort_and_deduplicate_storage_access_entry_point

    def run_fn(io, u, up, s, ts, sp, tails, lim, want_trace):
        wit = StorageDeduplicatorInstanceWitness(io, u, up, s, ts, sp, tails)
        return sort_and_deduplicate_storage_access_entry_point(engine, wit, lim, want_trace=want_trace, raise_on_unsatisfied=False)

    res, lo, hi, rec = storage_rows_local(run_fn, witness, limit, rank, world, push_offsets)
    mine = torch.from_numpy(rec.copy())
    if device is not None:
        mine = mine.to(device)
    allr = torch.zeros((world, len(rec)), dtype=torch.int64, device=mine.device)
    if world > 1:
        dist.all_gather_into_tensor(allr, mine.reshape(1, -1))
    else:
        allr.copy_(mine.reshape(1, -1))
    commit_fn = lambda e: engine.commit_encoding(np.ascontiguousarray(e, dtype=np.uint64).reshape(1, -1))[0]
    com, io, trace, st = storage_rows_finish(res, rank, world, allr.cpu().numpy(), witness.closed_form_input, push_offsets,
                                             engine.scale_accumulators, commit_fn)
    return SorterResult(com, io, trace, st), (lo, hi)
