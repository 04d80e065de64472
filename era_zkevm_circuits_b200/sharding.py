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


# ---- ONE sorter instance cut over the ranks by row range -----------------------------------------------------
# The loops of sort_and_deduplicate_storage_access_inner (storage_validity_by_grand_product/mod.rs:560-800) and
# repack_and_prove_events_rollbacks_inner (log_sorter/mod.rs:234-420) carry a small state from row to row: the two grand-product
# accumulators per repetition, the three queue states, the previous key / item and -- storage only -- the cycle counter and the
# state of the storage cell under construction (input.rs:37-52 / log_sorter/input.rs:26-38: exactly the hidden FSM record the
# reference itself uses to chain circuit instances).  Cut at row `lo`, everything in that record except the accumulators is known
# WITHOUT running rows [0, lo):
#   * queue heads / lengths: the host's queue-state hints (prev_tails[lo]) and `lo` itself;
#   * the result queue: the host's tail hints, indexed by the number of pushes before the cut (push_offsets, a hint like the
#     tails: verified after the exchange against the counts the ranks measured);
#   * previous key / item / timestamp and the storage cell state: a cell's state is reset by its first row, so replaying the rows
#     of the cell that straddles the cut, [cell_start, lo) -- for log_sorter just row lo - 1 -- through the same entry point
#     reproduces them exactly (its FSM output IS the record, accumulators and result queue apart);
#   * the accumulators are running products: every rank starts from 1 and the columns are scaled after ONE all-gather.
# Each rank then runs the stock entry point over its rows as a chained instance; the exchange carries {4 products, push
# count, status, the FSM output record} per rank; the fix-up is 8 columns (GP_NEW, GP_ACC) times 4 field elements.

def _np(x):
    """host copy of a numpy array / torch tensor slice"""
    return x.cpu().numpy() if type(x).__module__.startswith("torch") else np.asarray(x)


def _qs(cls, head, tail, length):
    """a QueueState4 / QueueState12 from head / tail element sequences"""
    q = cls()
    for i in range(len(q.head)):
        q.head[i] = int(head[i]); q.tail[i] = int(tail[i])
    q.length = int(length)
    return q


def _qs4(head, tail, length):
    from . import abi
    return _qs(abi.QueueState4, head, tail, length)


def _qs_list(q):
    return list(q.head) + list(q.tail) + [q.length]


def _packed_keys(recs):
    """[m, 13] u32: address (5 limbs) then key (8 limbs) of LogQuery records (numpy structured or byte rows)"""
    a = _np(recs)
    if a.dtype.names:
        return np.concatenate([a["address"], a["key"]], axis=1)
    return np.concatenate([a[:, :20].view(np.uint32), a[:, 20:52].view(np.uint32)], axis=1)


class _StorageCut:
    """storage_validity_by_grand_product: field names, the replay window, the encodings of the commitments"""
    name = "storage_validity"
    fsm_queues = ("current_unsorted_queue_state", "current_intermediate_sorted_queue_state", "current_final_sorted_queue_state")
    obs_queues = ("unsorted_log_queue_state", "intermediate_sorted_queue_state")
    final_queue = "final_sorted_queue_state"

    def __init__(self):
        from . import abi
        self.ClosedForm, self.Fsm, self.cols, self.chk, self.Queue = abi.StorageClosedForm, abi.StorageFsm, abi.ST_COLS, abi.ST_CHK, abi.QueueState4

    def result_tails(self, w):
        return w.result_queue_tails

    def arrays(self, w):
        """(unsorted records, unsorted prev tails, sorted records, sorted prev tails, extra per-row arrays of the sorted queue)"""
        return w.unsorted_queue_witness, w.unsorted_queue_prev_tails, w.intermediate_sorted_queue_witness, w.intermediate_sorted_queue_prev_tails, \
            (w.intermediate_sorted_queue_timestamps,)

    def cycle0(self, io):
        return 0 if io.start_flag else int(io.hidden_fsm_input.cycle_idx)

    def replay_start(self, w, lo):
        """first row of the cell that holds row lo - 1: rows [cs, lo) share its packed key (address, key)"""
        win = 256
        while True:
            a0 = max(0, lo - win)
            keys = _packed_keys(w.intermediate_sorted_queue_witness[a0:lo])
            other = np.flatnonzero(~(keys == keys[-1]).all(axis=1))
            if len(other) or a0 == 0:
                return a0 + (int(other[-1]) + 1 if len(other) else 0)
            win *= 4

    def fill_replay_fsm(self, f, w, io0, cs):
        f.cycle_idx = self.cycle0(io0) + cs
        prev = _packed_keys(w.intermediate_sorted_queue_witness[cs - 1:cs])[0]
        for i in range(5):
            f.previous_address[i] = int(prev[i]); f.previous_packed_key[8 + i] = int(prev[i])
        for i in range(8):
            f.previous_key[i] = int(prev[5 + i]); f.previous_packed_key[i] = int(prev[5 + i])
        f.previous_timestamp = int(_np(w.intermediate_sorted_queue_timestamps[cs - 1:cs]).view(np.uint32)[0])

    def fsm_encoding(self, f):
        """77 field elements, StorageDeduplicatorFSMInputOutput's encoding order (input.rs:37-52)"""
        e = [f.lhs_accumulator[0], f.lhs_accumulator[1], f.rhs_accumulator[0], f.rhs_accumulator[1]]
        for q in self.fsm_queues:
            e += _qs_list(getattr(f, q))
        e += [f.cycle_idx] + list(f.previous_packed_key) + list(f.previous_key) + list(f.previous_address) + [f.previous_timestamp]
        e += [f.this_cell_has_explicit_read_and_rollback_depth_zero] + list(f.this_cell_base_value) + list(f.this_cell_current_value) + [f.this_cell_current_depth]
        return np.array([int(x) for x in e], dtype=np.uint64)

    def obs_in_encoding(self, io):
        return np.array([io.shard_id_to_process & 0xFF] + _qs_list(io.unsorted_log_queue_state) + _qs_list(io.intermediate_sorted_queue_state), dtype=np.uint64)


class _EventsCut:
    """log_sorter: the FSM record is the previous key and item -- the replay window is row lo - 1 alone"""
    name = "log_sorter"
    fsm_queues = ("initial_unsorted_queue_state", "intermediate_sorted_queue_state", "final_result_queue_state")
    obs_queues = ("initial_log_queue_state", "intermediate_sorted_queue_state")
    final_queue = "final_queue_state"

    def __init__(self):
        from . import abi
        self.ClosedForm, self.Fsm, self.cols, self.chk, self.Queue = abi.EventsClosedForm, abi.EventsFsm, abi.EV_COLS, abi.EV_CHK, abi.QueueState4

    def result_tails(self, w):
        return w.result_queue_tails

    def arrays(self, w):
        return w.initial_queue_witness, w.initial_queue_prev_tails, w.intermediate_sorted_queue_witness, w.intermediate_sorted_queue_prev_tails, ()

    def replay_start(self, w, lo):
        return lo - 1

    def fill_replay_fsm(self, f, w, io0, cs):
        pass

    def fsm_encoding(self, f):
        """68 field elements (log_sorter/input.rs:26-38): accumulators, 3 queue states, previous key, the previous item's 36 variables"""
        from . import abi
        e = [f.lhs_accumulator[0], f.lhs_accumulator[1], f.rhs_accumulator[0], f.rhs_accumulator[1]]
        for q in self.fsm_queues:
            e += _qs_list(getattr(f, q))
        p = f.previous_item
        fl = int(p.flags)
        e += [f.previous_key] + list(p.address) + list(p.key) + list(p.read_value) + list(p.written_value)
        e += [fl & 0xFF, (fl >> 16) & 1, (fl >> 17) & 1, (fl >> 18) & 1, (fl >> 8) & 0xFF, p.tx_number_in_block, p.timestamp]
        return np.array([int(x) for x in e], dtype=np.uint64)

    def obs_in_encoding(self, io):
        return np.array(_qs_list(io.initial_log_queue_state) + _qs_list(io.intermediate_sorted_queue_state), dtype=np.uint64)


class _DecommitCut:
    """sort_decommittment_requests: full-state (12-element) queues; the carried first-encountered timestamp belongs to the run of
    equal code hashes that straddles the cut, so the replay window is that run (as the storage cell)"""
    name = "sort_decommittment_requests"
    fsm_queues = ("initial_queue_state", "sorted_queue_state", "final_queue_state")
    obs_queues = ("initial_queue_state", "sorted_queue_initial_state")
    final_queue = "final_queue_state"

    def __init__(self):
        from . import abi
        self.ClosedForm, self.Fsm, self.cols, self.chk, self.Queue = abi.DecommitSorterClosedForm, abi.DecommitSorterFsm, abi.DQ_COLS, abi.DQ_CHK, abi.QueueState12

    def arrays(self, w):
        return w.initial_queue_witness, w.initial_queue_prev_states, w.sorted_queue_witness, w.sorted_queue_prev_states, ()

    def result_tails(self, w):
        return w.result_queue_states

    @staticmethod
    def _fields(recs):
        """[m, 11] u32: code_hash (8 limbs), page, is_first, timestamp of DecommitQuery records (numpy structured or byte rows)"""
        a = _np(recs)
        if a.dtype.names:
            return np.concatenate([a["code_hash"], a["page"][:, None], a["is_first"][:, None], a["timestamp"][:, None]], axis=1).astype(np.uint32)
        return a[:, :44].view(np.uint32)

    def replay_start(self, w, lo):
        win = 256
        while True:
            a0 = max(0, lo - win)
            h = self._fields(w.sorted_queue_witness[a0:lo])[:, :8]
            other = np.flatnonzero(~(h == h[-1]).all(axis=1))
            if len(other) or a0 == 0:
                return a0 + (int(other[-1]) + 1 if len(other) else 0)
            win *= 4

    def fill_replay_fsm(self, f, w, io0, cs):
        prev = self._fields(w.sorted_queue_witness[cs - 1:cs])[0]
        f.previous_packed_key[0] = int(prev[10])
        for i in range(8):
            f.previous_packed_key[1 + i] = int(prev[i]); f.previous_record.code_hash[i] = int(prev[i])
        f.previous_record.page, f.previous_record.is_first, f.previous_record.timestamp = int(prev[8]), int(prev[9]) & 1, int(prev[10])

    def fsm_encoding(self, f):
        """100 field elements (sort_decommittment_requests/input.rs:26-37)"""
        e = []
        for q in self.fsm_queues:
            e += _qs_list(getattr(f, q))
        e += [f.lhs_accumulator[0], f.lhs_accumulator[1], f.rhs_accumulator[0], f.rhs_accumulator[1]]
        p = f.previous_record
        e += list(f.previous_packed_key) + [f.first_encountered_timestamp] + list(p.code_hash) + [p.page, p.is_first & 1, p.timestamp]
        return np.array([int(x) for x in e], dtype=np.uint64)

    def obs_in_encoding(self, io):
        return np.array(_qs_list(io.initial_queue_state) + _qs_list(io.sorted_queue_initial_state), dtype=np.uint64)


class _RamCut:
    """ram_permutation: no result queue; besides the accumulators the loop carries a COUNTER (num_nondeterministic_writes, :260-290):
    every rank counts from 0 and the column / FSM field get the sum of the lower ranks' counts added after the exchange"""
    name = "ram_permutation"
    fsm_queues = ("current_unsorted_queue_state", "current_sorted_queue_state")
    final_queue = None
    counters = (("NUM_NONDET_WRITES", "num_nondeterministic_writes"),)  # (trace column, FSM field)

    def __init__(self):
        from . import abi
        self.ClosedForm, self.Fsm, self.cols, self.chk, self.Queue = abi.RamClosedForm, abi.RamFsm, abi.RAM_COLS, abi.RAM_CHK, abi.QueueState12

    def obs_queue(self, io, k):
        return io.observable_input.sorted_queue_initial_state if k else io.observable_input.unsorted_queue_initial_state

    def arrays(self, w):
        return w.unsorted_queue_witness, w.unsorted_queue_prev_states, w.sorted_queue_witness, w.sorted_queue_prev_states, ()

    def result_tails(self, w):
        return None

    def replay_start(self, w, lo):
        return lo - 1  # previous sorting key / value / is_ptr: the sorted record at lo - 1

    def fill_replay_fsm(self, f, w, io0, cs):
        pass

    def masked_checks(self):
        return self.chk["GRAND_PRODUCT"] | self.chk["NONDET_COUNT"]

    def completion_checks(self, io, grand):
        bad = 0
        if grand[0] != grand[1] or grand[2] != grand[3]:
            bad |= self.chk["GRAND_PRODUCT"]
        if int(io.hidden_fsm_output.num_nondeterministic_writes) != int(io.observable_input.non_deterministic_bootloader_memory_snapshot_length):
            bad |= self.chk["NONDET_COUNT"]
        return bad

    def fsm_encoding(self, f):
        """69 field elements (ram_permutation/input.rs:52-62)"""
        e = [f.lhs_accumulator[0], f.lhs_accumulator[1], f.rhs_accumulator[0], f.rhs_accumulator[1]]
        e += _qs_list(f.current_unsorted_queue_state) + _qs_list(f.current_sorted_queue_state)
        e += list(f.previous_sorting_key) + list(f.previous_full_key) + list(f.previous_value) + [f.previous_is_ptr, f.num_nondeterministic_writes]
        return np.array([int(x) for x in e], dtype=np.uint64)

    def obs_in_encoding(self, io):
        o = io.observable_input
        return np.array(_qs_list(o.unsorted_queue_initial_state) + _qs_list(o.sorted_queue_initial_state) +
                        [o.non_deterministic_bootloader_memory_snapshot_length], dtype=np.uint64)


def _obs_queue(cut, io, k):
    return cut.obs_queue(io, k) if hasattr(cut, "obs_queue") else getattr(io, cut.obs_queues[k])


def _has_result_queue(cut):
    return len(cut.fsm_queues) > 2


def _start_state(cut, io):
    """the selection the entry points make between the observable input and the hidden FSM input (storage mod.rs:190-395)"""
    f = io.hidden_fsm_input
    start = bool(io.start_flag)
    uq0 = _obs_queue(cut, io, 0) if start else getattr(f, cut.fsm_queues[0])
    sq0 = _obs_queue(cut, io, 1) if start else getattr(f, cut.fsm_queues[1])
    rq0 = None
    if _has_result_queue(cut):
        width = len(cut.Queue().head)
        rq0 = _qs(cut.Queue, [0] * width, [0] * width, 0) if start else getattr(f, cut.fsm_queues[2])
    return uq0, sq0, rq0


def storage_start_state(io):
    return _start_state(_StorageCut(), io)


def closed_form_commitment(cut, commit_fn, io):
    """ClosedFormInputCompactForm::from_full_form + commit (fsm_input_output/mod.rs:281-326) of a finished closed form;
    commit_fn(elements [len] uint64) -> [4] is the variable-length Poseidon2 commitment (Engine.commit_encoding)."""
    compact = np.zeros(18, dtype=np.uint64)
    compact[0], compact[1] = int(bool(io.start_flag)), int(bool(io.completion_flag))
    compact[2:6] = commit_fn(cut.obs_in_encoding(io))
    if io.completion_flag:
        if cut.final_queue is not None:  # an observable output of `()` commits to zeros
            compact[6:10] = commit_fn(np.array(_qs_list(getattr(io, cut.final_queue)), dtype=np.uint64))
    else:
        compact[14:18] = commit_fn(cut.fsm_encoding(io.hidden_fsm_output))
    if not io.start_flag:
        compact[10:14] = commit_fn(cut.fsm_encoding(io.hidden_fsm_input))
    return commit_fn(compact)


def storage_closed_form_commitment(commit_fn, io):
    return closed_form_commitment(_StorageCut(), commit_fn, io)


def events_closed_form_commitment(commit_fn, io):
    return closed_form_commitment(_EventsCut(), commit_fn, io)


_LOCAL_WORDS = 16  # int64 words of the fixed part of a rank's exchange record


def rows_local(cut, run_fn, witness, limit, rank, world, push_offsets):
    """Phase 1 on rank `rank`: derive the FSM input at the rank's first row, run the rank's rows through the entry point.
    run_fn(io, unsorted, unsorted_prev_tails, sorted, sorted_prev_tails, extras, result_tails, limit, want_trace) ->
    SorterResult-like (closed_form_input, trace, status, commitment).  Returns (result of the rank's rows, lo, hi, exchange
    record int64)."""
    import ctypes as C
    from . import abi
    w = witness
    io0 = w.closed_form_input
    uq0, sq0, rq0 = _start_state(cut, io0)
    u, up, s, sp, extras = cut.arrays(w)
    n_active = min(limit, int(uq0.length), int(sq0.length), len(u), len(s))
    tails_hint = cut.result_tails(w)
    has_rq = _has_result_queue(cut)
    if world > 1 and has_rq and (tails_hint is None or push_offsets is None or len(push_offsets) != world):
        raise ValueError("a row-sharded run needs the result-queue tail hints and one push offset per rank")
    if world > 1 and n_active < world:
        raise ValueError("fewer active rows than ranks")
    lo, hi = row_range(n_active, rank, world)
    if rank == world - 1:
        hi = limit  # the padding rows after the queues run empty stay with the last rank
    k_lo = int(push_offsets[rank]) if (world > 1 and has_rq) else 0
    io = cut.ClosedForm.from_buffer_copy(bytes(io0))
    sl = lambda a, a0, a1: None if a is None else a[a0:a1]
    args = lambda a0, a1: (sl(u, a0, a1), sl(up, a0, a1), sl(s, a0, a1), sl(sp, a0, a1), tuple(sl(x, a0, a1) for x in extras))
    if rank > 0:
        cs = cut.replay_start(w, lo)
        mini = cut.ClosedForm.from_buffer_copy(bytes(io0))
        if cs > 0:
            mini.start_flag = 0
            f = mini.hidden_fsm_input
            C.memset(C.byref(f), 0, C.sizeof(f))
            f.lhs_accumulator[0] = f.lhs_accumulator[1] = f.rhs_accumulator[0] = f.rhs_accumulator[1] = 1
            setattr(f, cut.fsm_queues[0], _qs(cut.Queue, _np(up[cs:cs + 1]).view(np.uint64)[0], uq0.tail, int(uq0.length) - cs))
            setattr(f, cut.fsm_queues[1], _qs(cut.Queue, _np(sp[cs:cs + 1]).view(np.uint64)[0], sq0.tail, int(sq0.length) - cs))
            cut.fill_replay_fsm(f, w, io0, cs)
        m = run_fn(mini, *args(cs, lo), None, lo - cs, False)
        if m.status.code not in (abi.ZKC_OK, abi.ZKC_ERR_UNSATISFIED):
            raise RuntimeError(f"replay of rows [{cs}, {lo}) failed: code {m.status.code}")
        io.start_flag = 0
        io.hidden_fsm_input = m.closed_form_input.hidden_fsm_output
        f = io.hidden_fsm_input
        f.lhs_accumulator[0] = f.lhs_accumulator[1] = f.rhs_accumulator[0] = f.rhs_accumulator[1] = 1
        if has_rq:
            tail = _np(tails_hint[k_lo - 1:k_lo]).view(np.uint64)[0] if k_lo > 0 else rq0.tail
            setattr(f, cut.fsm_queues[2], _qs(cut.Queue, rq0.head, tail, int(rq0.length) + k_lo))
        for _col, field in getattr(cut, "counters", ()):
            setattr(f, field, 0)  # counted from 0: the lower ranks' counts are added after the exchange
    res = run_fn(io, *args(lo, hi), None if tails_hint is None else tails_hint[k_lo:], hi - lo, True)
    out = res.closed_form_input.hidden_fsm_output
    st = res.status
    failed = int(st.failed_checks)
    if world > 1:  # lhs == rhs (and the counters' final values) hold for the WHOLE loop: re-evaluated after the exchange
        failed &= ~(cut.masked_checks() if hasattr(cut, "masked_checks") else cut.chk["GRAND_PRODUCT"])
    code = int(st.code) if (failed or st.code != abi.ZKC_ERR_UNSATISFIED) else 0
    fsm_bytes = np.frombuffer(bytes(out) + (bytes(getattr(res.closed_form_input, cut.final_queue)) if cut.final_queue else b""), dtype=np.uint8)
    pad = (-len(fsm_bytes)) % 8
    rec = np.zeros(_LOCAL_WORDS + (len(fsm_bytes) + pad) // 8, dtype=np.int64)
    loc = np.array([out.lhs_accumulator[0], out.rhs_accumulator[0], out.lhs_accumulator[1], out.rhs_accumulator[1]], dtype=np.uint64)
    rec[0:4] = loc.view(np.int64)
    if has_rq:
        pushes_in = int(getattr(io.hidden_fsm_input, cut.fsm_queues[2]).length) if not io.start_flag else 0
        rec[4] = int(getattr(out, cut.fsm_queues[2]).length) - pushes_in  # pushes of this rank's rows (+ the finalisation push on the last)
    for j, (_col, field) in enumerate(getattr(cut, "counters", ())):
        rec[11 + j] = int(getattr(out, field))  # rank 0 counts from the instance's own value, the others from 0
    rec[5], rec[6], rec[7] = code, failed, int(st.first_bad_row) + lo if st.first_bad_row >= 0 else -1
    rec[8] = int(res.closed_form_input.completion_flag)
    rec[9], rec[10] = lo, hi
    rec[_LOCAL_WORDS:] = np.concatenate([fsm_bytes, np.zeros(pad, np.uint8)]).view(np.int64)
    return res, lo, hi, rec


def rows_finish(cut, res, rank, world, records, io0, push_offsets, scale_fn, commit_fn):
    """Phase 2: `records` [world, words] = every rank's exchange record.  Scales this rank's GP_NEW / GP_ACC columns by the
    product of everything before its rows, verifies the push offsets, assembles the WHOLE instance's closed form, status and
    commitment (identical on every rank).  Returns (commitment, closed form, trace of this rank's rows, status)."""
    import ctypes as C
    from . import abi
    if world == 1:
        return res.commitment, res.closed_form_input, res.trace, res.status
    records = np.asarray(records, dtype=np.int64).reshape(world, -1)
    totals = records[:, 0:4].copy().view(np.uint64)
    seed = [1, 1, 1, 1]  # rank 0 runs from the instance's own accumulators: its totals carry them
    for r in range(rank):
        seed = [s * int(t) % GL_P for s, t in zip(seed, totals[r])]
    K = cut.cols
    if rank > 0 and res.trace is not None and any(s != 1 for s in seed):
        scale_fn(res.trace[K["GP_NEW"]:K["GP_NEW"] + 4], np.array(seed, dtype=np.uint64))
        scale_fn(res.trace[K["GP_ACC"]:K["GP_ACC"] + 4], np.array(seed, dtype=np.uint64))
    counters = getattr(cut, "counters", ())
    for j, (col, _field) in enumerate(counters):
        offset = int(sum(int(records[r, 11 + j]) for r in range(rank)))
        if offset and res.trace is not None:
            res.trace[K[col]] += offset  # the column is a running count: additive fix-up (numpy and torch both add in place)
    io = cut.ClosedForm.from_buffer_copy(bytes(io0))
    last = records[world - 1]
    fsm_len = C.sizeof(cut.Fsm)
    tail_bytes = last[_LOCAL_WORDS:].view(np.uint8)
    io.hidden_fsm_output = cut.Fsm.from_buffer_copy(tail_bytes[:fsm_len].tobytes())
    if cut.final_queue is not None:
        setattr(io, cut.final_queue, cut.Queue.from_buffer_copy(tail_bytes[fsm_len:fsm_len + C.sizeof(cut.Queue)].tobytes()))
    io.completion_flag = int(last[8])
    for j, (_col, field) in enumerate(counters):
        setattr(io.hidden_fsm_output, field, int(sum(int(records[r, 11 + j]) for r in range(world))))
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
        if _has_result_queue(cut) and r + 1 < world and int(push_offsets[r]) + int(records[r, 4]) != int(push_offsets[r + 1]):
            failed |= cut.chk["QUEUE_HINT"]
            code = code or abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT
    if io.completion_flag:
        if hasattr(cut, "completion_checks"):
            failed |= cut.completion_checks(io, grand)
        elif grand[0] != grand[1] or grand[2] != grand[3]:
            failed |= cut.chk["GRAND_PRODUCT"]
    if failed and not code:
        code = abi.ZKC_ERR_UNSATISFIED
    st = abi.Status()
    st.code, st.failed_checks, st.first_bad_row = code, failed, first_bad
    return closed_form_commitment(cut, commit_fn, io), io, res.trace, st


def storage_rows_local(run_fn, witness, limit, rank, world, push_offsets):
    """rows_local for storage_validity; run_fn(io, u, up, s, ts, sp, tails, limit, want_trace)"""
    return rows_local(_StorageCut(), lambda io, u, up, s, sp, ex, tails, lim, wt: run_fn(io, u, up, s, ex[0], sp, tails, lim, wt),
                      witness, limit, rank, world, push_offsets)


def storage_rows_finish(res, rank, world, records, io0, push_offsets, scale_fn, commit_fn):
    return rows_finish(_StorageCut(), res, rank, world, records, io0, push_offsets, scale_fn, commit_fn)


def events_rows_local(run_fn, witness, limit, rank, world, push_offsets):
    """rows_local for log_sorter; run_fn(io, u, up, s, sp, tails, limit, want_trace)"""
    return rows_local(_EventsCut(), lambda io, u, up, s, sp, ex, tails, lim, wt: run_fn(io, u, up, s, sp, tails, lim, wt),
                      witness, limit, rank, world, push_offsets)


def events_rows_finish(res, rank, world, records, io0, push_offsets, scale_fn, commit_fn):
    return rows_finish(_EventsCut(), res, rank, world, records, io0, push_offsets, scale_fn, commit_fn)


def decommit_rows_local(run_fn, witness, limit, rank, world, push_offsets):
    """rows_local for sort_decommittment_requests; run_fn(io, u, up, s, sp, states, limit, want_trace)"""
    return rows_local(_DecommitCut(), lambda io, u, up, s, sp, ex, tails, lim, wt: run_fn(io, u, up, s, sp, tails, lim, wt),
                      witness, limit, rank, world, push_offsets)


def decommit_rows_finish(res, rank, world, records, io0, push_offsets, scale_fn, commit_fn):
    return rows_finish(_DecommitCut(), res, rank, world, records, io0, push_offsets, scale_fn, commit_fn)


def decommit_sorter_closed_form_commitment(commit_fn, io):
    return closed_form_commitment(_DecommitCut(), commit_fn, io)


def ram_rows_local(run_fn, witness, limit, rank, world):
    """rows_local for ram_permutation (no result queue: no push offsets); run_fn(io, u, up, s, sp, limit, want_trace)"""
    return rows_local(_RamCut(), lambda io, u, up, s, sp, ex, tails, lim, wt: run_fn(io, u, up, s, sp, lim, wt), witness, limit, rank, world, None)


def ram_rows_finish(res, rank, world, records, io0, scale_fn, commit_fn):
    return rows_finish(_RamCut(), res, rank, world, records, io0, None, scale_fn, commit_fn)


def ram_closed_form_commitment(commit_fn, io):
    return closed_form_commitment(_RamCut(), commit_fn, io)


def _row_sharded(engine, local_fn, finish_fn, run_fn, witness, limit, rank, world, push_offsets, device):
    import torch
    import torch.distributed as dist
    from .log_sorter import SorterResult
    res, lo, hi, rec = local_fn(run_fn, witness, limit, rank, world, push_offsets)
    mine = torch.from_numpy(rec.copy())
    if device is not None:
        mine = mine.to(device)
    allr = torch.zeros((world, len(rec)), dtype=torch.int64, device=mine.device)
    if world > 1:
        dist.all_gather_into_tensor(allr, mine.reshape(1, -1))  # the path's ONE collective
    else:
        allr.copy_(mine.reshape(1, -1))
    commit_fn = lambda e: engine.commit_encoding(np.ascontiguousarray(e, dtype=np.uint64).reshape(1, -1))[0]
    com, io, trace, st = finish_fn(res, rank, world, allr.cpu().numpy(), witness.closed_form_input, push_offsets, engine.scale_accumulators, commit_fn)
    return SorterResult(com, io, trace, st), (lo, hi)


def storage_validity_row_sharded(engine, witness, limit, rank, world, push_offsets=None, device=None):
    """sort_and_deduplicate_storage_access_entry_point of ONE instance over `world` ranks (one process per GPU, torch.distributed
    initialised by the caller): phase 1, the single all-gather, phase 2.  Every rank holds (or can slice) the whole witness;
    push_offsets[r] = result-queue pushes in the rows before rank r's first row (row_range over the active rows)."""
    from .storage_validity import StorageDeduplicatorInstanceWitness, sort_and_deduplicate_storage_access_entry_point

    def run_fn(io, u, up, s, ts, sp, tails, lim, want_trace):
        wit = StorageDeduplicatorInstanceWitness(io, u, up, s, ts, sp, tails)
        return sort_and_deduplicate_storage_access_entry_point(engine, wit, lim, want_trace=want_trace, raise_on_unsatisfied=False)

    return _row_sharded(engine, storage_rows_local, storage_rows_finish, run_fn, witness, limit, rank, world, push_offsets, device)


def log_sorter_row_sharded(engine, witness, limit, rank, world, push_offsets=None, device=None):
    """sort_and_deduplicate_events_entry_point of ONE instance over `world` ranks, as storage_validity_row_sharded"""
    from .log_sorter import EventsDeduplicatorInstanceWitness, sort_and_deduplicate_events_entry_point

    def run_fn(io, u, up, s, sp, tails, lim, want_trace):
        wit = EventsDeduplicatorInstanceWitness(io, u, up, s, sp, tails)
        return sort_and_deduplicate_events_entry_point(engine, wit, lim, want_trace=want_trace, raise_on_unsatisfied=False)

    return _row_sharded(engine, events_rows_local, events_rows_finish, run_fn, witness, limit, rank, world, push_offsets, device)


def sort_decommittments_row_sharded(engine, witness, limit, rank, world, push_offsets=None, device=None):
    """sort_and_deduplicate_code_decommittments_entry_point of ONE instance over `world` ranks, as storage_validity_row_sharded"""
    from .sort_decommittment_requests import CodeDecommittmentsDeduplicatorInstanceWitness, sort_and_deduplicate_code_decommittments_entry_point

    def run_fn(io, u, up, s, sp, states, lim, want_trace):
        wit = CodeDecommittmentsDeduplicatorInstanceWitness(io, u, up, s, sp, states)
        return sort_and_deduplicate_code_decommittments_entry_point(engine, wit, lim, want_trace=want_trace, raise_on_unsatisfied=False)

    return _row_sharded(engine, decommit_rows_local, decommit_rows_finish, run_fn, witness, limit, rank, world, push_offsets, device)


def ram_permutation_row_sharded(engine, witness, limit, rank, world, device=None):
    """ram_permutation_entry_point of ONE instance over `world` ranks: as storage_validity_row_sharded, without a result queue and
    with the non-deterministic-write counter fixed up additively"""
    from .ram_permutation import RamPermutationCircuitInstanceWitness, ram_permutation_entry_point

    def run_fn(io, u, up, s, sp, lim, want_trace):
        return ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io, u, up, s, sp), lim, want_trace=want_trace, raise_on_unsatisfied=False)

    local = lambda rf, w, lim, r, n, _offs: ram_rows_local(rf, w, lim, r, n)
    finish = lambda res, r, n, recs, io0, _offs, sc, cm: ram_rows_finish(res, r, n, recs, io0, sc, cm)
    return _row_sharded(engine, local, finish, run_fn, witness, limit, rank, world, None, device)
