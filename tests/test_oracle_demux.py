"""demux_log_queue oracle against the reference's own vector (/root/reference/src/demux_log_queue/mod.rs:482-923,
limit = 16: every enforcement holds), the sorter-side queue simulation as an independent model of the six output queues,
chaining over instances and negative cases."""
import numpy as np

import orc as O
import vectors as V
from era_zkevm_circuits_b200 import abi, synthetic

K = abi.DMX_COLS
CHK = abi.DMX_CHK


def instance(orc, recs):
    prev, fin = O.log_queue_simulate(orc, recs)
    return O.demux_closed_form(fin, True), prev


def kind_of(recs):
    """which output queue each record belongs to (-1: none), independently of the oracle"""
    aux = recs["flags"] & 0xFF
    shard = (recs["flags"] >> 8) & 0xFF
    small = (recs["address"][:, 1:] == 0).all(axis=1)
    k = np.full(len(recs), -1)
    k[(aux == 0) & (shard == 0)] = 0
    k[aux == 1] = 1
    k[aux == 2] = 2
    for q, a in ((3, 0x8010), (4, 2), (5, 1)):
        k[(aux == 3) & small & (recs["address"][:, 0] == a)] = q
    return k


def test_reference_vector_is_satisfied(orc):
    recs = V.demux_reference_vector()
    assert len(recs) == 16
    io, _ = instance(orc, recs)
    rc, out, trace, com, st, tails = O.demux_entry_point(orc, io, recs, 16)
    assert rc == abi.ZKC_OK and st.failed_checks == 0
    assert out.completion_flag == 1
    # all 16 records are rollup storage accesses: queue 0 becomes the input queue, the rest stay empty
    assert [len(t) for t in tails] == [16, 0, 0, 0, 0, 0]
    assert list(out.output_queue_states[0].tail) == list(io.initial_log_queue_state.tail)
    assert out.output_queue_states[0].length == 16 and all(out.output_queue_states[q].length == 0 for q in range(1, 6))
    assert trace[K["BITMASK"]].tolist() == [1] * 16 and trace[K["IS_BITMASK"]].tolist() == [1] * 16


def test_six_way_split_matches_queue_model_and_chains(orc):
    recs = synthetic.vm_log_queue_trace(600, seed=3)
    kinds = kind_of(recs)
    assert all((kinds == q).sum() > 5 for q in range(6))
    io, _ = instance(orc, recs)
    rc, out, trace, com, st, tails = O.demux_entry_point(orc, io, recs, 640)
    assert rc == abi.ZKC_OK, (hex(st.failed_checks), st.first_bad_row)
    assert out.completion_flag == 1
    for q in range(6):
        sub = recs[kinds == q]
        prev, fin = O.log_queue_simulate(orc, sub)
        assert out.output_queue_states[q].length == len(sub) == len(tails[q])
        assert list(out.output_queue_states[q].tail) == list(fin.tail)
        assert bytes(out.output_queue_states[q]) == bytes(out.hidden_fsm_output.output_queue_states[q])
        # tails after each push = the previous-tail column the downstream circuit consumes, shifted by one
        assert np.array_equal(tails[q][:-1], prev[1:])
        assert trace[K["BITMASK"] + q].sum() == len(sub)
    # trivial rows: nothing executes, the discarded push runs on queue 0's state
    assert trace[K["EXECUTE"]][600:].sum() == 0 and trace[K["BITMASK"]:K["BITMASK"] + 6, 600:].sum() == 0
    assert trace[K["EXEC_LEN"]][639] == out.output_queue_states[0].length
    # chained instances == whole
    rc, a, ta, _, st, t1 = O.demux_entry_point(orc, io, recs, 250)
    assert rc == abi.ZKC_OK and a.completion_flag == 0 and all(a.output_queue_states[q].length == 0 for q in range(6))
    nxt = abi.DemuxClosedForm.from_buffer_copy(bytes(a)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.hidden_fsm_output
    rc, b, tb, _, st, t2 = O.demux_entry_point(orc, nxt, recs[250:], 390)
    assert rc == abi.ZKC_OK
    assert bytes(b.hidden_fsm_output) == bytes(out.hidden_fsm_output)
    assert np.array_equal(np.concatenate([ta, tb], axis=1), trace)
    for q in range(6):
        assert np.array_equal(np.concatenate([t1[q], t2[q]]), tails[q])
    exp = abi.DemuxClosedForm.from_buffer_copy(bytes(b)); exp.start_flag = 0; exp.hidden_fsm_input = a.hidden_fsm_output
    rc, *_ = O.demux_entry_point(orc, exp, recs[250:], 390, compare_expected=True)
    assert rc == abi.ZKC_OK
    exp.output_queue_states[4].length += 1
    rc, *_ = O.demux_entry_point(orc, exp, recs[250:], 390, compare_expected=True)
    assert rc == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH


def test_edge_and_negative_cases(orc):
    e = np.zeros(0, dtype=abi.LOG_QUERY_DTYPE)
    io, _ = instance(orc, e)
    rc, out, trace, com, st, tails = O.demux_entry_point(orc, io, e, 4)
    assert rc == abi.ZKC_OK and out.completion_flag == 1 and all(len(t) == 0 for t in tails)
    recs = synthetic.vm_log_queue_trace(100, seed=9)
    # a porter-shard storage access
    r2 = recs.copy(); i = int(np.flatnonzero(kind_of(recs) == 0)[5]); r2["flags"][i] |= 1 << 8
    io, _ = instance(orc, r2)
    rc, _, _, _, st, _ = O.demux_entry_point(orc, io, r2, 128)
    assert st.failed_checks == CHK["PORTER_STORAGE"] and st.first_bad_row == i
    # an unknown aux byte
    r3 = recs.copy(); r3["flags"][17] = (int(r3["flags"][17]) & 0xFFFFFF00) | 7
    io, _ = instance(orc, r3)
    rc, _, _, _, st, tails = O.demux_entry_point(orc, io, r3, 128)
    assert st.failed_checks == CHK["BITMASK"] and st.first_bad_row == 17
    # a precompile call to an address that is none of the three: legal for the circuit, goes nowhere
    r4 = recs.copy(); j = int(np.flatnonzero(kind_of(recs) == 3)[0]); r4["address"][j, 0] = 0x8011
    io, _ = instance(orc, r4)
    rc, out, trace, _, st, tails = O.demux_entry_point(orc, io, r4, 128)
    assert rc == abi.ZKC_OK and sum(len(t) for t in tails) == 99 and trace[K["BITMASK"]:K["BITMASK"] + 6, j].sum() == 0
    # custom constants
    opts = abi.DemuxOptions(); opts.custom_constants = 1
    opts.aux_bytes[:] = [0, 1, 2, 3]; opts.precompile_addresses[:] = [0x8011, 2, 1]
    rc, out, trace, _, st, tails = O.demux_entry_point(orc, io, r4, 128, options=opts)
    assert rc == abi.ZKC_OK and trace[K["BITMASK"] + 3, j] == 1
    # non-trivial head
    io, _ = instance(orc, recs); io.initial_log_queue_state.head[0] = 9
    rc, _, _, _, st, _ = O.demux_entry_point(orc, io, recs, 128)
    assert st.failed_checks & CHK["TRIVIAL_HEAD"]


def test_row_relations_of_the_trace(orc):
    """The row-to-row relations zkc_demux_log_queue_check_trace evaluates on the device (dmx_check_kernel), restated in numpy and held
    against the oracle's trace: classification, the six execute bits, the state push_with_optimize selects, the output queues'
    tails / lengths.  Pins the evaluator's reading of mod.rs:268-447 without a GPU."""
    n, limit = 2000, 2100
    recs = synthetic.vm_log_queue_trace(n, seed=5)
    io, _ = instance(orc, recs)
    rc, out, T, _, st, _ = O.demux_entry_point(orc, io, recs, limit)
    assert rc == 0
    K = abi.DMX_COLS
    col = lambda name, i=0: T[K[name] + i]
    prev = lambda a, first: np.concatenate([np.array([first], dtype=np.uint64), a[:-1].astype(np.uint64)])
    ex, emp, ln = col("EXECUTE"), col("QUEUE_IS_EMPTY"), col("LEN")
    len0 = io.initial_log_queue_state.length
    assert np.array_equal(ln + ex, prev(ln, len0)) and np.array_equal(emp, prev(ln, len0) == 0) and np.array_equal(ex, 1 - emp)
    I = K["ITEM"]
    aux, shard, small = T[I + 29], T[I + 33], (T[I + 1] | T[I + 2] | T[I + 3] | T[I + 4]) == 0
    ia = [col("IS_AUX", i) for i in range(4)]
    iad = [col("IS_ADDRESS", i) for i in range(3)]
    assert all(np.array_equal(ia[i], aux == i) for i in range(4))
    assert all(np.array_equal(iad[i], small & (T[I] == a)) for i, a in enumerate((0x8010, 2, 1)))
    rollup = col("IS_ROLLUP_SHARD")
    assert np.array_equal(rollup, shard == 0) and np.array_equal(col("EXECUTE_PORTER_STORAGE"), ia[0] & (1 - rollup) & ex)
    bits = [ia[0] & rollup & ex, ia[1] & ex, ia[2] & ex, ia[3] & iad[0] & ex, ia[3] & iad[1] & ex, ia[3] & iad[2] & ex]
    assert all(np.array_equal(col("BITMASK", q), bits[q]) for q in range(6))
    assert np.array_equal(col("IS_BITMASK"), (ia[0] + ia[1] + ia[2] + ia[3]) == 1)
    sel, anyb = np.zeros(limit, int), np.zeros(limit, bool)
    for q in range(6):
        sel = np.where(bits[q] == 1, q, sel); anyb |= bits[q] == 1
    ql = [col("QUEUE_LENS", q) for q in range(6)]
    assert all(np.array_equal(ql[q], prev(ql[q], 0) + bits[q]) for q in range(6))
    rows = np.arange(limit)
    assert np.array_equal(col("EXEC_LEN"), np.stack([prev(ql[q], 0) for q in range(6)])[sel, rows])
    for i in range(4):
        pt = np.stack([prev(col("QUEUE_TAILS", 4 * q + i), 0) for q in range(6)])
        assert np.array_equal(col("EXEC_TAIL", i), pt[sel, rows])
        for q in range(6):
            assert np.array_equal(col("QUEUE_TAILS", 4 * q + i), np.where(anyb & (sel == q), col("PUSH_ROUND2", i), pt[q]))
        h = col("HEAD", i)
        assert np.array_equal(h[ex == 0], prev(h, 0)[ex == 0])
