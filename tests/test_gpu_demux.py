"""demux_log_queue: CUDA path through the C ABI vs the CPU oracle, bit-exact (trace, FSM output, observable output,
commitment, status).  Mirrors /root/reference/src/demux_log_queue/mod.rs:482-600 and widens it."""
import numpy as np
import pytest

import orc as O
import vectors as V
from era_zkevm_circuits_b200 import (LogDemuxerCircuitInstanceWitness as Witness, abi,
                                     demultiplex_storage_logs_enty_point as entry_point, synthetic)

pytestmark = pytest.mark.gpu
K = abi.DMX_COLS
CHK = abi.DMX_CHK


def instance(orc, recs):
    prev, fin = O.log_queue_simulate(orc, recs)
    return O.demux_closed_form(fin, True), prev


def assert_same(want, got, check_trace=True):
    rc, io, trace, com, st, tails = want
    assert got.status.code == rc, (got.status.code, hex(got.status.failed_checks), got.status.first_bad_row, rc, hex(st.failed_checks))
    assert got.status.failed_checks == st.failed_checks and got.status.first_bad_row == st.first_bad_row
    assert got.closed_form_input.completion_flag == io.completion_flag
    assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(io.hidden_fsm_output)
    assert bytes(got.closed_form_input.output_queue_states) == bytes(io.output_queue_states)
    assert got.commitment.tolist() == com.tolist()
    if check_trace:
        bad = np.argwhere(got.trace != trace)
        assert bad.size == 0, f"first differing (col,row): {bad[:5].tolist()}"


def hints(tails):
    return np.concatenate(tails) if sum(len(t) for t in tails) else np.zeros((0, 4), dtype=np.uint64), [len(t) for t in tails]


def run_both(engine, orc, io, recs, prev, limit, tails=None, **kw):
    want = O.demux_entry_point(orc, io, recs, limit, options=kw.get("options"))
    t, c = hints(tails) if tails is not None else (None, None)
    got = entry_point(engine, Witness(io, recs, prev, t, c), limit, raise_on_unsatisfied=False, **kw)
    return want, got


def test_reference_vector(engine, orc):
    recs = V.demux_reference_vector()
    io, prev = instance(orc, recs)
    want, got = run_both(engine, orc, io, recs, prev, 16)
    assert want[0] == abi.ZKC_OK
    assert_same(want, got)
    want, got = run_both(engine, orc, io, recs, prev, 16, tails=want[5])
    assert_same(want, got)


@pytest.mark.parametrize("n,limit", [(1, 1), (2, 5), (255, 256), (257, 257), (1000, 1024), (20000, 20000), (3000, 70000)])
def test_synthetic_bit_exact(engine, orc, n, limit):
    recs = synthetic.vm_log_queue_trace(n, seed=n)
    io, prev = instance(orc, recs)
    want, got = run_both(engine, orc, io, recs, prev, limit)
    assert want[0] == abi.ZKC_OK, (hex(want[4].failed_checks), want[4].first_bad_row)
    assert sum(want[1].output_queue_states[q].length for q in range(6)) == n
    assert_same(want, got)
    want2, got2 = run_both(engine, orc, io, recs, prev, limit, tails=want[5])
    assert_same(want2, got2)


def test_chained_instances_empty_and_options(engine, orc):
    recs = synthetic.vm_log_queue_trace(3000, seed=9)
    io, prev = instance(orc, recs)
    whole = entry_point(engine, Witness(io, recs, prev), 3000)
    a = entry_point(engine, Witness(io, recs, prev), 1100)
    assert a.closed_form_input.completion_flag == 0
    nxt = abi.DemuxClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    want, got = run_both(engine, orc, nxt, recs[1100:], prev[1100:], 1900)
    assert_same(want, got)
    assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(whole.closed_form_input.hidden_fsm_output)
    assert np.array_equal(np.concatenate([a.trace, got.trace], axis=1), whole.trace)
    want2, got2 = run_both(engine, orc, nxt, recs[1100:], prev[1100:], 1900, tails=want[5])
    assert_same(want2, got2)
    exp = abi.DemuxClosedForm.from_buffer_copy(bytes(nxt))
    exp.hidden_fsm_output = got.closed_form_input.hidden_fsm_output
    exp.output_queue_states = got.closed_form_input.output_queue_states
    exp.completion_flag = 1
    ok = entry_point(engine, Witness(exp, recs[1100:], prev[1100:]), 1900, compare_expected=True)
    assert ok.status.code == 0
    exp.output_queue_states[5].length += 1
    bad = entry_point(engine, Witness(exp, recs[1100:], prev[1100:]), 1900, compare_expected=True, raise_on_unsatisfied=False)
    assert bad.status.code == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH
    e = np.zeros(0, dtype=abi.LOG_QUERY_DTYPE)
    io0, prev0 = instance(orc, e)
    want, got = run_both(engine, orc, io0, e, prev0, 8)
    assert want[0] == abi.ZKC_OK
    assert_same(want, got)
    want, got = run_both(engine, orc, io, recs, prev, 0)
    assert_same(want, got)
    # custom constants: another keccak address, swapped aux bytes
    opts = abi.DemuxOptions(); opts.custom_constants = 1
    opts.aux_bytes[:] = [1, 0, 2, 3]; opts.precompile_addresses[:] = [0x8011, 2, 1]
    want, got = run_both(engine, orc, io, recs, prev, 3000, options=opts)
    assert want[1].output_queue_states[3].length == 0 and want[1].output_queue_states[1].length > 1000
    assert_same(want, got)
    opts.aux_bytes[:] = [1, 1, 2, 3]
    with pytest.raises(Exception, match="INVALID_ARGUMENT"):
        entry_point(engine, Witness(io, recs, prev), 3000, options=opts)


def test_negative_cases_match_oracle(engine, orc):
    recs = synthetic.vm_log_queue_trace(1500, seed=4)
    r2 = recs.copy(); r2["flags"][int(np.flatnonzero((recs["flags"] & 0xFF) == 0)[9])] |= 1 << 8
    r3 = recs.copy(); r3["flags"][700] = (int(r3["flags"][700]) & 0xFFFFFF00) | 9
    r4 = recs.copy(); r4["address"][int(np.flatnonzero((recs["flags"] & 0xFF) == 3)[4]), 3] = 1
    for rr, code in ((r2, abi.ZKC_ERR_UNSATISFIED), (r3, abi.ZKC_ERR_UNSATISFIED), (r4, abi.ZKC_OK)):
        io, prev = instance(orc, rr)
        want, got = run_both(engine, orc, io, rr, prev, 1536)
        assert want[0] == code
        assert_same(want, got)
        want, got = run_both(engine, orc, io, rr, prev, 1536, tails=want[5])
        assert_same(want, got)
    # corrupted hints
    io, prev = instance(orc, recs)
    want = O.demux_entry_point(orc, io, recs, 1536)
    t, c = hints(want[5])
    t2 = t.copy(); t2[c[0] + 3, 1] ^= 1  # 4th tail of the events queue
    r = entry_point(engine, Witness(io, recs, prev, t2, c), 1536, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT
    c2 = list(c); c2[2] -= 1
    r = entry_point(engine, Witness(io, recs, prev, np.delete(t, c[0] + c[1] + c[2] - 1, axis=0), c2), 1536, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT
    p2 = prev.copy(); p2[7, 0] ^= 1
    r = entry_point(engine, Witness(io, recs, p2), 1536, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT and r.status.first_bad_row in (6, 7)


def test_device_resident_feeds_the_sorters(engine, orc):
    """2^17 VM log records on the device; the events queue the demultiplexer leaves is the queue of exactly the event
    records (the queue simulation is the sorter-side witness builder)"""
    import torch
    n = 1 << 17
    recs = synthetic.vm_log_queue_trace(n, seed=21)
    tod = lambda a: torch.from_numpy(a.view(np.uint8).reshape(len(a), -1)).cuda()
    d = tod(recs)
    prev, fin = engine.log_queue_simulate(d)
    io = O.demux_closed_form(fin[0], True)
    got = entry_point(engine, Witness(io, d, prev), n, want_trace=False)
    out = got.closed_form_input
    assert got.status.code == 0 and out.completion_flag == 1
    assert sum(out.output_queue_states[q].length for q in range(6)) == n
    aux = recs["flags"] & 0xFF
    for q, sel in ((1, aux == 1), (2, aux == 2), (0, aux == 0)):
        _, f = engine.log_queue_simulate(tod(np.ascontiguousarray(recs[sel])))
        assert bytes(f[0]) == bytes(out.output_queue_states[q]), q


def test_check_trace_constraint_evaluation(engine, orc):
    """zkc_demux_log_queue_check_trace: the ORACLE's trace satisfies every relation with and without the round-function gates; a fault
    injected into any relation family is found at its row; the engine's own trace of a chained second instance passes"""
    from era_zkevm_circuits_b200 import demux_log_queue_check_trace
    V_ = abi.DMXV
    n, limit = 3000, 3100
    recs = synthetic.vm_log_queue_trace(n, seed=8)
    io, prev = instance(orc, recs)
    want = O.demux_entry_point(orc, io, recs, limit)
    assert want[0] == abi.ZKC_OK
    trace = want[2]
    for gates in (0, abi.GATES_GENERAL):
        viol, st = demux_log_queue_check_trace(engine, io, trace, limit, gates)
        assert viol == 0 and st.code == 0, (gates, viol, hex(st.failed_checks), st.first_bad_row)
    import torch
    viol, st = demux_log_queue_check_trace(engine, io, torch.from_numpy(trace.view(np.int64)).cuda(), limit, abi.GATES_GENERAL)
    assert viol == 0
    events = np.flatnonzero(trace[K["BITMASK"] + 1])
    faults = [
        (K["EXECUTE"], 17, 2, V_["BOOLEAN"], 0),
        (K["ITEM"] + 7, 40, 1 << 33, V_["BOOLEAN"], 0),
        (K["ENC"] + 3, 99, None, V_["ENCODING"], 0),
        (K["LEN"], 123, None, V_["QUEUE_LEN"], 0),
        (K["HEAD"] + 1, 3050, None, V_["QUEUE_LEN"], 0),
        (K["HEAD"] + 1, 150, None, V_["ROUND_FUNCTION"], 0),
        (K["IS_AUX"] + 2, 200, None, V_["FLAGS"], 0),
        (K["IS_ADDRESS"] + 1, 210, None, V_["FLAGS"], 0),
        (K["IS_ROLLUP_SHARD"], 220, None, V_["FLAGS"], 0),
        (K["BITMASK"] + 4, 230, None, V_["FLAGS"], 0),
        (K["IS_BITMASK"], 240, None, V_["FLAGS"], 0),
        (K["EXEC_TAIL"] + 2, 300, None, V_["OUTPUT_QUEUES"], 0),
        (K["EXEC_LEN"], 310, None, V_["OUTPUT_QUEUES"], 0),
        (K["QUEUE_LENS"] + 3, 320, None, V_["OUTPUT_QUEUES"], 0),
        (K["QUEUE_TAILS"] + 4 * 1 + 2, int(events[20]), None, V_["OUTPUT_QUEUES"], abi.GATES_GENERAL),
        (K["PUSH_ROUND0"] + 5, 400, None, V_["ROUND_FUNCTION"], 0),
        (K["PUSH_ROUND2"] + 9, 410, None, V_["ROUND_FUNCTION"], 0),
    ]
    for col, row, val, bit, gates in faults:
        bad = trace.copy()
        bad[col, row] = np.uint64(val) if val is not None else bad[col, row] ^ np.uint64(1)
        viol, st = demux_log_queue_check_trace(engine, io, bad, limit, gates)
        assert viol >= 1 and st.first_bad_row == row and st.failed_checks & bit, (col, row, viol, st.first_bad_row, hex(st.failed_checks))
    cut = 1500
    a = entry_point(engine, Witness(io, recs, prev, None, None), cut, raise_on_unsatisfied=False)
    nxt = abi.DemuxClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    b = entry_point(engine, Witness(nxt, recs[cut:], prev[cut:], None, None), limit - cut, raise_on_unsatisfied=False)
    assert b.status.code == 0
    viol, st = demux_log_queue_check_trace(engine, nxt, b.trace, limit - cut)
    assert viol == 0, (viol, hex(st.failed_checks), st.first_bad_row)
