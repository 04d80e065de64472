"""keccak256_round_function: CUDA path through the C ABI vs the CPU oracle, bit-exact, plus the reference's own
digest check (circuit output == Keccak256, /root/reference/src/keccak256_round_function/mod.rs:1096-1144)."""
import numpy as np
import pytest

import orc as O
from era_zkevm_circuits_b200 import (Keccak256RoundFunctionCircuitInstanceWitness, abi, keccak256_round_function_entry_point,
                                     synthetic)
from test_oracle_keccak import REFERENCE_CASES, keccak256, single_call_instance

pytestmark = pytest.mark.gpu
K = abi.KC_COLS
W = Keccak256RoundFunctionCircuitInstanceWitness


def assert_same(want, got, check_trace=True):
    rc, io, trace, com, st, states = want
    assert got.status.code == rc, (got.status.code, hex(got.status.failed_checks), got.status.first_bad_row)
    assert got.status.failed_checks == st.failed_checks and got.status.first_bad_row == st.first_bad_row
    assert got.closed_form_input.completion_flag == io.completion_flag
    assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(io.hidden_fsm_output)
    assert bytes(got.closed_form_input.final_memory_state) == bytes(io.final_memory_state)
    assert got.commitment.tolist() == com.tolist()
    if check_trace:
        bad = np.argwhere(got.trace != trace)
        assert bad.size == 0, f"first differing (col,row): {bad[:8].tolist()}"


def run_both(engine, orc, io, req, reads, limit, states=None, **kw):
    want = O.keccak_entry_point(orc, io, req, reads, limit)
    prev, _ = O.log_queue_simulate(orc, req)
    got = keccak256_round_function_entry_point(engine, W(io, req, prev, reads, states), limit, raise_on_unsatisfied=False, **kw)
    return want, got


@pytest.mark.parametrize("length,unalignment", REFERENCE_CASES + [(0, 0), (0, 7), (1, 31), (272, 0), (1023, 17)])
def test_reference_cases(engine, orc, length, unalignment):
    msg = np.random.default_rng(length * 32 + unalignment).integers(0, 256, length, dtype=np.uint8).tobytes()
    io, req, reads = single_call_instance(orc, msg, unalignment)
    limit = max(2, length // 136 + 2)
    want, got = run_both(engine, orc, io, req, reads, limit)
    assert want[0] == abi.ZKC_OK
    assert_same(want, got)
    row = int(np.flatnonzero(got.trace[K["WRITE_RESULT"]])[0])
    limbs = got.trace[K["RESULT"]:K["RESULT"] + 8, row].astype(np.uint32)
    assert int.from_bytes(limbs.astype("<u4").tobytes(), "little").to_bytes(32, "big") == keccak256(orc, msg)
    want2, got2 = run_both(engine, orc, io, req, reads, limit, states=want[5])
    assert_same(want2, got2)


@pytest.mark.parametrize("n_calls,max_len,extra", [(1, 300, 0), (40, 700, 5), (300, 1024, 0), (2000, 600, 100)])
def test_many_calls_bit_exact(engine, orc, n_calls, max_len, extra):
    reqs, reads, msgs = synthetic.keccak_calls(n_calls, seed=n_calls, max_len=max_len)
    _, rfin = O.log_queue_simulate(orc, reqs)
    io = O.keccak_closed_form(rfin)
    limit = sum(len(m) // 136 + 1 for m in msgs) + extra
    want, got = run_both(engine, orc, io, reqs, reads, limit)
    assert want[0] == abi.ZKC_OK, hex(want[4].failed_checks)
    assert_same(want, got)
    rows = np.flatnonzero(got.trace[K["WRITE_RESULT"]])
    assert len(rows) == n_calls
    for r, m in list(zip(rows, msgs))[:50]:
        limbs = got.trace[K["RESULT"]:K["RESULT"] + 8, r].astype(np.uint32)
        assert int.from_bytes(limbs.astype("<u4").tobytes(), "little").to_bytes(32, "big") == keccak256(orc, m)
    want2, got2 = run_both(engine, orc, io, reqs, reads, limit, states=want[5])
    assert_same(want2, got2)


def test_chained_instances_split_mid_call(engine, orc):
    reqs, reads, msgs = synthetic.keccak_calls(60, seed=5, max_len=900)
    _, rfin = O.log_queue_simulate(orc, reqs)
    io = O.keccak_closed_form(rfin)
    limit = sum(len(m) // 136 + 1 for m in msgs) + 3
    whole = O.keccak_entry_point(orc, io, reqs, reads, limit)
    rows = np.flatnonzero(whole[2][K["WRITE_RESULT"]])
    for cut in (int(rows[20]) - 1, int(rows[20]) + 1, int(rows[33])):
        prev, _ = O.log_queue_simulate(orc, reqs)
        a = keccak256_round_function_entry_point(engine, W(io, reqs, prev, reads), cut)
        want_a = O.keccak_entry_point(orc, io, reqs, reads, cut)
        assert_same(want_a, a)
        nxt = abi.KeccakClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
        nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
        used_req = len(reqs) - a.closed_form_input.hidden_fsm_output.log_queue_state.length
        used_reads = int(a.trace[K["QUERY"] + 3::K["QUERY_STRIDE"]][:6].sum())
        want_b = O.keccak_entry_point(orc, nxt, reqs[used_req:], reads[used_reads:], limit - cut)
        b = keccak256_round_function_entry_point(engine, W(nxt, reqs[used_req:], prev[used_req:], reads[used_reads:]), limit - cut)
        assert_same(want_b, b)
        assert bytes(b.closed_form_input.hidden_fsm_output) == bytes(whole[1].hidden_fsm_output)
        assert np.array_equal(np.concatenate([a.trace, b.trace], axis=1), whole[2])


def test_negative_and_empty(engine, orc):
    msg = b"hello world" * 20
    io, req, reads = single_call_instance(orc, msg, 3)
    bad = req.copy(); bad["address"][0][0] = 0x8011
    _, rfin = O.log_queue_simulate(orc, bad)
    want, got = run_both(engine, orc, O.keccak_closed_form(rfin), bad, reads, 4)
    assert want[4].failed_checks == abi.KC_CHK["ADDRESS"]
    assert_same(want, got)
    bad = req.copy(); bad["flags"][0] = abi.lq_flags(aux=1, rw=1)
    _, rfin = O.log_queue_simulate(orc, bad)
    want, got = run_both(engine, orc, O.keccak_closed_form(rfin), bad, reads, 4)
    assert_same(want, got)
    # empty request queue: can_finish_immediately (mod.rs:196-213)
    e = np.zeros(0, dtype=abi.LOG_QUERY_DTYPE)
    _, rfin = O.log_queue_simulate(orc, e)
    want, got = run_both(engine, orc, O.keccak_closed_form(rfin), e, np.zeros((0, 8), dtype=np.uint32), 6)
    assert want[0] == 0 and want[1].completion_flag == 1
    assert_same(want, got)
    # corrupted memory-queue hint
    want = O.keccak_entry_point(orc, io, req, reads, 4)
    s = want[5].copy(); s[1, 3] ^= 1
    prev, _ = O.log_queue_simulate(orc, req)
    r = keccak256_round_function_entry_point(engine, W(io, req, prev, reads, s), 4, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT


def test_check_trace_constraint_evaluation(engine, orc):
    """zkc_keccak256_round_function_check_trace: the ORACLE's trace (many calls, every length / unalignment class, full padding rounds)
    satisfies every relation with and without the queue permutations; a fault injected into any cell family is found at its cycle;
    the engine's trace of a chained second instance (cut inside a message: buffer, fill count and keccak state carried) passes"""
    from era_zkevm_circuits_b200 import keccak256_round_function_check_trace
    V_ = abi.KCV
    reqs, reads, msgs = synthetic.keccak_calls(120, seed=8, max_len=700)
    prev, rfin = O.log_queue_simulate(orc, reqs)
    io = O.keccak_closed_form(rfin)
    limit = sum(len(m) // 136 + 1 for m in msgs) + 12
    want = O.keccak_entry_point(orc, io, reqs, reads, limit)
    assert want[0] == abi.ZKC_OK
    trace = want[2]
    for gates in (0, abi.GATES_GENERAL):
        viol, st = keccak256_round_function_check_trace(engine, io, trace, limit, gates)
        assert viol == 0 and st.code == 0, (gates, viol, hex(st.failed_checks), st.first_bad_row)
    import torch
    viol, st = keccak256_round_function_check_trace(engine, io, torch.from_numpy(trace.view(np.int64)).cuda(), limit, abi.GATES_GENERAL)
    assert viol == 0
    Q, QS = K["QUERY"], K["QUERY_STRIDE"]
    pops = np.flatnonzero(trace[K["FLAGS_IN"]])
    writes = np.flatnonzero(trace[K["WRITE_RESULT"]])
    mid = int(np.flatnonzero((trace[K["FLAGS_IN"]] == 0) & (trace[Q + 3] == 1) & (trace[K["WRITE_RESULT"]] == 0))[30])  # reading, mid-message
    faults = [
        (K["FLAGS_IN"] + 1, mid, None, V_["FSM"], 0),
        (K["CALL_ITEM"] + 9, int(pops[5]), 1 << 33, V_["BOOLEAN"], 0),
        (K["CALL_ITEM"] + 3, mid, None, V_["BOOLEAN"], 0),
        (K["CALL_ITEM"], int(pops[6]), None, V_["ENFORCE"], 0),
        (K["REQ_LEN"], mid, None, V_["QUEUE"], 0),
        (K["REQ_HEAD"] + 2, mid, None, V_["QUEUE"], 0),
        (K["REQ_HEAD"] + 2, int(pops[7]), None, V_["ROUND_FUNCTION"], 0),
        (K["PARAMS"] + 4, mid, None, V_["PARAMS"], 0),
        (K["TS_WRITE"], mid, None, V_["PARAMS"], 0),
        (K["READ_NON_ZERO_LENGTH"], mid, None, V_["FSM"], 0),
        (Q + 2 * QS + 2, mid, None, V_["QUERIES"], 0),
        (Q + 4 + 1, mid, None, V_["SPONGE"] | V_["BUFFER"], abi.GATES_GENERAL),  # a word read from memory is a free input: what it feeds no longer matches
        (Q + 3 * QS + 24, mid, None, V_["MEMORY_QUEUE"], 0),
        (Q + 12 + 3, mid, None, V_["ROUND_FUNCTION"], 0),
        (Q + 12 + 3, limit - 3, None, V_["MEMORY_QUEUE"], abi.GATES_GENERAL),
        (K["CURRENTLY_FILLED"], mid, None, V_["FSM"], 0),
        (K["INPUT"] + 77, mid, None, V_["SPONGE"], 0),
        (K["STATE_OUT"] + 123, mid, None, V_["SPONGE"], 0),
        (K["RESULT"] + 2, int(writes[9]), None, V_["SPONGE"], abi.GATES_GENERAL),
        (K["WRITE_RESULT"], mid, None, V_["FSM"], 0),
        (K["WRITE_TAIL"] + 5, int(writes[10]), None, V_["ROUND_FUNCTION"], 0),
        (K["FLAGS_OUT"] + 3, mid, None, V_["FSM"], 0),
        (K["BUFFER_OUT"] + 20, mid, None, V_["BUFFER"], 0),
    ]
    for col, row, val, bit, gates in faults:
        bad = trace.copy()
        bad[col, row] = np.uint64(val) if val is not None else bad[col, row] ^ np.uint64(1)
        viol, st = keccak256_round_function_check_trace(engine, io, bad, limit, gates)
        assert viol >= 1 and st.first_bad_row == row and st.failed_checks & bit, (col, row, viol, st.first_bad_row, hex(st.failed_checks))
    cut = mid
    a = keccak256_round_function_entry_point(engine, W(io, reqs, prev, reads), cut)
    nxt = abi.KeccakClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    used_req = len(reqs) - a.closed_form_input.hidden_fsm_output.log_queue_state.length
    used_reads = int(a.trace[K["QUERY"] + 3::K["QUERY_STRIDE"]][:6].sum())
    b = keccak256_round_function_entry_point(engine, W(nxt, reqs[used_req:], prev[used_req:], reads[used_reads:]), limit - cut)
    assert b.status.code == 0
    viol, st = keccak256_round_function_check_trace(engine, nxt, b.trace, limit - cut)
    assert viol == 0, (viol, hex(st.failed_checks), st.first_bad_row)
