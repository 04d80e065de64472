"""sha256_round_function: CUDA path through the C ABI vs the CPU oracle, bit-exact, and the digests against hashlib."""
import hashlib

import numpy as np
import pytest

import orc as O
from era_zkevm_circuits_b200 import Sha256RoundFunctionCircuitInstanceWitness, abi, sha256_round_function_entry_point, synthetic
from test_oracle_sha256 import digest_of_row

pytestmark = pytest.mark.gpu
K = abi.SH_COLS
W = Sha256RoundFunctionCircuitInstanceWitness


def assert_same(want, got):
    rc, io, trace, com, st, states = want
    assert got.status.code == rc, (got.status.code, hex(got.status.failed_checks), got.status.first_bad_row)
    assert got.status.failed_checks == st.failed_checks and got.status.first_bad_row == st.first_bad_row
    assert got.closed_form_input.completion_flag == io.completion_flag
    assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(io.hidden_fsm_output)
    assert bytes(got.closed_form_input.final_memory_state) == bytes(io.final_memory_state)
    assert got.commitment.tolist() == com.tolist()
    bad = np.argwhere(got.trace != trace)
    assert bad.size == 0, f"first differing (col,row): {bad[:8].tolist()}"


@pytest.mark.parametrize("n_calls,max_rounds,extra", [(1, 1, 0), (1, 7, 2), (50, 16, 5), (3000, 16, 64)])
def test_bit_exact_and_digests(engine, orc, n_calls, max_rounds, extra):
    reqs, reads, msgs = synthetic.sha256_calls(n_calls, seed=n_calls + max_rounds, max_rounds=max_rounds)
    prev, rfin = O.log_queue_simulate(orc, reqs)
    io = O.sha256_closed_form(rfin)
    limit = len(reads) // 2 + extra
    want = O.sha256_entry_point(orc, io, reqs, reads, limit)
    assert want[0] == 0
    got = sha256_round_function_entry_point(engine, W(io, reqs, prev, reads), limit)
    assert_same(want, got)
    rows = np.flatnonzero(got.trace[K["WRITE_RESULT"]])
    assert len(rows) == n_calls
    for r, m in list(zip(rows, msgs))[:64]:
        assert digest_of_row(got.trace, r) == hashlib.sha256(m).digest()
    got2 = sha256_round_function_entry_point(engine, W(io, reqs, prev, reads, want[5]), limit)
    assert_same(want, got2)


def test_chained_and_negative(engine, orc):
    reqs, reads, msgs = synthetic.sha256_calls(40, seed=9, max_rounds=12)
    prev, rfin = O.log_queue_simulate(orc, reqs)
    io = O.sha256_closed_form(rfin)
    total = len(reads) // 2
    whole = O.sha256_entry_point(orc, io, reqs, reads, total + 2)
    for cut in (1, total // 3, total // 2 + 1, total - 1):
        a = sha256_round_function_entry_point(engine, W(io, reqs, prev, reads), cut)
        assert_same(O.sha256_entry_point(orc, io, reqs, reads, cut), a)
        nxt = abi.Sha256ClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
        nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
        used = len(reqs) - a.closed_form_input.hidden_fsm_output.log_queue_state.length
        want_b = O.sha256_entry_point(orc, nxt, reqs[used:], reads[2 * cut:], total + 2 - cut)
        b = sha256_round_function_entry_point(engine, W(nxt, reqs[used:], prev[used:], reads[2 * cut:]), total + 2 - cut)
        assert_same(want_b, b)
        assert bytes(b.closed_form_input.hidden_fsm_output) == bytes(whole[1].hidden_fsm_output)
        assert np.array_equal(np.concatenate([a.trace, b.trace], axis=1), whole[2])
    bad = reqs.copy(); bad["address"][5][0] = 0x8010
    prevb, rfinb = O.log_queue_simulate(orc, bad)
    want = O.sha256_entry_point(orc, O.sha256_closed_form(rfinb), bad, reads, total)
    got = sha256_round_function_entry_point(engine, W(O.sha256_closed_form(rfinb), bad, prevb, reads), total, raise_on_unsatisfied=False)
    assert want[4].failed_checks == abi.KC_CHK["ADDRESS"]
    assert_same(want, got)
    e = np.zeros(0, dtype=abi.LOG_QUERY_DTYPE)
    preve, rfine = O.log_queue_simulate(orc, e)
    want = O.sha256_entry_point(orc, O.sha256_closed_form(rfine), e, np.zeros((0, 8), dtype=np.uint32), 5)
    got = sha256_round_function_entry_point(engine, W(O.sha256_closed_form(rfine), e, preve, np.zeros((0, 8), dtype=np.uint32)), 5)
    assert_same(want, got)


def test_check_trace_constraint_evaluation(engine, orc):
    """zkc_sha256_round_function_check_trace: the ORACLE's trace satisfies every relation with and without the queue permutations; a
    fault injected into any relation family is found at its cycle; the engine's trace of a chained second instance (cut inside a
    message) passes"""
    from era_zkevm_circuits_b200 import sha256_round_function_check_trace
    V_ = abi.SHV
    reqs, reads, msgs = synthetic.sha256_calls(300, seed=8, max_rounds=9)
    prev, rfin = O.log_queue_simulate(orc, reqs)
    io = O.sha256_closed_form(rfin)
    total = len(reads) // 2
    limit = total + 30
    want = O.sha256_entry_point(orc, io, reqs, reads, limit)
    assert want[0] == abi.ZKC_OK
    trace = want[2]
    for gates in (0, abi.GATES_GENERAL):
        viol, st = sha256_round_function_check_trace(engine, io, trace, limit, gates)
        assert viol == 0 and st.code == 0, (gates, viol, hex(st.failed_checks), st.first_bad_row)
    import torch
    viol, st = sha256_round_function_check_trace(engine, io, torch.from_numpy(trace.view(np.int64)).cuda(), limit, abi.GATES_GENERAL)
    assert viol == 0
    pops = np.flatnonzero(trace[K["FLAGS_IN"]])
    writes = np.flatnonzero(trace[K["WRITE_RESULT"]])
    mid = int(np.flatnonzero((trace[K["FLAGS_IN"]] == 0) & (trace[K["SHOULD_READ"]] == 1))[40])  # a cycle in the middle of a message
    faults = [
        (K["FLAGS_IN"] + 1, mid, None, V_["FSM"], 0),
        (K["CALL_ITEM"] + 9, int(pops[5]), 1 << 33, V_["BOOLEAN"], 0),
        (K["CALL_ITEM"], int(pops[6]), None, V_["ENFORCE"], 0),
        (K["REQ_LEN"], mid, None, V_["QUEUE"], 0),
        (K["REQ_HEAD"] + 2, mid, None, V_["QUEUE"], 0),
        (K["REQ_HEAD"] + 2, int(pops[7]), None, V_["ROUND_FUNCTION"], 0),
        (K["PARAMS"] + 3, mid, None, V_["PARAMS"], 0),
        (K["PARAMS"] + 1, int(pops[8]), None, V_["PARAMS"], 0),
        (K["TS_WRITE"], mid, None, V_["PARAMS"], 0),
        (K["SHOULD_READ"], limit - 3, None, V_["FSM"], 0),
        (K["QUERY"] + 21, mid, None, V_["PARAMS"], 0),
        (K["QUERY"] + K["QUERY_STRIDE"] + 20, mid, None, V_["MEMORY_QUEUE"], 0),
        (K["QUERY"] + 8 + 3, mid, None, V_["ROUND_FUNCTION"], 0),
        (K["QUERY"] + 8 + 3, limit - 3, None, V_["MEMORY_QUEUE"], abi.GATES_GENERAL),
        (K["MESSAGE"] + 11, mid, None, V_["COMPRESSION"], 0),
        (K["NUM_ROUNDS"], mid, None, V_["PARAMS"], 0),
        (K["STATE_IN"] + 4, mid, None, V_["COMPRESSION"], 0),
        (K["STATE_OUT"] + 6, mid, None, V_["COMPRESSION"], 0),
        (K["RESULT"] + 2, mid, None, V_["COMPRESSION"], 0),
        (K["WRITE_RESULT"], mid, None, V_["FSM"], 0),
        (K["WRITE_TAIL"] + 5, int(writes[9]), None, V_["ROUND_FUNCTION"], 0),
        (K["WRITE_LEN"], mid, None, V_["MEMORY_QUEUE"], 0),
        (K["FLAGS_OUT"] + 2, mid, None, V_["FSM"], 0),
    ]
    for col, row, val, bit, gates in faults:
        bad = trace.copy()
        bad[col, row] = np.uint64(val) if val is not None else bad[col, row] ^ np.uint64(1)
        viol, st = sha256_round_function_check_trace(engine, io, bad, limit, gates)
        assert viol >= 1 and st.first_bad_row == row and st.failed_checks & bit, (col, row, viol, st.first_bad_row, hex(st.failed_checks))
    cut = mid
    a = sha256_round_function_entry_point(engine, W(io, reqs, prev, reads), cut)
    nxt = abi.Sha256ClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    used = len(reqs) - a.closed_form_input.hidden_fsm_output.log_queue_state.length
    b = sha256_round_function_entry_point(engine, W(nxt, reqs[used:], prev[used:], reads[2 * cut:]), limit - cut)
    assert b.status.code == 0
    viol, st = sha256_round_function_check_trace(engine, nxt, b.trace, limit - cut)
    assert viol == 0, (viol, hex(st.failed_checks), st.first_bad_row)
