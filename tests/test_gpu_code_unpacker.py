"""code_unpacker_sha256: CUDA path (one thread per request chaining its SHA-256 rounds) through the C ABI vs the CPU oracle,
bit-exact (trace, FSM output, observable output, commitment, status)."""
import numpy as np
import pytest

import orc as O
import vectors as V
from era_zkevm_circuits_b200 import (CodeDecommitterCircuitInstanceWitness as Witness, abi, synthetic,
                                     unpack_code_into_memory_entry_point as entry_point)

pytestmark = pytest.mark.gpu
K = abi.CU_COLS
CHK = abi.CU_CHK


def instance(orc, reqs):
    prev, fin = O.decommit_queue_simulate(orc, reqs)
    return O.code_unpacker_closed_form(fin, None, True), prev


def rounds_of(reqs):
    return (((reqs["code_hash"][:, 7] & 0xFFFF).astype(np.int64) + 1) // 2)


def assert_same(want, got, check_trace=True):
    rc, io, trace, com, st, states = want
    assert got.status.code == rc, (got.status.code, hex(got.status.failed_checks), got.status.first_bad_row, rc, hex(st.failed_checks), st.first_bad_row)
    assert got.status.failed_checks == st.failed_checks and got.status.first_bad_row == st.first_bad_row
    assert got.closed_form_input.completion_flag == io.completion_flag
    assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(io.hidden_fsm_output)
    assert bytes(got.closed_form_input.memory_queue_final_state) == bytes(io.memory_queue_final_state)
    assert got.commitment.tolist() == com.tolist()
    if check_trace:
        bad = np.argwhere(got.trace != trace)
        assert bad.size == 0, f"first differing (col,row): {bad[:5].tolist()}"


def run_both(engine, orc, io, reqs, prev, words, limit, states=None, **kw):
    want = O.code_unpacker_entry_point(orc, io, reqs, words, limit)
    got = entry_point(engine, Witness(io, reqs, prev, words, states), limit, raise_on_unsatisfied=False, **kw)
    return want, got


def test_reference_vector(engine, orc):
    """/root/reference/src/code_unpacker_sha256/mod.rs:472-700: 1 request, 33 words, limit 40"""
    reqs, words = V.code_unpacker_reference_vector()
    io, prev = instance(orc, reqs)
    want, got = run_both(engine, orc, io, reqs, prev, words, 40)
    assert want[0] == abi.ZKC_OK and want[1].completion_flag == 1
    assert_same(want, got)
    want, got = run_both(engine, orc, io, reqs, prev, words, 40, states=want[5])
    assert_same(want, got)


@pytest.mark.parametrize("n,max_words,extra", [(1, 1, 0), (1, 9, 3), (5, 15, 0), (40, 31, 7), (300, 63, 100), (3, 1001, 1)])
def test_synthetic_bit_exact(engine, orc, n, max_words, extra):
    reqs, words = synthetic.code_decommit_requests(n, seed=n + max_words, max_words=max_words)
    io, prev = instance(orc, reqs)
    limit = int(rounds_of(reqs).sum()) + extra
    want, got = run_both(engine, orc, io, reqs, prev, words, limit)
    assert want[0] == abi.ZKC_OK, (hex(want[4].failed_checks), want[4].first_bad_row)
    assert want[1].completion_flag == 1 and len(want[5]) == len(words)
    assert_same(want, got)
    want2, got2 = run_both(engine, orc, io, reqs, prev, words, limit, states=want[5])
    assert_same(want2, got2)


def test_chained_instances_and_edges(engine, orc):
    reqs, words = synthetic.code_decommit_requests(60, seed=9, max_words=41)
    io, prev = instance(orc, reqs)
    total = int(rounds_of(reqs).sum())
    whole = entry_point(engine, Witness(io, reqs, prev, words), total + 9)
    assert whole.closed_form_input.completion_flag == 1
    csum = np.cumsum(rounds_of(reqs))
    # cut inside a request, at a request boundary, and inside the idle tail
    for cut in (int(csum[20]) + 3, int(csum[33]), total + 4):
        a = entry_point(engine, Witness(io, reqs, prev, words), cut)
        popped = len(reqs) - a.closed_form_input.hidden_fsm_output.decommittment_requests_queue_state.length
        used = a.closed_form_input.hidden_fsm_output.memory_queue_state.length
        nxt = abi.CodeUnpackerClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
        nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
        want, got = run_both(engine, orc, nxt, reqs[popped:], prev[popped:], words[used:], total + 9 - cut)
        assert_same(want, got)
        assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(whole.closed_form_input.hidden_fsm_output)
        assert np.array_equal(np.concatenate([a.trace, got.trace], axis=1), whole.trace)
    exp = abi.CodeUnpackerClosedForm.from_buffer_copy(bytes(nxt))
    exp.hidden_fsm_output = got.closed_form_input.hidden_fsm_output
    exp.memory_queue_final_state = got.closed_form_input.memory_queue_final_state
    exp.completion_flag = 1
    ok = entry_point(engine, Witness(exp, reqs[popped:], prev[popped:], words[used:]), total + 9 - cut, compare_expected=True)
    assert ok.status.code == 0
    exp.memory_queue_final_state.length += 1
    with pytest.raises(Exception, match="FSM_OUTPUT_MISMATCH"):
        entry_point(engine, Witness(exp, reqs[popped:], prev[popped:], words[used:]), total + 9 - cut, compare_expected=True)
    # limit 0; empty queue (the first cycle pops from an empty queue: reported, idle afterwards)
    want, got = run_both(engine, orc, io, reqs, prev, words, 0)
    assert_same(want, got)
    e = np.zeros(0, dtype=abi.DECOMMIT_QUERY_DTYPE)
    io0, prev0 = instance(orc, e)
    want, got = run_both(engine, orc, io0, e, prev0, words[:0], 6)
    assert want[4].failed_checks & CHK["WITNESS_EXHAUSTED"]
    assert_same(want, got)


def test_negative_cases_match_oracle(engine, orc):
    reqs, words = synthetic.code_decommit_requests(30, seed=4, max_words=21)
    limit = int(rounds_of(reqs).sum()) + 3
    nw = (reqs["code_hash"][:, 7] & 0xFFFF).astype(np.int64)
    w2 = words.copy(); w2[int(nw[:7].sum()) + 2, 5] ^= 1 << 9
    r3 = reqs.copy(); r3["code_hash"][4][7] ^= 1 << 25
    r4 = reqs.copy(); r4["code_hash"][2][3] ^= 1
    for rr, ww in ((reqs, w2), (r3, words), (r4, words), (reqs, words[:-5])):
        io, prev = instance(orc, rr)
        want, got = run_both(engine, orc, io, rr, prev, ww, limit)
        assert want[0] == abi.ZKC_ERR_UNSATISFIED
        assert_same(want, got)
    # fewer requests supplied than the queue holds: the FSM asks for one nobody supplied
    io, prev = instance(orc, reqs)
    want, got = run_both(engine, orc, io, reqs[:10], prev[:10], words, limit)
    assert want[4].failed_checks & CHK["WITNESS_EXHAUSTED"] and want[4].first_bad_row == int(rounds_of(reqs)[:10].sum())
    assert_same(want, got)
    # corrupted hints
    want = O.code_unpacker_entry_point(orc, io, reqs, words, limit)
    t = want[5].copy(); t[50, 3] ^= 1
    r = entry_point(engine, Witness(io, reqs, prev, words, t), limit, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT
    p2 = prev.copy(); p2[7, 10] ^= 1
    r = entry_point(engine, Witness(io, reqs, p2, words), limit, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT


def test_device_resident(engine, orc):
    import torch
    reqs, words = synthetic.code_decommit_requests(2000, seed=13, max_words=127)
    io, prev = instance(orc, reqs)
    limit = int(rounds_of(reqs).sum()) + 17
    want = O.code_unpacker_entry_point(orc, io, reqs, words, limit)
    tod = lambda a: torch.from_numpy(a.view(np.uint8).reshape(len(a), -1)).cuda()
    t64 = lambda a: torch.from_numpy(a.view(np.int64)).cuda()
    w = Witness(io, tod(reqs), t64(prev), torch.from_numpy(words.view(np.int32)).cuda(), t64(want[5]))
    got = entry_point(engine, w, limit)
    torch.cuda.synchronize()
    assert got.commitment.tolist() == want[3].tolist() and got.status.code == 0
    assert np.array_equal(got.trace.cpu().numpy().view(np.uint64), want[2])


def test_check_trace_constraint_evaluation(engine, orc):
    """zkc_code_unpacker_check_trace: the ORACLE's trace satisfies every relation with and without the queue permutations; a fault
    injected into any relation family is found at its cycle; the engine's trace of a chained second instance (cut inside a bytecode)
    passes"""
    from era_zkevm_circuits_b200 import code_unpacker_check_trace
    V_ = abi.CUV
    reqs, words = synthetic.code_decommit_requests(80, seed=8, max_words=41)
    io, prev = instance(orc, reqs)
    limit = int(rounds_of(reqs).sum()) + 20
    want = O.code_unpacker_entry_point(orc, io, reqs, words, limit)
    assert want[0] == abi.ZKC_OK
    trace = want[2]
    for gates in (0, abi.GATES_GENERAL):
        viol, st = code_unpacker_check_trace(engine, io, trace, limit, gates)
        assert viol == 0 and st.code == 0, (gates, viol, hex(st.failed_checks), st.first_bad_row)
    import torch
    viol, st = code_unpacker_check_trace(engine, io, torch.from_numpy(trace.view(np.int64)).cuda(), limit, abi.GATES_GENERAL)
    assert viol == 0
    pops = np.flatnonzero(trace[K["FLAGS_IN"]])
    fins = np.flatnonzero(trace[K["FINALIZE"]])
    mid = int(np.flatnonzero((trace[K["FLAGS_IN"]] == 0) & (trace[K["PROCESS_SECOND_WORD"]] == 1))[50])  # in the middle of a bytecode
    faults = [
        (K["FLAGS_IN"] + 1, mid, None, V_["FSM"], 0),
        (K["REQUEST"] + 3, int(pops[5]), 1 << 33, V_["BOOLEAN"], 0),
        (K["REQ_LEN"], mid, None, V_["QUEUE"], 0),
        (K["REQ_HEAD"] + 7, mid, None, V_["QUEUE"], 0),
        (K["REQ_HEAD"] + 7, int(pops[6]), None, V_["ROUND_FUNCTION"], 0),
        (K["VERSION_MATCHES"], int(pops[7]), None, V_["LENGTH"], 0),
        (K["LENGTH_IN_ROUNDS"], int(pops[8]), None, V_["LENGTH"], 0),
        (K["LENGTH_IN_BITS"], mid, None, V_["SELECTS"], 0),
        (K["PAGE"], mid, None, V_["SELECTS"], 0),
        (K["HASH_TO_COMPARE"] + 2, mid, None, V_["SELECTS"], 0),
        (K["DECOMMIT"], limit - 3, None, V_["FSM"], 0),
        (K["NUM_ROUNDS_LEFT"], mid, None, V_["FSM"], 0),
        (K["WORD1"] + 4, int(fins[9]), None, V_["BOOLEAN"], 0),
        (K["INDEX1"], mid, None, V_["SELECTS"], 0),
        (K["MEM_TAIL0"] + 12, mid, None, V_["MEMORY_QUEUE"], 0),
        (K["MEM_TAIL1"] + 3, mid, None, V_["ROUND_FUNCTION"], 0),
        (K["MEM_TAIL1"] + 3, limit - 3, None, V_["MEMORY_QUEUE"], abi.GATES_GENERAL),
        (K["MESSAGE"] + 15, int(fins[10]), None, V_["COMPRESSION"], 0),
        (K["STATE_IN"] + 4, mid, None, V_["COMPRESSION"], 0),
        (K["STATE_NEW"] + 6, mid, None, V_["COMPRESSION"], 0),
        (K["WORD0"] + 1, int(fins[11]), None, V_["ENFORCE"], abi.GATES_GENERAL),
        (K["FLAGS_OUT"], int(fins[12]), None, V_["FSM"], 0),
    ]
    for col, row, val, bit, gates in faults:
        bad = trace.copy()
        bad[col, row] = np.uint64(val) if val is not None else bad[col, row] ^ np.uint64(1)
        viol, st = code_unpacker_check_trace(engine, io, bad, limit, gates)
        assert viol >= 1 and st.first_bad_row == row and st.failed_checks & bit, (col, row, viol, st.first_bad_row, hex(st.failed_checks))
    cut = mid
    a = entry_point(engine, Witness(io, reqs, prev, words, None), cut)
    nxt = abi.CodeUnpackerClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    used_req = len(reqs) - a.closed_form_input.hidden_fsm_output.decommittment_requests_queue_state.length
    used_words = int(a.closed_form_input.hidden_fsm_output.memory_queue_state.length)
    b = entry_point(engine, Witness(nxt, reqs[used_req:], prev[used_req:], words[used_words:], None), limit - cut)
    assert b.status.code == 0
    viol, st = code_unpacker_check_trace(engine, nxt, b.trace, limit - cut)
    assert viol == 0, (viol, hex(st.failed_checks), st.first_bad_row)
