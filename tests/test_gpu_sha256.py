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
