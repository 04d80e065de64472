"""sort_decommittment_requests oracle against the reference's own vector
(/root/reference/src/sort_decommittment_requests/mod.rs:420-1390, limit = 16, is_start = true: every enforcement holds),
an independent Python model of the deduplication, chaining over instances and negative cases."""
import numpy as np

import orc as O
import vectors as V
from era_zkevm_circuits_b200 import abi, synthetic

K = abi.DQ_COLS
CHK = abi.DQ_CHK


def instance(orc, u, s):
    up, ufin = O.decommit_queue_simulate(orc, u)
    sp, sfin = O.decommit_queue_simulate(orc, s)
    return O.decommit_sorter_closed_form(ufin, sfin, True), up, sp


def hash_of(q):
    return sum(int(q["code_hash"][i]) << (32 * i) for i in range(8))


def test_encoding_layout(orc):
    q = np.zeros(1, dtype=abi.DECOMMIT_QUERY_DTYPE)
    q["code_hash"][0] = np.arange(100, 108)
    q["page"], q["timestamp"], q["is_first"] = 0x04030201, 0x14131211, 1
    e = np.zeros(8, dtype=np.uint64)
    orc.orc_decommit_query_encode(O.p(q), O.p(e))
    # decommit_query/mod.rs:31-107
    assert int(e[0]) == 100 + (0x01 << 32) + (0x02 << 40) + (0x03 << 48)
    assert int(e[1]) == 101 + (0x04 << 32) + (0x11 << 40) + (0x12 << 48)
    assert int(e[2]) == 102 + (0x13 << 32) + (0x14 << 40) + (1 << 48)
    assert e[3:].tolist() == [103, 104, 105, 106, 107]


def test_reference_vector_is_satisfied(orc):
    u, s = V.sort_decommittments_reference_vector()
    assert len(u) == len(s) == 29
    io, _, _ = instance(orc, u, s)
    rc, out, trace, com, st, states = O.sort_decommittments_entry_point(orc, io, u, s, 16)
    # the reference test runs 16 of the 29 rows and asserts check_if_satisfied
    assert rc == abi.ZKC_OK and st.failed_checks == 0, (hex(st.failed_checks), st.first_bad_row)
    assert out.completion_flag == 0 and out.hidden_fsm_output.initial_queue_state.length == 13
    # rows 0..15 of the sorted vector hold hashes A A A B C C ...: A is flushed at row 3, B at row 4
    assert trace[K["ADD_TO_QUEUE"]].tolist() == [0, 0, 0, 1, 1] + [0] * 11
    assert trace[K["SAME_HASH"]].tolist() == [0, 1, 1, 0, 0] + [1] * 11
    assert out.hidden_fsm_output.final_queue_state.length == 2 == len(states)
    # the pushed records carry the timestamp of the FIRST request of their hash and is_first = 1
    assert trace[K["PUSH_ITEM"] + 10][3] == s[0]["timestamp"] and trace[K["PUSH_ITEM"] + 9][3] == 1
    assert trace[K["PUSH_ITEM"] + 10][4] == s[3]["timestamp"]
    assert out.hidden_fsm_output.first_encountered_timestamp == s[4]["timestamp"]
    assert list(out.hidden_fsm_output.previous_packed_key) == [int(s[15]["timestamp"])] + s[15]["code_hash"].tolist()


def model(s):
    """deduplicated queue the out-of-circuit sorter produces: one record per hash, first timestamp, is_first = 1"""
    out = []
    for q in s:
        if not out or hash_of(out[-1]) != hash_of(q):
            r = q.copy(); r["is_first"] = 1
            out.append(r)
    return np.array(out, dtype=abi.DECOMMIT_QUERY_DTYPE)


def test_deduplication_matches_model_and_chains(orc):
    u, s = synthetic.decommit_requests_trace(500, seed=5, n_hashes=40)
    assert int(u["is_first"].sum()) == 40
    io, _, _ = instance(orc, u, s)
    rc, out, trace, com, st, states = O.sort_decommittments_entry_point(orc, io, u, s, 512)
    assert rc == abi.ZKC_OK, (hex(st.failed_checks), st.first_bad_row)
    assert out.completion_flag == 1 and list(out.hidden_fsm_output.lhs_accumulator) == list(out.hidden_fsm_output.rhs_accumulator)
    want = model(s)
    assert out.final_queue_state.length == len(want) == 40 == len(states)
    # the result queue is exactly the queue of the model's records
    _, fin = O.decommit_queue_simulate(orc, want)
    assert list(fin.tail) == list(out.final_queue_state.tail) == states[-1].tolist()
    # chained instances == whole
    rc, a, ta, _, st, s1 = O.sort_decommittments_entry_point(orc, io, u, s, 177)
    assert rc == abi.ZKC_OK and a.completion_flag == 0
    nxt = abi.DecommitSorterClosedForm.from_buffer_copy(bytes(a)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.hidden_fsm_output
    rc, b, tb, com_b, st, s2 = O.sort_decommittments_entry_point(orc, nxt, u[177:], s[177:], 335)
    assert rc == abi.ZKC_OK, hex(st.failed_checks)
    assert bytes(b.hidden_fsm_output) == bytes(out.hidden_fsm_output) and bytes(b.final_queue_state) == bytes(out.final_queue_state)
    assert np.array_equal(np.concatenate([ta, tb], axis=1), trace)
    assert np.array_equal(np.concatenate([s1, s2]), states)
    # compare_expected
    exp = abi.DecommitSorterClosedForm.from_buffer_copy(bytes(b))
    exp.start_flag = 0; exp.hidden_fsm_input = a.hidden_fsm_output
    rc, *_ = O.sort_decommittments_entry_point(orc, exp, u[177:], s[177:], 335, compare_expected=True)
    assert rc == abi.ZKC_OK
    exp.hidden_fsm_output.first_encountered_timestamp ^= 1
    rc, *_ = O.sort_decommittments_entry_point(orc, exp, u[177:], s[177:], 335, compare_expected=True)
    assert rc == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH


def test_edge_cases(orc):
    # empty queues: nothing to do, completed at once, empty output
    e = np.zeros(0, dtype=abi.DECOMMIT_QUERY_DTYPE)
    io, _, _ = instance(orc, e, e)
    rc, out, trace, com, st, states = O.sort_decommittments_entry_point(orc, io, e, e, 8)
    assert rc == abi.ZKC_OK and out.completion_flag == 1 and out.final_queue_state.length == 0 and len(states) == 0
    # a single request; limit == queue length: the finalisation step flushes it
    u, s = synthetic.decommit_requests_trace(1, seed=1)
    io, _, _ = instance(orc, u, s)
    rc, out, trace, com, st, states = O.sort_decommittments_entry_point(orc, io, u, s, 1)
    assert rc == abi.ZKC_OK and out.completion_flag == 1 and out.final_queue_state.length == 1
    # limit 0
    rc, out, _, _, st, _ = O.sort_decommittments_entry_point(orc, io, u, s, 0)
    assert rc == abi.ZKC_OK and out.completion_flag == 0 and out.hidden_fsm_output.initial_queue_state.length == 1


def test_negative_cases(orc):
    u, s = synthetic.decommit_requests_trace(200, seed=4, n_hashes=16)
    s2 = s.copy(); s2[[10, 60]] = s2[[60, 10]]
    io2, _, _ = instance(orc, u, s2)
    rc, _, _, _, st, _ = O.sort_decommittments_entry_point(orc, io2, u, s2, 256)
    assert st.failed_checks & CHK["ORDER"] and st.first_bad_row <= 11
    # first request of a hash without the marker
    first = int(np.flatnonzero(s["is_first"])[3])
    s3 = s.copy(); s3["is_first"][first] = 0
    u3 = u.copy(); u3["is_first"][int(np.flatnonzero(u["timestamp"] == s[first]["timestamp"])[0])] = 0
    io3, _, _ = instance(orc, u3, s3)
    rc, _, _, _, st, _ = O.sort_decommittments_entry_point(orc, io3, u3, s3, 256)
    assert st.failed_checks == CHK["MUST_BE_FIRST"] and st.first_bad_row == first
    # a repeated request pointing to another page
    rep = int(np.flatnonzero(s["is_first"] == 0)[7])
    s4 = s.copy(); s4["page"][rep] += 8
    u4 = u.copy(); u4["page"][int(np.flatnonzero(u["timestamp"] == s[rep]["timestamp"])[0])] += 8
    io4, _, _ = instance(orc, u4, s4)
    rc, _, _, _, st, _ = O.sort_decommittments_entry_point(orc, io4, u4, s4, 256)
    assert st.failed_checks & CHK["SAME_MEMORY_PAGE"] and st.first_bad_row == rep
    # not a permutation
    s5 = s.copy(); s5["page"][rep] += 8
    io5, _, _ = instance(orc, u, s5)
    rc, _, _, _, st, _ = O.sort_decommittments_entry_point(orc, io5, u, s5, 256)
    assert st.failed_checks & CHK["GRAND_PRODUCT"]
    # queue lengths differ
    io6, _, _ = instance(orc, u, s[:-1])
    rc, _, _, _, st, _ = O.sort_decommittments_entry_point(orc, io6, u, s[:-1], 256)
    assert st.failed_checks & CHK["LENGTHS_EQUAL"]
    # non-trivial head in the observable input
    io7, _, _ = instance(orc, u, s)
    io7.initial_queue_state.head[3] = 1
    rc, _, _, _, st, _ = O.sort_decommittments_entry_point(orc, io7, u, s, 256)
    assert st.failed_checks & CHK["TRIVIAL_HEAD"]
