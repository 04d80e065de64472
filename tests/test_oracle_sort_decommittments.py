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


def test_row_relations_of_the_trace(orc):
    """The row-to-row relations zkc_sort_decommittments_check_trace evaluates on the device (dq_check_kernel), restated in numpy and
    held against the oracle's trace: key comparison, same-hash / first-marker / same-page flags, the record to add with the
    first-encountered timestamp, queue bookkeeping.  Pins the evaluator's reading of mod.rs:235-381 without a GPU."""
    n, limit = 2000, 2100
    u, s = synthetic.decommit_requests_trace(n, seed=5, n_hashes=60)
    io, _, _ = instance(orc, u, s)
    rc, out, T, _, st, _ = O.sort_decommittments_entry_point(orc, io, u, s, limit)
    assert rc == 0
    K = abi.DQ_COLS
    col = lambda name, i=0: T[K[name] + i]
    prev = lambda a, first: np.concatenate([np.array([first], dtype=np.uint64), a[:-1].astype(np.uint64)])
    S = K["SORTED_ITEM"]
    ts, page = T[S + 10], T[S + 8]
    cur = [ts] + [T[S + i] for i in range(8)]
    pk = [prev(c, 0) for c in cur]
    borrow, all_eq = np.zeros(limit, np.uint64), np.ones(limit, np.uint64)
    for i in range(9):
        df, bo, le = col("CMP_DIFF", i), col("CMP_BORROW", i), col("CMP_LIMB_EQ", i)
        assert np.array_equal(pk[i] + (bo << np.uint64(32)), df + cur[i] + borrow) and np.array_equal(le, df == 0)
        borrow, all_eq = bo, all_eq & le
    assert np.array_equal(col("KEYS_ARE_EQUAL"), all_eq)
    same = np.ones(limit, bool)
    for i in range(8):
        same &= prev(T[S + i], 0) == T[S + i]
    same_hash, pop, trivial = col("SAME_HASH"), col("SHOULD_POP"), col("ORIGINAL_IS_EMPTY")
    pit = prev(trivial, 1)
    assert np.array_equal(same_hash, same) and np.array_equal(col("PREVIOUS_IS_TRIVIAL"), pit)
    assert np.array_equal(col("ENFORCE_MUST_BE_FIRST"), (1 - same_hash) & pop) and np.array_equal(col("ENFORCE_SAME_MEMORY_PAGE"), same_hash & (1 - pit))
    add = col("ADD_TO_QUEUE")
    assert np.array_equal(add, (1 - pit) & (1 - same_hash)) and add.sum() == 60  # 60 hashes: the last one is pushed by the first padding row (zero hash, previous item not trivial)
    fts = col("FIRST_TIMESTAMP")
    assert np.array_equal(fts, np.where(same_hash == 0, ts, prev(fts, 0)))
    P = K["PUSH_ITEM"]
    for i in range(8):
        assert np.array_equal(T[P + i], prev(T[S + i], 0))
    assert np.array_equal(T[P + 8], prev(page, 0)) and (T[P + 9] == 1).all() and np.array_equal(T[P + 10], prev(fts, 0))
    assert np.array_equal(col("RESULT_LEN"), np.cumsum(add))
    for i in range(12):
        t = col("RESULT_TAIL", i)
        assert np.array_equal(t[add == 0], prev(t, 0)[add == 0])
    for k, base, q0 in ((0, K["UNSORTED_ITEM"], io.initial_queue_state), (1, K["SORTED_ITEM"], io.sorted_queue_initial_state)):
        ln = T[base + 31]
        assert np.array_equal(ln + pop, prev(ln, q0.length))
        for i in range(12):
            h = T[base + 19 + i]
            assert np.array_equal(h[pop == 0], prev(h, q0.head[i])[pop == 0])
