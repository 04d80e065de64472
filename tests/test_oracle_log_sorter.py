"""log_sorter oracle against the reference's own vector (/root/reference/src/log_sorter/mod.rs:494-816, limit = 16,
is_start = true: every enforcement holds) + rollback-collapsing behaviour + negative cases."""
import numpy as np

import orc as O
import vectors as V
from era_zkevm_circuits_b200 import abi, synthetic

K = abi.EV_COLS
CHK = abi.EV_CHK


def instance(orc, u, s):
    up, ufin = O.log_queue_simulate(orc, u)
    sp, sfin = O.log_queue_simulate(orc, s)
    return O.events_closed_form(ufin, sfin, True), up, sp


def test_encoding_layout(orc):
    q = np.zeros(1, dtype=abi.LOG_QUERY_DTYPE)
    q["key"][0] = [0x04030201, 0x08070605, 0x0C0B0A09, 0x100F0E0D, 0x14131211, 0x18171615, 0x1C1B1A19, 0x201F1E1D]
    q["address"][0] = [0x24232221, 0x28272625, 0x2C2B2A29, 0x302F2E2D, 0x34333231]
    q["read_value"][0] = np.arange(100, 108)
    q["written_value"][0] = np.arange(200, 208)
    q["timestamp"], q["tx_number_in_block"] = 77, 88
    q["flags"] = abi.lq_flags(aux=0xA1, shard=0xB2, rw=1, rollback=1, service=1)
    e = np.zeros(20, dtype=np.uint64)
    orc.orc_log_query_encode(O.p(q), O.p(e))
    stream = list(range(1, 0x35))  # key bytes 1..32 then address bytes 0x21..0x34
    for i in range(16):
        w = 100 + i if i < 8 else 200 + i - 8
        assert int(e[i]) == w + (stream[3 * i] << 32) + (stream[3 * i + 1] << 40) + (stream[3 * i + 2] << 48)
    assert int(e[16]) == 77 + (0x31 << 32) + (0x32 << 40) + (0x33 << 48)
    assert int(e[17]) == 88 + (0x34 << 32) + (0xA1 << 40) + (0xB2 << 48)
    assert int(e[18]) == 3 and int(e[19]) == 1


def test_reference_vector_is_satisfied(orc):
    u, s = V.log_sorter_reference_vector()
    io, _, _ = instance(orc, u, s)
    rc, out, trace, com, st, tails = O.log_sorter_entry_point(orc, io, u, s, 16)
    assert rc == abi.ZKC_OK and st.failed_checks == 0
    assert out.completion_flag == 1
    assert list(out.hidden_fsm_output.lhs_accumulator) == list(out.hidden_fsm_output.rhs_accumulator)
    # 4 distinct timestamps, none rolled back: all 4 reach the result queue (3 inside the loop with the one-row
    # delay + row 4 flushes the last)
    assert out.final_queue_state.length == 4 and len(tails) == 4
    assert trace[K["ADD_TO_QUEUE"]].tolist() == [0, 1, 1, 1, 1] + [0] * 11


def test_rollbacks_are_collapsed(orc):
    u, s = synthetic.events_trace(400, seed=3, rollback_pct=25)
    n_rb = int(((u["flags"] >> 17) & 1).sum())
    assert n_rb > 50
    io, _, _ = instance(orc, u, s)
    rc, out, trace, com, st, tails = O.log_sorter_entry_point(orc, io, u, s, 512)
    assert rc == abi.ZKC_OK, (hex(st.failed_checks), st.first_bad_row)
    assert out.final_queue_state.length == 400 - 2 * n_rb == len(tails)
    # chained halves == whole
    rc, a, ta, _, st, t1 = O.log_sorter_entry_point(orc, io, u, s, 150)
    nxt = abi.EventsClosedForm.from_buffer_copy(bytes(a)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.hidden_fsm_output
    rc, b, tb, _, st, t2 = O.log_sorter_entry_point(orc, nxt, u[150:], s[150:], 362)
    assert rc == abi.ZKC_OK, hex(st.failed_checks)
    assert bytes(b.hidden_fsm_output) == bytes(out.hidden_fsm_output)
    assert np.array_equal(np.concatenate([ta, tb], axis=1), trace)


def test_negative_cases(orc):
    u, s = synthetic.events_trace(200, seed=4, rollback_pct=20)
    io, _, _ = instance(orc, u, s)
    s2 = s.copy(); s2[[10, 60]] = s2[[60, 10]]
    io2, _, _ = instance(orc, u, s2)
    rc, _, _, _, st, _ = O.log_sorter_entry_point(orc, io2, u, s2, 256)
    assert st.failed_checks & CHK["ORDER"] and st.first_bad_row == 11
    rb = int(np.flatnonzero((s["flags"] >> 17) & 1)[0])
    s3 = s.copy(); s3["written_value"][rb][0] ^= 1
    u3 = u.copy(); j = int(np.flatnonzero((u["timestamp"] == s[rb]["timestamp"]) & ((u["flags"] >> 17) & 1 == 1))[0]); u3["written_value"][j][0] ^= 1
    io3, _, _ = instance(orc, u3, s3)
    rc, _, _, _, st, _ = O.log_sorter_entry_point(orc, io3, u3, s3, 256)
    assert st.failed_checks == CHK["SAME_BODY"] and st.first_bad_row == rb
    s4 = s.copy(); s4["flags"][5] ^= (1 << 16)
    u4 = u.copy(); j = int(np.flatnonzero(u["timestamp"] == s[5]["timestamp"])[0]); u4["flags"][j] ^= (1 << 16)
    io4, _, _ = instance(orc, u4, s4)
    rc, _, _, _, st, _ = O.log_sorter_entry_point(orc, io4, u4, s4, 256)
    assert st.failed_checks == CHK["UNSORTED_IS_WRITE"] | CHK["SORTED_IS_WRITE"]
    u5 = u.copy(); u5["key"][7][1] ^= 2
    io5, _, _ = instance(orc, u5, s)
    rc, _, _, _, st, _ = O.log_sorter_entry_point(orc, io5, u5, s, 256)
    assert st.failed_checks == CHK["GRAND_PRODUCT"]


def test_row_relations_of_the_trace(orc):
    """The row-to-row relations zkc_log_sorter_check_trace evaluates on the device (ev_check_kernel), restated in numpy and held against
    the oracle's trace: queue bookkeeping, the timestamp borrow chain, the flag algebra of :327-372, the pushed record, the result
    queue's length / tail selection.  Pins the evaluator's reading of log_sorter/mod.rs:283-403 without a GPU."""
    n, limit = 2000, 2100
    u, s = synthetic.events_trace(n, seed=5, rollback_pct=20)
    io, _, _ = instance(orc, u, s)
    rc, out, T, _, st, _ = O.log_sorter_entry_point(orc, io, u, s, limit)
    assert rc == 0
    K = abi.EV_COLS
    col = lambda name, i=0: T[K[name] + i]
    prev = lambda a, first: np.concatenate([np.array([first], dtype=np.uint64), a[:-1].astype(np.uint64)])
    pop, emp = col("SHOULD_POP"), col("ORIGINAL_IS_EMPTY")
    for base, q0 in ((K["UNSORTED_ITEM"], io.initial_log_queue_state), (K["SORTED_ITEM"], io.intermediate_sorted_queue_state)):
        ln = T[base + 60]
        assert np.array_equal(ln + pop, prev(ln, q0.length))
        for i in range(4):
            h = T[base + 56 + i]
            assert np.array_equal(h[pop == 0], prev(h, q0.head[i])[pop == 0])
    S = K["SORTED_ITEM"]
    ts, rollback = T[S + 35], T[S + 31]
    diff, borrow, keys_equal = col("CMP_DIFF"), col("CMP_BORROW"), col("KEYS_EQUAL")
    assert np.array_equal(ts + (borrow << np.uint64(32)), diff + prev(ts, 0)) and np.array_equal(keys_equal, diff == 0)  # current - previous
    same_nt, diff_nt, ike, ve = col("SAME_NONTRIVIAL_LOG"), col("DIFFERENT_NONTRIVIAL_LOG"), col("ITEM_KEYS_EQUAL"), col("VALUES_EQUAL")
    keq, veq = np.ones(limit, bool), np.ones(limit, bool)
    for i in range(8):
        keq &= T[S + 5 + i] == prev(T[S + 5 + i], 0)
        veq &= T[S + 21 + i] == prev(T[S + 21 + i], 0)
    pit = col("PREVIOUS_IS_TRIVIAL")
    assert np.array_equal(same_nt, pop & keys_equal) and np.array_equal(diff_nt, pop & (1 - keys_equal)) and np.array_equal(ike, keq) and np.array_equal(ve, veq)
    assert np.array_equal(col("SAME_BODY"), ike & ve) and np.array_equal(pit, prev(emp, 1))
    should_enforce, maybe_add, add = col("SHOULD_ENFORCE"), col("MAYBE_ADD"), col("ADD_TO_QUEUE")
    assert np.array_equal(should_enforce, keys_equal & (1 - pit)) and np.array_equal(maybe_add, (1 - keys_equal) | emp)
    assert np.array_equal(add, (1 - pit) & maybe_add & (1 - prev(rollback, 0)))
    # enforcements: ascending keys, a fresh key is not a rollback, a repeated key is one with the same body
    assert not ((pop & borrow) | (diff_nt & rollback) | (same_nt & (1 - rollback)) | (should_enforce & (1 - col("SAME_BODY")))).any()
    assert np.array_equal(col("RESULT_LEN"), np.cumsum(add))
    for i in range(4):
        t = col("RESULT_TAIL", i)
        assert np.array_equal(t, np.where(add == 1, col("PUSH_ROUND2", i), prev(t, 0)))
