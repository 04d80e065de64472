"""storage_validity oracle against the reference's vectors
(/root/reference/src/storage_validity_by_grand_product/test_input.rs, test mod.rs:1035-1160: limit = 16,
is_start = true; the reference asserts that every IN-LOOP enforcement holds -- its sorted vector is not a
permutation of the unsorted one, which only the entry point's lhs == rhs check would notice)."""
import numpy as np

import orc as O
import vectors as V
from era_zkevm_circuits_b200 import abi, synthetic

K = abi.ST_COLS
CHK = abi.ST_CHK


def instance(orc, u, s, ts, shard=0):
    up, ufin = O.log_queue_simulate(orc, u)
    sp, sfin = O.log_queue_simulate(orc, s, ts)
    return O.storage_closed_form(ufin, sfin, shard, True), up, sp


def test_reference_vector_loop_is_satisfied(orc):
    u, s, ts = V.storage_reference_vector()
    assert ts.tolist() == [27, 28, 22, 25, 26, 31, 19, 16, 13, 9, 10, 6, 32, 35, 36, 38]
    io, _, _ = instance(orc, u, s, ts)
    rc, out, trace, com, st, tails = O.storage_validity_entry_point(orc, io, u, s, ts, 16)
    # every enforcement inside sort_and_deduplicate_storage_access_inner holds; only the entry point's
    # grand-product equality fails (not a permutation)
    assert st.failed_checks == CHK["GRAND_PRODUCT"] and st.first_bad_row == -1
    assert out.completion_flag == 1
    assert trace[K["SHOULD_POP"]].sum() == 16


def test_synthetic_trace_is_valid_and_deduplicates(orc):
    n = 3000
    u, s, ts = synthetic.storage_trace(n, seed=7, n_cells=100)
    io, _, _ = instance(orc, u, s, ts)
    rc, out, trace, com, st, tails = O.storage_validity_entry_point(orc, io, u, s, ts, 3072)
    assert rc == abi.ZKC_OK, (hex(st.failed_checks), st.first_bad_row)
    assert out.completion_flag == 1
    assert list(out.hidden_fsm_output.lhs_accumulator) == list(out.hidden_fsm_output.rhs_accumulator)
    assert 0 < out.final_sorted_queue_state.length <= 100 and len(tails) == out.final_sorted_queue_state.length
    assert trace[K["WRITE_ROLLBACK"]].sum() > 50 and trace[K["READ_AT_DEPTH_ZERO_OF_SAME_CELL"]].sum() > 50
    assert trace[K["CELL_CURRENT_DEPTH"]].max() >= 2
    # chained halves == whole
    rc, a, ta, _, st, t1 = O.storage_validity_entry_point(orc, io, u, s, ts, 1300)
    assert rc == 0
    nxt = abi.StorageClosedForm.from_buffer_copy(bytes(a)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.hidden_fsm_output
    rc, b, tb, _, st, t2 = O.storage_validity_entry_point(orc, nxt, u[1300:], s[1300:], ts[1300:], 1772)
    assert rc == 0, hex(st.failed_checks)
    assert bytes(b.hidden_fsm_output) == bytes(out.hidden_fsm_output)
    assert np.array_equal(np.concatenate([ta, tb], axis=1), trace)
    assert np.array_equal(np.concatenate([t1, t2]), tails)


def test_negative_cases(orc):
    u, s, ts = synthetic.storage_trace(800, seed=8, n_cells=40)
    io, _, _ = instance(orc, u, s, ts)
    rd = int(np.flatnonzero(((s["flags"] >> 16) & 1) == 0)[30])
    s2 = s.copy(); s2["read_value"][rd][1] ^= 1
    u2 = u.copy(); u2["read_value"][ts[rd]][1] ^= 1
    io2, _, _ = instance(orc, u2, s2, ts)
    rc, _, tr, _, st, _ = O.storage_validity_entry_point(orc, io2, u2, s2, ts, 800)
    assert st.failed_checks & (CHK["READ_CONSISTENCY"]) or tr[K["NEW_NON_TRIVIAL_CELL"]][rd] == 1
    s3 = s.copy(); ts3 = ts.copy(); s3[[100, 500]] = s3[[500, 100]]; ts3[[100, 500]] = ts3[[500, 100]]
    io3, _, _ = instance(orc, u, s3, ts3)
    rc, _, _, _, st, _ = O.storage_validity_entry_point(orc, io3, u, s3, ts3, 800)
    assert st.failed_checks & CHK["KEY_ORDER"]
    io4, _, _ = instance(orc, u, s, ts, shard=1)
    rc, _, _, _, st, _ = O.storage_validity_entry_point(orc, io4, u, s, ts, 800)
    assert st.failed_checks == CHK["SHARD_ID"] and st.first_bad_row == 0
    ts5 = ts.copy(); ts5[3] += 1
    io5, _, _ = instance(orc, u, s, ts5)
    rc, _, _, _, st, _ = O.storage_validity_entry_point(orc, io5, u, s, ts5, 800)
    assert st.failed_checks & CHK["GRAND_PRODUCT"]


def test_row_relations_of_the_trace(orc):
    """The row-to-row relations zkc_storage_validity_check_trace evaluates on the device (st_check_kernel), restated in numpy and
    held against the oracle's trace: the per-cell state machine (base / current value, rollback depth, explicit-read flag), the push
    decision, the timestamp / cycle-counter bookkeeping.  Pins the evaluator's reading of mod.rs:560-800 without a GPU."""
    n, limit = 3000, 3100
    u, s, ts = synthetic.storage_trace(n, seed=8, n_cells=40)
    io, _, _ = instance(orc, u, s, ts)
    rc, out, T, _, st, _ = O.storage_validity_entry_point(orc, io, u, s, ts, limit)
    assert rc == 0
    f = io.hidden_fsm_input
    col = lambda name, i=0: T[K[name] + i]
    prev = lambda a, first: np.concatenate([[first], a[:-1]]).astype(np.uint64)
    S = K["SORTED_ITEM"]
    rw = T[S + 30]
    new_cell, read_same, wnr, wrb = col("NEW_NON_TRIVIAL_CELL"), col("READ_OF_SAME_CELL"), col("WRITE_NO_ROLLBACK"), col("WRITE_ROLLBACK")
    b_depth, b_flag = prev(col("CELL_CURRENT_DEPTH"), f.this_cell_current_depth), prev(col("CELL_HAS_READ_AT_DEPTH_ZERO"), 0)
    a_depth = col("CELL_CURRENT_DEPTH")
    assert np.array_equal(a_depth, np.where(new_cell == 1, rw, (b_depth + wnr - wrb) & 0xFFFFFFFF))
    depth_zero, r0 = col("ROLLBACK_DEPTH_IS_ZERO"), col("READ_AT_DEPTH_ZERO_OF_SAME_CELL")
    assert np.array_equal(depth_zero, a_depth == 0) and np.array_equal(r0, depth_zero & read_same)
    assert np.array_equal(col("CELL_HAS_READ_AT_DEPTH_ZERO"), np.where(new_cell == 1, 1 - rw, b_flag | r0))
    req, unchanged = np.ones(limit, bool), np.ones(limit, bool)
    for i in range(8):
        rv, wv = T[S + 13 + i], T[S + 21 + i]
        b_cur, b_base = prev(col("CELL_CURRENT_VALUE", i), 0), prev(col("CELL_BASE_VALUE", i), 0)
        cur1 = np.where(new_cell == 1, np.where(rw == 1, wv, rv), b_cur)
        req &= cur1 == rv
        assert np.array_equal(col("CELL_CURRENT_VALUE", i), np.where(new_cell == 1, cur1, np.where(wnr == 1, wv, np.where(wrb == 1, rv, b_cur))))
        assert np.array_equal(col("CELL_BASE_VALUE", i), np.where((new_cell | r0) == 1, rv, b_base))
        unchanged &= b_cur == b_base
    assert np.array_equal(col("READ_IS_EQUAL_TO_CURRENT"), req) and np.array_equal(col("CHECK_READ_CONSISTENCY"), read_same | wnr)
    viu, dz = col("VALUE_IS_UNCHANGED"), col("CURRENT_DEPTH_IS_ZERO")
    assert np.array_equal(viu, unchanged) and np.array_equal(dz, b_depth == 0)
    ubnr, ipr, sw, su = col("UNCHANGED_BUT_NOT_BY_ROLLBACK"), col("ISSUE_PROTECTIVE_READ"), col("SHOULD_WRITE"), col("SHOULD_UPDATE")
    assert np.array_equal(ubnr, viu & (1 - dz)) and np.array_equal(ipr, b_flag | ubnr) and np.array_equal(sw, 1 - viu) and np.array_equal(su, ipr | sw)
    keys_equal, trivial = col("KEYS_ARE_EQUAL"), col("ORIGINAL_IS_EMPTY")
    prev_trivial = prev(trivial, 1)  # start_flag: the first row has no previous item (:574-575)
    push = col("SHOULD_PUSH")
    assert np.array_equal(push, (1 - prev_trivial) & (1 - keys_equal) & su) and push.sum() == 40
    assert np.array_equal(col("RESULT_LEN"), np.cumsum(push))
    ots = col("ORIGINAL_TIMESTAMP")
    assert np.array_equal(ots, np.arange(limit, dtype=np.uint64)) and np.array_equal(col("UNSORTED_EXT19"), col("UNSORTED_ENC", 19) + (ots << np.uint64(8)))
    sts = T[S + 36]
    assert np.array_equal(prev(sts, f.previous_timestamp) + (col("PREVIOUS_TIMESTAMP_IS_LESS") << np.uint64(32)), col("TS_DIFF") + sts)
    nt = 1 - trivial
    assert np.array_equal(col("MUST_ENFORCE"), keys_equal & nt) and np.array_equal(new_cell, nt & (1 - keys_equal))
