"""storage_validity_by_grand_product: CUDA path through the C ABI vs the CPU oracle, bit-exact.  Mirrors
/root/reference/src/storage_validity_by_grand_product/mod.rs:1035-1160 (reference vectors) and widens it."""
import numpy as np
import pytest

import orc as O
import vectors as V
from era_zkevm_circuits_b200 import (StorageDeduplicatorInstanceWitness, abi,
                                     sort_and_deduplicate_storage_access_entry_point, synthetic)

pytestmark = pytest.mark.gpu
K = abi.ST_COLS
CHK = abi.ST_CHK


def instance(orc, u, s, ts, shard=0):
    up, ufin = O.log_queue_simulate(orc, u)
    sp, sfin = O.log_queue_simulate(orc, s, ts)
    return O.storage_closed_form(ufin, sfin, shard, True), up, sp


def assert_same(want, got, check_trace=True):
    rc, io, trace, com, st, tails = want
    assert got.status.code == rc, (got.status.code, hex(got.status.failed_checks), got.status.first_bad_row)
    assert got.status.failed_checks == st.failed_checks and got.status.first_bad_row == st.first_bad_row
    assert got.closed_form_input.completion_flag == io.completion_flag
    assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(io.hidden_fsm_output)
    assert bytes(got.closed_form_input.final_sorted_queue_state) == bytes(io.final_sorted_queue_state)
    assert got.commitment.tolist() == com.tolist()
    if check_trace:
        bad = np.argwhere(got.trace != trace)
        assert bad.size == 0, f"first differing (col,row): {bad[:8].tolist()}"


def run_both(engine, orc, io, u, up, s, ts, sp, limit, tails=None, **kw):
    want = O.storage_validity_entry_point(orc, io, u, s, ts, limit)
    w = StorageDeduplicatorInstanceWitness(io, u, up, s, ts, sp, tails)
    got = sort_and_deduplicate_storage_access_entry_point(engine, w, limit, raise_on_unsatisfied=False, **kw)
    return want, got


def test_reference_vector(engine, orc):
    u, s, ts = V.storage_reference_vector()
    io, up, sp = instance(orc, u, s, ts)
    want, got = run_both(engine, orc, io, u, up, s, ts, sp, 16)
    assert want[4].failed_checks == CHK["GRAND_PRODUCT"]  # the reference's vector is not a permutation
    assert_same(want, got)


@pytest.mark.parametrize("n,limit,cells", [(1, 1, 1), (3, 4, 1), (255, 256, 7), (257, 257, 300), (1000, 1024, 20), (30000, 30000, 500)])
def test_synthetic_bit_exact(engine, orc, n, limit, cells):
    u, s, ts = synthetic.storage_trace(n, seed=n, n_cells=cells)
    io, up, sp = instance(orc, u, s, ts)
    want, got = run_both(engine, orc, io, u, up, s, ts, sp, limit)
    assert want[0] == abi.ZKC_OK, (hex(want[4].failed_checks), want[4].first_bad_row)
    assert_same(want, got)
    want2, got2 = run_both(engine, orc, io, u, up, s, ts, sp, limit, tails=want[5])
    assert_same(want2, got2)


def test_chained_instances(engine, orc):
    n = 4000
    u, s, ts = synthetic.storage_trace(n, seed=9, n_cells=60)
    io, up, sp = instance(orc, u, s, ts)
    W = StorageDeduplicatorInstanceWitness
    whole = sort_and_deduplicate_storage_access_entry_point(engine, W(io, u, up, s, ts, sp), n)
    cut = 1777
    a = sort_and_deduplicate_storage_access_entry_point(engine, W(io, u, up, s, ts, sp), cut)
    assert a.closed_form_input.completion_flag == 0
    nxt = abi.StorageClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    want, got = run_both(engine, orc, nxt, u[cut:], up[cut:], s[cut:], ts[cut:], sp[cut:], n - cut)
    assert_same(want, got)
    assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(whole.closed_form_input.hidden_fsm_output)
    assert np.array_equal(np.concatenate([a.trace, got.trace], axis=1), whole.trace)
    exp = abi.StorageClosedForm.from_buffer_copy(bytes(nxt))
    exp.hidden_fsm_output = got.closed_form_input.hidden_fsm_output
    exp.final_sorted_queue_state = got.closed_form_input.final_sorted_queue_state
    exp.completion_flag = 1
    ok = sort_and_deduplicate_storage_access_entry_point(engine, W(exp, u[cut:], up[cut:], s[cut:], ts[cut:], sp[cut:]), n - cut,
                                                         compare_expected=True)
    assert ok.status.code == 0
    exp.hidden_fsm_output.this_cell_current_depth += 1
    bad = sort_and_deduplicate_storage_access_entry_point(engine, W(exp, u[cut:], up[cut:], s[cut:], ts[cut:], sp[cut:]), n - cut,
                                                          compare_expected=True, raise_on_unsatisfied=False)
    assert bad.status.code == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH


def test_negative_cases_match_oracle(engine, orc):
    u, s, ts = synthetic.storage_trace(2500, seed=8, n_cells=40)
    cases = []
    rd = int(np.flatnonzero(((s["flags"] >> 16) & 1) == 0)[200])
    s2 = s.copy(); s2["read_value"][rd][1] ^= 1; u2 = u.copy(); u2["read_value"][ts[rd]][1] ^= 1
    cases.append((u2, s2, ts, 0))
    s3 = s.copy(); ts3 = ts.copy(); s3[[100, 1500]] = s3[[1500, 100]]; ts3[[100, 1500]] = ts3[[1500, 100]]
    cases.append((u, s3, ts3, 0))
    cases.append((u, s, ts, 1))
    ts5 = ts.copy(); ts5[3] += 1
    cases.append((u, s, ts5, 0))
    rb = int(np.flatnonzero((s["flags"] >> 17) & 1)[0])
    s6 = s.copy(); s6["flags"][rb - 1] ^= 1 << 17  # two rollbacks in a row: depth underflow
    cases.append((u, s6, ts, 0))
    for uu, ss, tt, shard in cases:
        io, up, sp = instance(orc, uu, ss, tt, shard)
        want, got = run_both(engine, orc, io, uu, up, ss, tt, sp, 2560)
        assert want[0] == abi.ZKC_ERR_UNSATISFIED
        assert_same(want, got)


def test_device_resident(engine, orc):
    import torch
    n = 6000
    u, s, ts = synthetic.storage_trace(n, seed=13, n_cells=100)
    io, up, sp = instance(orc, u, s, ts)
    want = O.storage_validity_entry_point(orc, io, u, s, ts, n)
    prev, fin = engine.log_queue_simulate(s, ts)
    assert np.array_equal(prev, sp) and bytes(fin[0]) == bytes(io.intermediate_sorted_queue_state)
    tod = lambda a: torch.from_numpy(a.view(np.uint8).reshape(len(a), -1)).cuda()
    t64 = lambda a: torch.from_numpy(a.view(np.int64)).cuda()
    w = StorageDeduplicatorInstanceWitness(io, tod(u), t64(up), tod(s), torch.from_numpy(ts.view(np.int32)).cuda(), t64(sp), t64(want[5]))
    got = sort_and_deduplicate_storage_access_entry_point(engine, w, n)
    torch.cuda.synchronize()
    assert got.commitment.tolist() == want[3].tolist()
    assert np.array_equal(got.trace.cpu().numpy().view(np.uint64), want[2])


def test_check_trace_constraint_evaluation(engine, orc):
    """zkc_storage_validity_check_trace: the ORACLE's trace (reads, writes, rollbacks, protective reads over 40 cells) satisfies
    every relation with and without the round-function gates; a fault injected into any relation family is found at its row"""
    from era_zkevm_circuits_b200 import storage_validity_check_trace
    V_ = abi.STV
    u, s, ts = synthetic.storage_trace(3000, seed=8, n_cells=40)
    io, up, sp = instance(orc, u, s, ts)
    limit = 3100
    want = O.storage_validity_entry_point(orc, io, u, s, ts, limit)
    assert want[0] == abi.ZKC_OK
    trace = want[2]
    for gates in (0, abi.GATES_GENERAL):
        viol, st = storage_validity_check_trace(engine, io, trace, limit, gates)
        assert viol == 0 and st.code == 0, (gates, viol, hex(st.failed_checks), st.first_bad_row)
    import torch
    viol, st = storage_validity_check_trace(engine, io, torch.from_numpy(trace.view(np.int64)).cuda(), limit, abi.GATES_GENERAL)
    assert viol == 0
    pushes = np.flatnonzero(trace[K["SHOULD_PUSH"]])
    same = np.flatnonzero(trace[K["NON_TRIVIAL_AND_SAME_CELL"]])
    faults = [
        (K["SHOULD_POP"], 17, 2, V_["BOOLEAN"], 0),
        (K["UNSORTED_ITEM"] + 7, 40, 1 << 33, V_["BOOLEAN"], 0),
        (K["ORIGINAL_TIMESTAMP"], 55, None, V_["BOOLEAN"], 0),
        (K["UNSORTED_ENC"] + 3, 99, None, V_["ENCODING"], 0),
        (K["UNSORTED_EXT19"], 98, None, V_["ENCODING"], 0),
        (K["SORTED_ENC"] + 19, 97, None, V_["ENCODING"], 0),
        (K["SORTED_LEN"], 123, None, V_["QUEUE_LEN"], 0),
        (K["UNSORTED_HEAD"] + 1, 3050, None, V_["QUEUE_LEN"], 0),
        (K["UNSORTED_HEAD"] + 1, 150, None, V_["ROUND_FUNCTION"], 0),
        (K["GP_CHAIN"] + 45, 200, None, V_["GP_CHAIN"], 0),
        (K["GP_ACC"] + 2, 300, None, V_["GP_ACC"], 0),
        (K["CMP_DIFF"] + 9, 400, None, V_["COMPARISON"], 0),
        (K["TS_DIFF"], 410, None, V_["COMPARISON"], 0),
        (K["WRITE_ROLLBACK"], 500, None, V_["FLAGS"], 0),
        (K["SHOULD_UPDATE"], 510, None, V_["FLAGS"], 0),
        (K["SHARD_ID_IS_VALID"], 520, None, V_["FLAGS"], 0),
        (K["CELL_CURRENT_VALUE"] + 3, int(same[20]), None, V_["CELL_STATE"], 0),
        (K["CELL_BASE_VALUE"] + 1, 530, None, V_["CELL_STATE"], 0),
        (K["CELL_CURRENT_DEPTH"], int(same[30]), None, V_["CELL_STATE"], 0),
        (K["CELL_HAS_READ_AT_DEPTH_ZERO"], 540, None, V_["CELL_STATE"], 0),
        (K["READ_IS_EQUAL_TO_CURRENT"], 550, None, V_["CELL_STATE"], 0),
        (K["PUSH_ENC"] + 17, 600, None, V_["ENCODING"], 0),
        (K["RESULT_LEN"], 700, None, V_["RESULT_QUEUE"], 0),
        (K["RESULT_TAIL"] + 2, int(pushes[5]), None, V_["RESULT_QUEUE"], 0),
        (K["PUSH_ROUND1"] + 5, int(pushes[10]), None, V_["ROUND_FUNCTION"], 0),
    ]
    for col, row, val, bit, gates in faults:
        bad = trace.copy()
        bad[col, row] = np.uint64(val) if val is not None else bad[col, row] ^ np.uint64(1)
        viol, st = storage_validity_check_trace(engine, io, bad, limit, gates)
        assert viol >= 1 and st.first_bad_row == row and st.failed_checks & bit, (col, row, viol, st.first_bad_row, hex(st.failed_checks))
    # the engine's own trace of a chained second instance (start_flag = 0: cell state / previous key from the FSM input)
    cut = 1500
    a = sort_and_deduplicate_storage_access_entry_point(engine, StorageDeduplicatorInstanceWitness(io, u, up, s, ts, sp), cut, raise_on_unsatisfied=False)
    nxt = abi.StorageClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    b = sort_and_deduplicate_storage_access_entry_point(engine, StorageDeduplicatorInstanceWitness(nxt, u[cut:], up[cut:], s[cut:], ts[cut:], sp[cut:]),
                                                        limit - cut, raise_on_unsatisfied=False)
    assert b.status.code == 0
    viol, st = storage_validity_check_trace(engine, nxt, b.trace, limit - cut)
    assert viol == 0, (viol, hex(st.failed_checks), st.first_bad_row)


@pytest.mark.parametrize("world,on_dev", [(2, False), (4, False), (3, True)])
def test_one_instance_cut_by_rows_over_ranks(engine, orc, world, on_dev):
    """sharding.storage_rows_local / storage_rows_finish with the ENGINE as the backend (the ranks run one after the other on this
    GPU; tests/test_sharding_gloo.py runs the same phases over a real process group): the rank traces, with their accumulator
    columns scaled after the exchange, concatenate to the whole instance's trace; every rank ends with the whole instance's
    closed form, status and commitment.  Cells span ~150 rows, so every cut lands inside a cell."""
    import torch
    from era_zkevm_circuits_b200 import sharding
    n, limit = 6000, 6100
    u, s, ts = synthetic.storage_trace(n, seed=33, n_cells=40)
    io, up, sp = instance(orc, u, s, ts)
    want = O.storage_validity_entry_point(orc, io, u, s, ts, limit)
    assert want[0] == abi.ZKC_OK
    tails = want[5]
    if on_dev:
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)).cuda()
        i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
        w = StorageDeduplicatorInstanceWitness(io, dev(u), i64(up), dev(s), torch.from_numpy(ts.astype(np.uint32).view(np.int32)).cuda(), i64(sp), i64(tails))
    else:
        w = StorageDeduplicatorInstanceWitness(io, u, up, s, ts, sp, tails)
    cum = np.concatenate([[0], np.cumsum(want[2][K["SHOULD_PUSH"]])]).astype(np.int64)
    offs = [int(cum[sharding.row_range(n, r, world)[0]]) for r in range(world)]

    def run(io_, u_, up_, s_, ts_, sp_, tails_, lim, want_trace):
        return sort_and_deduplicate_storage_access_entry_point(engine, StorageDeduplicatorInstanceWitness(io_, u_, up_, s_, ts_, sp_, tails_), lim,
                                                               want_trace=want_trace, raise_on_unsatisfied=False)

    commit = lambda e: engine.commit_encoding(np.ascontiguousarray(e, dtype=np.uint64).reshape(1, -1))[0]
    locs = [sharding.storage_rows_local(run, w, limit, r, world, offs) for r in range(world)]
    recs = np.stack([l[3] for l in locs])
    traces = []
    for r in range(world):
        com, io_g, trace, st = sharding.storage_rows_finish(locs[r][0], r, world, recs, io, offs, engine.scale_accumulators, commit)
        assert st.code == 0, (r, st.code, hex(st.failed_checks), st.first_bad_row)
        assert com.tolist() == want[3].tolist()
        assert bytes(io_g.hidden_fsm_output) == bytes(want[1].hidden_fsm_output) and io_g.completion_flag == want[1].completion_flag
        assert bytes(io_g.final_sorted_queue_state) == bytes(want[1].final_sorted_queue_state)
        traces.append(trace.cpu().numpy().view(np.uint64) if on_dev else trace)
    bad = np.argwhere(np.concatenate(traces, axis=1) != want[2])
    assert bad.size == 0, f"first differing (col,row): {bad[:8].tolist()}"
    # a wrong push offset is a wrong result-queue tail for that rank: its own hint verification fails
    wrong = list(offs); wrong[1] += 1
    res, lo, hi, rec = sharding.storage_rows_local(run, w, limit, 1, world, wrong)
    assert res.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT
