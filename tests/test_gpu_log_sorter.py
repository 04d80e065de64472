"""log_sorter: CUDA path through the C ABI vs the CPU oracle, bit-exact (trace, FSM output, observable
output, commitment, status).  Mirrors /root/reference/src/log_sorter/mod.rs:494-635 and widens it."""
import numpy as np
import pytest

import orc as O
import vectors as V
from era_zkevm_circuits_b200 import (EventsDeduplicatorInstanceWitness, abi, sort_and_deduplicate_events_entry_point,
                                     synthetic)

pytestmark = pytest.mark.gpu
K = abi.EV_COLS
CHK = abi.EV_CHK


def instance(orc, u, s):
    up, ufin = O.log_queue_simulate(orc, u)
    sp, sfin = O.log_queue_simulate(orc, s)
    return O.events_closed_form(ufin, sfin, True), up, sp


def assert_same(want, got, check_trace=True):
    rc, io, trace, com, st, tails = want
    assert got.status.code == rc, (got.status.code, hex(got.status.failed_checks), got.status.first_bad_row)
    assert got.status.failed_checks == st.failed_checks and got.status.first_bad_row == st.first_bad_row
    assert got.closed_form_input.completion_flag == io.completion_flag
    assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(io.hidden_fsm_output)
    assert bytes(got.closed_form_input.final_queue_state) == bytes(io.final_queue_state)
    assert got.commitment.tolist() == com.tolist()
    if check_trace:
        bad = np.argwhere(got.trace != trace)
        assert bad.size == 0, f"first differing (col,row): {bad[:5].tolist()}"


def run_both(engine, orc, io, u, up, s, sp, limit, tails=None, **kw):
    want = O.log_sorter_entry_point(orc, io, u, s, limit)
    w = EventsDeduplicatorInstanceWitness(io, u, up, s, sp, tails)
    got = sort_and_deduplicate_events_entry_point(engine, w, limit, raise_on_unsatisfied=False, **kw)
    return want, got


def test_reference_vector(engine, orc):
    u, s = V.log_sorter_reference_vector()
    io, up, sp = instance(orc, u, s)
    want, got = run_both(engine, orc, io, u, up, s, sp, 16)
    assert want[0] == abi.ZKC_OK
    assert_same(want, got)
    # with the host-supplied result-queue tails (verified, no sequential chain on the device)
    want, got = run_both(engine, orc, io, u, up, s, sp, 16, tails=want[5])
    assert_same(want, got)


@pytest.mark.parametrize("n,limit,rb", [(1, 1, 0), (2, 2, 100), (255, 256, 10), (257, 257, 30), (1000, 1024, 10), (20000, 20000, 10)])
def test_synthetic_bit_exact(engine, orc, n, limit, rb):
    u, s = synthetic.events_trace(n, seed=n, rollback_pct=rb)
    io, up, sp = instance(orc, u, s)
    want, got = run_both(engine, orc, io, u, up, s, sp, limit)
    assert want[0] == abi.ZKC_OK, (hex(want[4].failed_checks), want[4].first_bad_row)
    assert_same(want, got)
    want2, got2 = run_both(engine, orc, io, u, up, s, sp, limit, tails=want[5])
    assert_same(want2, got2)


def test_chained_instances_and_empty(engine, orc):
    u, s = synthetic.events_trace(3000, seed=9, rollback_pct=15)
    io, up, sp = instance(orc, u, s)
    whole = sort_and_deduplicate_events_entry_point(engine, EventsDeduplicatorInstanceWitness(io, u, up, s, sp), 3000)
    a = sort_and_deduplicate_events_entry_point(engine, EventsDeduplicatorInstanceWitness(io, u, up, s, sp), 1100)
    assert a.closed_form_input.completion_flag == 0
    nxt = abi.EventsClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    want, got = run_both(engine, orc, nxt, u[1100:], up[1100:], s[1100:], sp[1100:], 1900)
    assert_same(want, got)
    assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(whole.closed_form_input.hidden_fsm_output)
    assert np.array_equal(np.concatenate([a.trace, got.trace], axis=1), whole.trace)
    exp = abi.EventsClosedForm.from_buffer_copy(bytes(nxt))
    exp.hidden_fsm_output = got.closed_form_input.hidden_fsm_output
    exp.final_queue_state = got.closed_form_input.final_queue_state
    exp.completion_flag = 1
    ok = sort_and_deduplicate_events_entry_point(engine, EventsDeduplicatorInstanceWitness(exp, u[1100:], up[1100:], s[1100:], sp[1100:]),
                                                 1900, compare_expected=True)
    assert ok.status.code == 0
    exp.final_queue_state.length += 1
    bad = sort_and_deduplicate_events_entry_point(engine, EventsDeduplicatorInstanceWitness(exp, u[1100:], up[1100:], s[1100:], sp[1100:]),
                                                  1900, compare_expected=True, raise_on_unsatisfied=False)
    assert bad.status.code == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH
    # empty queues
    e = np.zeros(0, dtype=abi.LOG_QUERY_DTYPE)
    io0, up0, sp0 = instance(orc, e, e)
    want, got = run_both(engine, orc, io0, e, up0, e, sp0, 8)
    assert want[0] == abi.ZKC_OK
    assert_same(want, got)


def test_negative_cases_match_oracle(engine, orc):
    u, s = synthetic.events_trace(1500, seed=4, rollback_pct=20)
    cases = []
    s2 = s.copy(); s2[[10, 60]] = s2[[60, 10]]; cases.append((u, s2))
    rb = int(np.flatnonzero((s["flags"] >> 17) & 1)[3])
    s3 = s.copy(); s3["key"][rb][5] ^= 1; cases.append((u, s3))
    s4 = s.copy(); s4["flags"][700] ^= 1 << 17; cases.append((u, s4))
    u5 = u.copy(); u5["flags"][9] ^= 1 << 16; cases.append((u5, s))
    for uu, ss in cases:
        io, up, sp = instance(orc, uu, ss)
        want, got = run_both(engine, orc, io, uu, up, ss, sp, 1536)
        assert want[0] == abi.ZKC_ERR_UNSATISFIED
        assert_same(want, got)
    # corrupted hints
    io, up, sp = instance(orc, u, s)
    want = O.log_sorter_entry_point(orc, io, u, s, 1536)
    t = want[5].copy(); t[100, 2] ^= 1
    r = sort_and_deduplicate_events_entry_point(engine, EventsDeduplicatorInstanceWitness(io, u, up, s, sp, t), 1536, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT
    sp2 = sp.copy(); sp2[7, 0] ^= 1
    r = sort_and_deduplicate_events_entry_point(engine, EventsDeduplicatorInstanceWitness(io, u, up, s, sp2), 1536, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT and r.status.first_bad_row in (6, 7)


def test_device_resident_and_queue_simulate(engine, orc):
    import torch
    n = 5000
    u, s = synthetic.events_trace(n, seed=13)
    io, up, sp = instance(orc, u, s)
    want = O.log_sorter_entry_point(orc, io, u, s, n)
    prev, fin = engine.log_queue_simulate(u)
    assert np.array_equal(prev, up) and bytes(fin[0]) == bytes(io.initial_log_queue_state)
    tod = lambda a: torch.from_numpy(a.view(np.uint8).reshape(len(a), -1)).cuda()
    t64 = lambda a: torch.from_numpy(a.view(np.int64)).cuda()
    w = EventsDeduplicatorInstanceWitness(io, tod(u), t64(up), tod(s), t64(sp), t64(want[5]))
    got = sort_and_deduplicate_events_entry_point(engine, w, n)
    torch.cuda.synchronize()
    assert got.commitment.tolist() == want[3].tolist()
    assert np.array_equal(got.trace.cpu().numpy().view(np.uint64), want[2])


def test_check_trace_constraint_evaluation(engine, orc):
    """zkc_log_sorter_check_trace: the ORACLE's trace (reference vector + synthetic, rollbacks included) satisfies every
    relation with and without the round-function gates; a fault injected into any relation family is found at its row"""
    from era_zkevm_circuits_b200 import log_sorter_check_trace
    V_ = abi.EVV
    u, s = V.log_sorter_reference_vector()
    io, up, sp = instance(orc, u, s)
    want = O.log_sorter_entry_point(orc, io, u, s, 16)
    assert log_sorter_check_trace(engine, io, want[2], 16)[0] == 0
    u, s = synthetic.events_trace(3000, seed=8, rollback_pct=20)
    io, up, sp = instance(orc, u, s)
    limit = 3100
    want = O.log_sorter_entry_point(orc, io, u, s, limit)
    assert want[0] == abi.ZKC_OK
    trace = want[2]
    for gates in (0, abi.GATES_GENERAL):
        viol, st = log_sorter_check_trace(engine, io, trace, limit, gates)
        assert viol == 0 and st.code == 0, (viol, hex(st.failed_checks), st.first_bad_row)
    import torch
    viol, st = log_sorter_check_trace(engine, io, torch.from_numpy(trace.view(np.int64)).cuda(), limit, abi.GATES_GENERAL)
    assert viol == 0
    pushes = np.flatnonzero(trace[K["ADD_TO_QUEUE"]])
    faults = [
        (K["SHOULD_POP"], 17, 2, V_["BOOLEAN"], 0),
        (K["UNSORTED_ITEM"] + 7, 40, 1 << 33, V_["BOOLEAN"], 0),
        (K["UNSORTED_ENC"] + 3, 99, None, V_["ENCODING"], 0),
        (K["SORTED_LEN"], 123, None, V_["QUEUE_LEN"], 0),
        (K["GP_CHAIN"] + 45, 200, None, V_["GP_CHAIN"], 0),
        (K["GP_ACC"] + 2, 300, None, V_["GP_ACC"], 0),
        (K["CMP_DIFF"], 400, None, V_["COMPARISON"], 0),
        (K["SAME_BODY"], 500, None, V_["FLAGS"], 0),
        (K["PUSH_ENC"] + 17, 600, None, V_["ENCODING"], 0),
        (K["RESULT_LEN"], 700, None, V_["RESULT_QUEUE"], 0),
        (K["PUSH_ROUND1"] + 5, int(pushes[10]), None, V_["ROUND_FUNCTION"], 0),
    ]
    for col, row, val, bit, gates in faults:
        bad = trace.copy()
        bad[col, row] = np.uint64(val) if val is not None else bad[col, row] ^ np.uint64(1)
        viol, st = log_sorter_check_trace(engine, io, bad, limit, gates)
        assert viol >= 1 and st.first_bad_row == row and st.failed_checks & bit, (col, row, viol, st.first_bad_row, hex(st.failed_checks))
    # the engine's own trace of a chained second instance (start_flag = 0: accumulators / previous item from the FSM input)
    cut = 1500
    a = sort_and_deduplicate_events_entry_point(engine, EventsDeduplicatorInstanceWitness(io, u, up, s, sp), cut, raise_on_unsatisfied=False)
    nxt = abi.EventsClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    b = sort_and_deduplicate_events_entry_point(engine, EventsDeduplicatorInstanceWitness(nxt, u[cut:], up[cut:], s[cut:], sp[cut:]), limit - cut,
                                                raise_on_unsatisfied=False)
    assert b.status.code == 0
    viol, st = log_sorter_check_trace(engine, nxt, b.trace, limit - cut)
    assert viol == 0, (viol, hex(st.failed_checks), st.first_bad_row)


def test_one_instance_cut_by_rows_over_ranks(engine, orc):
    """sharding.events_rows_local / events_rows_finish with the ENGINE as the backend, 3 virtual ranks on this GPU: rank traces (accumulator
    columns scaled after the exchange) concatenate to the whole instance's trace; every rank ends with the whole closed form + commitment"""
    from era_zkevm_circuits_b200 import sharding
    n, limit = 5000, 5100
    u, s = synthetic.events_trace(n, seed=12, rollback_pct=15)
    io, up, sp = instance(orc, u, s)
    want = O.log_sorter_entry_point(orc, io, u, s, limit)
    assert want[0] == abi.ZKC_OK
    w = EventsDeduplicatorInstanceWitness(io, u, up, s, sp, want[5])
    world = 3
    cum = np.concatenate([[0], np.cumsum(want[2][K["ADD_TO_QUEUE"]])]).astype(np.int64)
    offs = [int(cum[sharding.row_range(n, r, world)[0]]) for r in range(world)]

    def run(io_, u_, up_, s_, sp_, tails_, lim, want_trace):
        return sort_and_deduplicate_events_entry_point(engine, EventsDeduplicatorInstanceWitness(io_, u_, up_, s_, sp_, tails_), lim,
                                                       want_trace=want_trace, raise_on_unsatisfied=False)

    commit = lambda e: engine.commit_encoding(np.ascontiguousarray(e, dtype=np.uint64).reshape(1, -1))[0]
    locs = [sharding.events_rows_local(run, w, limit, r, world, offs) for r in range(world)]
    recs = np.stack([l[3] for l in locs])
    traces = []
    for r in range(world):
        com, io_g, trace, st = sharding.events_rows_finish(locs[r][0], r, world, recs, io, offs, engine.scale_accumulators, commit)
        assert st.code == 0, (r, st.code, hex(st.failed_checks), st.first_bad_row)
        assert com.tolist() == want[3].tolist()
        assert bytes(io_g.hidden_fsm_output) == bytes(want[1].hidden_fsm_output) and bytes(io_g.final_queue_state) == bytes(want[1].final_queue_state)
        traces.append(trace)
    bad = np.argwhere(np.concatenate(traces, axis=1) != want[2])
    assert bad.size == 0, f"first differing (col,row): {bad[:8].tolist()}"
