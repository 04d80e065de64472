"""sort_decommittment_requests: CUDA path through the C ABI vs the CPU oracle, bit-exact (trace, FSM output, observable
output, commitment, status).  Mirrors /root/reference/src/sort_decommittment_requests/mod.rs:420-563 and widens it."""
import numpy as np
import pytest

import orc as O
import vectors as V
from era_zkevm_circuits_b200 import (CodeDecommittmentsDeduplicatorInstanceWitness as Witness, abi,
                                     sort_and_deduplicate_code_decommittments_entry_point as entry_point, synthetic)

pytestmark = pytest.mark.gpu
K = abi.DQ_COLS
CHK = abi.DQ_CHK


def instance(orc, u, s):
    up, ufin = O.decommit_queue_simulate(orc, u)
    sp, sfin = O.decommit_queue_simulate(orc, s)
    return O.decommit_sorter_closed_form(ufin, sfin, True), up, sp


def assert_same(want, got, check_trace=True):
    rc, io, trace, com, st, states = want
    assert got.status.code == rc, (got.status.code, hex(got.status.failed_checks), got.status.first_bad_row, rc, hex(st.failed_checks))
    assert got.status.failed_checks == st.failed_checks and got.status.first_bad_row == st.first_bad_row
    assert got.closed_form_input.completion_flag == io.completion_flag
    assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(io.hidden_fsm_output)
    assert bytes(got.closed_form_input.final_queue_state) == bytes(io.final_queue_state)
    assert got.commitment.tolist() == com.tolist()
    if check_trace:
        bad = np.argwhere(got.trace != trace)
        assert bad.size == 0, f"first differing (col,row): {bad[:5].tolist()}"


def run_both(engine, orc, io, u, up, s, sp, limit, states=None, **kw):
    want = O.sort_decommittments_entry_point(orc, io, u, s, limit)
    got = entry_point(engine, Witness(io, u, up, s, sp, states), limit, raise_on_unsatisfied=False, **kw)
    return want, got


def test_reference_vector(engine, orc):
    u, s = V.sort_decommittments_reference_vector()
    io, up, sp = instance(orc, u, s)
    want, got = run_both(engine, orc, io, u, up, s, sp, 16)
    assert want[0] == abi.ZKC_OK
    assert_same(want, got)
    # with the host-supplied result-queue states (verified, no sequential chain on the device)
    want, got = run_both(engine, orc, io, u, up, s, sp, 16, states=want[5])
    assert_same(want, got)


@pytest.mark.parametrize("n,limit,hashes", [(1, 1, 1), (2, 2, 1), (2, 3, 2), (255, 256, 7), (257, 257, 300), (1000, 1024, 50),
                                            (20000, 20000, 1000), (5000, 6000, 1), (20000, 60000, 3), (300, 100000, 2)])
def test_synthetic_bit_exact(engine, orc, n, limit, hashes):
    u, s = synthetic.decommit_requests_trace(n, seed=n, n_hashes=hashes)
    io, up, sp = instance(orc, u, s)
    want, got = run_both(engine, orc, io, u, up, s, sp, limit)
    assert want[0] == abi.ZKC_OK, (hex(want[4].failed_checks), want[4].first_bad_row)
    assert want[1].final_queue_state.length == int(u["is_first"].sum()) <= min(hashes, n)
    assert_same(want, got)
    want2, got2 = run_both(engine, orc, io, u, up, s, sp, limit, states=want[5])
    assert_same(want2, got2)


def test_chained_instances_and_empty(engine, orc):
    u, s = synthetic.decommit_requests_trace(3000, seed=9, n_hashes=64)
    io, up, sp = instance(orc, u, s)
    whole = entry_point(engine, Witness(io, u, up, s, sp), 3000)
    assert whole.closed_form_input.completion_flag == 1 and whole.closed_form_input.final_queue_state.length == 64
    # cut inside a run of equal hashes and at a run boundary
    for cut in (1100, int(np.flatnonzero(s["is_first"])[20])):
        a = entry_point(engine, Witness(io, u, up, s, sp), cut)
        assert a.closed_form_input.completion_flag == 0
        nxt = abi.DecommitSorterClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
        nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
        want, got = run_both(engine, orc, nxt, u[cut:], up[cut:], s[cut:], sp[cut:], 3000 - cut)
        assert_same(want, got)
        assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(whole.closed_form_input.hidden_fsm_output)
        assert np.array_equal(np.concatenate([a.trace, got.trace], axis=1), whole.trace)
    exp = abi.DecommitSorterClosedForm.from_buffer_copy(bytes(nxt))
    exp.hidden_fsm_output = got.closed_form_input.hidden_fsm_output
    exp.final_queue_state = got.closed_form_input.final_queue_state
    exp.completion_flag = 1
    ok = entry_point(engine, Witness(exp, u[cut:], up[cut:], s[cut:], sp[cut:]), 3000 - cut, compare_expected=True)
    assert ok.status.code == 0
    exp.final_queue_state.length += 1
    bad = entry_point(engine, Witness(exp, u[cut:], up[cut:], s[cut:], sp[cut:]), 3000 - cut, compare_expected=True,
                      raise_on_unsatisfied=False)
    assert bad.status.code == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH
    # empty queues, limit 0
    e = np.zeros(0, dtype=abi.DECOMMIT_QUERY_DTYPE)
    io0, up0, sp0 = instance(orc, e, e)
    want, got = run_both(engine, orc, io0, e, up0, e, sp0, 8)
    assert want[0] == abi.ZKC_OK
    assert_same(want, got)
    want, got = run_both(engine, orc, io, u, up, s, sp, 0)
    assert_same(want, got)


def test_negative_cases_match_oracle(engine, orc):
    u, s = synthetic.decommit_requests_trace(1500, seed=4, n_hashes=32)
    cases = []
    s2 = s.copy(); s2[[10, 60]] = s2[[60, 10]]; cases.append((u, s2))
    first = int(np.flatnonzero(s["is_first"])[3])
    s3 = s.copy(); s3["is_first"][first] = 0; cases.append((u, s3))
    rep = int(np.flatnonzero(s["is_first"] == 0)[700])
    s4 = s.copy(); s4["page"][rep] += 8; cases.append((u, s4))
    for uu, ss in cases:
        io, up, sp = instance(orc, uu, ss)
        want, got = run_both(engine, orc, io, uu, up, ss, sp, 1536)
        assert want[0] == abi.ZKC_ERR_UNSATISFIED
        assert_same(want, got)
    io, up, sp = instance(orc, u, s)
    io.sorted_queue_initial_state.head[11] = 5
    want, got = run_both(engine, orc, io, u, up, s, sp, 1536)
    # (the queue witness no longer chains from the claimed head either, which the engine reports on top)
    assert want[4].failed_checks & CHK["TRIVIAL_HEAD"] and got.status.failed_checks & CHK["TRIVIAL_HEAD"] and got.status.code != 0
    # a sorted witness shorter than the queue it claims to be is refused before any launch
    io, up, sp = instance(orc, u, s)
    with pytest.raises(Exception, match="INVALID_ARGUMENT"):
        entry_point(engine, Witness(io, u, up, s[:-1], sp[:-1]), 1536)
    # corrupted hints
    want = O.sort_decommittments_entry_point(orc, io, u, s, 1536)
    t = want[5].copy(); t[10, 9] ^= 1
    r = entry_point(engine, Witness(io, u, up, s, sp, t), 1536, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT
    r = entry_point(engine, Witness(io, u, up, s, sp, want[5][:-2]), 1536, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT
    sp2 = sp.copy(); sp2[7, 10] ^= 1
    r = entry_point(engine, Witness(io, u, up, s, sp2), 1536, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT and r.status.first_bad_row in (6, 7)


def test_device_resident_and_queue_simulate(engine, orc):
    import torch
    n = 5000
    u, s = synthetic.decommit_requests_trace(n, seed=13, n_hashes=200)
    io, up, sp = instance(orc, u, s)
    want = O.sort_decommittments_entry_point(orc, io, u, s, n)
    prev, fin = engine.decommit_queue_simulate(u)
    assert np.array_equal(prev, up) and bytes(fin[0]) == bytes(io.initial_queue_state)
    tod = lambda a: torch.from_numpy(a.view(np.uint8).reshape(len(a), -1)).cuda()
    t64 = lambda a: torch.from_numpy(a.view(np.int64)).cuda()
    prev_d, fin_d = engine.decommit_queue_simulate(tod(np.concatenate([u, s])), n_queues=2)
    assert np.array_equal(prev_d.cpu().numpy().view(np.uint64), np.concatenate([up, sp]))
    assert bytes(fin_d[1]) == bytes(io.sorted_queue_initial_state)
    got = entry_point(engine, Witness(io, tod(u), t64(up), tod(s), t64(sp), t64(want[5])), n)
    torch.cuda.synchronize()
    assert got.commitment.tolist() == want[3].tolist()
    assert np.array_equal(got.trace.cpu().numpy().view(np.uint64), want[2])


def test_full_size_properties(engine, orc):
    """2^20 requests over 2^14 hashes, device resident, no oracle: the sorted queue is a permutation of the original one
    (grand products meet), one output record per hash, chained halves == whole"""
    import torch
    n = 1 << 20
    u, s = synthetic.decommit_requests_trace(n, seed=77, n_hashes=1 << 14)
    tod = lambda a: torch.from_numpy(a.view(np.uint8).reshape(len(a), -1)).cuda()
    du, ds = tod(u), tod(s)
    prev, fin = engine.decommit_queue_simulate(torch.cat([du, ds]), n_queues=2)
    io = O.decommit_sorter_closed_form(fin[0], fin[1], True)
    whole = entry_point(engine, Witness(io, du, prev[:n], ds, prev[n:]), n, want_trace=False)
    out = whole.closed_form_input
    assert whole.status.code == 0 and out.completion_flag == 1
    assert list(out.hidden_fsm_output.lhs_accumulator) == list(out.hidden_fsm_output.rhs_accumulator)
    assert out.final_queue_state.length == 1 << 14
    half = n // 2 + 12345
    a = entry_point(engine, Witness(io, du, prev[:n], ds, prev[n:]), half, want_trace=False)
    nxt = abi.DecommitSorterClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    b = entry_point(engine, Witness(nxt, du[half:], prev[half:n], ds[half:], prev[n + half:]), n - half, want_trace=False)
    assert bytes(b.closed_form_input.hidden_fsm_output) == bytes(out.hidden_fsm_output)
    assert bytes(b.closed_form_input.final_queue_state) == bytes(out.final_queue_state)
    # the result queue is the queue of the deduplicated records (first timestamp, is_first = 1)
    firsts = s[np.flatnonzero(np.r_[True, (s["code_hash"][1:] != s["code_hash"][:-1]).any(axis=1)])].copy()
    firsts["is_first"] = 1
    _, fin2 = engine.decommit_queue_simulate(tod(firsts))
    assert list(fin2[0].tail) == list(out.final_queue_state.tail)


def test_check_trace_constraint_evaluation(engine, orc):
    """zkc_sort_decommittments_check_trace: the ORACLE's trace satisfies every relation with and without the round-function gates; a
    fault injected into any relation family is found at its row; the engine's own trace of a chained second instance passes"""
    from era_zkevm_circuits_b200 import sort_decommittments_check_trace
    V_ = abi.DQV
    n, limit = 3000, 3100
    u, s = synthetic.decommit_requests_trace(n, seed=8, n_hashes=70)
    io, up, sp = instance(orc, u, s)
    want = O.sort_decommittments_entry_point(orc, io, u, s, limit)
    assert want[0] == abi.ZKC_OK
    trace = want[2]
    for gates in (0, abi.GATES_GENERAL):
        viol, st = sort_decommittments_check_trace(engine, io, trace, limit, gates)
        assert viol == 0 and st.code == 0, (gates, viol, hex(st.failed_checks), st.first_bad_row)
    import torch
    viol, st = sort_decommittments_check_trace(engine, io, torch.from_numpy(trace.view(np.int64)).cuda(), limit, abi.GATES_GENERAL)
    assert viol == 0
    pushes = np.flatnonzero(trace[K["ADD_TO_QUEUE"]])
    faults = [
        (K["SHOULD_POP"], 17, 2, V_["BOOLEAN"], 0),
        (K["UNSORTED_ITEM"] + 3, 40, 1 << 33, V_["BOOLEAN"], 0),
        (K["UNSORTED_ENC"] + 2, 99, None, V_["ENCODING"], 0),
        (K["SORTED_LEN"], 123, None, V_["QUEUE_LEN"], 0),
        (K["SORTED_HEAD"] + 5, 3050, None, V_["QUEUE_LEN"], 0),
        (K["SORTED_HEAD"] + 5, 150, None, V_["ROUND_FUNCTION"], 0),
        (K["GP_CHAIN"] + 21, 200, None, V_["GP_CHAIN"], 0),
        (K["GP_ACC"] + 2, 300, None, V_["GP_ACC"], 0),
        (K["CMP_DIFF"] + 4, 400, None, V_["COMPARISON"], 0),
        (K["SAME_HASH"], 500, None, V_["FLAGS"], 0),
        (K["ADD_TO_QUEUE"], 510, None, V_["FLAGS"], 0),
        (K["FIRST_TIMESTAMP"], 520, None, V_["FLAGS"], 0),
        (K["PUSH_ITEM"] + 10, 530, None, V_["RESULT_QUEUE"], 0),
        (K["PUSH_ENC"] + 6, 600, None, V_["ENCODING"], 0),
        (K["RESULT_LEN"], 700, None, V_["RESULT_QUEUE"], 0),
        (K["RESULT_TAIL"] + 7, 700 + int(np.flatnonzero(trace[K["ADD_TO_QUEUE"], 700:] == 0)[0]), None, V_["RESULT_QUEUE"], abi.GATES_GENERAL),
        (K["RESULT_TAIL"] + 7, int(pushes[10]), None, V_["ROUND_FUNCTION"], 0),
    ]
    for col, row, val, bit, gates in faults:
        bad = trace.copy()
        bad[col, row] = np.uint64(val) if val is not None else bad[col, row] ^ np.uint64(1)
        viol, st = sort_decommittments_check_trace(engine, io, bad, limit, gates)
        assert viol >= 1 and st.first_bad_row == row and st.failed_checks & bit, (col, row, viol, st.first_bad_row, hex(st.failed_checks))
    cut = 1500
    a = entry_point(engine, Witness(io, u, up, s, sp), cut, raise_on_unsatisfied=False)
    nxt = abi.DecommitSorterClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    b = entry_point(engine, Witness(nxt, u[cut:], up[cut:], s[cut:], sp[cut:]), limit - cut, raise_on_unsatisfied=False)
    assert b.status.code == 0
    viol, st = sort_decommittments_check_trace(engine, nxt, b.trace, limit - cut)
    assert viol == 0, (viol, hex(st.failed_checks), st.first_bad_row)


def test_one_instance_cut_by_rows_over_ranks(engine, orc):
    """sharding.decommit_rows_local / decommit_rows_finish with the ENGINE as the backend, 3 virtual ranks on this GPU (host buffers
    and device tensors): the rank traces concatenate to the whole instance's trace; every rank ends with the whole closed form +
    commitment"""
    import torch
    from era_zkevm_circuits_b200 import sharding
    n, limit = 5000, 5100
    u, s = synthetic.decommit_requests_trace(n, seed=12, n_hashes=40)
    io, up, sp = instance(orc, u, s)
    want = O.sort_decommittments_entry_point(orc, io, u, s, limit)
    assert want[0] == abi.ZKC_OK
    world = 3
    cum = np.concatenate([[0], np.cumsum(want[2][K["ADD_TO_QUEUE"]])]).astype(np.int64)
    offs = [int(cum[sharding.row_range(n, r, world)[0]]) for r in range(world)]

    def run(io_, u_, up_, s_, sp_, states_, lim, want_trace):
        return entry_point(engine, Witness(io_, u_, up_, s_, sp_, states_), lim, want_trace=want_trace, raise_on_unsatisfied=False)

    commit = lambda e: engine.commit_encoding(np.ascontiguousarray(e, dtype=np.uint64).reshape(1, -1))[0]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)).cuda()
    i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
    for on_dev in (False, True):
        w = Witness(io, dev(u), i64(up), dev(s), i64(sp), i64(want[5])) if on_dev else Witness(io, u, up, s, sp, want[5])
        locs = [sharding.decommit_rows_local(run, w, limit, r, world, offs) for r in range(world)]
        recs = np.stack([l[3] for l in locs])
        traces = []
        for r in range(world):
            com, io_g, trace, st = sharding.decommit_rows_finish(locs[r][0], r, world, recs, io, offs, engine.scale_accumulators, commit)
            assert st.code == 0, (r, st.code, hex(st.failed_checks), st.first_bad_row)
            assert com.tolist() == want[3].tolist()
            assert bytes(io_g.hidden_fsm_output) == bytes(want[1].hidden_fsm_output) and bytes(io_g.final_queue_state) == bytes(want[1].final_queue_state)
            traces.append(trace.cpu().numpy().view(np.uint64) if on_dev else trace)
        bad = np.argwhere(np.concatenate(traces, axis=1) != want[2])
        assert bad.size == 0, f"first differing (col,row): {bad[:8].tolist()}"
