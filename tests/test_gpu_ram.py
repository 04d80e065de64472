"""ram_permutation: CUDA path through the C ABI vs the CPU oracle on the same inputs, bit-exact
(witness trace, FSM output, commitment, status).  Mirrors the reference's own test
(/root/reference/src/ram_permutation/mod.rs:418-557) and widens it."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import orc as O
import vectors as V
from era_zkevm_circuits_b200 import (RamPermutationCircuitInstanceWitness, abi, ram_permutation_entry_point,
                                     synthetic)

pytestmark = pytest.mark.gpu
K = abi.RAM_COLS


def run_both(engine, orc, io, u, up, s, sp, limit, **kw):
    want = O.ram_entry_point(orc, io, u, s, limit, **{k: v for k, v in kw.items() if k in ("compare_expected",)})
    w = RamPermutationCircuitInstanceWitness(io, u, up, s, sp)
    got = ram_permutation_entry_point(engine, w, limit, raise_on_unsatisfied=False, **kw)
    return want, got


def assert_same(want, got, check_trace=True):
    rc, io, trace, com, st = want
    assert got.status.code == rc
    assert got.status.failed_checks == st.failed_checks
    assert got.status.first_bad_row == st.first_bad_row
    assert got.closed_form_input.completion_flag == io.completion_flag
    assert H.fsm_equal(got.closed_form_input.hidden_fsm_output, io.hidden_fsm_output)
    assert got.commitment.tolist() == com.tolist()
    if check_trace:
        bad = np.argwhere(got.trace != trace)
        assert bad.size == 0, f"first differing (col,row): {bad[:5].tolist()}"


def test_reference_vector(engine, orc):
    u, s = V.ram_reference_vector()
    io, up, sp = H.ram_instance(orc, u, s, 1)
    want, got = run_both(engine, orc, io, u, up, s, sp, 16)
    assert want[0] == abi.ZKC_OK
    assert_same(want, got)


@pytest.mark.parametrize("n,limit", [(1, 1), (255, 255), (256, 256), (257, 300), (1000, 1000), (5000, 8192), (1 << 16, 1 << 16)])
def test_synthetic_bit_exact(engine, orc, n, limit):
    u, s = synthetic.ram_trace(n, seed=n, n_cells=max(1, n // 64), n_nondet=min(3, n // 2))
    io, up, sp = H.ram_instance(orc, u, s, min(3, n // 2))
    want, got = run_both(engine, orc, io, u, up, s, sp, limit)
    assert want[0] == abi.ZKC_OK, hex(want[4].failed_checks)
    assert_same(want, got)


def test_empty_queue_and_zero_limit(engine, orc):
    u, s = synthetic.ram_trace(0)
    io, up, sp = H.ram_instance(orc, u, s, 0)
    want, got = run_both(engine, orc, io, u, up, s, sp, 8)
    assert want[0] == abi.ZKC_OK
    assert_same(want, got)
    u, s = synthetic.ram_trace(10, n_cells=2)
    io, up, sp = H.ram_instance(orc, u, s, 0)
    want, got = run_both(engine, orc, io, u, up, s, sp, 0, want_trace=False)
    assert_same(want, got, check_trace=False)


def test_chained_instances(engine, orc):
    u, s = synthetic.ram_trace(3000, seed=9, n_cells=40, n_nondet=2)
    io, up, sp = H.ram_instance(orc, u, s, 2)
    whole = ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io, u, up, s, sp), 3072)
    a = ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io, u, up, s, sp), 1000)
    assert a.closed_form_input.completion_flag == 0
    nxt = H.continue_io(a.closed_form_input)
    want, got = run_both(engine, orc, nxt, u[1000:], up[1000:], s[1000:], sp[1000:], 2072)
    assert_same(want, got)
    assert H.fsm_equal(got.closed_form_input.hidden_fsm_output, whole.closed_form_input.hidden_fsm_output)
    assert np.array_equal(np.concatenate([a.trace, got.trace], axis=1), whole.trace)
    # hook_compare_witness: expected output supplied by the host
    exp = abi.RamClosedForm.from_buffer_copy(bytes(nxt))
    exp.hidden_fsm_output = got.closed_form_input.hidden_fsm_output
    exp.completion_flag = 1
    ok = ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(exp, u[1000:], up[1000:], s[1000:], sp[1000:]),
                                     2072, compare_expected=True)
    assert ok.status.code == 0
    exp.hidden_fsm_output.previous_value[3] ^= 1
    bad = ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(exp, u[1000:], up[1000:], s[1000:], sp[1000:]),
                                      2072, compare_expected=True, raise_on_unsatisfied=False)
    assert bad.status.code == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH


def test_negative_cases_match_oracle(engine, orc):
    u, s = synthetic.ram_trace(2000, seed=11, n_cells=30, n_nondet=1)
    cases = []
    s2 = s.copy(); s2[[700, 701]] = s2[[701, 700]]; cases.append((u, s2, 1))          # order violated
    w = int(np.flatnonzero(s["rw_flag"] == 0)[5])
    s3 = s.copy(); s3["value"][w][2] ^= 4; cases.append((u, s3, 1))                   # read != last write
    s4 = s.copy(); s4["is_ptr"][w] ^= 1; cases.append((u, s4, 1))                     # pointer flag differs
    cases.append((u, s, 0))                                                           # snapshot length wrong
    u5 = u.copy(); u5["value"][100][0] ^= 1; cases.append((u5, s, 1))                 # not a permutation
    for uu, ss, nd in cases:
        io, up, sp = H.ram_instance(orc, uu, ss, nd)
        want, got = run_both(engine, orc, io, uu, up, ss, sp, 2048)
        assert want[0] == abi.ZKC_ERR_UNSATISFIED
        assert_same(want, got)
    with pytest.raises(Exception):
        io, up, sp = H.ram_instance(orc, u, s2, 1)
        ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io, u, up, s2, sp), 2048)


def test_bad_queue_witness_and_arguments(engine, orc):
    u, s = synthetic.ram_trace(600, seed=12, n_cells=9)
    io, up, sp = H.ram_instance(orc, u, s, 0)
    up2 = up.copy(); up2[300, 5] ^= 1
    r = ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io, u, up2, s, sp), 600, raise_on_unsatisfied=False)
    assert r.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT and r.status.first_bad_row in (299, 300)
    # fewer witness records than the queue length says: the reference would panic on the empty deque
    with pytest.raises(Exception):
        ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io, u[:100], up[:100], s[:100], sp[:100]), 600)


def test_device_resident_inputs_and_missing_prev_states(engine, orc):
    import torch
    n = 4096
    u, s = synthetic.ram_trace(n, seed=13, n_nondet=4)
    io, up, sp = H.ram_instance(orc, u, s, 4)
    want = O.ram_entry_point(orc, io, u, s, n)
    tod = lambda a: torch.from_numpy(a.view(np.uint8).reshape(len(a), -1)).cuda()
    du, ds = tod(u), tod(s)
    dup, dsp = torch.from_numpy(up.view(np.int64)).cuda(), torch.from_numpy(sp.view(np.int64)).cuda()
    got = ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io, du, dup, ds, dsp), n)
    torch.cuda.synchronize()
    assert got.commitment.tolist() == want[3].tolist()
    assert np.array_equal(got.trace.cpu().numpy().view(np.uint64), want[2])
    # no previous-state column: the head chain is rebuilt on the device
    got = ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io, u[:512], None, s[:512], None), 256)
    want = O.ram_entry_point(orc, io, u, s, 256)
    assert got.commitment.tolist() == want[3].tolist() and np.array_equal(got.trace, want[2])
    # queue simulation on the device == oracle
    prev, fin = engine.memory_queue_simulate(u)
    assert np.array_equal(prev, up) and bytes(fin[0]) == bytes(io.observable_input.unsorted_queue_initial_state)


@pytest.mark.parametrize("limit", [3100, 3101])  # even: row pairs per thread (128-bit loads); odd: one row per thread
def test_check_trace_accepts_valid_and_localises_corruption(engine, orc, limit):
    from era_zkevm_circuits_b200 import ram_permutation_check_trace
    n = 3000
    u, s = synthetic.ram_trace(n, seed=21, n_cells=50, n_nondet=2)
    io, up, sp = H.ram_instance(orc, u, s, 2)
    _, _, trace, _, _ = O.ram_entry_point(orc, io, u, s, limit)  # the ORACLE's trace satisfies the CUDA checker
    for gates in (0, abi.GATES_GENERAL):
        viol, st = ram_permutation_check_trace(engine, io, trace, limit, gates)
        assert viol == 0 and st.code == 0, (gates, hex(st.failed_checks), st.first_bad_row)
    probes = [(K["GP_CHAIN"] + 13, 1234, abi.RAMV["GP_CHAIN"], 0), (K["GP_ACC"] + 1, 77, abi.RAMV["GP_ACC"], 0),
              (K["SORTED_ENC"] + 3, 5, abi.RAMV["ENCODING"], 0), (K["CMP_DIFF"] + 1, 2999, abi.RAMV["COMPARISON"], 0),
              (K["UNSORTED_HEAD"] + 2, 100, abi.RAMV["ROUND_FUNCTION"], 0), (K["NUM_NONDET_WRITES"], 10, abi.RAMV["NONDET"], 0),
              (K["SAME_CELL"], 600, abi.RAMV["FLAGS"], 0), (K["CAN_POP"], 3050, abi.RAMV["BOOLEAN"], 0),
              (K["SORTED_LEN"], 40, abi.RAMV["QUEUE_LEN"], 0)]
    for col, row, bit, gates in probes:
        t = trace.copy()
        t[col, row] ^= 1
        viol, st = ram_permutation_check_trace(engine, io, t, limit, gates)
        assert viol >= 1 and st.code == abi.ZKC_ERR_UNSATISFIED
        assert st.first_bad_row == row and st.failed_checks & bit, (col, row, hex(st.failed_checks), st.first_bad_row)
    # a head corruption is invisible to the streaming pass only when the row pops (needs the round function)
    t = trace.copy(); t[K["UNSORTED_HEAD"] + 2, 100] ^= 1
    viol, st = ram_permutation_check_trace(engine, io, t, limit, abi.GATES_GENERAL)
    assert viol == 0
    # the gadget cells: byte decompositions, differences, inverse witnesses (where the inverted value is not zero), limb flags
    for name in ("UNSORTED_ENC_BYTES", "SORTED_ENC_BYTES", "UNSORTED_LEN_INV", "SORTED_LEN_INV", "TS_INV", "PAGE_DIFF", "PAGE_DIFF_INV", "CMP_DIFF_INV",
                 "CELL_DIFF", "CELL_DIFF_INV", "CELL_LIMB_EQ", "VALUE_DIFF", "VALUE_DIFF_INV", "VALUE_LIMB_EQ", "VALUE_ZERO_DIFF", "VALUE_ZERO_DIFF_INV",
                 "VALUE_ZERO_LIMB_EQ", "PTR_DIFF", "PTR_DIFF_INV"):
        col = K[name] + (1 if name in ("CMP_DIFF_INV", "CELL_DIFF", "CELL_DIFF_INV", "CELL_LIMB_EQ") else 0)
        rows = np.flatnonzero(trace[col, 1:n] != 0) + 1
        if name.endswith("_EQ") or not len(rows):
            rows = np.arange(1, n)
        row = int(rows[len(rows) // 2])
        t = trace.copy()
        t[col, row] ^= 1
        viol, st = ram_permutation_check_trace(engine, io, t, limit, abi.GATES_GENERAL)
        assert viol >= 1 and st.first_bad_row == row and st.failed_checks & abi.RAMV["GADGET_CELLS"], (name, row, hex(st.failed_checks), st.first_bad_row)
    # an unsatisfiable input (order violated) is caught as an enforcement failure of its own trace
    s2 = s.copy(); s2[[700, 701]] = s2[[701, 700]]
    io2, up2, sp2 = H.ram_instance(orc, u, s2, 2)
    r = ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io2, u, up2, s2, sp2), limit, raise_on_unsatisfied=False)
    viol, st = ram_permutation_check_trace(engine, io2, r.trace, limit, abi.GATES_GENERAL)
    assert viol >= 1 and st.failed_checks == abi.RAMV["ENFORCE"] and st.first_bad_row == r.status.first_bad_row


def test_full_size_properties(engine):
    """2^20 rows (BASELINE trace capacity): size-independent properties instead of an oracle run --
    lhs == rhs for a true permutation, completion, self-consistency of the device trace under the
    device constraint evaluator, and chained halves == whole."""
    import torch
    from era_zkevm_circuits_b200 import ram_permutation_check_trace
    n = 1 << 20
    u, s = synthetic.ram_trace(n, seed=0xC1, n_cells=1 << 10, n_nondet=7)
    tod = lambda a: torch.from_numpy(a.view(np.uint8).reshape(len(a), -1)).cuda()
    du, ds = tod(u), tod(s)
    up, ufin = engine.memory_queue_simulate(du)
    sp, sfin = engine.memory_queue_simulate(ds)
    io = O.ram_closed_form(ufin[0], sfin[0], True, 7)
    r = ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io, du, up, ds, sp), n)
    out = r.closed_form_input.hidden_fsm_output
    assert r.status.code == 0 and r.closed_form_input.completion_flag == 1
    assert list(out.lhs_accumulator) == list(out.rhs_accumulator) and out.num_nondeterministic_writes == 7
    viol, st = ram_permutation_check_trace(engine, io, r.trace, n, 0)
    assert viol == 0
    half = n // 2
    a = ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io, du, up, ds, sp), half, want_trace=False)
    b = ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(
        H.continue_io(a.closed_form_input), du[half:], up[half:], ds[half:], sp[half:]), half, want_trace=False)
    assert H.fsm_equal(b.closed_form_input.hidden_fsm_output, out)
    assert b.commitment.tolist() != r.commitment.tolist()  # different closed forms commit differently


def test_one_instance_cut_by_rows_over_ranks(engine, orc):
    """sharding.ram_rows_local / ram_rows_finish with the ENGINE as the backend, 3 virtual ranks on this GPU (host buffers and device
    tensors): accumulator columns scaled and the non-deterministic-write counter offset after the exchange; the rank traces
    concatenate to the whole instance's trace, every rank ends with the whole closed form + commitment"""
    import torch
    from era_zkevm_circuits_b200 import sharding
    n, limit = 6000, 6100
    u, s = synthetic.ram_trace(n, seed=33, n_cells=80, n_nondet=3500)
    io, up, sp = H.ram_instance(orc, u, s, 3500)
    want = O.ram_entry_point(orc, io, u, s, limit)
    assert want[0] == abi.ZKC_OK
    world = 3

    def run(io_, u_, up_, s_, sp_, lim, want_trace):
        return ram_permutation_entry_point(engine, RamPermutationCircuitInstanceWitness(io_, u_, up_, s_, sp_), lim, want_trace=want_trace,
                                           raise_on_unsatisfied=False)

    commit = lambda e: engine.commit_encoding(np.ascontiguousarray(e, dtype=np.uint64).reshape(1, -1))[0]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)).cuda()
    i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
    for on_dev in (False, True):
        w = RamPermutationCircuitInstanceWitness(io, dev(u), i64(up), dev(s), i64(sp)) if on_dev else RamPermutationCircuitInstanceWitness(io, u, up, s, sp)
        locs = [sharding.ram_rows_local(run, w, limit, r, world) for r in range(world)]
        recs = np.stack([l[3] for l in locs])
        assert int(recs[:, 11].sum()) == 3500 and np.count_nonzero(recs[:, 11]) >= 2
        traces = []
        for r in range(world):
            com, io_g, trace, st = sharding.ram_rows_finish(locs[r][0], r, world, recs, io, engine.scale_accumulators, commit)
            assert st.code == 0, (r, st.code, hex(st.failed_checks), st.first_bad_row)
            assert com.tolist() == want[3].tolist() and bytes(io_g.hidden_fsm_output) == bytes(want[1].hidden_fsm_output)
            traces.append(trace.cpu().numpy().view(np.uint64) if on_dev else trace)
        bad = np.argwhere(np.concatenate(traces, axis=1) != want[2])
        assert bad.size == 0, f"first differing (col,row): {bad[:8].tolist()}"
