"""main_vm: CUDA path (one thread per cycle from host-supplied snapshots) through the C ABI vs the sequential CPU
oracle: trace, every per-cycle state, FSM output, commitment and status bit-exact; plus the GPU out-of-circuit run
against the oracle's."""
import ctypes as C

import numpy as np
import pytest

import orc as O
from era_zkevm_circuits_b200 import (VmCircuitWitness, abi, isa as I, main_vm_entry_point, main_vm_initial_state,
                                     main_vm_simulate)

pytestmark = pytest.mark.gpu
K = abi.VM_COLS


def fresh(orc):
    isa = I.Isa()
    io = abi.VmClosedForm()
    io.start_flag = 1
    io.rollback_queue_tail_for_block[0] = 12345
    io.memory_queue_initial_tail[3] = 777
    io.memory_queue_initial_length = 5
    return isa, io, O.vm_initial_state(orc, io, isa.isa)


def assert_same(want, got, check_trace=True):
    rc, io, trace, com, st = want
    assert got.status.code == rc, (got.status.code, hex(got.status.failed_checks), got.status.first_bad_row)
    assert got.status.failed_checks == st.failed_checks and got.status.first_bad_row == st.first_bad_row
    assert got.closed_form_input.completion_flag == io.completion_flag
    a, b = np.zeros(243, dtype=np.uint64), np.zeros(243, dtype=np.uint64)
    lib = O.load()
    lib.orc_vm_flatten_state(C.byref(got.closed_form_input.hidden_fsm_output), O.p(a))
    lib.orc_vm_flatten_state(C.byref(io.hidden_fsm_output), O.p(b))
    assert np.array_equal(a, b)
    assert got.commitment.tolist() == com.tolist()
    if check_trace:
        bad = np.argwhere(got.trace != trace)
        assert bad.size == 0, f"first differing (col,row): {bad[:8].tolist()}"


def test_initial_state_matches_oracle(engine, orc):
    isa, io, st = fresh(orc)
    got = main_vm_initial_state(engine, io, isa.isa)
    assert bytes(got) == bytes(st)


@pytest.mark.parametrize("n_ops,cycles,seed", [(8, 1, 1), (64, 127, 2), (256, 700, 3), (1024, 5000, 4), (4096, 20000, 5)])
def test_random_programs_bit_exact(engine, orc, n_ops, cycles, seed):
    isa, io, st = fresh(orc)
    ops = I.random_program(isa, n_ops, seed=seed)
    rc, snaps, wit, status = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles)
    assert rc == 0, (hex(status.failed_checks), status.first_bad_row)
    want = O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles)
    assert want[0] == 0
    got = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, snaps, wit), cycles)
    assert_same(want, got)
    # the GPU out-of-circuit run reproduces the oracle's snapshots and witness
    d_snaps, d_wit, st2 = main_vm_simulate(engine, isa.isa, [st], I.pack_code(ops)[None], cycles)
    assert st2.code == 0
    hs, hw = d_snaps.cpu().numpy()[0], d_wit.cpu().numpy()[0]
    lib = O.load()
    for i in (0, 1, cycles // 2, cycles):
        a, b = np.zeros(243, dtype=np.uint64), np.zeros(243, dtype=np.uint64)
        lib.orc_vm_flatten_state(O.p(np.ascontiguousarray(hs[i])), O.p(a)); lib.orc_vm_flatten_state(O.p(np.ascontiguousarray(snaps[i])), O.p(b))
        assert np.array_equal(a, b), i
    assert np.array_equal(hw[:, :68], wit[:, :68])
    # device-resident inputs straight from the simulator
    got2 = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, d_snaps[0], d_wit[0]), cycles)
    assert got2.commitment.tolist() == want[3].tolist() and got2.status.code == 0
    assert np.array_equal(got2.trace.cpu().numpy().view(np.uint64), want[2])


def test_chained_instances_and_expected_output(engine, orc):
    isa, io, st = fresh(orc)
    ops = I.random_program(isa, 512, seed=9)
    cycles, cut = 3000, 1234
    rc, snaps, wit, _ = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles)
    whole = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, snaps, wit), cycles)
    a = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, snaps[:cut + 1], wit[:cut]), cut)
    nxt = abi.VmClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    want = O.vm_entry_point(orc, nxt, isa.isa, snaps[cut:], wit[cut:], cycles - cut)
    b = main_vm_entry_point(engine, VmCircuitWitness(nxt, isa.isa, snaps[cut:], wit[cut:]), cycles - cut)
    assert_same(want, b)
    assert np.array_equal(np.concatenate([a.trace, b.trace], axis=1), whole.trace)
    exp = abi.VmClosedForm.from_buffer_copy(bytes(nxt))
    exp.hidden_fsm_output = b.closed_form_input.hidden_fsm_output
    ok = main_vm_entry_point(engine, VmCircuitWitness(exp, isa.isa, snaps[cut:], wit[cut:]), cycles - cut, compare_expected=True)
    assert ok.status.code == 0
    exp.hidden_fsm_output.flags[1] ^= 1
    bad = main_vm_entry_point(engine, VmCircuitWitness(exp, isa.isa, snaps[cut:], wit[cut:]), cycles - cut, compare_expected=True,
                              raise_on_unsatisfied=False)
    assert bad.status.code == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH


def test_error_cases_match_oracle(engine, orc):
    isa, io, st = fresh(orc)
    ops = I.random_program(isa, 256, seed=13)
    cycles = 900
    rc, snaps, wit, _ = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles)
    # corrupted snapshot: a register limb, then a queue state element
    for idx, byte in ((123, 40), (500, 1100), (0, 36), (cycles, 44)):
        bad = snaps.copy(); bad[idx, byte] ^= 1
        want = O.vm_entry_point(orc, io, isa.isa, bad, wit, cycles)
        got = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, bad, wit), cycles, raise_on_unsatisfied=False)
        assert want[0] == abi.ZKC_ERR_SNAPSHOT_MISMATCH
        assert_same(want, got, check_trace=False)
    # corrupted oracle answer (a memory read value): the next state no longer matches
    r = int(np.flatnonzero(O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles)[2][K["SHOULD_READ_SRC0"]])[7])
    w2 = wit.copy(); w2[r, 36] ^= 1
    want = O.vm_entry_point(orc, io, isa.isa, snaps, w2, cycles)
    got = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, snaps, w2), cycles, raise_on_unsatisfied=False)
    assert want[0] == abi.ZKC_ERR_SNAPSHOT_MISMATCH and want[4].first_bad_row == r + 1
    assert_same(want, got, check_trace=False)
    # a program that raises an exception (ptr.add on a non-pointer): the panic cycle is reported as unsupported
    ops2 = [isa.encode(I.OP_ADD, 0, 0, src0=2, src1=3, dst0=4)] * 5 + [isa.encode(I.OP_PTR, 0, 1, src=I.MODE_IMM16, src1=2, dst0=5, imm0=1)]
    ops2 += [isa.encode(I.OP_NOP)] * 4
    rc, s2, w3, status = O.vm_run(orc, isa.isa, st, I.pack_code(ops2), 8)
    assert rc == abi.ZKC_ERR_UNSUPPORTED and status.first_bad_row == 6
    want = O.vm_entry_point(orc, io, isa.isa, s2, w3, 8)
    got = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, s2, w3), 8, raise_on_unsatisfied=False)
    assert want[0] == abi.ZKC_ERR_UNSUPPORTED
    assert_same(want, got, check_trace=False)


def test_batch_of_instances(engine, orc):
    from era_zkevm_circuits_b200 import main_vm_entry_point_batch
    n, cycles = 6, 500
    isa = I.Isa()
    ios, states, codes = [], [], []
    for i in range(n):
        io = abi.VmClosedForm(); io.start_flag = 1; io.rollback_queue_tail_for_block[1] = 100 + i
        ios.append(io); states.append(O.vm_initial_state(orc, io, isa.isa))
        codes.append(I.pack_code(I.random_program(isa, 128, seed=50 + i)))
    d_snaps, d_wit, st = main_vm_simulate(engine, isa.isa, states, np.stack(codes), cycles)
    assert st.code == 0
    import torch
    trace = torch.empty((n, K["NUM_COLS"], cycles), dtype=torch.int64, device="cuda")
    coms, out, statuses, rc = main_vm_entry_point_batch(engine, ios, isa.isa, d_snaps, d_wit, cycles, trace_out=trace)
    assert rc == 0
    hs, hw, ht = d_snaps.cpu().numpy(), d_wit.cpu().numpy(), trace.cpu().numpy().view(np.uint64)
    for i in range(n):
        want = O.vm_entry_point(orc, ios[i], isa.isa, hs[i], hw[i], cycles)
        assert want[0] == 0 and coms[i].tolist() == want[3].tolist()
        assert np.array_equal(ht[i], want[2])
    assert len({tuple(c) for c in coms.tolist()}) == n
    # one bad instance does not disturb the others
    hs2 = hs.copy(); hs2[3, 77, 40] ^= 1
    coms2, out2, statuses2, rc2 = main_vm_entry_point_batch(engine, ios, isa.isa, np.ascontiguousarray(hs2), np.ascontiguousarray(hw), cycles)
    assert rc2 == abi.ZKC_ERR_SNAPSHOT_MISMATCH and statuses2[3].first_bad_row == 77
    assert [s.code for s in statuses2] == [0, 0, 0, abi.ZKC_ERR_SNAPSHOT_MISMATCH, 0, 0]
    assert coms2[[0, 1, 2, 4, 5]].tolist() == coms[[0, 1, 2, 4, 5]].tolist()
