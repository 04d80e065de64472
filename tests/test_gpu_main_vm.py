"""main_vm: CUDA path (one thread per cycle from host-supplied snapshots, sponges as dense job launches) through the C
ABI vs the sequential CPU oracle: trace, every per-cycle state, FSM output, commitment and status bit-exact; plus the GPU
out-of-circuit run (two passes: rollback-queue resolution, then recording) against the oracle's."""
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


def with_tail(io, tail):
    io2 = abi.VmClosedForm.from_buffer_copy(bytes(io))
    for i in range(4):
        io2.rollback_queue_tail_for_block[i] = int(tail[i])
    return io2


def flat(lib, state_bytes):
    a = np.zeros(243, dtype=np.uint64)
    buf = np.ascontiguousarray(np.frombuffer(bytes(state_bytes), dtype=np.uint8))
    lib.orc_vm_flatten_state(O.p(buf), O.p(a))
    return a


def assert_same(want, got, check_trace=True):
    rc, io, trace, com, st = want
    assert got.status.code == rc, (got.status.code, hex(got.status.failed_checks), got.status.first_bad_row, rc, hex(st.failed_checks), st.first_bad_row)
    assert got.status.failed_checks == st.failed_checks and got.status.first_bad_row == st.first_bad_row
    assert got.closed_form_input.completion_flag == io.completion_flag
    lib = O.load()
    assert np.array_equal(flat(lib, got.closed_form_input.hidden_fsm_output), flat(lib, io.hidden_fsm_output))
    assert got.commitment.tolist() == com.tolist()
    if check_trace:
        bad = np.argwhere(got.trace != trace)
        assert bad.size == 0, f"first differing (col,row): {bad[:8].tolist()}"


def test_initial_state_matches_oracle(engine, orc):
    isa, io, st = fresh(orc)
    got = main_vm_initial_state(engine, io, isa.isa)
    assert bytes(got) == bytes(st)


@pytest.mark.parametrize("n_ops,cycles,seed,full", [(8, 1, 1, False), (64, 127, 2, False), (256, 700, 3, True), (1024, 5000, 4, True),
                                                    (4096, 20000, 5, True), (512, 3000, 6, True)])
def test_random_programs_bit_exact(engine, orc, n_ops, cycles, seed, full):
    isa, io, st = fresh(orc)
    ops = I.random_program(isa, n_ops, seed=seed, full=full)
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
    assert rc == 0, (hex(status.failed_checks), status.first_bad_row)
    io = with_tail(io, tail)
    want = O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles, cw=cw)
    assert want[0] == 0, (want[0], hex(want[4].failed_checks), want[4].first_bad_row)
    got = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, snaps, wit, cw), cycles)
    assert_same(want, got)
    if full:
        props = got.trace[K["PROPS"]]
        for op in (I.OP_UMA, I.OP_LOG, I.OP_NEAR_CALL, I.OP_RET):
            assert ((props >> np.uint64(op)) & np.uint64(1)).sum() > 0, op
    # the GPU out-of-circuit run reproduces the oracle's snapshots, witness, popped frames and resolved rollback tail
    sim = main_vm_simulate(engine, isa.isa, [st], I.pack_code(ops)[None], cycles)
    assert sim.status.code == 0, (sim.status.code, hex(sim.status.failed_checks), sim.status.first_bad_row)
    hs, hw = sim.snapshots.cpu().numpy()[0], sim.witness.cpu().numpy()[0]
    lib = O.load()
    for i in (0, 1, cycles // 2, cycles):
        assert np.array_equal(flat(lib, hs[i].tobytes()), flat(lib, snaps[i].tobytes())), i
    assert np.array_equal(hw, wit)
    assert sim.rollback_tails[0].tolist() == tail.tolist() and int(sim.n_callstack[0]) == len(cw)
    assert np.array_equal(sim.callstack_witness.cpu().numpy()[0, :len(cw)], cw)
    # device-resident inputs straight from the simulator
    got2 = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, sim.snapshots[0], sim.witness[0], sim.callstack_witness[0]), cycles)
    assert got2.commitment.tolist() == want[3].tolist() and got2.status.code == 0
    assert np.array_equal(got2.trace.cpu().numpy().view(np.uint64), want[2])


def test_hand_written_calls_logs_uma(engine, orc):
    """the programs of tests/test_oracle_main_vm_ops.py (every branch of near_call / ret / log / uma) on the GPU"""
    import test_oracle_main_vm_ops as T
    isa, io, st = T.fresh(orc, tail=4242)
    A, B, Cv = 0x1111 << 200 | 5, 0x2222 << 100 | 6, 0x3333
    T.set_reg(st, 2, 7); T.set_reg(st, 3, A); T.set_reg(st, 4, B); T.set_reg(st, 5, Cv); T.set_reg(st, 6, 9)
    T.set_reg(st, 7, 70); T.set_reg(st, 8, (3 | (10 << 32) | (64 << 64) | (10 << 96)), is_ptr=1)
    ops = [
        isa.encode(I.OP_LOG, I.LOG_STORAGE_WRITE, src0=2, src1=3),
        isa.encode(I.OP_LOG, I.LOG_STORAGE_READ, src0=2, dst0=10),
        isa.encode(I.OP_NEAR_CALL, src0=0, imm0=8, imm1=12),
        isa.encode(I.OP_LOG, I.LOG_STORAGE_READ, src0=2, dst0=11),
        isa.encode(I.OP_NEAR_CALL, src0=0, imm0=14, imm1=12),
        isa.encode(I.OP_LOG, I.LOG_STORAGE_READ, src0=6, dst0=12),
        isa.encode(I.OP_LOG, I.LOG_EVENT, 1, src0=2, src1=5),
        isa.encode(I.OP_JUMP, 0, 0, src=I.MODE_IMM16, imm0=20),
        isa.encode(I.OP_LOG, I.LOG_STORAGE_WRITE, src0=2, src1=4),
        isa.encode(I.OP_UMA, I.UMA_HEAP_WRITE, 1, src0=7, src1=3, dst0=9),
        isa.encode(I.OP_LOG, I.LOG_EVENT, 0, src0=3, src1=4),
        isa.encode(I.OP_RET, I.RET_REVERT),
        isa.encode(I.OP_JUMP, 0, 0, src=I.MODE_IMM16, imm0=3),
        isa.encode(I.OP_NOP),
        isa.encode(I.OP_LOG, I.LOG_STORAGE_WRITE, src0=6, src1=5),
        isa.encode(I.OP_RET, I.RET_OK),
    ] + [isa.encode(I.OP_NOP)] * 4 + [
        isa.encode(I.OP_LOG, I.LOG_PRECOMPILE, src0=2, src1=6, dst0=14),
        isa.encode(I.OP_UMA, I.UMA_HEAP_READ, 1, src0=7, dst0=13, dst1=7),
        isa.encode(I.OP_UMA, I.UMA_PTR_READ, 0, src0=8, dst0=15),
        isa.encode(I.OP_UMA, I.UMA_AUX_WRITE, 0, src0=2, src1=4),
        isa.encode(I.OP_PTR, 0, 1, src=I.MODE_IMM16, src1=2, dst0=11, imm0=1),  # exception -> the root frame panics
        isa.encode(I.OP_NOP),
    ]
    cycles = 26
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
    assert rc == 0
    io2 = with_tail(io, tail); io2.start_flag = 0; io2.hidden_fsm_input = O.vm_state_at(snaps, 0)
    want = O.vm_entry_point(orc, io2, isa.isa, snaps, wit, cycles, cw=cw)
    assert want[0] == abi.ZKC_ERR_UNSATISFIED and want[4].failed_checks == abi.VM_CHK["BOOTLOADER_EXIT"] and want[1].completion_flag == 1
    got = main_vm_entry_point(engine, VmCircuitWitness(io2, isa.isa, snaps, wit, cw), cycles, raise_on_unsatisfied=False)
    assert_same(want, got)
    # the same on the GPU simulator
    sim = main_vm_simulate(engine, isa.isa, [st], I.pack_code(ops)[None], cycles)
    assert sim.status.code == 0 and np.array_equal(sim.witness.cpu().numpy()[0], wit)
    assert sim.rollback_tails[0].tolist() == tail.tolist()
    assert np.array_equal(flat(O.load(), sim.snapshots.cpu().numpy()[0, cycles].tobytes()), flat(O.load(), snaps[cycles].tobytes()))
    # circuit enforcement failures are reported like the oracle reports them: a wrong claimed rollback head, a wrong popped frame
    for what, (row, byte) in {"rollback": (0, 144), "value": (1, 80)}.items():
        w2 = wit.copy(); w2[row, byte] ^= 1
        want = O.vm_entry_point(orc, io2, isa.isa, snaps, w2, cycles, cw=cw)
        got = main_vm_entry_point(engine, VmCircuitWitness(io2, isa.isa, snaps, w2, cw), cycles, raise_on_unsatisfied=False)
        assert want[0] != 0
        assert_same(want, got, check_trace=False)
    for byte in (0, 100, 250, 300):  # context fields / previous sponge state of the first popped frame
        cw2 = cw.copy(); cw2[0, byte] ^= 1
        want = O.vm_entry_point(orc, io2, isa.isa, snaps, wit, cycles, cw=cw2)
        got = main_vm_entry_point(engine, VmCircuitWitness(io2, isa.isa, snaps, wit, cw2), cycles, raise_on_unsatisfied=False)
        assert want[0] != 0, byte
        assert_same(want, got, check_trace=False)
    # a callstack index out of range
    ret_row = int(np.flatnonzero(want[2][K["OP_AUX"] + 43])[0]) if want[2] is not None else 0
    w2 = wit.copy(); w2[ret_row, 68:72] = np.frombuffer(np.uint32(1000).tobytes(), dtype=np.uint8)
    want = O.vm_entry_point(orc, io2, isa.isa, snaps, w2, cycles, cw=cw)
    got = main_vm_entry_point(engine, VmCircuitWitness(io2, isa.isa, snaps, w2, cw), cycles, raise_on_unsatisfied=False)
    assert want[4].failed_checks & abi.VM_CHK["CALLSTACK"]
    assert_same(want, got, check_trace=False)
    # no popped-frame table at all: every ret is reported, nothing is dereferenced
    want = O.vm_entry_point(orc, io2, isa.isa, snaps, wit, cycles, cw=None)
    got = main_vm_entry_point(engine, VmCircuitWitness(io2, isa.isa, snaps, wit, None), cycles, raise_on_unsatisfied=False)
    assert want[4].failed_checks & abi.VM_CHK["CALLSTACK"]
    assert_same(want, got, check_trace=False)


def test_chained_instances_and_expected_output(engine, orc):
    isa, io, st = fresh(orc)
    ops = I.random_program(isa, 512, seed=9)
    cycles, cut = 3000, 1234
    rc, snaps, wit, _, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
    io = with_tail(io, tail)
    whole = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, snaps, wit, cw), cycles)
    a = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, snaps[:cut + 1], wit[:cut], cw), cut)
    nxt = abi.VmClosedForm.from_buffer_copy(bytes(a.closed_form_input)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.closed_form_input.hidden_fsm_output
    want = O.vm_entry_point(orc, nxt, isa.isa, snaps[cut:], wit[cut:], cycles - cut, cw=cw)
    b = main_vm_entry_point(engine, VmCircuitWitness(nxt, isa.isa, snaps[cut:], wit[cut:], cw), cycles - cut)
    assert_same(want, b)
    assert np.array_equal(np.concatenate([a.trace, b.trace], axis=1), whole.trace)
    exp = abi.VmClosedForm.from_buffer_copy(bytes(nxt))
    exp.hidden_fsm_output = b.closed_form_input.hidden_fsm_output
    ok = main_vm_entry_point(engine, VmCircuitWitness(exp, isa.isa, snaps[cut:], wit[cut:], cw), cycles - cut, compare_expected=True)
    assert ok.status.code == 0
    exp.hidden_fsm_output.flags[1] ^= 1
    bad = main_vm_entry_point(engine, VmCircuitWitness(exp, isa.isa, snaps[cut:], wit[cut:], cw), cycles - cut, compare_expected=True,
                              raise_on_unsatisfied=False)
    assert bad.status.code == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH


def test_error_cases_match_oracle(engine, orc):
    isa, io, st = fresh(orc)
    ops = I.random_program(isa, 256, seed=13)
    cycles = 900
    rc, snaps, wit, _, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
    io0 = io
    io = with_tail(io, tail)
    # corrupted snapshot: a register limb, a queue state element, context words, the stack sponge
    for idx, byte in ((123, 40), (500, 1100), (0, 36), (cycles, 44), (300, 700), (301, 760), (640, 900), (77, 1000), (10, 812)):
        bad = snaps.copy(); bad[idx, byte] ^= 1
        want = O.vm_entry_point(orc, io, isa.isa, bad, wit, cycles, cw=cw)
        got = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, bad, wit, cw), cycles, raise_on_unsatisfied=False)
        assert want[0] == abi.ZKC_ERR_SNAPSHOT_MISMATCH, (idx, byte)
        assert_same(want, got, check_trace=False)
    # corrupted oracle answer (a memory read value): the next state no longer matches
    r = int(np.flatnonzero(O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles, cw=cw)[2][K["SHOULD_READ_SRC0"]])[7])
    w2 = wit.copy(); w2[r, 36] ^= 1
    want = O.vm_entry_point(orc, io, isa.isa, snaps, w2, cycles, cw=cw)
    got = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, snaps, w2, cw), cycles, raise_on_unsatisfied=False)
    assert want[0] == abi.ZKC_ERR_SNAPSHOT_MISMATCH and want[4].first_bad_row == r + 1
    assert_same(want, got, check_trace=False)


@pytest.mark.parametrize("want_trace", [True, False])
def test_snapshot_link_covers_every_state_word(engine, orc, want_trace):
    """One flipped bit in each 32-bit word of a mid-trace snapshot, one word at a time: with a trace the link check reads what the
    cycle produced back from the trace columns (vm_link_kernel), without one the cycle's thread compares -- same verdicts as the oracle,
    which compares the whole record (main_vm/mod.rs:126-208: each cycle starts from the state the previous one left)."""
    isa, io, st = fresh(orc)
    ops = I.random_program(isa, 256, seed=29)
    cycles = 600
    rc, snaps, wit, _, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
    io = with_tail(io, tail)
    n_words = snaps.shape[1] // 4
    caught = 0
    for w in range(n_words):
        idx = 1 + (w * 37) % (cycles - 1)
        bad = snaps.copy(); bad[idx, 4 * w + (w % 4)] ^= 1 << (w % 8)
        want = O.vm_entry_point(orc, io, isa.isa, bad, wit, cycles, cw=cw)
        got = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, bad, wit, cw), cycles, want_trace=want_trace, raise_on_unsatisfied=False)
        assert got.status.code == want[0] and got.status.first_bad_row == want[4].first_bad_row, (w, idx, got.status.code, want[0])
        assert got.status.failed_checks == want[4].failed_checks, (w, idx)
        caught += want[0] == abi.ZKC_ERR_SNAPSHOT_MISMATCH
    assert caught >= n_words - 8  # only padding words may go unnoticed


@pytest.mark.parametrize("want_trace", [True, False])
def test_snapshot_link_after_calls_and_rets(engine, orc, want_trace):
    """the same, on the snapshots right after far calls, near calls and rets (the cycles that replace the whole context record)"""
    isa, io, st = fresh(orc)
    io.default_aa_code_hash[7] = (1 << 24) | 5; io.default_aa_code_hash[2] = 0xA1
    ops = I.random_program(isa, 512, seed=8, far_calls=True)
    cycles = 3000
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True, gc=io)
    assert rc == 0
    io = with_tail(io, tail)
    good = O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles, cw=cw)
    assert good[0] == 0
    moved = np.flatnonzero(good[2][K["DEPTH_OUT"]][:-1] != np.concatenate([[0], good[2][K["DEPTH_OUT"]][:-2]]))  # rows that push / pop a frame
    assert len(moved) > 20
    n_words = snaps.shape[1] // 4
    for j, w in enumerate(range(0, n_words, 3)):
        idx = int(moved[j % len(moved)]) + 1
        bad = snaps.copy(); bad[idx, 4 * w] ^= 0x10
        want = O.vm_entry_point(orc, io, isa.isa, bad, wit, cycles, cw=cw)
        got = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, bad, wit, cw), cycles, want_trace=want_trace, raise_on_unsatisfied=False)
        assert got.status.code == want[0] and got.status.first_bad_row == want[4].first_bad_row, (w, idx, got.status.code, want[0])
        assert got.status.failed_checks == want[4].failed_checks, (w, idx)


def test_batch_of_instances(engine, orc):
    from era_zkevm_circuits_b200 import main_vm_entry_point_batch
    n, cycles = 6, 500
    isa = I.Isa()
    ios, states, codes = [], [], []
    for i in range(n):
        io = abi.VmClosedForm(); io.start_flag = 1; io.rollback_queue_tail_for_block[1] = 100 + i
        ios.append(io); states.append(O.vm_initial_state(orc, io, isa.isa))
        codes.append(I.pack_code(I.random_program(isa, 128, seed=50 + i)))
    sim = main_vm_simulate(engine, isa.isa, states, np.stack(codes), cycles)
    assert sim.status.code == 0
    ios = [with_tail(io, t) for io, t in zip(ios, sim.rollback_tails)]
    import torch
    trace = torch.empty((n, K["NUM_COLS"], cycles), dtype=torch.int64, device="cuda")
    coms, out, statuses, rc = main_vm_entry_point_batch(engine, ios, isa.isa, sim.snapshots, sim.witness, cycles, trace_out=trace,
                                                        callstack_witness=sim.callstack_witness)
    assert rc == 0, [(s.code, hex(s.failed_checks), s.first_bad_row) for s in statuses]
    hs, hw, hc, ht = sim.snapshots.cpu().numpy(), sim.witness.cpu().numpy(), sim.callstack_witness.cpu().numpy(), trace.cpu().numpy().view(np.uint64)
    for i in range(n):
        want = O.vm_entry_point(orc, ios[i], isa.isa, hs[i], hw[i], cycles, cw=hc[i])
        assert want[0] == 0 and coms[i].tolist() == want[3].tolist()
        assert np.array_equal(ht[i], want[2])
    assert len({tuple(c) for c in coms.tolist()}) == n
    # one bad instance does not disturb the others
    hs2 = hs.copy(); hs2[3, 77, 40] ^= 1
    coms2, out2, statuses2, rc2 = main_vm_entry_point_batch(engine, ios, isa.isa, np.ascontiguousarray(hs2), np.ascontiguousarray(hw), cycles,
                                                            callstack_witness=np.ascontiguousarray(hc))
    assert rc2 == abi.ZKC_ERR_SNAPSHOT_MISMATCH and statuses2[3].first_bad_row == 77
    assert [s.code for s in statuses2] == [0, 0, 0, abi.ZKC_ERR_SNAPSHOT_MISMATCH, 0, 0]
    assert coms2[[0, 1, 2, 4, 5]].tolist() == coms[[0, 1, 2, 4, 5]].tolist()


def test_compact_trace_layout(engine, orc):
    """COMPACT layout (159 dense columns + one record per enforced sponge relation) carries exactly the dense trace"""
    from era_zkevm_circuits_b200 import main_vm_entry_point_batch
    import torch
    n, cycles = 3, 9000  # 9000 rows x 3 instances: the chunked host pipeline is taken
    isa = I.Isa()
    ios, states, codes = [], [], []
    for i in range(n):
        io = abi.VmClosedForm(); io.start_flag = 1; io.rollback_queue_tail_for_block[2] = 7 + i
        ios.append(io); states.append(O.vm_initial_state(orc, io, isa.isa))
        codes.append(I.pack_code(I.random_program(isa, 512, seed=90 + i)))
    sim = main_vm_simulate(engine, isa.isa, states, np.stack(codes), cycles)
    assert sim.status.code == 0
    ios = [with_tail(io, t) for io, t in zip(ios, sim.rollback_tails)]
    dense = torch.empty((n, K["NUM_COLS"], cycles), dtype=torch.int64, device="cuda")
    coms, _, _, rc = main_vm_entry_point_batch(engine, ios, isa.isa, sim.snapshots, sim.witness, cycles, trace_out=dense,
                                               callstack_witness=sim.callstack_witness)
    assert rc == 0
    want = dense.cpu().numpy().view(np.uint64)
    n_jobs = int(want[:, K["SPONGE_ENFORCE"]:K["SPONGE_ENFORCE"] + 9].sum())
    # device-resident compact output
    comp = torch.empty((n, abi.VM_COMPACT_COLS, cycles), dtype=torch.int64, device="cuda")
    rec = torch.zeros((n_jobs + 10, 104), dtype=torch.uint8, device="cuda")
    coms2, _, st2, rc = main_vm_entry_point_batch(engine, ios, isa.isa, sim.snapshots, sim.witness, cycles, trace_out=comp,
                                                  callstack_witness=sim.callstack_witness, sponge_records_out=rec)
    assert rc == 0 and coms2.tolist() == coms.tolist() and st2[0].reserved == n_jobs
    records = rec.cpu().numpy().view(abi.VM_SPONGE_RECORD_DTYPE).reshape(-1)[:n_jobs]
    got = abi.vm_expand_compact_trace(comp.cpu().numpy().view(np.uint64), records, cycles)
    assert np.array_equal(got, want)
    # host buffers (the chunked H2D | kernels | D2H pipeline), dense and compact
    hs, hw, hc = (np.ascontiguousarray(x.cpu().numpy()) for x in (sim.snapshots, sim.witness, sim.callstack_witness))
    hdense = np.zeros((n, K["NUM_COLS"], cycles), dtype=np.uint64)
    coms3, _, _, rc = main_vm_entry_point_batch(engine, ios, isa.isa, hs, hw, cycles, trace_out=hdense, callstack_witness=hc)
    assert rc == 0 and coms3.tolist() == coms.tolist() and np.array_equal(hdense, want)
    hcomp = np.zeros((n, abi.VM_COMPACT_COLS, cycles), dtype=np.uint64)
    hrec = np.zeros(n_jobs, dtype=abi.VM_SPONGE_RECORD_DTYPE)
    coms4, _, st4, rc = main_vm_entry_point_batch(engine, ios, isa.isa, hs, hw, cycles, trace_out=hcomp, callstack_witness=hc,
                                                  sponge_records_out=hrec)
    assert rc == 0 and coms4.tolist() == coms.tolist() and st4[0].reserved == n_jobs
    assert np.array_equal(abi.vm_expand_compact_trace(hcomp, hrec, cycles), want)
    # too small a record buffer: the count still tells how many there were
    small = np.zeros(5, dtype=abi.VM_SPONGE_RECORD_DTYPE)
    _, _, st5, rc = main_vm_entry_point_batch(engine, ios, isa.isa, hs, hw, cycles, trace_out=hcomp, callstack_witness=hc,
                                              sponge_records_out=small)
    assert rc == 0 and st5[0].reserved == n_jobs


@pytest.mark.parametrize("n_ops,cycles,seed", [(1024, 20000, 1), (1024, 20000, 3), (512, 6000, 8)])
def test_far_calls_bit_exact(engine, orc, n_ops, cycles, seed):
    """far calls (normal / delegate / mimic, static, to deployed, undeployed and kernel addresses, with exceptions) in the
    random mix: the witness comes from the oracle's out-of-circuit run (the GPU run models one frame's pages)"""
    isa, io, st = fresh(orc)
    io.default_aa_code_hash[7] = (1 << 24) | 5; io.default_aa_code_hash[2] = 0xA1
    ops = I.random_program(isa, n_ops, seed=seed, far_calls=True)
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True, gc=io)
    assert rc == 0, (hex(status.failed_checks), status.first_bad_row)
    io = with_tail(io, tail)
    want = O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles, cw=cw)
    assert want[0] == 0, (want[0], hex(want[4].failed_checks), want[4].first_bad_row)
    assert want[2][K["OP_AUX"] + 46].sum() > 10
    got = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, snaps, wit, cw), cycles)
    assert_same(want, got)
    # corrupted far-call witnesses: the code hash read, the suggested decommit page
    row = int(np.flatnonzero((want[2][K["OP_AUX"] + 46] == 1) & (want[2][K["SPONGE_ENFORCE"] + 8] == 1))[0])
    for byte in (80, 76):
        w2 = wit.copy(); w2[row, byte] ^= 1
        want2 = O.vm_entry_point(orc, io, isa.isa, snaps, w2, cycles, cw=cw)
        got2 = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, snaps, w2, cw), cycles, raise_on_unsatisfied=False)
        assert want2[0] != 0
        assert_same(want2, got2, check_trace=False)
    # the GPU out-of-circuit run refuses far calls loudly
    sim = main_vm_simulate(engine, isa.isa, [st], I.pack_code(ops)[None], 2000)
    assert sim.status.code == abi.ZKC_ERR_UNSUPPORTED


def test_far_call_hand_written(engine, orc):
    import test_oracle_main_vm_ops as T
    isa, io, st = T.fresh(orc, tail=31)
    abi_reg = 100000 << 192
    T.set_reg(st, 2, abi_reg); T.set_reg(st, 3, 0x9001); T.set_reg(st, 4, 0x9002)
    ops = T.far_call_program(isa)
    cycles = 20
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
    assert rc == 0
    io2 = with_tail(io, tail); io2.start_flag = 0; io2.hidden_fsm_input = O.vm_state_at(snaps, 0)
    want = O.vm_entry_point(orc, io2, isa.isa, snaps, wit, cycles, cw=cw)
    assert want[0] == 0
    got = main_vm_entry_point(engine, VmCircuitWitness(io2, isa.isa, snaps, wit, cw), cycles)
    assert_same(want, got)
    # instance boundary right behind the far call: the computed final state carries the new frame + decommit queue
    want4 = O.vm_entry_point(orc, io2, isa.isa, snaps[:5], wit[:4], 4, cw=cw)
    got4 = main_vm_entry_point(engine, VmCircuitWitness(io2, isa.isa, snaps[:5], wit[:4], cw), 4)
    assert_same(want4, got4)
    bad = snaps.copy(); bad[4, 1176 - 96 + 8] ^= 1  # decommitment queue state of the snapshot behind the far call
    want5 = O.vm_entry_point(orc, io2, isa.isa, bad[:5], wit[:4], 4, cw=cw)
    got5 = main_vm_entry_point(engine, VmCircuitWitness(io2, isa.isa, bad[:5], wit[:4], cw), 4, raise_on_unsatisfied=False)
    assert want5[0] == abi.ZKC_ERR_SNAPSHOT_MISMATCH
    assert_same(want5, got5, check_trace=False)


def test_full_size_instance_properties(engine, orc):
    """BASELINE.json configs[1] size (one instance of 2^20 cycles, the C2 instruction mix): the oracle is too slow to
    replay it here, so parity rests on size-independent properties -- every snapshot link, sponge chain and queue join
    verifies (status OK); the instance cut in two chained instances (hidden_fsm_output -> hidden_fsm_input) reproduces
    the same trace and final state; the host-buffer path (chunked H2D | kernels | D2H pipeline, COMPACT layout) returns
    the same witness as the device-resident one; and a 2^12-cycle prefix of the same run is bit-exact against the oracle."""
    import torch
    from era_zkevm_circuits_b200 import main_vm_entry_point_batch
    cycles = 1 << 20
    isa = I.Isa()
    io = abi.VmClosedForm(); io.start_flag = 1; io.rollback_queue_tail_for_block[0] = 0xC2
    st = O.vm_initial_state(orc, io, isa.isa)
    code = I.pack_code(I.random_program(isa, 4096, seed=0xC2))
    sim = main_vm_simulate(engine, isa.isa, [st], code[None], cycles)
    assert sim.status.code == 0
    io = with_tail(io, sim.rollback_tails[0])
    n_cw = int(sim.n_callstack[0])
    cw = sim.callstack_witness[:, :max(1, n_cw)].contiguous()
    trace = torch.empty((1, K["NUM_COLS"], cycles), dtype=torch.int64, device="cuda")
    coms, out, sts, rc = main_vm_entry_point_batch(engine, [io], isa.isa, sim.snapshots, sim.witness, cycles, trace_out=trace, callstack_witness=cw)
    assert rc == 0 and sts[0].code == 0 and sts[0].failed_checks == 0
    # the instruction mix is what the bench claims
    props = trace[0, K["PROPS"]]
    share = {op: float(((props >> op) & 1).sum()) / cycles for op in range(16)}
    assert 0.30 < share[I.OP_ADD] + share[I.OP_SUB] < 0.60 and 0.05 < share[I.OP_UMA] < 0.15 and 0.01 < share[I.OP_LOG] < 0.05
    assert 0.005 < share[I.OP_NEAR_CALL] < 0.03 and share[I.OP_RET] > 0.005 and share[I.OP_FAR_CALL] == 0
    # queue bookkeeping: the memory queue grew by exactly the enforced memory relations
    fin = out[0].hidden_fsm_output
    assert fin.memory_queue_length == int(trace[0, K["MEMQ_LENGTH_OUT"], -1])
    # two chained instances == the whole
    cut = 333333
    ta = torch.empty((1, K["NUM_COLS"], cut), dtype=torch.int64, device="cuda")
    ca, oa, sa, rc = main_vm_entry_point_batch(engine, [io], isa.isa, sim.snapshots[:, :cut + 1].contiguous(), sim.witness[:, :cut].contiguous(), cut,
                                               trace_out=ta, callstack_witness=cw)
    assert rc == 0
    nxt = abi.VmClosedForm.from_buffer_copy(bytes(oa[0])); nxt.start_flag = 0; nxt.hidden_fsm_input = oa[0].hidden_fsm_output
    tb = torch.empty((1, K["NUM_COLS"], cycles - cut), dtype=torch.int64, device="cuda")
    cb, ob, sb, rc = main_vm_entry_point_batch(engine, [nxt], isa.isa, sim.snapshots[:, cut:].contiguous(), sim.witness[:, cut:].contiguous(), cycles - cut,
                                               trace_out=tb, callstack_witness=cw)
    assert rc == 0
    lib = O.load()
    assert np.array_equal(flat(lib, ob[0].hidden_fsm_output), flat(lib, fin))
    assert torch.equal(torch.cat([ta, tb], dim=2), trace)
    del ta, tb
    # host buffers, COMPACT layout
    hs, hw, hc = (np.ascontiguousarray(x.cpu().numpy()) for x in (sim.snapshots, sim.witness, cw))
    hcomp = np.zeros((1, abi.VM_COMPACT_COLS, cycles), dtype=np.uint64)
    hrec = np.zeros(int(cycles * 1.25), dtype=abi.VM_SPONGE_RECORD_DTYPE)
    c2, o2, s2, rc = main_vm_entry_point_batch(engine, [io], isa.isa, hs, hw, cycles, trace_out=hcomp, callstack_witness=hc, sponge_records_out=hrec)
    assert rc == 0 and c2.tolist() == coms.tolist()
    n_rec = s2[0].reserved
    assert n_rec == int(trace[0, K["SPONGE_ENFORCE"]:K["SPONGE_ENFORCE"] + 9].sum())
    dense = trace.cpu().numpy().view(np.uint64)
    assert np.array_equal(hcomp[0, :K["SPONGE_ENFORCE"]], dense[0, :K["SPONGE_ENFORCE"]]) and np.array_equal(hcomp[0, abi.VM_COMPACT_OP_AUX:], dense[0, K["OP_AUX"]:])
    rec = hrec[:n_rec]
    assert np.array_equal(dense[0, K["SPONGE_FINAL"] + 12 * rec["slot"].astype(np.int64) + 3, rec["row"]], rec["out"][:, 3])
    # a prefix of the same run against the oracle
    pre = 1 << 12
    want = O.vm_entry_point(orc, io, isa.isa, hs[0, :pre + 1], hw[0, :pre], pre, cw=hc[0])
    assert want[0] == 0 and np.array_equal(want[2], dense[0, :, :pre])


def test_check_trace_constraint_evaluation(engine, orc):
    """zkc_main_vm_check_trace: valid traces (the ORACLE's own, and the engine's, far calls included) satisfy every
    row-local relation; a fault injected into any relation family is found, at its row, with its family bit"""
    import torch
    from era_zkevm_circuits_b200 import main_vm_check_trace
    V = abi.VMV
    isa, io, st = fresh(orc)
    cycles = 6000
    ops = I.random_program(isa, 1024, seed=21, far_calls=True)
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
    assert rc == 0
    io = with_tail(io, tail)
    want = O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles, cw=cw)
    assert want[0] == 0
    trace = want[2]
    viol, stt = main_vm_check_trace(engine, isa.isa, trace, cycles)       # the oracle's trace, host memory
    assert viol == 0 and stt.code == 0, (viol, hex(stt.failed_checks), stt.first_bad_row)
    got = main_vm_entry_point(engine, VmCircuitWitness(io, isa.isa, torch.from_numpy(snaps).cuda(), torch.from_numpy(wit).cuda(),
                                                       torch.from_numpy(np.ascontiguousarray(cw)).cuda()), cycles)
    viol, stt = main_vm_check_trace(engine, isa.isa, got.trace, cycles)   # the engine's trace, device memory
    assert viol == 0
    props = trace[K["PROPS"]]
    is_op = lambda op: np.flatnonzero((props >> np.uint64(op)) & np.uint64(1))
    faults = [
        (K["CONDITION"], 17, 2, V["BOOLEAN"]),
        (K["SUB_PC"], 40, 5, V["RANGE"]),
        (K["VARIANT"], 99, None, V["DECODE"]),
        (K["PROPS"], 123, None, V["DECODE"]),
        (K["ERGS_COST"], 200, None, V["DECODE"]),
        (K["MASK_INTO_NOP"], int(np.flatnonzero(trace[K["MASK_INTO_NOP"]] == 0)[5]), 1, V["EXCEPTION_MASKS"]),
        (K["DST0"] + 3, int(is_op(I.OP_ADD)[3]), None, V["ADD_SUB"]),
        (K["DST0"] + 1, int(is_op(I.OP_SUB)[2]), None, V["ADD_SUB"]),
        (K["DST1"] + 2, int(is_op(I.OP_MUL)[1]), None, V["MUL_DIV"]),
        (K["DST0"] + 1, int(is_op(I.OP_DIV)[1]), None, V["MUL_DIV"]),
        (K["DST0"] + 8, int(is_op(I.OP_BINOP)[4]), None, V["BINOP"]),
        (K["SPONGE_FINAL"] + 12 * 6 + 2, int(np.flatnonzero(trace[K["SPONGE_ENFORCE"] + 6] == 0)[0]), 77, V["SPONGE"]),
        (K["OP_AUX"] + 5, int(is_op(I.OP_ADD)[0]), 9, V["SELECTION"]),
        (K["DST1"] + 1, int(is_op(I.OP_ADD)[7]), 9, V["SELECTION"]),
    ]
    for col, row, val, bit in faults:
        bad = trace.copy()
        bad[col, row] = np.uint64(val) if val is not None else bad[col, row] ^ np.uint64(1)
        viol, stt = main_vm_check_trace(engine, isa.isa, bad, cycles)
        assert viol == 1 and stt.first_bad_row == row and stt.failed_checks & bit, (col, row, viol, stt.first_bad_row, hex(stt.failed_checks))
    # a batch of instances
    two = np.ascontiguousarray(np.stack([trace, trace]))
    two[1, K["CONDITION"], 5] = 3
    viol, stt = main_vm_check_trace(engine, isa.isa, two, cycles, n_instances=2)
    assert viol == 1 and stt.first_bad_row == cycles + 5


def test_column_inputs_equal_record_inputs(engine, orc):
    """zkc_main_vm_entry_point_columns (snapshots / oracle answers as device columns) == the record entry point == the oracle,
    for a batch whose instance boundaries are not warp-aligned; hidden_fsm_output of every instance included"""
    import torch
    from era_zkevm_circuits_b200 import main_vm_entry_point_batch, main_vm_entry_point_columns, main_vm_rows_to_columns
    n, cycles = 5, 1237
    isa = I.Isa()
    ios, snaps, wits, cws = [], [], [], []
    for i in range(n):
        io = abi.VmClosedForm(); io.start_flag = 1; io.rollback_queue_tail_for_block[1] = 31 + i
        st = O.vm_initial_state(orc, io, isa.isa)
        rc, sn, wi, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(I.random_program(isa, 700, seed=300 + i, far_calls=(i == 2))), cycles, full=True)
        assert rc == 0
        ios.append(with_tail(io, tail)); snaps.append(sn); wits.append(wi); cws.append(cw)
    cap = max(1, max(len(c) for c in cws))
    cwa = np.zeros((n, cap, C.sizeof(abi.VmCallstackWitness)), dtype=np.uint8)
    for i, c in enumerate(cws):
        cwa[i, :len(c)] = c
    d_s, d_w, d_c = torch.from_numpy(np.stack(snaps)).cuda(), torch.from_numpy(np.stack(wits)).cuda(), torch.from_numpy(cwa).cuda()
    t_rows = torch.empty((n, K["NUM_COLS"], cycles), dtype=torch.int64, device="cuda")
    t_cols = torch.empty_like(t_rows)
    c1, o1, s1, rc1 = main_vm_entry_point_batch(engine, ios, isa.isa, d_s, d_w, cycles, trace_out=t_rows, callstack_witness=d_c)
    cols = main_vm_rows_to_columns(engine, d_s, d_w, cycles)
    # the columns are what the header says: word w of snapshot i of instance k at [w, k * (limit + 1) + i]
    hs = np.stack(snaps).view(np.uint32).reshape(n * (cycles + 1), -1)
    assert np.array_equal(cols.state_words.cpu().numpy().view(np.uint32)[:, :n * (cycles + 1)], hs.T)
    hw = np.stack(wits).view(np.uint32).reshape(n * cycles, -1)
    assert np.array_equal(cols.witness_words.cpu().numpy().view(np.uint32)[:, :n * cycles], hw.T)
    c2, o2, s2, rc2 = main_vm_entry_point_columns(engine, ios, isa.isa, cols, cycles, trace_out=t_cols, callstack_witness=d_c)
    assert rc1 == 0 and rc2 == 0 and c1.tolist() == c2.tolist() and torch.equal(t_rows, t_cols)
    lib = O.load()
    for i in range(n):
        want = O.vm_entry_point(orc, ios[i], isa.isa, snaps[i], wits[i], cycles, cw=cws[i])
        assert want[0] == 0 and c2[i].tolist() == want[3].tolist()
        assert np.array_equal(t_cols[i].cpu().numpy().view(np.uint64), want[2])
        for o in (o1, o2):
            assert np.array_equal(flat(lib, o[i].hidden_fsm_output), flat(lib, want[1].hidden_fsm_output)), i
            assert o[i].completion_flag == want[1].completion_flag
    # a broken link is found at its row through the columns too
    bad = cols.state_words.clone()
    bad[abi.VM_STATE_WORDS - 3, 2 * (cycles + 1) + 400] ^= 1  # decommitment queue state of instance 2, snapshot 400
    c3, o3, s3, rc3 = main_vm_entry_point_columns(engine, ios, isa.isa, type(cols)(bad, cols.witness_words), cycles, callstack_witness=d_c)
    assert rc3 == abi.ZKC_ERR_SNAPSHOT_MISMATCH and [s.code for s in s3] == [0, 0, abi.ZKC_ERR_SNAPSHOT_MISMATCH, 0, 0]
    assert s3[2].first_bad_row in (399, 400)


@pytest.mark.parametrize("n,cycles,segment", [(1, 3000, 0), (3, 5000, 1024), (2, 4096, 2048), (1, 1, 0)])
def test_stream_inputs_and_packed_trace(engine, orc, n, cycles, segment):
    """zkc_main_vm_entry_point_stream: segmented input streams (dense / sparse words, expanded on the device) in, PACKED trace
    (typed columns + aux / sponge records) out == the oracle's dense trace, FSM output and commitment, bit-exactly"""
    from era_zkevm_circuits_b200 import (main_vm_entry_point_stream, vm_encode_input_stream, vm_expand_packed_trace, vm_packed_trace_buffers)
    isa = I.Isa()
    ios, snaps, wits, cws = [], [], [], []
    for i in range(n):
        io = abi.VmClosedForm(); io.start_flag = 1; io.rollback_queue_tail_for_block[3] = 77 + i
        st = O.vm_initial_state(orc, io, isa.isa)
        rc, sn, wi, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(I.random_program(isa, 900, seed=500 + 7 * i + cycles, far_calls=(i == 1))), cycles, full=True)
        assert rc == 0
        ios.append(with_tail(io, tail)); snaps.append(sn); wits.append(wi); cws.append(cw)
    cap = max(1, max(len(c) for c in cws))
    cwa = np.zeros((n, cap, C.sizeof(abi.VmCallstackWitness)), dtype=np.uint8)
    for i, c in enumerate(cws):
        cwa[i, :len(c)] = c
    streams = [vm_encode_input_stream(engine.lib, snaps[i], wits[i], cycles, segment) for i in range(n)]
    out = vm_packed_trace_buffers(engine, n, cycles, aux_fraction=1.0, sponge_per_cycle=9.0)
    coms, o, sts, rc = main_vm_entry_point_stream(engine, ios, isa.isa, streams, cycles, callstack_witness=cwa, out=out)
    assert rc == 0, [(s.code, hex(s.failed_checks), s.first_bad_row) for s in sts]
    dense = vm_expand_packed_trace(engine.lib, out, n, cycles)
    lib = O.load()
    n_aux = 0
    for i in range(n):
        want = O.vm_entry_point(orc, ios[i], isa.isa, snaps[i], wits[i], cycles, cw=cws[i])
        assert want[0] == 0 and coms[i].tolist() == want[3].tolist()
        bad = np.argwhere(dense[i] != want[2])
        assert bad.size == 0, f"instance {i}: first differing (col,row): {bad[:8].tolist()}"
        assert np.array_equal(flat(lib, o[i].hidden_fsm_output), flat(lib, want[1].hidden_fsm_output))
        t = want[2]
        moved = np.ones(cycles, dtype=bool)
        moved[1:] = (t[K["FORWARD_TAIL_OUT"]:K["FORWARD_TAIL_OUT"] + 10, 1:] != t[K["FORWARD_TAIL_OUT"]:K["FORWARD_TAIL_OUT"] + 10, :-1]).any(axis=0)
        n_aux += int((moved | (t[K["OP_AUX"]:K["OP_AUX"] + 48] != 0).any(axis=0)).sum())
    assert out.n_aux_records == n_aux
    assert out.n_sponge_records == int(sum(O.vm_entry_point(orc, ios[i], isa.isa, snaps[i], wits[i], cycles, cw=cws[i])[2][K["SPONGE_ENFORCE"]:K["SPONGE_ENFORCE"] + 9].sum() for i in range(n)))
    # no witness wanted: commitments only
    coms2, _, _, rc2 = main_vm_entry_point_stream(engine, ios, isa.isa, streams, cycles, callstack_witness=cwa, out=None)
    assert rc2 == 0 and coms2.tolist() == coms.tolist()
    # record buffers too small: the counts still say how many there were
    small = vm_packed_trace_buffers(engine, n, cycles, aux_fraction=0.0, sponge_per_cycle=0.0)
    coms3, _, _, rc3 = main_vm_entry_point_stream(engine, ios, isa.isa, streams, cycles, callstack_witness=cwa, out=small)
    assert rc3 == 0 and small.n_aux_records == out.n_aux_records and small.n_sponge_records == out.n_sponge_records
    assert np.array_equal(small.cols32, out.cols32) and np.array_equal(small.cols8, out.cols8)
    for s in streams:
        s.free()


@pytest.mark.parametrize("n,cycles,seed,far", [(1, 1, 1, False), (1, 5000, 2, False), (3, 4097, 3, True)])
def test_gadget_cells_bit_exact(engine, orc, n, cycles, seed, far):
    """zkc_main_vm_gadget_cells: the oblivious add/sub, binop, mul/div, shift cells and the per-cycle relations, CUDA vs the oracle
    (which tests/test_oracle_main_vm_gadgets.py pins on Python integers), host and device buffers, batches"""
    import torch
    from era_zkevm_circuits_b200 import main_vm_gadget_cells
    isa, io, st = fresh(orc)
    traces = []
    for k in range(n):
        ops = I.random_program(isa, 1024, seed=seed + k, far_calls=far)
        rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
        assert rc == 0
        want = O.vm_entry_point(orc, with_tail(io, tail), isa.isa, snaps, wit, cycles, cw=cw)
        assert want[0] == 0
        traces.append(want[2])
    trace = np.ascontiguousarray(np.stack(traces))
    want = O.vm_gadget_cells(orc, trace, cycles, n)
    got = main_vm_gadget_cells(engine, trace, cycles, n)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"first differing (instance, column, row): {bad[:5].tolist()}"
    dev = main_vm_gadget_cells(engine, torch.from_numpy(trace.view(np.int64)).cuda(), cycles, n)
    assert np.array_equal(dev.cpu().numpy().view(np.uint64), want)
