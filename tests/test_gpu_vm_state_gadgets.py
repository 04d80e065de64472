"""zkc_main_vm_state_gadget_cells (the ptr / jump / context gadget-cell block, include/zkc_b200.h ZKC_VM_STATE_GADGET_COLUMNS): CUDA vs
the oracle (which tests/test_oracle_main_vm_gadgets.py pins on Python integers), host and device buffers, batches of instances."""
import numpy as np
import pytest

import orc as O
from era_zkevm_circuits_b200 import abi, isa as I
from era_zkevm_circuits_b200.engine import ZkcError
from test_gpu_main_vm import fresh, with_tail

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,cycles,seed,far", [(1, 1, 1, False), (1, 5000, 2, False), (3, 4097, 3, True)])
def test_state_gadget_cells_bit_exact(engine, orc, n, cycles, seed, far):
    import torch
    from era_zkevm_circuits_b200 import main_vm_state_gadget_cells
    isa, io, st = fresh(orc)
    traces, snapshots = [], []
    for k in range(n):
        ops = I.random_program(isa, 1024, seed=seed + k, far_calls=far)
        rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
        assert rc == 0
        want = O.vm_entry_point(orc, with_tail(io, tail), isa.isa, snaps, wit, cycles, cw=cw)
        assert want[0] == 0
        traces.append(want[2])
        snapshots.append(snaps)
    trace = np.ascontiguousarray(np.stack(traces))
    snaps = np.ascontiguousarray(np.stack(snapshots))
    want = O.vm_state_gadget_cells(orc, trace, snaps, cycles, n)
    assert want.shape == (n, abi.VMS_COLS["NUM_COLS"], cycles)
    got = main_vm_state_gadget_cells(engine, trace, snaps, cycles, n)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"first differing (instance, column, row): {bad[:5].tolist()}"
    dev = main_vm_state_gadget_cells(engine, torch.from_numpy(trace.view(np.int64)).cuda(), torch.from_numpy(snaps).cuda(), cycles, n)
    assert np.array_equal(dev.cpu().numpy().view(np.uint64), want)
    # single-instance form: [NUM_COLS, limit] in, [VMS NUM_COLS, limit] out
    one = main_vm_state_gadget_cells(engine, trace[0], snaps[0], cycles)
    assert np.array_equal(one, want[0])


def test_state_gadget_cells_argument_checks(engine):
    from era_zkevm_circuits_b200 import main_vm_state_gadget_cells
    trace = np.zeros((abi.VM_COLS["NUM_COLS"], 8), dtype=np.uint64)
    with pytest.raises(ZkcError):                                           # 8 snapshots for 8 cycles: one short
        main_vm_state_gadget_cells(engine, trace, np.zeros((8, 1176), dtype=np.uint8), 8)
    out = main_vm_state_gadget_cells(engine, trace, np.zeros((9, 1176), dtype=np.uint8), 8)
    assert out.shape == (abi.VMS_COLS["NUM_COLS"], 8)
    S = abi.VMS_COLS
    # all-zero inputs: no opcode applies; src1 = 0 is an integer with every limb zero; the increment of tx_number 0 is 1
    assert out[S["PTR_SRC1_IS_INTEGER"]].tolist() == [1] * 8 and out[S["PTR_ARGS_INVALID"]].tolist() == [1] * 8
    assert out[S["PTR_SRC1_LIMB_IS_ZERO"]:S["PTR_SRC1_LIMB_IS_ZERO"] + 8].min() == 1 and out[S["CTX_INCREMENTED_TX_NUMBER"]].tolist() == [1] * 8
    assert out[S["CTX_WRITE_LIKE"]].tolist() == [1] * 8 and out[S["CTX_RESULT_256"]:S["CTX_RESULT_256"] + 8].max() == 0


@pytest.mark.parametrize("n,cycles,seed,far", [(1, 1, 1, False), (1, 5000, 2, False), (3, 4097, 3, True)])
def test_memory_sponge_cells_bit_exact(engine, orc, n, cycles, seed, far):
    """zkc_main_vm_memory_sponge_cells: the fetch / src0 read / dst0 write relations of every cycle (three chained Poseidon2
    permutations per cycle), CUDA vs the oracle, host and device buffers, batches; and against the engine's OWN dense trace: wherever
    a slot is enforced there, the permutation output is the same value"""
    import torch
    from era_zkevm_circuits_b200 import main_vm_memory_sponge_cells
    isa, io, st = fresh(orc)
    traces, snapshots = [], []
    for k in range(n):
        ops = I.random_program(isa, 1024, seed=seed + k, far_calls=far)
        rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
        assert rc == 0
        want = O.vm_entry_point(orc, with_tail(io, tail), isa.isa, snaps, wit, cycles, cw=cw)
        assert want[0] == 0
        traces.append(want[2])
        snapshots.append(snaps)
    trace = np.ascontiguousarray(np.stack(traces))
    snaps = np.ascontiguousarray(np.stack(snapshots))
    want = O.vm_memory_sponge_cells(orc, trace, snaps, cycles, n)
    got = main_vm_memory_sponge_cells(engine, trace, snaps, cycles, n)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"first differing (instance, column, row): {bad[:5].tolist()}"
    dev = main_vm_memory_sponge_cells(engine, torch.from_numpy(trace.view(np.int64)).cuda(), torch.from_numpy(snaps).cuda(), cycles, n)
    assert np.array_equal(dev.cpu().numpy().view(np.uint64), want)
    K, Q = abi.VM_COLS, abi.VMQ_COLS
    for slot, name in enumerate(("FETCH", "SRC0", "DST0")):
        enforced = (trace[:, K["SPONGE_ENFORCE"] + slot] != 0) & ((got[:, Q["SELECTED"]] != 0) | (slot == 0))
        a = trace[:, K["SPONGE_FINAL"] + 12 * slot:K["SPONGE_FINAL"] + 12 * slot + 12]
        b = got[:, Q[name + "_FINAL"]:Q[name + "_FINAL"] + 12]
        assert np.array_equal(a.transpose(0, 2, 1)[enforced], b.transpose(0, 2, 1)[enforced]), name


@pytest.mark.parametrize("n,cycles,seed,far", [(1, 1, 1, False), (1, 5000, 2, False), (3, 4097, 3, True)])
def test_prestate_cells_bit_exact(engine, orc, n, cycles, seed, far):
    """zkc_main_vm_prestate_cells: the cells of create_prestate that are not DENSE columns (selector masks, the 15-step register select
    chains, operand locations, src0 selects, swap, erasure flags), CUDA vs the oracle (pinned on Python integers and on the DENSE
    trace's own results by tests/test_oracle_main_vm_gadgets.py), host and device buffers, batches"""
    import torch
    from era_zkevm_circuits_b200 import main_vm_prestate_cells
    isa, io, st = fresh(orc)
    traces, snapshots = [], []
    for k in range(n):
        ops = I.random_program(isa, 1024, seed=seed + k, far_calls=far)
        rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
        assert rc == 0
        want = O.vm_entry_point(orc, with_tail(io, tail), isa.isa, snaps, wit, cycles, cw=cw)
        assert want[0] == 0
        traces.append(want[2])
        snapshots.append(snaps)
    trace = np.ascontiguousarray(np.stack(traces))
    snaps = np.ascontiguousarray(np.stack(snapshots))
    want = O.vm_prestate_cells(orc, trace, snaps, cycles, n)
    assert want.shape == (n, abi.VMP_COLS["NUM_COLS"], cycles)
    got = main_vm_prestate_cells(engine, trace, snaps, cycles, n)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"first differing (instance, column, row): {bad[:5].tolist()}"
    dev = main_vm_prestate_cells(engine, torch.from_numpy(trace.view(np.int64)).cuda(), torch.from_numpy(snaps).cuda(), cycles, n)
    assert np.array_equal(dev.cpu().numpy().view(np.uint64), want)
    one = main_vm_prestate_cells(engine, trace[0], snaps[0], cycles)
    assert np.array_equal(one, want[0])
    # the block's last cells are the DENSE trace's operands before the erasure (base_structures/register/mod.rs:67-76)
    K, P = abi.VM_COLS, abi.VMP_COLS
    for src, swapped, flag in (("SRC0", "SRC0_SWAPPED", "SHOULD_ERASE_SRC0"), ("SRC1", "SRC1_SWAPPED", "SHOULD_ERASE_SRC1")):
        keep = np.ones((1, 9, 1), dtype=np.uint64); erase = got[:, P[flag]][:, None, :]
        keep = np.where(np.isin(np.arange(9), (0, 2, 3))[None, :, None], 1 - erase, keep)
        assert np.array_equal(trace[:, K[src]:K[src] + 9], got[:, P[swapped]:P[swapped] + 9] * keep), src


def test_prestate_cells_mutated_inputs_and_argument_checks(engine, orc):
    """user mode, pointer registers, every register index, the 16- / 32-bit wraps, skipped and pending cycles, any property bits"""
    from era_zkevm_circuits_b200 import main_vm_prestate_cells
    from test_oracle_main_vm_gadgets import mutated_prestate_inputs
    trace, snaps = mutated_prestate_inputs(orc, 1500, 7)
    want = O.vm_prestate_cells(orc, trace, snaps, 1500)
    got = main_vm_prestate_cells(engine, np.ascontiguousarray(trace), np.ascontiguousarray(snaps), 1500)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"first differing (column, row): {bad[:5].tolist()}"
    zero = np.zeros((abi.VM_COLS["NUM_COLS"], 8), dtype=np.uint64)
    with pytest.raises(ZkcError):                                           # 8 snapshots for 8 cycles: one short
        main_vm_prestate_cells(engine, zero, np.zeros((8, 1176), dtype=np.uint8), 8)
    out = main_vm_prestate_cells(engine, zero, np.zeros((9, 1176), dtype=np.uint8), 8)
    P = abi.VMP_COLS
    assert out.shape == (P["NUM_COLS"], 8)
    # all-zero inputs: the cycle executes, pc + 1 = 1, pages 1 / 2 / 3, no selector set, every chain stays zero
    assert out[P["EXECUTE_CYCLE"]].tolist() == [1] * 8 and out[P["PC_PLUS_ONE"]].tolist() == [1] * 8 and out[P["AUX_HEAP_PAGE"]].tolist() == [3] * 8
    assert out[P["SRC0_SELECTORS"]:P["DST1_SELECTORS"] + 15].max() == 0 and out[P["DRAFT_SRC0_CHAIN"]:P["DST0_REG_LOW_CHAIN"] + 15].max() == 0
    assert out[P["TIMESTAMPS"]:P["TIMESTAMPS"] + 4, 0].tolist() == [1, 2, 3, 4] and out[P["CAN_SKIP_READ"]].tolist() == [1] * 8


@pytest.mark.parametrize("n,cycles,seed,far", [(1, 1, 1, False), (1, 5000, 2, False), (3, 4097, 3, True)])
def test_writeback_cells_bit_exact(engine, orc, n, cycles, seed, far):
    """zkc_main_vm_writeback_cells: the register write-back of the state diffs (cycle.rs:158-433), CUDA vs the oracle (pinned on Python
    integers by tests/test_oracle_main_vm_gadgets.py), host and device buffers, batches; and the end of every register's select chain
    is that register in snapshot i + 1"""
    import torch
    from era_zkevm_circuits_b200 import main_vm_writeback_cells
    isa, io, st = fresh(orc)
    traces, snapshots = [], []
    for k in range(n):
        ops = I.random_program(isa, 1024, seed=seed + k, far_calls=far)
        rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
        assert rc == 0
        want = O.vm_entry_point(orc, with_tail(io, tail), isa.isa, snaps, wit, cycles, cw=cw)
        assert want[0] == 0
        traces.append(want[2])
        snapshots.append(snaps)
    trace = np.ascontiguousarray(np.stack(traces))
    snaps = np.ascontiguousarray(np.stack(snapshots))
    want = O.vm_writeback_cells(orc, isa.isa, trace, snaps, cycles, n)
    W = abi.VMW_COLS
    assert want.shape == (n, W["NUM_COLS"], cycles)
    got = main_vm_writeback_cells(engine, isa.isa, trace, snaps, cycles, n)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"first differing (instance, column, row): {bad[:5].tolist()}"
    dev = main_vm_writeback_cells(engine, isa.isa, torch.from_numpy(trace.view(np.int64)).cuda(), torch.from_numpy(snaps).cuda(), cycles, n)
    assert np.array_equal(dev.cpu().numpy().view(np.uint64), want)
    one = main_vm_writeback_cells(engine, isa.isa, trace[0], snaps[0], cycles)
    assert np.array_equal(one, want[0])
    words = np.frombuffer(snaps.tobytes(), dtype=np.uint32).reshape(n, -1, 294)[:, :cycles + 1]
    base = abi.VmState.registers.offset // 4
    for r in range(15):
        nxt = words[:, 1:, base + 9 * r:base + 9 * r + 9].transpose(0, 2, 1).astype(np.uint64)   # [n, 9, cycles]
        assert np.array_equal(got[:, W["IS_PTR_AFTER_DST1"] + r], nxt[:, 0] & 1), r
        assert np.array_equal(got[:, W["VALUE_AFTER_DST1"] + 8 * r:W["VALUE_AFTER_DST1"] + 8 * r + 8], nxt[:, 1:]), r


def test_writeback_cells_mutated_inputs_and_argument_checks(engine, orc):
    from era_zkevm_circuits_b200 import main_vm_writeback_cells
    from test_oracle_main_vm_gadgets import mutate_writeback_inputs, vm_trace_and_snapshots
    isa = I.Isa().isa
    trace, snaps = vm_trace_and_snapshots(orc, 1500, 7, True)
    trace = np.ascontiguousarray(trace.copy()); snaps = np.ascontiguousarray(snaps)
    mutate_writeback_inputs(trace, 7)
    want = O.vm_writeback_cells(orc, isa, trace, snaps, 1500)
    got = main_vm_writeback_cells(engine, isa, trace, snaps, 1500)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"first differing (column, row): {bad[:5].tolist()}"
    zero = np.zeros((abi.VM_COLS["NUM_COLS"], 8), dtype=np.uint64)
    with pytest.raises(ZkcError):                                           # 8 snapshots for 8 cycles: one short
        main_vm_writeback_cells(engine, isa, zero, np.zeros((8, 1176), dtype=np.uint8), 8)
    out = main_vm_writeback_cells(engine, isa, zero, np.zeros((9, 1176), dtype=np.uint8), 8)
    W = abi.VMW_COLS
    # all-zero inputs: nothing is written, dst0 would update a register (no memory access), no far call is a "non system" one
    assert out.shape == (W["NUM_COLS"], 8) and out[W["DST0_PERFORMS_REG_UPDATE"]].tolist() == [1] * 8 and out[W["FAR_CALL_NON_SYSTEM"]].tolist() == [1] * 8
    assert out[W["FAR_CALL_UPDATE"]].max() == 0 and out[W["WRITE_AS_DST0"]:].max() == 0
