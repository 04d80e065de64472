"""The row statement the CUDA kernel vm_prestate_kernel runs (era_zkevm_circuits_b200/csrc/main_vm_prestate_row.cuh, a
__host__ __device__ function) compiled by g++ (tests/cpp/prestate_row_host.cpp) and compared with the oracle on a box without a GPU:
the kernel's index arithmetic and every select of create_prestate, bit-exact, on programs that exercise every opcode family.  The
product has no CPU path -- this harness is test infrastructure; the GPU run of the same statement is tests/test_gpu_vm_state_gadgets.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import orc as O
from era_zkevm_circuits_b200 import abi
from test_oracle_main_vm_gadgets import vm_trace_and_snapshots

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_rows():
    out = os.path.join(ROOT, "build", "libprestate_row_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", out,
                           os.path.join(ROOT, "tests", "cpp", "prestate_row_host.cpp")])
    return C.CDLL(out).prestate_rows_host


@pytest.mark.parametrize("cycles,seed,far", [(3000, 33, False), (3000, 5, True), (257, 9, False)])
def test_kernel_row_statement_matches_oracle(orc, host_rows, cycles, seed, far):
    trace, snaps = vm_trace_and_snapshots(orc, cycles, seed, far)
    want = O.vm_prestate_cells(orc, trace, snaps, cycles)
    got = O.vm_prestate_cells(orc, trace, snaps, cycles, fn=host_rows)
    assert want.shape == (abi.VMP_COLS["NUM_COLS"], cycles)
    bad = np.argwhere(want != got)
    assert bad.size == 0, bad[:8]


def test_kernel_row_statement_instances(orc, host_rows):
    """[n, cols, limit] indexing: two different instances side by side equal the two single-instance results"""
    a = vm_trace_and_snapshots(orc, 500, 1, False)
    b = vm_trace_and_snapshots(orc, 500, 2, True)
    trace = np.stack([a[0], b[0]]); snaps = np.stack([a[1][:501], b[1][:501]])
    both = O.vm_prestate_cells(orc, trace, snaps, 500, 2, fn=host_rows)
    assert np.array_equal(both[0], O.vm_prestate_cells(orc, a[0], a[1], 500))
    assert np.array_equal(both[1], O.vm_prestate_cells(orc, b[0], b[1], 500))
    assert np.array_equal(both, O.vm_prestate_cells(orc, trace, snaps, 500, 2))


def test_kernel_row_statement_on_mutated_inputs(orc, host_rows):
    """user mode, pointer registers, every register index, the 16- and 32-bit wraps, skipped / pending cycles, any property bits"""
    from test_oracle_main_vm_gadgets import mutated_prestate_inputs
    for seed in (7, 8, 9):
        trace, snaps = mutated_prestate_inputs(orc, 1500, seed)
        want = O.vm_prestate_cells(orc, trace, snaps, 1500)
        got = O.vm_prestate_cells(orc, trace, snaps, 1500, fn=host_rows)
        bad = np.argwhere(want != got)
        assert bad.size == 0, bad[:8]


@pytest.fixture(scope="module")
def host_writeback_rows(host_rows):
    return C.CDLL(os.path.join(ROOT, "build", "libprestate_row_host.so")).writeback_rows_host


@pytest.mark.parametrize("cycles,seed,far", [(3000, 33, False), (3000, 5, True), (3000, 11, True), (257, 9, False)])
def test_writeback_row_statement_matches_oracle_and_the_next_snapshot(orc, host_writeback_rows, cycles, seed, far):
    """vm_writeback_kernel's row statement (csrc/main_vm_writeback_row.cuh) compiled by g++ against the oracle; and, for both, the END of
    each register's select chain is that register in snapshot i + 1 (every cycle: ordinary writes, encoded-but-unflagged dst1, far
    calls, far returns)"""
    from era_zkevm_circuits_b200 import isa as I
    isa = I.Isa().isa
    trace, snaps = vm_trace_and_snapshots(orc, cycles, seed, far)
    want = O.vm_writeback_cells(orc, isa, trace, snaps, cycles)
    got = O.vm_writeback_cells(orc, isa, trace, snaps, cycles, fn=host_writeback_rows)
    W = abi.VMW_COLS
    assert want.shape == (W["NUM_COLS"], cycles) and W["NUM_COLS"] == 513
    bad = np.argwhere(want != got)
    assert bad.size == 0, bad[:8]
    words = np.frombuffer(np.ascontiguousarray(snaps).tobytes(), dtype=np.uint32).reshape(-1, 294)[:cycles + 1]
    base = abi.VmState.registers.offset // 4
    for r in range(15):
        nxt = words[1:, base + 9 * r:base + 9 * r + 9].T.astype(np.uint64)        # [9, cycles]: is_pointer, 8 limbs of register r + 1 AFTER each cycle
        assert np.array_equal(got[W["IS_PTR_AFTER_DST1"] + r], nxt[0] & 1), r
        assert np.array_equal(got[W["VALUE_AFTER_DST1"] + 8 * r:W["VALUE_AFTER_DST1"] + 8 * r + 8], nxt[1:]), r
    if far:
        assert got[W["FAR_CALL_UPDATE"]].sum() > 0 and got[W["FAR_RETURN_UPDATE"]].sum() > 0 and got[W["ZERO_OUT"]:W["ZERO_OUT"] + 15].sum() > 0


def test_writeback_row_statement_instances(orc, host_writeback_rows):
    from era_zkevm_circuits_b200 import isa as I
    isa = I.Isa().isa
    a = vm_trace_and_snapshots(orc, 500, 1, False)
    b = vm_trace_and_snapshots(orc, 500, 2, True)
    trace = np.stack([a[0], b[0]]); snaps = np.stack([a[1][:501], b[1][:501]])
    both = O.vm_writeback_cells(orc, isa, trace, snaps, 500, 2, fn=host_writeback_rows)
    assert np.array_equal(both[0], O.vm_writeback_cells(orc, isa, a[0], a[1], 500))
    assert np.array_equal(both[1], O.vm_writeback_cells(orc, isa, b[0], b[1], 500))


def test_writeback_row_statement_on_mutated_inputs(orc, host_writeback_rows):
    """any property bits (several opcode kinds at once), system / constructor ABI bytes, kernel and user targets, every register index"""
    from era_zkevm_circuits_b200 import isa as I
    from test_oracle_main_vm_gadgets import mutate_writeback_inputs
    isa = I.Isa().isa
    for seed in (7, 8):
        trace, snaps = vm_trace_and_snapshots(orc, 1500, seed, True)
        trace = trace.copy()
        mutate_writeback_inputs(trace, seed)
        want = O.vm_writeback_cells(orc, isa, trace, snaps, 1500)
        got = O.vm_writeback_cells(orc, isa, trace, snaps, 1500, fn=host_writeback_rows)
        bad = np.argwhere(want != got)
        assert bad.size == 0, bad[:8]
        W = abi.VMW_COLS
        assert want[W["FAR_CALL_CLEANUP_REGISTER"]].sum() > 0 and (want[W["FAR_CALL_UPDATE"]] > want[W["FAR_CALL_CLEANUP_REGISTER"]]).sum() > 0
