"""The segmented input stream of the main_vm call (include/zkc_b200.h, zkc_vm_input_stream): the library's host encoder against
a numpy reference decoder -- no GPU needed (the encoder falls back to pageable memory without a device)."""
import ctypes as C

import numpy as np
import pytest

import orc as O
from era_zkevm_circuits_b200 import abi, isa as I
from era_zkevm_circuits_b200.main_vm import vm_decode_input_stream, vm_encode_input_stream, vm_packed_layout


def run(orc, cycles, seed):
    isa = I.Isa()
    io = abi.VmClosedForm(); io.start_flag = 1; io.rollback_queue_tail_for_block[0] = seed
    st = O.vm_initial_state(orc, io, isa.isa)
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(I.random_program(isa, 600, seed=seed)), cycles, full=True)
    assert rc == 0
    return snaps, wit


@pytest.mark.parametrize("cycles,segment", [(0, 0), (1, 0), (777, 0), (5000, 1024), (4096, 1024), (3000, 3000), (2049, 2048)])
def test_encode_decode_round_trip(orc, cycles, segment):
    lib = abi.load_library()
    snaps, wit = run(orc, max(cycles, 1), 7 + cycles)
    s = vm_encode_input_stream(lib, snaps, wit, cycles, segment)
    st = s.struct
    seg = segment or 1 << 16
    assert st.limit == cycles and st.segment_cycles == seg and st.n_segments == max(1, -(-cycles // seg))
    state, w = vm_decode_input_stream(s)
    assert np.array_equal(state, snaps[:cycles + 1].view(np.uint32).T)
    assert np.array_equal(w, wit[:cycles].view(np.uint32).T)
    if cycles >= 777:  # the point of the format: far fewer bytes than the records
        assert s.bytes < 0.4 * (snaps[:cycles + 1].nbytes + wit[:cycles].nbytes), s.bytes
    total = sum(st.segments[k].blob_bytes for k in range(st.n_segments))
    assert total == s.bytes
    s.free()


def test_every_change_costs_at_most_a_dense_word(orc):
    """a word goes dense when its change list would be longer: random snapshots degrade to (almost) the record size"""
    lib = abi.load_library()
    rng = np.random.default_rng(3)
    n = 300
    snaps = rng.integers(0, 256, (n + 1, C.sizeof(abi.VmState)), dtype=np.uint8)
    wit = rng.integers(0, 256, (n, C.sizeof(abi.VmCycleWitness)), dtype=np.uint8)
    s = vm_encode_input_stream(lib, snaps, wit, n, 128)
    state, w = vm_decode_input_stream(s)
    assert np.array_equal(state, snaps.view(np.uint32).T) and np.array_equal(w, wit.view(np.uint32).T)
    assert s.bytes < 1.02 * (snaps.nbytes + wit.nbytes) + 3 * 8192
    s.free()


def test_packed_layout_covers_every_column():
    lib = abi.load_library()
    kind, slot, counts = vm_packed_layout(lib)
    K = abi.VM_COLS
    assert int(counts.sum()) == K["NUM_COLS"] and counts[abi.VM_PK_SPONGE_RECORD] == 117 and counts[abi.VM_PK_AUX_RECORD] == 58
    for k in range(6):
        assert sorted(slot[kind == k].tolist()) == list(range(int(counts[k])))
    assert counts[abi.VM_PK_LIMB_RECORD] == 26 and kind[K["CODE_WORD"] + 7] == kind[K["SRC0_FROM_MEMORY"]] == kind[K["DST1"] + 8] == abi.VM_PK_LIMB_RECORD
    assert slot[K["CODE_WORD"] + 3] == 3 and slot[K["SRC0_FROM_MEMORY"] + 2] == 9 + 2 and slot[K["DST1"]] == 18
    assert kind[K["PROPS"]] == abi.VM_PK_U64 and kind[K["OPCODE"]] == abi.VM_PK_U32 and kind[K["IMM0"]] == abi.VM_PK_U16
    assert kind[K["CONDITION"]] == abi.VM_PK_U8 and kind[K["OP_AUX"] + 47] == abi.VM_PK_AUX_RECORD
    assert all(kind[K["FORWARD_TAIL_OUT"] + i] == abi.VM_PK_AUX_RECORD for i in range(10))
    # bytes per cycle of the four typed blocks
    assert int(counts[0]) + 2 * int(counts[1]) + 4 * int(counts[2]) + 8 * int(counts[3]) < 220
