"""bincode wire format of the reference's *CircuitInstanceWitness structs (era_zkevm_circuits_b200/wire.py): a hand-assembled
byte string of every primitive (what serde + bincode 1.x emit for bool / u32 / field element / U256 / Address / VecDeque),
write -> read round trips of whole witnesses, ingestion of a dump by the ORACLE (the ingested witness proves like the
original), and loud failures on malformed input."""
import struct

import numpy as np
import pytest

import helpers as H
import orc as O
import vectors as V
from era_zkevm_circuits_b200 import abi, synthetic, wire


def test_primitives_byte_for_byte():
    w = wire.Writer()
    w.boolean(True); w.u32(0x01020304); w.field(abi.GL_P - 1); w.u256(0); w.u256(0xABC << 200); w.h160(0x8010)
    want = (b"\x01" + b"\x04\x03\x02\x01" + struct.pack("<Q", abi.GL_P - 1) +
            struct.pack("<Q", 3) + b"0x0" +                                  # U256: hex string without leading zeros
            struct.pack("<Q", 2 + 53) + b"0xabc" + b"0" * 50 +
            struct.pack("<Q", 42) + b"0x" + b"0" * 36 + b"8010")             # H160: always 40 digits
    assert bytes(w.b) == want
    r = wire.Reader(want)
    assert (r.boolean(), r.u32(), r.field(), r.u256(), r.u256(), r.h160()) == (1, 0x01020304, abi.GL_P - 1, 0, 0xABC << 200, 0x8010)
    r.done()


def test_ram_permutation_dump_round_trip_and_ingestion(orc):
    from era_zkevm_circuits_b200 import RamPermutationCircuitInstanceWitness
    u, s = V.ram_reference_vector()
    io, up, sp = H.ram_instance(orc, u, s, nondet_len=1)
    dump = wire.write_ram_permutation_witness(RamPermutationCircuitInstanceWitness(io, u, up, s, sp))
    # header: start_flag, completion_flag, then the unsorted queue's head (12 field elements)
    assert dump[:2] == b"\x01\x00" and dump[2:2 + 96] == bytes(96)
    got = wire.read_ram_permutation_witness(dump)
    assert bytes(got.closed_form_input) == bytes(io)
    for a, b in ((got.unsorted_queue_witness, u), (got.sorted_queue_witness, s)):
        for name in ("timestamp", "memory_page", "index", "rw_flag", "is_ptr", "value"):
            assert np.array_equal(a[name], b[name]), name
    assert np.array_equal(got.unsorted_queue_prev_states, up) and np.array_equal(got.sorted_queue_prev_states, sp)
    assert wire.write_ram_permutation_witness(got) == dump
    # the ingested witness proves exactly like the original
    want = O.ram_entry_point(orc, io, u, s, 16)
    have = O.ram_entry_point(orc, got.closed_form_input, got.unsorted_queue_witness, got.sorted_queue_witness, 16)
    assert want[0] == have[0] == abi.ZKC_OK and have[3].tolist() == want[3].tolist() and np.array_equal(have[2], want[2])
    # a chained (not start) instance with a non-trivial FSM input
    u2, s2 = synthetic.ram_trace(300, seed=4, n_cells=30, n_nondet=2)
    io2, up2, sp2 = H.ram_instance(orc, u2, s2, 2)
    first = O.ram_entry_point(orc, io2, u2, s2, 100)
    nxt = abi.RamClosedForm.from_buffer_copy(bytes(first[1])); nxt.start_flag = 0
    nxt.hidden_fsm_input = first[1].hidden_fsm_output
    w2 = RamPermutationCircuitInstanceWitness(nxt, u2[100:], up2[100:], s2[100:], sp2[100:])
    back = wire.read_ram_permutation_witness(wire.write_ram_permutation_witness(w2))
    assert bytes(back.closed_form_input) == bytes(nxt) and len(back.unsorted_queue_witness) == 200


def test_events_deduplicator_dump_round_trip(orc):
    from era_zkevm_circuits_b200 import EventsDeduplicatorInstanceWitness
    u, s = synthetic.events_trace(200, seed=6, rollback_pct=20)
    up, ufin = O.log_queue_simulate(orc, u)
    sp, sfin = O.log_queue_simulate(orc, s)
    io = O.events_closed_form(ufin, sfin, True)
    first = O.log_sorter_entry_point(orc, io, u, s, 120)
    nxt = abi.EventsClosedForm.from_buffer_copy(bytes(first[1])); nxt.start_flag = 0
    nxt.hidden_fsm_input = first[1].hidden_fsm_output  # carries a previous_item: Address / U256 strings of a real record
    w = EventsDeduplicatorInstanceWitness(nxt, u[120:], up[120:], s[120:], sp[120:])
    dump = wire.write_events_deduplicator_witness(w)
    got = wire.read_events_deduplicator_witness(dump)
    assert bytes(got.closed_form_input) == bytes(nxt)
    assert got.initial_queue_witness.tobytes() == np.ascontiguousarray(u[120:]).tobytes()
    assert got.intermediate_sorted_queue_witness.tobytes() == np.ascontiguousarray(s[120:]).tobytes()
    assert np.array_equal(got.initial_queue_prev_tails, up[120:]) and np.array_equal(got.intermediate_sorted_queue_prev_tails, sp[120:])
    assert wire.write_events_deduplicator_witness(got) == dump
    have = O.log_sorter_entry_point(orc, got.closed_form_input, got.initial_queue_witness, got.intermediate_sorted_queue_witness, 100)
    want = O.log_sorter_entry_point(orc, nxt, u[120:], s[120:], 100)
    assert have[0] == want[0] == abi.ZKC_OK and have[3].tolist() == want[3].tolist()


def test_malformed_dumps_fail_loudly(orc):
    from era_zkevm_circuits_b200 import RamPermutationCircuitInstanceWitness
    u, s = V.ram_reference_vector()
    io, up, sp = H.ram_instance(orc, u, s, nondet_len=1)
    dump = bytearray(wire.write_ram_permutation_witness(RamPermutationCircuitInstanceWitness(io, u, up, s, sp)))
    with pytest.raises(wire.WireError):
        wire.read_ram_permutation_witness(bytes(dump) + b"\x00")            # trailing bytes
    with pytest.raises(wire.WireError):
        wire.read_ram_permutation_witness(bytes(dump[:-5]))                 # truncated
    bad = bytearray(dump); bad[0] = 2
    with pytest.raises(wire.WireError):
        wire.read_ram_permutation_witness(bytes(bad))                       # bool out of range
    bad = bytearray(dump); bad[2:10] = struct.pack("<Q", abi.GL_P)
    with pytest.raises(wire.WireError):
        wire.read_ram_permutation_witness(bytes(bad))                       # non-canonical field element


def test_storage_deduplicator_dump_round_trip_and_ingestion(orc):
    """a chained (start_flag = 0) storage_validity instance: every field of the 77-element FSM record is populated; the ingested dump
    proves like the original under the oracle"""
    from era_zkevm_circuits_b200 import StorageDeduplicatorInstanceWitness
    u, s, ts = synthetic.storage_trace(300, seed=4, n_cells=9)
    up, ufin = O.log_queue_simulate(orc, u)
    sp, sfin = O.log_queue_simulate(orc, s, ts)
    io = O.storage_closed_form(ufin, sfin, 0, True)
    first = O.storage_validity_entry_point(orc, io, u, s, ts, 120)
    nxt = abi.StorageClosedForm.from_buffer_copy(bytes(first[1])); nxt.start_flag = 0
    nxt.hidden_fsm_input = first[1].hidden_fsm_output
    w = StorageDeduplicatorInstanceWitness(nxt, u[120:], up[120:], s[120:], ts[120:], sp[120:])
    dump = wire.write_storage_deduplicator_witness(w)
    assert dump[:3] == b"\x00\x00\x00"  # start_flag, completion_flag, shard id
    got = wire.read_storage_deduplicator_witness(dump)
    assert bytes(got.closed_form_input) == bytes(nxt)
    for a, b in ((got.unsorted_queue_witness, w.unsorted_queue_witness), (got.intermediate_sorted_queue_witness, w.intermediate_sorted_queue_witness)):
        assert a.tobytes() == np.ascontiguousarray(b).tobytes()
    assert np.array_equal(got.intermediate_sorted_queue_timestamps, ts[120:])
    assert np.array_equal(got.unsorted_queue_prev_tails, up[120:]) and np.array_equal(got.intermediate_sorted_queue_prev_tails, sp[120:])
    assert wire.write_storage_deduplicator_witness(got) == dump
    want = O.storage_validity_entry_point(orc, nxt, u[120:], s[120:], ts[120:], 200)
    have = O.storage_validity_entry_point(orc, got.closed_form_input, got.unsorted_queue_witness, got.intermediate_sorted_queue_witness,
                                          got.intermediate_sorted_queue_timestamps, 200)
    assert want[0] == have[0] == 0 and np.array_equal(want[3], have[3]) and np.array_equal(want[2], have[2])
    with pytest.raises(wire.WireError):
        wire.read_storage_deduplicator_witness(dump[:-3])


def test_sha256_round_function_dump_round_trip_and_ingestion(orc):
    """a chained (start_flag = 0) sha256 instance: mid-message FSM state (inner state, call parameters) populated; the ingested dump
    proves like the original under the oracle"""
    from era_zkevm_circuits_b200 import Sha256RoundFunctionCircuitInstanceWitness
    reqs, reads, msgs = synthetic.sha256_calls(12, seed=3, max_rounds=9)
    prev, rfin = O.log_queue_simulate(orc, reqs)
    io = O.sha256_closed_form(rfin)
    total = len(reads) // 2
    cut = total // 2
    first = O.sha256_entry_point(orc, io, reqs, reads, cut)
    nxt = abi.Sha256ClosedForm.from_buffer_copy(bytes(first[1])); nxt.start_flag = 0
    nxt.hidden_fsm_input = first[1].hidden_fsm_output
    used = len(reqs) - first[1].hidden_fsm_output.log_queue_state.length
    w = Sha256RoundFunctionCircuitInstanceWitness(nxt, reqs[used:], prev[used:], np.ascontiguousarray(reads, dtype=np.uint32).reshape(-1, 8)[2 * cut:])
    dump = wire.write_sha256_round_function_witness(w)
    got = wire.read_sha256_round_function_witness(dump)
    assert bytes(got.closed_form_input) == bytes(nxt)
    assert got.requests_queue_witness.tobytes() == np.ascontiguousarray(w.requests_queue_witness).tobytes()
    assert np.array_equal(got.requests_queue_prev_tails, w.requests_queue_prev_tails) and np.array_equal(got.memory_reads_witness, w.memory_reads_witness)
    assert wire.write_sha256_round_function_witness(got) == dump
    want = O.sha256_entry_point(orc, nxt, w.requests_queue_witness, w.memory_reads_witness, total + 3 - cut)
    have = O.sha256_entry_point(orc, got.closed_form_input, got.requests_queue_witness, got.memory_reads_witness, total + 3 - cut)
    assert want[0] == have[0] == 0 and np.array_equal(want[3], have[3]) and np.array_equal(want[2], have[2])
    with pytest.raises(wire.WireError):
        wire.read_sha256_round_function_witness(dump + b"\x00")


def test_keccak256_round_function_dump_round_trip_and_ingestion(orc):
    """a chained keccak256 instance cut in the middle of a call: the 200-byte internal state, the byte buffer and its fill count
    are populated; the ingested dump proves like the original under the oracle"""
    from era_zkevm_circuits_b200 import Keccak256RoundFunctionCircuitInstanceWitness
    KC = abi.KC_COLS
    reqs, reads, msgs = synthetic.keccak_calls(10, seed=3, max_len=700)
    prev, rfin = O.log_queue_simulate(orc, reqs)
    io = O.keccak_closed_form(rfin)
    limit = sum(len(m) // 136 + 1 for m in msgs) + 5
    whole = O.keccak_entry_point(orc, io, reqs, reads, limit)
    cut = int(np.flatnonzero(whole[2][KC["WRITE_RESULT"]])[4]) - 1
    first = O.keccak_entry_point(orc, io, reqs, reads, cut)
    nxt = abi.KeccakClosedForm.from_buffer_copy(bytes(first[1])); nxt.start_flag = 0
    nxt.hidden_fsm_input = first[1].hidden_fsm_output
    assert any(nxt.hidden_fsm_input.keccak_internal_state) and nxt.hidden_fsm_input.completed == 0
    used_req = len(reqs) - first[1].hidden_fsm_output.log_queue_state.length
    used_reads = int(first[2][KC["QUERY"] + 3::KC["QUERY_STRIDE"]][:6].sum())
    w = Keccak256RoundFunctionCircuitInstanceWitness(nxt, reqs[used_req:], prev[used_req:], np.ascontiguousarray(reads, dtype=np.uint32).reshape(-1, 8)[used_reads:])
    dump = wire.write_keccak256_round_function_witness(w)
    got = wire.read_keccak256_round_function_witness(dump)
    assert bytes(got.closed_form_input) == bytes(nxt)
    assert got.requests_queue_witness.tobytes() == np.ascontiguousarray(w.requests_queue_witness).tobytes()
    assert np.array_equal(got.memory_reads_witness, w.memory_reads_witness) and wire.write_keccak256_round_function_witness(got) == dump
    want = O.keccak_entry_point(orc, nxt, w.requests_queue_witness, w.memory_reads_witness, limit - cut)
    have = O.keccak_entry_point(orc, got.closed_form_input, got.requests_queue_witness, got.memory_reads_witness, limit - cut)
    assert want[0] == have[0] == 0 and np.array_equal(want[3], have[3]) and np.array_equal(want[2], have[2])
    assert bytes(have[1].hidden_fsm_output) == bytes(whole[1].hidden_fsm_output)


def test_decommit_sorter_dump_round_trip_and_ingestion(orc):
    from era_zkevm_circuits_b200 import CodeDecommittmentsDeduplicatorInstanceWitness as W
    u, s = synthetic.decommit_requests_trace(300, seed=4, n_hashes=20)
    up, ufin = O.decommit_queue_simulate(orc, u)
    sp, sfin = O.decommit_queue_simulate(orc, s)
    io = O.decommit_sorter_closed_form(ufin, sfin, True)
    first = O.sort_decommittments_entry_point(orc, io, u, s, 120)
    nxt = abi.DecommitSorterClosedForm.from_buffer_copy(bytes(first[1])); nxt.start_flag = 0
    nxt.hidden_fsm_input = first[1].hidden_fsm_output  # carries previous_record / previous_packed_key / first_encountered_timestamp
    w = W(nxt, u[120:], up[120:], s[120:], sp[120:])
    dump = wire.write_decommit_sorter_witness(w)
    got = wire.read_decommit_sorter_witness(dump)
    assert bytes(got.closed_form_input) == bytes(nxt) and wire.write_decommit_sorter_witness(got) == dump
    assert got.sorted_queue_witness.tobytes() == np.ascontiguousarray(s[120:]).tobytes() and np.array_equal(got.initial_queue_prev_states, up[120:])
    want = O.sort_decommittments_entry_point(orc, nxt, u[120:], s[120:], 200)
    have = O.sort_decommittments_entry_point(orc, got.closed_form_input, got.initial_queue_witness, got.sorted_queue_witness, 200)
    assert want[0] == have[0] == 0 and np.array_equal(want[3], have[3]) and np.array_equal(want[2], have[2])


def test_code_decommitter_dump_round_trip_and_ingestion(orc):
    from era_zkevm_circuits_b200 import CodeDecommitterCircuitInstanceWitness as W
    reqs, words = synthetic.code_decommit_requests(12, seed=4, max_words=21)
    prev, fin = O.decommit_queue_simulate(orc, reqs)
    io = O.code_unpacker_closed_form(fin, None, True)
    per_request = [int(h[7]) & 0xFFFF for h in reqs["code_hash"]]
    limit = sum((n + 1) // 2 for n in per_request) + 3
    cut = limit // 2
    first = O.code_unpacker_entry_point(orc, io, reqs, words, cut)
    nxt = abi.CodeUnpackerClosedForm.from_buffer_copy(bytes(first[1])); nxt.start_flag = 0
    nxt.hidden_fsm_input = first[1].hidden_fsm_output  # mid-bytecode: SHA-256 state, hash to compare, UInt16 round counter
    used_req = len(reqs) - first[1].hidden_fsm_output.decommittment_requests_queue_state.length
    used_words = int(first[1].hidden_fsm_output.memory_queue_state.length)
    rest = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1, 8)[used_words:]
    split = [sum(per_request[:used_req]) - used_words] + per_request[used_req:]  # the request in progress first
    w = W(nxt, reqs[used_req:], prev[used_req:], rest)
    dump = wire.write_code_decommitter_witness(w, split)
    got, got_split = wire.read_code_decommitter_witness(dump)
    assert bytes(got.closed_form_input) == bytes(nxt) and got_split == split and wire.write_code_decommitter_witness(got, got_split) == dump
    assert np.array_equal(got.code_words, rest) and got.sorted_requests_queue_witness.tobytes() == np.ascontiguousarray(reqs[used_req:]).tobytes()
    want = O.code_unpacker_entry_point(orc, nxt, reqs[used_req:], rest, limit - cut)
    have = O.code_unpacker_entry_point(orc, got.closed_form_input, got.sorted_requests_queue_witness, got.code_words, limit - cut)
    assert want[0] == have[0] == 0 and np.array_equal(want[3], have[3]) and np.array_equal(want[2], have[2])


def test_log_demuxer_and_linear_hasher_dump_round_trips(orc):
    from era_zkevm_circuits_b200 import LinearHasherCircuitInstanceWitness, LogDemuxerCircuitInstanceWitness
    recs = synthetic.vm_log_queue_trace(200, seed=6)
    prev, fin = O.log_queue_simulate(orc, recs)
    io = O.demux_closed_form(fin, True)
    first = O.demux_entry_point(orc, io, recs, 90)
    nxt = abi.DemuxClosedForm.from_buffer_copy(bytes(first[1])); nxt.start_flag = 0
    nxt.hidden_fsm_input = first[1].hidden_fsm_output  # six non-trivial output queue states
    w = LogDemuxerCircuitInstanceWitness(nxt, recs[90:], prev[90:])
    dump = wire.write_log_demuxer_witness(w)
    got = wire.read_log_demuxer_witness(dump)
    assert bytes(got.closed_form_input) == bytes(nxt) and wire.write_log_demuxer_witness(got) == dump
    want = O.demux_entry_point(orc, nxt, recs[90:], 120)
    have = O.demux_entry_point(orc, got.closed_form_input, got.initial_queue_witness, 120)
    assert want[0] == have[0] == 0 and np.array_equal(want[3], have[3]) and np.array_equal(want[2], have[2])
    msgs = recs.copy(); msgs["tx_number_in_block"] &= 0xFFFF
    mprev, mfin = O.log_queue_simulate(orc, msgs)
    lio = O.linear_hasher_closed_form(mfin)
    done = O.linear_hasher_entry_point(orc, lio, msgs, 210)
    assert done[0] == 0
    lw = LinearHasherCircuitInstanceWitness(done[1], msgs, mprev)  # the finished closed form carries the 32 digest bytes
    dump = wire.write_linear_hasher_witness(lw)
    got = wire.read_linear_hasher_witness(dump)
    assert bytes(got.closed_form_input) == bytes(done[1]) and wire.write_linear_hasher_witness(got) == dump
    assert got.queue_witness.tobytes() == msgs.tobytes() and np.array_equal(got.queue_prev_tails, mprev)
    again = O.linear_hasher_entry_point(orc, got.closed_form_input, got.queue_witness, 210)
    assert again[0] == 0 and np.array_equal(again[3], done[3])


def test_main_vm_closed_form_dump_round_trip_and_ingestion(orc):
    """VmCircuitWitness.closed_form_input (circuit_inputs/main_vm.rs:9-71): byte layout of the leading fields assembled by hand, a
    write -> read round trip of a closed form whose FSM output sits inside a far call (non-trivial context, callstack sponge, queue
    states), the oracle W handed to a caller-supplied reader, and the ingested closed form chaining the next instance exactly like
    the original"""
    from era_zkevm_circuits_b200 import isa as I
    isa = I.Isa()
    io = abi.VmClosedForm(); io.start_flag = 1
    io.memory_queue_initial_tail[11] = 5; io.memory_queue_initial_length = 7
    io.zkporter_is_available = 1
    io.default_aa_code_hash[0], io.default_aa_code_hash[7] = 0x1234, 0x0100_0000
    st0 = O.vm_initial_state(orc, io, isa.isa)
    cycles, first = 3000, 1700
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st0, I.pack_code(I.random_program(isa, 1024, seed=3, far_calls=True)), cycles, full=True)
    assert rc == 0
    for k in range(4):
        io.rollback_queue_tail_for_block[k] = int(tail[k])
    a = O.vm_entry_point(orc, io, isa.isa, snaps[:first + 1], wit[:first], first, cw=cw)
    assert a[0] == 0
    done = a[1]                                                            # closed form after the first instance (FSM output filled in)
    assert done.hidden_fsm_output.context_stack_depth >= 1 and any(done.hidden_fsm_output.stack_sponge_state)
    dump = wire.write_vm_circuit_witness(done)
    # start_flag, completion_flag, rollback tail (4 field elements), memory queue tail (12) + length u32
    assert dump[:2] == bytes([1, done.completion_flag]) and dump[2:34] == b"".join(struct.pack("<Q", int(x)) for x in tail)
    assert dump[34:34 + 96] == bytes(88) + struct.pack("<Q", 5) and dump[130:134] == struct.pack("<I", 7)
    # decommitment tail (12) + length, then GlobalContext: bool + U256 as a hex string without leading zeros
    o = 134 + 96 + 4
    aa = "0x1000000" + "0" * 48 + "00001234"
    assert dump[o:o + 1] == b"\x01" and dump[o + 1:o + 9] == struct.pack("<Q", len(aa)) and dump[o + 9:o + 9 + len(aa)] == aa.encode()
    got, oracle = wire.read_vm_circuit_witness(dump)
    assert oracle is None and wire.write_vm_circuit_witness(got) == dump
    for name, _ in abi.VmState._fields_:
        if name == "_pad":
            continue
        x, y = getattr(got.hidden_fsm_output, name), getattr(done.hidden_fsm_output, name)
        assert bytes(x) == bytes(y) if hasattr(x, "_length_") or hasattr(x, "_fields_") else x == y, name
    assert bytes(got.memory_queue_final_state) == bytes(done.memory_queue_final_state)
    # W: whatever follows the closed form belongs to the harness' oracle type
    tagged, oracle = wire.read_vm_circuit_witness(dump + b"\x07\x00\x00\x00", read_oracle=lambda r: r.u32())
    assert oracle == 7 and wire.write_vm_circuit_witness(tagged) == dump
    with pytest.raises(wire.WireError):
        wire.read_vm_circuit_witness(dump + b"\x00")                       # trailing bytes without a reader for W
    with pytest.raises(wire.WireError):
        wire.read_vm_circuit_witness(dump[:-3])
    bad = bytearray(dump); bad[0] = 2
    with pytest.raises(wire.WireError):
        wire.read_vm_circuit_witness(bytes(bad))
    # ingestion: the next instance chained from the INGESTED closed form commits like the one chained from the original
    def chained(src):
        nxt = abi.VmClosedForm.from_buffer_copy(bytes(src)); nxt.start_flag = 0
        nxt.hidden_fsm_input = src.hidden_fsm_output
        return O.vm_entry_point(orc, nxt, isa.isa, snaps[first:], wit[first:], cycles - first, cw=cw)
    want, have = chained(done), chained(got)
    assert want[0] == have[0] == 0 and have[3].tolist() == want[3].tolist() and np.array_equal(have[2], want[2])
