"""Wire format of the reference's `*CircuitInstanceWitness` structs: what `bincode::serialize` (bincode 1.x,
default options: little-endian, fixed-width integers, u64 sequence lengths) produces for the serde derives of

    RamPermutationCircuitInstanceWitness   /root/reference/src/ram_permutation/input.rs:99-116
    EventsDeduplicatorInstanceWitness      /root/reference/src/log_sorter/input.rs:98-106
    StorageDeduplicatorInstanceWitness     /root/reference/src/storage_validity_by_grand_product/input.rs:128-136
    Sha256RoundFunctionCircuitInstanceWitness /root/reference/src/sha256_round_function/input.rs:85-89
    Keccak256RoundFunctionCircuitInstanceWitness /root/reference/src/keccak256_round_function/input.rs:95-99
    CodeDecommittmentsDeduplicatorInstanceWitness /root/reference/src/sort_decommittment_requests/input.rs:110-125
    CodeDecommitterCircuitInstanceWitness  /root/reference/src/code_unpacker_sha256/input.rs:134-140
    LogDemuxerCircuitInstanceWitness       /root/reference/src/demux_log_queue/input.rs:118-121
    LinearHasherCircuitInstanceWitness     /root/reference/src/linear_hasher/input.rs:71-80
    VmCircuitWitness (its closed_form_input; the witness oracle W is a type the HARNESS defines, see read_vm_circuit_witness)
                                           /root/reference/src/fsm_input_output/circuit_inputs/main_vm.rs:9-71

read into the host-side witness forms of this package (closed-form struct + struct-of-arrays queue witnesses), and written
back (test_harness-style dumps for the round-trip tests).

Field order = declaration order of the structs (serde derive); `()` occupies no bytes; `bool` is one byte; a field element
(GoldilocksField, a newtype over u64) is 8 bytes; fixed arrays `[T; N]` are tuples (no length prefix; boojum's BigArraySerde
for N > 32 likewise); `VecDeque` is a u64 length + elements; a tuple is its members back to back.  `U256` / `Address`
(ethereum_types, serde through impl-serde) are STRINGS: u64 length + "0x" + hex, U256 without leading zeros ("0x0" for zero),
H160 always 40 digits.

The layouts of the witness structs boojum derives (QueueStateWitness { head, tail: QueueTailStateWitness { tail, length } },
`(item_witness, previous_tail)` deque elements) follow the published boojum sources; boojum is not vendored in
/root/reference, so no dump of the reference itself is available here to pin this reader against: PARITY UNPINNED, like the
Poseidon2 constants.  The reader fails loudly (WireError) on trailing bytes, non-canonical field elements, bad booleans and
malformed hex strings, so a layout drift shows up as an error, not as a silently wrong witness.
"""
import struct
from dataclasses import dataclass

import numpy as np

from . import abi


class WireError(ValueError):
    pass


class Reader:
    def __init__(self, data: bytes):
        self.d, self.o = memoryview(data), 0

    def take(self, n):
        if self.o + n > len(self.d):
            raise WireError(f"unexpected end of input at byte {self.o} (+{n})")
        b = self.d[self.o:self.o + n]
        self.o += n
        return b

    def u8(self):
        return self.take(1)[0]

    def u16(self):
        return struct.unpack("<H", self.take(2))[0]

    def u32(self):
        return struct.unpack("<I", self.take(4))[0]

    def u64(self):
        return struct.unpack("<Q", self.take(8))[0]

    def boolean(self):
        v = self.u8()
        if v > 1:
            raise WireError(f"invalid bool {v} at byte {self.o - 1}")
        return v

    def field(self):
        v = self.u64()
        if v >= abi.GL_P:
            raise WireError(f"non-canonical field element {v:#x} at byte {self.o - 8}")
        return v

    def fields(self, n):
        return [self.field() for _ in range(n)]

    def hex_string(self, max_digits, exact=False):
        n = self.u64()
        if n < 3 or n > 2 + max_digits:
            raise WireError(f"hex string of length {n} at byte {self.o - 8}")
        s = bytes(self.take(n)).decode("ascii", "replace")
        if not s.startswith("0x") or (exact and n != 2 + max_digits):
            raise WireError(f"malformed hex string {s!r}")
        try:
            return int(s[2:], 16)
        except ValueError:
            raise WireError(f"malformed hex string {s!r}") from None

    def u256(self):
        return self.hex_string(64)

    def h160(self):
        return self.hex_string(40, exact=True)

    def done(self):
        if self.o != len(self.d):
            raise WireError(f"{len(self.d) - self.o} trailing bytes")


class Writer:
    def __init__(self):
        self.b = bytearray()

    def u8(self, v): self.b.append(int(v) & 0xFF)
    def u16(self, v): self.b += struct.pack("<H", int(v))
    def u32(self, v): self.b += struct.pack("<I", int(v))
    def u64(self, v): self.b += struct.pack("<Q", int(v))
    def boolean(self, v): self.u8(1 if v else 0)
    def field(self, v): self.u64(v)

    def fields(self, vs):
        for v in vs:
            self.u64(v)

    def string(self, s):
        raw = s.encode("ascii")
        self.u64(len(raw))
        self.b += raw

    def u256(self, v): self.string(hex(int(v)))
    def h160(self, v): self.string("0x%040x" % int(v))


def _limbs(v, n):
    return [(int(v) >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def _from_limbs(limbs):
    return sum(int(x) << (32 * i) for i, x in enumerate(limbs))


# ---- QueueState<F, N> ------------------------------------------------------------------------------------------------------
def _read_queue_state(r: Reader, q):
    n = len(q.head)
    for i, v in enumerate(r.fields(n)):
        q.head[i] = v
    for i, v in enumerate(r.fields(n)):
        q.tail[i] = v
    q.length = r.u32()


def _write_queue_state(w: Writer, q):
    w.fields(q.head)
    w.fields(q.tail)
    w.u32(q.length)


# ---- MemoryQuery / LogQuery witnesses --------------------------------------------------------------------------------------
def _read_memory_query(r: Reader, rec):
    rec["timestamp"], rec["memory_page"], rec["index"] = r.u32(), r.u32(), r.u32()
    rec["rw_flag"], rec["is_ptr"] = r.boolean(), r.boolean()
    rec["value"] = _limbs(r.u256(), 8)


def _write_memory_query(w: Writer, rec):
    w.u32(rec["timestamp"]); w.u32(rec["memory_page"]); w.u32(rec["index"])
    w.boolean(rec["rw_flag"]); w.boolean(rec["is_ptr"])
    w.u256(_from_limbs(rec["value"]))


def _read_log_query(r: Reader):
    """-> (address[5], key[8], read_value[8], written_value[8], tx_number, timestamp, flags) as a LOG_QUERY_DTYPE tuple"""
    address, key, rv, wv = _limbs(r.h160(), 5), _limbs(r.u256(), 8), _limbs(r.u256(), 8), _limbs(r.u256(), 8)
    aux, rw, rollback, service, shard = r.u8(), r.boolean(), r.boolean(), r.boolean(), r.u8()
    tx, ts = r.u32(), r.u32()
    return address, key, rv, wv, tx, ts, abi.lq_flags(aux, shard, rw, rollback, service)


def _write_log_query(w: Writer, rec):
    fl = int(rec["flags"])
    w.h160(_from_limbs(rec["address"])); w.u256(_from_limbs(rec["key"])); w.u256(_from_limbs(rec["read_value"]))
    w.u256(_from_limbs(rec["written_value"]))
    w.u8(fl & 0xFF); w.boolean((fl >> 16) & 1); w.boolean((fl >> 17) & 1); w.boolean((fl >> 18) & 1); w.u8((fl >> 8) & 0xFF)
    w.u32(rec["tx_number_in_block"]); w.u32(rec["timestamp"])


def _set_log_query_struct(dst, t):
    address, key, rv, wv, tx, ts, flags = t
    for i in range(5):
        dst.address[i] = address[i]
    for i in range(8):
        dst.key[i], dst.read_value[i], dst.written_value[i] = key[i], rv[i], wv[i]
    dst.tx_number_in_block, dst.timestamp, dst.flags = tx, ts, flags


def _log_query_struct_as_record(src):
    rec = np.zeros((), dtype=abi.LOG_QUERY_DTYPE)
    rec["address"], rec["key"] = list(src.address), list(src.key)
    rec["read_value"], rec["written_value"] = list(src.read_value), list(src.written_value)
    rec["tx_number_in_block"], rec["timestamp"], rec["flags"] = src.tx_number_in_block, src.timestamp, src.flags
    return rec


# ---- ram_permutation ---------------------------------------------------------------------------------------------------------
def _read_ram_fsm(r: Reader, f):
    for name in ("lhs_accumulator", "rhs_accumulator"):
        for i, v in enumerate(r.fields(2)):
            getattr(f, name)[i] = v
    _read_queue_state(r, f.current_unsorted_queue_state)
    _read_queue_state(r, f.current_sorted_queue_state)
    for i in range(3):
        f.previous_sorting_key[i] = r.u32()
    for i in range(2):
        f.previous_full_key[i] = r.u32()
    for i, v in enumerate(_limbs(r.u256(), 8)):
        f.previous_value[i] = v
    f.previous_is_ptr = r.boolean()
    f.num_nondeterministic_writes = r.u32()


def _write_ram_fsm(w: Writer, f):
    w.fields(f.lhs_accumulator); w.fields(f.rhs_accumulator)
    _write_queue_state(w, f.current_unsorted_queue_state)
    _write_queue_state(w, f.current_sorted_queue_state)
    for v in f.previous_sorting_key:
        w.u32(v)
    for v in f.previous_full_key:
        w.u32(v)
    w.u256(_from_limbs(f.previous_value))
    w.boolean(f.previous_is_ptr)
    w.u32(f.num_nondeterministic_writes)


def _read_memory_queue(r: Reader):
    n = r.u64()
    if n > (len(r.d) - r.o) // 8:
        raise WireError(f"queue witness claims {n} elements")
    recs = np.zeros(n, dtype=abi.MEMORY_QUERY_DTYPE)
    prev = np.zeros((n, 12), dtype=np.uint64)
    for k in range(n):
        _read_memory_query(r, recs[k])
        prev[k] = r.fields(12)
    return recs, prev


def _write_memory_queue(w: Writer, recs, prev):
    w.u64(len(recs))
    for rec, p in zip(recs, prev):
        _write_memory_query(w, rec)
        w.fields(p)


def read_ram_permutation_witness(data: bytes):
    """bincode bytes of RamPermutationCircuitInstanceWitness<GoldilocksField> -> ram_permutation.RamPermutationCircuitInstanceWitness"""
    from .ram_permutation import RamPermutationCircuitInstanceWitness
    r = Reader(data)
    io = abi.RamClosedForm()
    io.start_flag, io.completion_flag = r.boolean(), r.boolean()
    _read_queue_state(r, io.observable_input.unsorted_queue_initial_state)
    _read_queue_state(r, io.observable_input.sorted_queue_initial_state)
    io.observable_input.non_deterministic_bootloader_memory_snapshot_length = r.u32()
    # observable_output: `()`
    _read_ram_fsm(r, io.hidden_fsm_input)
    _read_ram_fsm(r, io.hidden_fsm_output)
    u, up = _read_memory_queue(r)
    s, sp = _read_memory_queue(r)
    r.done()
    return RamPermutationCircuitInstanceWitness(io, u, up, s, sp)


def write_ram_permutation_witness(w_) -> bytes:
    w = Writer()
    io = w_.closed_form_input
    w.boolean(io.start_flag); w.boolean(io.completion_flag)
    _write_queue_state(w, io.observable_input.unsorted_queue_initial_state)
    _write_queue_state(w, io.observable_input.sorted_queue_initial_state)
    w.u32(io.observable_input.non_deterministic_bootloader_memory_snapshot_length)
    _write_ram_fsm(w, io.hidden_fsm_input)
    _write_ram_fsm(w, io.hidden_fsm_output)
    _write_memory_queue(w, w_.unsorted_queue_witness, w_.unsorted_queue_prev_states)
    _write_memory_queue(w, w_.sorted_queue_witness, w_.sorted_queue_prev_states)
    return bytes(w.b)


# ---- log_sorter --------------------------------------------------------------------------------------------------------------
def _read_events_fsm(r: Reader, f):
    for name in ("lhs_accumulator", "rhs_accumulator"):
        for i, v in enumerate(r.fields(2)):
            getattr(f, name)[i] = v
    _read_queue_state(r, f.initial_unsorted_queue_state)
    _read_queue_state(r, f.intermediate_sorted_queue_state)
    _read_queue_state(r, f.final_result_queue_state)
    f.previous_key = r.u32()
    _set_log_query_struct(f.previous_item, _read_log_query(r))


def _write_events_fsm(w: Writer, f):
    w.fields(f.lhs_accumulator); w.fields(f.rhs_accumulator)
    _write_queue_state(w, f.initial_unsorted_queue_state)
    _write_queue_state(w, f.intermediate_sorted_queue_state)
    _write_queue_state(w, f.final_result_queue_state)
    w.u32(f.previous_key)
    _write_log_query(w, _log_query_struct_as_record(f.previous_item))


def _read_log_queue(r: Reader):
    n = r.u64()
    if n > (len(r.d) - r.o) // 8:
        raise WireError(f"queue witness claims {n} elements")
    recs = np.zeros(n, dtype=abi.LOG_QUERY_DTYPE)
    prev = np.zeros((n, 4), dtype=np.uint64)
    for k in range(n):
        recs[k] = _read_log_query(r)
        prev[k] = r.fields(4)
    return recs, prev


def _write_log_queue(w: Writer, recs, prev):
    w.u64(len(recs))
    for rec, p in zip(recs, prev):
        _write_log_query(w, rec)
        w.fields(p)


def read_events_deduplicator_witness(data: bytes):
    """bincode bytes of EventsDeduplicatorInstanceWitness<GoldilocksField> -> log_sorter.EventsDeduplicatorInstanceWitness"""
    from .log_sorter import EventsDeduplicatorInstanceWitness
    r = Reader(data)
    io = abi.EventsClosedForm()
    io.start_flag, io.completion_flag = r.boolean(), r.boolean()
    _read_queue_state(r, io.initial_log_queue_state)
    _read_queue_state(r, io.intermediate_sorted_queue_state)
    _read_queue_state(r, io.final_queue_state)
    _read_events_fsm(r, io.hidden_fsm_input)
    _read_events_fsm(r, io.hidden_fsm_output)
    u, up = _read_log_queue(r)
    s, sp = _read_log_queue(r)
    r.done()
    return EventsDeduplicatorInstanceWitness(io, u, up, s, sp)


def write_events_deduplicator_witness(w_) -> bytes:
    w = Writer()
    io = w_.closed_form_input
    w.boolean(io.start_flag); w.boolean(io.completion_flag)
    _write_queue_state(w, io.initial_log_queue_state)
    _write_queue_state(w, io.intermediate_sorted_queue_state)
    _write_queue_state(w, io.final_queue_state)
    _write_events_fsm(w, io.hidden_fsm_input)
    _write_events_fsm(w, io.hidden_fsm_output)
    _write_log_queue(w, w_.initial_queue_witness, w_.initial_queue_prev_tails)
    _write_log_queue(w, w_.intermediate_sorted_queue_witness, w_.intermediate_sorted_queue_prev_tails)
    return bytes(w.b)


# ---- storage_validity_by_grand_product ---------------------------------------------------------------------------------------
def _read_storage_fsm(r: Reader, f):
    """StorageDeduplicatorFSMInputOutput, storage_validity_by_grand_product/input.rs:37-52"""
    for name in ("lhs_accumulator", "rhs_accumulator"):
        for i, v in enumerate(r.fields(2)):
            getattr(f, name)[i] = v
    _read_queue_state(r, f.current_unsorted_queue_state)
    _read_queue_state(r, f.current_intermediate_sorted_queue_state)
    _read_queue_state(r, f.current_final_sorted_queue_state)
    f.cycle_idx = r.u32()
    for i in range(13):
        f.previous_packed_key[i] = r.u32()
    for i, v in enumerate(_limbs(r.u256(), 8)):
        f.previous_key[i] = v
    for i, v in enumerate(_limbs(r.h160(), 5)):
        f.previous_address[i] = v
    f.previous_timestamp = r.u32()
    f.this_cell_has_explicit_read_and_rollback_depth_zero = r.boolean()
    for i, v in enumerate(_limbs(r.u256(), 8)):
        f.this_cell_base_value[i] = v
    for i, v in enumerate(_limbs(r.u256(), 8)):
        f.this_cell_current_value[i] = v
    f.this_cell_current_depth = r.u32()


def _write_storage_fsm(w: Writer, f):
    w.fields(f.lhs_accumulator); w.fields(f.rhs_accumulator)
    _write_queue_state(w, f.current_unsorted_queue_state)
    _write_queue_state(w, f.current_intermediate_sorted_queue_state)
    _write_queue_state(w, f.current_final_sorted_queue_state)
    w.u32(f.cycle_idx)
    for v in f.previous_packed_key:
        w.u32(v)
    w.u256(_from_limbs(f.previous_key)); w.h160(_from_limbs(f.previous_address)); w.u32(f.previous_timestamp)
    w.boolean(f.this_cell_has_explicit_read_and_rollback_depth_zero)
    w.u256(_from_limbs(f.this_cell_base_value)); w.u256(_from_limbs(f.this_cell_current_value)); w.u32(f.this_cell_current_depth)


def read_storage_deduplicator_witness(data: bytes):
    """bincode bytes of StorageDeduplicatorInstanceWitness<GoldilocksField> (input.rs:128-136) ->
    storage_validity.StorageDeduplicatorInstanceWitness.  The sorted queue's elements are TimestampedStorageLogRecord witnesses
    (mod.rs:63-66): the LogQuery witness followed by its u32 timestamp, then the previous tail."""
    from .storage_validity import StorageDeduplicatorInstanceWitness
    r = Reader(data)
    io = abi.StorageClosedForm()
    io.start_flag, io.completion_flag = r.boolean(), r.boolean()
    io.shard_id_to_process = r.u8()
    _read_queue_state(r, io.unsorted_log_queue_state)
    _read_queue_state(r, io.intermediate_sorted_queue_state)
    _read_queue_state(r, io.final_sorted_queue_state)
    _read_storage_fsm(r, io.hidden_fsm_input)
    _read_storage_fsm(r, io.hidden_fsm_output)
    u, up = _read_log_queue(r)
    n = r.u64()
    if n > (len(r.d) - r.o) // 8:
        raise WireError(f"queue witness claims {n} elements")
    s = np.zeros(n, dtype=abi.LOG_QUERY_DTYPE)
    ts = np.zeros(n, dtype=np.uint32)
    sp = np.zeros((n, 4), dtype=np.uint64)
    for k in range(n):
        s[k] = _read_log_query(r)
        ts[k] = r.u32()
        sp[k] = r.fields(4)
    r.done()
    return StorageDeduplicatorInstanceWitness(io, u, up, s, ts, sp)


def write_storage_deduplicator_witness(w_) -> bytes:
    w = Writer()
    io = w_.closed_form_input
    w.boolean(io.start_flag); w.boolean(io.completion_flag)
    w.u8(io.shard_id_to_process)
    _write_queue_state(w, io.unsorted_log_queue_state)
    _write_queue_state(w, io.intermediate_sorted_queue_state)
    _write_queue_state(w, io.final_sorted_queue_state)
    _write_storage_fsm(w, io.hidden_fsm_input)
    _write_storage_fsm(w, io.hidden_fsm_output)
    _write_log_queue(w, w_.unsorted_queue_witness, w_.unsorted_queue_prev_tails)
    w.u64(len(w_.intermediate_sorted_queue_witness))
    for rec, t, p in zip(w_.intermediate_sorted_queue_witness, w_.intermediate_sorted_queue_timestamps, w_.intermediate_sorted_queue_prev_tails):
        _write_log_query(w, rec)
        w.u32(t)
        w.fields(p)
    return bytes(w.b)


# ---- sha256_round_function -----------------------------------------------------------------------------------------------------
def _read_sha256_fsm(r: Reader, f):
    """Sha256RoundFunctionFSMInputOutput (sha256_round_function/input.rs:24-57): internal_fsm, then the two queue states"""
    f.read_precompile_call, f.read_words_for_round, f.completed = r.boolean(), r.boolean(), r.boolean()
    for i in range(8):
        f.sha256_inner_state[i] = r.u32()
    f.timestamp_to_use_for_read, f.timestamp_to_use_for_write = r.u32(), r.u32()
    f.input_page, f.input_offset, f.output_page, f.output_offset, f.num_rounds = r.u32(), r.u32(), r.u32(), r.u32(), r.u32()  # mod.rs:44-50
    _read_queue_state(r, f.log_queue_state)
    _read_queue_state(r, f.memory_queue_state)


def _write_sha256_fsm(w: Writer, f):
    w.boolean(f.read_precompile_call); w.boolean(f.read_words_for_round); w.boolean(f.completed)
    for v in f.sha256_inner_state:
        w.u32(v)
    for v in (f.timestamp_to_use_for_read, f.timestamp_to_use_for_write, f.input_page, f.input_offset, f.output_page, f.output_offset, f.num_rounds):
        w.u32(v)
    _write_queue_state(w, f.log_queue_state)
    _write_queue_state(w, f.memory_queue_state)


def read_sha256_round_function_witness(data: bytes):
    """bincode bytes of Sha256RoundFunctionCircuitInstanceWitness<GoldilocksField> (input.rs:85-89) ->
    sha256_round_function.Sha256RoundFunctionCircuitInstanceWitness; memory_reads_witness (VecDeque<U256>) becomes [n, 8] u32 limbs"""
    from .sha256_round_function import Sha256RoundFunctionCircuitInstanceWitness
    r = Reader(data)
    io = abi.Sha256ClosedForm()
    io.start_flag, io.completion_flag = r.boolean(), r.boolean()
    _read_queue_state(r, io.initial_log_queue_state)
    _read_queue_state(r, io.initial_memory_queue_state)
    _read_queue_state(r, io.final_memory_state)
    _read_sha256_fsm(r, io.hidden_fsm_input)
    _read_sha256_fsm(r, io.hidden_fsm_output)
    reqs, prev = _read_log_queue(r)
    n = r.u64()
    if n > (len(r.d) - r.o) // 11:  # a U256 string is at least 8 + 3 bytes
        raise WireError(f"memory_reads_witness claims {n} elements")
    reads = np.zeros((n, 8), dtype=np.uint32)
    for k in range(n):
        reads[k] = _limbs(r.u256(), 8)
    r.done()
    return Sha256RoundFunctionCircuitInstanceWitness(io, reqs, prev, reads)


def write_sha256_round_function_witness(w_) -> bytes:
    w = Writer()
    io = w_.closed_form_input
    w.boolean(io.start_flag); w.boolean(io.completion_flag)
    _write_queue_state(w, io.initial_log_queue_state)
    _write_queue_state(w, io.initial_memory_queue_state)
    _write_queue_state(w, io.final_memory_state)
    _write_sha256_fsm(w, io.hidden_fsm_input)
    _write_sha256_fsm(w, io.hidden_fsm_output)
    _write_log_queue(w, w_.requests_queue_witness, w_.requests_queue_prev_tails)
    w.u64(len(w_.memory_reads_witness))
    for word in w_.memory_reads_witness:
        w.u256(_from_limbs(word))
    return bytes(w.b)


# ---- keccak256_round_function --------------------------------------------------------------------------------------------------
def _read_keccak_fsm(r: Reader, f):
    """Keccak256RoundFunctionFSMInputOutput (keccak256_round_function/input.rs:29-67): internal_fsm { 4 flags, the 5 x 5 x 8 byte
    state (nested fixed arrays: 200 bytes, no prefixes), 2 timestamps, Keccak256PrecompileCallParams (mod.rs:48-55), ByteBuffer
    { bytes [u8; 192] (BigArraySerde: a tuple), filled: u8 } (buffer/mod.rs:8-11) }, then the two queue states"""
    f.read_precompile_call, f.read_unaligned_words_for_round, f.padding_round, f.completed = r.boolean(), r.boolean(), r.boolean(), r.boolean()
    for i, b in enumerate(bytes(r.take(200))):
        f.keccak_internal_state[i] = b
    f.timestamp_to_use_for_read, f.timestamp_to_use_for_write = r.u32(), r.u32()
    f.input_page, f.input_memory_byte_offset, f.input_memory_byte_length, f.output_page, f.output_word_offset = r.u32(), r.u32(), r.u32(), r.u32(), r.u32()
    f.needs_full_padding_round = r.boolean()
    for i, b in enumerate(bytes(r.take(192))):
        f.buffer_bytes[i] = b
    f.buffer_filled = r.u8()
    _read_queue_state(r, f.log_queue_state)
    _read_queue_state(r, f.memory_queue_state)


def _write_keccak_fsm(w: Writer, f):
    for v in (f.read_precompile_call, f.read_unaligned_words_for_round, f.padding_round, f.completed):
        w.boolean(v)
    w.b += bytes(f.keccak_internal_state)
    for v in (f.timestamp_to_use_for_read, f.timestamp_to_use_for_write, f.input_page, f.input_memory_byte_offset, f.input_memory_byte_length,
              f.output_page, f.output_word_offset):
        w.u32(v)
    w.boolean(f.needs_full_padding_round)
    w.b += bytes(f.buffer_bytes)
    w.u8(f.buffer_filled)
    _write_queue_state(w, f.log_queue_state)
    _write_queue_state(w, f.memory_queue_state)


def read_keccak256_round_function_witness(data: bytes):
    """bincode bytes of Keccak256RoundFunctionCircuitInstanceWitness<GoldilocksField> (input.rs:95-99) ->
    keccak256_round_function.Keccak256RoundFunctionCircuitInstanceWitness; memory_reads_witness (VecDeque<U256>) becomes [n, 8] u32 limbs"""
    from .keccak256_round_function import Keccak256RoundFunctionCircuitInstanceWitness
    r = Reader(data)
    io = abi.KeccakClosedForm()
    io.start_flag, io.completion_flag = r.boolean(), r.boolean()
    _read_queue_state(r, io.initial_log_queue_state)
    _read_queue_state(r, io.initial_memory_queue_state)
    _read_queue_state(r, io.final_memory_state)
    _read_keccak_fsm(r, io.hidden_fsm_input)
    _read_keccak_fsm(r, io.hidden_fsm_output)
    reqs, prev = _read_log_queue(r)
    n = r.u64()
    if n > (len(r.d) - r.o) // 11:
        raise WireError(f"memory_reads_witness claims {n} elements")
    reads = np.zeros((n, 8), dtype=np.uint32)
    for k in range(n):
        reads[k] = _limbs(r.u256(), 8)
    r.done()
    return Keccak256RoundFunctionCircuitInstanceWitness(io, reqs, prev, reads)


def write_keccak256_round_function_witness(w_) -> bytes:
    w = Writer()
    io = w_.closed_form_input
    w.boolean(io.start_flag); w.boolean(io.completion_flag)
    _write_queue_state(w, io.initial_log_queue_state)
    _write_queue_state(w, io.initial_memory_queue_state)
    _write_queue_state(w, io.final_memory_state)
    _write_keccak_fsm(w, io.hidden_fsm_input)
    _write_keccak_fsm(w, io.hidden_fsm_output)
    _write_log_queue(w, w_.requests_queue_witness, w_.requests_queue_prev_tails)
    w.u64(len(w_.memory_reads_witness))
    for word in w_.memory_reads_witness:
        w.u256(_from_limbs(word))
    return bytes(w.b)


# ---- DecommitQuery witnesses: sort_decommittment_requests, code_unpacker_sha256 ------------------------------------------------------
def _read_decommit_query(r: Reader):
    """DecommitQuery witness (base_structures/decommit_query/mod.rs:22-27) as a DECOMMIT_QUERY_DTYPE tuple"""
    code_hash = _limbs(r.u256(), 8)
    page, is_first, ts = r.u32(), r.boolean(), r.u32()
    return code_hash, page, is_first, ts, 0


def _write_decommit_query(w: Writer, rec):
    w.u256(_from_limbs(rec["code_hash"])); w.u32(rec["page"]); w.boolean(rec["is_first"]); w.u32(rec["timestamp"])


def _read_decommit_queue(r: Reader):
    """FullStateCircuitQueueRawWitness<DecommitQuery>: u64 count, then (item, previous 12-element state)"""
    n = r.u64()
    if n > (len(r.d) - r.o) // 8:
        raise WireError(f"queue witness claims {n} elements")
    recs = np.zeros(n, dtype=abi.DECOMMIT_QUERY_DTYPE)
    prev = np.zeros((n, 12), dtype=np.uint64)
    for k in range(n):
        recs[k] = _read_decommit_query(r)
        prev[k] = r.fields(12)
    return recs, prev


def _write_decommit_queue(w: Writer, recs, prev):
    w.u64(len(recs))
    for rec, p in zip(recs, prev):
        _write_decommit_query(w, rec)
        w.fields(p)


def _read_decommit_sorter_fsm(r: Reader, f):
    """CodeDecommittmentsDeduplicatorFSMInputOutput, sort_decommittment_requests/input.rs:26-37"""
    _read_queue_state(r, f.initial_queue_state); _read_queue_state(r, f.sorted_queue_state); _read_queue_state(r, f.final_queue_state)
    for name in ("lhs_accumulator", "rhs_accumulator"):
        for i, v in enumerate(r.fields(2)):
            getattr(f, name)[i] = v
    for i in range(9):
        f.previous_packed_key[i] = r.u32()
    f.first_encountered_timestamp = r.u32()
    code_hash, page, is_first, ts, _ = _read_decommit_query(r)
    for i in range(8):
        f.previous_record.code_hash[i] = code_hash[i]
    f.previous_record.page, f.previous_record.is_first, f.previous_record.timestamp = page, is_first, ts


def _write_decommit_sorter_fsm(w: Writer, f):
    _write_queue_state(w, f.initial_queue_state); _write_queue_state(w, f.sorted_queue_state); _write_queue_state(w, f.final_queue_state)
    w.fields(f.lhs_accumulator); w.fields(f.rhs_accumulator)
    for v in f.previous_packed_key:
        w.u32(v)
    w.u32(f.first_encountered_timestamp)
    p = f.previous_record
    w.u256(_from_limbs(p.code_hash)); w.u32(p.page); w.boolean(p.is_first); w.u32(p.timestamp)


def read_decommit_sorter_witness(data: bytes):
    """bincode bytes of CodeDecommittmentsDeduplicatorInstanceWitness<GoldilocksField> (input.rs:110-125)"""
    from .sort_decommittment_requests import CodeDecommittmentsDeduplicatorInstanceWitness
    r = Reader(data)
    io = abi.DecommitSorterClosedForm()
    io.start_flag, io.completion_flag = r.boolean(), r.boolean()
    _read_queue_state(r, io.initial_queue_state); _read_queue_state(r, io.sorted_queue_initial_state)
    _read_queue_state(r, io.final_queue_state)
    _read_decommit_sorter_fsm(r, io.hidden_fsm_input); _read_decommit_sorter_fsm(r, io.hidden_fsm_output)
    u, up = _read_decommit_queue(r)
    s, sp = _read_decommit_queue(r)
    r.done()
    return CodeDecommittmentsDeduplicatorInstanceWitness(io, u, up, s, sp)


def write_decommit_sorter_witness(w_) -> bytes:
    w = Writer()
    io = w_.closed_form_input
    w.boolean(io.start_flag); w.boolean(io.completion_flag)
    _write_queue_state(w, io.initial_queue_state); _write_queue_state(w, io.sorted_queue_initial_state)
    _write_queue_state(w, io.final_queue_state)
    _write_decommit_sorter_fsm(w, io.hidden_fsm_input); _write_decommit_sorter_fsm(w, io.hidden_fsm_output)
    _write_decommit_queue(w, w_.initial_queue_witness, w_.initial_queue_prev_states)
    _write_decommit_queue(w, w_.sorted_queue_witness, w_.sorted_queue_prev_states)
    return bytes(w.b)


def _read_code_unpacker_fsm(r: Reader, f):
    """CodeDecommitterFSMInputOutput (code_unpacker_sha256/input.rs:23-65): internal_fsm, then the two queue states"""
    s = f.internal_fsm
    for i in range(8):
        s.sha256_inner_state[i] = r.u32()
    for i, v in enumerate(_limbs(r.u256(), 8)):
        s.hash_to_compare_against[i] = v
    s.current_index, s.current_page, s.timestamp = r.u32(), r.u32(), r.u32()
    s.num_rounds_left = struct.unpack("<H", r.take(2))[0]  # UInt16
    s.length_in_bits = r.u32()
    s.state_get_from_queue, s.state_decommit, s.finished = r.boolean(), r.boolean(), r.boolean()
    _read_queue_state(r, f.decommittment_requests_queue_state)
    _read_queue_state(r, f.memory_queue_state)


def _write_code_unpacker_fsm(w: Writer, f):
    s = f.internal_fsm
    for v in s.sha256_inner_state:
        w.u32(v)
    w.u256(_from_limbs(s.hash_to_compare_against))
    w.u32(s.current_index); w.u32(s.current_page); w.u32(s.timestamp)
    w.b += struct.pack("<H", int(s.num_rounds_left))
    w.u32(s.length_in_bits)
    w.boolean(s.state_get_from_queue); w.boolean(s.state_decommit); w.boolean(s.finished)
    _write_queue_state(w, f.decommittment_requests_queue_state)
    _write_queue_state(w, f.memory_queue_state)


def read_code_decommitter_witness(data: bytes):
    """bincode bytes of CodeDecommitterCircuitInstanceWitness<GoldilocksField> (input.rs:134-140); code_words (Vec<Vec<U256>>, one inner
    vector per request) is flattened into [total_words, 8] u32 limbs in pop order.  Returns (witness, words per request)."""
    from .code_unpacker_sha256 import CodeDecommitterCircuitInstanceWitness
    r = Reader(data)
    io = abi.CodeUnpackerClosedForm()
    io.start_flag, io.completion_flag = r.boolean(), r.boolean()
    _read_queue_state(r, io.memory_queue_initial_state); _read_queue_state(r, io.sorted_requests_queue_initial_state)
    _read_queue_state(r, io.memory_queue_final_state)
    _read_code_unpacker_fsm(r, io.hidden_fsm_input); _read_code_unpacker_fsm(r, io.hidden_fsm_output)
    reqs, prev = _read_decommit_queue(r)
    n_outer = r.u64()
    if n_outer > (len(r.d) - r.o) // 8:
        raise WireError(f"code_words claims {n_outer} vectors")
    words, per_request = [], []
    for _ in range(n_outer):
        n = r.u64()
        if n > (len(r.d) - r.o) // 11:
            raise WireError(f"a code vector claims {n} words")
        per_request.append(n)
        words += [_limbs(r.u256(), 8) for _ in range(n)]
    r.done()
    arr = np.array(words, dtype=np.uint32).reshape(-1, 8)
    return CodeDecommitterCircuitInstanceWitness(io, reqs, prev, arr), per_request


def write_code_decommitter_witness(w_, words_per_request) -> bytes:
    """words_per_request: how code_words splits over the requests in pop order (a request in progress on entry comes first)"""
    w = Writer()
    io = w_.closed_form_input
    w.boolean(io.start_flag); w.boolean(io.completion_flag)
    _write_queue_state(w, io.memory_queue_initial_state); _write_queue_state(w, io.sorted_requests_queue_initial_state)
    _write_queue_state(w, io.memory_queue_final_state)
    _write_code_unpacker_fsm(w, io.hidden_fsm_input); _write_code_unpacker_fsm(w, io.hidden_fsm_output)
    _write_decommit_queue(w, w_.sorted_requests_queue_witness, w_.sorted_requests_queue_prev_states)
    assert sum(words_per_request) == len(w_.code_words)
    w.u64(len(words_per_request))
    k = 0
    for n in words_per_request:
        w.u64(n)
        for word in w_.code_words[k:k + n]:
            w.u256(_from_limbs(word))
        k += n
    return bytes(w.b)


# ---- demux_log_queue, linear_hasher ------------------------------------------------------------------------------------------------
def _read_demux_fsm(r: Reader, f):
    """LogDemuxerFSMInputOutput, demux_log_queue/input.rs:24-32: the input queue, then storage / events / l1 messages / keccak256 /
    sha256 / ecrecover"""
    _read_queue_state(r, f.initial_log_queue_state)
    for q in range(6):
        _read_queue_state(r, f.output_queue_states[q])


def _write_demux_fsm(w: Writer, f):
    _write_queue_state(w, f.initial_log_queue_state)
    for q in range(6):
        _write_queue_state(w, f.output_queue_states[q])


def read_log_demuxer_witness(data: bytes):
    """bincode bytes of LogDemuxerCircuitInstanceWitness<GoldilocksField> (input.rs:118-121)"""
    from .demux_log_queue import LogDemuxerCircuitInstanceWitness
    r = Reader(data)
    io = abi.DemuxClosedForm()
    io.start_flag, io.completion_flag = r.boolean(), r.boolean()
    _read_queue_state(r, io.initial_log_queue_state)
    for q in range(6):
        _read_queue_state(r, io.output_queue_states[q])
    _read_demux_fsm(r, io.hidden_fsm_input); _read_demux_fsm(r, io.hidden_fsm_output)
    recs, prev = _read_log_queue(r)
    r.done()
    return LogDemuxerCircuitInstanceWitness(io, recs, prev)


def write_log_demuxer_witness(w_) -> bytes:
    w = Writer()
    io = w_.closed_form_input
    w.boolean(io.start_flag); w.boolean(io.completion_flag)
    _write_queue_state(w, io.initial_log_queue_state)
    for q in range(6):
        _write_queue_state(w, io.output_queue_states[q])
    _write_demux_fsm(w, io.hidden_fsm_input); _write_demux_fsm(w, io.hidden_fsm_output)
    _write_log_queue(w, w_.initial_queue_witness, w_.initial_queue_prev_tails)
    return bytes(w.b)


def read_linear_hasher_witness(data: bytes):
    """bincode bytes of LinearHasherCircuitInstanceWitness<GoldilocksField> (input.rs:71-80): the hidden FSM is `()` (no bytes), the
    observable output 32 digest bytes"""
    from .linear_hasher import LinearHasherCircuitInstanceWitness
    r = Reader(data)
    io = abi.LinearHasherClosedForm()
    io.start_flag, io.completion_flag = r.boolean(), r.boolean()
    _read_queue_state(r, io.queue_state)
    for i, b in enumerate(bytes(r.take(32))):
        io.keccak256_hash[i] = b
    recs, prev = _read_log_queue(r)
    r.done()
    return LinearHasherCircuitInstanceWitness(io, recs, prev)


def write_linear_hasher_witness(w_) -> bytes:
    w = Writer()
    io = w_.closed_form_input
    w.boolean(io.start_flag); w.boolean(io.completion_flag)
    _write_queue_state(w, io.queue_state)
    w.b += bytes(int(b) & 0xFF for b in io.keccak256_hash)
    _write_log_queue(w, w_.queue_witness, w_.queue_prev_tails)
    return bytes(w.b)


# ---- main_vm -----------------------------------------------------------------------------------------------------------------
# VmLocalStateWitness: base_structures/vm_state/mod.rs:92-109 (field order = declaration order); Callstack :callstack.rs:17-21,
# FullExecutionContext callstack.rs:45-49, ExecutionContextRecord saved_context.rs:36-59, VMRegister register/mod.rs:21-24.
# UInt256 -> U256 (hex string), UInt160 -> Address (40 hex digits), UInt32 / UInt16 / UInt8 -> u32 / u16 / u8, Boolean -> bool.
def _read_vm_context_record(r: Reader, c):
    for name in ("this_address", "caller", "code_address"):
        for i, v in enumerate(_limbs(r.h160(), 5)):
            getattr(c, name)[i] = v
    c.code_page, c.base_page, c.heap_upper_bound, c.aux_heap_upper_bound = r.u32(), r.u32(), r.u32(), r.u32()
    for i, v in enumerate(r.fields(4)):
        c.reverted_queue_head[i] = v
    for i, v in enumerate(r.fields(4)):
        c.reverted_queue_tail[i] = v
    c.reverted_queue_segment_len = r.u32()
    c.pc, c.sp, c.exception_handler_loc = r.u16(), r.u16(), r.u16()
    c.ergs_remaining = r.u32()
    c.is_static_execution, c.is_kernel_mode = r.boolean(), r.boolean()
    c.this_shard_id, c.caller_shard_id, c.code_shard_id = r.u8(), r.u8(), r.u8()
    for i in range(4):
        c.context_u128_value_composite[i] = r.u32()
    c.is_local_call = r.boolean()


def _write_vm_context_record(w: Writer, c):
    for name in ("this_address", "caller", "code_address"):
        w.h160(_from_limbs(getattr(c, name)))
    for v in (c.code_page, c.base_page, c.heap_upper_bound, c.aux_heap_upper_bound):
        w.u32(v)
    w.fields(c.reverted_queue_head); w.fields(c.reverted_queue_tail)
    w.u32(c.reverted_queue_segment_len)
    w.u16(c.pc); w.u16(c.sp); w.u16(c.exception_handler_loc)
    w.u32(c.ergs_remaining)
    w.boolean(c.is_static_execution); w.boolean(c.is_kernel_mode)
    w.u8(c.this_shard_id); w.u8(c.caller_shard_id); w.u8(c.code_shard_id)
    for v in c.context_u128_value_composite:
        w.u32(v)
    w.boolean(c.is_local_call)


def _read_vm_state(r: Reader, st):
    for i, v in enumerate(_limbs(r.u256(), 8)):
        st.previous_code_word[i] = v
    for k in range(15):
        st.registers[k].is_pointer = r.boolean()
        for i, v in enumerate(_limbs(r.u256(), 8)):
            st.registers[k].value[i] = v
    for i in range(3):
        st.flags[i] = r.boolean()
    st.timestamp, st.memory_page_counter, st.tx_number_in_block, st.previous_code_page = r.u32(), r.u32(), r.u32(), r.u32()
    st.previous_super_pc = r.u16()
    st.pending_exception = r.boolean()
    st.ergs_per_pubdata_byte = r.u32()
    # callstack: current_context { saved_context, log_queue_forward_tail, log_queue_forward_part_length }, depth, sponge state
    _read_vm_context_record(r, st.current_context)
    for i, v in enumerate(r.fields(4)):
        st.current_context.log_queue_forward_tail[i] = v
    st.current_context.log_queue_forward_part_length = r.u32()
    st.context_stack_depth = r.u32()
    for i, v in enumerate(r.fields(12)):
        st.stack_sponge_state[i] = v
    for i, v in enumerate(r.fields(12)):
        st.memory_queue_state[i] = v
    st.memory_queue_length = r.u32()
    for i, v in enumerate(r.fields(12)):
        st.code_decommittment_queue_state[i] = v
    st.code_decommittment_queue_length = r.u32()
    for i in range(4):
        st.context_composite_u128[i] = r.u32()


def _write_vm_state(w: Writer, st):
    w.u256(_from_limbs(st.previous_code_word))
    for k in range(15):
        w.boolean(st.registers[k].is_pointer)
        w.u256(_from_limbs(st.registers[k].value))
    for i in range(3):
        w.boolean(st.flags[i])
    for v in (st.timestamp, st.memory_page_counter, st.tx_number_in_block, st.previous_code_page):
        w.u32(v)
    w.u16(st.previous_super_pc)
    w.boolean(st.pending_exception)
    w.u32(st.ergs_per_pubdata_byte)
    _write_vm_context_record(w, st.current_context)
    w.fields(st.current_context.log_queue_forward_tail)
    w.u32(st.current_context.log_queue_forward_part_length)
    w.u32(st.context_stack_depth)
    w.fields(st.stack_sponge_state)
    w.fields(st.memory_queue_state); w.u32(st.memory_queue_length)
    w.fields(st.code_decommittment_queue_state); w.u32(st.code_decommittment_queue_length)
    for v in st.context_composite_u128:
        w.u32(v)


def read_vm_closed_form_input(r: Reader):
    """VmCircuitInputOutputWitness = ClosedFormInputWitness<F, VmLocalState, VmInputData, VmOutputData>
    (fsm_input_output/circuit_inputs/main_vm.rs:9-61, fsm_input_output/mod.rs:32-48) from an open Reader -> abi.VmClosedForm"""
    io = abi.VmClosedForm()
    io.start_flag, io.completion_flag = r.boolean(), r.boolean()
    # observable_input: VmInputData
    for i, v in enumerate(r.fields(4)):
        io.rollback_queue_tail_for_block[i] = v
    for i, v in enumerate(r.fields(12)):
        io.memory_queue_initial_tail[i] = v
    io.memory_queue_initial_length = r.u32()
    for i, v in enumerate(r.fields(12)):
        io.decommitment_queue_initial_tail[i] = v
    io.decommitment_queue_initial_length = r.u32()
    io.zkporter_is_available = r.boolean()                                  # per_block_context: GlobalContext, vm_state/mod.rs:159-162
    for i, v in enumerate(_limbs(r.u256(), 8)):
        io.default_aa_code_hash[i] = v
    # observable_output: VmOutputData
    _read_queue_state(r, io.log_queue_final_state)
    _read_queue_state(r, io.memory_queue_final_state)
    _read_queue_state(r, io.decommitment_queue_final_state)
    _read_vm_state(r, io.hidden_fsm_input)
    _read_vm_state(r, io.hidden_fsm_output)
    return io


def write_vm_closed_form_input(w: Writer, io):
    w.boolean(io.start_flag); w.boolean(io.completion_flag)
    w.fields(io.rollback_queue_tail_for_block)
    w.fields(io.memory_queue_initial_tail); w.u32(io.memory_queue_initial_length)
    w.fields(io.decommitment_queue_initial_tail); w.u32(io.decommitment_queue_initial_length)
    w.boolean(io.zkporter_is_available)
    w.u256(_from_limbs(io.default_aa_code_hash))
    _write_queue_state(w, io.log_queue_final_state)
    _write_queue_state(w, io.memory_queue_final_state)
    _write_queue_state(w, io.decommitment_queue_final_state)
    _write_vm_state(w, io.hidden_fsm_input)
    _write_vm_state(w, io.hidden_fsm_output)


def read_vm_circuit_witness(data: bytes, read_oracle=None):
    """bincode bytes of VmCircuitWitness<GoldilocksField, W> (circuit_inputs/main_vm.rs:64-71): `closed_form_input` followed by
    `witness_oracle: W`.  W is whatever type the harness implements WitnessOracle with -- it is not defined in the reference crate
    -- so its bytes are handed to `read_oracle(reader)` (which must consume them; its result is returned next to the closed form).
    With read_oracle = None the dump must end after the closed form (a W that serialises to nothing, e.g. a unit placeholder).
    The engine's own per-cycle form of the oracle's answers is zkc_vm_cycle_witness."""
    r = Reader(data)
    io = read_vm_closed_form_input(r)
    oracle = read_oracle(r) if read_oracle is not None else None
    r.done()
    return io, oracle


def write_vm_circuit_witness(io, oracle_bytes: bytes = b"") -> bytes:
    w = Writer()
    write_vm_closed_form_input(w, io)
    return bytes(w.b) + bytes(oracle_bytes)
