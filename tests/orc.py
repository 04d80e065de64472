"""ctypes loader of the CPU oracle (oracle/liborc.so) -- TEST INFRASTRUCTURE.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

from era_zkevm_circuits_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None
_vp = C.c_void_p


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ROOT, "oracle", "liborc.so")
    srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith((".c", ".h"))]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(path)
    for name in ("orc_gl_mul", "orc_gl_add", "orc_gl_sub"):
        getattr(lib, name).restype = C.c_uint64
        getattr(lib, name).argtypes = [C.c_uint64, C.c_uint64]
    lib.orc_gl_inv.restype = C.c_uint64
    lib.orc_gl_inv.argtypes = [C.c_uint64]
    lib.orc_poseidon2_constants.argtypes = [_vp]
    lib.orc_poseidon2_permutation.argtypes = [_vp]
    lib.orc_commit_encoding.argtypes = [_vp, C.c_size_t, _vp]
    lib.orc_produce_fs_challenges.argtypes = [_vp, C.c_uint32, _vp, C.c_uint32, C.c_int, C.c_int, _vp]
    lib.orc_memory_query_encode.argtypes = [_vp, _vp]
    lib.orc_accumulate_grand_products.argtypes = [_vp, _vp, _vp, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _vp, _vp]
    lib.orc_memory_queue_simulate.argtypes = [_vp, C.c_size_t, _vp, C.POINTER(abi.QueueState12)]
    lib.orc_ram_permutation_entry_point.restype = C.c_int
    lib.orc_ram_permutation_entry_point.argtypes = [C.POINTER(abi.RamClosedForm), _vp, C.c_size_t, _vp, C.c_size_t,
                                                    C.c_size_t, C.POINTER(abi.RamOptions), _vp, _vp, C.POINTER(abi.Status)]
    lib.orc_log_query_encode.argtypes = [_vp, _vp]
    lib.orc_log_queue_simulate.argtypes = [_vp, _vp, C.c_size_t, _vp, C.POINTER(abi.QueueState4)]
    lib.orc_log_sorter_entry_point.restype = C.c_int
    lib.orc_log_sorter_entry_point.argtypes = [C.POINTER(abi.EventsClosedForm), _vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t,
                                               C.POINTER(abi.SorterOptions), _vp, _vp, C.POINTER(C.c_size_t), _vp,
                                               C.POINTER(abi.Status)]
    lib.orc_storage_validity_entry_point.restype = C.c_int
    lib.orc_storage_validity_entry_point.argtypes = [C.POINTER(abi.StorageClosedForm), _vp, C.c_size_t, _vp, _vp, C.c_size_t,
                                                     C.c_size_t, C.POINTER(abi.SorterOptions), _vp, _vp,
                                                     C.POINTER(C.c_size_t), _vp, C.POINTER(abi.Status)]
    lib.orc_decommit_query_encode.argtypes = [_vp, _vp]
    lib.orc_decommit_queue_simulate.argtypes = [_vp, C.c_size_t, _vp, C.POINTER(abi.QueueState12)]
    lib.orc_sort_decommittments_entry_point.restype = C.c_int
    lib.orc_sort_decommittments_entry_point.argtypes = [C.POINTER(abi.DecommitSorterClosedForm), _vp, C.c_size_t, _vp, C.c_size_t,
                                                        C.c_size_t, C.POINTER(abi.SorterOptions), _vp, _vp,
                                                        C.POINTER(C.c_size_t), _vp, C.POINTER(abi.Status)]
    lib.orc_demux_log_queue_entry_point.restype = C.c_int
    lib.orc_demux_log_queue_entry_point.argtypes = [C.POINTER(abi.DemuxClosedForm), _vp, C.c_size_t, C.c_size_t,
                                                    C.POINTER(abi.DemuxOptions), _vp, _vp, C.POINTER(C.c_size_t), _vp,
                                                    C.POINTER(abi.Status)]
    lib.orc_code_unpacker_entry_point.restype = C.c_int
    lib.orc_code_unpacker_entry_point.argtypes = [C.POINTER(abi.CodeUnpackerClosedForm), _vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t,
                                                  C.POINTER(abi.SorterOptions), _vp, _vp, C.POINTER(C.c_size_t), _vp,
                                                  C.POINTER(abi.Status)]
    lib.orc_keccak_f1600.argtypes = [_vp]
    lib.orc_keccak256.argtypes = [_vp, C.c_size_t, _vp]
    lib.orc_keccak256_entry_point.restype = C.c_int
    lib.orc_keccak256_entry_point.argtypes = [C.POINTER(abi.KeccakClosedForm), _vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t,
                                              C.POINTER(abi.PrecompileOptions), _vp, _vp, C.POINTER(C.c_size_t), _vp,
                                              C.POINTER(abi.Status)]
    lib.orc_sha256_compress.argtypes = [_vp, _vp]
    lib.orc_sha256_entry_point.restype = C.c_int
    lib.orc_sha256_entry_point.argtypes = [C.POINTER(abi.Sha256ClosedForm), _vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t,
                                           C.POINTER(abi.PrecompileOptions), _vp, _vp, C.POINTER(C.c_size_t), _vp,
                                           C.POINTER(abi.Status)]
    lib.orc_vm_initial_bootloader_state.argtypes = [C.POINTER(abi.VmClosedForm), C.POINTER(abi.VmIsa), C.POINTER(abi.VmState)]
    lib.orc_main_vm_run.restype = C.c_int
    lib.orc_main_vm_run.argtypes = [C.POINTER(abi.VmIsa), C.POINTER(abi.VmClosedForm), C.POINTER(abi.VmState), _vp, C.c_size_t, C.c_size_t, _vp, _vp, _vp,
                                    C.c_size_t, C.POINTER(C.c_size_t), _vp, C.POINTER(abi.Status)]
    lib.orc_main_vm_entry_point.restype = C.c_int
    lib.orc_main_vm_entry_point.argtypes = [C.POINTER(abi.VmClosedForm), C.POINTER(abi.VmIsa), _vp, _vp, _vp, C.c_size_t, C.c_size_t,
                                            C.POINTER(abi.VmOptions), _vp, _vp, C.POINTER(abi.Status)]
    lib.orc_vm_flatten_state.restype = C.c_size_t
    lib.orc_vm_flatten_state.argtypes = [_vp, _vp]
    _LIB = lib
    return lib


def p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def poseidon2(lib, states):
    out = np.array(states, dtype=np.uint64, copy=True).reshape(-1, 12)
    for i in range(out.shape[0]):
        lib.orc_poseidon2_permutation(C.c_void_p(out[i].ctypes.data))
    return out


def commit_encoding(lib, row):
    row = np.ascontiguousarray(row, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    lib.orc_commit_encoding(p(row), len(row), p(out))
    return out


def memory_queue_simulate(lib, records):
    records = np.ascontiguousarray(records)
    prev = np.zeros((len(records), 12), dtype=np.uint64)
    fin = abi.QueueState12()
    lib.orc_memory_queue_simulate(p(records), len(records), p(prev), C.byref(fin))
    return prev, fin


def accumulate_grand_products(lib, lhs, rhs, ch, acc_in, flags=None, want_chain=False):
    enc_len, rows = lhs.shape
    lhs = np.ascontiguousarray(lhs, dtype=np.uint64); rhs = np.ascontiguousarray(rhs, dtype=np.uint64)
    ch = np.ascontiguousarray(ch, dtype=np.uint64); acc_in = np.ascontiguousarray(acc_in, dtype=np.uint64)
    acc = np.zeros((4, rows), dtype=np.uint64)
    chain = np.zeros((4 * enc_len, rows), dtype=np.uint64) if want_chain else None
    fin = np.zeros(4, dtype=np.uint64)
    if flags is not None:
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
    lib.orc_accumulate_grand_products(p(lhs), p(rhs), p(flags), enc_len, rows, p(ch), p(acc_in), p(acc), p(chain), p(fin))
    return acc, chain, fin


def ram_closed_form(unsorted_state, sorted_state, start=True, nondet_len=0, fsm_in=None):
    io = abi.RamClosedForm()
    io.start_flag = int(start)
    io.observable_input.unsorted_queue_initial_state = unsorted_state
    io.observable_input.sorted_queue_initial_state = sorted_state
    io.observable_input.non_deterministic_bootloader_memory_snapshot_length = nondet_len
    if fsm_in is not None:
        io.hidden_fsm_input = fsm_in
    return io


def ram_entry_point(lib, io, unsorted, sorted_, limit, want_trace=True, compare_expected=False, heap_page=0):
    """returns (rc, io_out, trace, commitment, status)"""
    io2 = abi.RamClosedForm.from_buffer_copy(bytes(io))
    unsorted = np.ascontiguousarray(unsorted); sorted_ = np.ascontiguousarray(sorted_)
    trace = np.zeros((abi.RAM_COLS["NUM_COLS"], limit), dtype=np.uint64) if want_trace else None
    com = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    opts = abi.RamOptions(heap_page, int(compare_expected))
    rc = lib.orc_ram_permutation_entry_point(C.byref(io2), p(unsorted), len(unsorted), p(sorted_), len(sorted_), limit,
                                             C.byref(opts), p(trace), p(com), C.byref(st))
    return rc, io2, trace, com, st


def log_queue_simulate(lib, records, extra_ts=None):
    records = np.ascontiguousarray(records)
    prev = np.zeros((len(records), 4), dtype=np.uint64)
    fin = abi.QueueState4()
    if extra_ts is not None:
        extra_ts = np.ascontiguousarray(extra_ts, dtype=np.uint32)
    lib.orc_log_queue_simulate(p(records), p(extra_ts), len(records), p(prev), C.byref(fin))
    return prev, fin


def events_closed_form(unsorted_state, sorted_state, start=True, fsm_in=None):
    io = abi.EventsClosedForm()
    io.start_flag = int(start)
    io.initial_log_queue_state = unsorted_state
    io.intermediate_sorted_queue_state = sorted_state
    if fsm_in is not None:
        io.hidden_fsm_input = fsm_in
    return io


def log_sorter_entry_point(lib, io, unsorted, sorted_, limit, want_trace=True, compare_expected=False):
    """returns (rc, io_out, trace, commitment, status, result_tails)"""
    io2 = abi.EventsClosedForm.from_buffer_copy(bytes(io))
    unsorted = np.ascontiguousarray(unsorted); sorted_ = np.ascontiguousarray(sorted_)
    trace = np.zeros((abi.EV_COLS["NUM_COLS"], limit), dtype=np.uint64) if want_trace else None
    tails = np.zeros((limit + 1, 4), dtype=np.uint64)
    n_tails = C.c_size_t()
    com = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    opts = abi.SorterOptions(int(compare_expected))
    rc = lib.orc_log_sorter_entry_point(C.byref(io2), p(unsorted), len(unsorted), p(sorted_), len(sorted_), limit,
                                        C.byref(opts), p(trace), p(tails), C.byref(n_tails), p(com), C.byref(st))
    return rc, io2, trace, com, st, tails[:n_tails.value].copy()


def decommit_queue_simulate(lib, records):
    """returns (prev_states [n, 12], final QueueState12)"""
    records = np.ascontiguousarray(records)
    prev = np.zeros((len(records), 12), dtype=np.uint64)
    fin = abi.QueueState12()
    lib.orc_decommit_queue_simulate(p(records), len(records), p(prev), C.byref(fin))
    return prev, fin


def decommit_sorter_closed_form(unsorted_state, sorted_state, start=True, fsm_in=None):
    io = abi.DecommitSorterClosedForm()
    io.start_flag = int(start)
    io.initial_queue_state = unsorted_state
    io.sorted_queue_initial_state = sorted_state
    if fsm_in is not None:
        io.hidden_fsm_input = fsm_in
    return io


def sort_decommittments_entry_point(lib, io, unsorted, sorted_, limit, want_trace=True, compare_expected=False):
    """returns (rc, io_out, trace, commitment, status, result_states)"""
    io2 = abi.DecommitSorterClosedForm.from_buffer_copy(bytes(io))
    unsorted = np.ascontiguousarray(unsorted); sorted_ = np.ascontiguousarray(sorted_)
    trace = np.zeros((abi.DQ_COLS["NUM_COLS"], limit), dtype=np.uint64) if want_trace else None
    states = np.zeros((limit + 1, 12), dtype=np.uint64)
    n_states = C.c_size_t()
    com = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    opts = abi.SorterOptions(int(compare_expected))
    rc = lib.orc_sort_decommittments_entry_point(C.byref(io2), p(unsorted), len(unsorted), p(sorted_), len(sorted_), limit,
                                                 C.byref(opts), p(trace), p(states), C.byref(n_states), p(com), C.byref(st))
    return rc, io2, trace, com, st, states[:n_states.value].copy()


def demux_closed_form(queue_state, start=True, fsm_in=None):
    io = abi.DemuxClosedForm()
    io.start_flag = int(start)
    io.initial_log_queue_state = queue_state
    if fsm_in is not None:
        io.hidden_fsm_input = fsm_in
    return io


def demux_entry_point(lib, io, records, limit, want_trace=True, compare_expected=False, options=None):
    """returns (rc, io_out, trace, commitment, status, tails): tails = list of 6 arrays [n_q, 4] (tail after each push)"""
    io2 = abi.DemuxClosedForm.from_buffer_copy(bytes(io))
    records = np.ascontiguousarray(records)
    trace = np.zeros((abi.DMX_COLS["NUM_COLS"], limit), dtype=np.uint64) if want_trace else None
    tails = np.zeros((6, max(limit, 1), 4), dtype=np.uint64)
    n_tails = (C.c_size_t * 6)()
    com = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    opts = options if options is not None else abi.DemuxOptions()
    opts.compare_expected = int(compare_expected)
    rc = lib.orc_demux_log_queue_entry_point(C.byref(io2), p(records), len(records), limit, C.byref(opts), p(trace), p(tails),
                                             n_tails, p(com), C.byref(st))
    return rc, io2, trace, com, st, [tails[q, :n_tails[q]].copy() for q in range(6)]


def linear_hasher_closed_form(queue_state, start=True):
    io = abi.LinearHasherClosedForm()
    io.start_flag = int(start)
    io.queue_state = queue_state
    return io


def linear_hasher_entry_point(lib, io, records, limit, want_trace=True, compare_expected=False):
    """returns (rc, io_out, trace, commitment, status, keccak_states [limit, 25])"""
    io2 = abi.LinearHasherClosedForm.from_buffer_copy(bytes(io))
    records = np.ascontiguousarray(records)
    trace = np.zeros((abi.LH_COLS["NUM_COLS"], limit), dtype=np.uint64) if want_trace else None
    states = np.zeros((max(limit, 1), 25), dtype=np.uint64)
    com = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    opts = abi.SorterOptions(int(compare_expected))
    lib.orc_linear_hasher_entry_point.restype = C.c_int
    lib.orc_linear_hasher_entry_point.argtypes = [C.POINTER(abi.LinearHasherClosedForm), _vp, C.c_size_t, C.c_size_t, C.POINTER(abi.SorterOptions),
                                                  _vp, _vp, _vp, C.POINTER(abi.Status)]
    rc = lib.orc_linear_hasher_entry_point(C.byref(io2), p(records), len(records), limit, C.byref(opts), p(trace), p(states), p(com), C.byref(st))
    return rc, io2, trace, com, st, states[:limit].copy()


def code_unpacker_closed_form(requests_state, memory_state=None, start=True, fsm_in=None):
    io = abi.CodeUnpackerClosedForm()
    io.start_flag = int(start)
    io.sorted_requests_queue_initial_state = requests_state
    if memory_state is not None:
        io.memory_queue_initial_state = memory_state
    if fsm_in is not None:
        io.hidden_fsm_input = fsm_in
    return io


def code_unpacker_entry_point(lib, io, requests, code_words, limit, want_trace=True, compare_expected=False):
    """returns (rc, io_out, trace, commitment, status, memory_states)"""
    io2 = abi.CodeUnpackerClosedForm.from_buffer_copy(bytes(io))
    requests = np.ascontiguousarray(requests)
    code_words = np.ascontiguousarray(code_words, dtype=np.uint32).reshape(-1, 8)
    trace = np.zeros((abi.CU_COLS["NUM_COLS"], limit), dtype=np.uint64) if want_trace else None
    states = np.zeros((2 * limit + 1, 12), dtype=np.uint64)
    n_states = C.c_size_t()
    com = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    opts = abi.SorterOptions(int(compare_expected))
    rc = lib.orc_code_unpacker_entry_point(C.byref(io2), p(requests), len(requests), p(code_words), len(code_words), limit,
                                           C.byref(opts), p(trace), p(states), C.byref(n_states), p(com), C.byref(st))
    return rc, io2, trace, com, st, states[:n_states.value].copy()


def storage_closed_form(unsorted_state, sorted_state, shard=0, start=True, fsm_in=None):
    io = abi.StorageClosedForm()
    io.start_flag = int(start)
    io.shard_id_to_process = shard
    io.unsorted_log_queue_state = unsorted_state
    io.intermediate_sorted_queue_state = sorted_state
    if fsm_in is not None:
        io.hidden_fsm_input = fsm_in
    return io


def storage_validity_entry_point(lib, io, unsorted, sorted_, sorted_ts, limit, want_trace=True, compare_expected=False):
    """returns (rc, io_out, trace, commitment, status, result_tails)"""
    io2 = abi.StorageClosedForm.from_buffer_copy(bytes(io))
    unsorted = np.ascontiguousarray(unsorted); sorted_ = np.ascontiguousarray(sorted_)
    sorted_ts = np.ascontiguousarray(sorted_ts, dtype=np.uint32)
    trace = np.zeros((abi.ST_COLS["NUM_COLS"], limit), dtype=np.uint64) if want_trace else None
    tails = np.zeros((limit + 1, 4), dtype=np.uint64)
    n_tails = C.c_size_t()
    com = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    opts = abi.SorterOptions(int(compare_expected))
    rc = lib.orc_storage_validity_entry_point(C.byref(io2), p(unsorted), len(unsorted), p(sorted_), p(sorted_ts), len(sorted_),
                                              limit, C.byref(opts), p(trace), p(tails), C.byref(n_tails), p(com), C.byref(st))
    return rc, io2, trace, com, st, tails[:n_tails.value].copy()


def keccak_closed_form(requests_state, memory_state=None, start=True, fsm_in=None):
    io = abi.KeccakClosedForm()
    io.start_flag = int(start)
    io.initial_log_queue_state = requests_state
    if memory_state is not None:
        io.initial_memory_queue_state = memory_state
    if fsm_in is not None:
        io.hidden_fsm_input = fsm_in
    return io


def keccak_entry_point(lib, io, requests, memory_reads, limit, want_trace=True, compare_expected=False):
    """returns (rc, io_out, trace, commitment, status, memory_states)"""
    io2 = abi.KeccakClosedForm.from_buffer_copy(bytes(io))
    requests = np.ascontiguousarray(requests)
    memory_reads = np.ascontiguousarray(memory_reads, dtype=np.uint32).reshape(-1, 8)
    trace = np.zeros((abi.KC_COLS["NUM_COLS"], limit), dtype=np.uint64) if want_trace else None
    states = np.zeros((7 * limit + 1, 12), dtype=np.uint64)
    n_states = C.c_size_t()
    com = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    opts = abi.PrecompileOptions(int(compare_expected), 0, 0, 0)
    rc = lib.orc_keccak256_entry_point(C.byref(io2), p(requests), len(requests), p(memory_reads), len(memory_reads), limit,
                                       C.byref(opts), p(trace), p(states), C.byref(n_states), p(com), C.byref(st))
    return rc, io2, trace, com, st, states[:n_states.value].copy()


def sha256_closed_form(requests_state, memory_state=None, start=True, fsm_in=None):
    io = abi.Sha256ClosedForm()
    io.start_flag = int(start)
    io.initial_log_queue_state = requests_state
    if memory_state is not None:
        io.initial_memory_queue_state = memory_state
    if fsm_in is not None:
        io.hidden_fsm_input = fsm_in
    return io


def sha256_entry_point(lib, io, requests, memory_reads, limit, want_trace=True, compare_expected=False):
    """returns (rc, io_out, trace, commitment, status, memory_states)"""
    io2 = abi.Sha256ClosedForm.from_buffer_copy(bytes(io))
    requests = np.ascontiguousarray(requests)
    memory_reads = np.ascontiguousarray(memory_reads, dtype=np.uint32).reshape(-1, 8)
    trace = np.zeros((abi.SH_COLS["NUM_COLS"], limit), dtype=np.uint64) if want_trace else None
    states = np.zeros((3 * limit + 1, 12), dtype=np.uint64)
    n_states = C.c_size_t()
    com = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    opts = abi.PrecompileOptions(int(compare_expected), 0, 0, 0)
    rc = lib.orc_sha256_entry_point(C.byref(io2), p(requests), len(requests), p(memory_reads), len(memory_reads), limit,
                                    C.byref(opts), p(trace), p(states), C.byref(n_states), p(com), C.byref(st))
    return rc, io2, trace, com, st, states[:n_states.value].copy()


def vm_initial_state(lib, io, isa):
    st = abi.VmState()
    lib.orc_vm_initial_bootloader_state(C.byref(io), C.byref(isa), C.byref(st))
    return st


def vm_run(lib, isa, initial, code_words, cycles, full=False, gc=None):
    """out-of-circuit run: returns (rc, snapshots [cycles + 1, 1176] uint8, witness [cycles, 176] uint8, status); full=True
    appends (callstack witness [n, 336] uint8, resolved rollback_queue_tail_for_block [4])"""
    code_words = np.ascontiguousarray(code_words, dtype=np.uint32)
    snaps = np.zeros((cycles + 1, C.sizeof(abi.VmState)), dtype=np.uint8)
    wit = np.zeros((cycles, C.sizeof(abi.VmCycleWitness)), dtype=np.uint8)
    cw = np.zeros((cycles + 1, C.sizeof(abi.VmCallstackWitness)), dtype=np.uint8)
    n_cw = C.c_size_t()
    tail = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    rc = lib.orc_main_vm_run(C.byref(isa), C.byref(gc) if gc is not None else None, C.byref(initial), p(code_words), len(code_words), cycles, p(snaps), p(wit), p(cw),
                             len(cw), C.byref(n_cw), p(tail), C.byref(st))
    if full:
        return rc, snaps, wit, st, cw[:n_cw.value].copy(), tail
    return rc, snaps, wit, st


def vm_state_at(snaps, i):
    return abi.VmState.from_buffer_copy(snaps[i].tobytes())


def vm_entry_point(lib, io, isa, snaps, wit, limit, want_trace=True, compare_expected=False, cw=None):
    io2 = abi.VmClosedForm.from_buffer_copy(bytes(io))
    trace = np.zeros((abi.VM_COLS["NUM_COLS"], limit), dtype=np.uint64) if want_trace else None
    com = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    opts = abi.VmOptions(int(compare_expected))
    n_cw = 0 if cw is None else len(cw)
    rc = lib.orc_main_vm_entry_point(C.byref(io2), C.byref(isa), p(np.ascontiguousarray(snaps)), p(np.ascontiguousarray(wit)),
                                     p(np.ascontiguousarray(cw)) if n_cw else None, n_cw, limit,
                                     C.byref(opts), p(trace), p(com), C.byref(st))
    return rc, io2, trace, com, st


def vm_gadget_cells(lib, trace, limit, n_instances=1):
    """orc_main_vm_gadget_cells: DENSE trace(s) -> [n?, VMG_COLS.NUM_COLS, limit]"""
    trace = np.ascontiguousarray(trace, dtype=np.uint64)
    out = np.zeros(tuple(trace.shape[:-2]) + (abi.VMG_COLS["NUM_COLS"], limit), dtype=np.uint64)
    lib.orc_main_vm_gadget_cells.restype = None
    lib.orc_main_vm_gadget_cells.argtypes = [_vp, C.c_size_t, C.c_size_t, _vp]
    lib.orc_main_vm_gadget_cells(p(trace), limit, n_instances, p(out))
    return out


def vm_state_gadget_cells(lib, trace, snaps, limit, n_instances=1):
    """orc_main_vm_state_gadget_cells: DENSE trace(s) + snapshots [n?, limit + 1] -> [n?, VMS_COLS.NUM_COLS, limit]"""
    trace = np.ascontiguousarray(trace, dtype=np.uint64)
    snaps = np.ascontiguousarray(snaps)
    assert snaps.nbytes >= (limit + 1) * n_instances * C.sizeof(abi.VmState)
    out = np.zeros(tuple(trace.shape[:-2]) + (abi.VMS_COLS["NUM_COLS"], limit), dtype=np.uint64)
    lib.orc_main_vm_state_gadget_cells.restype = None
    lib.orc_main_vm_state_gadget_cells.argtypes = [_vp, _vp, C.c_size_t, C.c_size_t, _vp]
    lib.orc_main_vm_state_gadget_cells(p(trace), p(snaps), limit, n_instances, p(out))
    return out


def vm_prestate_cells(lib, trace, snaps, limit, n_instances=1, fn=None):
    """orc_main_vm_prestate_cells: DENSE trace(s) + snapshots [n?, limit + 1] -> [n?, VMP_COLS.NUM_COLS, limit]; `fn` = another function
    of the same signature (the g++ build of the kernel's row statement, tests/cpp/prestate_row_host.cpp)"""
    trace = np.ascontiguousarray(trace, dtype=np.uint64)
    snaps = np.ascontiguousarray(snaps)
    assert snaps.nbytes >= (limit + 1) * n_instances * C.sizeof(abi.VmState)
    out = np.zeros(tuple(trace.shape[:-2]) + (abi.VMP_COLS["NUM_COLS"], limit), dtype=np.uint64)
    fn = fn or lib.orc_main_vm_prestate_cells
    fn.restype = None
    fn.argtypes = [_vp, _vp, C.c_size_t, C.c_size_t, _vp]
    fn(p(trace), p(snaps), limit, n_instances, p(out))
    return out


def vm_writeback_cells(lib, isa, trace, snaps, limit, n_instances=1, fn=None):
    """orc_main_vm_writeback_cells: ISA tables + DENSE trace(s) + snapshots [n?, limit + 1] -> [n?, VMW_COLS.NUM_COLS, limit]; `fn` as above"""
    trace = np.ascontiguousarray(trace, dtype=np.uint64)
    snaps = np.ascontiguousarray(snaps)
    assert snaps.nbytes >= (limit + 1) * n_instances * C.sizeof(abi.VmState)
    out = np.zeros(tuple(trace.shape[:-2]) + (abi.VMW_COLS["NUM_COLS"], limit), dtype=np.uint64)
    fn = fn or lib.orc_main_vm_writeback_cells
    fn.restype = None
    fn.argtypes = [C.POINTER(abi.VmIsa), _vp, _vp, C.c_size_t, C.c_size_t, _vp]
    fn(C.byref(isa), p(trace), p(snaps), limit, n_instances, p(out))
    return out


def vm_memory_sponge_cells(lib, trace, snaps, limit, n_instances=1):
    """orc_main_vm_memory_sponge_cells: DENSE trace(s) + snapshots [n?, limit + 1] -> [n?, VMQ_COLS.NUM_COLS, limit]"""
    trace = np.ascontiguousarray(trace, dtype=np.uint64)
    snaps = np.ascontiguousarray(snaps)
    assert snaps.nbytes >= (limit + 1) * n_instances * C.sizeof(abi.VmState)
    out = np.zeros(tuple(trace.shape[:-2]) + (abi.VMQ_COLS["NUM_COLS"], limit), dtype=np.uint64)
    lib.orc_main_vm_memory_sponge_cells.restype = None
    lib.orc_main_vm_memory_sponge_cells.argtypes = [_vp, _vp, C.c_size_t, C.c_size_t, _vp]
    lib.orc_main_vm_memory_sponge_cells(p(trace), p(snaps), limit, n_instances, p(out))
    return out
