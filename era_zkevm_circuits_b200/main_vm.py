"""Host-side mirror of `main_vm_entry_point` (/root/reference/src/main_vm/mod.rs:47-232) plus the out-of-circuit run
that produces its witness (the role of the external zk_evm crate)."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import abi
from .engine import Engine, ZkcError, on_device, ptr
from .log_sorter import SorterResult


@dataclass
class VmCircuitWitness:
    """fsm_input_output/circuit_inputs/main_vm.rs:64-71.  The WitnessOracle is flattened into per-cycle answers (+ the
    frames its rets pop, out of line), and the per-cycle VmLocalState snapshots make the instance data parallel."""
    closed_form_input: abi.VmClosedForm
    isa: abi.VmIsa
    snapshots: object  # [limit + 1, 1176] uint8 (numpy) or torch uint8 on the GPU: zkc_vm_state before each cycle + final
    witness_oracle: object  # [limit, 176] uint8: zkc_vm_cycle_witness per cycle
    callstack_witness: object = None  # [n, 336] uint8: zkc_vm_callstack_witness, indexed by witness.callstack_index


def _n_cw(cw):
    return 0 if cw is None else int(cw.shape[-2])


def main_vm_entry_point(engine: Engine, witness: VmCircuitWitness, limit: int, want_trace=True, compare_expected=False,
                        raise_on_unsatisfied=True, trace_out=None) -> SorterResult:
    w = witness
    dev = on_device(w.snapshots, w.witness_oracle)
    if w.callstack_witness is not None and _n_cw(w.callstack_witness):
        assert on_device(w.callstack_witness) == dev, "snapshots / witness / callstack witness must live on the same side"
    if trace_out is not None:
        dev |= 2 * on_device(trace_out)
    elif dev:
        dev = 3
    trace = trace_out
    if want_trace and trace is None:
        if dev & 2:
            import torch
            trace = torch.empty((abi.VM_COLS["NUM_COLS"], limit), dtype=torch.int64, device=w.snapshots.device)
        else:
            trace = np.empty((abi.VM_COLS["NUM_COLS"], limit), dtype=np.uint64)
    io = abi.VmClosedForm.from_buffer_copy(bytes(w.closed_form_input))
    opts = abi.VmOptions(int(compare_expected))
    commitment = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    n_cw = _n_cw(w.callstack_witness)
    rc = engine.lib.zkc_main_vm_entry_point(engine.h, C.byref(io), C.byref(w.isa), ptr(w.snapshots), ptr(w.witness_oracle),
                                            ptr(w.callstack_witness) if n_cw else None, n_cw, limit,
                                            C.byref(opts), dev, ptr(trace), ptr(commitment), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE) or (rc and raise_on_unsatisfied):
        raise ZkcError(rc, st, "main_vm_entry_point")
    return SorterResult(commitment, io, trace, st)


def main_vm_initial_state(engine: Engine, closed_form_input: abi.VmClosedForm, isa: abi.VmIsa) -> abi.VmState:
    """initial_bootloader_state, main_vm/loading.rs:13-226"""
    out = abi.VmState()
    rc = engine.lib.zkc_main_vm_initial_state(engine.h, C.byref(closed_form_input), C.byref(isa), C.byref(out))
    if rc:
        raise ZkcError(rc, what="zkc_main_vm_initial_state")
    return out


@dataclass
class VmSimulation:
    snapshots: object          # torch uint8 [n, cycles + 1, 1176]
    witness: object            # torch uint8 [n, cycles, 176]
    callstack_witness: object  # torch uint8 [n, capacity, 336]
    n_callstack: np.ndarray    # [n] frames popped per instance
    rollback_tails: np.ndarray  # [n, 4] resolved rollback_queue_tail_for_block per instance
    status: abi.Status


def main_vm_simulate(engine: Engine, isa: abi.VmIsa, initial_states, code, cycles: int, callstack_capacity=None) -> VmSimulation:
    """Out-of-circuit run of n independent VM instances on the GPU (two passes: the rollback queue is resolved first).
    initial_states: list of abi.VmState (bootloader start states); code: [n, code_words, 8] uint32."""
    import torch
    n = len(initial_states)
    code = np.ascontiguousarray(code, dtype=np.uint32).reshape(n, -1, 8)
    init = np.frombuffer(b"".join(bytes(s) for s in initial_states), dtype=np.uint8).reshape(n, -1)
    d_init = torch.from_numpy(init.copy()).cuda()
    d_code = torch.from_numpy(code.view(np.int32)).cuda()
    cap = callstack_capacity if callstack_capacity is not None else max(16, cycles // 16)
    snaps = torch.empty((n, cycles + 1, C.sizeof(abi.VmState)), dtype=torch.uint8, device="cuda")
    wit = torch.empty((n, cycles, C.sizeof(abi.VmCycleWitness)), dtype=torch.uint8, device="cuda")
    cw = torch.zeros((n, cap, C.sizeof(abi.VmCallstackWitness)), dtype=torch.uint8, device="cuda")
    n_cw = np.zeros(n, dtype=np.uint32)
    tails = np.zeros((n, 4), dtype=np.uint64)
    st = abi.Status()
    rc = engine.lib.zkc_main_vm_simulate(engine.h, C.byref(isa), ptr(d_init), ptr(d_code), code.shape[1], n, cycles, ptr(snaps),
                                         ptr(wit), ptr(cw), cap, ptr(n_cw), ptr(tails), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "zkc_main_vm_simulate")
    return VmSimulation(snaps, wit, cw, n_cw, tails, st)


def main_vm_entry_point_batch(engine: Engine, closed_form_inputs, isa: abi.VmIsa, snapshots, witness_oracle, limit: int,
                              trace_out=None, compare_expected=False, callstack_witness=None, sponge_records_out=None):
    """n independent instances in one set of launches.  closed_form_inputs: list of abi.VmClosedForm;
    snapshots [n, limit + 1, 1176], witness_oracle [n, limit, 176], callstack_witness [n, capacity, 336] (numpy or torch
    CUDA uint8); trace_out: optional [n, NUM_COLS, limit] uint64.  Returns (commitments [n, 4], closed forms (updated),
    statuses, rc)."""
    n = len(closed_form_inputs)
    ios = (abi.VmClosedForm * n)(*[abi.VmClosedForm.from_buffer_copy(bytes(c)) for c in closed_form_inputs])
    dev = on_device(snapshots, witness_oracle)
    if trace_out is not None:
        dev |= 2 * on_device(trace_out)
    commitments = np.zeros((n, 4), dtype=np.uint64)
    statuses = (abi.Status * n)()
    opts = abi.VmOptions(int(compare_expected))
    if sponge_records_out is not None:
        # COMPACT layout: trace_out is [n, VM_COMPACT_COLS, limit]; the sponge columns come back as records
        # (abi.VM_SPONGE_RECORD_DTYPE array or torch uint8 [capacity, 104]); statuses[0].reserved = how many
        assert trace_out is not None and on_device(sponge_records_out) == on_device(trace_out)
        opts.trace_layout = abi.VM_TRACE_COMPACT
        opts.sponge_records_capacity = len(sponge_records_out)
        opts.sponge_records = ptr(sponge_records_out)
    n_cw = _n_cw(callstack_witness)
    rc = engine.lib.zkc_main_vm_entry_point_batch(engine.h, C.cast(ios, C.c_void_p), n, C.byref(isa), ptr(snapshots),
                                                  ptr(witness_oracle), ptr(callstack_witness) if n_cw else None, n_cw, limit,
                                                  C.byref(opts), dev, ptr(trace_out), ptr(commitments), C.cast(statuses, C.c_void_p))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, statuses[0], "main_vm_entry_point_batch")
    return commitments, ios, statuses, rc


@dataclass
class VmColumnInputs:
    """zkc_vm_columns: torch int32 CUDA tensors state_words [294, state_stride], witness_words [44, witness_stride]; word w
    of snapshot i of instance k at state_words[w, k * (limit + 1) + i], of cycle i's oracle answers at witness_words[w, k * limit + i]"""
    state_words: object
    witness_words: object

    def struct(self):
        return abi.VmColumns(self.state_words.data_ptr(), self.state_words.shape[1], self.witness_words.data_ptr(), self.witness_words.shape[1])


def main_vm_rows_to_columns(engine: Engine, snapshots, witness_oracle, limit: int) -> VmColumnInputs:
    """device records [n, limit + 1, 1176] / [n, limit, 176] (torch CUDA uint8) -> device columns"""
    import torch
    assert on_device(snapshots, witness_oracle)
    n = int(snapshots.shape[0]) if snapshots.dim() == 3 else 1
    ss, ws = (max(n * (limit + 1), 1) + 31) // 32 * 32, (max(n * limit, 1) + 31) // 32 * 32
    st = torch.empty((abi.VM_STATE_WORDS, ss), dtype=torch.int32, device=snapshots.device)
    wt = torch.empty((abi.VM_WITNESS_WORDS, ws), dtype=torch.int32, device=snapshots.device)
    rc = engine.lib.zkc_main_vm_rows_to_columns(engine.h, ptr(snapshots), ptr(witness_oracle), n, limit, ptr(st), ss, ptr(wt), ws)
    if rc:
        raise ZkcError(rc, what="zkc_main_vm_rows_to_columns")
    return VmColumnInputs(st, wt)


def main_vm_entry_point_columns(engine: Engine, closed_form_inputs, isa: abi.VmIsa, columns: VmColumnInputs, limit: int, trace_out=None,
                                compare_expected=False, callstack_witness=None, sponge_records_out=None):
    """main_vm_entry_point_batch with the snapshots / oracle answers already in HBM as columns (and the popped frames as device
    records [n, capacity, 336]).  Returns (commitments [n, 4], closed forms (updated), statuses, rc)."""
    n = len(closed_form_inputs)
    ios = (abi.VmClosedForm * n)(*[abi.VmClosedForm.from_buffer_copy(bytes(c)) for c in closed_form_inputs])
    commitments = np.zeros((n, 4), dtype=np.uint64)
    statuses = (abi.Status * n)()
    opts = abi.VmOptions(int(compare_expected))
    if sponge_records_out is not None:
        assert trace_out is not None and on_device(sponge_records_out) == on_device(trace_out)
        opts.trace_layout = abi.VM_TRACE_COMPACT
        opts.sponge_records_capacity = len(sponge_records_out)
        opts.sponge_records = ptr(sponge_records_out)
    n_cw = _n_cw(callstack_witness)
    assert not n_cw or on_device(callstack_witness)
    cs = columns.struct()
    rc = engine.lib.zkc_main_vm_entry_point_columns(engine.h, C.cast(ios, C.c_void_p), n, C.byref(isa), C.byref(cs),
                                                    ptr(callstack_witness) if n_cw else None, n_cw, limit, C.byref(opts),
                                                    int(trace_out is not None and on_device(trace_out)), ptr(trace_out), ptr(commitments),
                                                    C.cast(statuses, C.c_void_p))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, statuses[0], "main_vm_entry_point_columns")
    return commitments, ios, statuses, rc


def main_vm_check_trace(engine: Engine, isa: abi.VmIsa, trace, limit: int, n_instances: int = 1):
    """Constraint evaluation of finished main_vm traces (DENSE layout): the row-local relations of vm_cycle.  trace:
    [NUM_COLS, limit] or [n, NUM_COLS, limit] uint64 (numpy: host, torch CUDA: device).  Returns (violating rows, status)."""
    st = abi.Status()
    viol = C.c_uint64()
    rc = engine.lib.zkc_main_vm_check_trace(engine.h, C.byref(isa), ptr(trace), limit, n_instances, on_device(trace), C.byref(viol), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "main_vm_check_trace")
    return viol.value, st


def main_vm_gadget_cells(engine: Engine, trace, limit: int, n_instances: int = 1):
    """The cells the add/sub, binop, mul/div and shift gadgets allocate on every cycle whatever the opcode, and the relations
    vm_cycle enforces once per cycle (include/zkc_b200.h, ZKC_VM_GADGET_COLUMNS), from finished DENSE traces [NUM_COLS, limit] /
    [n, NUM_COLS, limit].  Returns [VMG_COLS.NUM_COLS, limit] / [n, ..] uint64 in the memory space of `trace`."""
    dev = on_device(trace)
    shape = tuple(trace.shape[:-2]) + (abi.VMG_COLS["NUM_COLS"], limit)
    if dev:
        import torch
        out = torch.empty(shape, dtype=trace.dtype, device=trace.device)
    else:
        trace = np.ascontiguousarray(trace, dtype=np.uint64)
        out = np.empty(shape, dtype=np.uint64)
    rc = engine.lib.zkc_main_vm_gadget_cells(engine.h, ptr(trace), limit, n_instances, dev, ptr(out))
    if rc:
        raise ZkcError(rc, what="zkc_main_vm_gadget_cells")
    return out


def _trace_and_snapshots_block(engine: Engine, fn_name: str, n_cols: int, trace, snapshots, limit: int, n_instances: int, isa=None):
    """shared by the per-cycle blocks computed from finished DENSE traces + the snapshots the entry point took"""
    dev = on_device(trace)
    if dev != on_device(snapshots):
        raise ZkcError(abi.ZKC_ERR_INVALID_ARGUMENT, what=f"{fn_name}: trace and snapshots in different memory spaces")
    shape = tuple(trace.shape[:-2]) + (n_cols, limit)
    need = (limit + 1) * n_instances * C.sizeof(abi.VmState)
    if dev:
        import torch
        have = snapshots.numel() * snapshots.element_size()
        out = torch.empty(shape, dtype=trace.dtype, device=trace.device)
    else:
        trace = np.ascontiguousarray(trace, dtype=np.uint64)
        snapshots = np.ascontiguousarray(snapshots)
        have = snapshots.nbytes
        out = np.empty(shape, dtype=np.uint64)
    rows = 1
    for d in trace.shape[:-2]:
        rows *= int(d)
    if have < need or trace.shape[-2] != abi.VM_COLS["NUM_COLS"] or trace.shape[-1] != limit or rows != n_instances:
        raise ZkcError(abi.ZKC_ERR_INVALID_ARGUMENT, what=f"{fn_name}: trace / snapshots do not cover n_instances x limit cycles")
    lead = (engine.h,) if isa is None else (engine.h, C.byref(isa))   # the write-back block also takes the ISA tables (host)
    rc = getattr(engine.lib, fn_name)(*lead, ptr(trace), ptr(snapshots), limit, n_instances, dev, ptr(out))
    if rc:
        raise ZkcError(rc, what=fn_name)
    return out


def main_vm_state_gadget_cells(engine: Engine, trace, snapshots, limit: int, n_instances: int = 1):
    """The cells the ptr, jump and context gadgets allocate on every cycle whatever the opcode (include/zkc_b200.h,
    ZKC_VM_STATE_GADGET_COLUMNS; opcodes/ptr.rs:6-183, jump.rs:3-38, context.rs:7-307), from finished DENSE traces
    [NUM_COLS, limit] / [n, NUM_COLS, limit] and the snapshots the entry point took ([limit + 1] / [n, limit + 1] records,
    abi.VM_STATE_DTYPE or a byte tensor on the device).  Returns [VMS_COLS.NUM_COLS, limit] / [n, ..] uint64 in the memory
    space of `trace`."""
    return _trace_and_snapshots_block(engine, "zkc_main_vm_state_gadget_cells", abi.VMS_COLS["NUM_COLS"], trace, snapshots, limit, n_instances)


def main_vm_memory_sponge_cells(engine: Engine, trace, snapshots, limit: int, n_instances: int = 1):
    """The three memory-queue relations every cycle evaluates whatever its opcode -- opcode fetch, src0 read, dst0 write
    (main_vm/utils.rs:129-233, :388-522, cycle.rs:799-935, :937-957): initial state, permutation output, selected tail and length
    per step (include/zkc_b200.h, ZKC_VM_MEMORY_SPONGE_COLUMNS).  Same arguments as main_vm_state_gadget_cells; returns
    [VMQ_COLS.NUM_COLS, limit] / [n, ..] uint64 in the memory space of `trace`."""
    return _trace_and_snapshots_block(engine, "zkc_main_vm_memory_sponge_cells", abi.VMQ_COLS["NUM_COLS"], trace, snapshots, limit, n_instances)


def main_vm_prestate_cells(engine: Engine, trace, snapshots, limit: int, n_instances: int = 1):
    """The cells create_prestate allocates on the way to the values the DENSE trace names (main_vm/pre_state.rs:71-519,
    main_vm/utils.rs:106-120, :237-386, decoded_opcode.rs:192-202): cycle control, the opcode select inside the code word, the four
    15-bit register selector masks, the 15-step register select chains, operand locations, the src0 selects, the operand swap and the
    pointer-erasure flags (include/zkc_b200.h, ZKC_VM_PRESTATE_COLUMNS).  Same arguments as main_vm_state_gadget_cells; returns
    [VMP_COLS.NUM_COLS, limit] / [n, ..] uint64 in the memory space of `trace`."""
    return _trace_and_snapshots_block(engine, "zkc_main_vm_prestate_cells", abi.VMP_COLS["NUM_COLS"], trace, snapshots, limit, n_instances)


def main_vm_writeback_cells(engine: Engine, isa, trace, snapshots, limit: int, n_instances: int = 1):
    """The register write-back of the state diffs (cycle.rs:158-433): the dst0 update flags, per register write_as_dst0, the far call /
    far return specific updates, pointer-marker removal and zero-out flags, both is_pointer dot products with their selects, and the
    value select chain dst0 -> far call -> far return -> zero-out -> dst1 (include/zkc_b200.h, ZKC_VM_WRITEBACK_COLUMNS).  `isa`: the
    abi.VmIsa tables (calling-convention register lists); the other arguments as main_vm_state_gadget_cells.  Returns
    [VMW_COLS.NUM_COLS, limit] / [n, ..] uint64 in the memory space of `trace`."""
    return _trace_and_snapshots_block(engine, "zkc_main_vm_writeback_cells", abi.VMW_COLS["NUM_COLS"], trace, snapshots, limit, n_instances, isa=isa)


# ---- transport forms over PCIe (include/zkc_b200.h, "transport forms of the main_vm call") ---------------------------------
class VmInputStreamHandle:
    """a zkc_vm_input_stream of ONE instance, in (pinned) host memory owned by the library"""

    def __init__(self, lib, ptr_, nbytes):
        self.lib, self.ptr, self.bytes = lib, ptr_, nbytes

    @property
    def struct(self):
        return self.ptr.contents

    def free(self):
        if self.ptr:
            self.lib.zkc_vm_input_stream_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def vm_encode_input_stream(lib, snapshots, witness_oracle, limit: int, segment_cycles: int = 0) -> VmInputStreamHandle:
    """records of one instance (numpy uint8 [limit + 1, 1176], [limit, 176]) -> segmented stream (host encoder of the library)"""
    snapshots, witness_oracle = np.ascontiguousarray(snapshots), np.ascontiguousarray(witness_oracle)
    assert snapshots.shape[0] >= limit + 1 and witness_oracle.shape[0] >= limit
    out = C.POINTER(abi.VmInputStream)()
    nbytes = C.c_uint64()
    rc = lib.zkc_vm_encode_input_stream(ptr(snapshots), ptr(witness_oracle), limit, segment_cycles, C.byref(out), C.byref(nbytes))
    if rc:
        raise ZkcError(rc, what="zkc_vm_encode_input_stream")
    return VmInputStreamHandle(lib, out, nbytes.value)


def vm_decode_input_stream(stream: VmInputStreamHandle):
    """reference decoder of the stream format (numpy; tests): returns (state_words [294, limit + 1], witness_words [44, limit])"""
    st = stream.struct
    limit = int(st.limit)
    state = np.zeros((abi.VM_STATE_WORDS, limit + 1), dtype=np.uint32)
    wit = np.zeros((abi.VM_WITNESS_WORDS, limit), dtype=np.uint32)
    for k in range(st.n_segments):
        seg = st.segments[k]
        raw = np.frombuffer((C.c_uint8 * seg.blob_bytes).from_address(seg.blob), dtype=np.uint8)
        h = abi.VmSegmentHeader.from_buffer_copy(raw[:C.sizeof(abi.VmSegmentHeader)].tobytes())
        assert h.magic == abi.VM_SEGMENT_MAGIC and h.blob_bytes == seg.blob_bytes and h.first_cycle == k * st.segment_cycles
        n, f = h.n_cycles, h.first_cycle
        arr = lambda off, cnt, dt: np.frombuffer(raw, dtype=dt, count=cnt, offset=off)
        dsw, dww = arr(h.off_dense_state_word, h.n_dense_state, "<u2"), arr(h.off_dense_witness_word, h.n_dense_witness, "<u2")
        state[dsw, f:f + n + 1] = arr(h.off_dense_state, h.n_dense_state * (n + 1), "<u4").reshape(h.n_dense_state, n + 1)
        wit[dww, f:f + n] = arr(h.off_dense_witness, h.n_dense_witness * n, "<u4").reshape(h.n_dense_witness, n)
        so = arr(h.off_sparse_state_offsets, abi.VM_STATE_WORDS + 1, "<u4")
        si, sv = arr(h.off_sparse_state_index, h.n_sparse_state, "<u4"), arr(h.off_sparse_state_value, h.n_sparse_state, "<u4")
        for w in range(abi.VM_STATE_WORDS):
            lo, hi = int(so[w]), int(so[w + 1])
            if hi > lo:
                assert si[lo] == 0
                runs = np.diff(np.append(si[lo:hi], n + 1).astype(np.int64))
                state[w, f:f + n + 1] = np.repeat(sv[lo:hi], runs)
        wo = arr(h.off_sparse_witness_offsets, abi.VM_WITNESS_WORDS + 1, "<u4")
        wi, wv = arr(h.off_sparse_witness_index, h.n_sparse_witness, "<u4"), arr(h.off_sparse_witness_value, h.n_sparse_witness, "<u4")
        for w in range(abi.VM_WITNESS_WORDS):
            lo, hi = int(wo[w]), int(wo[w + 1])
            if hi > lo:
                assert w not in dww
                wit[w, f + wi[lo:hi].astype(np.int64)] = wv[lo:hi]
    return state, wit


_PK_LAYOUT = None


def vm_packed_layout(lib):
    """(kind [276], slot [276], counts [6]) of the PACKED trace"""
    global _PK_LAYOUT
    if _PK_LAYOUT is None:
        kind, slot, counts = np.zeros(abi.VM_COLS["NUM_COLS"], np.uint8), np.zeros(abi.VM_COLS["NUM_COLS"], np.uint16), np.zeros(7, np.uint32)
        lib.zkc_vm_packed_layout(ptr(kind), ptr(slot), ptr(counts))
        _PK_LAYOUT = (kind, slot, counts)
    return _PK_LAYOUT


@dataclass
class VmPackedTraceBuffers:
    """host buffers of a zkc_vm_packed_trace (numpy views, optionally of pinned memory): cols8 [N8, rows] ... and the records"""
    cols8: np.ndarray
    cols16: np.ndarray
    cols32: np.ndarray
    cols64: np.ndarray
    aux_records: np.ndarray
    sponge_records: np.ndarray
    limb_records: np.ndarray
    n_aux_records: int = 0
    n_sponge_records: int = 0
    n_limb_records: int = 0

    @property
    def nbytes_used(self):
        return (self.cols8.nbytes + self.cols16.nbytes + self.cols32.nbytes + self.cols64.nbytes +
                self.n_aux_records * abi.VM_AUX_RECORD_DTYPE.itemsize + self.n_sponge_records * abi.VM_SPONGE_RECORD_DTYPE.itemsize +
                self.n_limb_records * abi.VM_LIMB_RECORD_DTYPE.itemsize)


def vm_packed_trace_buffers(engine: Engine, n: int, limit: int, aux_fraction=0.3, sponge_per_cycle=1.5, limb_per_cycle=1.0, alloc=None) -> VmPackedTraceBuffers:
    """alloc(shape, dtype) -> array (default numpy.zeros; pass a pinned allocator for asynchronous copies)"""
    kind, slot, counts = vm_packed_layout(engine.lib)
    rows = n * limit
    alloc = alloc or (lambda shape, dt: np.zeros(shape, dtype=dt))
    return VmPackedTraceBuffers(alloc((int(counts[0]), rows), np.uint8), alloc((int(counts[1]), rows), np.uint16), alloc((int(counts[2]), rows), np.uint32),
                                alloc((int(counts[3]), rows), np.uint64), alloc((int(rows * aux_fraction) + 64,), abi.VM_AUX_RECORD_DTYPE),
                                alloc((int(rows * sponge_per_cycle) + 64,), abi.VM_SPONGE_RECORD_DTYPE),
                                alloc((int(rows * limb_per_cycle) + 64 + n,), abi.VM_LIMB_RECORD_DTYPE))


def main_vm_entry_point_stream(engine: Engine, closed_form_inputs, isa: abi.VmIsa, streams, limit: int, callstack_witness=None,
                               out: Optional[VmPackedTraceBuffers] = None, compare_expected=False):
    """zkc_main_vm_entry_point_stream: host buffers in the transport forms.  streams: list of VmInputStreamHandle (one per
    instance); callstack_witness: numpy [n, capacity, 336]; out: VmPackedTraceBuffers or None.  Returns (commitments,
    closed forms, statuses, rc)."""
    n = len(closed_form_inputs)
    assert len(streams) == n
    ios = (abi.VmClosedForm * n)(*[abi.VmClosedForm.from_buffer_copy(bytes(c)) for c in closed_form_inputs])
    commitments = np.zeros((n, 4), dtype=np.uint64)
    statuses = (abi.Status * n)()
    opts = abi.VmOptions(int(compare_expected))
    sp = (C.POINTER(abi.VmInputStream) * n)(*[s.ptr for s in streams])
    pk = None
    if out is not None:
        opts.trace_layout = abi.VM_TRACE_PACKED
        pk = abi.VmPackedTrace(out.cols8.ctypes.data, out.cols16.ctypes.data, out.cols32.ctypes.data, out.cols64.ctypes.data,
                               out.aux_records.ctypes.data, len(out.aux_records), 0, out.sponge_records.ctypes.data, len(out.sponge_records), 0,
                               out.limb_records.ctypes.data, len(out.limb_records), 0)
    n_cw = _n_cw(callstack_witness)
    assert not n_cw or not on_device(callstack_witness)
    rc = engine.lib.zkc_main_vm_entry_point_stream(engine.h, C.cast(ios, C.c_void_p), n, C.byref(isa), C.cast(sp, C.c_void_p),
                                                   ptr(callstack_witness) if n_cw else None, n_cw, limit, C.byref(opts),
                                                   C.byref(pk) if pk is not None else None, ptr(commitments), C.cast(statuses, C.c_void_p))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, statuses[0], "main_vm_entry_point_stream")
    if out is not None:
        out.n_aux_records, out.n_sponge_records, out.n_limb_records = int(pk.n_aux_records), int(pk.n_sponge_records), int(pk.n_limb_records)
    return commitments, ios, statuses, rc


def vm_expand_packed_trace(lib, out: VmPackedTraceBuffers, n: int, limit: int) -> np.ndarray:
    """PACKED -> DENSE [n, 276, limit] uint64, bit-exactly (what the host shim does while it assigns the cells)"""
    kind, slot, counts = vm_packed_layout(lib)
    K = abi.VM_COLS
    rows = n * limit
    dense = np.zeros((n, K["NUM_COLS"], limit), dtype=np.uint64)
    blocks = {abi.VM_PK_U8: out.cols8, abi.VM_PK_U16: out.cols16, abi.VM_PK_U32: out.cols32, abi.VM_PK_U64: out.cols64}
    for c in range(K["NUM_COLS"]):
        if int(kind[c]) in blocks:
            dense[:, c, :] = blocks[int(kind[c])][int(slot[c])].reshape(n, limit)
    assert out.n_aux_records <= len(out.aux_records) and out.n_sponge_records <= len(out.sponge_records) and \
        out.n_limb_records <= len(out.limb_records), "record capacity exceeded"
    rec = out.sponge_records[:out.n_sponge_records]
    r, s_ = rec["row"].astype(np.int64), rec["slot"].astype(np.int64)
    flat = dense.transpose(1, 0, 2).reshape(K["NUM_COLS"], rows)  # [col, g] view
    flat[K["SPONGE_ENFORCE"] + s_, r] = 1
    for j in range(12):
        flat[K["SPONGE_FINAL"] + 12 * s_ + j, r] = rec["out"][:, j]
    # limb records: src0 memory operand / dst1 (zeros elsewhere), code word (held until the next record of its instance)
    limb = out.limb_records[:out.n_limb_records]
    lrow, lkind = limb["row"].astype(np.int64), limb["kind"]
    for k_, base, width in ((abi.VM_LIMB_SRC0_FROM_MEMORY, K["SRC0_FROM_MEMORY"], 9), (abi.VM_LIMB_DST1, K["DST1"], 9)):
        sel = lkind == k_
        flat[base:base + width, lrow[sel]] = limb["v"][sel][:, :width].T
    if rows:
        sel = lkind == abi.VM_LIMB_CODE_WORD
        cw_rows, cw_vals = lrow[sel], limb["v"][sel][:, :8]
        order = np.argsort(cw_rows, kind="stable")
        cw_rows, cw_vals = cw_rows[order], cw_vals[order]
        has = np.zeros(rows, dtype=np.int64)
        has[cw_rows] = np.arange(1, len(cw_rows) + 1)
        assert (has[::limit] > 0).all(), "row 0 of an instance without a code-word record"
        flat[K["CODE_WORD"]:K["CODE_WORD"] + 8, :] = cw_vals[np.maximum.accumulate(has) - 1].T
    aux = out.aux_records[:out.n_aux_records]
    aux = aux[np.argsort(aux["row"], kind="stable")]
    ar = aux["row"].astype(np.int64)
    flat[K["OP_AUX"]:K["OP_AUX"] + 48, ar] = aux["op_aux"].T
    # queue ends: the record at or before each row (row 0 of every instance has one)
    if rows:
        has = np.zeros(rows, dtype=np.int64)
        has[ar] = np.arange(1, len(ar) + 1)
        last = np.maximum.accumulate(has)
        assert (last > 0).all() and (has[::limit] > 0).all(), "row 0 of an instance without a record"
        flat[K["FORWARD_TAIL_OUT"]:K["FORWARD_TAIL_OUT"] + 10, :] = aux["queue_ends"][last - 1].T
    return np.ascontiguousarray(flat.reshape(K["NUM_COLS"], n, limit).transpose(1, 0, 2))
