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


def main_vm_check_trace(engine: Engine, isa: abi.VmIsa, trace, limit: int, n_instances: int = 1):
    """Constraint evaluation of finished main_vm traces (DENSE layout): the row-local relations of vm_cycle.  trace:
    [NUM_COLS, limit] or [n, NUM_COLS, limit] uint64 (numpy: host, torch CUDA: device).  Returns (violating rows, status)."""
    st = abi.Status()
    viol = C.c_uint64()
    rc = engine.lib.zkc_main_vm_check_trace(engine.h, C.byref(isa), ptr(trace), limit, n_instances, on_device(trace), C.byref(viol), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "main_vm_check_trace")
    return viol.value, st
