"""Host-side mirror of `ram_permutation_entry_point`
(/root/reference/src/ram_permutation/mod.rs:31-210): same argument meaning
(`closed_form_input_witness`, `limit`), returns the 4-element input commitment, and raises where
the reference would panic / produce an unsatisfiable system."""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import abi
from .engine import Engine, ZkcError, check_hint_rows, on_device, ptr


@dataclass
class RamPermutationCircuitInstanceWitness:
    """ram_permutation/input.rs:99-116.  The two `FullStateCircuitQueueRawWitness` deques are kept
    as struct-of-arrays: the records in pop order and, per record, the queue state before its push."""
    closed_form_input: abi.RamClosedForm
    unsorted_queue_witness: object  # [n] MEMORY_QUERY_DTYPE (numpy) or torch uint8 [n, 64] on the GPU
    unsorted_queue_prev_states: Optional[object]  # [n, 12] uint64 or None
    sorted_queue_witness: object
    sorted_queue_prev_states: Optional[object]


@dataclass
class RamPermutationResult:
    commitment: np.ndarray  # [4] uint64: the public inputs
    closed_form_input: abi.RamClosedForm  # completion_flag and hidden_fsm_output filled in
    trace: Optional[object]  # [ZKC_RAM_NUM_COLS, limit] uint64, column-major witness
    status: abi.Status = field(default_factory=abi.Status)


def ram_permutation_entry_point(engine: Engine, witness: RamPermutationCircuitInstanceWitness, limit: int,
                                want_trace=True, compare_expected=False, bootloader_heap_page=0,
                                raise_on_unsatisfied=True, trace_out=None) -> RamPermutationResult:
    w = witness
    check_hint_rows("ram_permutation_entry_point", w.unsorted_queue_witness, w.unsorted_queue_prev_states)
    check_hint_rows("ram_permutation_entry_point", w.sorted_queue_witness, w.sorted_queue_prev_states)
    dev = on_device(w.unsorted_queue_witness, w.sorted_queue_witness, w.unsorted_queue_prev_states,
                    w.sorted_queue_prev_states)
    if trace_out is not None:
        dev |= 2 * on_device(trace_out)
    elif dev:
        dev = 3
    n_u, n_s = len(w.unsorted_queue_witness), len(w.sorted_queue_witness)
    trace = trace_out
    if want_trace and trace is None:
        if dev & 2:
            import torch
            trace = torch.empty((abi.RAM_COLS["NUM_COLS"], limit), dtype=torch.int64,
                                device=w.unsorted_queue_witness.device)
        else:
            trace = np.empty((abi.RAM_COLS["NUM_COLS"], limit), dtype=np.uint64)
    io = abi.RamClosedForm.from_buffer_copy(bytes(w.closed_form_input))
    opts = abi.RamOptions(bootloader_heap_page, int(compare_expected))
    commitment = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    rc = engine.lib.zkc_ram_permutation_entry_point(
        engine.h, C.byref(io), ptr(w.unsorted_queue_witness), ptr(w.unsorted_queue_prev_states), n_u,
        ptr(w.sorted_queue_witness), ptr(w.sorted_queue_prev_states), n_s, limit, C.byref(opts), dev,
        ptr(trace), ptr(commitment), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE) or (rc and raise_on_unsatisfied):
        raise ZkcError(rc, st, "ram_permutation_entry_point")
    return RamPermutationResult(commitment, io, trace, st)


def ram_permutation_check_trace(engine: Engine, closed_form_input: abi.RamClosedForm, trace, limit: int, gates=0,
                                bootloader_heap_page=0):
    """Constraint evaluation of a finished witness trace (the reference's `check_if_satisfied`,
    ram_permutation/mod.rs:556).  Returns (violating_rows, status)."""
    st = abi.Status()
    viol = C.c_uint64()
    opts = abi.RamOptions(bootloader_heap_page, 0)
    io = abi.RamClosedForm.from_buffer_copy(bytes(closed_form_input))
    rc = engine.lib.zkc_ram_permutation_check_trace(engine.h, C.byref(io), ptr(trace), limit, C.byref(opts), gates,
                                                    on_device(trace), C.byref(viol), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "ram_permutation_check_trace")
    return viol.value, st
