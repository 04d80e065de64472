"""Host-side mirror of `sort_and_deduplicate_code_decommittments_entry_point`
(/root/reference/src/sort_decommittment_requests/mod.rs:40-233)."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import abi
from .engine import Engine, ZkcError, check_hint_rows, on_device, ptr
from .log_sorter import SorterResult


@dataclass
class CodeDecommittmentsDeduplicatorInstanceWitness:
    """sort_decommittment_requests/input.rs:114-131; the two FullStateCircuitQueueRawWitness deques as struct-of-arrays"""
    closed_form_input: abi.DecommitSorterClosedForm
    initial_queue_witness: object  # [n] DECOMMIT_QUERY_DTYPE or torch uint8 [n, 48]
    initial_queue_prev_states: object  # [n, 12] uint64: queue state before each record was pushed
    sorted_queue_witness: object
    sorted_queue_prev_states: object
    result_queue_states: Optional[object] = None  # [pushes, 12]: result-queue tail after every executed push (optional hint)


def sort_and_deduplicate_code_decommittments_entry_point(engine: Engine, witness: CodeDecommittmentsDeduplicatorInstanceWitness,
                                                         limit: int, want_trace=True, compare_expected=False,
                                                         raise_on_unsatisfied=True, trace_out=None) -> SorterResult:
    w = witness
    check_hint_rows("sort_and_deduplicate_code_decommittments_entry_point", w.initial_queue_witness, w.initial_queue_prev_states)
    check_hint_rows("sort_and_deduplicate_code_decommittments_entry_point", w.sorted_queue_witness, w.sorted_queue_prev_states)
    dev = on_device(w.initial_queue_witness, w.sorted_queue_witness, w.initial_queue_prev_states, w.sorted_queue_prev_states,
                    w.result_queue_states)
    if trace_out is not None:
        dev |= 2 * on_device(trace_out)
    elif dev:
        dev = 3
    trace = trace_out
    if want_trace and trace is None:
        if dev & 2:
            import torch
            trace = torch.empty((abi.DQ_COLS["NUM_COLS"], limit), dtype=torch.int64, device=w.initial_queue_witness.device)
        else:
            trace = np.empty((abi.DQ_COLS["NUM_COLS"], limit), dtype=np.uint64)
    io = abi.DecommitSorterClosedForm.from_buffer_copy(bytes(w.closed_form_input))
    opts = abi.SorterOptions(int(compare_expected))
    commitment = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    n_states = 0 if w.result_queue_states is None else len(w.result_queue_states)
    rc = engine.lib.zkc_sort_decommittments_entry_point(
        engine.h, C.byref(io), ptr(w.initial_queue_witness), ptr(w.initial_queue_prev_states), len(w.initial_queue_witness),
        ptr(w.sorted_queue_witness), ptr(w.sorted_queue_prev_states), len(w.sorted_queue_witness),
        ptr(w.result_queue_states), n_states, limit, C.byref(opts), dev, ptr(trace), ptr(commitment), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE) or (rc and raise_on_unsatisfied):
        raise ZkcError(rc, st, "sort_and_deduplicate_code_decommittments_entry_point")
    return SorterResult(commitment, io, trace, st)


def sort_decommittments_check_trace(engine: Engine, closed_form_input: abi.DecommitSorterClosedForm, trace, limit: int, gates: int = 0):
    """Constraint evaluation of a finished sort_decommittment_requests trace [DQ_COLS.NUM_COLS, limit] (numpy: host, torch CUDA:
    device): every row-local relation of sort_and_deduplicate_code_decommittments_inner (mod.rs:235-381).  Returns (violating rows,
    status); status.failed_checks holds abi.DQV bits."""
    st = abi.Status()
    viol = C.c_uint64()
    io = abi.DecommitSorterClosedForm.from_buffer_copy(bytes(closed_form_input))
    rc = engine.lib.zkc_sort_decommittments_check_trace(engine.h, C.byref(io), ptr(trace), limit, gates, on_device(trace), C.byref(viol), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "sort_decommittments_check_trace")
    return viol.value, st
