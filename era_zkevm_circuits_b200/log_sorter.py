"""Host-side mirror of `sort_and_deduplicate_events_entry_point`
(/root/reference/src/log_sorter/mod.rs:34-232)."""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import abi
from .engine import Engine, ZkcError, check_hint_rows, on_device, ptr


@dataclass
class EventsDeduplicatorInstanceWitness:
    """log_sorter/input.rs:98-106; the two CircuitQueueRawWitness deques as struct-of-arrays"""
    closed_form_input: abi.EventsClosedForm
    initial_queue_witness: object  # [n] LOG_QUERY_DTYPE or torch uint8 [n, 128]
    initial_queue_prev_tails: object  # [n, 4] uint64
    intermediate_sorted_queue_witness: object
    intermediate_sorted_queue_prev_tails: object
    result_queue_tails: Optional[object] = None  # [pushes, 4]: tail after every executed push (optional hint)


@dataclass
class SorterResult:
    commitment: np.ndarray
    closed_form_input: object
    trace: Optional[object]
    status: abi.Status = field(default_factory=abi.Status)


def sort_and_deduplicate_events_entry_point(engine: Engine, witness: EventsDeduplicatorInstanceWitness, limit: int,
                                            want_trace=True, compare_expected=False, raise_on_unsatisfied=True,
                                            trace_out=None) -> SorterResult:
    w = witness
    check_hint_rows("sort_and_deduplicate_events_entry_point", w.initial_queue_witness, w.initial_queue_prev_tails)
    check_hint_rows("sort_and_deduplicate_events_entry_point", w.intermediate_sorted_queue_witness, w.intermediate_sorted_queue_prev_tails)
    dev = on_device(w.initial_queue_witness, w.intermediate_sorted_queue_witness, w.initial_queue_prev_tails,
                    w.intermediate_sorted_queue_prev_tails, w.result_queue_tails)
    if trace_out is not None:
        dev |= 2 * on_device(trace_out)
    elif dev:
        dev = 3
    trace = trace_out
    if want_trace and trace is None:
        if dev & 2:
            import torch
            trace = torch.empty((abi.EV_COLS["NUM_COLS"], limit), dtype=torch.int64, device=w.initial_queue_witness.device)
        else:
            trace = np.empty((abi.EV_COLS["NUM_COLS"], limit), dtype=np.uint64)
    io = abi.EventsClosedForm.from_buffer_copy(bytes(w.closed_form_input))
    opts = abi.SorterOptions(int(compare_expected))
    commitment = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    n_tails = 0 if w.result_queue_tails is None else len(w.result_queue_tails)
    rc = engine.lib.zkc_log_sorter_entry_point(
        engine.h, C.byref(io), ptr(w.initial_queue_witness), ptr(w.initial_queue_prev_tails), len(w.initial_queue_witness),
        ptr(w.intermediate_sorted_queue_witness), ptr(w.intermediate_sorted_queue_prev_tails),
        len(w.intermediate_sorted_queue_witness), ptr(w.result_queue_tails), n_tails, limit, C.byref(opts), dev,
        ptr(trace), ptr(commitment), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE) or (rc and raise_on_unsatisfied):
        raise ZkcError(rc, st, "sort_and_deduplicate_events_entry_point")
    return SorterResult(commitment, io, trace, st)


def log_sorter_check_trace(engine: Engine, closed_form_input: abi.EventsClosedForm, trace, limit: int, gates: int = 0):
    """Constraint evaluation of a finished log_sorter trace [EV_COLS.NUM_COLS, limit] (numpy: host, torch CUDA: device): every
    row-local relation of repack_and_prove_events_rollbacks_inner.  Returns (violating rows, status)."""
    st = abi.Status()
    viol = C.c_uint64()
    io = abi.EventsClosedForm.from_buffer_copy(bytes(closed_form_input))
    rc = engine.lib.zkc_log_sorter_check_trace(engine.h, C.byref(io), ptr(trace), limit, gates, on_device(trace), C.byref(viol), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "log_sorter_check_trace")
    return viol.value, st
