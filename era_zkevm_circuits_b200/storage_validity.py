"""Host-side mirror of `sort_and_deduplicate_storage_access_entry_point`
(/root/reference/src/storage_validity_by_grand_product/mod.rs:166-507)."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import abi
from .engine import Engine, ZkcError, check_hint_rows, on_device, ptr
from .log_sorter import SorterResult



@dataclass
class StorageDeduplicatorInstanceWitness:
    """storage_validity_by_grand_product/input.rs:128-136; queue witnesses as struct-of-arrays.  The sorted
    queue holds TimestampedStorageLogRecord: the LogQuery record plus its queue-position timestamp."""
    closed_form_input: abi.StorageClosedForm
    unsorted_queue_witness: object  # [n] LOG_QUERY_DTYPE or torch uint8 [n, 128]
    unsorted_queue_prev_tails: object  # [n, 4] uint64
    intermediate_sorted_queue_witness: object
    intermediate_sorted_queue_timestamps: object  # [n] uint32
    intermediate_sorted_queue_prev_tails: object
    result_queue_tails: Optional[object] = None


def sort_and_deduplicate_storage_access_entry_point(engine: Engine, witness: StorageDeduplicatorInstanceWitness, limit: int,
                                                    want_trace=True, compare_expected=False, raise_on_unsatisfied=True,
                                                    trace_out=None) -> SorterResult:
    w = witness
    check_hint_rows("sort_and_deduplicate_storage_access_entry_point", w.unsorted_queue_witness, w.unsorted_queue_prev_tails)
    check_hint_rows("sort_and_deduplicate_storage_access_entry_point", w.intermediate_sorted_queue_witness, w.intermediate_sorted_queue_prev_tails,
                    w.intermediate_sorted_queue_timestamps)
    dev = on_device(w.unsorted_queue_witness, w.intermediate_sorted_queue_witness, w.unsorted_queue_prev_tails,
                    w.intermediate_sorted_queue_prev_tails, w.intermediate_sorted_queue_timestamps, w.result_queue_tails)
    if trace_out is not None:
        dev |= 2 * on_device(trace_out)
    elif dev:
        dev = 3
    trace = trace_out
    if want_trace and trace is None:
        if dev & 2:
            import torch
            trace = torch.empty((abi.ST_COLS["NUM_COLS"], limit), dtype=torch.int64, device=w.unsorted_queue_witness.device)
        else:
            trace = np.empty((abi.ST_COLS["NUM_COLS"], limit), dtype=np.uint64)
    io = abi.StorageClosedForm.from_buffer_copy(bytes(w.closed_form_input))
    opts = abi.SorterOptions(int(compare_expected))
    commitment = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    ts = w.intermediate_sorted_queue_timestamps
    if ts is not None and not (dev & 1):
        ts = np.ascontiguousarray(ts, dtype=np.uint32)
    n_tails = 0 if w.result_queue_tails is None else len(w.result_queue_tails)
    rc = engine.lib.zkc_storage_validity_entry_point(
        engine.h, C.byref(io), ptr(w.unsorted_queue_witness), ptr(w.unsorted_queue_prev_tails), len(w.unsorted_queue_witness),
        ptr(w.intermediate_sorted_queue_witness), ptr(ts), ptr(w.intermediate_sorted_queue_prev_tails),
        len(w.intermediate_sorted_queue_witness), ptr(w.result_queue_tails), n_tails, limit, C.byref(opts), dev, ptr(trace),
        ptr(commitment), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE) or (rc and raise_on_unsatisfied):
        raise ZkcError(rc, st, "sort_and_deduplicate_storage_access_entry_point")
    return SorterResult(commitment, io, trace, st)


def storage_validity_check_trace(engine: Engine, closed_form_input: abi.StorageClosedForm, trace, limit: int, gates: int = 0):
    """Constraint evaluation of a finished storage_validity trace [ST_COLS.NUM_COLS, limit] (numpy: host, torch CUDA: device):
    every row-local relation of sort_and_deduplicate_storage_access_inner (mod.rs:560-800), the per-cell state machine
    included.  Returns (violating rows, status); status.failed_checks holds abi.STV bits."""
    st = abi.Status()
    viol = C.c_uint64()
    io = abi.StorageClosedForm.from_buffer_copy(bytes(closed_form_input))
    rc = engine.lib.zkc_storage_validity_check_trace(engine.h, C.byref(io), ptr(trace), limit, gates, on_device(trace), C.byref(viol), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "storage_validity_check_trace")
    return viol.value, st
