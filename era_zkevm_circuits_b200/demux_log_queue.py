"""Host-side mirror of `demultiplex_storage_logs_enty_point` (sic; /root/reference/src/demux_log_queue/mod.rs:38-217)."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import abi
from .engine import Engine, ZkcError, check_hint_rows, on_device, ptr
from .log_sorter import SorterResult


@dataclass
class LogDemuxerCircuitInstanceWitness:
    """demux_log_queue/input.rs:124-129; the CircuitQueueRawWitness deque as struct-of-arrays"""
    closed_form_input: abi.DemuxClosedForm
    initial_queue_witness: object  # [n] LOG_QUERY_DTYPE or torch uint8 [n, 128]
    initial_queue_prev_tails: object  # [n, 4] uint64
    # optional hint: the six output queues' tails after each of their pushes, concatenated queue after queue
    # ([sum(counts), 4]) + the six counts
    output_queue_tails: Optional[object] = None
    output_queue_counts: Optional[Sequence[int]] = None


def demultiplex_storage_logs_enty_point(engine: Engine, witness: LogDemuxerCircuitInstanceWitness, limit: int, want_trace=True,
                                        compare_expected=False, raise_on_unsatisfied=True, trace_out=None,
                                        options: Optional[abi.DemuxOptions] = None) -> SorterResult:
    w = witness
    check_hint_rows("demultiplex_storage_logs_enty_point", w.initial_queue_witness, w.initial_queue_prev_tails)
    dev = on_device(w.initial_queue_witness, w.initial_queue_prev_tails, w.output_queue_tails)
    if trace_out is not None:
        dev |= 2 * on_device(trace_out)
    elif dev:
        dev = 3
    trace = trace_out
    if want_trace and trace is None:
        if dev & 2:
            import torch
            trace = torch.empty((abi.DMX_COLS["NUM_COLS"], limit), dtype=torch.int64, device=w.initial_queue_witness.device)
        else:
            trace = np.empty((abi.DMX_COLS["NUM_COLS"], limit), dtype=np.uint64)
    io = abi.DemuxClosedForm.from_buffer_copy(bytes(w.closed_form_input))
    opts = abi.DemuxOptions.from_buffer_copy(bytes(options)) if options is not None else abi.DemuxOptions()
    opts.compare_expected = int(compare_expected)
    commitment = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    counts = (C.c_size_t * 6)(*([0] * 6 if w.output_queue_tails is None else [int(c) for c in w.output_queue_counts]))
    if w.output_queue_tails is not None:
        assert sum(counts) == len(w.output_queue_tails), "output_queue_counts must add up to the rows of output_queue_tails"
    rc = engine.lib.zkc_demux_log_queue_entry_point(
        engine.h, C.byref(io), ptr(w.initial_queue_witness), ptr(w.initial_queue_prev_tails), len(w.initial_queue_witness),
        ptr(w.output_queue_tails), counts, limit, C.byref(opts), dev, ptr(trace), ptr(commitment), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE) or (rc and raise_on_unsatisfied):
        raise ZkcError(rc, st, "demultiplex_storage_logs_enty_point")
    return SorterResult(commitment, io, trace, st)


def demux_log_queue_check_trace(engine: Engine, closed_form_input: abi.DemuxClosedForm, trace, limit: int, gates: int = 0,
                                options: Optional[abi.DemuxOptions] = None):
    """Constraint evaluation of a finished demux_log_queue trace [DMX_COLS.NUM_COLS, limit] (numpy: host, torch CUDA: device): every
    row-local relation of demultiplex_storage_logs_inner and push_with_optimize (mod.rs:268-447).  Returns (violating rows, status);
    status.failed_checks holds abi.DMXV bits."""
    st = abi.Status()
    viol = C.c_uint64()
    io = abi.DemuxClosedForm.from_buffer_copy(bytes(closed_form_input))
    opts = abi.DemuxOptions.from_buffer_copy(bytes(options)) if options is not None else abi.DemuxOptions()
    rc = engine.lib.zkc_demux_log_queue_check_trace(engine.h, C.byref(io), C.byref(opts), ptr(trace), limit, gates, on_device(trace), C.byref(viol), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "demux_log_queue_check_trace")
    return viol.value, st
