"""Host-side mirror of `unpack_code_into_memory_entry_point` (/root/reference/src/code_unpacker_sha256/mod.rs:33-148)."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import abi
from .engine import Engine, ZkcError, on_device, ptr
from .log_sorter import SorterResult


@dataclass
class CodeDecommitterCircuitInstanceWitness:
    """code_unpacker_sha256/input.rs:152-160: the requests queue's raw witness as struct-of-arrays + the code words"""
    closed_form_input: abi.CodeUnpackerClosedForm
    sorted_requests_queue_witness: object  # [n] DECOMMIT_QUERY_DTYPE or torch uint8 [n, 48]
    sorted_requests_queue_prev_states: object  # [n, 12] uint64
    code_words: object  # [total_words, 8] uint32 little-endian limbs, flattened over the requests in pop order
    memory_queue_states: Optional[object] = None  # [pushes, 12]: memory queue tail after each write (optional hint)


def unpack_code_into_memory_entry_point(engine: Engine, witness: CodeDecommitterCircuitInstanceWitness, limit: int, want_trace=True,
                                        compare_expected=False, raise_on_unsatisfied=True, trace_out=None) -> SorterResult:
    w = witness
    dev = on_device(w.sorted_requests_queue_witness, w.sorted_requests_queue_prev_states, w.code_words, w.memory_queue_states)
    if trace_out is not None:
        dev |= 2 * on_device(trace_out)
    elif dev:
        dev = 3
    trace = trace_out
    if want_trace and trace is None:
        if dev & 2:
            import torch
            trace = torch.empty((abi.CU_COLS["NUM_COLS"], limit), dtype=torch.int64, device=w.sorted_requests_queue_witness.device)
        else:
            trace = np.empty((abi.CU_COLS["NUM_COLS"], limit), dtype=np.uint64)
    words = w.code_words
    if not (dev & 1):
        words = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1, 8)
    io = abi.CodeUnpackerClosedForm.from_buffer_copy(bytes(w.closed_form_input))
    opts = abi.SorterOptions(int(compare_expected))
    commitment = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    n_states = 0 if w.memory_queue_states is None else len(w.memory_queue_states)
    rc = engine.lib.zkc_code_unpacker_entry_point(
        engine.h, C.byref(io), ptr(w.sorted_requests_queue_witness), ptr(w.sorted_requests_queue_prev_states),
        len(w.sorted_requests_queue_witness), ptr(words), len(words), ptr(w.memory_queue_states), n_states, limit, C.byref(opts), dev,
        ptr(trace), ptr(commitment), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE) or (rc and raise_on_unsatisfied):
        raise ZkcError(rc, st, "unpack_code_into_memory_entry_point")
    return SorterResult(commitment, io, trace, st)


def code_unpacker_check_trace(engine: Engine, closed_form_input: abi.CodeUnpackerClosedForm, trace, limit: int, gates: int = 0):
    """Constraint evaluation of a finished code_unpacker_sha256 trace [CU_COLS.NUM_COLS, limit] (numpy: host, torch CUDA: device):
    every relation of unpack_code_into_memory_inner (mod.rs:191-447), the compression and the hash comparison included.  Returns
    (violating rows, status); status.failed_checks holds abi.CUV bits."""
    st = abi.Status()
    viol = C.c_uint64()
    io = abi.CodeUnpackerClosedForm.from_buffer_copy(bytes(closed_form_input))
    rc = engine.lib.zkc_code_unpacker_check_trace(engine.h, C.byref(io), ptr(trace), limit, gates, on_device(trace), C.byref(viol), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "code_unpacker_check_trace")
    return viol.value, st
