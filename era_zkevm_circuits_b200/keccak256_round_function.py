"""Host-side mirror of `keccak256_round_function_entry_point`
(/root/reference/src/keccak256_round_function/mod.rs:673-794)."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import abi
from .engine import Engine, ZkcError, on_device, ptr
from .log_sorter import SorterResult


@dataclass
class Keccak256RoundFunctionCircuitInstanceWitness:
    """keccak256_round_function/input.rs:92-99"""
    closed_form_input: abi.KeccakClosedForm
    requests_queue_witness: object  # [n] LOG_QUERY_DTYPE (precompile calls in pop order)
    requests_queue_prev_tails: object  # [n, 4] uint64
    memory_reads_witness: object  # [n_words, 8] uint32 little-endian limbs, FIFO order
    memory_queue_states: Optional[object] = None  # [pushes, 12]: memory queue tail after each push (optional hint)


def keccak256_round_function_entry_point(engine: Engine, witness: Keccak256RoundFunctionCircuitInstanceWitness, limit: int,
                                         want_trace=True, compare_expected=False, raise_on_unsatisfied=True,
                                         trace_out=None) -> SorterResult:
    w = witness
    dev = on_device(w.requests_queue_witness, w.requests_queue_prev_tails, w.memory_reads_witness, w.memory_queue_states)
    if trace_out is not None:
        dev |= 2 * on_device(trace_out)
    elif dev:
        dev = 3
    trace = trace_out
    if want_trace and trace is None:
        if dev & 2:
            import torch
            trace = torch.empty((abi.KC_COLS["NUM_COLS"], limit), dtype=torch.int64, device=w.requests_queue_witness.device)
        else:
            trace = np.empty((abi.KC_COLS["NUM_COLS"], limit), dtype=np.uint64)
    reads = w.memory_reads_witness
    if not (dev & 1):
        reads = np.ascontiguousarray(reads, dtype=np.uint32).reshape(-1, 8)
    io = abi.KeccakClosedForm.from_buffer_copy(bytes(w.closed_form_input))
    opts = abi.PrecompileOptions(int(compare_expected), 0, 0, 0)
    commitment = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    n_states = 0 if w.memory_queue_states is None else len(w.memory_queue_states)
    rc = engine.lib.zkc_keccak256_round_function_entry_point(
        engine.h, C.byref(io), ptr(w.requests_queue_witness), ptr(w.requests_queue_prev_tails), len(w.requests_queue_witness),
        ptr(reads), len(reads), ptr(w.memory_queue_states), n_states, limit, C.byref(opts), dev, ptr(trace), ptr(commitment),
        C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE) or (rc and raise_on_unsatisfied):
        raise ZkcError(rc, st, "keccak256_round_function_entry_point")
    return SorterResult(commitment, io, trace, st)


def keccak256_round_function_check_trace(engine: Engine, closed_form_input: abi.KeccakClosedForm, trace, limit: int, gates: int = 0,
                                         options: Optional[abi.PrecompileOptions] = None):
    """Constraint evaluation of a finished keccak256_round_function trace [KC_COLS.NUM_COLS, limit] (numpy: host, torch CUDA: device):
    the cycle function re-run on every cycle from the previous cycle's cells and compared cell by cell, the pop and the memory queue
    bookkeeping.  Returns (violating rows, status); status.failed_checks holds abi.KCV bits."""
    st = abi.Status()
    viol = C.c_uint64()
    io = abi.KeccakClosedForm.from_buffer_copy(bytes(closed_form_input))
    opts = abi.PrecompileOptions.from_buffer_copy(bytes(options)) if options is not None else abi.PrecompileOptions()
    rc = engine.lib.zkc_keccak256_round_function_check_trace(engine.h, C.byref(io), C.byref(opts), ptr(trace), limit, gates, on_device(trace),
                                                             C.byref(viol), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "keccak256_round_function_check_trace")
    return viol.value, st
