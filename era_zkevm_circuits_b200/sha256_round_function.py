"""Host-side mirror of `sha256_round_function_entry_point`
(/root/reference/src/sha256_round_function/mod.rs:343-470)."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import abi
from .engine import Engine, ZkcError, on_device, ptr
from .log_sorter import SorterResult


@dataclass
class Sha256RoundFunctionCircuitInstanceWitness:
    """sha256_round_function/input.rs (same shape as the keccak witness)"""
    closed_form_input: abi.Sha256ClosedForm
    requests_queue_witness: object
    requests_queue_prev_tails: object
    memory_reads_witness: object  # [2 * rounds, 8] uint32 little-endian limbs of the big-endian memory words
    memory_queue_states: Optional[object] = None


def sha256_round_function_entry_point(engine: Engine, witness: Sha256RoundFunctionCircuitInstanceWitness, limit: int,
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
            trace = torch.empty((abi.SH_COLS["NUM_COLS"], limit), dtype=torch.int64, device=w.requests_queue_witness.device)
        else:
            trace = np.empty((abi.SH_COLS["NUM_COLS"], limit), dtype=np.uint64)
    reads = w.memory_reads_witness
    if not (dev & 1):
        reads = np.ascontiguousarray(reads, dtype=np.uint32).reshape(-1, 8)
    io = abi.Sha256ClosedForm.from_buffer_copy(bytes(w.closed_form_input))
    opts = abi.PrecompileOptions(int(compare_expected), 0, 0, 0)
    commitment = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    n_states = 0 if w.memory_queue_states is None else len(w.memory_queue_states)
    rc = engine.lib.zkc_sha256_round_function_entry_point(
        engine.h, C.byref(io), ptr(w.requests_queue_witness), ptr(w.requests_queue_prev_tails), len(w.requests_queue_witness),
        ptr(reads), len(reads), ptr(w.memory_queue_states), n_states, limit, C.byref(opts), dev, ptr(trace), ptr(commitment),
        C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE) or (rc and raise_on_unsatisfied):
        raise ZkcError(rc, st, "sha256_round_function_entry_point")
    return SorterResult(commitment, io, trace, st)


def sha256_round_function_check_trace(engine: Engine, closed_form_input: abi.Sha256ClosedForm, trace, limit: int, gates: int = 0,
                                      options: Optional[abi.PrecompileOptions] = None):
    """Constraint evaluation of a finished sha256_round_function trace [SH_COLS.NUM_COLS, limit] (numpy: host, torch CUDA: device):
    every relation of sha256_precompile_inner (mod.rs:146-330), the compression included.  Returns (violating rows, status);
    status.failed_checks holds abi.SHV bits."""
    st = abi.Status()
    viol = C.c_uint64()
    io = abi.Sha256ClosedForm.from_buffer_copy(bytes(closed_form_input))
    opts = abi.PrecompileOptions.from_buffer_copy(bytes(options)) if options is not None else abi.PrecompileOptions()
    rc = engine.lib.zkc_sha256_round_function_check_trace(engine.h, C.byref(io), C.byref(opts), ptr(trace), limit, gates, on_device(trace),
                                                          C.byref(viol), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "sha256_round_function_check_trace")
    return viol.value, st
