"""Host-side mirror of `linear_hasher_entry_point` (/root/reference/src/linear_hasher/mod.rs:35-214)."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import abi
from .engine import Engine, ZkcError, on_device, ptr
from .log_sorter import SorterResult


@dataclass
class LinearHasherCircuitInstanceWitness:
    """linear_hasher/input.rs:74-87; the CircuitQueueRawWitness deque as struct-of-arrays"""
    closed_form_input: abi.LinearHasherClosedForm
    queue_witness: object  # [n] LOG_QUERY_DTYPE or torch uint8 [n, 128]
    queue_prev_tails: object  # [n, 4] uint64
    # optional hint: the keccak state (25 lanes) after every cycle, [limit, 25] uint64 (what an out-of-circuit run of the hasher holds)
    keccak_states: Optional[object] = None


def linear_hasher_entry_point(engine: Engine, witness: LinearHasherCircuitInstanceWitness, limit: int, want_trace=True,
                              compare_expected=False, raise_on_unsatisfied=True, trace_out=None) -> SorterResult:
    w = witness
    dev = on_device(w.queue_witness, w.queue_prev_tails, w.keccak_states)
    if trace_out is not None:
        dev |= 2 * on_device(trace_out)
    elif dev:
        dev = 3
    trace = trace_out
    if want_trace and trace is None:
        if dev & 2:
            import torch
            trace = torch.empty((abi.LH_COLS["NUM_COLS"], limit), dtype=torch.int64, device=w.queue_witness.device)
        else:
            trace = np.empty((abi.LH_COLS["NUM_COLS"], limit), dtype=np.uint64)
    if w.keccak_states is not None:
        assert len(w.keccak_states) >= limit, "keccak_states holds the state after EVERY cycle: [limit, 25]"
    io = abi.LinearHasherClosedForm.from_buffer_copy(bytes(w.closed_form_input))
    opts = abi.SorterOptions(int(compare_expected))
    commitment = np.zeros(4, dtype=np.uint64)
    st = abi.Status()
    rc = engine.lib.zkc_linear_hasher_entry_point(engine.h, C.byref(io), ptr(w.queue_witness), ptr(w.queue_prev_tails), len(w.queue_witness),
                                                  ptr(w.keccak_states), limit, C.byref(opts), dev, ptr(trace), ptr(commitment), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE) or (rc and raise_on_unsatisfied):
        raise ZkcError(rc, st, "linear_hasher_entry_point")
    return SorterResult(commitment, io, trace, st)


def linear_hasher_check_trace(engine: Engine, closed_form_input: abi.LinearHasherClosedForm, trace, limit: int, gates: int = 0):
    """Constraint evaluation of a finished linear_hasher trace [LH_COLS.NUM_COLS, limit] (numpy: host, torch CUDA: device): every
    relation of the loop of linear_hasher_entry_point (mod.rs:103-171), the keccak sponge included.  Returns (violating rows, status);
    status.failed_checks holds abi.LHV bits."""
    st = abi.Status()
    viol = C.c_uint64()
    io = abi.LinearHasherClosedForm.from_buffer_copy(bytes(closed_form_input))
    rc = engine.lib.zkc_linear_hasher_check_trace(engine.h, C.byref(io), ptr(trace), limit, gates, on_device(trace), C.byref(viol), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "linear_hasher_check_trace")
    return viol.value, st
