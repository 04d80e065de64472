"""Engine: one context per GPU / process over the C ABI (include/zkc_b200.h).

torch is plumbing only: callers may hand in torch CUDA tensors (device-resident inputs, on_device=1)
and bind the engine to torch's current stream; numpy arrays are treated as host buffers and the
library does the H2D/D2H copies itself.
"""
import ctypes as C

import numpy as np

from . import abi


class ZkcError(RuntimeError):
    def __init__(self, code, status=None, what=""):
        self.code = code
        self.status = status
        msg = f"{what}: {abi.CODE_NAMES.get(code, code)}"
        if status is not None:
            msg += f" (first_bad_row={status.first_bad_row}, failed_checks=0x{status.failed_checks:x}, cuda={status.cuda_error})"
        super().__init__(msg)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def ptr(x):
    """raw pointer of a numpy array / torch tensor / None"""
    if x is None:
        return None
    if _is_torch(x):
        assert x.is_contiguous()
        return C.c_void_p(x.data_ptr())
    assert x.flags["C_CONTIGUOUS"]
    return C.c_void_p(x.ctypes.data)


def check_hint_rows(what, records, *hints):
    """the C ABI takes queue-state hints without a length: they must cover every record (a short hint array would be
    read out of bounds)"""
    n = 0 if records is None else len(records)
    for h in hints:
        if h is not None and len(h) < n:
            raise ValueError(f"{what}: a queue-state hint array has {len(h)} rows for {n} records")


def on_device(*xs):
    flags = {bool(x.is_cuda) if _is_torch(x) else False for x in xs if x is not None}
    if len(flags) > 1:
        raise ValueError("mixing host and device buffers in one call")
    return int(flags.pop()) if flags else 0


class Engine:
    def __init__(self, device=0, stream=None):
        self.lib = abi.load_library()
        h = C.c_void_p()
        rc = self.lib.zkc_create(device, C.byref(h))
        if rc != abi.ZKC_OK:
            raise ZkcError(rc, what="zkc_create (an sm_100 GPU is required; there is no CPU path)")
        self.h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.zkc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        """stream: a torch.cuda.Stream, a raw cudaStream_t integer or None (legacy default stream)"""
        raw = getattr(stream, "cuda_stream", stream) or 0
        self.lib.zkc_set_stream(self.h, C.c_void_p(raw))

    @property
    def launches(self):
        return int(self.lib.zkc_launch_count(self.h))

    def version(self):
        return self.lib.zkc_version().decode()

    def profile(self, enable=True):
        self.lib.zkc_profile_enable(self.h, int(enable))

    def profile_reset(self):
        self.lib.zkc_profile_reset(self.h)

    def profile_query(self, name):
        ms, n = C.c_double(), C.c_uint64()
        self.lib.zkc_profile_query(self.h, name.encode(), C.byref(ms), C.byref(n))
        return ms.value, n.value

    # ---- primitives ---------------------------------------------------------------------------
    def poseidon2_permute(self, states):
        """states: [n, 12] uint64 (numpy -> numpy, torch cuda int64/uint64 -> same)"""
        n = states.shape[0]
        if _is_torch(states):
            import torch
            out = torch.empty_like(states)
        else:
            states = np.ascontiguousarray(states, dtype=np.uint64)
            out = np.empty_like(states)
        rc = self.lib.zkc_poseidon2_permute(self.h, ptr(states), ptr(out), n, on_device(states))
        if rc:
            raise ZkcError(rc, what="zkc_poseidon2_permute")
        return out

    def field_ops(self, a, b, c):
        """element-wise (a*b, a+b, a-b, a*b+c) over Goldilocks; canonical uint64 numpy inputs"""
        a, b, c = (np.ascontiguousarray(x, dtype=np.uint64) for x in (a, b, c))
        outs = [np.empty_like(a) for _ in range(4)]
        rc = self.lib.zkc_field_ops(self.h, ptr(a), ptr(b), ptr(c), a.size, *(ptr(o) for o in outs))
        if rc:
            raise ZkcError(rc, what="zkc_field_ops")
        return outs

    def commit_encoding(self, inputs):
        """inputs [n_items, len] uint64 -> [n_items, 4]; fsm_input_output/mod.rs:281-326"""
        n, ln = inputs.shape
        if _is_torch(inputs):
            import torch
            out = torch.empty((n, 4), dtype=inputs.dtype, device=inputs.device)
        else:
            inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
            out = np.empty((n, 4), dtype=np.uint64)
        rc = self.lib.zkc_commit_encoding(self.h, ptr(inputs), ln, n, ptr(out), on_device(inputs))
        if rc:
            raise ZkcError(rc, what="zkc_commit_encoding")
        return out

    def accumulate_grand_products(self, lhs_enc, rhs_enc, challenges, acc_in, should_acc=None, want_chain=False):
        """utils.rs:81-137.  lhs_enc/rhs_enc column-major [enc_len, rows]; challenges [2, enc_len+1];
        acc_in = (lhs0, lhs1, rhs0, rhs1).  Returns (acc_out [4, rows], chain or None, acc_final[4])."""
        enc_len, rows = lhs_enc.shape
        dev = on_device(lhs_enc, rhs_enc, should_acc)
        ch = np.ascontiguousarray(challenges, dtype=np.uint64)
        ai = np.ascontiguousarray(acc_in, dtype=np.uint64)
        fin = np.zeros(4, dtype=np.uint64)
        if dev:
            import torch
            acc = torch.empty((4, rows), dtype=lhs_enc.dtype, device=lhs_enc.device)
            chain = torch.empty((4 * enc_len, rows), dtype=lhs_enc.dtype, device=lhs_enc.device) if want_chain else None
        else:
            lhs_enc = np.ascontiguousarray(lhs_enc, dtype=np.uint64)
            rhs_enc = np.ascontiguousarray(rhs_enc, dtype=np.uint64)
            if should_acc is not None:
                should_acc = np.ascontiguousarray(should_acc, dtype=np.uint8)
            acc = np.empty((4, rows), dtype=np.uint64)
            chain = np.empty((4 * enc_len, rows), dtype=np.uint64) if want_chain else None
        rc = self.lib.zkc_accumulate_grand_products(self.h, ptr(lhs_enc), ptr(rhs_enc), ptr(should_acc), enc_len, rows,
                                                    ptr(ch), ptr(ai), ptr(acc), ptr(chain), ptr(fin), dev)
        if rc:
            raise ZkcError(rc, what="zkc_accumulate_grand_products")
        return acc, chain, fin

    def scale_accumulators(self, acc, factors):
        """acc[c, :] *= factors[c] in place (zkc_scale_accumulators): the fix-up of a row range accumulated from the neutral
        element once the product of everything before it is known (sharding.distributed_grand_products)."""
        n_cols, rows = acc.shape
        f = np.ascontiguousarray(factors, dtype=np.uint64)
        rc = self.lib.zkc_scale_accumulators(self.h, ptr(acc), n_cols, rows, ptr(f), int(on_device(acc)))
        if rc:
            raise ZkcError(rc, what="zkc_scale_accumulators")
        return acc

    def memory_queue_simulate(self, records, n_queues=1):
        """push every record into `n_queues` empty memory queues (records split evenly, in order).
        Returns (prev_states [n, 12] uint64, final_states: list/array of QueueState12)."""
        n = len(records)
        assert n % n_queues == 0
        dev = on_device(records)
        final = (abi.QueueState12 * n_queues)()
        if dev:
            import torch
            prev = torch.empty((n, 12), dtype=torch.int64, device=records.device)
            fin_d = torch.empty((n_queues, C.sizeof(abi.QueueState12)), dtype=torch.uint8, device=records.device)
            rc = self.lib.zkc_memory_queue_simulate(self.h, ptr(records), n // n_queues, n_queues, ptr(prev), ptr(fin_d), 1)
            if rc:
                raise ZkcError(rc, what="zkc_memory_queue_simulate")
            host = fin_d.cpu().numpy()  # keep alive across the memmove
            C.memmove(final, host.ctypes.data, C.sizeof(final))
        else:
            prev = np.empty((n, 12), dtype=np.uint64)
            rc = self.lib.zkc_memory_queue_simulate(self.h, ptr(records), n // n_queues, n_queues, ptr(prev),
                                                    C.cast(final, C.c_void_p), 0)
            if rc:
                raise ZkcError(rc, what="zkc_memory_queue_simulate")
        return prev, final

    def decommit_queue_simulate(self, records, n_queues=1):
        """push every DecommitQuery record into `n_queues` empty full-state queues (records split evenly, in order).
        Returns (prev_states [n, 12] uint64, final_states: array of QueueState12)."""
        n = len(records)
        assert n % n_queues == 0
        dev = on_device(records)
        final = (abi.QueueState12 * n_queues)()
        if dev:
            import torch
            prev = torch.empty((n, 12), dtype=torch.int64, device=records.device)
            fin_d = torch.empty((n_queues, C.sizeof(abi.QueueState12)), dtype=torch.uint8, device=records.device)
            rc = self.lib.zkc_decommit_queue_simulate(self.h, ptr(records), n // n_queues, n_queues, ptr(prev), ptr(fin_d), 1)
            if rc:
                raise ZkcError(rc, what="zkc_decommit_queue_simulate")
            host = fin_d.cpu().numpy()
            C.memmove(final, host.ctypes.data, C.sizeof(final))
        else:
            prev = np.empty((n, 12), dtype=np.uint64)
            rc = self.lib.zkc_decommit_queue_simulate(self.h, ptr(records), n // n_queues, n_queues, ptr(prev),
                                                      C.cast(final, C.c_void_p), 0)
            if rc:
                raise ZkcError(rc, what="zkc_decommit_queue_simulate")
        return prev, final

    def log_queue_simulate(self, records, extra_timestamps=None, n_queues=1):
        """push every LogQuery record into `n_queues` empty 4-wide queues; returns (prev_tails [n, 4], final states)"""
        n = len(records)
        assert n % n_queues == 0
        dev = on_device(records, extra_timestamps)
        final = (abi.QueueState4 * n_queues)()
        if dev:
            import torch
            prev = torch.empty((n, 4), dtype=torch.int64, device=records.device)
            fin_d = torch.empty((n_queues, C.sizeof(abi.QueueState4)), dtype=torch.uint8, device=records.device)
            rc = self.lib.zkc_log_queue_simulate(self.h, ptr(records), ptr(extra_timestamps), n // n_queues, n_queues,
                                                 ptr(prev), ptr(fin_d), 1)
            if rc:
                raise ZkcError(rc, what="zkc_log_queue_simulate")
            host = fin_d.cpu().numpy()
            C.memmove(final, host.ctypes.data, C.sizeof(final))
        else:
            prev = np.empty((n, 4), dtype=np.uint64)
            if extra_timestamps is not None:
                extra_timestamps = np.ascontiguousarray(extra_timestamps, dtype=np.uint32)
            rc = self.lib.zkc_log_queue_simulate(self.h, ptr(records), ptr(extra_timestamps), n // n_queues, n_queues,
                                                 ptr(prev), C.cast(final, C.c_void_p), 0)
            if rc:
                raise ZkcError(rc, what="zkc_log_queue_simulate")
        return prev, final
