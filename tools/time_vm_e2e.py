#!/usr/bin/env python3
"""End-to-end timing of the main_vm call over HOST buffers, both transport forms:
  python tools/time_vm_e2e.py [log2 cycles] [steps] [segment_cycles]
records: pinned zkc_vm_state / zkc_vm_cycle_witness records in, COMPACT trace out (round 1's e2e)
stream : segmented input stream in, PACKED trace out (zkc_main_vm_entry_point_stream)"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from era_zkevm_circuits_b200 import (Engine, abi, isa as I, main_vm_entry_point_batch, main_vm_initial_state, main_vm_simulate)  # noqa: E402
from era_zkevm_circuits_b200.main_vm import main_vm_entry_point_stream, vm_encode_input_stream, vm_packed_trace_buffers  # noqa: E402

cycles = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
segment = int(sys.argv[3]) if len(sys.argv) > 3 else 0
eng = Engine(0)
isa = I.Isa()
io = abi.VmClosedForm(); io.start_flag = 1
st0 = main_vm_initial_state(eng, io, isa.isa)
sim = main_vm_simulate(eng, isa.isa, [st0], np.stack([I.pack_code(I.random_program(isa, 4096, seed=0xC2))]), cycles)
assert sim.status.code == 0
for k in range(4):
    io.rollback_queue_tail_for_block[k] = int(sim.rollback_tails[0][k])
n_cw = max(1, int(sim.n_callstack.max()))
cw = sim.callstack_witness[:, :n_cw].contiguous().cpu().numpy()
snaps, wit = sim.snapshots.cpu().numpy(), sim.witness.cpu().numpy()
del sim
torch.cuda.empty_cache()


def pinned(shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = eng.lib.zkc_host_alloc(max(n, 1))
    assert p
    return np.frombuffer((C.c_uint8 * n).from_address(p), dtype=dtype).reshape(shape)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


# ---- records in, COMPACT out
hs, hw, hc = pinned(snaps.shape, np.uint8), pinned(wit.shape, np.uint8), pinned(cw.shape, np.uint8)
hs[:] = snaps; hw[:] = wit; hc[:] = cw
htrace = pinned((1, abi.VM_COMPACT_COLS, cycles), np.uint64)
hrec = pinned((int(cycles * 1.25), 104), np.uint8)
res = {}


def step_records():
    coms, out, sts, rc = main_vm_entry_point_batch(eng, [io], isa.isa, hs, hw, cycles, trace_out=htrace, callstack_witness=hc, sponge_records_out=hrec)
    assert rc == 0
    res["records"] = coms[0].tolist()


ms = timed(step_records)
print(f"records: {ms:.2f} ms/step = {cycles / ms / 1e3:.1f} M cycles/s; h2d {(hs.nbytes + hw.nbytes) / 1e6:.0f} MB d2h {htrace.nbytes / 1e6:.0f}+ MB", flush=True)

# ---- stream in, PACKED out
t0 = time.perf_counter()
stream = vm_encode_input_stream(eng.lib, snaps[0], wit[0], cycles, segment)
t_enc = time.perf_counter() - t0
out = vm_packed_trace_buffers(eng, 1, cycles, alloc=pinned)


def step_stream():
    coms, ios, sts, rc = main_vm_entry_point_stream(eng, [io], isa.isa, [stream], cycles, callstack_witness=hc, out=out)
    assert rc == 0, (rc, sts[0].code, hex(sts[0].failed_checks), sts[0].first_bad_row)
    res["stream"] = coms[0].tolist()


ms = timed(step_stream)
print(f"stream : {ms:.2f} ms/step = {cycles / ms / 1e3:.1f} M cycles/s; h2d {stream.bytes / 1e6:.0f} MB ({stream.bytes / cycles:.0f} B/cycle) "
      f"d2h {out.nbytes_used / 1e6:.0f} MB ({out.nbytes_used / cycles:.0f} B/cycle); host encode {t_enc:.2f} s; "
      f"aux records {out.n_aux_records} sponge records {out.n_sponge_records}", flush=True)
assert res["records"] == res["stream"]
print("commitments equal")
