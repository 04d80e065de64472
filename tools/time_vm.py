#!/usr/bin/env python3
"""Per-kernel timing of one main_vm step (device-resident inputs) for kernel experiments:
  [ZKC_B200_LIB=path/to/variant.so] python tools/time_vm.py [instances] [log2 cycles] [steps] [rows|columns]
rows: the record entry point (transposition to columns inside the step); columns (default): inputs already columns in HBM"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from era_zkevm_circuits_b200 import (Engine, abi, isa as I, main_vm_entry_point_batch, main_vm_entry_point_columns, main_vm_initial_state,  # noqa: E402
                                     main_vm_rows_to_columns, main_vm_simulate)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cycles = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 18)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mode = sys.argv[4] if len(sys.argv) > 4 else "columns"
eng = Engine(0)
isa = I.Isa()
ios, states, codes = [], [], []
progs = [I.pack_code(I.random_program(isa, 4096, seed=0xC2 + k)) for k in range(4)]
for i in range(n):
    io = abi.VmClosedForm(); io.start_flag = 1; io.rollback_queue_tail_for_block[0] = i
    ios.append(io); states.append(main_vm_initial_state(eng, io, isa.isa)); codes.append(progs[i % 4])
sim = main_vm_simulate(eng, isa.isa, states, np.stack(codes), cycles)
assert sim.status.code == 0
for io, t in zip(ios, sim.rollback_tails):
    for k in range(4):
        io.rollback_queue_tail_for_block[k] = int(t[k])
cw = sim.callstack_witness[:, :max(1, int(sim.n_callstack.max()))].contiguous()
trace = torch.empty((n, abi.VM_COLS["NUM_COLS"], cycles), dtype=torch.int64, device="cuda")


cols = main_vm_rows_to_columns(eng, sim.snapshots, sim.witness, cycles) if mode == "columns" else None


def step():
    if cols is not None:
        coms, out, statuses, rc = main_vm_entry_point_columns(eng, ios, isa.isa, cols, cycles, trace_out=trace, callstack_witness=cw)
    else:
        coms, out, statuses, rc = main_vm_entry_point_batch(eng, ios, isa.isa, sim.snapshots, sim.witness, cycles, trace_out=trace, callstack_witness=cw)
    assert rc == 0
    return coms


for _ in range(3):
    c0 = step()
eng.profile_reset(); eng.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(steps):
    step()
e1.record(); torch.cuda.synchronize()
eng.profile(False)
ms = e0.elapsed_time(e1) / steps
torch.cuda.synchronize(); e0.record()  # and without per-kernel profiling: vm_link runs beside the sponge kernels on the side stream
for _ in range(steps):
    step()
e1.record(); torch.cuda.synchronize()
ms_free = e0.elapsed_time(e1) / steps
out = {k: eng.profile_query(k) for k in ("vm_rows_to_columns", "vm_cycles", "vm_link", "vm_sponge", "vm_sponge_far", "vm_sponge_trace", "vm_finalize")}
print(os.environ.get("ZKC_B200_LIB", "default"), mode, f"step {ms_free:.3f} ms = {n * cycles / ms_free / 1e3:.1f} M cycles/s (kernel by kernel: {ms:.3f} ms) |",
      " ".join(f"{k} {v[0] / steps:.3f}" for k, v in out.items()), "| commitment", hex(int(c0[0][0])))
