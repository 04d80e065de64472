#!/usr/bin/env python3
"""Short driver for ncu captures of the main_vm kernels: 64 instances x 2^12 cycles, device-resident inputs.
  ncu --set full --clock-control none --import-source on -k regex:vm_ -o gpurun_out/prof_vm python tools/profile_vm.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from era_zkevm_circuits_b200 import Engine, abi, isa as I, main_vm_entry_point_batch, main_vm_initial_state, main_vm_simulate  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cycles = 1 << 12
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
for _ in range(3):
    coms, out, statuses, rc = main_vm_entry_point_batch(eng, ios, isa.isa, sim.snapshots, sim.witness, cycles, trace_out=trace,
                                                        callstack_witness=cw)
    assert rc == 0
from era_zkevm_circuits_b200 import main_vm_check_trace  # noqa: E402
viol, _ = main_vm_check_trace(eng, isa.isa, trace, cycles, n)
assert viol == 0
torch.cuda.synchronize()
print("ok")
