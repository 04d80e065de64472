#!/usr/bin/env python3
"""Timing of zkc_main_vm_check_trace on a 2^20-row trace (device-resident): [ZKC_B200_LIB=variant.so] python tools/time_check.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from era_zkevm_circuits_b200 import Engine, abi, isa as I, main_vm_check_trace, main_vm_entry_point_batch, main_vm_initial_state, main_vm_simulate  # noqa: E402

n, cycles = 16, 1 << 16
eng = Engine(0)
isa = I.Isa()
ios, states = [], []
code = I.pack_code(I.random_program(isa, 4096, seed=0xC2))
for i in range(n):
    io = abi.VmClosedForm(); io.start_flag = 1; io.rollback_queue_tail_for_block[0] = i
    ios.append(io); states.append(main_vm_initial_state(eng, io, isa.isa))
sim = main_vm_simulate(eng, isa.isa, states, np.stack([code] * n), cycles)
for io, t in zip(ios, sim.rollback_tails):
    for k in range(4):
        io.rollback_queue_tail_for_block[k] = int(t[k])
cw = sim.callstack_witness[:, :max(1, int(sim.n_callstack.max()))].contiguous()
trace = torch.empty((n, abi.VM_COLS["NUM_COLS"], cycles), dtype=torch.int64, device="cuda")
coms, out, statuses, rc = main_vm_entry_point_batch(eng, ios, isa.isa, sim.snapshots, sim.witness, cycles, trace_out=trace, callstack_witness=cw)
assert rc == 0
flush = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
for _ in range(3):
    viol, _ = main_vm_check_trace(eng, isa.isa, trace, cycles, n)
assert viol == 0
eng.profile_reset(); eng.profile(True)
for _ in range(10):
    flush.fill_(1)  # 256 MB > L2
    main_vm_check_trace(eng, isa.isa, trace, cycles, n)
eng.profile(False)
ms, k = eng.profile_query("vm_check")
print(os.environ.get("ZKC_B200_LIB", "default"), f"vm_check {ms / k:.3f} ms per 2^20 rows = {n * cycles * 2208 / (ms / k) / 1e6:.0f} GB/s")
