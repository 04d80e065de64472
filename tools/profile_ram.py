#!/usr/bin/env python3
"""Short driver for ncu captures: one ram_permutation instance of 2^20 rows (witness generation + both
constraint-evaluation variants), device-resident inputs.  Usage (on the GPU box):
  ncu --set full --clock-control none --import-source on -k regex:ram_ -o gpurun_out/prof_ram python tools/profile_ram.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from era_zkevm_circuits_b200 import (Engine, RamPermutationCircuitInstanceWitness, abi, ram_permutation_check_trace,  # noqa: E402
                                     ram_permutation_entry_point, synthetic)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
eng = Engine(0)
u, s = synthetic.ram_trace(n, seed=0xC1, n_nondet=7)
# 64 sub-queues hashed in parallel would not be one instance; keep setup short instead: chain on 2^20 rows takes ~20 s
both = torch.from_numpy(np.concatenate([u, s]).view(np.uint8).reshape(2 * n, 64)).cuda()
prev, fin = eng.memory_queue_simulate(both, n_queues=2)
io = abi.RamClosedForm()
io.start_flag = 1
io.observable_input.unsorted_queue_initial_state = fin[0]
io.observable_input.sorted_queue_initial_state = fin[1]
io.observable_input.non_deterministic_bootloader_memory_snapshot_length = 7
trace = torch.empty((abi.RAM_COLS["NUM_COLS"], n), dtype=torch.int64, device="cuda")
w = RamPermutationCircuitInstanceWitness(io, both[:n], prev[:n], both[n:], prev[n:])
for _ in range(2):
    r = ram_permutation_entry_point(eng, w, n, trace_out=trace)
    v1, _ = ram_permutation_check_trace(eng, io, trace, n, abi.GATES_GENERAL)
    v2, _ = ram_permutation_check_trace(eng, io, trace, n, 0)
    assert r.status.code == 0 and v1 == 0 and v2 == 0
torch.cuda.synchronize()
print("ok")
