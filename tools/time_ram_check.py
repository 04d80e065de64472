#!/usr/bin/env python3
"""Timing of zkc_ram_permutation_check_trace (general-purpose gates) on a 2^k-row trace rewritten before every launch (cold HBM):
  [ZKC_B200_LIB=variant.so] python tools/time_ram_check.py [log2 rows]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from era_zkevm_circuits_b200 import (Engine, RamPermutationCircuitInstanceWitness, abi, ram_permutation_check_trace,  # noqa: E402
                                     ram_permutation_entry_point, synthetic)

rn = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
eng = Engine(0)
u, s = synthetic.ram_trace(rn, seed=0xC1, n_cells=1 << 10, n_nondet=7)
d_both = torch.from_numpy(np.concatenate([u, s]).view(np.uint8).reshape(2 * rn, 64)).cuda()
prev, fin = eng.memory_queue_simulate(d_both, n_queues=2)
rio = abi.RamClosedForm(); rio.start_flag = 1
rio.observable_input.unsorted_queue_initial_state = fin[0]
rio.observable_input.sorted_queue_initial_state = fin[1]
rio.observable_input.non_deterministic_bootloader_memory_snapshot_length = 7
rtrace = torch.empty((abi.RAM_COLS["NUM_COLS"], rn), dtype=torch.int64, device="cuda")
rw = RamPermutationCircuitInstanceWitness(rio, d_both[:rn], prev[:rn], d_both[rn:], prev[rn:])
assert ram_permutation_entry_point(eng, rw, rn, trace_out=rtrace).status.code == 0
for _ in range(3):
    assert ram_permutation_check_trace(eng, rio, rtrace, rn, abi.GATES_GENERAL)[0] == 0
eng.profile_reset(); eng.profile(True)
for _ in range(10):
    ram_permutation_entry_point(eng, rw, rn, trace_out=rtrace)  # rewrites the trace: the check reads cold HBM
    assert ram_permutation_check_trace(eng, rio, rtrace, rn, abi.GATES_GENERAL)[0] == 0
eng.profile(False)
ms, k = eng.profile_query("ram_check")
rms, rk = eng.profile_query("ram_rows")
ims, ik = eng.profile_query("ram_inverse")
nb = abi.RAM_COLS["NUM_COLS"] * 8
print(os.environ.get("ZKC_B200_LIB", "default"), f"ram_check {ms / k:.3f} ms per 2^{rn.bit_length() - 1} rows = {rn * nb / (ms / k) / 1e6:.0f} GB/s | "
      f"ram_rows {rms / max(rk, 1):.3f} ms ram_inverse {ims / max(ik, 1):.3f} ms")
