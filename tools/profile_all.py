#!/usr/bin/env python3
"""One or two launches of every kernel family, for ncu captures (round 2 evidence):
  ncu --set full --clock-control none --import-source on -k regex:'<names>' -o gpurun_out/prof python tools/profile_all.py [log2 cycles]
main_vm (columns, 2^k cycles) + its constraint evaluator + gadget cells; ram_permutation + evaluator; log_sorter + evaluator;
storage_validity + evaluator + the generic allocation-check evaluator; keccak256 / sha256 round functions; linear_hasher.
Queue-state hints come from the engine's own un-hinted first run (verified chains), as in tools/bench_configs.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from era_zkevm_circuits_b200 import (Engine, EventsDeduplicatorInstanceWitness, LinearHasherCircuitInstanceWitness,  # noqa: E402
                                     RamPermutationCircuitInstanceWitness, abi, isa as I, linear_hasher_entry_point, log_sorter_check_trace,
                                     main_vm_check_trace, main_vm_entry_point_columns, main_vm_gadget_cells, main_vm_initial_state,
                                     main_vm_rows_to_columns, main_vm_simulate, ram_permutation_check_trace, ram_permutation_entry_point,
                                     sort_and_deduplicate_events_entry_point, synthetic)

k = int(sys.argv[1]) if len(sys.argv) > 1 else 18
eng = Engine(0)
eng.set_stream(torch.cuda.current_stream())
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)).cuda()
i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()

# ---- main_vm -------------------------------------------------------------------------------------------------------------
cycles = 1 << k
isa = I.Isa()
io = abi.VmClosedForm(); io.start_flag = 1
st0 = main_vm_initial_state(eng, io, isa.isa)
sim = main_vm_simulate(eng, isa.isa, [st0], np.stack([I.pack_code(I.random_program(isa, 4096, seed=0xC2))]), cycles)
assert sim.status.code == 0
for j in range(4):
    io.rollback_queue_tail_for_block[j] = int(sim.rollback_tails[0][j])
cw = sim.callstack_witness[:, :max(1, int(sim.n_callstack.max()))].contiguous()
cols = main_vm_rows_to_columns(eng, sim.snapshots, sim.witness, cycles)
trace = torch.empty((1, abi.VM_COLS["NUM_COLS"], cycles), dtype=torch.int64, device="cuda")
for _ in range(2):
    coms, out, sts, rc = main_vm_entry_point_columns(eng, [io], isa.isa, cols, cycles, trace_out=trace, callstack_witness=cw)
    assert rc == 0
viol, _ = main_vm_check_trace(eng, isa.isa, trace, cycles, 1)
assert viol == 0
g = main_vm_gadget_cells(eng, trace, cycles, 1)
del g, trace, cols, sim
torch.cuda.empty_cache()

# ---- ram_permutation ---------------------------------------------------------------------------------------------------------
rn = 1 << k
u, s = synthetic.ram_trace(rn, seed=0xC1, n_cells=1 << 10, n_nondet=7)
both = dev(np.concatenate([u, s]))
prev, fin = eng.memory_queue_simulate(both, n_queues=2)
rio = abi.RamClosedForm(); rio.start_flag = 1
rio.observable_input.unsorted_queue_initial_state = fin[0]
rio.observable_input.sorted_queue_initial_state = fin[1]
rio.observable_input.non_deterministic_bootloader_memory_snapshot_length = 7
rtrace = torch.empty((abi.RAM_COLS["NUM_COLS"], rn), dtype=torch.int64, device="cuda")
r = ram_permutation_entry_point(eng, RamPermutationCircuitInstanceWitness(rio, both[:rn], prev[:rn], both[rn:], prev[rn:]), rn, trace_out=rtrace)
assert r.status.code == 0
assert ram_permutation_check_trace(eng, rio, rtrace, rn, abi.GATES_GENERAL)[0] == 0
del rtrace, both, prev

# ---- log_sorter ----------------------------------------------------------------------------------------------------------------
en = 1 << (k - 2)
eu, es = synthetic.events_trace(en, seed=0xC4, rollback_pct=10)
d_eu, d_es = dev(eu), dev(es)
eup, eufin = eng.log_queue_simulate(d_eu)
esp, esfin = eng.log_queue_simulate(d_es)
eio = abi.EventsClosedForm(); eio.start_flag = 1
eio.initial_log_queue_state = eufin[0]; eio.intermediate_sorted_queue_state = esfin[0]
etrace = torch.empty((abi.EV_COLS["NUM_COLS"], en), dtype=torch.int64, device="cuda")
w = EventsDeduplicatorInstanceWitness(eio, d_eu, eup, d_es, esp)
got = sort_and_deduplicate_events_entry_point(eng, w, en, trace_out=etrace)  # un-hinted: the chain kernel runs once
assert got.status.code == 0
K = abi.EV_COLS
tails = etrace[K["RESULT_TAIL"]:K["RESULT_TAIL"] + 4].t()[etrace[K["ADD_TO_QUEUE"]] != 0].contiguous()
w.result_queue_tails = tails
got = sort_and_deduplicate_events_entry_point(eng, w, en, trace_out=etrace)
assert got.status.code == 0
assert log_sorter_check_trace(eng, eio, etrace, en, abi.GATES_GENERAL)[0] == 0
assert log_sorter_check_trace(eng, eio, etrace, en, 0)[0] == 0
del etrace

# ---- storage_validity + evaluator, the generic allocation-check evaluator --------------------------------------------------------
from era_zkevm_circuits_b200 import (StorageDeduplicatorInstanceWitness, check_trace_columns, sort_and_deduplicate_storage_access_entry_point,  # noqa: E402
                                     storage_validity_check_trace)
su, ss, sts = synthetic.storage_trace(en, seed=0xC4, n_cells=1 << 10)
d_ts = torch.from_numpy(sts.astype(np.uint32).view(np.int32)).cuda()
spu, sfu = eng.log_queue_simulate(dev(su))
sps, sfs = eng.log_queue_simulate(dev(ss), d_ts)
sio = abi.StorageClosedForm(); sio.start_flag = 1
sio.unsorted_log_queue_state = sfu[0]; sio.intermediate_sorted_queue_state = sfs[0]
sw = StorageDeduplicatorInstanceWitness(sio, dev(su), spu, dev(ss), d_ts, sps, None)
strace = torch.empty((abi.ST_COLS["NUM_COLS"], en), dtype=torch.int64, device="cuda")
assert sort_and_deduplicate_storage_access_entry_point(eng, sw, en, trace_out=strace).status.code == 0
KS = abi.ST_COLS
sw.result_queue_tails = strace[KS["RESULT_TAIL"]:KS["RESULT_TAIL"] + 4].t()[strace[KS["SHOULD_PUSH"]] != 0].contiguous()
assert sort_and_deduplicate_storage_access_entry_point(eng, sw, en, trace_out=strace).status.code == 0
assert storage_validity_check_trace(eng, sio, strace, en, abi.GATES_GENERAL)[0] == 0
assert storage_validity_check_trace(eng, sio, strace, en, 0)[0] == 0
assert check_trace_columns(eng, "storage_validity", strace)[0] == 0
del strace

# ---- linear_hasher ---------------------------------------------------------------------------------------------------------------
ln = 1 << (k - 4)
recs = synthetic.vm_log_queue_trace(ln, seed=5)
recs["tx_number_in_block"] &= 0xFFFF
d_recs = dev(recs)
lp, lfin = eng.log_queue_simulate(d_recs)
lio = abi.LinearHasherClosedForm(); lio.start_flag = 1; lio.queue_state = lfin[0]
ltrace = torch.empty((abi.LH_COLS["NUM_COLS"], ln), dtype=torch.int64, device="cuda")
got = linear_hasher_entry_point(eng, LinearHasherCircuitInstanceWitness(lio, d_recs, lp), ln, trace_out=ltrace)  # chain rebuilt on the device
assert got.status.code == 0
KL = abi.LH_COLS
so = ltrace[KL["STATE_OUT"]:KL["STATE_OUT"] + 50].t().contiguous()
states = (so[:, 0::2] | (so[:, 1::2] << 32)).contiguous()
got = linear_hasher_entry_point(eng, LinearHasherCircuitInstanceWitness(lio, d_recs, lp, states), ln, trace_out=ltrace)
assert got.status.code == 0
torch.cuda.synchronize()
print("ok")
