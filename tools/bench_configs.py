#!/usr/bin/env python3
"""Secondary configurations of BASELINE.json (the bench.py line is configs[1]); one JSON line per measurement.

  python tools/bench_configs.py c1            ram_permutation 2^16 rows: GPU vs the CPU oracle, bit-exact, both timed
  python tools/bench_configs.py c3            keccak256 + sha256 round-function circuits, 2^18 cycles each
  python tools/bench_configs.py c4 [log2rows] storage_validity + log_sorter entry points (default 2^20 rows: building the
                                              input queues is a sequential hash chain of 3 permutations per record)
  torchrun --nproc-per-node N tools/bench_configs.py gp [log2rows]
                                              ONE grand product of 2^22 rows (ENC = 20) cut over N ranks by row ranges:
                                              local accumulation, one all-gather of 4 x u64 per rank, re-seeded pass
Timing: CUDA events on the engine's stream, warm-up first, device-resident inputs."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from era_zkevm_circuits_b200 import (Engine, EventsDeduplicatorInstanceWitness, Keccak256RoundFunctionCircuitInstanceWitness,  # noqa: E402
                                     RamPermutationCircuitInstanceWitness, Sha256RoundFunctionCircuitInstanceWitness,
                                     StorageDeduplicatorInstanceWitness, abi, keccak256_round_function_entry_point,
                                     ram_permutation_entry_point, sha256_round_function_entry_point, sharding,
                                     sort_and_deduplicate_events_entry_point, sort_and_deduplicate_storage_access_entry_point, synthetic,
                                     CodeDecommittmentsDeduplicatorInstanceWitness, sort_and_deduplicate_code_decommittments_entry_point,
                                     LogDemuxerCircuitInstanceWitness, demultiplex_storage_logs_enty_point,
                                     CodeDecommitterCircuitInstanceWitness, unpack_code_into_memory_entry_point)


def timed(fn, steps=5, warmup=2):
    for _ in range(warmup):
        out = fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)).cuda()


def once(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


def pushes_from_trace(trace, groups, flag_cols, state_cols, width):
    """queue states after every executed push, in push order: row-major over (row, push slot).  trace [cols, rows] int64 on
    the GPU; flag_cols / state_cols: per push slot, the column of its execute flag and the first column of its state"""
    rows = trace.shape[1]
    flags = torch.stack([trace[c] for c in flag_cols], dim=1) != 0                                   # [rows, slots]
    states = torch.stack([trace[c:c + width].t() for c in state_cols], dim=1)                        # [rows, slots, width]
    return states[flags].contiguous()


def c1(eng):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    import orc as O  # the oracle: checker + CPU baseline of this configuration
    lib = O.load()
    n = 1 << 16
    u, s = synthetic.ram_trace(n, seed=0xC1, n_cells=1 << 10, n_nondet=7)
    io, up, sp = H.ram_instance(lib, u, s, 7)
    t0 = time.perf_counter()
    rc, io_ref, trace_ref, com_ref, st_ref = O.ram_entry_point(lib, io, u, s, n)
    cpu_s = time.perf_counter() - t0
    w = RamPermutationCircuitInstanceWitness(io, dev(u), torch.from_numpy(up.view(np.int64)).cuda(), dev(s), torch.from_numpy(sp.view(np.int64)).cuda())
    trace = torch.empty((abi.RAM_COLS["NUM_COLS"], n), dtype=torch.int64, device="cuda")
    ms, got = timed(lambda: ram_permutation_entry_point(eng, w, n, trace_out=trace))
    same = bool(np.array_equal(trace.cpu().numpy().view(np.uint64), trace_ref)) and got.commitment.tolist() == com_ref.tolist()
    print(json.dumps({"config": "C1 ram_permutation, 2^16 rows", "gpu_ms": ms, "gpu_rows_per_s": n / ms * 1e3, "cpu_oracle_s_1_thread": cpu_s,
                      "cpu_rows_per_s": n / cpu_s, "bit_exact_vs_oracle": same, "status": got.status.code}))


def c3(eng, log2cycles=18):
    cycles = 1 << log2cycles
    reqs, reads, msgs = synthetic.keccak_calls(cycles // 4, seed=0xC3)  # ~4.3 cycles per call at lengths < 1024
    prev, fin = eng.log_queue_simulate(dev(reqs))
    io = abi.KeccakClosedForm(); io.start_flag = 1; io.initial_log_queue_state = fin[0]
    w = Keccak256RoundFunctionCircuitInstanceWitness(io, dev(reqs), prev, torch.from_numpy(reads.view(np.int32)).cuda(), None)
    trace = torch.empty((abi.KC_COLS["NUM_COLS"], cycles), dtype=torch.int64, device="cuda")
    K = abi.KC_COLS
    ms0, got = once(lambda: keccak256_round_function_entry_point(eng, w, cycles, trace_out=trace, raise_on_unsatisfied=False))
    # the memory queue the circuit writes is a hash chain: without the host's intermediate states it is rebuilt sequentially;
    # the out-of-circuit run has them (here: taken from the first run's trace) and every push is then verified in parallel
    fc = [K["QUERY"] + q * K["QUERY_STRIDE"] + 3 for q in range(6)] + [K["WRITE_RESULT"]]
    sc = [K["QUERY"] + q * K["QUERY_STRIDE"] + 12 for q in range(6)] + [K["WRITE_TAIL"]]
    w.memory_queue_states = pushes_from_trace(trace, None, fc, sc, 12)
    ref = trace.clone()
    ms, got = timed(lambda: keccak256_round_function_entry_point(eng, w, cycles, trace_out=trace, raise_on_unsatisfied=False))
    print(json.dumps({"config": f"C3 keccak256_round_function, 2^{log2cycles} cycles", "gpu_ms_with_queue_states": ms, "cycles_per_s": cycles / ms * 1e3,
                      "gpu_ms_without_queue_states_sequential_chain": ms0, "calls": len(reqs), "memory_pushes": len(w.memory_queue_states),
                      "status": got.status.code, "completed": int(got.closed_form_input.completion_flag), "trace_columns": K["NUM_COLS"],
                      "same_trace_both_ways": bool(torch.equal(ref, trace))}))
    del ref
    reqs, reads, msgs = synthetic.sha256_calls(cycles // 9, seed=0xC3)  # ~8.5 rounds per call
    prev, fin = eng.log_queue_simulate(dev(reqs))
    io = abi.Sha256ClosedForm(); io.start_flag = 1; io.initial_log_queue_state = fin[0]
    w = Sha256RoundFunctionCircuitInstanceWitness(io, dev(reqs), prev, torch.from_numpy(reads.view(np.int32)).cuda(), None)
    trace = torch.empty((abi.SH_COLS["NUM_COLS"], cycles), dtype=torch.int64, device="cuda")
    K = abi.SH_COLS
    ms0, got = once(lambda: sha256_round_function_entry_point(eng, w, cycles, trace_out=trace, raise_on_unsatisfied=False))
    sr = K["SHOULD_READ"]
    w.memory_queue_states = pushes_from_trace(trace, None, [sr, sr, K["WRITE_RESULT"]],
                                              [K["QUERY"] + 8, K["QUERY"] + K["QUERY_STRIDE"] + 8, K["WRITE_TAIL"]], 12)
    ref = trace.clone()
    ms, got = timed(lambda: sha256_round_function_entry_point(eng, w, cycles, trace_out=trace, raise_on_unsatisfied=False))
    print(json.dumps({"config": f"C3 sha256_round_function, 2^{log2cycles} cycles", "gpu_ms_with_queue_states": ms, "cycles_per_s": cycles / ms * 1e3,
                      "gpu_ms_without_queue_states_sequential_chain": ms0, "calls": len(reqs), "memory_pushes": len(w.memory_queue_states),
                      "status": got.status.code, "completed": int(got.closed_form_input.completion_flag), "trace_columns": K["NUM_COLS"],
                      "same_trace_both_ways": bool(torch.equal(ref, trace))}))


def c4(eng, log2rows):
    n = 1 << log2rows
    u, s = synthetic.events_trace(n, seed=0xC4, rollback_pct=10)
    t0 = time.perf_counter()
    prev, fin = eng.log_queue_simulate(dev(np.concatenate([u, s])), n_queues=2)
    setup = time.perf_counter() - t0
    io = abi.EventsClosedForm(); io.start_flag = 1
    io.initial_log_queue_state = fin[0]; io.intermediate_sorted_queue_state = fin[1]
    w = EventsDeduplicatorInstanceWitness(io, dev(u), prev[:n], dev(s), prev[n:], None)
    trace = torch.empty((abi.EV_COLS["NUM_COLS"], n), dtype=torch.int64, device="cuda")
    K = abi.EV_COLS
    ms0, got = once(lambda: sort_and_deduplicate_events_entry_point(eng, w, n, trace_out=trace, raise_on_unsatisfied=False))
    # result queue tails after each executed push (+ the finalisation push): the hint that makes the pushes row-parallel
    tails = pushes_from_trace(trace, None, [K["ADD_TO_QUEUE"]], [K["RESULT_TAIL"]], 4)
    fin_tail = torch.tensor(np.array(list(got.closed_form_input.final_queue_state.tail), dtype=np.uint64).view(np.int64), device="cuda").reshape(1, 4)
    if int(got.closed_form_input.final_queue_state.length) == len(tails) + 1:
        tails = torch.cat([tails, fin_tail])
    w.result_queue_tails = tails.contiguous()
    ms, got = timed(lambda: sort_and_deduplicate_events_entry_point(eng, w, n, trace_out=trace, raise_on_unsatisfied=False), steps=3, warmup=1)
    print(json.dumps({"config": f"C4 log_sorter, 2^{log2rows} rows", "gpu_ms_with_result_tails": ms, "rows_per_s": n / ms * 1e3,
                      "gpu_ms_without_result_tails_sequential_chain": ms0, "result_pushes": len(tails), "status": got.status.code,
                      "failed_checks": got.status.failed_checks, "completed": int(got.closed_form_input.completion_flag), "queue_setup_s": setup,
                      "trace_GB": trace.numel() * 8 / 1e9}))
    del trace, w
    u, s, ts = synthetic.storage_trace(n, seed=0xC4, n_cells=1 << 16)
    d_ts = torch.from_numpy(np.concatenate([np.zeros(n, dtype=np.uint32), ts.astype(np.uint32)]).view(np.int32)).cuda()
    # the unsorted queue holds plain records, the sorted one timestamped records: two chains
    pu, fu = eng.log_queue_simulate(dev(u))
    psd, fs = eng.log_queue_simulate(dev(s), d_ts[n:].contiguous())
    io = abi.StorageClosedForm(); io.start_flag = 1
    io.unsorted_log_queue_state = fu[0]; io.intermediate_sorted_queue_state = fs[0]
    w = StorageDeduplicatorInstanceWitness(io, dev(u), pu, dev(s), d_ts[n:].contiguous(), psd, None)
    trace = torch.empty((abi.ST_COLS["NUM_COLS"], n), dtype=torch.int64, device="cuda")
    K = abi.ST_COLS
    ms0, got = once(lambda: sort_and_deduplicate_storage_access_entry_point(eng, w, n, trace_out=trace, raise_on_unsatisfied=False))
    tails = pushes_from_trace(trace, None, [K["SHOULD_PUSH"]], [K["RESULT_TAIL"]], 4)
    fin_tail = torch.tensor(np.array(list(got.closed_form_input.final_sorted_queue_state.tail), dtype=np.uint64).view(np.int64), device="cuda").reshape(1, 4)
    if int(got.closed_form_input.final_sorted_queue_state.length) == len(tails) + 1:
        tails = torch.cat([tails, fin_tail])
    w.result_queue_tails = tails.contiguous()
    ms, got = timed(lambda: sort_and_deduplicate_storage_access_entry_point(eng, w, n, trace_out=trace, raise_on_unsatisfied=False), steps=3, warmup=1)
    print(json.dumps({"config": f"C4 storage_validity_by_grand_product, 2^{log2rows} rows", "gpu_ms_with_result_tails": ms, "rows_per_s": n / ms * 1e3,
                      "gpu_ms_without_result_tails_sequential_chain": ms0, "result_pushes": len(tails),
                      "status": got.status.code, "failed_checks": got.status.failed_checks, "completed": int(got.closed_form_input.completion_flag),
                      "trace_GB": trace.numel() * 8 / 1e9}))


def dq(eng, log2rows):
    """sort_decommittment_requests, 2^log2rows requests over 2^14 code hashes, device resident"""
    n = 1 << log2rows
    u, s = synthetic.decommit_requests_trace(n, seed=0xC4, n_hashes=1 << 14)
    prev, fin = eng.decommit_queue_simulate(dev(np.concatenate([u, s])), n_queues=2)
    io = abi.DecommitSorterClosedForm(); io.start_flag = 1
    io.initial_queue_state = fin[0]; io.sorted_queue_initial_state = fin[1]
    w = CodeDecommittmentsDeduplicatorInstanceWitness(io, dev(u), prev[:n], dev(s), prev[n:], None)
    trace = torch.empty((abi.DQ_COLS["NUM_COLS"], n), dtype=torch.int64, device="cuda")
    K = abi.DQ_COLS
    run = lambda: sort_and_deduplicate_code_decommittments_entry_point(eng, w, n, trace_out=trace, raise_on_unsatisfied=False)
    ms0, got = once(run)
    states = pushes_from_trace(trace, None, [K["ADD_TO_QUEUE"]], [K["RESULT_TAIL"]], 12)
    fin_tail = torch.tensor(np.array(list(got.closed_form_input.final_queue_state.tail), dtype=np.uint64).view(np.int64), device="cuda").reshape(1, 12)
    if int(got.closed_form_input.final_queue_state.length) == len(states) + 1:
        states = torch.cat([states, fin_tail])
    w.result_queue_states = states.contiguous()
    ms, got = timed(run, steps=5, warmup=2)
    eng.profile(True)
    run()
    prof = {k: eng.profile_query(k)[0] for k in ("dq_rows", "dq_push", "dq_finalize", "dq_prologue")}
    eng.profile(False)
    # algorithmic bytes per row: 2 records (48 B) + 2 previous states (96 B) read, 173 trace columns written
    alg = n * (2 * 48 + 2 * 96 + 8 * abi.DQ_COLS["NUM_COLS"])
    print(json.dumps({"config": f"sort_decommittment_requests, 2^{log2rows} rows, 2^14 hashes", "gpu_ms_with_result_states": ms,
                      "rows_per_s": n / ms * 1e3, "algorithmic_GB_per_s": alg / ms / 1e6,
                      "gpu_ms_without_result_states_sequential_chain": ms0, "result_pushes": len(states), "status": got.status.code,
                      "failed_checks": got.status.failed_checks, "completed": int(got.closed_form_input.completion_flag),
                      "kernel_ms": prof, "trace_GB": trace.numel() * 8 / 1e9}))


def dmx(eng, log2rows):
    """demux_log_queue, 2^log2rows VM log records, device resident"""
    n = 1 << log2rows
    recs = synthetic.vm_log_queue_trace(n, seed=0xC4)
    t0 = time.perf_counter()
    prev, fin = eng.log_queue_simulate(dev(recs))
    setup = time.perf_counter() - t0
    io = abi.DemuxClosedForm(); io.start_flag = 1
    io.initial_log_queue_state = fin[0]
    w = LogDemuxerCircuitInstanceWitness(io, dev(recs), prev, None, None)
    trace = torch.empty((abi.DMX_COLS["NUM_COLS"], n), dtype=torch.int64, device="cuda")
    K = abi.DMX_COLS
    run = lambda: demultiplex_storage_logs_enty_point(eng, w, n, trace_out=trace, raise_on_unsatisfied=False)
    ms0, got = once(run)
    # per queue: the tail after each of its pushes, from the trace of the un-hinted run
    tails, counts = [], []
    for q in range(6):
        rows = torch.nonzero(trace[K["BITMASK"] + q]).flatten()
        tails.append(trace[K["QUEUE_TAILS"] + 4 * q:K["QUEUE_TAILS"] + 4 * q + 4].t()[rows])
        counts.append(len(rows))
    w.output_queue_tails = torch.cat(tails).contiguous(); w.output_queue_counts = counts
    ms, got = timed(run, steps=5, warmup=2)
    eng.profile(True)
    run()
    prof = {k: eng.profile_query(k)[0] for k in ("dmx_rows", "dmx_push", "dmx_finalize", "dmx_prologue")}
    eng.profile(False)
    print(json.dumps({"config": f"demux_log_queue, 2^{log2rows} rows", "gpu_ms_with_output_tails": ms, "rows_per_s": n / ms * 1e3,
                      "gpu_ms_without_output_tails_sequential_chains": ms0, "pushes_per_queue": counts, "status": got.status.code,
                      "failed_checks": got.status.failed_checks, "completed": int(got.closed_form_input.completion_flag),
                      "kernel_ms": prof, "queue_setup_s": setup, "trace_GB": trace.numel() * 8 / 1e9}))


def cu(eng):
    """code_unpacker_sha256, ~2^17 cycles, device resident: many short bytecodes vs few long ones (one thread per request)"""
    for n, max_words in ((4096, 127), (512, 1023)):
        reqs, words = synthetic.code_decommit_requests(n, seed=0xC4, max_words=max_words)
        cycles = int((((reqs["code_hash"][:, 7] & 0xFFFF).astype(np.int64) + 1) // 2).sum())
        d_reqs = dev(reqs)
        prev, fin = eng.decommit_queue_simulate(d_reqs)
        io = abi.CodeUnpackerClosedForm(); io.start_flag = 1
        io.sorted_requests_queue_initial_state = fin[0]
        w = CodeDecommitterCircuitInstanceWitness(io, d_reqs, prev, torch.from_numpy(words.view(np.int32)).cuda(), None)
        trace = torch.empty((abi.CU_COLS["NUM_COLS"], cycles), dtype=torch.int64, device="cuda")
        K = abi.CU_COLS
        run = lambda: unpack_code_into_memory_entry_point(eng, w, cycles, trace_out=trace, raise_on_unsatisfied=False)
        ms0, got = once(run)
        states = pushes_from_trace(trace, None, [K["DECOMMIT"], K["PROCESS_SECOND_WORD"]], [K["MEM_TAIL0"], K["MEM_TAIL1"]], 12)
        w.memory_queue_states = states.contiguous()
        ms, got = timed(run, steps=5, warmup=2)
        eng.profile(True)
        run()
        prof = {k: eng.profile_query(k)[0] for k in ("cu_chain", "cu_rows", "cu_memq", "cu_tail", "cu_plan", "cu_prologue", "cu_finalize")}
        eng.profile(False)
        print(json.dumps({"config": f"code_unpacker_sha256, {n} requests, {len(words)} words, {cycles} cycles", "gpu_ms_with_memory_states": ms,
                          "cycles_per_s": cycles / ms * 1e3, "MB_of_code_per_s": len(words) * 32 / ms / 1e3,
                          "gpu_ms_without_memory_states_sequential_chain": ms0, "memory_pushes": len(states), "status": got.status.code,
                          "failed_checks": got.status.failed_checks, "completed": int(got.closed_form_input.completion_flag), "kernel_ms": prof}))
        del trace, w


def gp(eng, log2rows):
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    n, enc = 1 << log2rows, 20
    lo, hi = sharding.row_range(n, rank, world)
    g = torch.Generator(device="cuda"); g.manual_seed(0xC4 + rank)
    P = 0xFFFFFFFF00000001
    lhs = (torch.randint(0, 1 << 62, (enc, hi - lo), dtype=torch.int64, device="cuda", generator=g))
    rhs = (torch.randint(0, 1 << 62, (enc, hi - lo), dtype=torch.int64, device="cuda", generator=g))
    ch = np.random.default_rng(7).integers(0, P, size=(2, enc + 1), dtype=np.uint64)

    def local(acc_in):
        acc, _, fin = eng.accumulate_grand_products(lhs, rhs, ch, acc_in)
        return acc, fin

    def step():
        return sharding.distributed_grand_products(local, rank, world, device="cuda", scale_fn=eng.scale_accumulators)

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps):
        acc, fin, grand = step()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": f"C4 grand-product scan, 2^{log2rows} rows x ENC 20, {world} GPU(s), rows cut by range", "ms": float(ms.item()),
                          "rows_per_s": n / float(ms.item()) * 1e3, "algorithmic_GBps": n * (352 + (64 if world > 1 else 0)) / float(ms.item()) / 1e6,
                          "passes": "one accumulation pass (352 B / row) + a 4-column fix-up multiply (64 B / row) on the ranks behind the first",
                          "collective": "one all-gather of 4 x u64 per rank", "grand_totals": [hex(int(x)) for x in grand]}))
    if world > 1:
        dist.destroy_process_group()


def c4rows(eng, log2rows):
    """ONE storage_validity instance of 2^log2rows rows cut by row range over the ranks (sharding.storage_validity_row_sharded):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_configs.py c4rows 20
    Every rank builds the same synthetic instance and its hints (queue states: two device chains, ~20 s each at 2^20 rows; the
    result-queue tails and the push offsets come from one whole-instance run, which is also the 1-GPU time the cut is compared
    with); the timed step is phase 1 (cell replay + the rank's rows) + the all-gather + phase 2 (fix-up, closed form, commitment)."""
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    n = 1 << log2rows
    u, s, ts = synthetic.storage_trace(n, seed=0xC4, n_cells=1 << max(4, log2rows - 4))
    d_ts2 = torch.from_numpy(np.concatenate([np.zeros(n, dtype=np.uint32), ts.astype(np.uint32)]).view(np.int32)).cuda()
    d_ts = d_ts2[n:].contiguous()
    t0 = time.perf_counter()
    prev, fin = eng.log_queue_simulate(dev(np.concatenate([u, s])), d_ts2, n_queues=2)  # both chains side by side (timestamp 0 = plain record)
    pu, psd = prev[:n].contiguous(), prev[n:].contiguous()
    setup = time.perf_counter() - t0
    io = abi.StorageClosedForm(); io.start_flag = 1
    io.unsorted_log_queue_state = fin[0]; io.intermediate_sorted_queue_state = fin[1]
    w = StorageDeduplicatorInstanceWitness(io, dev(u), pu, dev(s), d_ts, psd, None)
    trace = torch.empty((abi.ST_COLS["NUM_COLS"], n), dtype=torch.int64, device="cuda")
    K = abi.ST_COLS
    whole = sort_and_deduplicate_storage_access_entry_point(eng, w, n, trace_out=trace, raise_on_unsatisfied=False)  # un-hinted result queue: builds the tails
    assert whole.status.code == 0, (whole.status.code, hex(whole.status.failed_checks))
    tails = pushes_from_trace(trace, None, [K["SHOULD_PUSH"]], [K["RESULT_TAIL"]], 4)
    fin_tail = torch.tensor(np.array(list(whole.closed_form_input.final_sorted_queue_state.tail), dtype=np.uint64).view(np.int64), device="cuda").reshape(1, 4)
    if int(whole.closed_form_input.final_sorted_queue_state.length) == len(tails) + 1:
        tails = torch.cat([tails, fin_tail])
    w.result_queue_tails = tails.contiguous()
    cum = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), torch.cumsum((trace[K["SHOULD_PUSH"]] != 0).to(torch.int64), 0)])
    offs = [int(cum[sharding.row_range(n, r, world)[0]]) for r in range(world)]
    ms_whole, whole = timed(lambda: sort_and_deduplicate_storage_access_entry_point(eng, w, n, trace_out=trace, raise_on_unsatisfied=False), steps=3, warmup=1)
    assert whole.status.code == 0
    del trace
    torch.cuda.empty_cache()

    def step():
        return sharding.storage_validity_row_sharded(eng, w, n, rank, world, offs, device="cuda")

    for _ in range(2):
        got, (lo, hi) = step()
    assert got.status.code == 0, (rank, got.status.code, hex(got.status.failed_checks), got.status.first_bad_row)
    assert got.commitment.tolist() == whole.commitment.tolist(), "the cut instance must end with the whole instance's commitment"
    assert bytes(got.closed_form_input.hidden_fsm_output) == bytes(whole.closed_form_input.hidden_fsm_output)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": f"C4 storage_validity_by_grand_product, ONE instance of 2^{log2rows} rows cut by row range over {world} GPU(s)",
                          "ms": float(ms.item()), "rows_per_s": n / float(ms.item()) * 1e3, "ms_whole_instance_one_gpu": ms_whole,
                          "speedup_vs_one_gpu": ms_whole / float(ms.item()), "push_offsets": offs, "result_pushes": len(tails),
                          "same_commitment_and_fsm_output_as_the_whole_instance": True, "queue_hint_setup_s": setup,
                          "collective": "one all-gather of {4 products, push count, status, FSM output record} per rank; fix-up: 8 columns x 4 field elements"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "c1"
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    eng = Engine(int(os.environ.get("LOCAL_RANK", "0")))
    eng.set_stream(torch.cuda.current_stream())
    if what == "c1":
        c1(eng)
    elif what == "c3":
        c3(eng, int(sys.argv[2]) if len(sys.argv) > 2 else 18)
    elif what == "c4":
        c4(eng, int(sys.argv[2]) if len(sys.argv) > 2 else 20)
    elif what == "c4rows":
        c4rows(eng, int(sys.argv[2]) if len(sys.argv) > 2 else 20)
    elif what == "dq":
        dq(eng, int(sys.argv[2]) if len(sys.argv) > 2 else 20)
    elif what == "dmx":
        dmx(eng, int(sys.argv[2]) if len(sys.argv) > 2 else 18)
    elif what == "cu":
        cu(eng)
    elif what == "gp":
        gp(eng, int(sys.argv[2]) if len(sys.argv) > 2 else 22)
