#!/usr/bin/env python3
"""bench.py -- one JSON line per run (driver contract).

Workload (BASELINE.json configs[1]): main_vm, ONE instance of 2^20 cycles per GPU per step (--instances / --cycles
reshape the same 2^20 cycles into a batch of shorter instances: a production main_vm instance holds ~5.6 k cycles and
instances only communicate through their closed-form inputs).  A "step" = one main_vm entry point call: every cycle
evaluated from its VmLocalState snapshot + oracle answers, memory-queue sponges, witness trace written to HBM, FSM
output and the commitment.
`value` = cycles/s over all ranks with inputs resident in HBM; `e2e` = the same call with pinned HOST buffers (H2D of
snapshots + witness, D2H of the trace + closed forms inside the timed region).  The second half of the headline metric
(constraint-evaluation GB/s vs HBM peak) is measured in the same run on the streaming constraint evaluator of the
ram_permutation trace and reported in `roofline`.  `--impl reference` times the CPU oracle (oracle/, a C restatement of
the Rust reference, which cannot be compiled in this image) on all host threads.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_INSTANCES = 1
CYCLES_PER_INSTANCE = 1 << 20
PROGRAM_LEN = 1 << 12
METRIC = "main_vm cycles/sec witness-gen at 2^20 rows; constraint-eval GB/s vs HBM peak"  # BASELINE.json's string; a row = one cycle
UNIT = "cycles/s"



def workload(n, cycles):
    return (f"main_vm, {n} instance(s) x {cycles} cycles per GPU per step, synthetic ISA table + random programs with the C2 mix "
            "of SURVEY 8d (40 % add/sub, 15 % binop, 10 % mul/div, 10 % shifts, 10 % UMA heap r/w, 5 % jumps, 5 % context/ptr, "
            "3 % log, 2 % near_call/ret; 30 % stack/code/immediate operands; far calls are not part of that mix)")


# ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of vm_cycles_kernel per launch (profiles/README.md); None until captured
VM_CYCLES_TRAFFIC = int((373.3e6 + 382.1e6) * 4)
VM_CYCLES_TRAFFIC_NOTE = ("ncu --set full dram__bytes_read+write of vm_cycles_kernel<false> at 2^18 cycles (profiles/r02_final_ncu_summary.txt, the final "
                          "kernel of round 2), scaled x4 to 2^20 cycles: 1.10x the algorithmic bytes")


def sorter_check_rooflines(eng, log2rows, peak):
    """Constraint evaluation of the two LogQuery sorter traces (C4's circuits) at 2^log2rows rows, device resident: the traces come from
    the entry points themselves (queue-state hints built on the device), the evaluators stream them once (general-purpose gates)."""
    import torch
    from era_zkevm_circuits_b200 import (EventsDeduplicatorInstanceWitness, StorageDeduplicatorInstanceWitness, abi, log_sorter_check_trace, synthetic,
                                         sort_and_deduplicate_events_entry_point, sort_and_deduplicate_storage_access_entry_point,
                                         storage_validity_check_trace)
    n = 1 << log2rows
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)).cuda()
    out = {}

    def measure(name, key, ncols, rewrite, check, row_kernels=()):
        for _ in range(2):
            viol, st_ = check()
            assert viol == 0, (name, hex(st_.failed_checks), st_.first_bad_row)
        eng.profile_reset(); eng.profile(True)
        for _ in range(5):
            rewrite()  # rewrites the trace: the next check reads cold HBM
            check()
        eng.profile(False)
        ms, k = eng.profile_query(key)
        gbs = ncols * 8 * n / (ms / k * 1e-3) / 1e9
        out[name] = {"rows": n, "algorithmic_bytes_per_row": ncols * 8, "avg_launch_ms": ms / k, "achieved": gbs, "unit": "GB/s", "peak": peak,
                     "frac": gbs / peak, "bound": "hbm"}
        for kern, perms in row_kernels:  # the witness-generation kernels of the same circuit: Poseidon2 permutations per row (integer roofline)
            kms, kk = eng.profile_query(kern)
            if kk:
                out.setdefault("_int", {})[kern + "_kernel"] = {"permutations_per_launch": perms * n, "ms_per_launch": kms / kk}

    u, s = synthetic.events_trace(n, seed=0xC4, rollback_pct=10)
    prev, fin = eng.log_queue_simulate(dev(np.concatenate([u, s])), n_queues=2)
    io = abi.EventsClosedForm(); io.start_flag = 1
    io.initial_log_queue_state = fin[0]; io.intermediate_sorted_queue_state = fin[1]
    w = EventsDeduplicatorInstanceWitness(io, dev(u), prev[:n], dev(s), prev[n:], None)
    trace = torch.empty((abi.EV_COLS["NUM_COLS"], n), dtype=torch.int64, device="cuda")
    run = lambda: sort_and_deduplicate_events_entry_point(eng, w, n, trace_out=trace, raise_on_unsatisfied=False)
    assert run().status.code == 0
    K = abi.EV_COLS
    w.result_queue_tails = trace[K["RESULT_TAIL"]:K["RESULT_TAIL"] + 4].t()[trace[K["ADD_TO_QUEUE"]] != 0].contiguous()  # hints: the pushes go row-parallel
    measure("ev_check_kernel<false> (log_sorter trace)", "ev_check", K["NUM_COLS"], run, lambda: log_sorter_check_trace(eng, io, trace, n, abi.GATES_GENERAL),
            row_kernels=(("ev_rows", 8), ("ev_push", 1)))  # 2 pops x 3 + rounds 0-1 of the push; round 2 of the push
    del trace, w
    u, s, ts = synthetic.storage_trace(n, seed=0xC4, n_cells=1 << 12)
    d_ts = torch.from_numpy(ts.astype(np.uint32).view(np.int32)).cuda()
    pu, fu = eng.log_queue_simulate(dev(u))
    psd, fs = eng.log_queue_simulate(dev(s), d_ts)
    sio = abi.StorageClosedForm(); sio.start_flag = 1
    sio.unsorted_log_queue_state = fu[0]; sio.intermediate_sorted_queue_state = fs[0]
    sw = StorageDeduplicatorInstanceWitness(sio, dev(u), pu, dev(s), d_ts, psd, None)
    strace = torch.empty((abi.ST_COLS["NUM_COLS"], n), dtype=torch.int64, device="cuda")
    srun = lambda: sort_and_deduplicate_storage_access_entry_point(eng, sw, n, trace_out=strace, raise_on_unsatisfied=False)
    assert srun().status.code == 0
    K = abi.ST_COLS
    sw.result_queue_tails = strace[K["RESULT_TAIL"]:K["RESULT_TAIL"] + 4].t()[strace[K["SHOULD_PUSH"]] != 0].contiguous()
    measure("st_check_kernel<false> (storage_validity trace)", "st_check", K["NUM_COLS"], srun,
            lambda: storage_validity_check_trace(eng, sio, strace, n, abi.GATES_GENERAL),
            row_kernels=(("st_rows", 6), ("st_push_rows", 2), ("st_push", 1)))
    return out


def measured_peaks():
    try:
        j = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for t, ln in self.lines:
            if t0 is not None and not (t0 <= t <= t1 + 0.2):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_array(eng, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = eng.lib.zkc_host_alloc(max(n, 1))
    if not p:
        raise MemoryError("zkc_host_alloc")
    buf = (C.c_uint8 * n).from_address(p)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


# ------------------------------------------------------------------------------------------- reference arm / cpu baseline
CPU_CHUNK = 1 << 14  # cycles per chained chunk instance of the CPU legs


def oracle_vm_inputs(cycles, seed=0xC2):
    """the workload of the GPU arm on the CPU: ONE instance of `cycles` cycles (out-of-circuit run by the C oracle, untimed)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc as O  # oracle: the thing MEASURED by the callers is the CPU baseline itself
    from era_zkevm_circuits_b200 import abi, isa as I
    lib = O.load()
    isa = I.Isa()
    io = abi.VmClosedForm(); io.start_flag = 1
    st = O.vm_initial_state(lib, io, isa.isa)
    rc, snaps, wit, status, cw, tail = O.vm_run(lib, isa.isa, st, I.pack_code(I.random_program(isa, PROGRAM_LEN, seed=seed)), cycles, full=True)
    assert rc == 0
    for k in range(4):
        io.rollback_queue_tail_for_block[k] = int(tail[k])
    return lib, isa, io, snaps, wit, np.ascontiguousarray(cw)


def oracle_vm_chunked(lib, isa, io, snaps, wit, cw, first, cycles, threads, chunk=CPU_CHUNK, repeat=1):
    """times the CPU oracle's main_vm entry point over cycles [first, first + cycles) of ONE instance, evaluated the way the
    reference parallelises a long run: as chained instances of `chunk` cycles (hidden_fsm_input of a chunk = the state the
    previous one ends in, main_vm/mod.rs:96-97 -- here the run's own snapshot), `threads` chunks at a time, each writing its
    witness trace.  Returns (cycles/s, seconds)."""
    import itertools
    import orc as O
    from era_zkevm_circuits_b200 import abi
    ncols = abi.VM_COLS["NUM_COLS"]
    jobs = [(s0, min(chunk, first + cycles - s0)) for s0 in range(first, first + cycles, chunk)] * repeat
    counter = itertools.count()
    traces = [np.zeros((ncols, chunk), dtype=np.uint64) for _ in range(threads)]

    def work(slot):
        while True:
            j = next(counter)
            if j >= len(jobs):
                return
            s0, ln = jobs[j]
            io2 = abi.VmClosedForm.from_buffer_copy(bytes(io))
            if s0:
                io2.start_flag = 0
                io2.hidden_fsm_input = abi.VmState.from_buffer_copy(snaps[s0].tobytes())
            com = np.zeros(4, dtype=np.uint64)
            st = abi.Status()
            opts = abi.VmOptions(0)
            tr = traces[slot] if ln == chunk else np.zeros((ncols, ln), dtype=np.uint64)
            rc = lib.orc_main_vm_entry_point(C.byref(io2), C.byref(isa.isa), O.p(snaps[s0:]), O.p(wit[s0:]), O.p(cw) if len(cw) else None, len(cw),
                                             ln, C.byref(opts), O.p(tr), O.p(com), C.byref(st))
            assert rc == 0, (rc, hex(st.failed_checks), st.first_bad_row)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return cycles * repeat / dt, dt


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    n, cycles = args.instances, args.cycles
    total = n * cycles
    lib, isa, io, snaps, wit, cw = oracle_vm_inputs(cycles)  # one instance's inputs; n > 1 evaluates them n times per step (same work per instance)
    chunk = min(CPU_CHUNK, cycles)
    for _ in range(min(args.warmup, 1)):
        oracle_vm_chunked(lib, isa, io, snaps, wit, cw, 0, cycles, cores, chunk, repeat=n)
    tot, dt = 0, 0.0
    for _ in range(args.steps):
        v, t = oracle_vm_chunked(lib, isa, io, snaps, wit, cw, 0, cycles, cores, chunk, repeat=n)
        tot += total
        dt += t
    v = tot / dt
    sample = (f"the whole workload per step: {n} instance(s) x {cycles} cycles, evaluated as {(total + chunk - 1) // chunk} chained chunk instances of "
              f"{chunk} cycles on {cores} host threads (C oracle of main_vm_entry_point, a restatement of the Rust reference, which cannot be "
              f"compiled in this image; every witness cell written; the out-of-circuit run that produces the inputs is not timed, as on the GPU arm)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload(n, cycles), "evaluated_as": f"{(total + chunk - 1) // chunk} chained chunks x {chunk} cycles on {cores} threads",
                   "row_is": "one main_vm cycle (BASELINE.json's '2^20 rows' = 2^20 cycles of one instance)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from era_zkevm_circuits_b200 import (Engine, RamPermutationCircuitInstanceWitness, abi, isa as I, main_vm_entry_point_batch,
                                         main_vm_entry_point_columns, main_vm_initial_state, main_vm_rows_to_columns, main_vm_simulate,
                                         ram_permutation_check_trace, ram_permutation_entry_point, synthetic)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream)

    # ---- inputs: out-of-circuit run of this rank's instances on the GPU (setup, untimed) -------------------------------------
    n, cycles = args.instances, args.cycles
    isa = I.Isa()
    ios, states, codes = [], [], []
    distinct_programs = [I.pack_code(I.random_program(isa, PROGRAM_LEN, seed=0xC2 + rank * 7 + k)) for k in range(8)]
    for i in range(n):
        io = abi.VmClosedForm(); io.start_flag = 1
        io.rollback_queue_tail_for_block[0] = rank * n + i  # distinct instances -> distinct commitments
        ios.append(io)
        states.append(main_vm_initial_state(eng, io, isa.isa))
        codes.append(distinct_programs[i % 8])
    sim = main_vm_simulate(eng, isa.isa, states, np.stack(codes), cycles)
    st = sim.status
    assert st.code == 0, (st.code, hex(st.failed_checks), st.first_bad_row)
    d_snaps, d_wit = sim.snapshots, sim.witness
    n_cw = max(1, int(sim.n_callstack.max()))
    d_cw = sim.callstack_witness[:, :n_cw].contiguous()
    for io, t in zip(ios, sim.rollback_tails):  # the block's rollback tail is an output of the out-of-circuit run
        for k in range(4):
            io.rollback_queue_tail_for_block[k] = int(t[k])
    ncols = abi.VM_COLS["NUM_COLS"]
    trace = torch.empty((n, ncols, cycles), dtype=torch.int64, device="cuda")
    gathered = torch.zeros((world * n, 4), dtype=torch.int64, device="cuda") if world > 1 else None

    # inputs resident in HBM in the layout the cycle kernel reads: columns (struct of arrays, zkc_vm_columns) -- what the GPU
    # out-of-circuit run / the stream expansion produce; the record (array of structs) entry point transposes first
    d_cols = main_vm_rows_to_columns(eng, d_snaps, d_wit, cycles)

    def step_device():
        coms, out, statuses, rc = main_vm_entry_point_columns(eng, ios, isa.isa, d_cols, cycles, trace_out=trace, callstack_witness=d_cw)
        assert rc == 0, [(x.code, hex(x.failed_checks), x.first_bad_row) for x in statuses][:4]
        if world > 1:  # the only exchange of the sharded job: 4 x u64 commitment per instance
            c = torch.from_numpy(coms.view(np.int64)).cuda(non_blocking=True)
            dist.all_gather_into_tensor(gathered, c)
        return coms

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t0, t1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    for _ in range(args.warmup):
        step_device()
    eng.profile_reset()
    eng.profile(True)
    l0 = eng.launches
    ms, t0, t1 = timed(step_device, args.steps)
    launches = eng.launches - l0
    eng.profile(False)
    prof = {k: eng.profile_query(k) for k in ("vm_cycles", "vm_link", "vm_sponge", "vm_sponge_far", "vm_sponge_trace", "vm_finalize")}
    value = n * cycles * world * args.steps / (ms / 1e3)
    # the timed region is K steps of a few ms: keep the same step running (untimed) until the sampler (10 Hz) has seen >= 1.5 s
    # of this load, and report the clocks over [start of the timed region, end of that tail]
    # (the step holds a collective when n_gpus > 1: every rank runs the SAME number of tail steps, derived from the max-over-ranks time)
    for _ in range(int(1500.0 / max(ms / args.steps, 0.05)) + 1):
        step_device()
    torch.cuda.synchronize()
    clocks = sampler.stop(t0, time.perf_counter()) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region + 1.5 s of the same step (untimed), nvidia-smi at 10 Hz"

    # ---- constraint evaluation of the main_vm trace itself: the row-local relations of vm_cycle, one stream over 276 columns -----
    from era_zkevm_circuits_b200 import main_vm_check_trace
    for _ in range(3):
        vviol, _ = main_vm_check_trace(eng, isa.isa, trace, cycles, n)
    assert vviol == 0
    eng.profile_reset()
    eng.profile(True)
    for _ in range(max(5, min(args.steps, 20))):
        step_device()  # rewrites the 2.3 GB trace: the next check reads cold HBM
        vviol, _ = main_vm_check_trace(eng, isa.isa, trace, cycles, n)
    eng.profile(False)
    vchk_ms, vchk_n = eng.profile_query("vm_check")

    # ---- constraint evaluation (second half of the metric): streaming evaluator over a 2^20-row ram_permutation trace --------
    rn = 1 << 20
    u, s = synthetic.ram_trace(rn, seed=0xC1 + rank, n_cells=1 << 10, n_nondet=7)
    d_both = torch.from_numpy(np.concatenate([u, s]).view(np.uint8).reshape(2 * rn, 64)).cuda()
    prev, fin = eng.memory_queue_simulate(d_both, n_queues=2)  # the two 2^20-long hash chains of the instance (setup, ~20 s)
    rio = abi.RamClosedForm(); rio.start_flag = 1
    rio.observable_input.unsorted_queue_initial_state = fin[0]
    rio.observable_input.sorted_queue_initial_state = fin[1]
    rio.observable_input.non_deterministic_bootloader_memory_snapshot_length = 7
    rtrace = torch.empty((abi.RAM_COLS["NUM_COLS"], rn), dtype=torch.int64, device="cuda")
    rw = RamPermutationCircuitInstanceWitness(rio, d_both[:rn], prev[:rn], d_both[rn:], prev[rn:])
    r = ram_permutation_entry_point(eng, rw, rn, trace_out=rtrace)
    assert r.status.code == 0
    for _ in range(3):
        ram_permutation_check_trace(eng, rio, rtrace, rn, abi.GATES_GENERAL)
    eng.profile_reset()
    eng.profile(True)
    for _ in range(max(5, min(args.steps, 20))):
        viol, _ = ram_permutation_check_trace(eng, rio, rtrace, rn, abi.GATES_GENERAL)
        assert viol == 0
        r = ram_permutation_entry_point(eng, rw, rn, trace_out=rtrace)  # rewrites 1.1 GB: the next check reads cold HBM
    eng.profile(False)
    chk_ms, chk_n = eng.profile_query("ram_check")
    rows_ms, rows_n = eng.profile_query("ram_rows")
    del d_both, prev, rtrace

    # ---- constraint evaluation of the two LogQuery sorter traces (C4's circuits), 2^17 rows each; rank 0 of a 1-GPU run only -------
    sorter_eval = None
    if world == 1:
        try:
            sorter_eval = sorter_check_rooflines(eng, 17, measured_peaks()[0])
        except Exception as ex:  # reported, never fatal for the headline numbers
            sorter_eval = {"error": repr(ex)[:300]}

    # ---- e2e: pinned host inputs -> H2D -> kernels -> D2H of witness columns + closed forms ------------------------------------
    # The reference-facing call over HOST buffers in the C ABI's transport forms (include/zkc_b200.h): the out-of-circuit run's
    # snapshots / oracle answers as a segmented input stream (dense words + change lists: what a VM run emits natively), the
    # witness back as the PACKED trace (typed columns + aux / sponge records; abi.vm_expand_packed_trace rebuilds the dense
    # trace bit-exactly).  The stream is encoded once (setup, host time reported); every timed step copies it H2D, expands,
    # evaluates, packs and copies the witness D2H, segment by segment on three streams.
    from era_zkevm_circuits_b200.main_vm import main_vm_entry_point_stream, vm_encode_input_stream, vm_packed_trace_buffers
    snaps_h, wit_h = d_snaps.cpu().numpy(), d_wit.cpu().numpy()
    hc = pinned_array(eng, (n, n_cw, C.sizeof(abi.VmCallstackWitness)), np.uint8)
    hc[:] = d_cw.cpu().numpy()
    t_enc = time.perf_counter()
    streams = [vm_encode_input_stream(eng.lib, snaps_h[i], wit_h[i], cycles, 0) for i in range(n)]
    t_enc = time.perf_counter() - t_enc
    pk = vm_packed_trace_buffers(eng, n, cycles, alloc=lambda shape, dt: pinned_array(eng, shape, dt))
    stream_bytes = sum(st_.bytes for st_ in streams)

    def step_e2e():
        coms, _ios, statuses, rc = main_vm_entry_point_stream(eng, ios, isa.isa, streams, cycles, callstack_witness=hc, out=pk)
        assert rc == 0, [(x.code, hex(x.failed_checks), x.first_bad_row) for x in statuses][:4]
        assert pk.n_aux_records <= len(pk.aux_records) and pk.n_sponge_records <= len(pk.sponge_records)
        if world > 1:
            c = torch.from_numpy(coms.view(np.int64)).cuda(non_blocking=True)
            dist.all_gather_into_tensor(gathered, c)
        return coms

    e2e_steps = max(1, min(args.steps, 10))
    c_e2e = step_e2e()
    assert np.array_equal(c_e2e, step_device()), "stream / record forms disagree on the commitments"
    ms_e2e, _, _ = timed(step_e2e, e2e_steps)
    e2e_value = n * cycles * world * e2e_steps / (ms_e2e / 1e3)
    h2d = int(stream_bytes + hc.nbytes + n * (C.sizeof(abi.VmClosedForm) + 64) + C.sizeof(abi.VmIsa))
    d2h = int(pk.nbytes_used + n * (C.sizeof(abi.VmClosedForm) + 64 + 32 + C.sizeof(abi.Status)))
    e2e_records = None
    if args.e2e_records and world == 1:
        # round 1's transport for comparison: 1 176-byte snapshot + 176-byte witness records in, COMPACT trace out
        hs = pinned_array(eng, (n, cycles + 1, C.sizeof(abi.VmState)), np.uint8)
        hw = pinned_array(eng, (n, cycles, C.sizeof(abi.VmCycleWitness)), np.uint8)
        hs[:] = snaps_h; hw[:] = wit_h
        htrace = pinned_array(eng, (n, abi.VM_COMPACT_COLS, cycles), np.uint64)
        hrec = pinned_array(eng, (int(n * cycles * 1.25), 104), np.uint8)

        def step_records():
            coms, out, statuses, rc = main_vm_entry_point_batch(eng, ios, isa.isa, hs, hw, cycles, trace_out=htrace, callstack_witness=hc,
                                                                sponge_records_out=hrec)
            assert rc == 0

        step_records()
        ms_rec, _, _ = timed(step_records, 3)
        e2e_records = {"value": n * cycles * 3 / (ms_rec / 1e3), "unit": UNIT, "ms_per_step": ms_rec / 3,
                       "h2d_bytes_per_step": int(hs.nbytes + hw.nbytes + hc.nbytes), "d2h_bytes_per_step": int(htrace.nbytes),
                       "note": "record form in, COMPACT trace out (round 1's e2e path)"}
    cpu_snaps, cpu_wit, cpu_cw = snaps_h[0], wit_h[0], np.ascontiguousarray(hc[0])
    del snaps_h, wit_h

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the FULL named witness of instance 0: the dense trace of the timed step + the five oblivious gadget-cell blocks ---------
    # (arithmetic gadgets 724, ptr / jump / context 87, per-cycle memory sponges 112, create_prestate 428, register write-back 513 columns per cycle).  A
    # secondary figure, measured after everything above and never fatal for it; 1-GPU runs only.
    full_witness = None
    if world == 1:
        try:
            from era_zkevm_circuits_b200 import main_vm_gadget_cells, main_vm_memory_sponge_cells, main_vm_prestate_cells, main_vm_state_gadget_cells, main_vm_writeback_cells
            t0_, s0_ = trace[0], d_snaps[0]
            blocks = (("vm_gadgets", main_vm_gadget_cells, abi.VMG_COLS["NUM_COLS"], lambda: main_vm_gadget_cells(eng, t0_, cycles)),
                      ("vm_state_gadgets", main_vm_state_gadget_cells, abi.VMS_COLS["NUM_COLS"], lambda: main_vm_state_gadget_cells(eng, t0_, s0_, cycles)),
                      ("vm_memory_sponges", main_vm_memory_sponge_cells, abi.VMQ_COLS["NUM_COLS"], lambda: main_vm_memory_sponge_cells(eng, t0_, s0_, cycles)),
                      ("vm_prestate", main_vm_prestate_cells, abi.VMP_COLS["NUM_COLS"], lambda: main_vm_prestate_cells(eng, t0_, s0_, cycles)),
                      ("vm_writeback", main_vm_writeback_cells, abi.VMW_COLS["NUM_COLS"], lambda: main_vm_writeback_cells(eng, isa.isa, t0_, s0_, cycles)))
            per_block, total_ms = {}, 0.0
            for name, _fn, cols, call in blocks:
                out = call(); del out                      # warm-up (first allocation of the output block)
                eng.profile_reset(); eng.profile(True)
                for _ in range(5):
                    out = call(); del out
                eng.profile(False)
                k_ms, k_n = eng.profile_query(name)
                per_block[name + "_kernel"] = {"columns": cols, "avg_launch_ms": k_ms / k_n if k_n else None,
                                               "write_GBps": cols * 8 * cycles / (k_ms / k_n * 1e-3) / 1e9 if k_n and k_ms else None}
                total_ms += k_ms / k_n if k_n else 0.0
            torch.cuda.synchronize()
            cols_total = ncols + sum(b[2] for b in blocks)
            step_ms = ms / args.steps / n + total_ms       # one instance: its share of the timed step + its five block kernels
            full_witness = {"columns_per_cycle": cols_total, "ms_per_instance": step_ms, "value": cycles / (step_ms * 1e-3), "unit": UNIT,
                            "kernels": per_block,
                            "note": "device-resident: the timed main_vm step (per instance) + zkc_main_vm_gadget_cells + zkc_main_vm_state_gadget_cells + "
                                    "zkc_main_vm_memory_sponge_cells + zkc_main_vm_prestate_cells + zkc_main_vm_writeback_cells over the trace it wrote (kernel times from CUDA events inside the library)"}
        except Exception as ex:  # reported, never fatal for the headline numbers
            full_witness = {"error": repr(ex)[:300]}

    peak, peak_src = measured_peaks()
    chk_bytes = abi.RAM_COLS["NUM_COLS"] * 8 * rn
    chk_gbs = chk_bytes / (chk_ms / chk_n * 1e-3) / 1e9 if chk_n else None
    cyc_ms, cyc_n = prof["vm_cycles"]
    # algorithmic bytes of one cycle in vm_cycles_kernel: its snapshot (read once) + oracle answers in, the trace columns it
    # writes (the 9 + 108 sponge columns belong to vm_sponge_trace_kernel) -- DESIGN.md section 5
    vm_cols = ncols - 117
    vm_bytes_per_cycle = C.sizeof(abi.VmState) + C.sizeof(abi.VmCycleWitness) + vm_cols * 8
    vm_gbs = vm_bytes_per_cycle * n * cycles / (cyc_ms / cyc_n * 1e-3) / 1e9 if cyc_n else None
    roofline = {"kernel": "vm_cycles_kernel (main_vm witness generation: one thread per cycle, warp-cooperative snapshot diff)",
                "bound": "hbm", "achieved": vm_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": vm_gbs / peak if vm_gbs else None, "frac_of_nominal_8000": vm_gbs / 8000.0 if vm_gbs else None,
                "algorithmic_bytes_per_cycle": vm_bytes_per_cycle, "avg_launch_ms": cyc_ms / cyc_n if cyc_n else None,
                "share_of_step": (cyc_ms / cyc_n) / (ms / args.steps) if cyc_n else None,
                "traffic": VM_CYCLES_TRAFFIC, "traffic_note": VM_CYCLES_TRAFFIC_NOTE}
    constraint_eval = {"kernel": "ram_check_kernel<false, 2> (constraint evaluation of a 2^20-row ram_permutation trace, 222 columns incl. the gadget cells: "
                                 "63 field multiplications per row; latency-bound, profiles/README.md)",
                       "bound": "hbm", "achieved": chk_gbs, "peak": peak, "unit": "GB/s",
                       "frac": chk_gbs / peak if chk_gbs else None, "frac_of_nominal_8000": chk_gbs / 8000.0 if chk_gbs else None,
                       "algorithmic_bytes_per_row": abi.RAM_COLS["NUM_COLS"] * 8, "avg_launch_ms": chk_ms / chk_n if chk_n else None,
                       "traffic": int((473.7e6 + 24.0e6) * 4),
                       "traffic_note": "ncu --set full dram__bytes_read+write of this kernel at 2^18 rows (profiles/r02_run19_ncu_ram_check_summary.txt), "
                                       "scaled x4 to 2^20 rows: 1.07x the algorithmic bytes (no re-reads)"}
    vchk_gbs = ncols * 8 * n * cycles / (vchk_ms / vchk_n * 1e-3) / 1e9 if vchk_n else None
    constraint_eval_vm = {"kernel": "vm_check_kernel<2> (constraint evaluation of the main_vm trace the timed step wrote: decoding, exception masks, add/sub, "
                                    "mul/div, bitwise relations, selection, sponge columns; every cell read once, row pairs with 128-bit loads)",
                          "bound": "hbm", "achieved": vchk_gbs, "peak": peak, "unit": "GB/s", "frac": vchk_gbs / peak if vchk_gbs else None,
                          "frac_of_nominal_8000": vchk_gbs / 8000.0 if vchk_gbs else None, "algorithmic_bytes_per_row": ncols * 8,
                          "avg_launch_ms": vchk_ms / vchk_n if vchk_n else None, "traffic": None}
    # ---- integer-issue roofline of the Poseidon2-bound kernels: permutations/s against the issue-slot peak ----------
    # One permutation costs ~19 000 thread instructions (472 64x64 multiplications with the 13-instruction Goldilocks reduction +
    # the lazy 96-bit linear layers; measured: smsp__inst_executed of vm_sponge_kernel x 32 / jobs, profiles/r01_ncu_full_vm_kernels_raw.csv);
    # the issue peak is SMs x 4 schedulers x 32 lanes x SM clock thread instructions/s.
    P2_INSTR = 19000.0
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    sm_hz = float((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
    p2_peak = sm_count * 128 * sm_hz / P2_INSTR
    sponge_ms, sponge_n = prof["vm_sponge"]
    sponge_perms = pk.n_sponge_records  # executed relations of one step (far-call slots included: ~0 in this mix)
    int_roofline = {"unit": "Poseidon2 permutations/s", "bound": "integer issue slots", "thread_instructions_per_permutation": P2_INSTR,
                    "peak": p2_peak, "peak_note": f"{sm_count} SMs x 128 lanes x {sm_hz / 1e6:.0f} MHz / {P2_INSTR:.0f}", "kernels": {}}
    if sponge_n:
        a = sponge_perms / (sponge_ms / args.steps * 1e-3)
        int_roofline["kernels"]["vm_sponge_kernel (5 launches per step)"] = {"permutations_per_step": sponge_perms, "ms_per_step": sponge_ms / args.steps,
                                                                              "achieved": a, "frac": a / p2_peak}
    if rows_n:
        a = 2.0 * rn / (rows_ms / rows_n * 1e-3)
        int_roofline["kernels"]["ram_rows_kernel (2 queue pops per row + 4 x 8 FMA chains, scan)"] = {
            "permutations_per_launch": 2 * rn, "ms_per_launch": rows_ms / rows_n, "achieved": a, "frac": a / p2_peak}
    if isinstance(sorter_eval, dict):
        for kern, v in (sorter_eval.pop("_int", None) or {}).items():
            a = v["permutations_per_launch"] / (v["ms_per_launch"] * 1e-3)
            int_roofline["kernels"][kern + " (2^17 rows)"] = dict(v, achieved=a, frac=a / p2_peak)
    kernels = {k + "_kernel": {"avg_launch_ms": v[0] / v[1], "launches_per_step": v[1] / args.steps,
                               "share_of_step": v[0] / ms} for k, v in prof.items() if v[1]}
    kernels["ram_rows_kernel (ram_permutation witness generation, 2^20 rows)"] = {"avg_launch_ms": rows_ms / rows_n if rows_n else None}
    cores = os.cpu_count() or 1
    # CPU baseline on the SAME inputs (the snapshots / oracle answers the GPU arm just evaluated), a bounded sample of them:
    # one chunk on one thread, then 4 chunks per thread on all threads
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc as O  # oracle: the thing MEASURED here is the CPU baseline itself
    olib = O.load()
    chunk = min(CPU_CHUNK, cycles)
    n1 = min(cycles, 16 * chunk)
    cpu1, t_1 = oracle_vm_chunked(olib, isa, ios[0], cpu_snaps, cpu_wit, cpu_cw, 0, n1, 1, chunk)
    n_sample = cycles
    oracle_vm_chunked(olib, isa, ios[0], cpu_snaps, cpu_wit, cpu_cw, 0, min(cycles, cores * chunk), cores, chunk)  # warm-up
    reps = max(1, int(round(3.0 * (1 << 20) / cycles)))  # ~20 core-seconds of work on a 16-thread box
    cpun, t_n = oracle_vm_chunked(olib, isa, ios[0], cpu_snaps, cpu_wit, cpu_cw, 0, n_sample, cores, chunk, repeat=reps)
    cpu_baseline = {"value": cpun, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{reps} pass(es) over cycles [0, {n_sample}) of instance 0 of the GPU arm's own inputs as {n_sample // chunk} chained chunk instances "
                              f"of {chunk} cycles on {cores} threads ({t_n:.1f} s); single thread: {cpu1:.0f} cycles/s ({n1} cycles, {t_1:.1f} s); C oracle of "
                              f"main_vm_entry_point incl. witness trace (a restatement: boojum synthesis of the same cycles does more work per cycle)",
                    "single_thread_value": cpu1}
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": workload(n, cycles), "row_is": "one main_vm cycle (BASELINE.json's '2^20 rows' = 2^20 cycles of one instance)",
                   "cycles_per_gpu_per_step": n * cycles, "instances_per_gpu": n, "trace_columns": ncols,
                   "l2_policy": f"snapshots + witness ({(d_snaps.numel() + d_wit.numel()) / 1e9:.2f} GB) and trace ({trace.numel() * 8 / 1e9:.2f} GB) "
                                "per step exceed the 126 MB L2",
                   "step": "main_vm entry point: start state, all cycles (witness columns to HBM), memory-queue sponges, FSM output + "
                           "commitment [+ NCCL all-gather of the 4-element commitments when n_gpus > 1]"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                "note": "zkc_main_vm_entry_point_stream: pinned host segmented input stream in (snapshots + oracle answers as dense words + change lists), "
                        "full witness out as the PACKED trace (typed columns + aux / sponge records) + closed forms; H2D | expand + kernels + pack | D2H "
                        "pipelined per segment",
                "input_stream_bytes_per_cycle": stream_bytes / (n * cycles), "packed_trace_bytes_per_cycle": pk.nbytes_used / (n * cycles),
                "host_encode_s_setup": t_enc, "aux_records_per_step": pk.n_aux_records, "sponge_records_per_step": pk.n_sponge_records,
                "records_form": e2e_records},
        "roofline": roofline, "constraint_eval": constraint_eval_vm, "constraint_eval_ram_permutation": constraint_eval, "constraint_eval_sorters": sorter_eval, "int_roofline": int_roofline, "kernels": kernels, "cpu_baseline": cpu_baseline,
        "full_witness": full_witness,
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instances", type=int, default=0,
                    help="main_vm instances per GPU per step; 0 = BASELINE.json's shapes: 1 (configs[1], one 2^20-cycle instance on one GPU) and, on "
                         "8 GPUs, 8 (configs[4]: 64 instances of 2^20 cycles sharded over 8 GPUs, NCCL gather of the commitments)")
    ap.add_argument("--cycles", type=int, default=CYCLES_PER_INSTANCE, help="cycles per instance")
    ap.add_argument("--e2e-records", action="store_true", help="also time round 1's record-form host call (N = 1 only)")
    args = ap.parse_args()
    if args.instances <= 0:
        args.instances = 8 if int(os.environ.get("WORLD_SIZE", "1")) == 8 and args.gpus == 8 else N_INSTANCES
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
