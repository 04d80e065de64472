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
CPU_CYCLES = 1 << 12  # the CPU legs run the same programs in instances of this many cycles (cost per cycle is the same)
METRIC = "main_vm cycles/sec witness-gen at 2^20 cycles; constraint-eval GB/s vs HBM peak"
UNIT = "cycles/s"



def workload(n, cycles):
    return (f"main_vm, {n} instance(s) x {cycles} cycles per GPU per step, synthetic ISA table + random programs with the C2 mix "
            "of SURVEY 8d (40 % add/sub, 15 % binop, 10 % mul/div, 10 % shifts, 10 % UMA heap r/w, 5 % jumps, 5 % context/ptr, "
            "3 % log, 2 % near_call/ret; 30 % stack/code/immediate operands; far calls are not part of that mix)")


# ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of vm_cycles_kernel per launch (profiles/README.md); None until captured
VM_CYCLES_TRAFFIC = (416582912 + 368345600) * 4
VM_CYCLES_TRAFFIC_NOTE = ("ncu --set full dram__bytes_read+write of vm_cycles_kernel at 2^18 cycles (profiles/r01_ncu_full_vm_kernels_raw.csv, "
                          "final kernel of the round), scaled x4 to 2^20 cycles: 1.14x the algorithmic bytes")


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
def oracle_vm_job(instances, cycles, threads, seed=0xC2):
    """times the CPU oracle's main_vm entry point on `instances` independent instances, `threads` at a time.
    Returns (cycles/s, seconds).  Input construction (the out-of-circuit run) is not timed."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc as O  # oracle: the thing MEASURED here is the CPU baseline itself
    from era_zkevm_circuits_b200 import abi, isa as I
    lib = O.load()
    isa = I.Isa()
    distinct = min(instances, 4)  # a few distinct programs, reused round-robin: the work per instance is the same
    jobs = []
    for i in range(distinct):
        io = abi.VmClosedForm(); io.start_flag = 1; io.rollback_queue_tail_for_block[0] = i
        st = O.vm_initial_state(lib, io, isa.isa)
        rc, snaps, wit, status, cw, tail = O.vm_run(lib, isa.isa, st, I.pack_code(I.random_program(isa, PROGRAM_LEN, seed=seed + i)), cycles, full=True)
        assert rc == 0
        for k in range(4):
            io.rollback_queue_tail_for_block[k] = int(tail[k])
        jobs.append((io, snaps, wit, np.ascontiguousarray(cw)))
    ncols = abi.VM_COLS["NUM_COLS"]
    traces = [np.zeros((ncols, cycles), dtype=np.uint64) for _ in range(min(threads, instances))]

    def work(slot, count):
        for k in range(count):
            io, snaps, wit, cw = jobs[(slot + k) % distinct]
            io2 = abi.VmClosedForm.from_buffer_copy(bytes(io))
            com = np.zeros(4, dtype=np.uint64)
            st = abi.Status()
            opts = abi.VmOptions(0)
            rc = lib.orc_main_vm_entry_point(C.byref(io2), C.byref(isa.isa), O.p(snaps), O.p(wit), O.p(cw) if len(cw) else None, len(cw),
                                             cycles, C.byref(opts), O.p(traces[slot]), O.p(com), C.byref(st))
            assert rc == 0, (rc, hex(st.failed_checks), st.first_bad_row)

    per = [instances // threads + (1 if i < instances % threads else 0) for i in range(threads)]
    ts = [threading.Thread(target=work, args=(i, c)) for i, c in enumerate(per) if c]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return instances * cycles / dt, dt


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    inst = 64 * cores  # ~1-2 s of CPU work per step
    for _ in range(args.warmup):
        oracle_vm_job(inst, CPU_CYCLES, cores)
    tot, dt = 0, 0.0
    for _ in range(args.steps):
        v, t = oracle_vm_job(inst, CPU_CYCLES, cores)
        tot += inst * CPU_CYCLES
        dt += t
    v = tot / dt
    sample = (f"{inst} independent instances x 2^12 cycles per step on {cores} threads (C oracle of the same entry point, witness "
              f"trace written; bounded sample: the same programs cut into 2^12-cycle instances so that every host thread has work)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": {"workload": workload(args.instances, args.cycles)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from era_zkevm_circuits_b200 import (Engine, RamPermutationCircuitInstanceWitness, abi, isa as I, main_vm_entry_point_batch,
                                         main_vm_initial_state, main_vm_simulate, ram_permutation_check_trace,
                                         ram_permutation_entry_point, synthetic)
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

    def step_device():
        coms, out, statuses, rc = main_vm_entry_point_batch(eng, ios, isa.isa, d_snaps, d_wit, cycles, trace_out=trace, callstack_witness=d_cw)
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
    prof = {k: eng.profile_query(k) for k in ("vm_cycles", "vm_sponge", "vm_sponge_far", "vm_sponge_trace", "vm_finalize")}
    value = n * cycles * world * args.steps / (ms / 1e3)
    clocks = sampler.stop(t0, t1) if rank == 0 else None

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

    # ---- e2e: pinned host inputs -> H2D -> kernels -> D2H of witness columns + closed forms ------------------------------------
    hs = pinned_array(eng, (n, cycles + 1, C.sizeof(abi.VmState)), np.uint8)
    hw = pinned_array(eng, (n, cycles, C.sizeof(abi.VmCycleWitness)), np.uint8)
    hc = pinned_array(eng, (n, n_cw, C.sizeof(abi.VmCallstackWitness)), np.uint8)
    hs[:] = d_snaps.cpu().numpy(); hw[:] = d_wit.cpu().numpy(); hc[:] = d_cw.cpu().numpy()
    # e2e returns the witness in the COMPACT layout of the C ABI: 159 dense columns + one 104-byte record per enforced sponge
    # relation instead of 117 mostly-zero sponge columns (same values, fewer bytes over PCIe)
    htrace = pinned_array(eng, (n, abi.VM_COMPACT_COLS, cycles), np.uint64)
    rec_cap = int(n * cycles * 1.25)
    hrec = pinned_array(eng, (rec_cap, 104), np.uint8)
    n_records = [0]

    def step_e2e():
        coms, out, statuses, rc = main_vm_entry_point_batch(eng, ios, isa.isa, hs, hw, cycles, trace_out=htrace, callstack_witness=hc,
                                                            sponge_records_out=hrec)
        assert rc == 0 and statuses[0].reserved <= rec_cap
        n_records[0] = statuses[0].reserved
        if world > 1:
            c = torch.from_numpy(coms.view(np.int64)).cuda(non_blocking=True)
            dist.all_gather_into_tensor(gathered, c)

    e2e_steps = max(1, min(args.steps, 5))
    step_e2e()
    ms_e2e, _, _ = timed(step_e2e, e2e_steps)
    e2e_value = n * cycles * world * e2e_steps / (ms_e2e / 1e3)
    h2d = int(hs.nbytes + hw.nbytes + hc.nbytes + n * (C.sizeof(abi.VmClosedForm) + 64) + C.sizeof(abi.VmIsa))
    d2h = int(htrace.nbytes + n_records[0] * 104 + n * (C.sizeof(abi.VmClosedForm) + 64 + 32 + C.sizeof(abi.Status)))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

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
    constraint_eval = {"kernel": "ram_check_kernel<false> (constraint evaluation of a 2^20-row ram_permutation trace, streaming relations)",
                       "bound": "hbm", "achieved": chk_gbs, "peak": peak, "unit": "GB/s",
                       "frac": chk_gbs / peak if chk_gbs else None, "frac_of_nominal_8000": chk_gbs / 8000.0 if chk_gbs else None,
                       "algorithmic_bytes_per_row": abi.RAM_COLS["NUM_COLS"] * 8, "avg_launch_ms": chk_ms / chk_n if chk_n else None,
                       "traffic": (281072896 + 4868864) * 4,
                       "traffic_note": "ncu --set full dram__bytes_read+write of this kernel at 2^18 rows (profiles/r01_ncu_full_ram_kernels_raw.csv), "
                                       "scaled x4 to 2^20 rows: equals the algorithmic bytes (no re-reads)"}
    vchk_gbs = ncols * 8 * n * cycles / (vchk_ms / vchk_n * 1e-3) / 1e9 if vchk_n else None
    constraint_eval_vm = {"kernel": "vm_check_kernel (constraint evaluation of the main_vm trace: decoding, exception masks, add/sub, mul/div, "
                                    "bitwise relations, selection, sponge columns; every cell read once)",
                          "bound": "hbm", "achieved": vchk_gbs, "peak": peak, "unit": "GB/s", "frac": vchk_gbs / peak if vchk_gbs else None,
                          "frac_of_nominal_8000": vchk_gbs / 8000.0 if vchk_gbs else None, "algorithmic_bytes_per_row": ncols * 8,
                          "avg_launch_ms": vchk_ms / vchk_n if vchk_n else None, "traffic": None}
    kernels = {k + "_kernel": {"avg_launch_ms": v[0] / v[1], "launches_per_step": v[1] / args.steps,
                               "share_of_step": v[0] / ms} for k, v in prof.items() if v[1]}
    kernels["ram_rows_kernel (ram_permutation witness generation, 2^20 rows)"] = {"avg_launch_ms": rows_ms / rows_n if rows_n else None}
    cores = os.cpu_count() or 1
    cpu1, t_1 = oracle_vm_job(256, CPU_CYCLES, 1)
    cpun, t_n = oracle_vm_job(512 * cores, CPU_CYCLES, cores)
    cpu_baseline = {"value": cpun, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{512 * cores} instances x 2^12 cycles of the same workload on {cores} threads ({t_n:.1f} s); single thread: "
                              f"{cpu1:.0f} cycles/s ({t_1:.1f} s); C oracle of main_vm_entry_point incl. witness trace",
                    "single_thread_value": cpu1}
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": workload(n, cycles), "cycles_per_gpu_per_step": n * cycles, "instances_per_gpu": n, "trace_columns": ncols,
                   "l2_policy": f"snapshots + witness ({(hs.nbytes + hw.nbytes + hc.nbytes) / 1e9:.2f} GB) and trace ({trace.numel() * 8 / 1e9:.2f} GB) "
                                "per step exceed the 126 MB L2",
                   "step": "main_vm entry point: start state, all cycles (witness columns to HBM), memory-queue sponges, FSM output + "
                           "commitment [+ NCCL all-gather of the 4-element commitments when n_gpus > 1]"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / e2e_steps, "note": "pinned host snapshots + witness in; full witness out in the COMPACT layout (159 dense columns + sponge records) + closed forms; "
                        "H2D | kernels | D2H pipelined over 16 row chunks", "sponge_records_per_step": n_records[0]},
        "roofline": roofline, "constraint_eval": constraint_eval, "constraint_eval_main_vm": constraint_eval_vm, "kernels": kernels, "cpu_baseline": cpu_baseline,
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instances", type=int, default=N_INSTANCES, help="main_vm instances per GPU per step")
    ap.add_argument("--cycles", type=int, default=CYCLES_PER_INSTANCE, help="cycles per instance")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
