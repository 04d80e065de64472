#!/usr/bin/env python3
"""bench.py -- one JSON line per run (driver contract).

A "step" is one pass of the hot path over one batch of synthetic input PER GPU:
witness generation of one circuit instance (entry point, witness columns written to HBM) followed by
constraint evaluation of the finished trace.  `value` = loop iterations ("rows"/"cycles") per second,
whole job over all ranks, inputs resident in HBM; `e2e` = the same through the reference-facing call
with pinned HOST buffers (H2D of the inputs and D2H of the witness columns + FSM output inside the
timed region).  `--impl reference` times the CPU oracle (oracle/, a C restatement of the Rust
reference, which cannot be compiled in this image) on all host threads.
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

WORKLOAD_ROWS = 1 << 20
METRIC = "ram_permutation witness-gen + constraint-eval rows/sec at 2^20 rows per instance"
UNIT = "rows/s"


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
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
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
def oracle_ram_job(rows_per_instance, instances, threads):
    """times the CPU oracle on `instances` independent ram_permutation instances, `threads` at a time.
    Returns (rows/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc as O  # oracle: the thing MEASURED here is the CPU baseline itself
    import helpers as H
    from era_zkevm_circuits_b200 import abi, synthetic
    lib = O.load()
    u, s = synthetic.ram_trace(rows_per_instance, seed=0xC1, n_nondet=7)
    io, _, _ = H.ram_instance(lib, u, s, 7)
    traces = [np.zeros((abi.RAM_COLS["NUM_COLS"], rows_per_instance), dtype=np.uint64) for _ in range(min(threads, instances))]

    def work(slot, count):
        for _ in range(count):
            io2 = abi.RamClosedForm.from_buffer_copy(bytes(io))
            com = np.zeros(4, dtype=np.uint64)
            st = abi.Status()
            opts = abi.RamOptions(0, 0)
            rc = lib.orc_ram_permutation_entry_point(C.byref(io2), O.p(u), len(u), O.p(s), len(s), rows_per_instance,
                                                     C.byref(opts), O.p(traces[slot]), O.p(com), C.byref(st))
            assert rc == 0

    per = [instances // threads + (1 if i < instances % threads else 0) for i in range(threads)]
    ts = [threading.Thread(target=work, args=(i, c)) for i, c in enumerate(per) if c]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return rows_per_instance * instances / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rows = 1 << 16
    inst = cores  # one instance per host thread per step
    for _ in range(args.warmup):
        oracle_ram_job(rows, inst, cores)
    tot, dt = 0, 0.0
    for _ in range(args.steps):
        tot += rows * inst
        dt += oracle_ram_job(rows, inst, cores)[1]  # the job's own clock: input construction is not timed
    v = tot / dt
    sample = f"{inst} independent instances x 2^16 rows per step on {cores} threads (C oracle, witness trace written)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "ram_permutation, 2^20 rows per instance (reference arm: bounded sample of 2^16-row instances)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from era_zkevm_circuits_b200 import (Engine, RamPermutationCircuitInstanceWitness, abi, ram_permutation_check_trace,
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

    n = WORKLOAD_ROWS
    ncols = abi.RAM_COLS["NUM_COLS"]
    u, s = synthetic.ram_trace(n, seed=0xC1 + rank, n_cells=1 << 10, n_nondet=7)
    both = np.concatenate([u, s])
    d_both = torch.from_numpy(both.view(np.uint8).reshape(2 * n, 64)).cuda()
    prev, fin = eng.memory_queue_simulate(d_both, n_queues=2)  # device-side hash chains (setup, untimed)
    torch.cuda.synchronize()
    du, ds, dup, dsp = d_both[:n], d_both[n:], prev[:n], prev[n:]
    io = abi.RamClosedForm()
    io.start_flag = 1
    io.observable_input.unsorted_queue_initial_state = fin[0]
    io.observable_input.sorted_queue_initial_state = fin[1]
    io.observable_input.non_deterministic_bootloader_memory_snapshot_length = 7
    trace = torch.empty((ncols, n), dtype=torch.int64, device="cuda")
    w_dev = RamPermutationCircuitInstanceWitness(io, du, dup, ds, dsp)
    gathered = torch.zeros((world, 4), dtype=torch.int64, device="cuda") if world > 1 else None

    def step_device():
        r = ram_permutation_entry_point(eng, w_dev, n, trace_out=trace)
        viol, _ = ram_permutation_check_trace(eng, io, trace, n, abi.GATES_GENERAL)
        assert viol == 0 and r.status.code == 0
        if world > 1:  # the only exchange of the sharded job: 4 x u64 commitment per instance
            c = torch.from_numpy(r.commitment.view(np.int64)).cuda(non_blocking=True)
            dist.all_gather_into_tensor(gathered, c.reshape(1, 4))
        return r

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)  # let nvidia-smi come up so that samples fall inside the timed region
    for _ in range(args.warmup):
        step_device()
    eng.profile_reset()
    eng.profile(True)
    l0 = eng.launches
    ms = timed(step_device, args.steps)
    launches = eng.launches - l0
    eng.profile(False)
    prof = {k: eng.profile_query(k) for k in ("ram_rows", "ram_check", "ram_prologue", "ram_finalize")}
    clocks = sampler.stop() if rank == 0 else None
    value = n * world * args.steps / (ms / 1e3)

    # ---- e2e: pinned host inputs -> H2D -> kernels -> D2H of witness columns + FSM output -------------------------------
    hu = pinned_array(eng, (n,), abi.MEMORY_QUERY_DTYPE); hu[:] = u
    hs = pinned_array(eng, (n,), abi.MEMORY_QUERY_DTYPE); hs[:] = s
    hup = pinned_array(eng, (n, 12), np.uint64); hup[:] = dup.cpu().numpy().view(np.uint64)
    hsp = pinned_array(eng, (n, 12), np.uint64); hsp[:] = dsp.cpu().numpy().view(np.uint64)
    htrace = pinned_array(eng, (ncols, n), np.uint64)
    w_host = RamPermutationCircuitInstanceWitness(io, hu, hup, hs, hsp)

    def step_e2e():
        r = ram_permutation_entry_point(eng, w_host, n, trace_out=htrace)
        assert r.status.code == 0
        if world > 1:
            c = torch.from_numpy(r.commitment.view(np.int64)).cuda(non_blocking=True)
            dist.all_gather_into_tensor(gathered, c.reshape(1, 4))

    e2e_steps = max(1, min(args.steps, 5))
    step_e2e()
    ms_e2e = timed(step_e2e, e2e_steps)
    e2e_value = n * world * e2e_steps / (ms_e2e / 1e3)
    h2d = int(hu.nbytes + hs.nbytes + hup.nbytes + hsp.nbytes + C.sizeof(abi.RamClosedForm))
    d2h = int(htrace.nbytes + C.sizeof(abi.RamClosedForm) + 32 + C.sizeof(abi.Status))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    chk_ms, chk_n = prof["ram_check"]
    rows_ms, rows_n = prof["ram_rows"]
    chk_bytes = ncols * 8 * n
    rows_bytes = (2 * 64 + 2 * 96 + ncols * 8) * n
    chk_gbs = chk_bytes / (chk_ms / chk_n * 1e-3) / 1e9 if chk_n else None
    rows_gbs = rows_bytes / (rows_ms / rows_n * 1e-3) / 1e9 if rows_n else None
    roofline = {"kernel": "ram_check_kernel<false> (constraint evaluation, streaming relations)", "bound": "hbm",
                "achieved": chk_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": chk_gbs / peak if chk_gbs else None, "frac_of_nominal_8000": chk_gbs / 8000.0 if chk_gbs else None,
                "algorithmic_bytes_per_row": ncols * 8, "avg_launch_ms": chk_ms / chk_n if chk_n else None, "traffic": None}
    kernels = {
        "ram_rows_kernel (witness generation: 2 Poseidon2/row + scan; integer-ALU bound)": {
            "avg_launch_ms": rows_ms / rows_n if rows_n else None, "algorithmic_bytes_per_row": rows_bytes // n,
            "achieved_gbs": rows_gbs, "frac_of_hbm_peak": rows_gbs / peak if rows_gbs else None,
            "poseidon2_per_s": 2 * n / (rows_ms / rows_n * 1e-3) if rows_n else None},
        "ram_prologue_kernel": {"avg_launch_ms": prof["ram_prologue"][0] / max(1, prof["ram_prologue"][1])},
        "ram_finalize_kernel": {"avg_launch_ms": prof["ram_finalize"][0] / max(1, prof["ram_finalize"][1])},
    }

    cores = os.cpu_count() or 1
    cpu1, t1 = oracle_ram_job(1 << 16, 2, 1)
    cpun, tn = oracle_ram_job(1 << 16, 2 * cores, cores)
    cpu_baseline = {"value": cpun, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{2 * cores} instances x 2^16 rows of the same synthetic workload on {cores} threads ({tn:.1f} s); "
                              f"single thread: {cpu1:.0f} rows/s ({t1:.1f} s); C oracle incl. witness trace, no constraint eval",
                    "single_thread_value": cpu1}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": "ram_permutation single instance per GPU, limit = 2^20 rows (C1 trace shape at the reference's "
                               "max_trace_len); main_vm (configs[1]) is not built yet",
                   "rows_per_gpu_per_step": n, "trace_columns": ncols,
                   "l2_policy": "inputs (335 MB) and trace (1.1 GB) per step exceed the 126 MB L2",
                   "step": "entry point (witness columns to HBM) + constraint evaluation (general gates) [+ NCCL all-gather "
                           "of the 4-element commitments when n_gpus > 1]"},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / e2e_steps, "note": "pinned host inputs; full witness trace copied back to the host"},
        "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline,
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
