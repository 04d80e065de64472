#!/usr/bin/env python3
"""One line per kernel launch of an `ncu -i X.ncu-rep --page raw --csv` export with the metrics the roofline claims rest on:
  python tools/ncu_raw_summary.py <raw.csv> > profiles/rNN_ncu_summary.txt
duration, DRAM bytes read / written (and their sum as GB/s over the duration), registers, active warps, issue-slot utilisation,
lanes active per instruction, L1 / L2 hit rates, the top stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active)."""
import csv
import re
import sys

csv.field_size_limit(1 << 30)
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, default=0.0):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


def scaled(r, name, to):
    """value converted to `to` (byte / us) from the unit ncu chose for the column"""
    i = col.get(name)
    if i is None:
        return 0.0
    u = units[i].lower()
    v = val(r, name)
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
    return v * mult.get(u, 1)


stall_cols = [h for h in hdr if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio", h) or re.match(r"smsp__average_warp.*_issue_stalled_.*_per_warp_active\.pct", h)]
print(f"{'kernel':<44}{'grid':>8}{'blk':>5}{'regs':>5}{'us':>9}{'rd MB':>9}{'wr MB':>9}{'GB/s':>8}{'warps%':>7}{'issue%':>7}{'lanes':>6}{'L1hit':>6}{'L2hit':>6}  top stalls (warps per issue)")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    us = scaled(r, "gpu__time_duration.sum", "us")
    rd, wr = scaled(r, "dram__bytes_read.sum", "byte"), scaled(r, "dram__bytes_write.sum", "byte")
    stalls = sorted(((val(r, h), re.sub(r"smsp__average_warps?_issue_stalled_|_per_issue_active\.ratio|_per_warp_active\.pct", "", h)) for h in stall_cols if "per_issue_active" in h), reverse=True)[:4]
    print(f"{name[-43:]:<44}{int(val(r, 'launch__grid_size')):>8}{int(val(r, 'launch__block_size')):>5}{int(val(r, 'launch__registers_per_thread')):>5}{us:>9.1f}"
          f"{rd / 1e6:>9.1f}{wr / 1e6:>9.1f}{(rd + wr) / us / 1e3 if us else 0:>8.0f}{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):>7.1f}"
          f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):>7.1f}{val(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'):>6.1f}"
          f"{val(r, 'l1tex__t_sector_hit_rate.pct'):>6.1f}{val(r, 'lts__t_sector_hit_rate.pct'):>6.1f}  " + ", ".join(f"{n} {v:.2f}" for v, n in stalls))
