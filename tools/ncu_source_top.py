#!/usr/bin/env python3
"""Top source lines of an `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` export by warp-stall samples:
  python tools/ncu_source_top.py <export.csv> [n]
prints file, line, samples, share, instructions executed, the dominant stall reasons and the source text."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
csv.field_size_limit(1 << 30)
fname, hdr, lines = "?", None, []
for r in csv.reader(open(path, errors="replace")):
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No" and len(r) > 4:
        hdr = r
    elif hdr and r[0].isdigit() and len(r) >= len(hdr) and r[4].isdigit():
        lines.append((fname, int(r[0]), r[1], dict(zip(hdr[4:], r[4:]))))
total = sum(int(m["# Samples"]) for *_, m in lines) or 1
print(f"{len(lines)} source lines with metrics, {total} warp-stall samples")
stalls = [h for h in (hdr or []) if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(m.get(s, "0") or 0) for *_, m in lines) for s in stalls}
print("stall reasons over the kernel:", ", ".join(f"{k[6:]} {v * 100 / total:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for f, ln, src, m in sorted(lines, key=lambda x: -int(x[3]["# Samples"]))[:top]:
    s = int(m["# Samples"])
    why = sorted(((int(m.get(k, "0") or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{f}:{ln:<5} {s:>7} {s * 100 / total:5.1f}%  inst {m['Instructions Executed']:>9}  {' '.join(f'{k}={v}' for v, k in why if v):<28} | {src.strip()[:110]}")
