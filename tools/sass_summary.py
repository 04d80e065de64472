#!/usr/bin/env python3
"""Per-kernel SASS summary of the built engine (cuobjdump -sass): instruction count, the memory mnemonics that matter for the
roofline claims (LDG.E.128 row-pair streaming, UBLKCP / UTMALDG bulk + tensor TMA copies, SYNCS mbarrier waits, STG widths,
local-memory spills) and two sample lines of each TMA instruction.  usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "era_zkevm_circuits_b200", "libzkc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: re.sub(r"\((bool|int|unsigned int)\)", "", subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()).split("(")[0].replace("void ", "")
KEYS = ["LDG.E.128", "LDG.E.64", "LDG.E ", "STG.E.128", "STG.E.64", "STG.E ", "UBLKCP", "UTMALDG", "SYNCS", "LDS", "STS", "LDL", "STL", "SHFL", "IMAD.WIDE.U32", "ATOMG", "RED"]
name, stats, samples = None, {}, collections.defaultdict(list)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = demangle(m.group(1))
        stats[name] = collections.Counter()
        continue
    if name is None or not re.search(r"/\*[0-9a-f]{4,}\*/", line):
        continue
    ins = line.split("*/", 1)[1].strip()
    if not ins or ins.startswith("/*"):
        continue
    stats[name]["instr"] += 1
    for k in KEYS:
        if k in ins:
            stats[name][k.strip()] += 1
    for k in ("UBLKCP", "UTMALDG", "SYNCS"):
        if k in ins and len(samples[(name, k)]) < 2:
            samples[(name, k)].append(ins.split(";")[0])
print(f"# {os.path.basename(lib)}: {len(stats)} kernels (cuobjdump -sass, sm_100a)")
print(f"{'kernel':<58}{'instr':>7} " + " ".join(f"{k.strip():>9}" for k in KEYS))
for n, c in sorted(stats.items(), key=lambda kv: -kv[1]["instr"]):
    print(f"{n[-57:]:<58}{c['instr']:>7} " + " ".join(f"{c[k.strip()]:>9}" for k in KEYS))
print("\n# TMA / mbarrier instructions as emitted")
for (n, k), ls in samples.items():
    for l in ls:
        print(f"{n[-50:]:<52} {l}")
