"""include/zkc_b200_vm_variables.json (tools/gen_vm_variables.py): the machine-readable map from the named witness cells of a
main_vm cycle to the allocation site in the reference.  Checked against the header (every column of the six blocks is covered
exactly once, names / offsets / widths agree with the enums the engine compiles) and, where /root/reference is present (this
container, not the GPU box), that every cited file:line exists and the cited lines are not blank."""
import json
import os
import re
import sys

from era_zkevm_circuits_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


def load():
    return json.load(open(os.path.join(ROOT, "include", "zkc_b200_vm_variables.json")))


def test_committed_file_is_what_the_generator_writes():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_vm_variables as G
    assert G.build() == load()


def test_blocks_cover_every_column_once_and_mirror_the_header():
    doc = load()
    want = {"dense": ("ZKC_VM_", abi.VM_COLS, None), "gadget": ("ZKC_VMG_", abi.VMG_COLS, abi.VMG_WIDTHS),
            "state_gadget": ("ZKC_VMS_", abi.VMS_COLS, abi.VMS_WIDTHS), "memory_sponge": ("ZKC_VMQ_", abi.VMQ_COLS, abi.VMQ_WIDTHS),
            "prestate": ("ZKC_VMP_", abi.VMP_COLS, abi.VMP_WIDTHS),
            "writeback": ("ZKC_VMW_", abi.VMW_COLS, abi.VMW_WIDTHS)}
    assert [b["block"] for b in doc["blocks"]] == list(want)
    assert sum(b["num_columns"] for b in doc["blocks"]) == 276 + 724 + 87 + 112 + 428 + 513
    for b in doc["blocks"]:
        prefix, cols, widths = want[b["block"]]
        assert b["num_columns"] == cols["NUM_COLS"]
        nxt = 0
        for c in b["columns"]:
            name = c["name"][len(prefix):]
            assert c["name"].startswith(prefix) and c["column"] == cols[name] == nxt, c
            assert widths is None or c["width"] == widths[name], c
            assert re.fullmatch(r"src/[\w/\.]+\.rs:\d+(-\d+)?", c["reference"]), c
            nxt += c["width"]
        assert nxt == b["num_columns"], b["block"]
    # the header declares the entry point of every block
    text = open(os.path.join(ROOT, "include", "zkc_b200.h")).read()
    for b in doc["blocks"]:
        assert re.search(r"\b" + b["entry_point"] + r"\s*\(", text) and ("enum " + b["enum"]) in text, b["block"]


def test_every_cited_line_exists_in_the_reference():
    if not os.path.isdir(REFERENCE):
        return  # the GPU box has no reference tree
    cache = {}
    for b in load()["blocks"]:
        for c in b["columns"]:
            f, a, z = re.fullmatch(r"(src/[\w/\.]+\.rs):(\d+)(?:-(\d+))?", c["reference"]).groups()
            lines = cache.setdefault(f, open(os.path.join(REFERENCE, f)).read().split("\n"))
            a, z = int(a), int(z or a)
            assert 1 <= a <= z <= len(lines), c
            assert lines[a - 1].strip() and lines[z - 1].strip(), (c["name"], c["reference"], "cites a blank line")
