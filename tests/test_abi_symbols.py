"""The C-ABI library loads without a GPU and exports every symbol include/zkc_b200.h declares
(no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from era_zkevm_circuits_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "zkc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zkc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(abi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), s
    assert set(abi.SIGNATURES) == set(syms)


def test_struct_layouts_match_header():
    # sizes the CUDA side static-asserts on (csrc/ram_permutation.cu)
    assert C.sizeof(abi.QueueState12) == 200
    assert C.sizeof(abi.RamInputData) == 408
    assert C.sizeof(abi.RamFsm) == 32 + 400 + 64
    assert C.sizeof(abi.Status) == 24
    assert abi.MEMORY_QUERY_DTYPE.itemsize == 64


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    lib = abi.load_library()
    h = C.c_void_p()
    assert lib.zkc_create(0, C.byref(h)) == abi.ZKC_ERR_NO_DEVICE
    assert lib.zkc_version().startswith(b"zkc_b200")


def test_vm_columns_mirror_header():
    """abi.VM_COLS / VM_CHK are hand-written mirrors of the header's enum zkc_vm_col and ZKC_VM_CHK_* defines"""
    text = open(os.path.join(ROOT, "include", "zkc_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", text[text.index("enum zkc_vm_col {"):], flags=re.S)
    body = body[:body.index("};")]
    cols = {m.group(1): int(m.group(2)) for m in re.finditer(r"ZKC_VM_([A-Z0-9_]+)\s*=\s*(\d+)", body)}
    assert cols == abi.VM_COLS
    chk = {m.group(1): 1 << int(m.group(2)) for m in re.finditer(r"#define ZKC_VM_CHK_([A-Z_]+) \(1u << (\d+)\)", text)}
    assert chk == abi.VM_CHK
    assert abi.VM_COMPACT_COLS == abi.VM_COLS["NUM_COLS"] - 117 and abi.VM_COMPACT_OP_AUX == abi.VM_COLS["OP_AUX"] - 117
    assert C.sizeof(abi.VmOptions) == 24 and C.sizeof(abi.VmCycleWitness) == 176 and C.sizeof(abi.VmCallstackWitness) == 336


def test_decommit_sorter_mirrors_header():
    """abi.DQ_COLS / DQ_CHK / struct sizes against the header (gcc computes the sizes)"""
    import subprocess
    import tempfile
    text = open(os.path.join(ROOT, "include", "zkc_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", text[text.index("enum zkc_decommit_sorter_col {"):], flags=re.S)
    body = body[:body.index("};")]
    cols = {m.group(1): int(m.group(2)) for m in re.finditer(r"ZKC_DQ_([A-Z0-9_]+)\s*=\s*(\d+)", body)}
    assert cols == abi.DQ_COLS
    chk = {m.group(1): 1 << int(m.group(2)) for m in re.finditer(r"#define ZKC_DQ_CHK_([A-Z_]+) \(1u << (\d+)\)", text)}
    assert chk == abi.DQ_CHK
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "%s"\nint main(void){printf("%%zu %%zu %%zu\\n", sizeof(zkc_decommit_query), '
                             'sizeof(zkc_decommit_sorter_fsm), sizeof(zkc_decommit_sorter_closed_form));return 0;}\n'
                             % os.path.join(ROOT, "include", "zkc_b200.h"))
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-o", exe, src])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [abi.DECOMMIT_QUERY_DTYPE.itemsize, C.sizeof(abi.DecommitSorterFsm), C.sizeof(abi.DecommitSorterClosedForm)]
    assert C.sizeof(abi.DecommitQuery) == 48


def test_demux_mirrors_header():
    import subprocess
    import tempfile
    text = open(os.path.join(ROOT, "include", "zkc_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", text[text.index("enum zkc_demux_col {"):], flags=re.S)
    body = body[:body.index("};")]
    cols = {m.group(1): int(m.group(2)) for m in re.finditer(r"ZKC_DMX_([A-Z0-9_]+)\s*=\s*(\d+)", body)}
    assert cols == abi.DMX_COLS
    chk = {m.group(1): 1 << int(m.group(2)) for m in re.finditer(r"#define ZKC_DMX_CHK_([A-Z_]+) \(1u << (\d+)\)", text)}
    assert chk == abi.DMX_CHK
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "%s"\nint main(void){printf("%%zu %%zu %%zu\\n", sizeof(zkc_demux_fsm), '
                             'sizeof(zkc_demux_closed_form), sizeof(zkc_demux_options));return 0;}\n'
                             % os.path.join(ROOT, "include", "zkc_b200.h"))
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-o", exe, src])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(abi.DemuxFsm), C.sizeof(abi.DemuxClosedForm), C.sizeof(abi.DemuxOptions)]


@pytest.mark.parametrize("enum,prefix,cols,chk", [
    ("zkc_ram_col", "RAM", "RAM_COLS", "RAM_CHK"), ("zkc_events_col", "EV", "EV_COLS", "EV_CHK"),
    ("zkc_storage_col", "ST", "ST_COLS", "ST_CHK"), ("zkc_keccak_col", "KC", "KC_COLS", "KC_CHK"),
    ("zkc_sha256_col", "SH", "SH_COLS", None)])
def test_column_enums_mirror_header(enum, prefix, cols, chk):
    """every name the Python mirror uses has the header's value (the mirrors may name a subset of the columns)"""
    text = open(os.path.join(ROOT, "include", "zkc_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", text[text.index("enum %s {" % enum):], flags=re.S)
    body = body[:body.index("};")]
    header = {m.group(1): int(m.group(2)) for m in re.finditer(r"ZKC_%s_([A-Z0-9_]+)\s*=\s*(\d+)" % prefix, body)}
    mirror = getattr(abi, cols)
    assert mirror["NUM_COLS"] == header["NUM_COLS"]
    for name, value in mirror.items():
        if name in header:
            assert header[name] == value, name
    assert len(set(mirror) & set(header)) >= min(len(mirror), 8)
    if chk:
        hchk = {m.group(1): 1 << int(m.group(2)) for m in re.finditer(r"#define ZKC_%s_CHK_([A-Z_0-9]+) \(1u << (\d+)\)" % prefix, text)}
        for name, value in getattr(abi, chk).items():
            assert hchk[name] == value, name


def test_code_unpacker_mirrors_header():
    import subprocess
    import tempfile
    text = open(os.path.join(ROOT, "include", "zkc_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", text[text.index("enum zkc_code_unpacker_col {"):], flags=re.S)
    body = body[:body.index("};")]
    cols = {m.group(1): int(m.group(2)) for m in re.finditer(r"ZKC_CU_([A-Z0-9_]+)\s*=\s*(\d+)", body)}
    assert cols == abi.CU_COLS
    chk = {m.group(1): 1 << int(m.group(2)) for m in re.finditer(r"#define ZKC_CU_CHK_([A-Z_]+) \(1u << (\d+)\)", text)}
    assert chk == abi.CU_CHK
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "%s"\nint main(void){printf("%%zu %%zu %%zu\\n", sizeof(zkc_code_decommittment_fsm), '
                             'sizeof(zkc_code_unpacker_fsm), sizeof(zkc_code_unpacker_closed_form));return 0;}\n'
                             % os.path.join(ROOT, "include", "zkc_b200.h"))
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-o", exe, src])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(abi.CodeDecommittmentFsm), C.sizeof(abi.CodeUnpackerFsm), C.sizeof(abi.CodeUnpackerClosedForm)]


def test_evaluator_violation_bits_mirror_header():
    """the relation-family bits of the ten trace evaluators (ZKC_<X>V_*): abi.py's hand-written dicts against the header's defines"""
    text = open(os.path.join(ROOT, "include", "zkc_b200.h")).read()
    for prefix, mirror in (("RAMV", abi.RAMV), ("EVV", abi.EVV), ("STV", abi.STV), ("DQV", abi.DQV), ("DMXV", abi.DMXV), ("SHV", abi.SHV),
                           ("CUV", abi.CUV), ("LHV", abi.LHV), ("KCV", abi.KCV)):
        bits = {m.group(1): 1 << int(m.group(2)) for m in re.finditer(r"#define ZKC_%s_([A-Z_]+) \(1u << (\d+)\)" % prefix, text)}
        assert bits and bits == mirror, (prefix, bits, mirror)
        assert len(set(bits.values())) == len(bits)
