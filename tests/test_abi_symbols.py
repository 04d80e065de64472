"""The C-ABI library loads without a GPU and exports every symbol include/zkc_b200.h declares
(no compute calls here)."""
import ctypes as C
import os
import re

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
