"""The reference's own in-repo test inputs, restated as numpy records."""
import numpy as np

from era_zkevm_circuits_b200 import abi

BOOTLOADER_HEAP_PAGE = 10  # zkevm_opcode_defs::BOOTLOADER_HEAP_PAGE (un-vendored dependency)


def _u256(v):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)]


def _mq(ts, page, index, rw, is_ptr, value):
    q = np.zeros((), dtype=abi.MEMORY_QUERY_DTYPE)
    q["timestamp"], q["memory_page"], q["index"], q["rw_flag"], q["is_ptr"] = ts, page, index, rw, is_ptr
    q["value"] = _u256(value)
    return q


def ram_reference_vector():
    """witness_input_unsorted / witness_input_sorted, /root/reference/src/ram_permutation/mod.rs:559-634"""
    unsorted = np.array([
        _mq(1025, 30, 0, 0, 0, 1125899906842626),
        _mq(1024, 30, 0, 1, 0, 1125899906842626),
        _mq(0, BOOTLOADER_HEAP_PAGE, 695, 1, 0, 12345678),
    ], dtype=abi.MEMORY_QUERY_DTYPE)
    sorted_ = np.array([
        _mq(0, BOOTLOADER_HEAP_PAGE, 695, 1, 0, 12345678),
        _mq(1024, 30, 0, 1, 0, 1125899906842626),
        _mq(1025, 30, 0, 0, 0, 1125899906842626),
    ], dtype=abi.MEMORY_QUERY_DTYPE)
    return unsorted, sorted_
