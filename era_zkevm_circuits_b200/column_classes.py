"""Per-column allocation classes of every trace (zkc_col_class: FIELD / BOOLEAN / U8 / U16 / U32) for the generic
allocation-check evaluator zkc_check_trace_columns: what gadget type the reference allocates each cell as (Boolean::allocate,
UInt8 / UInt16 / UInt32::allocate_checked, Num) at the reference line the column enum in include/zkc_b200.h cites.  A class is
the STRONGEST range the reference's own allocation enforces; where a value is a Num in the reference (hash outputs, encodings,
accumulators) it is FIELD.  tests/test_gpu_column_checks.py runs every table over oracle traces (a wrong class fails there)."""
import ctypes as C

import numpy as np

from . import abi
from .engine import Engine, ZkcError, on_device, ptr

F, B, U8, U16, U32 = 0, 1, 2, 3, 4
CLASS_NAMES = ("FIELD", "BOOLEAN", "U8", "U16", "U32")

LOG_ITEM = [U32] * 29 + [U8, B, B, B, U8, U32, U32]      # flatten order, base_structures/log_query/mod.rs:62-101
DECOMMIT_ITEM = [U32] * 8 + [U32, B, U32]                # base_structures/decommit_query/mod.rs:133-149
MEMORY_ITEM = [U32, U32, U32, B, B] + [U32] * 8          # base_structures/memory_query/mod.rs:52-68


def _table(parts, total):
    out = []
    for p in parts:
        out += p
    assert len(out) == total, (len(out), total)
    return np.array(out, dtype=np.uint8)


def _gadget_block():
    cls = []
    for name, width in abi.VMG_WIDTHS.items():
        if name.endswith("_BYTES") or name == "BINOP_ALL_RESULTS":
            k = U8
        elif name == "BINOP_COMPOSITE":
            k = F                                          # and | or << 16 | xor << 32: a BinopTable row, 40 bits
        elif name == "SHIFT_INVERTED":
            k = U16                                        # 256 - shift: 256 for a zero shift
        elif name in ("SHIFT_AMOUNT", "SHIFT_FULL"):
            k = U8
        elif width == 1 or name.endswith("_IS_ZERO") or name == "ADDREL_CARRY":
            k = B
        else:
            k = U32
        cls += [k] * width
    return np.array(cls, dtype=np.uint8)


def _state_gadget_block():
    """ZKC_VM_STATE_GADGET_COLUMNS: booleans everywhere except the UInt32 results / selected limbs, the UInt16 jump destination"""
    u32 = {"PTR_ADD_RESULT", "PTR_SUB_RESULT", "PTR_SHRINK_RESULT", "PTR_LOW_IF_ADD", "PTR_LOW_IF_ADD_OR_SUB", "PTR_96_128_IF_SHRINK", "PTR_HIGHEST_128",
           "PTR_LOWEST32", "PTR_96_128", "CTX_INCREMENTED_TX_NUMBER", "CTX_META_HIGHEST", "CTX_LOW_U32", "CTX_RESULT_128", "CTX_RESULT_160_THIS",
           "CTX_RESULT_160_CALLER", "CTX_RESULT_160_CODE", "CTX_RESULT_256"}
    cls = []
    for name, width in abi.VMS_WIDTHS.items():
        if name == "PTR_DST0":
            cls += [B] + [U32] * 8
        elif name == "JUMP_DST":
            cls += [U16]
        else:
            cls += [U32 if name in u32 else B] * width
    return np.array(cls, dtype=np.uint8)


def _memory_sponge_block():
    """ZKC_VM_MEMORY_SPONGE_COLUMNS: field elements (encodings, sponge states) but the SELECTED flag and the three queue lengths"""
    cls = []
    for name, width in abi.VMQ_WIDTHS.items():
        cls += [B if name == "SELECTED" else U32 if name.endswith("_LENGTH_AFTER") else F] * width
    return np.array(cls, dtype=np.uint8)


def _prestate_block():
    """ZKC_VM_PRESTATE_COLUMNS: Booleans but the UInt16 pc / indices / low halves, the UInt32 timestamps / pages / opcode halves, and the
    VMRegister groups (is_pointer + eight UInt32 limbs)"""
    u16 = {"PC_PLUS_ONE", "SRC0_REG_LOWEST", "DST0_REG_LOWEST", "SRC_INDEX_FOR_ABSOLUTE", "SRC_INDEX_FOR_RELATIVE", "DST_INDEX_FOR_ABSOLUTE",
           "DST_INDEX_FOR_RELATIVE_WITH_PUSH", "DST_INDEX_FOR_RELATIVE", "DST_INDEX_SOMEWHAT_RELATIVE"}
    u32 = {"TIMESTAMPS", "NEXT_CYCLE_TIMESTAMP", "OPCODE_SELECT_CHAIN", "DST0_REG_LOW_CHAIN", "STACK_PAGE", "HEAP_PAGE", "AUX_HEAP_PAGE"}
    registers = {"DRAFT_SRC0_CHAIN", "SRC1_REGISTER_CHAIN", "SRC0_AFTER_USE_REG", "SRC0_AFTER_USE_IMM", "SRC0_SWAPPED", "SRC1_SWAPPED"}
    cls = []
    for name, width in abi.VMP_WIDTHS.items():
        if name in registers:
            cls += ([B] + [U32] * 8) * (width // 9)
        else:
            cls += [U16 if name in u16 else U32 if name in u32 else B] * width
    return np.array(cls, dtype=np.uint8)


def _writeback_block():
    """ZKC_VM_WRITEBACK_COLUMNS: Booleans but the r2 word of a far call (0..3, a UInt32 limb) and the UInt32 limbs of the value chains"""
    cls = []
    for name, width in abi.VMW_WIDTHS.items():
        cls += [U32 if name.startswith("VALUE_AFTER_") or name == "FAR_CALL_NEW_R2_LOW" else B] * width
    return np.array(cls, dtype=np.uint8)


TABLES = {
    "ram_permutation": lambda: _table([[B] * 3, MEMORY_ITEM, [F] * 8, [F] * 12, [U32], MEMORY_ITEM, [F] * 8, [F] * 12, [U32], [B] * 3, [U32],
                                       [U32] * 3, [B] * 3, [B] * 3, [B] * 10, [F] * 32, [F] * 4, [F] * 4, [U8] * 24, [F] * 2, [F], [F, F], [F] * 3,
                                       [F] * 2, [F] * 2, [B] * 2, [F] * 8, [F] * 8, [B] * 8, [U32] * 8, [F] * 8, [B] * 8, [F, F]], abi.RAM_COLS["NUM_COLS"]),
    "log_sorter": lambda: _table([[B] * 3, LOG_ITEM, [F] * 20, [F] * 4, [U32], LOG_ITEM, [F] * 20, [F] * 4, [U32], [F] * 88, [U32], [B] * 11,
                                  [F] * 20, [F] * 36, [F] * 4, [U32]], abi.EV_COLS["NUM_COLS"]),
    "storage_validity": lambda: _table([[B] * 3, [U32], LOG_ITEM, [F] * 20, [F], [F] * 4, [U32], LOG_ITEM + [U32], [F] * 20, [F] * 4, [U32], [B],
                                        [F] * 88, [U32] * 13, [B] * 13, [B] * 13, [B, B], [U32], [B] * 10, [F] * 20, [F] * 36, [F] * 4, [U32],
                                        [U32] * 8, [U32] * 8, [U32], [B] * 10], abi.ST_COLS["NUM_COLS"]),
    "sort_decommittment_requests": lambda: _table([[B] * 3, DECOMMIT_ITEM, [F] * 8, [F] * 12, [U32], DECOMMIT_ITEM, [F] * 8, [F] * 12, [U32], [F] * 40,
                                                   [U32] * 9, [B] * 9, [B] * 9, [B] * 6, DECOMMIT_ITEM, [F] * 8, [F] * 12, [U32], [U32]],
                                                  abi.DQ_COLS["NUM_COLS"]),
    "demux_log_queue": lambda: _table([[B] * 2, LOG_ITEM, [F] * 20, [F] * 4, [U32], [B] * 4, [B] * 3, [B, B], [B] * 6, [B], [F] * 4, [U32], [F] * 36,
                                       [F] * 24, [U32] * 6], abi.DMX_COLS["NUM_COLS"]),
    "keccak256_round_function": lambda: _table([[B] * 4, LOG_ITEM, [F] * 4, [U32], [U32] * 5 + [B], [U32, U32], [B] * 3,
                                                ([U32, U8, U8, B] + [U32] * 8 + [F] * 12 + [U32, U32, U32, U8]) * 6, [B], [U8], [B] * 3, [U8] * 136,
                                                [U8] * 200, [B], [U32] * 8, [F] * 12, [U32], [B] * 4, [U8] * 192], abi.KC_COLS["NUM_COLS"]),
    "sha256_round_function": lambda: _table([[B] * 3, LOG_ITEM, [F] * 4, [U32], [U32] * 5, [U32, U32], [B, B], ([U32] * 8 + [F] * 12 + [U32, U32]) * 2,
                                             [U32] * 16, [U32], [U32] * 8, [U32] * 8, [B], [U32] * 8, [F] * 12, [U32], [B] * 3], abi.SH_COLS["NUM_COLS"]),
    "code_unpacker_sha256": lambda: _table([[B] * 3, DECOMMIT_ITEM, [F] * 12, [U32], [B], [U16], [U16], [U32], [U32], [U32], [U32] * 8, [B], [U16],
                                            [B] * 3, [U32] * 16, [U32] * 3, [F] * 12 + [U32], [F] * 12 + [U32], [U32] * 16, [U32] * 24, [B] * 3],
                                           abi.CU_COLS["NUM_COLS"]),
    "linear_hasher": lambda: _table([[B] * 2, LOG_ITEM, [F] * 20, [F] * 4, [U32], [B, B], [U8] * 88, [B] * 3, [U32] * 100, [B]], abi.LH_COLS["NUM_COLS"]),
    "main_vm_gadget_cells": _gadget_block,
    "main_vm_state_gadget_cells": _state_gadget_block,
    "main_vm_memory_sponge_cells": _memory_sponge_block,
    "main_vm_prestate_cells": _prestate_block,
    "main_vm_writeback_cells": _writeback_block,
}


def column_classes(circuit: str) -> np.ndarray:
    return TABLES[circuit]()


def check_trace_columns(engine: Engine, circuit_or_classes, trace):
    """zkc_check_trace_columns over a finished trace [n_cols, rows] (numpy: host, torch CUDA: device).  Returns
    (violating rows, status, first bad column)."""
    cls = column_classes(circuit_or_classes) if isinstance(circuit_or_classes, str) else np.ascontiguousarray(circuit_or_classes, dtype=np.uint8)
    n_cols, rows = trace.shape
    assert len(cls) == n_cols
    st = abi.Status()
    viol, col = C.c_uint64(), C.c_uint32()
    rc = engine.lib.zkc_check_trace_columns(engine.h, ptr(trace), n_cols, rows, ptr(cls), on_device(trace), C.byref(viol), C.byref(col), C.byref(st))
    if rc in (abi.ZKC_ERR_INVALID_ARGUMENT, abi.ZKC_ERR_CUDA, abi.ZKC_ERR_NO_DEVICE):
        raise ZkcError(rc, st, "check_trace_columns")
    return viol.value, st, col.value
