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


def _limbs(v, n):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def log_queries_from_fixture(items):
    """tests/golden/*_vector.json (extracted from the reference by tools/extract_reference_vectors.py) -> records.
    Address::from_low_u64_le(v) puts v's little-endian bytes into the LAST 8 bytes of the 20-byte big-endian address."""
    out = np.zeros(len(items), dtype=abi.LOG_QUERY_DTYPE)
    extra = np.zeros(len(items), dtype=np.uint32)
    for i, q in enumerate(items):
        addr_bytes = bytes(12) + int(q["address_low_u64_le"]).to_bytes(8, "little")
        out[i]["address"] = _limbs(int.from_bytes(addr_bytes, "big"), 5)
        out[i]["key"] = _limbs(int(q["key"]), 8)
        out[i]["read_value"] = _limbs(int(q["read_value"]), 8)
        out[i]["written_value"] = _limbs(int(q["written_value"]), 8)
        out[i]["tx_number_in_block"] = q["tx_number_in_block"]
        out[i]["timestamp"] = q["timestamp"]
        out[i]["flags"] = abi.lq_flags(q["aux_byte"], q["shard_id"], q["rw_flag"], q["rollback"], q["is_service"])
        extra[i] = q.get("extra_timestamp", 0)
    return out, extra


def _fixture(name):
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)))


def log_sorter_reference_vector():
    """witness_input_unsorted / witness_input_sorted, /root/reference/src/log_sorter/mod.rs:637-816"""
    f = _fixture("log_sorter_vector.json")
    return log_queries_from_fixture(f["unsorted"])[0], log_queries_from_fixture(f["sorted"])[0]


def storage_reference_vector():
    """generate_test_input_unsorted / generate_test_input_sorted,
    /root/reference/src/storage_validity_by_grand_product/test_input.rs; returns (unsorted, sorted, sorted_timestamps)"""
    f = _fixture("storage_validity_vector.json")
    s, ts = log_queries_from_fixture(f["sorted"])
    return log_queries_from_fixture(f["unsorted"])[0], s, ts


def decommit_queries_from_fixture(items):
    out = np.zeros(len(items), dtype=abi.DECOMMIT_QUERY_DTYPE)
    for i, q in enumerate(items):
        out[i]["code_hash"] = _limbs(int(q["code_hash"]), 8)
        out[i]["page"], out[i]["is_first"], out[i]["timestamp"] = q["page"], q["is_first"], q["timestamp"]
    return out


def sort_decommittments_reference_vector():
    """witness_input_unsorted / witness_input_sorted, /root/reference/src/sort_decommittment_requests/mod.rs:565-1390"""
    f = _fixture("sort_decommittments_vector.json")
    return decommit_queries_from_fixture(f["unsorted"]), decommit_queries_from_fixture(f["sorted"])


def demux_reference_vector():
    """witness_input_unsorted, /root/reference/src/demux_log_queue/mod.rs:602-923"""
    return log_queries_from_fixture(_fixture("demux_log_queue_vector.json")["records"])[0]


def code_unpacker_reference_vector():
    """test_code_unpacker_inner, /root/reference/src/code_unpacker_sha256/mod.rs:472-700: one request (versioned code hash,
    page 2368, timestamp 40973) and its 33 bytecode words.  Returns (requests [1], code_words [33, 8] uint32 limbs)."""
    f = _fixture("code_unpacker_vector.json")
    req = np.zeros(1, dtype=abi.DECOMMIT_QUERY_DTYPE)
    req["code_hash"][0] = _limbs(int(f["code_hash"]), 8)
    req["page"], req["is_first"], req["timestamp"] = f["page"], f["is_first"], f["timestamp"]
    words = np.array([_limbs(int(w), 8) for w in f["code_words"]], dtype=np.uint32)
    return req, words
