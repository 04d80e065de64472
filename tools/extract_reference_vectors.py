#!/usr/bin/env python3
"""Extracts the reference's own in-repo LogQuery test vectors into JSON fixtures (tests/golden/):
  /root/reference/src/log_sorter/mod.rs:637-816                       -> log_sorter_vector.json
  /root/reference/src/storage_validity_by_grand_product/test_input.rs -> storage_validity_vector.json
  /root/reference/src/sort_decommittment_requests/mod.rs:565-1390      -> sort_decommittments_vector.json
  /root/reference/src/demux_log_queue/mod.rs:602-923                   -> demux_log_queue_vector.json
  /root/reference/src/code_unpacker_sha256/mod.rs:632-700              -> code_unpacker_vector.json
Run in the build container (the reference is not present on the GPU box); only the extracted DATA is
committed, no reference source.  Values are kept as decimal strings / ints exactly as written there."""
import json
import os
import re
import sys

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

BOOL = {"bool_true": 1, "bool_false": 0}
U8 = {"zero_8": 0, "one_8": 1}


def parse_queries(text):
    """every `LogQuery::<F> { ... }` literal, in order; plus the enclosing outer timestamp if any"""
    out = []
    pat = re.compile(r"(?:TimestampedStorageLogRecord::<F>\s*\{\s*timestamp:\s*UInt32::allocated_constant\(cs,\s*(\d+)\),\s*record:\s*)?"
                     r"LogQuery::<F>\s*\{(.*?)\n\s*\}", re.S)
    for m in pat.finditer(text):
        body = m.group(2)
        q = {}
        a = re.search(r"address:.*?from_low_u64_le\((\d+)\)", body, re.S)
        q["address_low_u64_le"] = int(a.group(1))
        for f in ("key", "read_value", "written_value"):
            v = re.search(f + r":\s*UInt256::allocated_constant\(\s*cs,\s*U256::from_dec_str\(\s*\"(\d+)\",?\s*\)", body, re.S)
            q[f] = v.group(1)
        for f in ("rw_flag", "rollback", "is_service"):
            q[f] = BOOL[re.search(f + r":\s*(\w+)", body).group(1)]
        for f in ("aux_byte", "shard_id"):
            q[f] = U8[re.search(f + r":\s*(\w+)", body).group(1)]
        for f in ("tx_number_in_block", "timestamp"):
            v = re.search(f + r":\s*(zero_32|UInt32::allocated_constant\(cs,\s*(\d+)\))", body)
            q[f] = 0 if v.group(1) == "zero_32" else int(v.group(2))
        if m.group(1) is not None:
            q["extra_timestamp"] = int(m.group(1))
        out.append(q)
    return out


def parse_decommit_queries(text):
    """every `DecommitQuery::<F> { ... }` literal, in order"""
    out = []
    for m in re.finditer(r"DecommitQuery::<F>\s*\{(.*?)\n\s*\};", text, re.S):
        body = m.group(1)
        q = {"code_hash": re.search(r"from_dec_str\(\s*\"(\d+)\"", body, re.S).group(1),
             "page": int(re.search(r"page:\s*UInt32::allocated_constant\(cs,\s*(\d+)\)", body).group(1)),
             "is_first": BOOL[re.search(r"is_first:\s*(\w+)", body).group(1)],
             "timestamp": int(re.search(r"timestamp:\s*UInt32::allocated_constant\(cs,\s*(\d+)\)", body).group(1))}
        out.append(q)
    return out


def split_fn(text, name):
    i = text.index("fn " + name)
    j = text.find("\n    fn ", i + 1)
    k = text.find("\npub fn ", i + 1)
    ends = [e for e in (j, k) if e > 0]
    return text[i:min(ends)] if ends else text[i:]


def main():
    ls = open(os.path.join(REF, "log_sorter", "mod.rs")).read()
    fixture = {"source": "reference src/log_sorter/mod.rs witness_input_unsorted / witness_input_sorted (test :494)",
               "unsorted": parse_queries(split_fn(ls, "witness_input_unsorted")),
               "sorted": parse_queries(split_fn(ls, "witness_input_sorted"))}
    assert len(fixture["unsorted"]) == len(fixture["sorted"]) == 4, (len(fixture["unsorted"]), len(fixture["sorted"]))
    json.dump(fixture, open(os.path.join(OUT, "log_sorter_vector.json"), "w"), indent=1)
    sv = open(os.path.join(REF, "storage_validity_by_grand_product", "test_input.rs")).read()
    fixture = {"source": "reference src/storage_validity_by_grand_product/test_input.rs (test mod.rs:1035)",
               "unsorted": parse_queries(split_fn(sv, "generate_test_input_unsorted")),
               "sorted": parse_queries(split_fn(sv, "generate_test_input_sorted"))}
    assert len(fixture["unsorted"]) == len(fixture["sorted"]) == 16, (len(fixture["unsorted"]), len(fixture["sorted"]))
    assert all("extra_timestamp" in q for q in fixture["sorted"])
    json.dump(fixture, open(os.path.join(OUT, "storage_validity_vector.json"), "w"), indent=1)
    sd = open(os.path.join(REF, "sort_decommittment_requests", "mod.rs")).read()
    fixture = {"source": "reference src/sort_decommittment_requests/mod.rs witness_input_unsorted / witness_input_sorted (test :420, limit 16)",
               "unsorted": parse_decommit_queries(split_fn(sd, "witness_input_unsorted")),
               "sorted": parse_decommit_queries(split_fn(sd, "witness_input_sorted"))}
    assert len(fixture["unsorted"]) == len(fixture["sorted"]) == 29, (len(fixture["unsorted"]), len(fixture["sorted"]))
    json.dump(fixture, open(os.path.join(OUT, "sort_decommittments_vector.json"), "w"), indent=1)
    dm = open(os.path.join(REF, "demux_log_queue", "mod.rs")).read()
    fixture = {"source": "reference src/demux_log_queue/mod.rs witness_input_unsorted (test :482, limit 16)",
               "records": parse_queries(split_fn(dm, "witness_input_unsorted"))}
    assert len(fixture["records"]) >= 8, len(fixture["records"])
    json.dump(fixture, open(os.path.join(OUT, "demux_log_queue_vector.json"), "w"), indent=1)
    cu = open(os.path.join(REF, "code_unpacker_sha256", "mod.rs")).read()
    req = re.search(r"page:\s*(\d+),\s*is_first:\s*(true|false),\s*timestamp:\s*(\d+)", split_fn(cu, "create_request_queue_witness"))
    fixture = {"source": "reference src/code_unpacker_sha256/mod.rs test_code_unpacker_inner (:472, limit 40): get_code_hash_witness, "
                         "get_byte_code_witness, create_request_queue_witness",
               "code_hash": re.search(r'from_dec_str\(\s*"(\d+)"', split_fn(cu, "get_code_hash_witness"), re.S).group(1),
               "page": int(req.group(1)), "is_first": int(req.group(2) == "true"), "timestamp": int(req.group(3)),
               "code_words": re.findall(r'"(\d+)"', split_fn(cu, "get_byte_code_witness"))}
    assert len(fixture["code_words"]) == 33, len(fixture["code_words"])
    json.dump(fixture, open(os.path.join(OUT, "code_unpacker_vector.json"), "w"), indent=1)
    print("ok")


if __name__ == "__main__":
    sys.exit(main())
