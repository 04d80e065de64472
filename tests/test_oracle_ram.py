"""ram_permutation oracle against the reference's own test vector
(/root/reference/src/ram_permutation/mod.rs:418-634: limit = 16, is_start = true; the reference
asserts `check_if_satisfied`, i.e. every enforcement of the loop holds) plus negative cases and the
self-generated golden fixture (NOT reference-pinned: values depend on the unpinned Poseidon2)."""
import hashlib
import json
import os

import numpy as np

import helpers as H
import orc as O
import vectors as V
from era_zkevm_circuits_b200 import abi, synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ram_permutation.json")
K = abi.RAM_COLS


def run_ref_vector(orc, limit=16):
    u, s = V.ram_reference_vector()
    io, _, _ = H.ram_instance(orc, u, s, nondet_len=1)
    return O.ram_entry_point(orc, io, u, s, limit)


def test_reference_vector_is_satisfied(orc):
    rc, io, trace, com, st = run_ref_vector(orc)
    assert rc == abi.ZKC_OK and st.failed_checks == 0 and st.first_bad_row == -1
    assert io.completion_flag == 1
    out = io.hidden_fsm_output
    assert list(out.lhs_accumulator) == list(out.rhs_accumulator)  # it is a permutation
    assert out.num_nondeterministic_writes == 1  # the bootloader-heap write at timestamp 0
    assert trace[K["CAN_POP"]].tolist() == [1, 1, 1] + [0] * 13
    assert trace[K["IS_NONDET_WRITE"]].tolist() == [1] + [0] * 15
    # padding rows keep the accumulators (utils.rs:134-135)
    assert len(set(trace[K["GP_ACC"]][2:].tolist())) == 1
    # queue fully drained: head == tail
    assert list(out.current_unsorted_queue_state.head) == list(out.current_unsorted_queue_state.tail)
    assert out.current_unsorted_queue_state.length == 0


def test_encoding_layout(orc):
    # memory_query/mod.rs:103-221
    u, _ = V.ram_reference_vector()
    q = u[:1].copy()
    q["value"][0] = [0x11111111, 0x22222222, 0x33333333, 0x44444444, 0x55555555, 0xA4A3A2A1, 0xB4B3B2B1, 0xC4C3C2C1]
    q["rw_flag"], q["is_ptr"], q["index"] = 1, 1, 7
    e = np.zeros(8, dtype=np.uint64)
    orc.orc_memory_query_encode(O.p(q), O.p(e))
    assert int(e[0]) == 1025 and int(e[1]) == 30
    assert int(e[2]) == 7 + (1 << 32) + (1 << 33)
    assert int(e[3]) == 0x11111111 + (0xA1 << 32) + (0xA2 << 40) + (0xA3 << 48)
    assert int(e[4]) == 0x22222222 + (0xA4 << 32) + (0xB1 << 40) + (0xB2 << 48)
    assert int(e[5]) == 0x33333333 + (0xB3 << 32) + (0xB4 << 40) + (0xC1 << 48)
    assert int(e[6]) == 0x44444444 + (0xC2 << 32) + (0xC3 << 40) + (0xC4 << 48)
    assert int(e[7]) == 0x55555555


def test_negative_cases(orc):
    u, s = V.ram_reference_vector()
    chk = abi.RAM_CHK
    # sorted side out of order
    s2 = s[[1, 0, 2]].copy()
    io, _, _ = H.ram_instance(orc, u, s2, 1)
    rc, _, _, _, st = O.ram_entry_point(orc, io, u, s2, 16)
    assert rc == abi.ZKC_ERR_UNSATISFIED and st.failed_checks & chk["ASCENDING"] and st.first_bad_row == 1
    # a read that disagrees with the previous write
    s3 = s.copy(); s3["value"][2][0] ^= 1
    u3 = u.copy(); u3["value"][0][0] ^= 1
    io, _, _ = H.ram_instance(orc, u3, s3, 1)
    rc, _, _, _, st = O.ram_entry_point(orc, io, u3, s3, 16)
    assert st.failed_checks == chk["READ_CONSISTENT"] and st.first_bad_row == 2
    # not a permutation
    s4 = s.copy(); s4["value"][1][3] = 99; s4["value"][2][3] = 99
    io, _, _ = H.ram_instance(orc, u, s4, 1)
    rc, _, _, _, st = O.ram_entry_point(orc, io, u, s4, 16)
    assert st.failed_checks == chk["GRAND_PRODUCT"] and st.first_bad_row == -1
    # uninitialised read of a non-zero value (first access of the cell is a read)
    s5 = s.copy(); s5["rw_flag"][1] = 0
    u5 = u.copy(); u5["rw_flag"][1] = 0
    io, _, _ = H.ram_instance(orc, u5, s5, 1)
    rc, _, _, _, st = O.ram_entry_point(orc, io, u5, s5, 16)
    assert st.failed_checks == chk["UNINIT_READ_ZERO"] and st.first_bad_row == 1
    # wrong snapshot length
    io, _, _ = H.ram_instance(orc, u, s, 0)
    rc, _, _, _, st = O.ram_entry_point(orc, io, u, s, 16)
    assert st.failed_checks == chk["NONDET_COUNT"]
    # limit smaller than the queue: not completed, no final enforcement
    io, _, _ = H.ram_instance(orc, u, s4, 1)
    rc, io2, _, _, st = O.ram_entry_point(orc, io, u, s4, 2)
    assert rc == abi.ZKC_OK and io2.completion_flag == 0


def test_chained_instances_equal_one_instance(orc):
    u, s = synthetic.ram_trace(300, seed=5, n_cells=16, n_nondet=2)
    io, _, _ = H.ram_instance(orc, u, s, 2)
    rc, whole, tr, com, st = O.ram_entry_point(orc, io, u, s, 320)
    assert rc == 0, hex(st.failed_checks)
    rc, a, tra, _, st = O.ram_entry_point(orc, io, u, s, 128)
    assert rc == 0 and a.completion_flag == 0
    rc, b, trb, com_b, st = O.ram_entry_point(orc, H.continue_io(a), u[128:], s[128:], 192)
    assert rc == 0, hex(st.failed_checks)
    assert b.completion_flag == 1
    assert H.fsm_equal(b.hidden_fsm_output, whole.hidden_fsm_output)
    assert np.array_equal(np.concatenate([tra, trb], axis=1), tr)
    # hook_compare_witness
    exp = abi.RamClosedForm.from_buffer_copy(bytes(H.continue_io(a)))
    exp.hidden_fsm_output = b.hidden_fsm_output; exp.completion_flag = 1
    rc, *_ = O.ram_entry_point(orc, exp, u[128:], s[128:], 192, compare_expected=True)
    assert rc == 0
    exp.hidden_fsm_output.num_nondeterministic_writes += 1
    rc, *_ = O.ram_entry_point(orc, exp, u[128:], s[128:], 192, compare_expected=True)
    assert rc == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH


def test_synthetic_trace_is_valid(orc):
    u, s = synthetic.ram_trace(1 << 12, n_nondet=5)
    io, _, _ = H.ram_instance(orc, u, s, 5)
    rc, io2, trace, com, st = O.ram_entry_point(orc, io, u, s, 1 << 12)
    assert rc == 0, (hex(st.failed_checks), st.first_bad_row)
    assert io2.completion_flag == 1 and io2.hidden_fsm_output.num_nondeterministic_writes == 5
    assert trace[K["CHECK_EQUALITY"]].sum() > 500 and trace[K["SORTED_ITEM"] + 4].sum() > 0


def golden_payload(orc):
    rc, io, trace, com, st = run_ref_vector(orc)
    u, s = synthetic.ram_trace(64, seed=0xC1, n_cells=8, n_nondet=1)
    io2, _, _ = H.ram_instance(orc, u, s, 1)
    rc2, io2o, trace2, com2, st2 = O.ram_entry_point(orc, io2, u, s, 64)
    return {
        "note": "self-generated by tests/test_oracle_ram.py::golden_payload from oracle/ -- NOT reference-pinned "
                "(Poseidon2 of the un-vendored boojum crate has no known-answer vector in the reference)",
        "ref_vector": {"rc": rc, "commitment": [int(x) for x in com],
                       "lhs": [int(x) for x in io.hidden_fsm_output.lhs_accumulator],
                       "tail": [int(x) for x in io.hidden_fsm_output.current_unsorted_queue_state.tail],
                       "trace_sha256": hashlib.sha256(trace.tobytes()).hexdigest()},
        "synthetic64": {"rc": rc2, "commitment": [int(x) for x in com2],
                        "trace_sha256": hashlib.sha256(trace2.tobytes()).hexdigest()},
    }


def test_golden_fixture(orc):
    got = golden_payload(orc)
    if os.environ.get("ZKC_REGEN_GOLDEN"):
        json.dump(got, open(GOLDEN, "w"), indent=1)
    want = json.load(open(GOLDEN))
    assert got == want


def test_gadget_cells_against_big_integers(orc):
    """columns 134..221 (include/zkc_b200_ram_variables.json): byte decompositions, differences and ZeroCheckGate witnesses,
    re-derived with Python integers from the item columns of the same trace"""
    P = abi.GL_P
    u, s = synthetic.ram_trace(400, seed=5, n_cells=25, n_nondet=3)
    io, _, _ = H.ram_instance(orc, u, s, 3)
    limit = 420
    rc, io2, tr, com, st = O.ram_entry_point(orc, io, u, s, limit)
    assert rc == abi.ZKC_OK
    col = lambda name, i=0: [int(v) for v in tr[K[name] + i]]

    def zero_check(x, inv, flag):
        for a, b, f in zip(x, inv, flag):
            assert f == int(a == 0) and (a * b) % P == (0 if a == 0 else 1) and (b < P) and (a != 0 or b == 0)

    for side in ("UNSORTED", "SORTED"):
        for l in range(3):
            limb = col(side + "_ITEM", 5 + 5 + l)
            by = [col(side + "_ENC_BYTES", 4 * l + b) for b in range(4)]
            assert all(max(b) < 256 for b in by)
            assert limb == [sum(by[b][r] << (8 * b) for b in range(4)) for r in range(limit)]
    n = len(u)
    len_before = [max(n - r, 0) for r in range(limit)]
    zero_check(len_before, col("UNSORTED_LEN_INV"), col("UNSORTED_IS_EMPTY"))
    zero_check(len_before, col("SORTED_LEN_INV"), col("SORTED_IS_EMPTY"))
    zero_check(col("SORTED_ITEM", 0), col("TS_INV"), col("TS_IS_ZERO"))
    page, index = col("SORTED_ITEM", 1), col("SORTED_ITEM", 2)
    assert col("PAGE_DIFF") == [(p - 10) % P for p in page]
    zero_check(col("PAGE_DIFF"), col("PAGE_DIFF_INV"), col("PAGE_IS_BOOTLOADER_HEAP"))
    for i in range(3):
        zero_check(col("CMP_DIFF", i), col("CMP_DIFF_INV", i), col("CMP_LIMB_EQ", i))
    prev = lambda xs, first: [first] + xs[:-1]
    for i, (cur, first) in enumerate([(index, 0), (page, 0)]):
        assert col("CELL_DIFF", i) == [(a - b) % P for a, b in zip(cur, prev(cur, first))]
        zero_check(col("CELL_DIFF", i), col("CELL_DIFF_INV", i), col("CELL_LIMB_EQ", i))
    assert col("SAME_CELL") == [a & b for a, b in zip(col("CELL_LIMB_EQ", 0), col("CELL_LIMB_EQ", 1))]
    veq, vz = [1] * limit, [1] * limit
    for i in range(8):
        v = col("SORTED_ITEM", 5 + i)
        assert col("VALUE_DIFF", i) == [(a - b) % P for a, b in zip(v, prev(v, 0))] and col("VALUE_ZERO_DIFF", i) == v
        zero_check(col("VALUE_DIFF", i), col("VALUE_DIFF_INV", i), col("VALUE_LIMB_EQ", i))
        zero_check(v, col("VALUE_ZERO_DIFF_INV", i), col("VALUE_ZERO_LIMB_EQ", i))
        veq = [a & b for a, b in zip(veq, col("VALUE_LIMB_EQ", i))]
        vz = [a & b for a, b in zip(vz, col("VALUE_ZERO_LIMB_EQ", i))]
    assert veq == col("VALUE_EQUAL") and vz == col("VALUE_IS_ZERO")
    isp = col("SORTED_ITEM", 4)
    assert col("PTR_DIFF") == [(b - a) % P for a, b in zip(isp, prev(isp, 0))]
    zero_check(col("PTR_DIFF"), col("PTR_DIFF_INV"), col("PTR_EQUALITY"))


def test_variable_map_covers_every_column():
    doc = json.load(open(os.path.join(os.path.dirname(GOLDEN), "..", "..", "include", "zkc_b200_ram_variables.json")))
    seen = np.zeros(K["NUM_COLS"], dtype=int)
    for c in doc["columns"]:
        assert K[c["name"][len("ZKC_RAM_"):]] == c["column"] and ".rs:" in c["reference"]
        seen[c["column"]:c["column"] + c["width"]] += 1
    assert (seen == 1).all() and len(doc["host_resolved"]) >= 5
