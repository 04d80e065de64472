"""keccak256 precompile oracle.  PINNED: the reference's own tests (keccak256_round_function/mod.rs:1096-1144) compare
the circuit's digest with sha3::Keccak256 for ten (length, unalignment) cases, limit = 2; reproduced here against
(a) standard Keccak-256 known answers, (b) hashlib's SHA3-256 driving the same permutation with the 0x06 domain byte."""
import hashlib

import numpy as np
import pytest

import orc as O
from era_zkevm_circuits_b200 import abi, synthetic

K = abi.KC_COLS
REFERENCE_CASES = [(50, 0), (135, 0), (200, 0), (180, 0), (136, 0), (50, 31), (135, 31), (136, 31), (200, 31), (166, 22)]


def keccak256(orc, msg: bytes) -> bytes:
    out = np.zeros(32, dtype=np.uint8)
    buf = np.frombuffer(msg, dtype=np.uint8).copy() if msg else np.zeros(1, dtype=np.uint8)
    orc.orc_keccak256(O.p(buf), len(msg), O.p(out))
    return out.tobytes()


def sha3_256_via_oracle_permutation(orc, msg: bytes) -> bytes:
    st = np.zeros(25, dtype=np.uint64)
    padded = bytearray(msg) + bytearray(136 - len(msg) % 136)
    padded[len(msg)] ^= 0x06
    padded[-1] ^= 0x80
    for off in range(0, len(padded), 136):
        blk = np.frombuffer(bytes(padded[off:off + 136]), dtype="<u8")
        st[:17] ^= blk
        orc.orc_keccak_f1600(O.p(st))
    return st[:4].astype("<u8").tobytes()


def test_permutation_against_hashlib_and_kats(orc):
    for msg in [b"", b"abc", bytes(range(135)), bytes(range(136)), bytes(200), b"x" * 1000]:
        assert sha3_256_via_oracle_permutation(orc, msg) == hashlib.sha3_256(msg).digest()
    assert keccak256(orc, b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert keccak256(orc, b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"


def single_call_instance(orc, msg, unalignment, in_page=123, out_page=456):
    req = np.array([synthetic.precompile_call(abi.KECCAK256_PRECOMPILE_ADDRESS, unalignment, len(msg), 0, in_page, out_page, 1)])
    _, rfin = O.log_queue_simulate(orc, req)
    reads = synthetic.bytes_to_u256_words(msg, unalignment) if msg else np.zeros((0, 8), dtype=np.uint32)
    return O.keccak_closed_form(rfin), req, reads


@pytest.mark.parametrize("length,unalignment", REFERENCE_CASES)
def test_reference_cases(orc, length, unalignment):
    """test_for_length_and_unalignment, mod.rs:1000-1094: one call, limit = 2, digest written to memory == Keccak256"""
    msg = np.random.default_rng(length * 32 + unalignment).integers(0, 256, length, dtype=np.uint8).tobytes()
    io, req, reads = single_call_instance(orc, msg, unalignment)
    rc, out, trace, com, st, states = O.keccak_entry_point(orc, io, req, reads, 2)
    assert rc == abi.ZKC_OK, hex(st.failed_checks)
    assert out.completion_flag == 1
    row = int(np.flatnonzero(trace[K["WRITE_RESULT"]])[0])
    limbs = trace[K["RESULT"]:K["RESULT"] + 8, row].astype(np.uint32)
    digest = int.from_bytes(limbs.astype("<u4").tobytes(), "little").to_bytes(32, "big")
    assert digest == keccak256(orc, msg)
    assert out.final_memory_state.length == len(reads) + 1 == len(states)


@pytest.mark.parametrize("length,unalignment", [(0, 0), (0, 5), (1, 0), (31, 1), (135, 0), (136, 0), (137, 3), (271, 31), (272, 0), (1023, 17), (2000, 9)])
def test_more_lengths(orc, length, unalignment):
    msg = np.random.default_rng(length).integers(0, 256, length, dtype=np.uint8).tobytes()
    io, req, reads = single_call_instance(orc, msg, unalignment)
    limit = length // 136 + 3
    rc, out, trace, com, st, states = O.keccak_entry_point(orc, io, req, reads, limit)
    assert rc == abi.ZKC_OK, hex(st.failed_checks)
    rows = np.flatnonzero(trace[K["WRITE_RESULT"]])
    assert len(rows) == 1 and rows[0] == length // 136  # one permutation per 136-byte block incl. the padding block
    limbs = trace[K["RESULT"]:K["RESULT"] + 8, rows[0]].astype(np.uint32)
    assert int.from_bytes(limbs.astype("<u4").tobytes(), "little").to_bytes(32, "big") == keccak256(orc, msg)
    assert out.completion_flag == 1 and out.hidden_fsm_output.completed == 1


def test_many_calls_and_chaining(orc):
    reqs, reads, msgs = synthetic.keccak_calls(40, seed=3, max_len=700)
    _, rfin = O.log_queue_simulate(orc, reqs)
    io = O.keccak_closed_form(rfin)
    limit = sum(len(m) // 136 + 1 for m in msgs) + 5
    rc, out, trace, com, st, states = O.keccak_entry_point(orc, io, reqs, reads, limit)
    assert rc == abi.ZKC_OK, hex(st.failed_checks)
    rows = np.flatnonzero(trace[K["WRITE_RESULT"]])
    assert len(rows) == 40
    for r, m in zip(rows, msgs):
        limbs = trace[K["RESULT"]:K["RESULT"] + 8, r].astype(np.uint32)
        assert int.from_bytes(limbs.astype("<u4").tobytes(), "little").to_bytes(32, "big") == keccak256(orc, m)
    assert out.completion_flag == 1 and out.final_memory_state.length == len(reads) + 40
    # split in the middle of a call: chained instances == whole
    cut = int(rows[17]) - 1
    rc, a, ta, _, st, s1 = O.keccak_entry_point(orc, io, reqs, reads, cut)
    assert rc == 0 and a.completion_flag == 0
    nxt = abi.KeccakClosedForm.from_buffer_copy(bytes(a)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.hidden_fsm_output
    used_req = len(reqs) - a.hidden_fsm_output.log_queue_state.length
    used_reads = int(ta[K["QUERY"] + 3::K["QUERY_STRIDE"]][:6].sum())
    rc, b, tb, _, st, s2 = O.keccak_entry_point(orc, nxt, reqs[used_req:], reads[used_reads:], limit - cut)
    assert rc == 0, hex(st.failed_checks)
    assert bytes(b.hidden_fsm_output) == bytes(out.hidden_fsm_output)
    assert np.array_equal(np.concatenate([ta, tb], axis=1), trace)
    assert np.array_equal(np.concatenate([s1, s2]), states)


def test_negative_cases(orc):
    msg = b"hello world" * 20
    io, req, reads = single_call_instance(orc, msg, 3)
    bad = req.copy(); bad["address"][0][0] = 0x8011
    _, rfin = O.log_queue_simulate(orc, bad)
    rc, _, _, _, st, _ = O.keccak_entry_point(orc, O.keccak_closed_form(rfin), bad, reads, 4)
    assert st.failed_checks == abi.KC_CHK["ADDRESS"] and st.first_bad_row == 0
    rc, _, _, _, st, _ = O.keccak_entry_point(orc, io, req, reads[:2], 4)
    assert st.failed_checks & abi.KC_CHK["WITNESS_EXHAUSTED"]
