"""sha256 precompile oracle.  The reference has no test for this circuit; pinned against
hashlib.sha256: the U256 written to memory, read as a big-endian 32-byte word, is the SHA-256 digest of the
(pre-padded) message the calls feed block by block."""
import hashlib

import numpy as np
import pytest

import orc as O
from era_zkevm_circuits_b200 import abi, synthetic

K = abi.SH_COLS


def digest_of_row(trace, row):
    limbs = trace[K["RESULT"]:K["RESULT"] + 8, row].astype(np.uint32)
    return int.from_bytes(limbs.astype("<u4").tobytes(), "little").to_bytes(32, "big")


def test_compression_against_hashlib(orc):
    for msg in [b"", b"abc", b"a" * 55, b"a" * 56, b"a" * 64, bytes(range(200))]:
        st = np.array([0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19], dtype=np.uint32)
        padded = synthetic.sha256_pad(msg)
        for off in range(0, len(padded), 64):
            blk = np.frombuffer(padded[off:off + 64], dtype=">u4").astype(np.uint32)
            orc.orc_sha256_compress(O.p(st), O.p(blk))
        assert st.astype(">u4").tobytes() == hashlib.sha256(msg).digest()


@pytest.mark.parametrize("n_calls,max_rounds", [(1, 1), (1, 5), (30, 16)])
def test_digests_and_chaining(orc, n_calls, max_rounds):
    reqs, reads, msgs = synthetic.sha256_calls(n_calls, seed=n_calls + max_rounds, max_rounds=max_rounds)
    _, rfin = O.log_queue_simulate(orc, reqs)
    io = O.sha256_closed_form(rfin)
    total = len(reads) // 2
    rc, out, trace, com, st, states = O.sha256_entry_point(orc, io, reqs, reads, total + 3)
    assert rc == abi.ZKC_OK, hex(st.failed_checks)
    rows = np.flatnonzero(trace[K["WRITE_RESULT"]])
    assert len(rows) == n_calls
    for r, m in zip(rows, msgs):
        assert digest_of_row(trace, r) == hashlib.sha256(m).digest()
    assert out.completion_flag == 1 and out.final_memory_state.length == len(reads) + n_calls == len(states)
    if total > 3:
        cut = total // 2
        rc, a, ta, _, st, s1 = O.sha256_entry_point(orc, io, reqs, reads, cut)
        nxt = abi.Sha256ClosedForm.from_buffer_copy(bytes(a)); nxt.start_flag = 0
        nxt.hidden_fsm_input = a.hidden_fsm_output
        used_req = len(reqs) - a.hidden_fsm_output.log_queue_state.length
        rc, b, tb, _, st, s2 = O.sha256_entry_point(orc, nxt, reqs[used_req:], reads[2 * cut:], total + 3 - cut)
        assert rc == 0, hex(st.failed_checks)
        assert bytes(b.hidden_fsm_output) == bytes(out.hidden_fsm_output)
        assert np.array_equal(np.concatenate([ta, tb], axis=1), trace)


def test_negative(orc):
    reqs, reads, msgs = synthetic.sha256_calls(3, seed=1, max_rounds=2)
    bad = reqs.copy(); bad["address"][1][0] = 0x8010
    _, rfin = O.log_queue_simulate(orc, bad)
    rc, _, _, _, st, _ = O.sha256_entry_point(orc, O.sha256_closed_form(rfin), bad, reads, 10)
    assert st.failed_checks == abi.KC_CHK["ADDRESS"]
    e = np.zeros(0, dtype=abi.LOG_QUERY_DTYPE)
    _, rfin = O.log_queue_simulate(orc, e)
    rc, out, _, _, st, _ = O.sha256_entry_point(orc, O.sha256_closed_form(rfin), e, np.zeros((0, 8), dtype=np.uint32), 4)
    assert rc == 0 and out.completion_flag == 1
