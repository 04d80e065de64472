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


def test_cycle_relations_of_the_trace(orc):
    """The cycle-to-cycle relations zkc_sha256_round_function_check_trace evaluates on the device (sh_check_kernel), restated in numpy
    and held against the oracle's trace: FSM flags, parameter / timestamp selects, offsets, round counter, state chaining, message
    words, result word, memory-queue bookkeeping.  Pins the evaluator's reading of mod.rs:146-330 without a GPU."""
    reqs, reads, msgs = synthetic.sha256_calls(200, seed=3, max_rounds=9)
    _, rfin = O.log_queue_simulate(orc, reqs)
    io = O.sha256_closed_form(rfin)
    limit = len(reads) // 2 + 20
    rc, out, T, com, st, states = O.sha256_entry_point(orc, io, reqs, reads, limit)
    assert rc == 0
    col = lambda name, i=0: T[K[name] + i]
    prev = lambda a, first: np.concatenate([np.array([first], dtype=np.uint64), a[:-1].astype(np.uint64)])
    rpc, rwr_in, comp = col("FLAGS_IN", 0), col("FLAGS_IN", 1), col("FLAGS_IN", 2)
    fo = [col("FLAGS_OUT", i) for i in range(3)]
    assert all(np.array_equal(x[1:], y[:-1]) for x, y in zip((rpc, rwr_in, comp), fo)) and (rpc[0], rwr_in[0], comp[0]) == (1, 0, 0)
    I = K["CALL_ITEM"]
    key, call_ts = [T[I + 5 + i] for i in range(8)], T[I + 35]
    P = [col("PARAMS", i) for i in range(5)]
    Q0, Q1 = K["QUERY"], K["QUERY"] + K["QUERY_STRIDE"]
    carried = [prev(P[0], 0), prev(T[Q1 + 21], 0), prev(P[2], 0), prev(P[3], 0), prev(col("NUM_ROUNDS"), 0)]
    from_call = [key[4], key[0], key[5], key[2], key[6]]
    assert all(np.array_equal(P[i], np.where(rpc == 1, from_call[i], carried[i])) for i in range(5))
    tsr, tsw = col("TS_READ"), col("TS_WRITE")
    assert np.array_equal(tsr, np.where(rpc == 1, call_ts, prev(tsr, 0))) and np.array_equal(tsw, np.where(rpc == 1, (tsr + 1) & 0xFFFFFFFF, prev(tsw, 0)))
    reset, should_read, rwr = col("RESET_BUFFER"), col("SHOULD_READ"), rpc | rwr_in
    assert np.array_equal(reset, rpc | comp) and np.array_equal(should_read, P[4] != 0)
    assert np.array_equal(T[Q0 + 21], (P[1] + rwr) & 0xFFFFFFFF) and np.array_equal(T[Q1 + 21], (T[Q0 + 21] + rwr) & 0xFFFFFFFF)
    rounds = col("NUM_ROUNDS")
    write = col("WRITE_RESULT")
    assert np.array_equal(rounds, (P[4] - rwr) & 0xFFFFFFFF) and np.array_equal(write, rwr & (rounds == 0))
    ln = col("REQ_LEN")
    empty = ln == 0
    assert np.array_equal(ln + rpc, prev(ln, rfin.length))
    assert np.array_equal(fo[0], write & ~empty) and np.array_equal(fo[2], (write & empty) | comp) and np.array_equal(fo[1], 1 - (fo[0] | fo[2]))
    IV = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
    assert all(np.array_equal(col("STATE_IN", i), np.where(reset == 1, IV[i], prev(col("STATE_OUT", i), 0))) for i in range(8))
    assert all(np.array_equal(col("RESULT", 7 - k), col("STATE_OUT", k)) for k in range(8))
    assert all(np.array_equal(col("MESSAGE", 8 * q + i), T[K["QUERY"] + K["QUERY_STRIDE"] * q + 7 - i]) for q in range(2) for i in range(8))
    l0, l1, lw = T[Q0 + 20], T[Q1 + 20], col("WRITE_LEN")
    assert np.array_equal(l0, prev(lw, 0) + should_read) and np.array_equal(l1, l0 + should_read) and np.array_equal(lw, l1 + write)
    for i in range(12):
        t0, t1, tw = T[Q0 + 8 + i], T[Q1 + 8 + i], col("WRITE_TAIL", i)
        idle = should_read == 0
        assert np.array_equal(t0[idle], prev(tw, 0)[idle]) and np.array_equal(t1[idle], t0[idle]) and np.array_equal(tw[write == 0], t1[write == 0])
    assert all((T[K["QUERY"] + K["QUERY_STRIDE"] * q + i][should_read == 0] == 0).all() for q in range(2) for i in range(8))
