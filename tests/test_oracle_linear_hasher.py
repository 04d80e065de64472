"""linear_hasher oracle (/root/reference/src/linear_hasher/mod.rs:35-214).  The reference holds no vector for this circuit;
the oracle is pinned on what the circuit is FOR: its observable output must be Keccak-256 of the concatenated 88-byte
L2 -> L1 message serialisations (log_query/mod.rs:645-686), checked here against a serialisation written independently in
Python and the Keccak-256 the keccak256 precompile tests pin on the standard KATs (orc_keccak256)."""
import numpy as np
import pytest

import orc as O
from era_zkevm_circuits_b200 import abi, synthetic

K = abi.LH_COLS
CHK = abi.LH_CHK


def messages(n, seed):
    recs = synthetic.vm_log_queue_trace(n, seed=seed)
    recs["tx_number_in_block"] &= 0xFFFF  # u16 on the wire (log_query/mod.rs:661-668)
    return recs


def serialise(r) -> bytes:
    """shard_id, is_service, tx_number (2 bytes BE), address (20 BE), key (32 BE), written_value (32 BE)"""
    fl = int(r["flags"])
    be = lambda limbs: int.from_bytes(np.asarray(limbs, dtype="<u4").tobytes(), "little").to_bytes(4 * len(limbs), "big")
    out = bytes([(fl >> 8) & 0xFF, (fl >> 18) & 1]) + int(r["tx_number_in_block"]).to_bytes(2, "big") + be(r["address"]) + be(r["key"]) + be(r["written_value"])
    assert len(out) == abi.LH_MESSAGE_BYTES
    return out


def keccak256(orc, msg: bytes) -> bytes:
    out = np.zeros(32, dtype=np.uint8)
    buf = np.frombuffer(msg, dtype=np.uint8).copy() if msg else np.zeros(1, dtype=np.uint8)
    orc.orc_keccak256(O.p(buf), len(msg), O.p(out))
    return out.tobytes()


def instance(orc, recs):
    prev, fin = O.log_queue_simulate(orc, recs)
    return O.linear_hasher_closed_form(fin), prev


def digest_of(io):
    return bytes(int(io.keccak256_hash[i]) for i in range(32))


@pytest.mark.parametrize("n,limit", [(0, 0), (0, 4), (1, 1), (1, 3), (2, 2), (3, 8), (17, 17), (34, 40), (200, 256), (1000, 1000)])
def test_digest_is_keccak256_of_the_serialised_queue(orc, n, limit):
    recs = messages(n, seed=100 + n)
    io, _ = instance(orc, recs)
    rc, out, trace, com, st, states = O.linear_hasher_entry_point(orc, io, recs, limit)
    assert rc == abi.ZKC_OK, (hex(st.failed_checks), st.first_bad_row)
    assert out.completion_flag == 1
    stream = b"".join(serialise(r) for r in recs)
    assert digest_of(out) == keccak256(orc, stream)
    if n == 0:
        assert digest_of(out).hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    if limit:
        assert trace[K["BYTES"]:K["BYTES"] + 88, :n].T.astype(np.uint8).tobytes() == stream
        # the absorption schedule: one full block per 136 bytes of the stream, one padded block at the last item
        assert int(trace[K["ABSORB_FULL"]].sum()) == len(stream) // 136 and int(trace[K["ABSORB_LAST"]].sum()) == (1 if n else 0)
        assert trace[K["DONE"]].tolist() == [int(c + 1 >= n) for c in range(limit)]


def test_enforcements(orc):
    recs = messages(20, seed=7)
    io, _ = instance(orc, recs)
    rc, out, _, _, st, _ = O.linear_hasher_entry_point(orc, io, recs, 19)  # the single instance must finish the queue (:176)
    assert rc == abi.ZKC_ERR_UNSATISFIED and st.failed_checks == CHK["NOT_COMPLETED"] and out.completion_flag == 0
    io2 = O.linear_hasher_closed_form(io.queue_state, start=False)           # :66
    rc, _, _, _, st, _ = O.linear_hasher_entry_point(orc, io2, recs, 20)
    assert rc == abi.ZKC_ERR_UNSATISFIED and st.failed_checks == CHK["START_FLAG"]
    bad = recs.copy(); bad["tx_number_in_block"][5] = 0x10000                # into_bytes truncates to two bytes and enforces the rest zero
    io3, _ = instance(orc, bad)
    rc, _, _, _, st, _ = O.linear_hasher_entry_point(orc, io3, bad, 20)
    assert rc == abi.ZKC_ERR_UNSATISFIED and st.failed_checks == CHK["TX_NUMBER_RANGE"] and st.first_bad_row == 5
    io4 = O.linear_hasher_closed_form(io.queue_state); io4.queue_state.tail[0] ^= 1  # the witness is not the committed queue (:173)
    rc, _, _, _, st, _ = O.linear_hasher_entry_point(orc, io4, recs, 20)
    assert rc == abi.ZKC_ERR_UNSATISFIED and st.failed_checks == CHK["QUEUE_CONSISTENCY"]
    exp = O.linear_hasher_closed_form(io.queue_state)
    rc, good, _, _, _, _ = O.linear_hasher_entry_point(orc, io, recs, 20)
    exp.completion_flag = 1
    for i in range(32):
        exp.keccak256_hash[i] = good.keccak256_hash[i]
    assert O.linear_hasher_entry_point(orc, exp, recs, 20, compare_expected=True)[0] == abi.ZKC_OK
    exp.keccak256_hash[3] ^= 1
    assert O.linear_hasher_entry_point(orc, exp, recs, 20, compare_expected=True)[0] == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH


def test_cycle_relations_of_the_trace(orc):
    """The cycle-to-cycle flag relations zkc_linear_hasher_check_trace evaluates on the device (lh_check_kernel), restated in numpy and
    held against the oracle's trace (the sponge itself is pinned by the digest tests above): queue bookkeeping, now_empty /
    is_last_serialization / done / continue_to_absorb, the absorption conditions as functions of the cycle index, state carry-over."""
    K = abi.LH_COLS
    recs = messages(500, seed=5)
    io, _ = instance(orc, recs)
    limit = 520
    rc, out, T, com, st, states = O.linear_hasher_entry_point(orc, io, recs, limit)
    assert rc == 0
    col = lambda name, i=0: T[K[name] + i]
    prev = lambda a, first: np.concatenate([np.array([first], dtype=np.uint64), a[:-1].astype(np.uint64)])
    emp, pop, ln, len0 = col("QUEUE_IS_EMPTY"), col("SHOULD_POP"), col("LEN"), io.queue_state.length
    assert np.array_equal(ln + pop, prev(ln, len0)) and np.array_equal(emp, prev(ln, len0) == 0) and np.array_equal(pop, 1 - emp)
    now, last, done, cont = col("NOW_EMPTY"), col("IS_LAST_SERIALIZATION"), col("DONE"), col("CONTINUE_TO_ABSORB")
    done_prev = prev(done, int(len0 == 0))
    assert np.array_equal(now, ln == 0) and np.array_equal(last, pop & now) and np.array_equal(done, done_prev | last) and np.array_equal(cont, 1 - done_prev)
    lb = (np.arange(limit) * abi.LH_MESSAGE_BYTES) % 136
    full, final = col("ABSORB_FULL"), col("ABSORB_LAST")
    assert np.array_equal(full, (lb + abi.LH_MESSAGE_BYTES >= 136) & (cont == 1)) and np.array_equal(final, cont & last)
    for i in range(50):
        mid, so = col("STATE_MID", i), col("STATE_OUT", i)
        assert np.array_equal(mid[full == 0], prev(so, 0)[full == 0]) and np.array_equal(so[final == 0], mid[final == 0])
    assert (T[K["BYTES"]:K["BYTES"] + abi.LH_MESSAGE_BYTES] < 256).all()
