"""code_unpacker_sha256 oracle: the FSM's SHA-256 over the unpacked words against hashlib (a request is satisfiable exactly
when its versioned hash is the SHA-256 of its code), memory writes, chaining over instances, negative cases.  The reference's
own test (/root/reference/src/code_unpacker_sha256/mod.rs:480-...) builds one request the same way."""
import hashlib

import numpy as np

import orc as O
import vectors as V
from era_zkevm_circuits_b200 import abi, synthetic

K = abi.CU_COLS
CHK = abi.CU_CHK


def instance(orc, reqs):
    prev, fin = O.decommit_queue_simulate(orc, reqs)
    return O.code_unpacker_closed_form(fin, None, True), prev


def test_reference_vector_is_satisfied(orc):
    """the reference's own test: 1 request, 33 words, limit 40; every enforcement holds (the hash check among them), the requests
    queue ends empty and the memory queue equals the queue of the 33 writes (mod.rs:600-615)"""
    reqs, words = V.code_unpacker_reference_vector()
    assert int(reqs["code_hash"][0][7]) == (abi.CODE_HASH_VERSION_TOP16 << 16) | 33
    io, _ = instance(orc, reqs)
    rc, out, trace, com, st, states = O.code_unpacker_entry_point(orc, io, reqs, words, 40)
    assert rc == abi.ZKC_OK and st.failed_checks == 0
    assert out.completion_flag == 1 and out.hidden_fsm_output.decommittment_requests_queue_state.length == 0
    assert trace[K["FINALIZE"]].tolist() == [0] * 16 + [1] + [0] * 23
    mq = np.zeros(33, dtype=abi.MEMORY_QUERY_DTYPE)
    mq["timestamp"], mq["memory_page"], mq["index"], mq["rw_flag"], mq["value"] = 40973, 2368, np.arange(33), 1, words
    prev = np.zeros((33, 12), dtype=np.uint64)
    fin = abi.QueueState12()
    orc.orc_memory_queue_simulate(O.p(mq), 33, O.p(prev), O.C.byref(fin))
    assert list(fin.tail) == list(out.memory_queue_final_state.tail) and out.memory_queue_final_state.length == 33
    # a flipped bit of the code is caught by the hash comparison of the finalizing round
    w2 = words.copy(); w2[20, 2] ^= 1
    rc, _, _, _, st, _ = O.code_unpacker_entry_point(orc, io, reqs, w2, 40)
    assert st.failed_checks == CHK["HASH"] and st.first_bad_row == 16


def test_single_request_sha256_and_memory_writes(orc):
    reqs, words = synthetic.code_decommit_requests(1, seed=1, max_words=9)
    n_words = int(reqs["code_hash"][0][7]) & 0xFFFF
    assert n_words % 2 == 1 and len(words) == n_words
    rounds = (n_words + 1) // 2
    io, _ = instance(orc, reqs)
    rc, out, trace, com, st, states = O.code_unpacker_entry_point(orc, io, reqs, words, rounds + 2)
    assert rc == abi.ZKC_OK, (hex(st.failed_checks), st.first_bad_row)
    assert out.completion_flag == 1 and out.hidden_fsm_output.internal_fsm.finished == 1
    # one memory write per word, indices 0 .. n_words-1 on the request's page at its timestamp
    assert len(states) == n_words == out.memory_queue_final_state.length
    idx = np.concatenate([np.stack([trace[K["INDEX0"]][:rounds], trace[K["INDEX1"]][:rounds]], axis=1).reshape(-1)[:n_words]])
    assert idx.tolist() == list(range(n_words))
    assert set(trace[K["PAGE"]][:rounds].tolist()) == {int(reqs["page"][0])}
    # the state after the finalizing round is the SHA-256 of the code
    code = b"".join(int(sum(int(l) << (32 * k) for k, l in enumerate(w))).to_bytes(32, "big") for w in words)
    digest = hashlib.sha256(code).digest()
    got = b"".join(int(trace[K["STATE_NEW"] + i][rounds - 1]).to_bytes(4, "big") for i in range(8))
    assert got == digest
    assert trace[K["FINALIZE"]].tolist() == [0] * (rounds - 1) + [1, 0, 0]
    # idle afterwards
    assert trace[K["FLAGS_OUT"] + 2][rounds - 1:].tolist() == [1, 1, 1] and trace[K["DECOMMIT"]][rounds:].sum() == 0


def test_many_requests_chain_and_compare(orc):
    reqs, words = synthetic.code_decommit_requests(40, seed=7, max_words=31)
    total_rounds = int((((reqs["code_hash"][:, 7] & 0xFFFF) + 1) // 2).sum())
    io, _ = instance(orc, reqs)
    rc, out, trace, com, st, states = O.code_unpacker_entry_point(orc, io, reqs, words, total_rounds + 5)
    assert rc == abi.ZKC_OK, (hex(st.failed_checks), st.first_bad_row)
    assert out.completion_flag == 1 and len(states) == len(words)
    assert trace[K["FINALIZE"]].sum() == 40 and trace[K["FLAGS_IN"]].sum() == 40
    # chained instances == whole (cut inside a request)
    cut = total_rounds // 2 + 1
    rc, a, ta, _, st, s1 = O.code_unpacker_entry_point(orc, io, reqs, words, cut)
    assert rc == abi.ZKC_OK and a.completion_flag == 0
    popped = len(reqs) - a.hidden_fsm_output.decommittment_requests_queue_state.length
    nxt = abi.CodeUnpackerClosedForm.from_buffer_copy(bytes(a)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.hidden_fsm_output
    rc, b, tb, _, st, s2 = O.code_unpacker_entry_point(orc, nxt, reqs[popped:], words[len(s1):], total_rounds + 5 - cut)
    assert rc == abi.ZKC_OK, (hex(st.failed_checks), st.first_bad_row)
    assert bytes(b.hidden_fsm_output) == bytes(out.hidden_fsm_output) and bytes(b.memory_queue_final_state) == bytes(out.memory_queue_final_state)
    assert np.array_equal(np.concatenate([ta, tb], axis=1), trace) and np.array_equal(np.concatenate([s1, s2]), states)
    exp = abi.CodeUnpackerClosedForm.from_buffer_copy(bytes(b)); exp.start_flag = 0; exp.hidden_fsm_input = a.hidden_fsm_output
    rc, *_ = O.code_unpacker_entry_point(orc, exp, reqs[popped:], words[len(s1):], total_rounds + 5 - cut, compare_expected=True)
    assert rc == abi.ZKC_OK
    exp.hidden_fsm_output.internal_fsm.current_index ^= 1
    rc, *_ = O.code_unpacker_entry_point(orc, exp, reqs[popped:], words[len(s1):], total_rounds + 5 - cut, compare_expected=True)
    assert rc == abi.ZKC_ERR_FSM_OUTPUT_MISMATCH


def test_negative_cases(orc):
    reqs, words = synthetic.code_decommit_requests(6, seed=3, max_words=15)
    rounds = ((reqs["code_hash"][:, 7] & 0xFFFF) + 1) // 2
    limit = int(rounds.sum()) + 2
    # a flipped code bit: the hash of request 2 no longer matches, at its finalizing round
    w2 = words.copy(); w2[int((reqs["code_hash"][:2, 7] & 0xFFFF).sum()) + 1, 3] ^= 4
    io, _ = instance(orc, reqs)
    rc, _, _, _, st, _ = O.code_unpacker_entry_point(orc, io, reqs, w2, limit)
    assert st.failed_checks == CHK["HASH"] and st.first_bad_row == int(rounds[:3].sum()) - 1
    # wrong version byte
    r3 = reqs.copy(); r3["code_hash"][1][7] ^= 1 << 24
    io3, _ = instance(orc, r3)
    rc, _, _, _, st, _ = O.code_unpacker_entry_point(orc, io3, r3, words, limit)
    assert st.failed_checks & CHK["VERSION"] and st.first_bad_row == int(rounds[0])
    # an even number of words
    r4 = reqs.copy(); r4["code_hash"][0][7] += 1
    io4, _ = instance(orc, r4)
    rc, _, _, _, st, _ = O.code_unpacker_entry_point(orc, io4, r4, words, limit)
    assert st.failed_checks & CHK["LENGTH"] and st.first_bad_row == 0
    # the code words run dry
    rc, _, _, _, st, _ = O.code_unpacker_entry_point(orc, io, reqs, words[:-3], limit)
    assert st.failed_checks & CHK["WITNESS_EXHAUSTED"]
    # empty requests queue at the start: the first cycle pops from an empty queue
    e = np.zeros(0, dtype=abi.DECOMMIT_QUERY_DTYPE)
    io0, _ = instance(orc, e)
    rc, out, _, _, st, _ = O.code_unpacker_entry_point(orc, io0, e, words[:0], 4)
    assert st.failed_checks & CHK["WITNESS_EXHAUSTED"] and st.first_bad_row == 0


def test_cycle_relations_of_the_trace(orc):
    """The cycle-to-cycle relations zkc_code_unpacker_check_trace evaluates on the device (cu_check_kernel), restated in numpy and held
    against the oracle's trace: FSM flags, versioned-hash decomposition, selects, round counter, indices, SHA-256 block with the
    padding selected in on finalize, state chaining, the hash comparison, memory-queue bookkeeping.  Pins the evaluator's reading of
    mod.rs:191-447 without a GPU."""
    reqs, words = synthetic.code_decommit_requests(60, seed=4, max_words=41)
    io, _ = instance(orc, reqs)
    limit = int((((reqs["code_hash"][:, 7] & 0xFFFF).astype(np.int64) + 1) // 2).sum()) + 15
    rc, out, T, com, st, states = O.code_unpacker_entry_point(orc, io, reqs, words, limit)
    assert rc == 0
    col = lambda name, i=0: T[K[name] + i]
    prev = lambda a, first: np.concatenate([np.array([first], dtype=np.uint64), a[:-1].astype(np.uint64)])
    get, dc_in, fin_in = col("FLAGS_IN", 0), col("FLAGS_IN", 1), col("FLAGS_IN", 2)
    fo = [col("FLAGS_OUT", i) for i in range(3)]
    assert all(np.array_equal(x[1:], y[:-1]) for x, y in zip((get, dc_in, fin_in), fo)) and (get[0], dc_in[0], fin_in[0]) == (1, 0, 0)
    R = K["REQUEST"]
    h7 = T[R + 7]
    assert np.array_equal(col("VERSION_MATCHES"), (h7 >> np.uint64(16)) == abi.CODE_HASH_VERSION_TOP16)
    liw, lir, bits = col("LENGTH_IN_WORDS"), col("LENGTH_IN_ROUNDS"), col("LENGTH_IN_BITS")
    assert np.array_equal(liw, np.where(get == 1, h7 & 0xFFFF, 1)) and np.array_equal(2 * lir, liw + 1)
    assert np.array_equal(bits, np.where(get == 1, (liw * 256) & 0xFFFFFFFF, prev(bits, 0)))
    assert np.array_equal(col("TIMESTAMP"), np.where(get == 1, T[R + 10], prev(col("TIMESTAMP"), 0)))
    assert np.array_equal(col("PAGE"), np.where(get == 1, T[R + 8], prev(col("PAGE"), 0)))
    for i in range(8):
        assert np.array_equal(col("HASH_TO_COMPARE", i), np.where(get == 1, T[R + i] if i < 7 else 0 * h7, prev(col("HASH_TO_COMPARE", i), 0)))
    dcm, nrl = col("DECOMMIT"), col("NUM_ROUNDS_LEFT")
    sel = np.where(get == 1, lir, prev(nrl, 0))
    assert np.array_equal(dcm, dc_in | get) and np.array_equal(nrl, np.where(dcm == 1, (sel - 1) & 0xFFFF, sel))
    last, fz, second = col("LAST_ROUND"), col("FINALIZE"), col("PROCESS_SECOND_WORD")
    assert np.array_equal(last, nrl == 0) and np.array_equal(fz, last & dcm) and np.array_equal(second, (1 - last) & dcm)
    i0, i1, iout = col("INDEX0"), col("INDEX1"), col("INDEX_OUT")
    assert np.array_equal(i0, np.where(get == 1, 0, prev(iout, 0))) and np.array_equal(i1, i0 + dcm) and np.array_equal(iout, i1 + second)
    assert all((col("WORD0", i)[dcm == 0] == 0).all() and (col("WORD1", i)[second == 0] == 0).all() for i in range(8))
    l0, l1 = T[K["MEM_TAIL0"] + 12], T[K["MEM_TAIL1"] + 12]
    assert np.array_equal(l0, prev(l1, 0) + dcm) and np.array_equal(l1, l0 + second)
    for i in range(12):
        t0, t1 = T[K["MEM_TAIL0"] + i], T[K["MEM_TAIL1"] + i]
        assert np.array_equal(t0[dcm == 0], prev(t1, 0)[dcm == 0]) and np.array_equal(t1[second == 0], t0[second == 0])
    pad = [np.full(limit, 1 << 31, dtype=np.uint64)] + [np.zeros(limit, np.uint64)] * 6 + [bits]
    for i in range(8):
        assert np.array_equal(col("MESSAGE", i), col("WORD0", 7 - i)) and np.array_equal(col("MESSAGE", 8 + i), np.where(fz == 1, pad[i], col("WORD1", 7 - i)))
    IV = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
    for i in range(8):
        si, sn, so = col("STATE_IN", i), col("STATE_NEW", i), col("STATE_OUT", i)
        assert np.array_equal(si, np.where(get == 1, IV[i], prev(so, 0))) and np.array_equal(so, np.where(dcm == 1, sn, si))
        digest_limb = col("STATE_NEW", 7 - i) if i < 7 else 0 * si
        assert np.array_equal(digest_limb[fz == 1], col("HASH_TO_COMPARE", i)[fz == 1])
    ln = col("REQ_LEN")
    empty = ln == 0
    assert np.array_equal(ln + get, prev(ln, io.sorted_requests_queue_initial_state.length))
    assert np.array_equal(fo[0], fz & ~empty) and np.array_equal(fo[1], second) and np.array_equal(fo[2], fin_in | (fz & empty))
