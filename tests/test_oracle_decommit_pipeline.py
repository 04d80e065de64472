"""The decommitment pipeline end to end on the oracle: the VM's decommitment queue (requests in execution order, repeated
hashes) -> sort_decommittment_requests (deduplicated queue) -> code_unpacker_sha256 (code into memory, hash checked).  What the
reference's scheduler checks between these circuits (`scheduler/mod.rs:79-93`: the output queue state of one circuit is the
input queue state of the next) is asserted here on the queue states the oracles pass along."""
import numpy as np

import orc as O
from era_zkevm_circuits_b200 import abi, synthetic


def test_sorter_output_queue_is_the_unpacker_input_queue(orc):
    # 12 distinct bytecodes, 60 requests over them in execution order
    uniq, words = synthetic.code_decommit_requests(12, seed=21, max_words=17)
    n_words = (uniq["code_hash"][:, 7] & 0xFFFF).astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(n_words)])
    rng = np.random.default_rng(5)
    which = np.concatenate([np.arange(12), rng.integers(0, 12, size=48)])
    rng.shuffle(which)
    u = np.zeros(60, dtype=abi.DECOMMIT_QUERY_DTYPE)
    first = {}
    for i, k in enumerate(which):
        first.setdefault(int(k), i)
        u[i]["code_hash"] = uniq["code_hash"][k]
        u[i]["timestamp"] = 5000 + 8 * i
        u[i]["is_first"] = int(first[int(k)] == i)
        u[i]["page"] = 4096 + 32 * first[int(k)]
    s = u[np.lexsort([u["timestamp"]] + [u["code_hash"][:, i] for i in range(8)])]
    # circuit 1: sort + deduplicate
    up, ufin = O.decommit_queue_simulate(orc, u)
    sp, sfin = O.decommit_queue_simulate(orc, s)
    rc, out, trace, com, st, states = O.sort_decommittments_entry_point(orc, O.decommit_sorter_closed_form(ufin, sfin, True), u, s, 64)
    assert rc == abi.ZKC_OK and out.completion_flag == 1 and out.final_queue_state.length == 12
    # the deduplicated records, as the out-of-circuit sorter hands them on: first request of every hash, in hash order
    K = abi.DQ_COLS
    rows = np.flatnonzero(trace[K["ADD_TO_QUEUE"]])
    dedup = np.zeros(12, dtype=abi.DECOMMIT_QUERY_DTYPE)
    for j, r in enumerate(rows):
        item = trace[K["PUSH_ITEM"]:K["PUSH_ITEM"] + 11, r]
        dedup[j]["code_hash"], dedup[j]["page"], dedup[j]["is_first"], dedup[j]["timestamp"] = item[:8], item[8], item[9], item[10]
    # limit (64) > queue length (60): the last record is flushed by the first trivial row of the loop, not by the
    # finalisation step (mod.rs:364-375)
    assert len(rows) == 12
    assert (dedup["is_first"] == 1).all() and len({bytes(d["code_hash"]) for d in dedup}) == 12
    assert sorted(dedup["timestamp"].tolist()) == sorted(5000 + 8 * first[k] for k in range(12))
    prev, fin = O.decommit_queue_simulate(orc, dedup)
    assert list(fin.tail) == list(out.final_queue_state.tail) and fin.length == 12
    # circuit 2: unpack; its input queue state IS circuit 1's output queue state
    order = [int(np.flatnonzero((uniq["code_hash"] == d["code_hash"]).all(axis=1))[0]) for d in dedup]
    code = np.concatenate([words[offs[k]:offs[k + 1]] for k in order])
    io = O.code_unpacker_closed_form(out.final_queue_state, None, True)
    limit = int(((n_words + 1) // 2).sum()) + 4
    rc, cu, tr, _, st, mem = O.code_unpacker_entry_point(orc, io, dedup, code, limit)
    assert rc == abi.ZKC_OK, (hex(st.failed_checks), st.first_bad_row)
    assert cu.completion_flag == 1 and cu.memory_queue_final_state.length == int(n_words.sum()) == len(mem)
    # every bytecode landed on the page its first request named
    C = abi.CU_COLS
    pages = sorted(set(tr[C["PAGE"]][tr[C["DECOMMIT"]] == 1].tolist()))
    assert pages == sorted(4096 + 32 * first[k] for k in range(12))
