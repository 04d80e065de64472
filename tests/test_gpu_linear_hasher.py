"""linear_hasher: CUDA path through the C ABI vs the CPU oracle, bit-exact (trace, observable output, commitment, status),
with the out-of-circuit hasher's keccak states as hints (row-parallel) and without (sequential chain on the device)."""
import numpy as np
import pytest

import orc as O
from era_zkevm_circuits_b200 import LinearHasherCircuitInstanceWitness as Witness, abi, linear_hasher_entry_point as entry_point
from test_oracle_linear_hasher import digest_of, instance, keccak256, messages, serialise

pytestmark = pytest.mark.gpu
CHK = abi.LH_CHK


def assert_same(want, got, check_trace=True):
    rc, io, trace, com, st, states = want
    assert got.status.code == rc, (got.status.code, hex(got.status.failed_checks), got.status.first_bad_row, rc, hex(st.failed_checks))
    assert got.status.failed_checks == st.failed_checks and got.status.first_bad_row == st.first_bad_row
    assert got.closed_form_input.completion_flag == io.completion_flag
    assert digest_of(got.closed_form_input) == digest_of(io)
    assert got.commitment.tolist() == com.tolist()
    if check_trace:
        bad = np.argwhere(got.trace != trace)
        assert bad.size == 0, f"first differing (col,row): {bad[:5].tolist()}"


@pytest.mark.parametrize("n,limit", [(0, 0), (0, 3), (1, 1), (2, 5), (17, 17), (255, 256), (1000, 1024), (20000, 20000)])
def test_bit_exact_with_and_without_state_hints(engine, orc, n, limit):
    recs = messages(n, seed=n + 1)
    io, prev = instance(orc, recs)
    want = O.linear_hasher_entry_point(orc, io, recs, limit)
    assert want[0] == abi.ZKC_OK, (hex(want[4].failed_checks), want[4].first_bad_row)
    got = entry_point(engine, Witness(io, recs, prev), limit, raise_on_unsatisfied=False)  # no hints: the device rebuilds the chain
    assert_same(want, got)
    got = entry_point(engine, Witness(io, recs, prev, want[5]), limit, raise_on_unsatisfied=False)
    assert_same(want, got)
    assert digest_of(got.closed_form_input) == keccak256(orc, b"".join(serialise(r) for r in recs))


def test_device_resident_and_wrong_hints(engine, orc):
    import torch
    recs = messages(5000, seed=3)
    io, prev = instance(orc, recs)
    limit = 5120
    want = O.linear_hasher_entry_point(orc, io, recs, limit)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)).cuda()
    got = entry_point(engine, Witness(io, d(recs), d(prev), d(want[5])), limit)
    assert got.commitment.tolist() == want[3].tolist() and np.array_equal(got.trace.cpu().numpy().view(np.uint64), want[2])
    states = want[5].copy(); states[1234, 7] ^= 1
    bad = entry_point(engine, Witness(io, recs, prev, states), limit, raise_on_unsatisfied=False)
    assert bad.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT and bad.status.failed_checks & CHK["STATE_HINT"] and bad.status.first_bad_row == 1234
    p2 = prev.copy(); p2[77, 1] ^= 1
    bad = entry_point(engine, Witness(io, recs, p2, want[5]), limit, raise_on_unsatisfied=False)
    assert bad.status.code == abi.ZKC_ERR_QUEUE_WITNESS_INCONSISTENT and bad.status.failed_checks & CHK["QUEUE_HINT"]


def test_enforcements_match_oracle(engine, orc):
    recs = messages(40, seed=9)
    io, prev = instance(orc, recs)
    for io_k, recs_k, limit in [(io, recs, 39), (O.linear_hasher_closed_form(io.queue_state, start=False), recs, 40)]:
        want = O.linear_hasher_entry_point(orc, io_k, recs_k, limit)
        got = entry_point(engine, Witness(io_k, recs_k, prev), limit, raise_on_unsatisfied=False)
        assert want[0] == abi.ZKC_ERR_UNSATISFIED
        assert_same(want, got)
    bad = recs.copy(); bad["tx_number_in_block"][5] = 0x10000
    io3, prev3 = instance(orc, bad)
    want = O.linear_hasher_entry_point(orc, io3, bad, 40)
    got = entry_point(engine, Witness(io3, bad, prev3), 40, raise_on_unsatisfied=False)
    assert_same(want, got)


def test_check_trace_constraint_evaluation(engine, orc):
    """zkc_linear_hasher_check_trace: the ORACLE's trace satisfies every relation with and without the pop's permutations; a fault
    injected into any relation family is found at its cycle (a byte of the serialisation also breaks the sponge of the cycle that
    absorbs it)"""
    from era_zkevm_circuits_b200 import linear_hasher_check_trace
    K = abi.LH_COLS
    V_ = abi.LHV
    recs = messages(700, seed=8)
    io, prev = instance(orc, recs)
    limit = 730
    want = O.linear_hasher_entry_point(orc, io, recs, limit)
    assert want[0] == abi.ZKC_OK
    trace = want[2]
    for gates in (0, abi.GATES_GENERAL):
        viol, st = linear_hasher_check_trace(engine, io, trace, limit, gates)
        assert viol == 0 and st.code == 0, (gates, viol, hex(st.failed_checks), st.first_bad_row)
    import torch
    viol, st = linear_hasher_check_trace(engine, io, torch.from_numpy(trace.view(np.int64)).cuda(), limit, abi.GATES_GENERAL)
    assert viol == 0
    full = np.flatnonzero(trace[K["ABSORB_FULL"]])
    faults = [
        (K["SHOULD_POP"], 17, 2, V_["BOOLEAN"], 0),
        (K["ITEM"] + 7, 40, 1 << 33, V_["BOOLEAN"], 0),
        (K["ITEM"] + 34, 41, 1 << 16, V_["ENFORCE"], 0),
        (K["ENC"] + 3, 99, None, V_["ENCODING"], 0),
        (K["BYTES"] + 50, 100, None, V_["ENCODING"], 0),
        (K["LEN"], 123, None, V_["QUEUE"], 0),
        (K["HEAD"] + 1, 715, None, V_["QUEUE"], 0),
        (K["HEAD"] + 1, 150, None, V_["ROUND_FUNCTION"], 0),
        (K["NOW_EMPTY"], 200, None, V_["FLAGS"], 0),
        (K["IS_LAST_SERIALIZATION"], 699, None, V_["FLAGS"], 0),
        (K["CONTINUE_TO_ABSORB"], 710, None, V_["FLAGS"], 0),
        (K["ABSORB_FULL"], int(full[30]), None, V_["FLAGS"], 0),
        (K["DONE"], 300, None, V_["FLAGS"], 0),
        (K["STATE_MID"] + 13, int(full[40]), None, V_["SPONGE"], 0),
        (K["STATE_OUT"] + 27, 400, None, V_["SPONGE"], 0),
    ]
    for col, row, val, bit, gates in faults:
        bad = trace.copy()
        bad[col, row] = np.uint64(val) if val is not None else bad[col, row] ^ np.uint64(1)
        viol, st = linear_hasher_check_trace(engine, io, bad, limit, gates)
        assert viol >= 1 and st.first_bad_row == row and st.failed_checks & bit, (col, row, viol, st.first_bad_row, hex(st.failed_checks))
    # the engine's own trace (device resident, state hints from the oracle)
    i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)).cuda()
    got = entry_point(engine, Witness(io, dev(recs), i64(prev), i64(want[5])), limit)
    viol, st = linear_hasher_check_trace(engine, io, got.trace, limit)
    assert viol == 0, (viol, hex(st.failed_checks), st.first_bad_row)
