"""Small valid ORACLE traces of every circuit (test infrastructure): name -> column-major trace [n_cols, rows], for the tests of
the generic allocation-check evaluator (era_zkevm_circuits_b200/column_classes.py)."""
import numpy as np

import helpers as H
import orc as O
from era_zkevm_circuits_b200 import abi, isa as I, synthetic


def oracle_traces(orc):
    out = {}
    u, s = synthetic.ram_trace(600, seed=2, n_cells=40, n_nondet=3)
    io, _, _ = H.ram_instance(orc, u, s, 3)
    r = O.ram_entry_point(orc, io, u, s, 640); assert r[0] == 0
    out["ram_permutation"] = r[2]
    u, s = synthetic.events_trace(500, seed=3, rollback_pct=20)
    _, ufin = O.log_queue_simulate(orc, u); _, sfin = O.log_queue_simulate(orc, s)
    r = O.log_sorter_entry_point(orc, O.events_closed_form(ufin, sfin, True), u, s, 520); assert r[0] == 0
    out["log_sorter"] = r[2]
    u, s, ts = synthetic.storage_trace(800, seed=7, n_cells=60)
    _, ufin = O.log_queue_simulate(orc, u); _, sfin = O.log_queue_simulate(orc, s, ts)
    r = O.storage_validity_entry_point(orc, O.storage_closed_form(ufin, sfin, 0, True), u, s, ts, 820); assert r[0] == 0
    out["storage_validity"] = r[2]
    u, s = synthetic.decommit_requests_trace(400, seed=5, n_hashes=30)
    _, ufin = O.decommit_queue_simulate(orc, u); _, sfin = O.decommit_queue_simulate(orc, s)
    r = O.sort_decommittments_entry_point(orc, O.decommit_sorter_closed_form(ufin, sfin, True), u, s, 420); assert r[0] == 0
    out["sort_decommittment_requests"] = r[2]
    recs = synthetic.vm_log_queue_trace(500, seed=5)
    _, rfin = O.log_queue_simulate(orc, recs)
    r = O.demux_entry_point(orc, O.demux_closed_form(rfin, True), recs, 520); assert r[0] == 0
    out["demux_log_queue"] = r[2]
    reqs, reads, _ = synthetic.keccak_calls(20, seed=3, max_len=500)
    _, rfin = O.log_queue_simulate(orc, reqs)
    r = O.keccak_entry_point(orc, O.keccak_closed_form(rfin), reqs, reads, 120); assert r[0] == 0
    out["keccak256_round_function"] = r[2]
    reqs, reads, _ = synthetic.sha256_calls(20, seed=4, max_rounds=9)
    _, rfin = O.log_queue_simulate(orc, reqs)
    r = O.sha256_entry_point(orc, O.sha256_closed_form(rfin), reqs, reads, len(reads) // 2 + 4); assert r[0] == 0
    out["sha256_round_function"] = r[2]
    creq, cwords = synthetic.code_decommit_requests(12, seed=9, max_words=33)
    _, cfin = O.decommit_queue_simulate(orc, creq)
    climit = int(((((creq["code_hash"][:, 7] & 0xFFFF).astype(np.int64)) + 1) // 2).sum()) + 4
    r = O.code_unpacker_entry_point(orc, O.code_unpacker_closed_form(cfin, None, True), creq, cwords, climit); assert r[0] == 0
    out["code_unpacker_sha256"] = r[2]
    recs = synthetic.vm_log_queue_trace(300, seed=6); recs["tx_number_in_block"] &= 0xFFFF
    _, lfin = O.log_queue_simulate(orc, recs)
    r = O.linear_hasher_entry_point(orc, O.linear_hasher_closed_form(lfin), recs, 320); assert r[0] == 0
    out["linear_hasher"] = r[2]
    isa = I.Isa()
    vio = abi.VmClosedForm(); vio.start_flag = 1
    st0 = O.vm_initial_state(orc, vio, isa.isa)
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st0, I.pack_code(I.random_program(isa, 512, seed=7)), 2000, full=True)
    for k in range(4):
        vio.rollback_queue_tail_for_block[k] = int(tail[k])
    r = O.vm_entry_point(orc, vio, isa.isa, snaps, wit, 2000, cw=cw); assert r[0] == 0
    out["main_vm_gadget_cells"] = O.vm_gadget_cells(orc, r[2], 2000)
    out["main_vm_state_gadget_cells"] = O.vm_state_gadget_cells(orc, r[2], snaps, 2000)
    out["main_vm_memory_sponge_cells"] = O.vm_memory_sponge_cells(orc, r[2], snaps, 2000)
    out["main_vm_prestate_cells"] = O.vm_prestate_cells(orc, r[2], snaps, 2000)
    out["main_vm_writeback_cells"] = O.vm_writeback_cells(orc, isa.isa, r[2], snaps, 2000)
    return out
