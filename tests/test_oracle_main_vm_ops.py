"""main_vm oracle: uma / near_call / ret / log (value-level restatement of main_vm/opcodes/{uma,log,call_ret}.rs and
call_ret_impl/{near_call,ret}.rs).  PARITY UNPINNED against the reference (no main_vm test in the reference, ISA
tables un-vendored): the semantics are pinned here against independent Python models -- a bytearray heap, Python
frames, and the log-queue hash chain of the sorter circuits' oracle."""
import numpy as np

import orc as O
from era_zkevm_circuits_b200 import abi, isa as I

K = abi.VM_COLS
M256 = (1 << 256) - 1


def reg(state, r):
    return sum(int(v) << (32 * i) for i, v in enumerate(state.registers[r - 1].value))


def set_reg(st, r, v, is_ptr=0):
    st.registers[r - 1].is_pointer = is_ptr
    for i in range(8):
        st.registers[r - 1].value[i] = (v >> (32 * i)) & 0xFFFFFFFF


def fresh(orc, tail=0):
    isa = I.Isa()
    io = abi.VmClosedForm()
    io.start_flag = 1
    io.rollback_queue_tail_for_block[0] = tail
    return isa, io, O.vm_initial_state(orc, io, isa.isa)


def run_full(orc, isa, io, st, ops, cycles):
    """out-of-circuit run + the circuit over its witness; returns everything"""
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
    assert rc == 0, (rc, hex(status.failed_checks), status.first_bad_row)
    io2 = abi.VmClosedForm.from_buffer_copy(bytes(io))
    for i in range(4):
        io2.rollback_queue_tail_for_block[i] = int(tail[i])
    # registers may have been preset: run the circuit as a continuation from snapshot 0
    io2.start_flag = 0
    io2.hidden_fsm_input = O.vm_state_at(snaps, 0)
    res = O.vm_entry_point(orc, io2, isa.isa, snaps, wit, cycles, cw=cw)
    return snaps, wit, cw, tail, res


def test_uma_heap_against_bytearray_model(orc):
    isa, io, st = fresh(orc)
    heap, aux = bytearray(1 << 16), bytearray(1 << 16)
    vals = [int.from_bytes(np.random.default_rng(i).bytes(32), "big") for i in range(6)]
    plan = [  # (variant, increment, offset, value register)
        (I.UMA_HEAP_WRITE, 0, 64, 0), (I.UMA_HEAP_READ, 0, 64, None), (I.UMA_HEAP_WRITE, 1, 45, 1), (I.UMA_HEAP_READ, 1, 40, None),
        (I.UMA_HEAP_READ, 0, 70, None), (I.UMA_AUX_WRITE, 0, 7, 2), (I.UMA_AUX_READ, 0, 0, None), (I.UMA_AUX_READ, 1, 31, None),
        (I.UMA_HEAP_READ, 0, 0, None), (I.UMA_HEAP_WRITE, 0, 96, 3), (I.UMA_HEAP_READ, 0, 95, None),
    ]
    ops, expect = [], []
    for i, v in enumerate(vals[:4]):
        set_reg(st, 10 + i, v)
    for k, (variant, inc, off, vr) in enumerate(plan):
        # offset via an immediate add into r2, then the access with src0 = r2 (src1 = value), dst0 = r3, dst1 = r4
        ops.append(isa.encode(I.OP_ADD, 0, 0, src=I.MODE_IMM16, src1=0, dst0=2, imm0=off))
        ops.append(isa.encode(I.OP_UMA, variant, inc, src0=2, src1=(10 + vr) if vr is not None else 0, dst0=3, dst1=4))
        mem = heap if variant in (I.UMA_HEAP_READ, I.UMA_HEAP_WRITE) else aux
        if variant in (I.UMA_HEAP_WRITE, I.UMA_AUX_WRITE):
            mem[off:off + 32] = vals[vr].to_bytes(32, "big")
            expect.append(("w", inc, off, None))
        else:
            expect.append(("r", inc, off, int.from_bytes(mem[off:off + 32], "big")))
    snaps, wit, cw, tail, res = run_full(orc, isa, io, st, ops, len(ops))
    assert res[0] == 0, (res[0], hex(res[4].failed_checks), res[4].first_bad_row)
    trace = res[2]
    s = lambda i: O.vm_state_at(snaps, i)
    qlen = 0
    for k, (kind, inc, off, val) in enumerate(expect):
        after = s(2 * k + 2)
        before_len, after_len = s(2 * k + 1).memory_queue_length, after.memory_queue_length
        unaligned = off % 32 != 0
        if kind == "r":
            assert reg(after, 3) == val, (k, hex(reg(after, 3)), hex(val))
            # dst1 = r4 is ENCODED in every access: write_as_dst1 is the selector bit (cycle.rs:330, :341-347), so without a flagged
            # candidate (uma.rs:944: only a read with the increment flag has one) the register takes the zero dot product
            assert reg(after, 4) == (off + 32 if inc else 0) and after.registers[3].is_pointer == 0
            assert after_len - before_len == (2 if unaligned else 1)
        else:
            if inc:
                assert reg(after, 3) == off + 32
            assert reg(after, 4) == 0 and after.registers[3].is_pointer == 0
            assert after_len - before_len == (4 if unaligned else 2)
        assert trace[K["OP_AUX"] + 2, 2 * k + 1] == off % 32 and trace[K["OP_AUX"] + 1, 2 * k + 1] == off // 32
    # heap bound untouched (bootloader frame starts at 2^24), ergs: 2 per add + 6 per uma
    assert s(len(ops)).current_context.ergs_remaining == 0xFFFFFFFF - len(plan) * 8
    # sponge slots: an unaligned write uses slots 1..4
    row = 2 * 2 + 1
    assert trace[K["SPONGE_ENFORCE"]:K["SPONGE_ENFORCE"] + 9, row].tolist() == [0, 1, 1, 1, 1, 0, 0, 0, 0]


def test_uma_fat_pointer_and_exceptions(orc):
    isa, io, st = fresh(orc)
    heap_page = 8 + 2
    data = bytes(range(1, 129))
    set_reg(st, 10, int.from_bytes(data[0:32], "big")); set_reg(st, 11, int.from_bytes(data[32:64], "big"))
    set_reg(st, 12, int.from_bytes(data[64:96], "big")); set_reg(st, 13, int.from_bytes(data[96:128], "big"))
    ptr = lambda off, start, length: off | (heap_page << 32) | (start << 64) | (length << 96)
    set_reg(st, 5, ptr(3, 64, 10), is_ptr=1)     # 7 bytes in bounds
    set_reg(st, 6, ptr(10, 64, 10), is_ptr=1)    # offset == length: nothing to read
    set_reg(st, 7, ptr(0, 33, 64), is_ptr=1)     # fully in bounds, unaligned
    set_reg(st, 8, ptr(0, 33, 64), is_ptr=0)     # not a pointer
    set_reg(st, 9, 1 << 40)                      # heap offset with high bits
    ops = []
    for i in range(4):  # fill heap[0..128)
        ops.append(isa.encode(I.OP_ADD, 0, 0, src=I.MODE_IMM16, src1=0, dst0=2, imm0=32 * i))
        ops.append(isa.encode(I.OP_UMA, I.UMA_HEAP_WRITE, 0, src0=2, src1=10 + i))
    base = len(ops)
    ops += [isa.encode(I.OP_UMA, I.UMA_PTR_READ, 1, src0=5, dst0=3, dst1=4),
            isa.encode(I.OP_UMA, I.UMA_PTR_READ, 0, src0=6, dst0=3),
            isa.encode(I.OP_UMA, I.UMA_PTR_READ, 0, src0=7, dst0=3),
            isa.encode(I.OP_UMA, I.UMA_PTR_READ, 0, src0=8, dst0=14),
            isa.encode(I.OP_NOP)]
    rc, snaps, wit, status = O.vm_run(orc, isa.isa, st, I.pack_code(ops), base + 4)
    assert rc == 0
    s = lambda i: O.vm_state_at(snaps, i)
    assert reg(s(base + 1), 3) == int.from_bytes(data[67:74] + bytes(25), "big")
    r4 = s(base + 1).registers[3]
    assert r4.is_pointer == 1 and r4.value[0] == 35 and list(r4.value[1:4]) == [heap_page, 64, 10]
    assert reg(s(base + 2), 3) == 0 and s(base + 2).memory_queue_length == s(base + 1).memory_queue_length
    assert reg(s(base + 3), 3) == int.from_bytes(data[33:65], "big")
    assert s(base + 4).pending_exception == 1 and reg(s(base + 4), 14) == 0
    # heap access with a dirty offset: panic, and the growth penalty burns every erg
    isa, io, st = fresh(orc)
    set_reg(st, 9, 1 << 40)
    rc, snaps, _, _ = O.vm_run(orc, isa.isa, st, I.pack_code([isa.encode(I.OP_UMA, I.UMA_HEAP_READ, 0, src0=9, dst0=3)]), 1)
    s1 = O.vm_state_at(snaps, 1)
    assert s1.pending_exception == 1 and s1.current_context.ergs_remaining == 0 and s1.memory_queue_length == 1
    # heap growth is paid for: a near-call frame with bound B reading beyond it
    isa, io, st = fresh(orc)
    st.current_context.heap_upper_bound = 100
    set_reg(st, 2, 200)
    rc, snaps, _, _ = O.vm_run(orc, isa.isa, st, I.pack_code([isa.encode(I.OP_UMA, I.UMA_HEAP_READ, 0, src0=2, dst0=3)]), 1)
    s1 = O.vm_state_at(snaps, 1)
    assert s1.current_context.heap_upper_bound == 232 and s1.current_context.ergs_remaining == 0xFFFFFFFF - 6 - 132


def test_near_call_and_ret_against_python_frames(orc):
    isa, io, st = fresh(orc, tail=777)
    set_reg(st, 2, 1000)  # ergs to pass
    ops = [
        isa.encode(I.OP_NEAR_CALL, src0=2, imm0=4, imm1=9),                 # 0: call 4, eh 9, pass 1000 ergs
        isa.encode(I.OP_ADD, 0, 0, src=I.MODE_IMM16, dst0=5, imm0=111),     # 1: after ok return
        isa.encode(I.OP_NEAR_CALL, src0=0, imm0=6, imm1=9),                 # 2: pass all ergs, callee reverts
        isa.encode(I.OP_NOP),                                               # 3
        isa.encode(I.OP_ADD, 0, 0, src=I.MODE_IMM16, dst0=6, imm0=5),       # 4: callee body
        isa.encode(I.OP_RET, I.RET_OK),                                     # 5
        isa.encode(I.OP_RET, I.RET_REVERT),                                 # 6: -> eh 9
        isa.encode(I.OP_NOP), isa.encode(I.OP_NOP),
        isa.encode(I.OP_NEAR_CALL, src0=0, imm0=12, imm1=14),               # 9: exception handler: call 12
        isa.encode(I.OP_RET, I.RET_OK, 1, imm0=20),                         # 10: never reached
        isa.encode(I.OP_NOP),
        isa.encode(I.OP_RET, I.RET_OK, 1, imm0=16),                         # 12: ret to label 16
        isa.encode(I.OP_NOP), isa.encode(I.OP_NOP), isa.encode(I.OP_NOP),
        isa.encode(I.OP_PTR, 0, 1, src=I.MODE_IMM16, src1=2, dst0=11, imm0=1),  # 16: ptr.add on an integer -> exception
        isa.encode(I.OP_NOP),                                               # 17: masked into ret.panic (far return from the root)
    ]
    cycles = 14
    snaps, wit, cw, tail, res = run_full(orc, isa, io, st, ops, cycles)
    assert res[0] == abi.ZKC_ERR_UNSATISFIED and res[4].failed_checks == abi.VM_CHK["BOOTLOADER_EXIT"]  # root panicked: pc != 0
    s = lambda i: O.vm_state_at(snaps, i)
    c = lambda i: s(i).current_context
    e0 = 0xFFFFFFFF
    # 0: near call
    assert (c(1).pc, c(1).exception_handler_loc, c(1).is_local_call, s(1).context_stack_depth) == (4, 9, 1, 2)
    assert c(1).ergs_remaining == 1000 and list(s(1).flags) == [0, 0, 0]
    # callee: add (2), ret ok (5): returns 1000 - 7 to the caller's e0 - 25 - 1000
    assert (c(3).pc, s(3).context_stack_depth, c(3).is_local_call) == (1, 1, 0)
    assert c(3).ergs_remaining == e0 - 25 - 1000 + (1000 - 7) and reg(s(3), 6) == 5
    assert list(s(3).stack_sponge_state) == list(s(0).stack_sponge_state)
    assert list(s(1).stack_sponge_state) != list(s(0).stack_sponge_state)
    # 1: add, 2: near call passing everything, 6: revert -> eh 9 of the callee frame
    assert reg(s(4), 5) == 111 and c(5).pc == 6 and c(5).ergs_remaining == c(4).ergs_remaining - 25
    assert (c(6).pc, s(6).context_stack_depth) == (9, 1) and list(s(6).flags) == [0, 0, 0]
    assert c(6).ergs_remaining == c(4).ergs_remaining - 25 - 5
    # 9: near call 12; 12: ret to label 16
    assert c(7).pc == 12 and (c(8).pc, s(8).context_stack_depth) == (16, 1)
    # 16: exception, 17: panic out of the root frame
    assert s(9).pending_exception == 1
    fin = s(10)
    assert fin.context_stack_depth == 0 and fin.pending_exception == 0 and list(fin.flags) == [1, 0, 0]
    assert fin.current_context.pc == 0xFFFF  # the root's exception handler location
    assert fin.registers[0].is_pointer == 1 and all(reg(fin, r) == 0 for r in range(1, 16))
    assert all(fin.registers[r].is_pointer == 0 for r in range(1, 15))
    # afterwards the VM idles: every cycle is skipped, timestamps stop
    assert s(11).previous_code_page == 0 and s(11).timestamp == s(10).timestamp  # the popped (empty) frame's code page
    assert bytes(s(12)) == bytes(s(11)) and bytes(s(14)) == bytes(s(11))
    assert res[2][K["SHOULD_SKIP_CYCLE"], 10:].tolist() == [1] * 4
    assert res[1].completion_flag == 1
    # popped frames: 4 rets (ok, revert, label, root panic)
    assert len(cw) == 4 and res[2][K["OP_AUX"] + 43].sum() == 4 and res[2][K["OP_AUX"] + 42].sum() == 3
    # rollback queue: nothing was logged, so every frame's tail is where its fate puts it: the ok / label frames start
    # at the parent's head (the block tail), the reverted frame at the forward tail (empty queue: zeros)
    w = lambda i: abi.VmCycleWitness.from_buffer_copy(wit[i].tobytes())
    # the root itself panics at the end: its segment must sit at the (empty) forward tail, and so must every ok child's
    assert list(tail) == [0, 0, 0, 0] and all(list(w(i).rollback) == [0, 0, 0, 0] for i in (0, 4, 6))
    # without the root's panic (9 cycles) the block tail stays as given and the fates differ
    snaps9, wit9, cw9, tail9, res9 = run_full(orc, isa, io, st, ops, 9)
    assert res9[0] == 0 and list(tail9) == [777, 0, 0, 0]
    w9 = lambda i: abi.VmCycleWitness.from_buffer_copy(wit9[i].tobytes())
    assert list(w9(0).rollback) == [777, 0, 0, 0] and list(w9(4).rollback) == [0, 0, 0, 0] and list(w9(6).rollback) == [777, 0, 0, 0]


def lq(address, key, read, written, tx, ts, flags):
    q = np.zeros(1, dtype=abi.LOG_QUERY_DTYPE)
    q["address"][0, 0] = address
    for i in range(8):
        q["key"][0, i] = (key >> (32 * i)) & 0xFFFFFFFF
        q["read_value"][0, i] = (read >> (32 * i)) & 0xFFFFFFFF
        q["written_value"][0, i] = (written >> (32 * i)) & 0xFFFFFFFF
    q["tx_number_in_block"], q["timestamp"], q["flags"] = tx, ts, flags
    return q


def test_log_storage_events_and_rollback_queue(orc):
    isa, io, st = fresh(orc, tail=4242)
    A, B, Cv = 0x1111 << 200 | 5, 0x2222 << 100 | 6, 0x3333
    set_reg(st, 2, 7)      # key
    set_reg(st, 3, A); set_reg(st, 4, B); set_reg(st, 5, Cv); set_reg(st, 6, 9)  # second key
    ops = [
        isa.encode(I.OP_LOG, I.LOG_STORAGE_WRITE, src0=2, src1=3),            # 0: [7] = A
        isa.encode(I.OP_LOG, I.LOG_STORAGE_READ, src0=2, dst0=10),            # 1: r10 = A
        isa.encode(I.OP_NEAR_CALL, src0=0, imm0=8, imm1=12),                  # 2: frame that reverts
        isa.encode(I.OP_LOG, I.LOG_STORAGE_READ, src0=2, dst0=11),            # 3: r11 = A again (the write of the frame is undone)
        isa.encode(I.OP_NEAR_CALL, src0=0, imm0=14, imm1=12),                 # 4: frame that returns ok
        isa.encode(I.OP_LOG, I.LOG_STORAGE_READ, src0=6, dst0=12),            # 5: r12 = C (kept)
        isa.encode(I.OP_LOG, I.LOG_EVENT, 1, src0=2, src1=5),                 # 6: event (first message)
        isa.encode(I.OP_JUMP, 0, 0, src=I.MODE_IMM16, imm0=20),               # 7
        isa.encode(I.OP_LOG, I.LOG_STORAGE_WRITE, src0=2, src1=4),            # 8: [7] = B
        isa.encode(I.OP_LOG, I.LOG_STORAGE_READ, src0=2, dst0=13),            # 9: r13 = B
        isa.encode(I.OP_LOG, I.LOG_EVENT, 0, src0=3, src1=4),                 # 10
        isa.encode(I.OP_RET, I.RET_REVERT),                                   # 11 -> eh 12
        isa.encode(I.OP_JUMP, 0, 0, src=I.MODE_IMM16, imm0=3),                # 12: exception handler: continue at 3
        isa.encode(I.OP_NOP),
        isa.encode(I.OP_LOG, I.LOG_STORAGE_WRITE, src0=6, src1=5),            # 14: [9] = C
        isa.encode(I.OP_RET, I.RET_OK),                                       # 15
    ] + [isa.encode(I.OP_NOP)] * 4 + [isa.encode(I.OP_LOG, I.LOG_PRECOMPILE, src0=2, src1=6, dst0=14), isa.encode(I.OP_NOP)]
    cycles = 17
    snaps, wit, cw, tail, res = run_full(orc, isa, io, st, ops, cycles)
    assert res[0] == 0, (res[0], hex(res[4].failed_checks), res[4].first_bad_row)  # every rollback / forward queue join holds
    s = lambda i: O.vm_state_at(snaps, i)
    fin = s(cycles)
    assert reg(fin, 10) == A and reg(fin, 13) == B and reg(fin, 11) == A and reg(fin, 12) == Cv and reg(fin, 14) == 1
    # execution order: 0 1 2 | 8 9 10 11 | 12 3 4 | 14 15 | 5 6 7 20(precompile)
    ctx = fin.current_context
    # forward queue: write A, read, [write B, read, event, + 2 rollbacks appended by the revert], read, write C, read, event, precompile
    assert ctx.log_queue_forward_part_length == 2 + 3 + 2 + 1 + 1 + 1 + 1 + 1
    # rollback segment of the root: write A, (merged from the ok frame) write C, event
    assert ctx.reverted_queue_segment_len == 3
    # the forward queue is the sorter circuits' hash chain over exactly these records (timestamps: cycle ts + 1)
    ts = lambda cyc: 1024 + 4 * cyc + 1
    ST, EV, PC = 0, 1, 3
    f = abi.lq_flags
    this = 0x8001
    recs = [
        lq(this, 7, 0, A, 0, ts(0), f(ST, 0, 1)), lq(this, 7, A, A, 0, ts(1), f(ST, 0, 0)),
        lq(this, 7, A, B, 0, ts(3), f(ST, 0, 1)), lq(this, 7, B, B, 0, ts(4), f(ST, 0, 0)), lq(this, A, 0, B, 0, ts(5), f(EV, 0, 1)),
        # the revert appends the frame's rollback segment, most recent first, with the rollback flag set
        lq(this, A, 0, B, 0, ts(5), f(EV, 0, 1, 1)), lq(this, 7, A, B, 0, ts(3), f(ST, 0, 1, 1)),
        lq(this, 7, A, A, 0, ts(8), f(ST, 0, 0)),
        lq(this, 9, 0, Cv, 0, ts(10), f(ST, 0, 1)), lq(this, 9, Cv, Cv, 0, ts(12), f(ST, 0, 0)),
        lq(this, 7, 0, Cv, 0, ts(13), f(EV, 0, 1, 0, 1)),
    ]
    pre = lq(this, 7 | (10 << 128) | (10 << 160), 0, 0, 0, ts(15), f(PC, 0, 0))
    recs.append(pre)
    prev, final = O.log_queue_simulate(orc, np.concatenate(recs))
    assert list(final.tail) == list(ctx.log_queue_forward_tail)
    # the root's rollback segment, read from its head, is [event, write C, write A] with the rollback flag, ending at the block tail
    seg = [lq(this, 7, 0, Cv, 0, ts(13), f(EV, 0, 1, 1, 1)), lq(this, 9, 0, Cv, 0, ts(10), f(ST, 0, 1, 1)), lq(this, 7, 0, A, 0, ts(0), f(ST, 0, 1, 1))]
    chain = np.array(list(ctx.reverted_queue_head), dtype=np.uint64)
    lib = orc
    for r in seg:
        enc = np.zeros(20, dtype=np.uint64)
        lib.orc_log_query_encode(O.p(r), O.p(enc))
        lib.orc_log_queue_absorb(O.p(chain), O.p(enc), None)
    assert list(chain) == list(ctx.reverted_queue_tail) == list(tail)
    assert list(tail) != [4242, 0, 0, 0]  # the block's rollback tail is an output of the resolution
    # ergs: storage writes cost nothing extra at 0 ergs per pubdata byte; precompile burns src1[0] = 9
    t = res[2]
    assert t[K["OP_AUX"] + 30, 15] == 9 and t[K["OP_AUX"] + 28, 15] == 1
    # a wrong claimed rollback head is caught at its cycle
    w2 = wit.copy()
    w2[0, 144] ^= 1
    io2 = abi.VmClosedForm.from_buffer_copy(bytes(res[1])); io2.start_flag = 0; io2.hidden_fsm_input = s(0)
    bad = O.vm_entry_point(orc, io2, isa.isa, snaps, w2, cycles, cw=cw)
    assert bad[0] == abi.ZKC_ERR_SNAPSHOT_MISMATCH and bad[4].failed_checks & abi.VM_CHK["ROLLBACK_QUEUE"] and bad[4].first_bad_row == 0


def test_precompile_and_l1_message_costs(orc):
    isa, io, st = fresh(orc)
    st.ergs_per_pubdata_byte = 3
    set_reg(st, 2, 7); set_reg(st, 3, 0xABC); set_reg(st, 4, 50)
    ops = [isa.encode(I.OP_LOG, I.LOG_TO_L1, src0=2, src1=3),              # burns 3 * 88
           isa.encode(I.OP_LOG, I.LOG_STORAGE_WRITE, src0=2, src1=3),      # first write: 3 * 64
           isa.encode(I.OP_LOG, I.LOG_STORAGE_WRITE, src0=2, src1=4),      # repeated write: refund 64 -> 0
           isa.encode(I.OP_LOG, I.LOG_PRECOMPILE, src0=2, src1=4, dst0=9)]  # burns 50
    rc, snaps, wit, status = O.vm_run(orc, isa.isa, st, I.pack_code(ops), 4)
    assert rc == 0
    e = [O.vm_state_at(snaps, i).current_context.ergs_remaining for i in range(5)]
    assert [e[i] - e[i + 1] for i in range(4)] == [40 + 3 * 88, 40 + 3 * 64, 40, 40 + 50]
    # not enough ergs for the burn: nothing is logged, ergs go to zero, precompile result 0
    isa, io, st = fresh(orc)
    st.current_context.ergs_remaining = 60
    set_reg(st, 2, 7); set_reg(st, 4, 50)
    rc, snaps, wit, status = O.vm_run(orc, isa.isa, st, I.pack_code([isa.encode(I.OP_LOG, I.LOG_PRECOMPILE, src0=2, src1=4, dst0=9)]), 1)
    s1 = O.vm_state_at(snaps, 1)
    assert s1.current_context.ergs_remaining == 0 and reg(s1, 9) == 0 and s1.current_context.log_queue_forward_part_length == 0


def test_random_programs_with_every_opcode(orc):
    for seed in (1, 2, 3):
        isa, io, st = fresh(orc, tail=seed)
        ops = I.random_program(isa, 512, seed=seed)
        cycles = 3000
        rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
        assert rc == 0, (hex(status.failed_checks), status.first_bad_row)
        for i in range(4):
            io.rollback_queue_tail_for_block[i] = int(tail[i])
        rc, out, trace, com, status = O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles, cw=cw)
        assert rc == 0, (rc, hex(status.failed_checks), status.first_bad_row)
        props = trace[K["PROPS"]]
        for op in (I.OP_UMA, I.OP_LOG, I.OP_NEAR_CALL, I.OP_RET, I.OP_MUL, I.OP_PTR):
            assert ((props >> np.uint64(op)) & np.uint64(1)).sum() > 3, op
        assert trace[K["OP_AUX"] + 45].sum() > 0  # some frames revert
        assert trace[K["SPONGE_ENFORCE"] + 4].sum() > 10
        # chained instances == whole
        cut = 1234
        rc, a, ta, _, _ = O.vm_entry_point(orc, io, isa.isa, snaps[:cut + 1], wit[:cut], cut, cw=cw)
        nxt = abi.VmClosedForm.from_buffer_copy(bytes(a)); nxt.start_flag = 0
        nxt.hidden_fsm_input = a.hidden_fsm_output
        rc, b, tb, com_b, _ = O.vm_entry_point(orc, nxt, isa.isa, snaps[cut:], wit[cut:], cycles - cut, cw=cw)
        assert rc == 0 and bytes(b.hidden_fsm_output) == bytes(out.hidden_fsm_output)
        assert np.array_equal(np.concatenate([ta, tb], axis=1), trace)


def far_call_program(isa):
    return [
        isa.encode(I.OP_CONTEXT, 1, 0, dst0=10),                                   # 0: r10 = caller
        isa.encode(I.OP_SUB, 0, 1, src0=10, src1=0, dst0=11),                      # 1: EQ iff caller == 0 (the root frame)
        isa.encode(I.OP_JUMP, 0, 0, src=I.MODE_IMM16, cond=I.COND_NE, imm0=12),    # 2: callee -> 12
        isa.encode(I.OP_FAR_CALL, I.FAR_CALL_NORMAL, 0, src0=2, src1=3, imm0=10),  # 3: call r3 with ABI r2, eh 10
        isa.encode(I.OP_ADD, 0, 0, src=I.MODE_IMM16, dst0=6, imm0=77),             # 4: after the ok return
        isa.encode(I.OP_FAR_CALL, I.FAR_CALL_NORMAL, 0, src0=2, src1=4, imm0=10),  # 5: registers were cleaned: address 0, no code
        isa.encode(I.OP_NOP), isa.encode(I.OP_NOP), isa.encode(I.OP_NOP), isa.encode(I.OP_NOP),
        isa.encode(I.OP_ADD, 0, 0, src=I.MODE_IMM16, dst0=7, imm0=99),             # 10: exception handler
        isa.encode(I.OP_JUMP, 0, 0, src=I.MODE_IMM16, imm0=20),                    # 11
        isa.encode(I.OP_ADD, 0, 0, src=I.MODE_IMM16, dst0=8, imm0=5),              # 12: callee body
        isa.encode(I.OP_LOG, I.LOG_STORAGE_WRITE, src0=5, src1=8),                 # 13: [0] = 5 in the callee's storage
        isa.encode(I.OP_UMA, I.UMA_HEAP_WRITE, 0, src0=5, src1=8),                 # 14: heap[0] = 5 on the callee's own heap page
        isa.encode(I.OP_RET, I.RET_OK),                                            # 15
    ] + [isa.encode(I.OP_NOP)] * 8


def test_far_call_decommit_and_return(orc):
    isa, io, st = fresh(orc, tail=31)
    abi_reg = 100000 << 192                     # ergs_passed in bits 192..224, forwarding mode 0 (use heap), normal call
    set_reg(st, 2, abi_reg); set_reg(st, 3, 0x9001); set_reg(st, 4, 0x9002); set_reg(st, 9, 0xABCDEF)
    ops = far_call_program(isa)
    cycles = 20
    snaps, wit, cw, tail, res = run_full(orc, isa, io, st, ops, cycles)
    assert res[0] == 0, (res[0], hex(res[4].failed_checks), res[4].first_bad_row)
    s = lambda i: O.vm_state_at(snaps, i)
    c = lambda i: s(i).current_context
    words = len(I.pack_code(ops))
    # cycle 3: the far call
    callee = c(4)
    assert s(4).context_stack_depth == 2 and s(4).memory_page_counter == 16 + 8 and s(4).pending_exception == 0
    assert (callee.this_address[0], callee.caller[0], callee.code_address[0]) == (0x9001, 0x8001, 0x9001)
    assert (callee.base_page, callee.code_page, callee.pc, callee.exception_handler_loc) == (16, 16, 0, 10)
    assert (callee.is_kernel_mode, callee.is_local_call, callee.is_static_execution) == (1, 0, 0)
    assert (callee.heap_upper_bound, callee.aux_heap_upper_bound, callee.ergs_remaining) == (4096, 4096, 100000)
    assert s(4).code_decommittment_queue_length == 1 and callee.log_queue_forward_part_length == 1
    # r1 = (empty) calldata pointer into the caller's heap, r2 = flags, r3..r15 cleaned for a non-system call
    r1 = s(4).registers[0]
    assert r1.is_pointer == 1 and list(r1.value[:4]) == [0, 8 + 2, 0, 0]
    assert all(reg(s(4), r) == 0 for r in range(2, 16))
    # caller frame: 63/64 rule after the decommit cost
    e_before = c(3).ergs_remaining
    after_decommit = e_before - 100 - 4 * words
    assert len(cw) == 2
    saved = abi.VmCallstackWitness.from_buffer_copy(cw[0].tobytes()).context
    assert saved.ergs_remaining == after_decommit - 100000 and saved.pc == 4
    # the callee runs the same program from 0: caller != 0 -> body at 12, writes its own storage + heap, returns
    assert c(7).pc == 12 and reg(s(8), 8) == 5
    ret_state = s(11)
    assert ret_state.context_stack_depth == 1 and ret_state.current_context.pc == 4 and list(ret_state.flags) == [0, 0, 0]
    assert ret_state.registers[0].is_pointer == 1 and list(ret_state.registers[0].value[:4]) == [0, 16 + 2, 0, 0]
    assert all(reg(ret_state, r) == 0 for r in range(2, 16)) and list(ret_state.context_composite_u128) == [0, 0, 0, 0]
    assert ret_state.current_context.reverted_queue_segment_len == 1  # the callee's storage write, handed to the root
    # cycle 11: add; cycle 12: far call to address 0 (kernel space, no code): exception, frame with the unmapped page
    assert reg(s(12), 6) == 77
    bad = s(13)
    assert bad.context_stack_depth == 2 and bad.pending_exception == 1 and bad.current_context.code_page == 0
    assert bad.memory_page_counter == 16 + 16 and bad.code_decommittment_queue_length == 1
    # next cycle: the pending exception is a ret.panic out of the callee: back in the root at its handler
    back = s(14)
    assert back.context_stack_depth == 1 and back.current_context.pc == 10 and list(back.flags) == [1, 0, 0]
    assert reg(s(15), 7) == 99
    t = res[2]
    assert t[K["OP_AUX"] + 46].sum() == 2 and t[K["OP_AUX"] + 47].tolist()[12] == 1
    assert t[K["SPONGE_ENFORCE"] + 5:K["SPONGE_ENFORCE"] + 9, 3].tolist() == [1, 1, 1, 1]
    assert t[K["SPONGE_ENFORCE"] + 5:K["SPONGE_ENFORCE"] + 9, 12].tolist() == [1, 1, 1, 0]  # the code hash is read, nothing is decommitted
    # delegate / mimic calls keep / forge the caller
    for variant, want_this, want_caller in ((I.FAR_CALL_DELEGATE, 0x8001, 0), (I.FAR_CALL_MIMIC, 0x9001, 0x7777)):
        isa, io, st = fresh(orc)
        set_reg(st, 2, abi_reg); set_reg(st, 3, 0x9001); set_reg(st, 15, 0x7777)
        rc, sn, _, _ = O.vm_run(orc, isa.isa, st, I.pack_code([isa.encode(I.OP_FAR_CALL, variant, 1, src0=2, src1=3, imm0=3)]), 1)
        cc = O.vm_state_at(sn, 1).current_context
        assert rc == 0 and (cc.this_address[0], cc.caller[0], cc.code_address[0], cc.is_static_execution) == (want_this, want_caller, 0x9001, 1)
    # a user-space address without code runs the default account code hash of the block
    isa, io, st = fresh(orc)
    io.default_aa_code_hash[7] = (1 << 24) | 3; io.default_aa_code_hash[0] = 0xAA
    set_reg(st, 2, abi_reg); set_reg(st, 3, 0x12340002)
    rc, sn, wt, _ = O.vm_run(orc, isa.isa, st, I.pack_code([isa.encode(I.OP_FAR_CALL, 0, 0, src0=2, src1=3, imm0=3)]), 1, gc=io)
    s1 = O.vm_state_at(sn, 1)
    assert rc == 0 and s1.pending_exception == 0 and s1.current_context.is_kernel_mode == 0 and s1.code_decommittment_queue_length == 1
    assert s1.current_context.code_page == 16  # unknown code: a fresh page


def test_encoded_dst1_register_is_written_whatever_the_gadget_flags(orc):
    """cycle.rs:330, :341-347: write_as_dst1 = the decoded dst1 selector bit (should_update_dst1, :177-187, is never read).  An add that
    encodes dst1 = r5 zeroes r5 (and its pointer marker); the same add masked into a NOP by a false condition decodes dst1 = 0
    (decoded_opcode.rs:170-177) and leaves r6 alone; mul writes its high half there; dst1 wins over dst0 on the same register."""
    isa, io, st = fresh(orc)
    set_reg(st, 5, 0x1234 << 200, is_ptr=1); set_reg(st, 6, 77, is_ptr=1); set_reg(st, 7, 5); set_reg(st, 8, (1 << 255) + 9)
    ops = [isa.encode(I.OP_ADD, 0, 0, src0=7, src1=7, dst0=2, dst1=5),                       # r2 = 10, r5 <- (not a pointer, 0)
           isa.encode(I.OP_ADD, 0, 0, cond=I.COND_EQ, src0=7, src1=7, dst0=3, dst1=6),       # eq flag is clear: NOP, r3 / r6 untouched
           isa.encode(I.OP_MUL, 0, 0, src0=8, src1=8, dst0=9, dst1=10),                      # r9 = low, r10 = high of r8 * r8
           isa.encode(I.OP_ADD, 0, 0, src0=7, src1=7, dst0=11, dst1=11)]                     # dst0 = dst1 = r11: dst1 (zero) is applied last
    snaps, wit, cw, tail, res = run_full(orc, isa, io, st, ops, len(ops))
    assert res[0] == 0, (res[0], hex(res[4].failed_checks), res[4].first_bad_row)
    s = lambda i: O.vm_state_at(snaps, i)
    assert reg(s(1), 2) == 10 and reg(s(1), 5) == 0 and s(1).registers[4].is_pointer == 0
    assert reg(s(2), 3) == 0 and reg(s(2), 6) == 77 and s(2).registers[5].is_pointer == 1
    sq = ((1 << 255) + 9) ** 2
    assert reg(s(3), 9) == sq % (1 << 256) and reg(s(3), 10) == sq >> 256
    assert reg(s(4), 11) == 0
    trace = res[2]
    assert trace[K["DST1_REG"]].tolist()[:4] == [5, 0, 10, 11] and trace[K["DST1_UPDATE_REGISTER"]].tolist()[:4] == [0, 0, 1, 0]
    assert trace[K["DST1"]:K["DST1"] + 9, 0].max() == 0
