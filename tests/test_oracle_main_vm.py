"""main_vm oracle (value-level restatement of vm_cycle): arithmetic / addressing subset; uma / log / calls are in
test_oracle_main_vm_ops.py.  PARITY UNPINNED against the
reference (no main_vm test, un-vendored ISA tables): the semantics are pinned here against Python big-int arithmetic
on hand-written programs, instruction by instruction."""
import ctypes as C

import numpy as np

import orc as O
from era_zkevm_circuits_b200 import abi, isa as I

M256 = (1 << 256) - 1
K = abi.VM_COLS


def reg(state, r):
    return sum(int(v) << (32 * i) for i, v in enumerate(state.registers[r - 1].value))


def fresh(orc):
    isa = I.Isa()
    io = abi.VmClosedForm()
    io.start_flag = 1
    st = O.vm_initial_state(orc, io, isa.isa)
    return isa, io, st


def run(orc, isa, st, ops, cycles=None):
    cycles = cycles or len(ops)
    rc, snaps, wit, status = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles)
    return rc, snaps, wit, status


def set_reg(st, r, v, is_ptr=0):
    st.registers[r - 1].is_pointer = is_ptr
    for i in range(8):
        st.registers[r - 1].value[i] = (v >> (32 * i)) & 0xFFFFFFFF


def test_initial_bootloader_state(orc):
    isa, io, st = fresh(orc)
    c = st.current_context
    assert (c.base_page, c.code_page, c.pc, c.sp, c.is_kernel_mode) == (8, 8, 0, 0, 1)
    assert c.ergs_remaining == 0xFFFFFFFF and c.exception_handler_loc == 0xFFFF and c.this_address[0] == 0x8001
    assert st.context_stack_depth == 1 and st.timestamp == 1024 and st.registers[0].is_pointer == 1
    assert any(st.stack_sponge_state)  # 4 Poseidon2 rounds over the empty-frame encoding


def test_arithmetic_against_python_ints(orc):
    isa, io, st = fresh(orc)
    a, b = (1 << 256) - 0x1234567890ABCDEF, (1 << 200) + 77
    set_reg(st, 2, a); set_reg(st, 3, b); set_reg(st, 4, 0); set_reg(st, 5, 13)
    ops = [
        isa.encode(I.OP_ADD, 0, 1, src0=2, src1=3, dst0=6),                 # r6 = a + b (overflow -> LT flag)
        isa.encode(I.OP_SUB, 0, 1, src0=3, src1=2, dst0=7),                 # r7 = b - a (borrow)
        isa.encode(I.OP_SUB, 0, 3, src0=3, src1=2, dst0=8),                 # swapped: r8 = a - b
        isa.encode(I.OP_MUL, 0, 0, src0=2, src1=3, dst0=9, dst1=10),        # r9, r10 = lo, hi
        isa.encode(I.OP_DIV, 0, 0, src0=2, src1=5, dst0=11, dst1=12),       # q, r
        isa.encode(I.OP_DIV, 0, 1, src0=2, src1=4, dst0=13, dst1=14),       # division by zero: 0, 0, LT set
        isa.encode(I.OP_BINOP, 0, 0, src0=2, src1=3, dst0=15),              # xor
        isa.encode(I.OP_SHIFT, 0, 0, src0=2, src1=5, dst0=2),               # shl 13
        isa.encode(I.OP_SHIFT, 1, 0, src0=3, src1=5, dst0=3),               # shr 13
    ]
    rc, snaps, wit, status = run(orc, isa, st, ops)
    assert rc == 0
    s = lambda i: O.vm_state_at(snaps, i)
    assert reg(s(1), 6) == (a + b) & M256 and list(s(1).flags) == [1, 0, 0]
    assert reg(s(2), 7) == (b - a) & M256 and list(s(2).flags) == [1, 0, 0]
    assert reg(s(3), 8) == a - b and list(s(3).flags) == [0, 0, 1]
    assert reg(s(4), 9) == (a * b) & M256 and reg(s(4), 10) == (a * b) >> 256
    assert reg(s(5), 11) == a // 13 and reg(s(5), 12) == a % 13
    assert reg(s(6), 13) == 0 and reg(s(6), 14) == 0 and list(s(6).flags) == [1, 0, 0]
    assert reg(s(7), 15) == a ^ b
    assert reg(s(8), 2) == (a << 13) & M256 and reg(s(9), 3) == b >> 13
    # rotations
    isa, io, st = fresh(orc)
    set_reg(st, 2, a); set_reg(st, 5, 77)
    ops = [isa.encode(I.OP_SHIFT, 2, 0, src0=2, src1=5, dst0=6), isa.encode(I.OP_SHIFT, 3, 0, src0=2, src1=5, dst0=7),
           isa.encode(I.OP_BINOP, 1, 1, src0=2, src1=4, dst0=8), isa.encode(I.OP_BINOP, 2, 0, src0=2, src1=5, dst0=9)]
    rc, snaps, _, _ = run(orc, isa, st, ops)
    s = lambda i: O.vm_state_at(snaps, i)
    assert reg(s(1), 6) == ((a << 77) | (a >> (256 - 77))) & M256
    assert reg(s(2), 7) == ((a >> 77) | (a << (256 - 77))) & M256
    assert reg(s(3), 8) == 0 and list(s(3).flags) == [0, 1, 0]  # and with zero register sets EQ
    assert reg(s(4), 9) == a | 77


def test_addressing_modes_stack_and_memory_queue(orc):
    isa, io, st = fresh(orc)
    set_reg(st, 2, 1000); set_reg(st, 3, 7)
    ops = [
        isa.encode(I.OP_ADD, 0, 0, dst=I.MODE_PUSH_POP, src0=2, src1=3, imm1=1),      # push (r2 + r3): stack[sp=0] = 1007, sp -> 1
        isa.encode(I.OP_ADD, 0, 0, dst=I.MODE_PUSH_POP, src0=2, src1=2, imm1=2),      # stack[1] = 2000, sp -> 3
        isa.encode(I.OP_ADD, 0, 0, src=I.MODE_STACK_OFFSET, src1=3, dst0=4, imm0=3),  # r4 = stack[sp - 3] + 7 = 1014
        isa.encode(I.OP_ADD, 0, 0, src=I.MODE_PUSH_POP, src1=0, dst0=5, imm0=2),      # pop 2: r5 = stack[3 - 2] = 2000, sp -> 1
        isa.encode(I.OP_ADD, 0, 0, src=I.MODE_STACK_ABS, dst=I.MODE_STACK_ABS, src1=3, imm0=0, imm1=9),  # stack[9] = stack[0] + 7
        isa.encode(I.OP_ADD, 0, 0, src=I.MODE_STACK_ABS, src1=0, dst0=6, imm0=9),
        isa.encode(I.OP_ADD, 0, 0, src=I.MODE_IMM16, src1=3, dst0=7, imm0=0xBEEF),
        isa.encode(I.OP_ADD, 0, 0, src=I.MODE_CODE, src1=0, dst0=8, imm0=1),          # r8 = code word 1
    ]
    rc, snaps, wit, _ = run(orc, isa, st, ops)
    assert rc == 0
    s = lambda i: O.vm_state_at(snaps, i)
    assert s(1).current_context.sp == 1 and s(2).current_context.sp == 3 and s(4).current_context.sp == 1
    assert reg(s(3), 4) == 1014 and reg(s(4), 5) == 2000 and reg(s(6), 6) == 1014 and reg(s(7), 7) == 0xBEEF + 7
    code = I.pack_code(ops)
    assert reg(s(8), 8) == sum(int(v) << (32 * i) for i, v in enumerate(code[1]))
    # memory queue growth: 2 code-word fetches (8 instructions = 2 words) + operand reads + writes
    lens = [s(i).memory_queue_length for i in range(9)]
    assert lens == [0, 2, 3, 4, 5, 8, 9, 9, 10]
    # timestamps advance by 4 per executed cycle
    assert s(8).timestamp == 1024 + 32


def test_conditions_jump_context_ptr(orc):
    isa, io, st = fresh(orc)
    set_reg(st, 2, 5); set_reg(st, 3, 5)
    ops = [
        isa.encode(I.OP_SUB, 0, 1, src0=2, src1=3, dst0=4),                                # EQ set
        isa.encode(I.OP_ADD, 0, 0, cond=I.COND_NE, src0=2, src1=3, dst0=5),                # skipped (masked into NOP)
        isa.encode(I.OP_ADD, 0, 0, cond=I.COND_EQ, src0=2, src1=3, dst0=6),                # executed
        isa.encode(I.OP_JUMP, 0, 0, src=I.MODE_IMM16, imm0=6),                             # pc -> 6
        isa.encode(I.OP_ADD, 0, 0, src0=2, src1=3, dst0=7),                                # skipped by the jump
        isa.encode(I.OP_ADD, 0, 0, src0=2, src1=3, dst0=7),
        isa.encode(I.OP_CONTEXT, 4, 0, dst0=8),                                            # ergs left
        isa.encode(I.OP_CONTEXT, 0, 0, dst0=9),                                            # this
        isa.encode(I.OP_CONTEXT, 7, 0, src0=2),                                            # set context u128
        isa.encode(I.OP_PTR, 0, 1, src=I.MODE_IMM16, src1=1, dst0=10, imm0=9),             # ptr.add r1, 9
        isa.encode(I.OP_PTR, 0, 1, src=I.MODE_IMM16, src1=2, dst0=11, imm0=9),             # not a pointer -> pending exception
    ]
    rc, snaps, wit, status = run(orc, isa, st, ops, cycles=9)
    assert rc == 0, (rc, hex(status.failed_checks), status.first_bad_row)
    s = lambda i: O.vm_state_at(snaps, i)
    assert list(s(1).flags) == [0, 1, 0]
    assert reg(s(2), 5) == 0 and reg(s(3), 6) == 10
    assert s(4).current_context.pc == 6  # the two instructions behind the jump never execute
    assert reg(s(5), 8) == 0xFFFFFFFF - 5 * 2  # a condition-masked instruction still pays its price (decoded_opcode.rs:391-394)
    assert reg(s(6), 9) == 0x8001 and list(s(7).context_composite_u128) == [5, 0, 0, 0]
    assert reg(s(8), 10) == (7 << 32) + 9 and s(8).registers[9].is_pointer == 1 and reg(s(8), 7) == 0
    assert s(9).pending_exception == 1 and reg(s(9), 11) == 0
    # the pending exception is masked into ret.panic next cycle: the root frame is left
    rc, snaps, wit, status = run(orc, isa, st, ops, cycles=10)
    assert rc == 0 and O.vm_state_at(snaps, 10).context_stack_depth == 0


def test_entry_point_on_random_program(orc):
    isa, io, st = fresh(orc)
    ops = I.random_program(isa, 256, seed=11, full=False)
    cycles = 700
    rc, snaps, wit, status = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles)
    assert rc == 0, (hex(status.failed_checks), status.first_bad_row)
    rc, out, trace, com, status = O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles)
    assert rc == 0, (rc, hex(status.failed_checks), status.first_bad_row)
    assert bytes(out.hidden_fsm_output) == snaps[cycles].tobytes() and out.completion_flag == 0
    assert trace[K["SHOULD_READ_SRC0"]].sum() > 50 and trace[K["PERFORM_DST0_MEMORY_WRITE"]].sum() > 20
    assert trace[K["MASK_INTO_NOP"]].sum() > 5 and trace[K["MASK_INTO_PANIC"]].sum() == 0
    # chained instances == whole
    cut = 300
    rc, a, ta, _, _ = O.vm_entry_point(orc, io, isa.isa, snaps[:cut + 1], wit[:cut], cut)
    nxt = abi.VmClosedForm.from_buffer_copy(bytes(a)); nxt.start_flag = 0
    nxt.hidden_fsm_input = a.hidden_fsm_output
    rc, b, tb, com_b, _ = O.vm_entry_point(orc, nxt, isa.isa, snaps[cut:], wit[cut:], cycles - cut)
    assert rc == 0 and bytes(b.hidden_fsm_output) == bytes(out.hidden_fsm_output)
    assert np.array_equal(np.concatenate([ta, tb], axis=1), trace)
    # a corrupted snapshot is detected at its cycle
    bad = snaps.copy(); bad[123, 40] ^= 1
    rc, _, _, _, status = O.vm_entry_point(orc, io, isa.isa, bad, wit, cycles)
    assert rc == abi.ZKC_ERR_SNAPSHOT_MISMATCH and status.first_bad_row == 123
