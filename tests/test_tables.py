"""The reference's lookup tables (src/tables/*.rs), restated here in Python from their generating functions, against what the
oracle computes in their place (the engine computes the same expressions; the GPU parity tests tie the two together):

    create_shift_to_num_converter_table     tables/bitshift.rs:12-40            -> SHIFT_CONSTANT of the gadget block, all 256 shifts
    create_conditionals_resolution_table    tables/conditional.rs:21-58         -> the CONDITION column of executed traces
    create_integer_to_bitmask_table         tables/integer_to_boolean_mask.rs:22-44 (register index, sub-pc)
                                                                                -> operand / destination register selection, opcode
                                                                                   selection inside the 32-byte code word
    create_uma_ptr_read_bitmask_table       tables/uma_ptr_read_cleanup.rs:12-40 -> fat-pointer reads past the slice end, all 32 keys

zkevm_opcode_defs (the ISA tables of tables/opcodes_decoding.rs) is un-vendored: the decoding table is host-supplied data
(zkc_vm_isa), so there is nothing of the reference to restate for it here."""
import numpy as np

import orc as O
from era_zkevm_circuits_b200 import abi, isa as I
from test_oracle_main_vm_ops import fresh, reg, set_reg

K, G = abi.VM_COLS, abi.VMG_COLS


def shift_to_num_converter_table():
    """tables/bitshift.rs:18-33: key = shift + (idx << 8) -> two consecutive 32-bit limbs of 1 << shift"""
    rows = {}
    for shift in range(256):
        modulus = 1 << shift
        for idx in range(4):
            y = modulus & 0xFFFFFFFF; modulus >>= 32
            z = modulus & 0xFFFFFFFF; modulus >>= 32
            rows[shift + (idx << 8)] = (y, z)
    assert len(rows) == 1024
    return rows


def test_bitshift_table_is_the_shift_constant(orc):
    table = shift_to_num_converter_table()
    for variant, name in ((0, "shl"), (3, "ror")):
        trace = np.zeros((K["NUM_COLS"], 256), dtype=np.uint64)
        trace[K["PROPS"]] = (1 << I.OP_SHIFT) | (1 << (16 + variant))
        trace[K["SRC1"] + 1] = np.arange(256)                                  # shift amount = low byte of src1
        trace[K["SRC0"] + 1:K["SRC0"] + 9] = 0x9E3779B9
        g = O.vm_gadget_cells(orc, trace, 256)
        for s in range(256):
            full = (256 - s) if (variant == 3 and s) else s                    # shifts.rs:61-74: ror by s = rol by 256 - s
            assert int(g[G["SHIFT_FULL"], s]) == full, (name, s)
            limbs = [int(x) for x in g[G["SHIFT_CONSTANT"]:G["SHIFT_CONSTANT"] + 8, s]]
            want = [w for idx in range(4) for w in table[full + (idx << 8)]]   # get_shift_constant: 4 lookups, shifts.rs:200-221
            assert limbs == want, (name, s)


def conditionals_resolution_table():
    """tables/conditional.rs:29-51 by condition NAME (variant_index belongs to zkevm_opcode_defs): (name, flags) -> resolution,
    flags = of | eq << 1 | gt << 2 (integer_into_flags, :7-13)"""
    rows = {}
    for name in ("Always", "Gt", "Lt", "Eq", "Ge", "Le", "Ne", "GtOrLt"):
        for i in range(8):
            of, eq, gt = bool(i & 1), bool(i & 2), bool(i & 4)
            rows[name, i] = int({"Always": True, "Lt": of, "Eq": eq, "Gt": gt, "Ge": gt or eq, "Le": of or eq, "Ne": not eq, "GtOrLt": gt or of}[name])
    return rows


def test_conditional_table_resolves_the_executed_conditions(orc):
    table = conditionals_resolution_table()
    names = {I.COND_ALWAYS: "Always", I.COND_GT: "Gt", I.COND_LT: "Lt", I.COND_EQ: "Eq", I.COND_GE: "Ge", I.COND_LE: "Le", I.COND_NE: "Ne",
             I.COND_GT_OR_LT: "GtOrLt"}
    isa, io, st = fresh(orc)
    for c, name in names.items():
        assert [isa.isa.condition_table[c][f] for f in range(8)] == [table[name, f] for f in range(8)], name
    cycles = 6000
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(I.random_program(isa, 2048, seed=11)), cycles, full=True)
    assert rc == 0
    for k in range(4):
        io.rollback_queue_tail_for_block[k] = int(tail[k])
    res = O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles, cw=cw)
    assert res[0] == 0
    trace, seen = res[2], set()
    for r in range(cycles):
        if int(trace[K["SHOULD_SKIP_CYCLE"], r]) or int(trace[K["PENDING_EXCEPTION_IN"], r]):
            continue
        f = O.vm_state_at(snaps, r).flags
        flags = int(f[0]) | int(f[1]) << 1 | int(f[2]) << 2
        c = int(trace[K["CONDITION_IDX"], r])
        assert int(trace[K["CONDITION"], r]) == table[names[c], flags], (r, c, flags)
        seen.add((c, int(trace[K["CONDITION"], r])))
    assert len({c for c, _ in seen}) == 8 and {v for _, v in seen} == {0, 1}


def integer_to_bitmask(a):
    """tables/integer_to_boolean_mask.rs:33-41: 0 -> no bit, a -> bit a - 1"""
    return 0 if a == 0 else 1 << (a - 1)


def test_register_index_and_sub_pc_bitmasks(orc):
    isa, io, st = fresh(orc)
    vals = [(0xA5A5 << 64) | (1000 + 17 * r) for r in range(1, 16)]
    for r, v in enumerate(vals, start=1):
        set_reg(st, r, v)
    # src1 = r<k> for every index 0..15 (r0 reads as zero: no bit of the mask selects anything), result into r<k> of the next op
    ops = [isa.encode(I.OP_ADD, 0, 0, src=I.MODE_IMM16, src1=k, dst0=0, imm0=0) for k in range(16)]
    # dst0 = r<k>: exactly register k changes (none for index 0)
    ops += [isa.encode(I.OP_ADD, 0, 0, src=I.MODE_IMM16, src1=0, dst0=k, imm0=77 + k) for k in range(16)]
    cycles = len(ops)
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(ops), cycles, full=True)
    assert rc == 0
    io.start_flag = 0
    io.hidden_fsm_input = O.vm_state_at(snaps, 0)
    for k in range(4):
        io.rollback_queue_tail_for_block[k] = int(tail[k])
    res = O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles, cw=cw)
    assert res[0] == 0
    trace = res[2]
    limbs = lambda col, r: sum(int(trace[col + 1 + i, r]) << (32 * i) for i in range(8))
    for k in range(16):
        mask = integer_to_bitmask(k)
        want = sum(v for r, v in enumerate(vals) if (mask >> r) & 1)             # the 15-way select under the one-hot mask
        assert int(trace[K["SRC1_REG"], k]) == k and limbs(K["SRC1"], k) == want, k
    for k in range(16):
        r = 16 + k
        before, after = O.vm_state_at(snaps, r), O.vm_state_at(snaps, r + 1)
        mask = integer_to_bitmask(k)
        assert int(trace[K["DST0_REG"], r]) == k
        for j in range(1, 16):
            changed = reg(before, j) != reg(after, j)
            assert changed == bool((mask >> (j - 1)) & 1), (k, j)
            if changed:
                assert reg(after, j) == 77 + k
    # sub-pc: the 64-bit opcode of cycle i sits in limbs (6 - 2 s, 7 - 2 s) of the code word, s = pc & 3 (pre_state.rs:183-214: the
    # default is the highest pair, bit s - 1 of the sub-pc mask selects the others)
    for r in range(cycles):
        s = int(trace[K["SUB_PC"], r])
        assert s == r % 4 and int(trace[K["SUPER_PC"], r]) == r // 4
        pairs = [(6, 7), (4, 5), (2, 3), (0, 1)]
        sel = pairs[0]
        for bit in range(3):
            if (integer_to_bitmask(s) >> bit) & 1:
                sel = pairs[bit + 1]
        assert [int(trace[K["OPCODE"] + i, r]) for i in range(2)] == [int(trace[K["CODE_WORD"] + sel[0], r]), int(trace[K["CODE_WORD"] + sel[1], r])], r


def uma_ptr_read_cleanup_mask(a):
    """tables/uma_ptr_read_cleanup.rs:26-36: the low `a` bits cleared out of 32"""
    full = (1 << 32) - 1
    return full if a == 0 else full - ((1 << a) - 1)


def test_uma_fat_pointer_read_cleanup_mask(orc):
    """a fat-pointer read whose 32-byte window ends `a` bytes past the slice: the last `a` bytes come back zero -- bit (31 - j) of
    the table's mask keeps big-endian byte j (uma.rs:566-588)"""
    heap_page = 8 + 2
    data = bytes(range(101, 101 + 96))
    for a in range(32):
        isa, io, st = fresh(orc)
        for i in range(3):
            set_reg(st, 10 + i, int.from_bytes(data[32 * i:32 * i + 32], "big"))
        start, offset = 5, 3
        length = offset + 32 - a                                              # the window [offset, offset + 32) overshoots by a
        set_reg(st, 5, offset | (heap_page << 32) | (start << 64) | (length << 96), is_ptr=1)
        ops = []
        for i in range(3):
            ops.append(isa.encode(I.OP_ADD, 0, 0, src=I.MODE_IMM16, src1=0, dst0=2, imm0=32 * i))
            ops.append(isa.encode(I.OP_UMA, I.UMA_HEAP_WRITE, 0, src0=2, src1=10 + i))
        ops.append(isa.encode(I.OP_UMA, I.UMA_PTR_READ, 0, src0=5, dst0=3))
        rc, snaps, wit, status = O.vm_run(orc, isa.isa, st, I.pack_code(ops), len(ops))
        assert rc == 0
        got = reg(O.vm_state_at(snaps, len(ops)), 3).to_bytes(32, "big")
        window = data[start + offset:start + offset + 32]
        mask = uma_ptr_read_cleanup_mask(a)
        want = bytes(b if (mask >> (31 - j)) & 1 else 0 for j, b in enumerate(window))
        assert got == want, (a, got.hex(), want.hex())
        assert want[32 - a:] == bytes(a) and (a == 32 or want[:32 - a] == window[:32 - a])
