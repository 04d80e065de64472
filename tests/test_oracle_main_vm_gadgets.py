"""Gadget cells of main_vm (include/zkc_b200.h, ZKC_VM_GADGET_COLUMNS; oracle/main_vm_gadgets.c): the oblivious results of the
add/sub, binop, mul/div and shift gadgets and the per-cycle relations, re-derived with Python integers from the operand /
property columns of a trace that exercises every opcode family -- an independent statement of
opcodes/{add_sub,binop,mul_div,shifts}.rs and opcodes/mod.rs:101-180."""
import numpy as np

import orc as O
from era_zkevm_circuits_b200 import abi, isa as I

K, G, W = abi.VM_COLS, abi.VMG_COLS, abi.VMG_WIDTHS
M256 = (1 << 256) - 1


def vm_trace(orc, cycles=4000, seed=33):
    isa = I.Isa()
    io = abi.VmClosedForm(); io.start_flag = 1
    st = O.vm_initial_state(orc, io, isa.isa)
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(I.random_program(isa, 1024, seed=seed)), cycles, full=True)
    assert rc == 0
    for k in range(4):
        io.rollback_queue_tail_for_block[k] = int(tail[k])
    want = O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles, cw=cw)
    assert want[0] == 0
    return want[2]


def u256(cols, base, r):
    return sum(int(cols[base + i, r]) << (32 * i) for i in range(8))


def test_gadget_cells_against_python_integers(orc):
    trace = vm_trace(orc)
    cycles = trace.shape[1]
    g = O.vm_gadget_cells(orc, trace, cycles)
    assert g.shape == (G["NUM_COLS"], cycles) and sum(W.values()) == G["NUM_COLS"] == 724
    props = trace[K["PROPS"]]
    bit = lambda r, n: (int(props[r]) >> n) & 1
    seen = set()
    for r in range(cycles):
        a, b = u256(trace, K["SRC0"] + 1, r), u256(trace, K["SRC1"] + 1, r)
        val = lambda name: u256(g, G[name], r)
        one = lambda name, i=0: int(g[G[name] + i, r])
        t = {op: bit(r, getattr(I, "OP_" + op)) for op in ("ADD", "SUB", "MUL", "DIV", "SHIFT", "BINOP")}
        seen |= {k for k, v in t.items() if v}
        # byte views
        assert sum(one("SRC0_BYTES", i) << (8 * i) for i in range(32)) == a and sum(one("SRC1_BYTES", i) << (8 * i) for i in range(32)) == b
        # add_sub.rs
        assert val("ADD_RESULT") == (a + b) & M256 and one("ADD_OF") == (a + b) >> 256
        assert val("SUB_RESULT") == (a - b) & M256 and one("SUB_UF") == int(a < b)
        res = val("ADD_RESULT") if t["ADD"] else val("SUB_RESULT")
        assert val("ADDSUB_RESULT") == res and one("ADDSUB_RESULT_IS_ZERO") == int(res == 0)
        # the shuffled relation holds for BOTH opcodes: a' + b' = c' + of * 2^256 with a' = src1
        assert b + val("ADDSUB_NEW_B") == val("ADDSUB_NEW_C") + (one("ADDSUB_NEW_OF") << 256)
        assert one("ADDSUB_GT") == int(not (one("ADDSUB_NEW_OF") or res == 0)) and one("ADDSUB_APPLY_ANY") == (t["ADD"] | t["SUB"])
        # binop.rs
        assert val("BINOP_AND") == a & b and val("BINOP_OR") == a | b and val("BINOP_XOR") == a ^ b
        for i in (0, 13, 31):
            x, y = (a >> (8 * i)) & 0xFF, (b >> (8 * i)) & 0xFF
            assert one("BINOP_COMPOSITE", i) == (x & y) | ((x | y) << 16) | ((x ^ y) << 32)
            assert [one("BINOP_ALL_RESULTS", 3 * i + j) for j in range(3)] == [x & y, x | y, x ^ y]
        # mul_div.rs
        assert val("MUL_LOW") + (val("MUL_HIGH") << 256) == a * b
        q, rm = (0, a) if b == 0 else divmod(a, b)
        assert val("DIV_QUOTIENT") == q and val("DIV_REMAINDER") == rm
        assert val("DIV_SUB_RESULT") == (rm - b) & M256 and one("DIV_REMAINDER_IS_LESS") == int(rm < b)
        assert one("DIV_DIVISOR_IS_ZERO") == int(b == 0) and one("DIV_MASK_REMAINDER") == (t["DIV"] & int(b == 0))
        assert val("MULDIV_RESULT_1") == (0 if (t["DIV"] and b == 0) else (val("MUL_HIGH") if t["MUL"] else rm))
        # the MulDivRelation of the gadget: a * b + rem = low + 2^256 high in both cases (:273-303)
        assert val("MULDIV_A_TO_ENFORCE") * b + val("MULDIV_REM_TO_ENFORCE") == val("MULDIV_MUL_LOW_TO_ENFORCE") + (val("MULDIV_MUL_HIGH_TO_ENFORCE") << 256)
        # shifts.rs
        sh = b & 0xFF
        is_shr, is_rol, is_ror = bit(r, 16 + 1), bit(r, 16 + 2), bit(r, 16 + 3)  # ZKC_VAR_SHIFT_{SHR, ROL, ROR}
        full = 256 - sh if (is_ror and sh) else sh
        assert one("SHIFT_AMOUNT") == sh and one("SHIFT_FULL") == full and val("SHIFT_CONSTANT") == 1 << full
        assert val("SHIFT_RSHIFT_Q") == a >> full and val("SHIFT_RSHIFT_R") == a & ((1 << full) - 1)
        assert val("SHIFT_LSHIFT_LOW") + (val("SHIFT_LSHIFT_HIGH") << 256) == a << full
        assert val("SHIFT_A_TO_ENFORCE") * (1 << full) + val("SHIFT_REM_TO_ENFORCE") == val("SHIFT_MUL_LOW_TO_ENFORCE") + (val("SHIFT_MUL_HIGH_TO_ENFORCE") << 256)
        right = is_shr and not (is_ror or is_rol)
        want = (a >> full) if right else (((a << full) & M256) + (((a << full) >> 256) if (is_ror or is_rol) else 0))
        assert val("SHIFT_RESULT") == want
        # the relations vm_cycle enforces: selected AddSubRelation with its carry chain, selected MulDivRelation with its partial products
        ra, rb, rc_ = val("ADDREL_A"), val("ADDREL_B"), val("ADDREL_C")
        assert ra + rb == rc_ + (one("ADDREL_OF") << 256)
        carry = 0
        for i in range(8):
            carry = (((ra >> (32 * i)) & 0xFFFFFFFF) + ((rb >> (32 * i)) & 0xFFFFFFFF) + carry) >> 32
            assert one("ADDREL_CARRY", i) == carry
        ma, mb = val("MULREL_A"), val("MULREL_B")
        assert ma * mb + val("MULREL_REM") == val("MULREL_LOW") + (val("MULREL_HIGH") << 256)
        part = [(val("MULREL_REM") >> (32 * i)) & 0xFFFFFFFF for i in range(8)] + [0] * 8
        for ai in range(8):
            ov = 0
            for bi in range(8):
                tt = ((ma >> (32 * ai)) & 0xFFFFFFFF) * ((mb >> (32 * bi)) & 0xFFFFFFFF) + part[ai + bi] + ov
                part[ai + bi], ov = tt & 0xFFFFFFFF, tt >> 32
                assert one("MULREL_PARTIAL_LOW", 8 * ai + bi) == part[ai + bi] and one("MULREL_PARTIAL_HIGH", 8 * ai + bi) == ov
            part[ai + 8] += ov
            assert one("MULREL_ROW_END", ai) == part[ai + 8] < (1 << 32)
        assert sum(p << (32 * i) for i, p in enumerate(part)) == val("MULREL_LOW") + (val("MULREL_HIGH") << 256)
        # which candidate is enforced (cycle.rs:632-668: the last pushed -- shifts -- unless add_sub / mul_div applies)
        if t["MUL"] or t["DIV"]:
            assert (ra, rb, rc_) == (b, val("DIV_SUB_RESULT"), rm) and ma == val("MULDIV_A_TO_ENFORCE") and val("RANGE_CHECK") == val("DIV_SUB_RESULT")
        elif t["ADD"] or t["SUB"]:
            assert (ra, rb, rc_) == (b, val("ADDSUB_NEW_B"), val("ADDSUB_NEW_C")) and val("RANGE_CHECK") == res
        else:
            assert (ra, rb, rc_) == (1 << full, val("SHIFT_SUB_RESULT"), val("SHIFT_RSHIFT_R")) and val("RANGE_CHECK") == val("SHIFT_SUB_RESULT")
        if not (t["MUL"] or t["DIV"]):
            assert ma == val("SHIFT_A_TO_ENFORCE") and mb == 1 << full
        # the selected path of the DENSE trace is the gadget's own result
        d0 = u256(trace, K["DST0"] + 1, r)
        if t["ADD"] or t["SUB"]:
            assert d0 == res
        if t["BINOP"]:
            assert d0 == val("BINOP_RESULT")
        if t["MUL"] or t["DIV"]:
            assert d0 == val("MULDIV_RESULT_0") and u256(trace, K["DST1"] + 1, r) == val("MULDIV_RESULT_1")
        if t["SHIFT"]:
            assert d0 == val("SHIFT_RESULT")
    assert seen == {"ADD", "SUB", "MUL", "DIV", "SHIFT", "BINOP"}
