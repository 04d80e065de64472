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


# ---- the second block: ptr / jump / context (ZKC_VM_STATE_GADGET_COLUMNS; oracle/main_vm_gadgets.c state_gadget_row) ------------
S, SW = abi.VMS_COLS, abi.VMS_WIDTHS


def vm_trace_and_snapshots(orc, cycles=4000, seed=33, far=False):
    isa = I.Isa()
    io = abi.VmClosedForm(); io.start_flag = 1
    st = O.vm_initial_state(orc, io, isa.isa)
    rc, snaps, wit, status, cw, tail = O.vm_run(orc, isa.isa, st, I.pack_code(I.random_program(isa, 1024, seed=seed, far_calls=far)), cycles, full=True)
    assert rc == 0
    for k in range(4):
        io.rollback_queue_tail_for_block[k] = int(tail[k])
    want = O.vm_entry_point(orc, io, isa.isa, snaps, wit, cycles, cw=cw)
    assert want[0] == 0
    return want[2], snaps


def state_gadget_reference(trace, snaps, r):
    """apply_ptr / apply_jump / apply_context on Python integers (256-bit operands as ONE integer, not limbs): name -> value(s)"""
    props = int(trace[K["PROPS"], r])
    bit = lambda n: (props >> n) & 1
    var = lambda v: bit(16 + v)                                            # ZKC_VM_BIT_VARIANT
    a, b = u256(trace, K["SRC0"] + 1, r), u256(trace, K["SRC1"] + 1, r)
    a_ptr, b_ptr = int(trace[K["SRC0"], r]), int(trace[K["SRC1"], r])
    M32 = 0xFFFFFFFF
    out = {}
    # ptr.rs
    apply = bit(I.OP_PTR)
    v_add, v_sub, v_pack, v_shrink = var(0), var(1), var(2), var(3)
    offset = b & M32
    out["PTR_SRC1_IS_INTEGER"] = 1 - b_ptr
    out["PTR_ARGS_VALID"] = int(a_ptr and not b_ptr); out["PTR_ARGS_INVALID"] = 1 - out["PTR_ARGS_VALID"]
    out["PTR_SRC1_LIMB_IS_ZERO"] = [int((b >> (32 * i)) & M32 == 0) for i in range(8)]
    out["PTR_SRC1_32_256_IS_ZERO"] = int(b >> 32 == 0); out["PTR_SRC1_32_256_IS_NONZERO"] = int(b >> 32 != 0)
    out["PTR_SRC1_0_128_IS_ZERO"] = int(b & ((1 << 128) - 1) == 0); out["PTR_SRC1_0_128_IS_NONZERO"] = 1 - out["PTR_SRC1_0_128_IS_ZERO"]
    out["PTR_ARITH_VARIANT"] = v_add | v_sub
    out["PTR_TOO_LARGE_OFFSET"] = int(b >> 32 != 0) & (v_add | v_sub)
    out["PTR_DIRTY_PACK"] = out["PTR_SRC1_0_128_IS_NONZERO"] & v_pack
    ptr_offset, ptr_len = a & M32, (a >> 96) & M32                        # FatPointer: offset | page | start | length
    out["PTR_ADD_RESULT"] = (ptr_offset + offset) & M32; out["PTR_ADD_OF"] = (ptr_offset + offset) >> 32
    out["PTR_SUB_RESULT"] = (ptr_offset - offset) & M32; out["PTR_SUB_UF"] = int(ptr_offset < offset)
    out["PTR_SHRINK_RESULT"] = (ptr_len - offset) & M32; out["PTR_SHRINK_UF"] = int(ptr_len < offset)
    out["PTR_ADD_PANIC"] = v_add & out["PTR_ADD_OF"]; out["PTR_SUB_PANIC"] = v_sub & out["PTR_SUB_UF"]; out["PTR_SHRINK_PANIC"] = v_shrink & out["PTR_SHRINK_UF"]
    panic = int(any(out[k] for k in ("PTR_ARGS_INVALID", "PTR_TOO_LARGE_OFFSET", "PTR_DIRTY_PACK", "PTR_ADD_PANIC", "PTR_SUB_PANIC", "PTR_SHRINK_PANIC")))
    out["PTR_ANY_PANIC"] = panic; out["PTR_SHOULD_PANIC"] = apply & panic; out["PTR_OK"] = 1 - panic; out["PTR_UPDATE_REGISTER"] = apply & (1 - panic)
    out["PTR_LOW_IF_ADD"] = out["PTR_ADD_RESULT"] if v_add else ptr_offset
    out["PTR_LOW_IF_ADD_OR_SUB"] = out["PTR_SUB_RESULT"] if v_sub else out["PTR_LOW_IF_ADD"]
    out["PTR_96_128_IF_SHRINK"] = out["PTR_SHRINK_RESULT"] if v_shrink else ptr_len
    hi = (b if v_pack else a) >> 128
    out["PTR_HIGHEST_128"] = [(hi >> (32 * i)) & M32 for i in range(4)]
    out["PTR_LOWEST32"] = ptr_offset if v_pack else out["PTR_LOW_IF_ADD_OR_SUB"]
    out["PTR_96_128"] = ptr_len if v_pack else out["PTR_96_128_IF_SHRINK"]
    dst = out["PTR_LOWEST32"] | (a & (((1 << 64) - 1) << 32)) | (out["PTR_96_128"] << 96) | (hi << 128)
    out["PTR_DST0"] = [a_ptr] + [(dst >> (32 * i)) & M32 for i in range(8)]
    # jump.rs
    out["JUMP_DST"] = a & 0xFFFF
    # context.rs
    st = abi.VmState.from_buffer_copy(snaps[r].tobytes())
    c = st.current_context
    apply = bit(I.OP_CONTEXT)
    is_set = var(7) | var(8) | var(9)
    out["CTX_WRITE_TO_CONTEXT"] = apply & var(7); out["CTX_SET_PUBDATA_ERGS"] = apply & var(8); out["CTX_INCREMENT_TX"] = apply & var(9)
    out["CTX_READ_ONLY"] = is_set; out["CTX_WRITE_LIKE"] = 1 - is_set; out["CTX_WRITE_TO_DST0"] = apply & (1 - is_set)
    out["CTX_INCREMENTED_TX_NUMBER"] = (st.tx_number_in_block + 1) & M32; out["CTX_TX_OF"] = (st.tx_number_in_block + 1) >> 32
    meta_hi = c.this_shard_id | c.caller_shard_id << 8 | c.code_shard_id << 16
    out["CTX_META_HIGHEST"] = meta_hi
    low = int(trace[K["DIRTY_ERGS_LEFT"], r]) if var(4) else int(trace[K["NEW_SP"], r])
    out["CTX_LOW_U32"] = low
    word = lambda limbs: sum(int(x) << (32 * i) for i, x in enumerate(limbs))
    limbs = lambda x, n: [(x >> (32 * i)) & M32 for i in range(n)]
    res = word(c.context_u128_value_composite) if var(6) else low
    out["CTX_RESULT_128"] = limbs(res, 4)
    res = word(c.this_address) if var(0) else res
    out["CTX_RESULT_160_THIS"] = limbs(res, 5)
    res = word(c.caller) if var(1) else res
    out["CTX_RESULT_160_CALLER"] = limbs(res, 5)
    res = word(c.code_address) if var(2) else res
    out["CTX_RESULT_160_CODE"] = limbs(res, 5)
    if var(3):
        res = st.ergs_per_pubdata_byte | c.heap_upper_bound << 64 | c.aux_heap_upper_bound << 96 | meta_hi << 224
    out["CTX_RESULT_256"] = limbs(res, 8)
    return out


def test_state_gadget_cells_against_python_integers(orc):
    assert sum(SW.values()) == S["NUM_COLS"] == 87
    seen_ops, seen_ctx, seen_ptr = set(), set(), 0
    for far, seed in ((False, 33), (True, 5)):
        trace, snaps = vm_trace_and_snapshots(orc, 3000, seed, far)
        cycles = trace.shape[1]
        g = O.vm_state_gadget_cells(orc, trace, snaps, cycles)
        assert g.shape == (S["NUM_COLS"], cycles)
        for r in range(cycles):
            want = state_gadget_reference(trace, snaps, r)
            assert set(want) == set(SW)
            for name, v in want.items():
                got = [int(x) for x in g[S[name]:S[name] + SW[name], r]]
                assert got == (v if isinstance(v, list) else [v]), (r, name, got, v)
            props = int(trace[K["PROPS"], r])
            d0 = [int(x) for x in trace[K["DST0"]:K["DST0"] + 9, r]]
            # the selected path of the DENSE trace is the gadget's own candidate
            if (props >> I.OP_PTR) & 1:
                seen_ops.add("PTR")
                if want["PTR_UPDATE_REGISTER"]:
                    seen_ptr += 1
                    assert d0 == want["PTR_DST0"], (r, d0, want["PTR_DST0"])
            if (props >> I.OP_CONTEXT) & 1:
                seen_ops.add("CONTEXT")
                seen_ctx |= {v for v in range(10) if (props >> (16 + v)) & 1}
                if want["CTX_WRITE_TO_DST0"]:
                    assert d0 == [0] + want["CTX_RESULT_256"], (r, d0, want["CTX_RESULT_256"])
            if (props >> I.OP_JUMP) & 1:
                seen_ops.add("JUMP")
                assert int(trace[K["PC_OUT"], r]) == want["JUMP_DST"]
    assert seen_ops == {"PTR", "CONTEXT", "JUMP"} and len(seen_ctx) >= 8 and seen_ptr > 0, (seen_ops, seen_ctx, seen_ptr)


# ---- the memory-queue relations of every cycle (ZKC_VM_MEMORY_SPONGE_COLUMNS; oracle/main_vm_gadgets.c memq_step) -------------
Q, QW = abi.VMQ_COLS, abi.VMQ_WIDTHS


def memory_query_encoding(ts, page, index, rw, is_ptr, value):
    """MemoryQuery::encode (base_structures/memory_query/mod.rs:103-221) from the 256-bit value as ONE integer: bytes of limbs 5-7
    spread over the high halves of the elements that carry limbs 0-3"""
    limb = lambda i: (value >> (32 * i)) & 0xFFFFFFFF
    byte = lambda i, k: (limb(i) >> (8 * k)) & 0xFF
    return [ts, page, index | rw << 32 | is_ptr << 33,
            limb(0) | byte(5, 0) << 32 | byte(5, 1) << 40 | byte(5, 2) << 48,
            limb(1) | byte(5, 3) << 32 | byte(6, 0) << 40 | byte(6, 1) << 48,
            limb(2) | byte(6, 2) << 32 | byte(6, 3) << 40 | byte(7, 0) << 48,
            limb(3) | byte(7, 1) << 32 | byte(7, 2) << 40 | byte(7, 3) << 48,
            limb(4)]


def test_memory_sponge_cells_against_python_and_the_dense_trace(orc):
    import ref_poseidon2 as R
    assert sum(QW.values()) == Q["NUM_COLS"] == 112
    rc_ = R.constants()
    own = (1 << I.OP_UMA) | (1 << I.OP_LOG) | (1 << I.OP_NEAR_CALL) | (1 << I.OP_FAR_CALL) | (1 << I.OP_RET)
    seen = dict(fetch=0, fetch_skipped=0, src0=0, dst0=0, not_selected=0)
    for far, seed, cycles in ((False, 33, 2500), (True, 5, 2500)):
        trace, snaps = vm_trace_and_snapshots(orc, cycles, seed, far)
        g = O.vm_memory_sponge_cells(orc, trace, snaps, cycles)
        assert g.shape == (Q["NUM_COLS"], cycles)
        col = lambda name, r, n=12: [int(x) for x in g[Q[name]:Q[name] + n, r]]
        for r in range(cycles):
            st, nxt = O.vm_state_at(snaps, r), O.vm_state_at(snaps, r + 1)
            props = int(trace[K["PROPS"], r])
            selected = int(props & own == 0)
            assert int(g[Q["SELECTED"], r]) == selected
            seen["not_selected"] += 1 - selected
            state, length = [int(x) for x in st.memory_queue_state], st.memory_queue_length
            steps = (
                ("FETCH", 0, int(trace[K["SHOULD_READ_OPCODE"], r]),
                 (st.timestamp, st.current_context.code_page, int(trace[K["SUPER_PC"], r]), 0, 0,
                  u256(trace, K["CODE_WORD"], r) if int(trace[K["SHOULD_READ_OPCODE"], r]) else 0)),
                ("SRC0", 1, int(trace[K["SHOULD_READ_SRC0"], r]),
                 (st.timestamp, int(trace[K["SRC0_PAGE"], r]), int(trace[K["SRC0_INDEX"], r]), 0, int(trace[K["SRC0_FROM_MEMORY"], r]),
                  u256(trace, K["SRC0_FROM_MEMORY"] + 1, r))),
                ("DST0", 2, int(trace[K["PERFORM_DST0_MEMORY_WRITE"], r]),
                 (st.timestamp + 3, int(trace[K["DST0_PAGE"], r]), int(trace[K["DST0_INDEX"], r]), 1, int(trace[K["DST0"], r]),
                  u256(trace, K["DST0"] + 1, r))),
            )
            for name, slot, execute, query in steps:
                init = memory_query_encoding(*query) + state[8:]
                assert col(name + "_INIT", r) == init, (r, name)
                final = col(name + "_FINAL", r)
                # the permutation itself: the second Python restatement on a sample of rows, the DENSE trace's own slot wherever
                # the relation is enforced there (every executed access of a cycle without an opcode that brings its own sponges)
                if r % 40 == slot:
                    assert R.permutation(init, rc_) == final, (r, name)
                enforced = int(trace[K["SPONGE_ENFORCE"] + slot, r])
                if slot == 0 or selected:
                    assert enforced == execute, (r, name, enforced, execute)
                    if enforced:
                        assert [int(x) for x in trace[K["SPONGE_FINAL"] + 12 * slot:K["SPONGE_FINAL"] + 12 * slot + 12, r]] == final, (r, name)
                if execute:
                    state, length = final, length + 1
                    seen[name.lower()] += 1
                elif slot == 0:
                    seen["fetch_skipped"] += 1
                assert col(name + "_STATE_AFTER", r) == state and int(g[Q[name + "_LENGTH_AFTER"], r]) == length, (r, name)
            if selected:  # no opcode touches the memory queue afterwards: this is the next cycle's queue
                assert state == [int(x) for x in nxt.memory_queue_state] and length == nxt.memory_queue_length == int(trace[K["MEMQ_LENGTH_OUT"], r]), r
    assert min(seen.values()) > 0, seen


# ---- the create_prestate block (ZKC_VM_PRESTATE_COLUMNS; oracle/main_vm_gadgets.c orc_main_vm_prestate_cells) ---------------------
P, PW = abi.VMP_COLS, abi.VMP_WIDTHS
SRC_MODE = lambda m: I.TYPE_BITS + I.VARIANT_BITS + I.FLAG_BITS + m
DST_MODE = lambda m: I.TYPE_BITS + I.VARIANT_BITS + I.FLAG_BITS + I.SRC_BITS + m
FLAG = lambda f: I.TYPE_BITS + I.VARIANT_BITS + f


def prestate_reference(trace, snaps, r):
    """create_prestate (pre_state.rs:71-519, utils.rs:106-120, :237-386) on Python integers: a register is (is_pointer, ONE 256-bit
    integer); name -> value(s) in column order"""
    col = lambda name, i=0: int(trace[K[name] + i, r])
    props = col("PROPS")
    bit = lambda n: (props >> n) & 1
    st = abi.VmState.from_buffer_copy(snaps[r].tobytes())
    c = st.current_context
    M16, M32 = 0xFFFF, 0xFFFFFFFF
    word = lambda limbs: sum(int(x) << (32 * i) for i, x in enumerate(limbs))
    cells = lambda reg: [reg[0]] + [(reg[1] >> (32 * i)) & M32 for i in range(8)]
    regs = [(int(x.is_pointer) & 1, word(x.value)) for x in st.registers]
    out = {}
    skip, pending = col("SHOULD_SKIP_CYCLE"), col("PENDING_EXCEPTION_IN")
    out["EXECUTE_CYCLE"] = 1 - skip
    out["SHOULD_TRY_TO_READ_OPCODE"] = int(not skip and not pending)
    out["PENDING_EXCEPTION_TAKEN_DOWN"] = 0
    out["PC_PLUS_ONE"], out["PC_PLUS_ONE_OF"] = (c.pc + 1) % 65536, (c.pc + 1) // 65536
    out["CODE_PAGES_ARE_EQUAL"] = int(st.previous_code_page == c.code_page)
    out["SUPER_PC_ARE_EQUAL"] = int(c.pc >> 2 == st.previous_super_pc)             # split_pc: super_pc = pc >> 2, sub_pc = pc & 3
    out["CAN_SKIP_READ"] = out["CODE_PAGES_ARE_EQUAL"] & out["SUPER_PC_ARE_EQUAL"]
    out["SHOULD_READ_FOR_NEW_PC"] = 1 - out["CAN_SKIP_READ"]
    out["TIMESTAMPS"] = [(st.timestamp + k) & M32 for k in (1, 2, 3, 4)]
    out["NEXT_CYCLE_TIMESTAMP"] = st.timestamp if skip else (st.timestamp + 4) & M32
    code_word = word(col("CODE_WORD", i) for i in range(8))
    sub_pc = c.pc & 3
    out["SUBPC_BITMASK"] = [int(sub_pc == k) for k in (1, 2, 3)]
    # the opcode is the sub_pc-th 64-bit lane counted from the TOP of the big-endian word; the chain holds the running selection
    chain, sel = [], code_word >> 192
    for k in (1, 2, 3):
        if sub_pc == k:
            sel = (code_word >> (192 - 64 * k)) & (2 ** 64 - 1)
        chain += [sel & M32, sel >> 32]
    out["OPCODE_SELECT_CHAIN"] = chain
    idx = [col("SRC0_REG"), col("SRC1_REG"), col("DST0_REG"), col("DST1_REG")]
    for name, i in zip(("SRC0_SELECTORS", "SRC1_SELECTORS", "DST0_SELECTORS", "DST1_SELECTORS"), idx):
        out[name] = [int(i == k + 1) for k in range(15)]
    pick = lambda i, k: regs[i - 1] if 1 <= i <= k + 1 else (0, 0)                  # after k + 1 steps of the chain
    out["DRAFT_SRC0_CHAIN"] = sum((cells(pick(idx[0], k)) for k in range(15)), [])
    out["SRC1_REGISTER_CHAIN"] = sum((cells(pick(idx[1], k)) for k in range(15)), [])
    out["DST0_REG_LOW_CHAIN"] = [pick(idx[2], k)[1] & M32 for k in range(15)]
    draft_src0, src1_reg = pick(idx[0], 14), pick(idx[1], 14)
    out["SRC0_REG_LOWEST"], out["DST0_REG_LOWEST"] = draft_src0[1] & M16, pick(idx[2], 14)[1] & M16
    out["STACK_PAGE"], out["HEAP_PAGE"], out["AUX_HEAP_PAGE"] = [(c.base_page + k) & M32 for k in (1, 2, 3)]
    s_code, s_abs, s_rel, s_pp = (bit(SRC_MODE(m)) for m in (I.MODE_CODE, I.MODE_STACK_ABS, I.MODE_STACK_OFFSET, I.MODE_PUSH_POP))
    src_abs = (out["SRC0_REG_LOWEST"] + col("IMM0")) & M16
    src_rel = (c.sp - src_abs) & M16
    out["SRC_ABSOLUTE_MODE"] = s_code | s_abs
    out["SRC_INDEX_FOR_ABSOLUTE"], out["SRC_INDEX_FOR_RELATIVE"] = src_abs, src_rel
    out["SRC_USE_STACK"] = s_abs | s_rel | s_pp
    out["SRC_DID_READ_UNMASKED"] = out["SRC_USE_STACK"] | s_code
    out["NOT_NOP"] = 1 - bit(I.OP_NOP)
    sp1 = src_rel if s_pp else c.sp & M16
    d_abs, d_rel, d_pp = (bit(DST_MODE(m)) for m in (I.MODE_STACK_ABS, I.MODE_STACK_OFFSET, I.MODE_PUSH_POP))
    dst_abs = (out["DST0_REG_LOWEST"] + col("IMM1")) & M16
    out["DST_INDEX_FOR_ABSOLUTE"] = dst_abs
    out["DST_INDEX_FOR_RELATIVE_WITH_PUSH"], out["DST_INDEX_FOR_RELATIVE"] = (sp1 + dst_abs) & M16, (sp1 - dst_abs) & M16
    out["DST_DID_WRITE_UNMASKED"] = d_abs | d_rel | d_pp
    out["DST_INDEX_SOMEWHAT_RELATIVE"] = sp1 if d_pp else out["DST_INDEX_FOR_RELATIVE"]
    from_memory = (col("SRC0_FROM_MEMORY") & 1, word(col("SRC0_FROM_MEMORY", 1 + i) for i in range(8)))
    src0 = draft_src0 if bit(SRC_MODE(I.MODE_REG)) else from_memory
    out["SRC0_AFTER_USE_REG"] = cells(src0)
    src0 = (0, col("IMM0")) if bit(SRC_MODE(I.MODE_IMM16)) else src0
    out["SRC0_AFTER_USE_IMM"] = cells(src0)
    asym = bit(I.OP_SUB) | bit(I.OP_DIV) | bit(I.OP_SHIFT)
    out["SWAP_IS_ASSYMMETRIC"], out["SWAP_T0"], out["SWAP_T1"] = asym, asym & bit(FLAG(1)), bit(I.OP_PTR) & bit(FLAG(0))
    swap = out["SWAP_T0"] | out["SWAP_T1"]
    a, b = (src1_reg, src0) if swap else (src0, src1_reg)
    out["SRC0_SWAPPED"], out["SRC1_SWAPPED"] = cells(a), cells(b)
    keeps = bit(I.OP_RET) | bit(I.OP_PTR) | bit(I.OP_UMA) | bit(I.OP_FAR_CALL)
    out["NOT_KERNEL_MODE"], out["KEEPS_POINTERS"], out["SHOULD_ERASE"] = 1 - (c.is_kernel_mode & 1), keeps, 1 - keeps
    out["SHOULD_ERASE_SRC0"] = a[0] & (1 - keeps) & out["NOT_KERNEL_MODE"]
    out["SHOULD_ERASE_SRC1"] = b[0] & out["NOT_KERNEL_MODE"]
    return out, swap, sp1


def erased(cells9, flag):
    """VMRegister::conditionally_erase_fat_pointer_data (base_structures/register/mod.rs:67-76): is_pointer and limbs 1, 2"""
    return [0 if flag and i in (0, 2, 3) else v for i, v in enumerate(cells9)]


def test_prestate_cells_against_python_integers_and_the_dense_trace(orc):
    assert sum(PW.values()) == P["NUM_COLS"] == 428
    seen = {"swap": 0, "erase0": 0, "erase1": 0, "push": 0, "pop": 0, "imm": 0, "mem": 0, "skip_read": 0, "sub_pc": set(), "masked": 0}
    for far, seed in ((False, 33), (True, 5)):
        trace, snaps = vm_trace_and_snapshots(orc, 3000, seed, far)
        cycles = trace.shape[1]
        g = O.vm_prestate_cells(orc, trace, snaps, cycles)
        assert g.shape == (P["NUM_COLS"], cycles)
        for r in range(cycles):
            want, swap, sp1 = prestate_reference(trace, snaps, r)
            assert list(want) != [] and set(want) == set(PW)
            for name, v in want.items():
                got = [int(x) for x in g[P[name]:P[name] + PW[name], r]]
                assert got == (v if isinstance(v, list) else [v]), (r, name, got, v)
            col = lambda name, i=0: int(trace[K[name] + i, r])
            # the block's cells lead to the RESULTS the DENSE trace names
            assert col("SWAP_OPERANDS") == swap
            assert col("SHOULD_READ_OPCODE") == want["SHOULD_TRY_TO_READ_OPCODE"] & want["SHOULD_READ_FOR_NEW_PC"]
            assert col("SUPER_PC") == abi.VmState.from_buffer_copy(snaps[r].tobytes()).current_context.pc >> 2
            if not col("SHOULD_SKIP_CYCLE") and not col("PENDING_EXCEPTION_IN"):   # else mask_into_nop / mask_into_panic, :216-221
                assert [col("OPCODE"), col("OPCODE", 1)] == want["OPCODE_SELECT_CHAIN"][4:], r
            else:
                seen["masked"] += 1
            assert col("SRC0_PAGE") == (want["STACK_PAGE"] if want["SRC_USE_STACK"] else
                                        abi.VmState.from_buffer_copy(snaps[r].tobytes()).current_context.code_page)
            assert col("SRC0_INDEX") == (want["SRC_INDEX_FOR_ABSOLUTE"] if want["SRC_ABSOLUTE_MODE"] else want["SRC_INDEX_FOR_RELATIVE"])
            assert col("SHOULD_READ_SRC0") == want["SRC_DID_READ_UNMASKED"] & want["NOT_NOP"]
            assert col("SP_AFTER_SRC0") == sp1
            props = col("PROPS")
            d_abs, d_pp = (props >> DST_MODE(I.MODE_STACK_ABS)) & 1, (props >> DST_MODE(I.MODE_PUSH_POP)) & 1
            assert col("DST0_PAGE") == want["STACK_PAGE"]
            assert col("DST0_INDEX") == (want["DST_INDEX_FOR_ABSOLUTE"] if d_abs else want["DST_INDEX_SOMEWHAT_RELATIVE"])
            assert col("DST0_PERFORMS_MEMORY_ACCESS") == want["DST_DID_WRITE_UNMASKED"] & want["NOT_NOP"]
            assert col("NEW_SP") == (want["DST_INDEX_FOR_RELATIVE_WITH_PUSH"] if d_pp else sp1)
            assert [col("SRC0", i) for i in range(9)] == erased(want["SRC0_SWAPPED"], want["SHOULD_ERASE_SRC0"]), r
            assert [col("SRC1", i) for i in range(9)] == erased(want["SRC1_SWAPPED"], want["SHOULD_ERASE_SRC1"]), r
            seen["swap"] += swap; seen["erase0"] += want["SHOULD_ERASE_SRC0"]; seen["erase1"] += want["SHOULD_ERASE_SRC1"]
            seen["push"] += d_pp; seen["pop"] += (props >> SRC_MODE(I.MODE_PUSH_POP)) & 1
            seen["imm"] += (props >> SRC_MODE(I.MODE_IMM16)) & 1
            seen["mem"] += col("SHOULD_READ_SRC0"); seen["skip_read"] += want["CAN_SKIP_READ"]
            seen["sub_pc"].add(col("SUB_PC"))
    # every random program runs in kernel mode, to the end of its cycles and never pops a source operand: those are the mutation test's
    assert seen["sub_pc"] == {0, 1, 2, 3} and all(v for k, v in seen.items() if k not in ("erase0", "erase1", "masked", "pop")), str(seen)


def mutated_prestate_inputs(orc, cycles=1500, seed=7):
    """a real trace + snapshots with the block's INPUTS redrawn at random (the block is a pure function of them): user mode, pointer
    registers, every register index, 16-bit wrap of pc / sp / indices, 32-bit wrap of timestamp / pages, skipped and pending cycles"""
    trace, snaps = vm_trace_and_snapshots(orc, cycles, seed, True)
    trace, snaps = trace.copy(), snaps[:cycles + 1].copy()
    rng = np.random.default_rng(seed)
    st = snaps.view(abi.VM_STATE_DTYPE).reshape(-1) if hasattr(abi, "VM_STATE_DTYPE") and snaps.dtype != abi.VM_STATE_DTYPE else snaps.reshape(-1)
    n = cycles
    edge16 = lambda: rng.choice([0, 1, 0xFFFF, 0xFFFE], n) * (rng.random(n) < 0.3) + rng.integers(0, 1 << 16, n) * (rng.random(n) < 0.7)
    words = np.frombuffer(st.tobytes(), dtype=np.uint32).reshape(len(st), -1).copy()
    V = abi.VmState
    off = lambda *path: sum(getattr(t, f).offset for t, f in path) // 4
    C_ = type(V().current_context)
    ctx = V.current_context.offset // 4
    for r in range(15):
        words[:n, V.registers.offset // 4 + 9 * r] = rng.integers(0, 2, n)
    for f, vals in (("pc", edge16() & 0xFFFF), ("sp", edge16() & 0xFFFF), ("is_kernel_mode", rng.integers(0, 2, n)),
                    ("base_page", np.where(rng.random(n) < 0.2, 0xFFFFFFFF - rng.integers(0, 3, n), rng.integers(0, 1 << 32, n)))):
        words[:n, ctx + getattr(C_, f).offset // 4] = vals
    words[:n, V.timestamp.offset // 4] = np.where(rng.random(n) < 0.2, 0xFFFFFFFF - rng.integers(0, 4, n), rng.integers(0, 1 << 32, n))
    pc = words[:n, ctx + C_.pc.offset // 4]
    same = rng.random(n) < 0.5
    words[:n, V.previous_code_page.offset // 4] = np.where(same, words[:n, ctx + C_.code_page.offset // 4], rng.integers(0, 1 << 32, n))
    words[:n, V.previous_super_pc.offset // 4] = np.where(rng.random(n) < 0.5, pc >> 2, rng.integers(0, 1 << 14, n))
    snaps = np.frombuffer(words.tobytes(), dtype=snaps.dtype).reshape(snaps.shape)
    trace[K["SUPER_PC"]], trace[K["SUB_PC"]] = pc >> 2, pc & 3
    for name in ("SRC0_REG", "SRC1_REG", "DST0_REG", "DST1_REG"):
        trace[K[name]] = rng.integers(0, 16, n)
    trace[K["IMM0"]], trace[K["IMM1"]] = edge16() & 0xFFFF, edge16() & 0xFFFF
    trace[K["SHOULD_SKIP_CYCLE"]], trace[K["PENDING_EXCEPTION_IN"]] = rng.random(n) < 0.2, rng.random(n) < 0.2
    trace[K["SRC0_FROM_MEMORY"]] = rng.integers(0, 2, n)
    for i in range(8):
        trace[K["SRC0_FROM_MEMORY"] + 1 + i] = rng.integers(0, 1 << 32, n)
        trace[K["CODE_WORD"] + i] = rng.integers(0, 1 << 32, n)
    props = rng.integers(0, 1 << 38, n).astype(np.uint64)                    # any bit pattern: every cell is bitwise in the property bits
    trace[K["PROPS"]] = props
    pb = lambda k: (props >> np.uint64(k)) & np.uint64(1)
    trace[K["SWAP_OPERANDS"]] = ((pb(I.OP_SUB) | pb(I.OP_DIV) | pb(I.OP_SHIFT)) & pb(FLAG(1))) | (pb(I.OP_PTR) & pb(FLAG(0)))
    return trace, snaps


def test_prestate_cells_on_mutated_inputs(orc):
    trace, snaps = mutated_prestate_inputs(orc)
    cycles = trace.shape[1]
    g = O.vm_prestate_cells(orc, trace, snaps, cycles)
    erase = [0, 0]
    for r in range(cycles):
        want, _, _ = prestate_reference(trace, snaps, r)
        for name, v in want.items():
            got = [int(x) for x in g[P[name]:P[name] + PW[name], r]]
            assert got == (v if isinstance(v, list) else [v]), (r, name, got, v)
        erase[0] += want["SHOULD_ERASE_SRC0"]; erase[1] += want["SHOULD_ERASE_SRC1"]
    assert min(erase) > 5, erase
    assert int(g[P["PC_PLUS_ONE_OF"]].sum()) > 0 and int((g[P["NEXT_CYCLE_TIMESTAMP"]] < 4).sum()) > 0   # both wraps happened


# ---- the register write-back block (ZKC_VM_WRITEBACK_COLUMNS; oracle/main_vm_gadgets.c orc_main_vm_writeback_cells) -----------------
WB, WBW = abi.VMW_COLS, abi.VMW_WIDTHS


def writeback_reference(isa, trace, snaps, r):
    """cycle.rs:158-433 on Python integers: a register is [is_pointer, one 256-bit integer]; each register's chain as a list of
    (flag, candidate) pairs applied in the reference's order; name -> value(s) in column order"""
    col = lambda name, i=0: int(trace[K[name] + i, r])
    props = col("PROPS")
    bit = lambda n: (props >> n) & 1
    st = abi.VmState.from_buffer_copy(snaps[r].tobytes())
    nx = abi.VmState.from_buffer_copy(snaps[r + 1].tobytes())
    word = lambda limbs: sum(int(x) << (32 * i) for i, x in enumerate(limbs))
    limbs = lambda x: [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
    capable = any(bit(o) for o in (I.OP_ADD, I.OP_SUB, I.OP_MUL, I.OP_DIV, I.OP_BINOP, I.OP_SHIFT, I.OP_PTR))
    upd, memw, mem_access = col("DST0_UPDATE_REGISTER"), col("PERFORM_DST0_MEMORY_WRITE"), col("DST0_PERFORMS_MEMORY_ACCESS")
    out = {"DST0_UPDATE_POTENTIALLY_TO_MEMORY": int(capable and (upd or memw)), "CAN_UPDATE_DST0_AS_REGISTER_ONLY": int(not capable and upd),
           "DST0_PERFORMS_REG_UPDATE": 1 - mem_access}
    out["DST0_REG_UPDATE_T"] = out["DST0_PERFORMS_REG_UPDATE"] & out["DST0_UPDATE_POTENTIALLY_TO_MEMORY"]
    assert upd == out["CAN_UPDATE_DST0_AS_REGISTER_ONLY"] | out["DST0_REG_UPDATE_T"]                      # cycle.rs:310
    abi_word, target = word(col("SRC0", 1 + i) for i in range(8)), word(col("SRC1", 1 + i) for i in range(8))
    far_call = bit(I.OP_FAR_CALL)
    target_is_kernel = (target & ((1 << 160) - 1)) >> 16 == 0
    system = int((abi_word >> 248) & 0xFF != 0 and target_is_kernel)
    constructor = int((abi_word >> 240) & 0xFF != 0 and st.current_context.is_kernel_mode & 1)
    far_return = bit(I.OP_RET) & (1 - (st.current_context.is_local_call & 1))
    cleanup = far_call & (1 - system)
    out.update(FAR_CALL_UPDATE=far_call, FAR_CALL_NON_SYSTEM=1 - system, FAR_CALL_CLEANUP_REGISTER=cleanup, FAR_RETURN_UPDATE=far_return,
               FAR_CALL_NEW_R2_LOW=constructor + 2 * system)
    dst0 = [col("DST0") & 1, word(col("DST0", 1 + i) for i in range(8))]
    dst1 = [col("DST1") & 1, word(col("DST1", 1 + i) for i in range(8))]
    abi_regs = set(range(isa.call_system_abi_registers[0], isa.call_system_abi_registers[1]))
    reserved = set(range(isa.call_reserved_range[0], isa.call_reserved_range[1])) | {isa.call_implicit_parameter_reg_idx}
    per = {k: [] for k in WBW if WBW[k] >= 15 or k in ("VALUE_AFTER_FAR_RETURN",)}
    for k in range(15):
        reg = [st.registers[k].is_pointer & 1, word(st.registers[k].value)]
        new_r1 = [nx.registers[0].is_pointer & 1, word(nx.registers[0].value)]
        write0, write1 = int(upd and col("DST0_REG") == k + 1), int(col("DST1_REG") == k + 1)
        ptr_side = [(write0, dst0[0])]                                   # (flag, is_pointer candidate), cycle.rs:339-362
        val_side = [(write0, dst0[1], "VALUE_AFTER_DST0")]               # (flag, value candidate, column after the step), :343-375, :415-433
        if k == 0:
            ptr_side += [(far_call, new_r1[0]), (far_return, new_r1[0])]
            val_side += [(far_call, new_r1[1], "VALUE_AFTER_FAR_CALL"), (far_return, new_r1[1], "VALUE_AFTER_FAR_RETURN")]
        if k == 1:
            ptr_side += [(far_call, 0)]
            val_side += [(far_call, constructor + 2 * system, "VALUE_AFTER_FAR_CALL")]
        marker = int((k in abi_regs | reserved and far_call) or (k >= 1 and far_return))
        zero = int((k in abi_regs and cleanup) or (k in reserved and far_call) or (k >= 1 and far_return))
        ptr_side.append((marker, 0))
        val_side += [(zero, 0, "VALUE_AFTER_ZERO_OUT"), (write1, dst1[1], "VALUE_AFTER_DST1")]
        any0, as0 = int(any(f for f, _ in ptr_side)), sum(f * v for f, v in ptr_side)
        after0 = as0 if any0 else reg[0]
        as1 = write1 * dst1[0]
        for name, v in (("WRITE_AS_DST0", write0), ("REMOVE_PTR_MARKER", marker), ("ZERO_OUT", zero), ("ANY_PTR_UPDATE_AS_DST0", any0),
                        ("IS_PTR_AS_DST0", as0), ("IS_PTR_AFTER_DST0", after0), ("IS_PTR_AS_DST1", as1), ("IS_PTR_AFTER_DST1", as1 if write1 else after0)):
            per[name].append(v)
        value = reg[1]
        for flag, cand, name in val_side:
            value = cand if flag else value
            per[name] += limbs(value)
    out.update(per)
    return out


def mutate_writeback_inputs(trace, seed):
    """the block is a pure function of its inputs: any property bits, ABI bytes, targets, register indices, flags (in place)"""
    rng = np.random.default_rng(seed)
    n = trace.shape[1]
    trace[K["PROPS"]] = rng.integers(0, 1 << 38, n).astype(np.uint64)
    for name in ("DST0_REG", "DST1_REG"):
        trace[K[name]] = rng.integers(0, 16, n)
    upd, mem = rng.integers(0, 2, n), rng.integers(0, 2, n)
    trace[K["DST0_UPDATE_REGISTER"]], trace[K["DST0_PERFORMS_MEMORY_ACCESS"]], trace[K["PERFORM_DST0_MEMORY_WRITE"]] = upd & (1 - mem), mem, mem & rng.integers(0, 2, n)
    trace[K["SRC0"] + 8] = rng.integers(0, 1 << 32, n) * (rng.random(n) < 0.7)
    for i in range(1, 6):
        trace[K["SRC1"] + i] = rng.integers(0, 1 << 32, n) * (rng.random(n) < 0.3) if i > 1 else rng.integers(0, 1 << 16, n) << (16 * (rng.random(n) < 0.3))
    trace[K["DST0"]], trace[K["DST1"]] = rng.integers(0, 2, n), rng.integers(0, 2, n)


def test_writeback_cells_against_python_integers(orc):
    assert sum(WBW.values()) == WB["NUM_COLS"] == 513
    isa = I.Isa().isa
    seen = {"far_call": 0, "far_return": 0, "dst1_zero": 0, "write0": 0, "memory": 0}
    for far, seed, mutate in ((False, 33, False), (True, 5, False), (True, 7, True)):
        trace, snaps = vm_trace_and_snapshots(orc, 1500, seed, far)
        trace, snaps = trace.copy(), snaps[:1501]
        if mutate:
            mutate_writeback_inputs(trace, seed)
        cycles = trace.shape[1]
        g = O.vm_writeback_cells(orc, isa, trace, snaps, cycles)
        for r in range(cycles):
            try:
                want = writeback_reference(isa, trace, snaps, r)
            except AssertionError:
                assert mutate          # the cycle.rs:310 identity ties the three flags together on REAL traces only
                continue
            assert set(want) == set(WBW)
            for name, v in want.items():
                got = [int(x) for x in g[WB[name]:WB[name] + WBW[name], r]]
                assert got == (v if isinstance(v, list) else [v]), (r, name, got, v)
            if not mutate:
                seen["far_call"] += want["FAR_CALL_UPDATE"]; seen["far_return"] += want["FAR_RETURN_UPDATE"]; seen["write0"] += sum(want["WRITE_AS_DST0"])
                seen["dst1_zero"] += int(int(trace[K["DST1_REG"], r]) != 0 and not int(trace[K["DST1_UPDATE_REGISTER"], r]))
                seen["memory"] += int(trace[K["PERFORM_DST0_MEMORY_WRITE"], r])
    assert all(seen.values()), seen
