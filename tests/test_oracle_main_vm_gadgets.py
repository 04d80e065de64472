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
