#!/usr/bin/env python3
"""Writes include/zkc_b200_vm_variables.json: the machine-readable map from the named witness cells of one main_vm cycle to the
place in the reference that allocates them (file:line into matter-labs/era-zkevm_circuits @ 8bf2454, src/main_vm/...), for the
six outputs of the engine: the DENSE trace (enum zkc_vm_col) and the five oblivious blocks (ZKC_VM_GADGET_COLUMNS,
ZKC_VM_STATE_GADGET_COLUMNS, ZKC_VM_MEMORY_SPONGE_COLUMNS, ZKC_VM_PRESTATE_COLUMNS, ZKC_VM_WRITEBACK_COLUMNS).  Column numbers and widths are taken from include/zkc_b200.h (the
single source); this file only attaches (reference, what) to every group.  tests/test_vm_variables.py checks the result against
the header and, when /root/reference is present, that every cited line exists.

usage: python tools/gen_vm_variables.py            (rewrites the JSON)"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

P, C, O = "src/main_vm/pre_state.rs", "src/main_vm/cycle.rs", "src/main_vm/opcodes/"

# ---- DENSE trace: group name -> (reference, what).  Widths follow from the enum (next group's first column). ----------------------
DENSE = {
    "SHOULD_SKIP_CYCLE": (P + ":88-89", "execution_has_ended = callstack.is_empty(cs); should_skip_cycle"),
    "PENDING_EXCEPTION_IN": (P + ":90", "current_state.pending_exception on entry"),
    "SHOULD_READ_OPCODE": (P + ":128-129", "should_try_to_read_opcode AND should_read_for_new_pc"),
    "SUPER_PC": (P + ":113", "split_pc: pc >> 2 (main_vm/utils.rs:46-104)"),
    "SUB_PC": (P + ":113", "split_pc: pc & 3"),
    "CODE_WORD": (P + ":177-182", "UInt256::conditionally_select(should_read_opcode, code_word, previous_code_word)"),
    "OPCODE": (P + ":184-221", "the 64-bit opcode selected by the sub-pc bitmask, after mask_into_nop / mask_into_panic: low, high u32"),
    "VARIANT": ("src/main_vm/decoded_opcode.rs:395-527", "opcode variant bits (11-bit table key) of the masked opcode"),
    "CONDITION_IDX": ("src/main_vm/decoded_opcode.rs:395-527", "condition selector of the opcode"),
    "CONDITION": ("src/main_vm/decoded_opcode.rs:52-59", "initial_decoding.condition: partially_decode_from_integer_and_resolve_condition (:395-527), a table lookup over (condition, flags) (tables/conditional.rs:28-51)"),
    "ERGS_COST": ("src/main_vm/decoded_opcode.rs:68", "masked_ergs_cost: the opcode price of the decoding table, 0 for a skipped cycle"),
    "OUT_OF_ERGS": ("src/main_vm/decoded_opcode.rs:78", "ergs_left.overflowing_sub(cost) underflow"),
    "KERNEL_MODE_EXCEPTION": ("src/main_vm/decoded_opcode.rs:87", "kernel-only opcode outside kernel mode"),
    "STATIC_EXCEPTION": ("src/main_vm/decoded_opcode.rs:89-92", "state-changing opcode in a static context"),
    "CALLSTACK_IS_FULL": (P + ":258", "callstack.is_full(cs) (vm_state/callstack.rs:38-42)"),
    "EXPLICIT_PANIC": ("src/main_vm/decoded_opcode.rs:84", "ret.panic variant"),
    "MASK_INTO_PANIC": ("src/main_vm/decoded_opcode.rs:94-106", "any exception: the opcode becomes the panic encoding"),
    "MASK_INTO_NOP": ("src/main_vm/decoded_opcode.rs:135-139", "condition not met / skipped cycle: the opcode becomes nop"),
    "PROPS": ("src/main_vm/decoded_opcode.rs:152-165", "the 48 property Booleans of the masked opcode (opcode_bitmask.rs:24-130) as one integer: bit k = Boolean k"),
    "DIRTY_ERGS_LEFT": (P + ":271-288", "dirty_ergs_left = preliminary_ergs_left (also AfterDecodingCarryParts, :513)"),
    "SRC0_REG": ("src/main_vm/decoded_opcode.rs:176", "split_register_encoding_byte: 4-bit src0 register index (its one-hot mask, :192-193, drives the 15-way select of pre_state.rs:303-309)"),
    "SRC1_REG": ("src/main_vm/decoded_opcode.rs:176", "src1 register index"),
    "DST0_REG": ("src/main_vm/decoded_opcode.rs:177", "dst0 register index"),
    "DST1_REG": ("src/main_vm/decoded_opcode.rs:177", "dst1 register index"),
    "IMM0": ("src/main_vm/decoded_opcode.rs:204", "imm0 (UInt16)"),
    "IMM1": ("src/main_vm/decoded_opcode.rs:205", "imm1 (UInt16)"),
    "SRC0_PAGE": (P + ":351-358", "resolve_memory_region_and_index_for_source: page (main_vm/utils.rs:233-304)"),
    "SRC0_INDEX": (P + ":351-358", "... index"),
    "SHOULD_READ_SRC0": (P + ":351-358", "... should_read_memory_for_src0"),
    "SP_AFTER_SRC0": (P + ":351-358", "... new_sp_after_src0"),
    "DST0_PAGE": (P + ":361-367", "resolve_memory_region_and_index_for_dest: page (main_vm/utils.rs:306-384)"),
    "DST0_INDEX": (P + ":361-367", "... index"),
    "DST0_PERFORMS_MEMORY_ACCESS": (P + ":361-367", "... should_write_memory_for_dst0"),
    "NEW_SP": (P + ":361-370", "new_sp, stored into the draft state's saved_context.sp"),
    "SRC0_FROM_MEMORY": (P + ":374-391", "may_be_read_memory_for_source_operand: the VMRegister read (is_pointer, 8 limbs); zero when not read"),
    "SWAP_OPERANDS": (P + ":418-446", "swap_operands"),
    "SRC0": (P + ":451-481", "src0 after the swap and conditionally_erase_fat_pointer_data: is_pointer, 8 limbs (= src0_view.is_ptr / u32x8_view)"),
    "SRC1": (P + ":453-482", "src1 likewise"),
    "DST0": (C + ":191-222", "dst0 = dot_product over diffs_accumulator.dst_0_values: is_pointer, 8 limbs"),
    "DST1": (C + ":224-246", "dst1 = dot_product over dst_1_values"),
    "PERFORM_DST0_MEMORY_WRITE": (C + ":248-254", "dst0_performs_memory_access AND dst0_update_potentially_to_memory"),
    "DST0_UPDATE_REGISTER": (C + ":298-304", "the register-file update flag of dst0"),
    "DST1_UPDATE_REGISTER": (C + ":185-187", "OR of the should_update_dst1 flags the gadgets push (collected by the reference, never read): whether DST1 holds a candidate; the register WRITE is gated by the dst1 selector alone (:330), see ZKC_VMW_IS_PTR_AS_DST1 / VALUE_AFTER_DST1"),
    "FLAGS_OUT": (C + ":603-606", "new_state.flags after the candidates: overflow_or_less_than, equal, greater_than"),
    "PENDING_EXCEPTION_OUT": (C + ":615-616", "Boolean::multi_or over pending_exceptions"),
    "PC_OUT": (C + ":438-445", "new pc after new_pc_candidates"),
    "ERGS_OUT": (C + ":448-463", "new ergs_remaining after new_ergs_left_candidates"),
    "HEAP_BOUND_OUT": (C + ":487-502", "heap_upper_bound after new_heap_bounds"),
    "AUX_HEAP_BOUND_OUT": (C + ":505-520", "aux_heap_upper_bound after new_aux_heap_bounds"),
    "MEMQ_LENGTH_OUT": (C + ":525-531", "memory_queue_length after memory_queue_candidates"),
    "DEPTH_OUT": (C + ":609-612", "callstack.context_stack_depth after the callstack candidates"),
    "FORWARD_TAIL_OUT": (C + ":547-566", "log_queue_forward_tail (4) and log_queue_forward_part_length after log_queue_forward_candidates"),
    "ROLLBACK_HEAD_OUT": (C + ":570-600", "reverted_queue_head (4) and reverted_queue_segment_len after log_queue_rollback_candidates"),
    "SPONGE_ENFORCE": (C + ":673-757", "should_enforce of the nine Poseidon2 relations (engine slot numbering, include/zkc_b200.h)"),
    "SPONGE_FINAL": (C + ":937-957", "R(initial_state) of every ENFORCED relation, 9 x 12 (zero otherwise)"),
    "OP_AUX": (O + "uma.rs:18-1001", "cells of the executed uma / log (log.rs:16-467) / call-ret (call_ret.rs:24-512) gadget, layout in include/zkc_b200.h"),
}

A, B, M, S = O + "add_sub.rs", O + "binop.rs", O + "mul_div.rs", O + "shifts.rs"
GADGET = {
    "SRC0_BYTES": ("src/main_vm/register_input_view.rs:27-53", "src0_view.u8x32_view: decompose_into_bytes_unchecked of the 8 limbs"),
    "SRC1_BYTES": ("src/main_vm/register_input_view.rs:27-53", "src1_view.u8x32_view"),
    "ADD_RESULT": (A + ":18-22", "allocate_addition_result_unchecked: limbs (:168-224)"), "ADD_OF": (A + ":18-22", "... overflow"),
    "SUB_RESULT": (A + ":24-28", "allocate_subtraction_result_unchecked: limbs (:226-282)"), "SUB_UF": (A + ":24-28", "... underflow"),
    "ADDSUB_RESULT": (A + ":52-57", "UInt32::parallel_select(apply_add, add, sub)"),
    "ADDSUB_NEW_B": (A + ":91-96", "new_b of the shuffled relation"), "ADDSUB_NEW_C": (A + ":98-103", "new_c"),
    "ADDSUB_NEW_OF": (A + ":105", "Boolean::conditionally_select(apply_add, of, uf)"),
    "ADDSUB_LIMB_IS_ZERO": (A + ":115", "result.map(is_zero)"), "ADDSUB_RESULT_IS_ZERO": (A + ":116", "multi_and of the limb flags"),
    "ADDSUB_GT": (A + ":119", "NOT (new_of OR result_is_zero)"), "ADDSUB_APPLY_ANY": (A + ":134", "apply_add OR apply_sub"),
    "ADDSUB_UPDATE_FLAGS": (A + ":147", "apply_any AND set_flags"),
    "BINOP_COMPOSITE": (B + ":132-142", "32 lookups into the and | or << 16 | xor << 32 table"),
    "BINOP_ALL_RESULTS": (B + ":145-174", "the 96 decomposed chunks (and, or, xor per byte)"),
    "BINOP_AND": (B + ":76-89", "and_chunks: UInt32::from_le_bytes per limb"), "BINOP_OR": (B + ":77-89", "or_chunks"), "BINOP_XOR": (B + ":78-89", "xor_chunks"),
    "BINOP_RESULT": (B + ":91-92", "parallel_select(is_and, and, xor) then (is_or, or, .)"),
    "BINOP_LIMB_IS_ZERO": (B + ":94", "result.map(is_zero)"), "BINOP_RESULT_IS_ZERO": (B + ":95", "multi_and"),
    "BINOP_UPDATE_FLAGS": (B + ":116", "should_apply AND should_set_flags"),
    "MUL_LOW": (M + ":237-238", "allocate_mul_result_unchecked: low limbs (:20-91)"), "MUL_HIGH": (M + ":237-238", "... high limbs"),
    "DIV_QUOTIENT": (M + ":239-240", "allocate_div_result_unchecked: quotient (:93-174; divisor 0 -> (0, a), :119-123)"),
    "DIV_REMAINDER": (M + ":239-240", "... remainder"),
    "MULDIV_RESULT_0": (M + ":253-258", "parallel_select(apply_mul, mul_low, quotient)"),
    "MULDIV_RESULT_1_UNMASKED": (M + ":259-264", "parallel_select(apply_mul, mul_high, remainder)"),
    "MULDIV_REM_TO_ENFORCE": (M + ":275-280", "rem_to_enforce"), "MULDIV_A_TO_ENFORCE": (M + ":281-282", "a_to_enforce"),
    "MULDIV_MUL_LOW_TO_ENFORCE": (M + ":284-285", "mul_low_to_enforce"), "MULDIV_MUL_HIGH_TO_ENFORCE": (M + ":286-291", "mul_high_to_enforce"),
    "MUL_HIGH_IS_ZERO": (M + ":302", "all_limbs_are_zero(mul_high)"), "MUL_LOW_IS_ZERO": (M + ":303", "all_limbs_are_zero(mul_low)"),
    "MUL_OF": (M + ":304", "NOT high_is_zero"), "MUL_GT": (M + ":306-310", "NOT of AND NOT eq"),
    "DIV_DIVISOR_IS_ZERO": (M + ":313", "all_limbs_are_zero(src1)"), "DIV_QUOTIENT_IS_ZERO": (M + ":316", ""), "DIV_REMAINDER_IS_ZERO": (M + ":317", ""),
    "DIV_SUB_RESULT": (M + ":322-323", "allocate_subtraction_result_unchecked(remainder, divisor)"),
    "DIV_REMAINDER_IS_LESS": (M + ":322-323", "its borrow"), "DIV_MASK_REMAINDER": (M + ":346", "apply_div AND divisor_is_zero"),
    "MULDIV_RESULT_1": (M + ":347", "result_1 masked"), "DIV_EQ": (M + ":350-353", ""), "DIV_GT": (M + ":354-357", ""),
    "MULDIV_OF": (M + ":359", ""), "MULDIV_EQ": (M + ":360", ""), "MULDIV_GT": (M + ":361", ""),
    "MULDIV_APPLY_ANY": (M + ":369", ""), "MULDIV_SET_FLAGS": (M + ":396", ""),
    "SHIFT_AMOUNT": (S + ":57-58", "src1_view.u8x32_view[0]"), "SHIFT_IS_ZERO": (S + ":62", ""), "SHIFT_INVERTED": (S + ":63-65", "256 - shift"),
    "SHIFT_CHANGE_FLAG": (S + ":67-70", "is_ror AND NOT shift_is_zero"), "SHIFT_FULL": (S + ":71-74", "full_shift"),
    "SHIFT_CONSTANT": (S + ":76", "get_shift_constant (:200-221, tables/bitshift.rs): 2^full_shift as 8 limbs"),
    "SHIFT_IS_RIGHT": (S + ":78-81", ""), "SHIFT_RSHIFT_Q": (S + ":82", "allocate_div_result_unchecked(reg, 2^s): quotient"),
    "SHIFT_RSHIFT_R": (S + ":82", "... remainder"), "SHIFT_APPLY_LEFT": (S + ":84-87", ""),
    "SHIFT_LSHIFT_LOW": (S + ":88", "allocate_mul_result_unchecked(reg, 2^s): low"), "SHIFT_LSHIFT_HIGH": (S + ":88", "... high"),
    "SHIFT_REM_TO_ENFORCE": (S + ":99-100", ""), "SHIFT_A_TO_ENFORCE": (S + ":101", ""), "SHIFT_MUL_LOW_TO_ENFORCE": (S + ":103", ""),
    "SHIFT_MUL_HIGH_TO_ENFORCE": (S + ":104-105", ""), "SHIFT_SUB_RESULT": (S + ":117-120", "allocate_subtraction_result_unchecked(rshift_r, 2^s)"),
    "SHIFT_REMAINDER_IS_LESS": (S + ":117-120", "its borrow"), "SHIFT_TEMP_RESULT": (S + ":136", ""), "SHIFT_RESULT": (S + ":138-152", "final_result"),
    "SHIFT_RESULT_IS_ZERO": (S + ":155", ""), "SHIFT_SET_FLAGS": (S + ":164", ""),
    "RANGE_CHECK": (C + ":619-629", "the selected conditional range check (8 x UInt32::from_variable_checked)"),
    "ADDREL_A": (C + ":632-647", "the ONE enforced AddSubRelation after the selects (opcodes/mod.rs:46-65): a"), "ADDREL_B": (C + ":632-647", "b"),
    "ADDREL_C": (C + ":632-647", "c"), "ADDREL_OF": (C + ":632-647", "of"),
    "ADDREL_CARRY": (O + "mod.rs:101-125", "enforce_addition_relation: the 8 intermediate carries"),
    "MULREL_A": (C + ":651-667", "the ONE enforced MulDivRelation after the selects (opcodes/mod.rs:75-98): a"), "MULREL_B": (C + ":651-667", "b"),
    "MULREL_REM": (C + ":651-667", "rem"), "MULREL_LOW": (C + ":651-667", "mul_low"), "MULREL_HIGH": (C + ":651-667", "mul_high"),
    "MULREL_PARTIAL_LOW": (O + "mod.rs:146-165", "enforce_mul_relation: low word of the 64 UInt32::fma_with_carry"),
    "MULREL_PARTIAL_HIGH": (O + "mod.rs:146-165", "... high word"), "MULREL_ROW_END": (O + "mod.rs:166-169", "the 8 row-end additions"),
}

T, X = O + "ptr.rs", O + "context.rs"
STATE = {
    "PTR_SRC1_IS_INTEGER": (T + ":53", "src_1.is_ptr.negated"), "PTR_ARGS_VALID": (T + ":56", ""), "PTR_ARGS_INVALID": (T + ":57", ""),
    "PTR_SRC1_LIMB_IS_ZERO": (T + ":60-63", "src1 u32x8_view.map(is_zero)"), "PTR_SRC1_32_256_IS_ZERO": (T + ":64", ""),
    "PTR_SRC1_0_128_IS_ZERO": (T + ":65", ""), "PTR_SRC1_32_256_IS_NONZERO": (T + ":67", ""), "PTR_ARITH_VARIANT": (T + ":70", ""),
    "PTR_TOO_LARGE_OFFSET": (T + ":71", ""), "PTR_SRC1_0_128_IS_NONZERO": (T + ":74", ""), "PTR_DIRTY_PACK": (T + ":75-76", ""),
    "PTR_ADD_RESULT": (T + ":79", "src0[0].overflowing_add(src1[0])"), "PTR_ADD_OF": (T + ":79", ""), "PTR_ADD_PANIC": (T + ":80", ""),
    "PTR_SUB_RESULT": (T + ":82", "src0[0].overflowing_sub(src1[0])"), "PTR_SUB_UF": (T + ":82", ""), "PTR_SUB_PANIC": (T + ":83", ""),
    "PTR_SHRINK_RESULT": (T + ":85", "src0[3].overflowing_sub(src1[0])"), "PTR_SHRINK_UF": (T + ":85", ""), "PTR_SHRINK_PANIC": (T + ":86", ""),
    "PTR_ANY_PANIC": (T + ":88-98", ""), "PTR_SHOULD_PANIC": (T + ":100", ""), "PTR_OK": (T + ":101", ""), "PTR_UPDATE_REGISTER": (T + ":102", ""),
    "PTR_LOW_IF_ADD": (T + ":107-112", ""), "PTR_LOW_IF_ADD_OR_SUB": (T + ":115-120", ""), "PTR_96_128_IF_SHRINK": (T + ":123-128", ""),
    "PTR_HIGHEST_128": (T + ":130-145", ""), "PTR_LOWEST32": (T + ":147-152", ""), "PTR_96_128": (T + ":154-159", ""),
    "PTR_DST0": (T + ":161-175", "the dst0 candidate: is_pointer of src0 + 8 limbs (existing variables, listed for convenience)"),
    "JUMP_DST": (O + "jump.rs:27-33", "UInt16::from_le_bytes of src0 bytes 0, 1"),
    "CTX_WRITE_TO_CONTEXT": (X + ":115", ""), "CTX_SET_PUBDATA_ERGS": (X + ":116", ""), "CTX_INCREMENT_TX": (X + ":117", ""),
    "CTX_READ_ONLY": (X + ":120-123", "the reference's name for 'one of the three state-setting variants'"), "CTX_WRITE_LIKE": (X + ":124", ""),
    "CTX_WRITE_TO_DST0": (X + ":126", ""), "CTX_INCREMENTED_TX_NUMBER": (X + ":131-133", "tx_number_in_block.overflowing_add(1)"),
    "CTX_TX_OF": (X + ":131-133", ""), "CTX_META_HIGHEST": (X + ":145-165", "UInt32::from_le_bytes(this, caller, code shard ids, 0)"),
    "CTX_LOW_U32": (X + ":203-208", "select(is_retrieve_ergs_left, preliminary_ergs_left, sp)"), "CTX_RESULT_128": (X + ":214-223", ""),
    "CTX_RESULT_160_THIS": (X + ":235-245", ""), "CTX_RESULT_160_CALLER": (X + ":247-257", ""), "CTX_RESULT_160_CODE": (X + ":259-269", ""),
    "CTX_RESULT_256": (X + ":284-285", "parallel_select(is_retrieve_meta, meta_as_register, result_256): the dst0 candidate's value"),
}

U = "src/main_vm/utils.rs"
MEMQ = {
    "SELECTED": (C + ":673-721", "no opcode with its own sponges applies: enforce_sponges slots 1 / 2 run on the src0 read / dst0 write candidates"),
    "FETCH_INIT": (U + ":194-210", "may_be_read_memory_for_code: query.encode (8) || memory_queue_state[8..12]"),
    "FETCH_FINAL": (U + ":212-213", "R::compute_round_function(initial_state), on every cycle"),
    "FETCH_STATE_AFTER": (U + ":225-230", "Num::parallel_select(should_access, final_state_candidate, current state)"),
    "FETCH_LENGTH_AFTER": (U + ":216-223", "UInt32::conditionally_select(should_access, length + 1, length)"),
    "SRC0_INIT": (U + ":458-492", "may_be_read_memory_for_source_operand: query.encode, initial_state"),
    "SRC0_FINAL": (C + ":937-957", "enforce_sponges: R(initial_state) of the slot's selected candidate (this one when SELECTED)"),
    "SRC0_STATE_AFTER": (U + ":506-511", "parallel_select(should_access, simulated_final_state, current state)"),
    "SRC0_LENGTH_AFTER": (U + ":497-504", ""),
    "DST0_INIT": (C + ":846-884", "may_be_write_memory: query.encode, initial_state"),
    "DST0_FINAL": (C + ":937-957", "enforce_sponges: R(initial_state) (when SELECTED)"),
    "DST0_STATE_AFTER": (C + ":898-903", "parallel_select(should_write_dst0, simulated_final_state, current tail)"),
    "DST0_LENGTH_AFTER": (C + ":889-896", ""),
}

D = "src/main_vm/decoded_opcode.rs"
PRESTATE = {
    "EXECUTE_CYCLE": (P + ":91", "should_skip_cycle.negated"),
    "SHOULD_TRY_TO_READ_OPCODE": (P + ":98", "execute_cycle.mask_negated(pending_exception)"),
    "PENDING_EXCEPTION_TAKEN_DOWN": (P + ":103-105", "pending_exception.mask_negated(execute_pending_exception_at_this_cycle): the flag masked by itself"),
    "PC_PLUS_ONE": (P + ":111", "current_pc.overflowing_add(1): UInt16 result"), "PC_PLUS_ONE_OF": (P + ":111", "... its (dropped) overflow bit"),
    "CODE_PAGES_ARE_EQUAL": (U + ":114", "should_read_memory: UInt32::equals(previous_code_page, current_code_page)"),
    "SUPER_PC_ARE_EQUAL": (U + ":115", "UInt16::equals(super_pc, previous_super_pc)"),
    "CAN_SKIP_READ": (U + ":117", "Boolean::multi_and"), "SHOULD_READ_FOR_NEW_PC": (U + ":119", "can_skip.negated"),
    "TIMESTAMPS": (P + ":144-150", "four increment_unchecked of current_state.timestamp: first decommit / precompile read, second ... write, dst write, next cycle"),
    "NEXT_CYCLE_TIMESTAMP": (P + ":151-156", "UInt32::conditionally_select(should_skip_cycle, timestamp, next_cycle_timestamp)"),
    "SUBPC_BITMASK": (P + ":185", "subpc_spread.spread_into_bits::<3> (the spread is the VMSubPCToBitmaskTable lookup of split_pc, main_vm/utils.rs:95-104)"),
    "OPCODE_SELECT_CHAIN": (P + ":188-206", "three <[UInt32; 2]>::conditionally_select steps over code_word.inner[6..8] / [4..6] / [2..4] / [0..2]: (low, high) after each; "
                                            "the last pair is the opcode BEFORE mask_into_nop / mask_into_panic"),
    "SRC0_SELECTORS": (D + ":192-193", "reg_idx_into_bitspread(src0_encoding).spread_into_bits::<REGISTERS_COUNT>: 15 Booleans"),
    "SRC1_SELECTORS": (D + ":195-196", ""), "DST0_SELECTORS": (D + ":198-199", ""), "DST1_SELECTORS": (D + ":201-202", ""),
    "DRAFT_SRC0_CHAIN": (P + ":303-309", "draft_src0 after each of the 15 VMRegister::conditionally_select steps: 15 x (is_pointer, 8 limbs)"),
    "SRC1_REGISTER_CHAIN": (P + ":312-318", "src1_register after each step, likewise"),
    "DST0_REG_LOW_CHAIN": (P + ":320-328", "current_dst0_reg_low after each of the 15 UInt32::conditionally_select steps"),
    "SRC0_REG_LOWEST": (P + ":310", "draft_src0.value.inner[0].low_u16"), "DST0_REG_LOWEST": (P + ":329", "current_dst0_reg_low.low_u16"),
    "STACK_PAGE": (P + ":342", "base_page.increment_unchecked"), "HEAP_PAGE": (P + ":343", "stack_page.increment_unchecked"),
    "AUX_HEAP_PAGE": (P + ":344", "heap_page.increment_unchecked"),
    "SRC_ABSOLUTE_MODE": (U + ":260", "resolve_memory_region_and_index_for_source: multi_or(use_code, use_stack_absolute)"),
    "SRC_INDEX_FOR_ABSOLUTE": (U + ":261", "register_low_value.overflowing_add(imm0)"), "SRC_INDEX_FOR_RELATIVE": (U + ":262", "current_sp.overflowing_sub(index_for_absolute)"),
    "SRC_USE_STACK": (U + ":272-279", ""), "SRC_DID_READ_UNMASKED": (U + ":280", "did_read before the NOP rule of :287"),
    "NOT_NOP": (U + ":286", "is_nop.negated (allocated again at :353 with the same value)"),
    "DST_INDEX_FOR_ABSOLUTE": (U + ":327", "resolve_memory_region_and_index_for_dest: register_low_value.overflowing_add(imm1)"),
    "DST_INDEX_FOR_RELATIVE_WITH_PUSH": (U + ":328", "current_sp.overflowing_add(index_for_absolute), current_sp = new_sp_after_src0"),
    "DST_INDEX_FOR_RELATIVE": (U + ":329", "current_sp.overflowing_sub(index_for_absolute)"),
    "DST_DID_WRITE_UNMASKED": (U + ":340-347", "did_write before the NOP rule of :354"),
    "DST_INDEX_SOMEWHAT_RELATIVE": (U + ":356-361", "index_with_somewhat_relative_addressing: a push writes at the current sp"),
    "SRC0_AFTER_USE_REG": (P + ":405", "VMRegister::conditionally_select(use_reg, draft_src0, src0_register_from_mem)"),
    "SRC0_AFTER_USE_IMM": (P + ":408-412", "VMRegister::conditionally_select(use_imm, imm_as_reg, src0)"),
    "SWAP_IS_ASSYMMETRIC": (P + ":431", "multi_or(is_sub, is_div, is_shift)"), "SWAP_T0": (P + ":435", ""), "SWAP_T1": (P + ":443", ""),
    "SRC0_SWAPPED": (P + ":451-452", "VMRegister::conditionally_select(swap_operands, selected_src1, selected_src0), before the erasure"),
    "SRC1_SWAPPED": (P + ":453-454", ""),
    "NOT_KERNEL_MODE": (P + ":458", ""), "KEEPS_POINTERS": (P + ":474-475", "multi_or(is_ret, is_ptr, is_uma, is_far_call)"), "SHOULD_ERASE": (P + ":474-475", "its negation"),
    "SHOULD_ERASE_SRC0": (P + ":459-477", "multi_and(src0.is_pointer, should_erase, not_kernel_mode)"),
    "SHOULD_ERASE_SRC1": (P + ":479", "multi_and(src1.is_pointer, not_kernel_mode)"),
}

FC, RT = O + "call_ret_impl/far_call.rs", O + "call_ret_impl/ret.rs"
WRITEBACK = {
    "DST0_UPDATE_POTENTIALLY_TO_MEMORY": (C + ":172", "Boolean::multi_or over the flags of the memory-capable dst0 candidates"),
    "CAN_UPDATE_DST0_AS_REGISTER_ONLY": (C + ":189", "Boolean::multi_or over the flags of the register-only dst0 candidates"),
    "DST0_PERFORMS_REG_UPDATE": (C + ":298", "dst0_performs_memory_access.negated"),
    "DST0_REG_UPDATE_T": (C + ":299-302", "multi_and(dst0_performs_reg_update, dst0_update_potentially_to_memory); ZKC_VM_DST0_UPDATE_REGISTER = can_update_dst0_as_register_only | t (:304)"),
    "FAR_CALL_UPDATE": (FC + ":1042-1043", "execute: the flag of the far call's specific updates of r1 / r2"),
    "FAR_CALL_NON_SYSTEM": (FC + ":1045", "far_call_abi.system_call.negated (system_call after the kernel-target mask, :430-431)"),
    "FAR_CALL_CLEANUP_REGISTER": (FC + ":1046", "multi_and(execute, non_system_call)"),
    "FAR_RETURN_UPDATE": (RT + ":442", "update_specific_registers_on_ret = multi_and(execute, is_far_return)"),
    "FAR_CALL_NEW_R2_LOW": (FC + ":1021-1028", "r2_low = constructor_call + 2 * system_call (Num::fma)"),
    "WRITE_AS_DST0": (C + ":328", "per register: multi_and(dst0_update_register, dst0 selector); write_as_dst1 (:330) IS the dst1 selector, ZKC_VMP_DST1_SELECTORS"),
    "REMOVE_PTR_MARKER": (C + ":365", "multi_or over remove_ptr_on_specific_registers[idx] (far_call.rs:1064-1067, ret.rs:461); no cell when no list names the register"),
    "ZERO_OUT": (C + ":377", "multi_or over specific_registers_zeroing[idx] (far_call.rs:1051-1054, ret.rs:454); no cell when no list names the register"),
    "ANY_PTR_UPDATE_AS_DST0": (C + ":381", "multi_or(write_as_dst0, the specific updates' flags, remove_ptr_marker)"),
    "IS_PTR_AS_DST0": (C + ":390-397", "dot_product over (flag, is_pointer) of dst0, the specific updates, (marker, false)"),
    "IS_PTR_AFTER_DST0": (C + ":399-404", "Boolean::conditionally_select(any_ptr_update_as_dst0, is_ptr_as_dst0, registers[idx].is_pointer)"),
    "IS_PTR_AS_DST1": (C + ":408-415", "dot_product over (write_as_dst1, dst1_is_ptr)"),
    "IS_PTR_AFTER_DST1": (C + ":416-421", "Boolean::conditionally_select(any_ptr_update_as_dst1, ...): the marker in the next state"),
    "VALUE_AFTER_DST0": (C + ":430-431", "UInt256::conditionally_select(write_as_dst0, dst0_value, registers[idx].value): 15 x 8 limbs"),
    "VALUE_AFTER_FAR_CALL": (C + ":430-431", "... the far call's specific update (r1 = final_fat_ptr.into_register, far_call.rs:1008; r2, :1030-1039): 2 x 8 limbs"),
    "VALUE_AFTER_FAR_RETURN": (C + ":430-431", "... the far return's specific update of r1 (ret.rs:441-445): 8 limbs"),
    "VALUE_AFTER_ZERO_OUT": (C + ":430-431", "... (zero_out_reg, zero_u256): 15 x 8 limbs"),
    "VALUE_AFTER_DST1": (C + ":430-431", "... (write_as_dst1, dst1_value), the last of the chain: the value in the next state: 15 x 8 limbs"),
}


def dense_layout():
    text = open(os.path.join(ROOT, "include", "zkc_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", text[text.index("enum zkc_vm_col {"):], flags=re.S)
    body = body[:body.index("};")]
    cols = [(m.group(1), int(m.group(2))) for m in re.finditer(r"ZKC_VM_([A-Z0-9_]+)\s*=\s*(\d+)", body)]
    n = dict(cols)["NUM_COLS"]
    cols = sorted((c for c in cols if c[0] != "NUM_COLS"), key=lambda c: c[1])
    return [(name, first, (cols[i + 1][1] if i + 1 < len(cols) else n) - first) for i, (name, first) in enumerate(cols)], n


def block(prefix, layout, table):
    out = []
    for name, first, width in layout:
        ref, what = table[name]
        out.append({"column": first, "name": prefix + name, "width": width, "reference": ref, "what": what})
    return out


def build():
    from era_zkevm_circuits_b200 import abi
    dense, n_dense = dense_layout()
    x = lambda cols, widths: [(k, cols[k], widths[k]) for k in widths]
    return {
        "circuit": "main_vm: one vm_cycle (src/main_vm/cycle.rs:28-795, pre_state.rs:71-519)",
        "trace": "column-major uint64 block[column * limit + cycle]; the six blocks are the outputs of zkc_main_vm_entry_point (DENSE), "
                 "zkc_main_vm_gadget_cells, zkc_main_vm_state_gadget_cells, zkc_main_vm_memory_sponge_cells, zkc_main_vm_prestate_cells, zkc_main_vm_writeback_cells (include/zkc_b200.h)",
        "provenance": "values the reference's own source names, one group per allocation site; boojum is un-vendored, so which INTERNAL cells its "
                      "gadgets add (selects, range-check decompositions, Poseidon2 round cells) is not listed here: they stay host-resolved",
        "not_produced": ["non-selected cells of apply_uma / apply_log / apply_calls_and_ret", "the in-circuit permutations of slots 3..8 when not enforced",
                         "intermediate cells of the scalar state-diff select chains (pc, ergs, flags, queues; the register write-back chains are in the writeback block, the 15-way register selects of create_prestate in the prestate block)", "lookup / range-check decompositions"],
        "blocks": [
            {"block": "dense", "entry_point": "zkc_main_vm_entry_point", "enum": "zkc_vm_col", "num_columns": n_dense, "columns": block("ZKC_VM_", dense, DENSE)},
            {"block": "gadget", "entry_point": "zkc_main_vm_gadget_cells", "enum": "zkc_vm_gadget_col", "num_columns": abi.VMG_COLS["NUM_COLS"],
             "columns": block("ZKC_VMG_", x(abi.VMG_COLS, abi.VMG_WIDTHS), GADGET)},
            {"block": "state_gadget", "entry_point": "zkc_main_vm_state_gadget_cells", "enum": "zkc_vm_state_gadget_col", "num_columns": abi.VMS_COLS["NUM_COLS"],
             "columns": block("ZKC_VMS_", x(abi.VMS_COLS, abi.VMS_WIDTHS), STATE)},
            {"block": "memory_sponge", "entry_point": "zkc_main_vm_memory_sponge_cells", "enum": "zkc_vm_memory_sponge_col", "num_columns": abi.VMQ_COLS["NUM_COLS"],
             "columns": block("ZKC_VMQ_", x(abi.VMQ_COLS, abi.VMQ_WIDTHS), MEMQ)},
            {"block": "prestate", "entry_point": "zkc_main_vm_prestate_cells", "enum": "zkc_vm_prestate_col", "num_columns": abi.VMP_COLS["NUM_COLS"],
             "columns": block("ZKC_VMP_", x(abi.VMP_COLS, abi.VMP_WIDTHS), PRESTATE)},
            {"block": "writeback", "entry_point": "zkc_main_vm_writeback_cells", "enum": "zkc_vm_writeback_col", "num_columns": abi.VMW_COLS["NUM_COLS"],
             "columns": block("ZKC_VMW_", x(abi.VMW_COLS, abi.VMW_WIDTHS), WRITEBACK)},
        ],
    }


if __name__ == "__main__":
    doc = build()
    path = os.path.join(ROOT, "include", "zkc_b200_vm_variables.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
        f.write("\n")
    print(path, sum(len(b["columns"]) for b in doc["blocks"]), "groups,", sum(b["num_columns"] for b in doc["blocks"]), "columns")
