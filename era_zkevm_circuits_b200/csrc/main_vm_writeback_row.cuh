// One row of the register write-back block (include/zkc_b200.h, ZKC_VM_WRITEBACK_COLUMNS): what vm_cycle allocates when it applies
// dst0 / dst1 and the far call / far return register conventions to the 15 registers.
//   update flags, write_as_dst0 / write_as_dst1     /root/reference/src/main_vm/cycle.rs:160-189, :303-330
//   specific updates, pointer markers, zero-out     /root/reference/src/main_vm/cycle.rs:349-375
//   the conventions themselves                      /root/reference/src/main_vm/opcodes/call_ret_impl/far_call.rs:1021-1070, ret.rs:442-464
//   is_pointer dot products + selects               /root/reference/src/main_vm/cycle.rs:377-412
//   value select chain                              /root/reference/src/main_vm/cycle.rs:415-433
// __host__ __device__ like main_vm_prestate_row.cuh: the statement the kernel runs is compiled by g++ and compared with the oracle where
// there is no GPU (tests/cpp/prestate_row_host.cpp).  The calling-convention register lists arrive as two 15-bit masks (bit r =
// register r + 1); every step is a mask blend, no lane branches on the opcode.
#pragma once
#include "main_vm_prestate_row.cuh"

namespace zkc {

struct vm_writeback_masks {
    uint32_t system_abi;   // CALL_SYSTEM_ABI_REGISTERS: zeroed by a far call unless it is a system call, pointer marker always removed
    uint32_t reserved;     // CALL_RESERVED_RANGE + CALL_IMPLICIT_PARAMETER_REG_IDX: zeroed and unmarked by every far call
};

ZKC_ROW_FN void vm_writeback_row(const uint64_t *t, const zkc_vm_state *st, vm_writeback_masks lists, uint64_t *out, size_t limit) {
#define IN(col) ZKC_ROW_LD(t + (size_t)(col) * limit)
#define OUT(col, i) out[(size_t)((col) + (i)) * limit]
    const uint64_t props = IN(ZKC_VM_PROPS);
#define KIND(n) ((uint32_t)(props >> ZKC_VM_BIT_TYPE(n)) & 1u)
    const zkc_vm_state *nx = st + 1;
    // ---- the update flags of dst0 (cycle.rs:160-189, :303-310) ----------------------------------------------------------------
    const uint32_t to_memory_capable = KIND(ZKC_OP_ADD) | KIND(ZKC_OP_SUB) | KIND(ZKC_OP_MUL) | KIND(ZKC_OP_DIV) | KIND(ZKC_OP_BINOP) |
                                       KIND(ZKC_OP_SHIFT) | KIND(ZKC_OP_PTR);
    const uint32_t update_register = (uint32_t)IN(ZKC_VM_DST0_UPDATE_REGISTER) & 1u;
    const uint32_t any_update = update_register | ((uint32_t)IN(ZKC_VM_PERFORM_DST0_MEMORY_WRITE) & 1u);
    const uint32_t reg_update = ((uint32_t)IN(ZKC_VM_DST0_PERFORMS_MEMORY_ACCESS) & 1u) ^ 1u;
    const uint32_t potentially = to_memory_capable & any_update;
    OUT(ZKC_VMW_DST0_UPDATE_POTENTIALLY_TO_MEMORY, 0) = potentially;
    OUT(ZKC_VMW_CAN_UPDATE_DST0_AS_REGISTER_ONLY, 0) = (to_memory_capable ^ 1u) & update_register;
    OUT(ZKC_VMW_DST0_PERFORMS_REG_UPDATE, 0) = reg_update;
    OUT(ZKC_VMW_DST0_REG_UPDATE_T, 0) = reg_update & potentially;
    // ---- far call / far return flags (far_call.rs:396-431, :1021-1046; ret.rs:442; call_ret.rs:133) ----------------------------
    const uint32_t far_call = KIND(ZKC_OP_FAR_CALL);
    const uint32_t abi_top = (uint32_t)IN(ZKC_VM_SRC0 + 8);   // limb 7 of src0: [ergs | forwarding | shard | constructor | system] bytes 28..31
    const uint32_t target_high = ((uint32_t)IN(ZKC_VM_SRC1 + 1) >> 16) | (uint32_t)IN(ZKC_VM_SRC1 + 2) | (uint32_t)IN(ZKC_VM_SRC1 + 3) |
                                 (uint32_t)IN(ZKC_VM_SRC1 + 4) | (uint32_t)IN(ZKC_VM_SRC1 + 5);
    const uint32_t system_call = (((abi_top >> 24) & 0xFFu) != 0u) & (target_high == 0u);
    const uint32_t constructor_call = (((abi_top >> 16) & 0xFFu) != 0u) & (ZKC_ROW_LD(&st->current_context.is_kernel_mode) & 1u);
    const uint32_t cleanup = far_call & (system_call ^ 1u);
    const uint32_t far_return = KIND(ZKC_OP_RET) & ((ZKC_ROW_LD(&st->current_context.is_local_call) & 1u) ^ 1u);
    const uint32_t r2_low = constructor_call + 2u * system_call;
    OUT(ZKC_VMW_FAR_CALL_UPDATE, 0) = far_call;
    OUT(ZKC_VMW_FAR_CALL_NON_SYSTEM, 0) = system_call ^ 1u;
    OUT(ZKC_VMW_FAR_CALL_CLEANUP_REGISTER, 0) = cleanup;
    OUT(ZKC_VMW_FAR_RETURN_UPDATE, 0) = far_return;
    OUT(ZKC_VMW_FAR_CALL_NEW_R2_LOW, 0) = r2_low;
    // ---- the two dot products of the DENSE trace and their targets ---------------------------------------------------------------
    uint32_t d0[9], d1[9];
#pragma unroll
    for (int w = 0; w < 9; w++) {
        d0[w] = (uint32_t)IN(ZKC_VM_DST0 + w) & (w ? 0xFFFFFFFFu : 1u);
        d1[w] = (uint32_t)IN(ZKC_VM_DST1 + w) & (w ? 0xFFFFFFFFu : 1u);
    }
    const uint32_t i_dst0 = (uint32_t)IN(ZKC_VM_DST0_REG), i_dst1 = (uint32_t)IN(ZKC_VM_DST1_REG);
    const uint32_t far_lists = lists.system_abi | lists.reserved;
    // ---- the 15 registers (cycle.rs:322-433) -------------------------------------------------------------------------------------------
#pragma unroll
    for (uint32_t r = 0; r < ZKC_VM_REGISTERS; r++) {
        const uint32_t write0 = update_register & (i_dst0 == r + 1u), write1 = i_dst1 == r + 1u;
        const uint32_t listed = (far_lists >> r) & 1u, in_abi = (lists.system_abi >> r) & 1u, in_reserved = (lists.reserved >> r) & 1u;
        const uint32_t not_r1 = r != 0u;
        const uint32_t marker = (listed & far_call) | (not_r1 & far_return);
        const uint32_t zero_out = (in_abi & cleanup) | (in_reserved & far_call) | (not_r1 & far_return);
        const uint32_t call_sets = far_call & (r < 2u), return_sets = far_return & (r == 0u);
        const uint32_t call_hint = call_sets & (uint32_t)(r == 0u);   // r1 of a far call / far return: the next snapshot's register
        const uint32_t *before = &st->registers[r].is_pointer, *after = &nx->registers[r].is_pointer;
        // is_pointer: dot product over (write0, dst0), (far call, new r), (far return, new r1), (marker, false); select; dst1 likewise
        const uint32_t new_marker = r == 0u ? (ZKC_ROW_LD(after) & 1u) : 0u;   // new r2 of a far call is an integer
        const uint32_t as0 = (write0 & d0[0]) + (call_hint & new_marker) + (return_sets & new_marker);   // a dot product: a sum, not an OR
        const uint32_t any0 = write0 | call_sets | return_sets | marker;
        const uint32_t after0 = blend32(any0, as0, ZKC_ROW_LD(before) & 1u);
        const uint32_t as1 = write1 & d1[0];
        OUT(ZKC_VMW_WRITE_AS_DST0, r) = write0;
        OUT(ZKC_VMW_REMOVE_PTR_MARKER, r) = marker;
        OUT(ZKC_VMW_ZERO_OUT, r) = zero_out;
        OUT(ZKC_VMW_ANY_PTR_UPDATE_AS_DST0, r) = any0;
        OUT(ZKC_VMW_IS_PTR_AS_DST0, r) = as0;
        OUT(ZKC_VMW_IS_PTR_AFTER_DST0, r) = after0;
        OUT(ZKC_VMW_IS_PTR_AS_DST1, r) = as1;
        OUT(ZKC_VMW_IS_PTR_AFTER_DST1, r) = blend32(write1, as1, after0);
#pragma unroll
        for (int w = 0; w < 8; w++) {
            uint32_t v = blend32(write0, d0[1 + w], ZKC_ROW_LD(before + 1 + w));
            OUT(ZKC_VMW_VALUE_AFTER_DST0, 8 * r + w) = v;
            if (r < 2u) {
                const uint32_t set = r == 0u ? ZKC_ROW_LD(after + 1 + w) : (w == 0 ? r2_low : 0u);
                v = blend32(call_sets, set, v);
                OUT(ZKC_VMW_VALUE_AFTER_FAR_CALL, 8 * r + w) = v;
            }
            if (r == 0u) {
                v = blend32(return_sets, ZKC_ROW_LD(after + 1 + w), v);
                OUT(ZKC_VMW_VALUE_AFTER_FAR_RETURN, w) = v;
            }
            v = blend32(zero_out, 0u, v);
            OUT(ZKC_VMW_VALUE_AFTER_ZERO_OUT, 8 * r + w) = v;
            v = blend32(write1, d1[1 + w], v);
            OUT(ZKC_VMW_VALUE_AFTER_DST1, 8 * r + w) = v;
        }
    }
#undef KIND
#undef OUT
#undef IN
}

}  // namespace zkc
