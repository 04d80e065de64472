// One row of the create_prestate block (include/zkc_b200.h, ZKC_VM_PRESTATE_COLUMNS): the cells the cycle's preamble allocates on its
// way to the values the DENSE trace names.
//   create_prestate                              /root/reference/src/main_vm/pre_state.rs:71-519
//   split_pc, should_read_memory                 /root/reference/src/main_vm/utils.rs:23-120
//   resolve_memory_region_and_index_for_source   /root/reference/src/main_vm/utils.rs:237-305
//   resolve_memory_region_and_index_for_dest     /root/reference/src/main_vm/utils.rs:307-386
//   reg_idx_into_bitspread + spread_into_bits    /root/reference/src/main_vm/decoded_opcode.rs:192-202
// The function is __host__ __device__ so that the SAME statement the kernel runs can be compiled by g++ and compared with the oracle
// where there is no GPU (tests/cpp/prestate_row_host.cpp, tests/test_prestate_row_host.py); vm_prestate_kernel
// (main_vm_prestate.cu) is a one-thread-per-cycle wrapper around it.  Nothing branches on the opcode: every select is a mask blend,
// so a warp's 32 cycles stay converged whatever they execute.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "../../include/zkc_b200.h"

#if defined(__CUDACC__)
#define ZKC_ROW_FN __host__ __device__ __forceinline__
#else
#define ZKC_ROW_FN static inline
#endif
#if defined(__CUDA_ARCH__)
#define ZKC_ROW_LD(p) __ldg(p)
#else
#define ZKC_ROW_LD(p) (*(p))
#endif

namespace zkc {

struct prestate_reg {   // VMRegister as nine 32-bit cells: is_pointer, value limbs (base_structures/register/mod.rs:21-24)
    uint32_t w[9];
};

// b ? x : y on a 0 / 1 flag without a branch
ZKC_ROW_FN uint32_t blend32(uint32_t flag, uint32_t x, uint32_t y) { const uint32_t m = 0u - (flag & 1u); return (x & m) | (y & ~m); }

// t: column 0 of this row in the DENSE trace (stride limit); st: the snapshot the cycle starts from; out: column 0 of this row in the
// block (stride limit)
ZKC_ROW_FN void vm_prestate_row(const uint64_t *t, const zkc_vm_state *st, uint64_t *out, size_t limit) {
#define IN(col) ZKC_ROW_LD(t + (size_t)(col) * limit)
#define OUT(col, i) out[(size_t)((col) + (i)) * limit]
    const uint64_t props = IN(ZKC_VM_PROPS);
#define PROP(n) ((uint32_t)(props >> (n)) & 1u)
    const zkc_vm_context *cx = &st->current_context;

    // ---- cycle control (pre_state.rs:88-156) -------------------------------------------------------------------------------
    const uint32_t skip = (uint32_t)IN(ZKC_VM_SHOULD_SKIP_CYCLE) & 1u, pending = (uint32_t)IN(ZKC_VM_PENDING_EXCEPTION_IN) & 1u;
    const uint32_t execute = skip ^ 1u;
    OUT(ZKC_VMP_EXECUTE_CYCLE, 0) = execute;
    OUT(ZKC_VMP_SHOULD_TRY_TO_READ_OPCODE, 0) = execute & (pending ^ 1u);
    OUT(ZKC_VMP_PENDING_EXCEPTION_TAKEN_DOWN, 0) = pending & (pending ^ 1u);    // the flag masked by itself: always 0
    const uint32_t pc_sum = ZKC_ROW_LD(&cx->pc) + 1u;
    OUT(ZKC_VMP_PC_PLUS_ONE, 0) = pc_sum & 0xFFFFu;
    OUT(ZKC_VMP_PC_PLUS_ONE_OF, 0) = pc_sum >> 16;
    const uint32_t same_page = ZKC_ROW_LD(&st->previous_code_page) == ZKC_ROW_LD(&cx->code_page);
    const uint32_t same_super_pc = (uint32_t)IN(ZKC_VM_SUPER_PC) == ZKC_ROW_LD(&st->previous_super_pc);
    OUT(ZKC_VMP_CODE_PAGES_ARE_EQUAL, 0) = same_page;
    OUT(ZKC_VMP_SUPER_PC_ARE_EQUAL, 0) = same_super_pc;
    OUT(ZKC_VMP_CAN_SKIP_READ, 0) = same_page & same_super_pc;
    OUT(ZKC_VMP_SHOULD_READ_FOR_NEW_PC, 0) = (same_page & same_super_pc) ^ 1u;
    const uint32_t ts = ZKC_ROW_LD(&st->timestamp);
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) OUT(ZKC_VMP_TIMESTAMPS, k) = (uint32_t)(ts + 1u + k);   // increment_unchecked: 32-bit wrap
    OUT(ZKC_VMP_NEXT_CYCLE_TIMESTAMP, 0) = blend32(skip, ts, ts + 4u);

    // ---- the 64-bit opcode inside the 256-bit code word (pre_state.rs:183-214) ---------------------------------------------
    {
        const uint32_t sub_pc = (uint32_t)IN(ZKC_VM_SUB_PC) & 3u;
        uint32_t lo = (uint32_t)IN(ZKC_VM_CODE_WORD + 6), hi = (uint32_t)IN(ZKC_VM_CODE_WORD + 7);
#pragma unroll
        for (uint32_t k = 0; k < 3; k++) {   // VMSubPCToBitmaskTable: sub_pc s > 0 sets bit s - 1 (tables/integer_to_boolean_mask.rs)
            const uint32_t bit = sub_pc == k + 1u;
            lo = blend32(bit, (uint32_t)IN(ZKC_VM_CODE_WORD + 4 - 2 * k), lo);
            hi = blend32(bit, (uint32_t)IN(ZKC_VM_CODE_WORD + 5 - 2 * k), hi);
            OUT(ZKC_VMP_SUBPC_BITMASK, k) = bit;
            OUT(ZKC_VMP_OPCODE_SELECT_CHAIN, 2 * k) = lo;
            OUT(ZKC_VMP_OPCODE_SELECT_CHAIN, 2 * k + 1) = hi;
        }
    }

    // ---- register selectors, the three 15-step select chains (decoded_opcode.rs:192-202, pre_state.rs:303-329) --------------
    const uint32_t i_src0 = (uint32_t)IN(ZKC_VM_SRC0_REG), i_src1 = (uint32_t)IN(ZKC_VM_SRC1_REG), i_dst0 = (uint32_t)IN(ZKC_VM_DST0_REG),
                   i_dst1 = (uint32_t)IN(ZKC_VM_DST1_REG);
    prestate_reg draft_src0, src1_reg;
#pragma unroll
    for (int w = 0; w < 9; w++) draft_src0.w[w] = src1_reg.w[w] = 0u;
    uint32_t dst0_low = 0u;
#pragma unroll
    for (uint32_t r = 0; r < ZKC_VM_REGISTERS; r++) {   // register r + 1 is selected by index r + 1; index 0 selects nothing
        const uint32_t s0 = i_src0 == r + 1u, s1 = i_src1 == r + 1u, d0 = i_dst0 == r + 1u, d1 = i_dst1 == r + 1u;
        OUT(ZKC_VMP_SRC0_SELECTORS, r) = s0; OUT(ZKC_VMP_SRC1_SELECTORS, r) = s1;
        OUT(ZKC_VMP_DST0_SELECTORS, r) = d0; OUT(ZKC_VMP_DST1_SELECTORS, r) = d1;
        const uint32_t *reg = &st->registers[r].is_pointer;   // is_pointer, value[8]: nine consecutive words
#pragma unroll
        for (int w = 0; w < 9; w++) {
            const uint32_t x = ZKC_ROW_LD(reg + w) & (w ? 0xFFFFFFFFu : 1u);
            draft_src0.w[w] = blend32(s0, x, draft_src0.w[w]);
            src1_reg.w[w] = blend32(s1, x, src1_reg.w[w]);
            if (w == 1) dst0_low = blend32(d0, x, dst0_low);
            OUT(ZKC_VMP_DRAFT_SRC0_CHAIN, 9 * r + w) = draft_src0.w[w];
            OUT(ZKC_VMP_SRC1_REGISTER_CHAIN, 9 * r + w) = src1_reg.w[w];
        }
        OUT(ZKC_VMP_DST0_REG_LOW_CHAIN, r) = dst0_low;
    }
    const uint32_t src0_lowest = draft_src0.w[1] & 0xFFFFu, dst0_lowest = dst0_low & 0xFFFFu;   // low_u16, :310 / :329
    OUT(ZKC_VMP_SRC0_REG_LOWEST, 0) = src0_lowest;
    OUT(ZKC_VMP_DST0_REG_LOWEST, 0) = dst0_lowest;
    {
        const uint32_t base_page = ZKC_ROW_LD(&cx->base_page);                 // :341-343, increment_unchecked three times
        OUT(ZKC_VMP_STACK_PAGE, 0) = (uint32_t)(base_page + 1u);
        OUT(ZKC_VMP_HEAP_PAGE, 0) = (uint32_t)(base_page + 2u);
        OUT(ZKC_VMP_AUX_HEAP_PAGE, 0) = (uint32_t)(base_page + 3u);
    }

    // ---- operand locations (utils.rs:237-386) ---------------------------------------------------------------------------------
    const uint32_t imm0 = (uint32_t)IN(ZKC_VM_IMM0), imm1 = (uint32_t)IN(ZKC_VM_IMM1), sp = ZKC_ROW_LD(&cx->sp) & 0xFFFFu;
    OUT(ZKC_VMP_NOT_NOP, 0) = PROP(ZKC_VM_BIT_TYPE(ZKC_OP_NOP)) ^ 1u;
    uint32_t sp_after_src0;
    {
        const uint32_t code = PROP(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_CODE_PAGE)), absolute = PROP(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_STACK_ABSOLUTE)),
                       relative = PROP(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_STACK_OFFSET)), push_pop = PROP(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_STACK_PUSH_POP));
        const uint32_t index_abs = (src0_lowest + imm0) & 0xFFFFu, index_rel = (sp - index_abs) & 0xFFFFu;   // UInt16 overflowing add / sub
        const uint32_t stack = absolute | relative | push_pop;
        OUT(ZKC_VMP_SRC_ABSOLUTE_MODE, 0) = code | absolute;
        OUT(ZKC_VMP_SRC_INDEX_FOR_ABSOLUTE, 0) = index_abs;
        OUT(ZKC_VMP_SRC_INDEX_FOR_RELATIVE, 0) = index_rel;
        OUT(ZKC_VMP_SRC_USE_STACK, 0) = stack;
        OUT(ZKC_VMP_SRC_DID_READ_UNMASKED, 0) = stack | code;
        sp_after_src0 = blend32(push_pop, index_rel, sp);
    }
    {
        const uint32_t absolute = PROP(ZKC_VM_BIT_DST_MODE(ZKC_MODE_STACK_ABSOLUTE)), relative = PROP(ZKC_VM_BIT_DST_MODE(ZKC_MODE_STACK_OFFSET)),
                       push_pop = PROP(ZKC_VM_BIT_DST_MODE(ZKC_MODE_STACK_PUSH_POP));
        const uint32_t index_abs = (dst0_lowest + imm1) & 0xFFFFu;
        const uint32_t index_push = (sp_after_src0 + index_abs) & 0xFFFFu, index_rel = (sp_after_src0 - index_abs) & 0xFFFFu;
        OUT(ZKC_VMP_DST_INDEX_FOR_ABSOLUTE, 0) = index_abs;
        OUT(ZKC_VMP_DST_INDEX_FOR_RELATIVE_WITH_PUSH, 0) = index_push;
        OUT(ZKC_VMP_DST_INDEX_FOR_RELATIVE, 0) = index_rel;
        OUT(ZKC_VMP_DST_DID_WRITE_UNMASKED, 0) = absolute | relative | push_pop;
        OUT(ZKC_VMP_DST_INDEX_SOMEWHAT_RELATIVE, 0) = blend32(push_pop, sp_after_src0, index_rel);   // a push writes at the current sp
    }

    // ---- src0 selects, operand swap, pointer erasure (pre_state.rs:403-479) --------------------------------------------------
    {
        const uint32_t use_reg = PROP(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_REG_ONLY)), use_imm = PROP(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_IMM16));
        const uint32_t swap = (uint32_t)IN(ZKC_VM_SWAP_OPERANDS) & 1u;
        uint32_t a_is_pointer = 0u, b_is_pointer = 0u;
#pragma unroll
        for (int w = 0; w < 9; w++) {
            const uint32_t from_memory = (uint32_t)IN(ZKC_VM_SRC0_FROM_MEMORY + w) & (w ? 0xFFFFFFFFu : 1u);
            const uint32_t after_reg = blend32(use_reg, draft_src0.w[w], from_memory);
            const uint32_t after_imm = blend32(use_imm, w == 1 ? imm0 : 0u, after_reg);   // VMRegister::from_imm: limb 0 = imm, not a pointer
            const uint32_t a = blend32(swap, src1_reg.w[w], after_imm), b = blend32(swap, after_imm, src1_reg.w[w]);
            OUT(ZKC_VMP_SRC0_AFTER_USE_REG, w) = after_reg;
            OUT(ZKC_VMP_SRC0_AFTER_USE_IMM, w) = after_imm;
            OUT(ZKC_VMP_SRC0_SWAPPED, w) = a;
            OUT(ZKC_VMP_SRC1_SWAPPED, w) = b;
            if (w == 0) { a_is_pointer = a; b_is_pointer = b; }
        }
        const uint32_t is_ptr = PROP(ZKC_VM_BIT_TYPE(ZKC_OP_PTR));
        const uint32_t asymmetric = PROP(ZKC_VM_BIT_TYPE(ZKC_OP_SUB)) | PROP(ZKC_VM_BIT_TYPE(ZKC_OP_DIV)) | PROP(ZKC_VM_BIT_TYPE(ZKC_OP_SHIFT));
        OUT(ZKC_VMP_SWAP_IS_ASSYMMETRIC, 0) = asymmetric;
        OUT(ZKC_VMP_SWAP_T0, 0) = asymmetric & PROP(ZKC_VM_BIT_FLAG(ZKC_VM_SWAP_OPERANDS_FLAG_IDX));
        OUT(ZKC_VMP_SWAP_T1, 0) = is_ptr & PROP(ZKC_VM_BIT_FLAG(ZKC_VM_SWAP_OPERANDS_PTR_FLAG_IDX));
        const uint32_t not_kernel = (ZKC_ROW_LD(&cx->is_kernel_mode) & 1u) ^ 1u;
        const uint32_t keeps = PROP(ZKC_VM_BIT_TYPE(ZKC_OP_RET)) | is_ptr | PROP(ZKC_VM_BIT_TYPE(ZKC_OP_UMA)) | PROP(ZKC_VM_BIT_TYPE(ZKC_OP_FAR_CALL));
        OUT(ZKC_VMP_NOT_KERNEL_MODE, 0) = not_kernel;
        OUT(ZKC_VMP_KEEPS_POINTERS, 0) = keeps;
        OUT(ZKC_VMP_SHOULD_ERASE, 0) = keeps ^ 1u;
        OUT(ZKC_VMP_SHOULD_ERASE_SRC0, 0) = a_is_pointer & (keeps ^ 1u) & not_kernel;
        OUT(ZKC_VMP_SHOULD_ERASE_SRC1, 0) = b_is_pointer & not_kernel;
    }
#undef PROP
#undef OUT
#undef IN
}

}  // namespace zkc
