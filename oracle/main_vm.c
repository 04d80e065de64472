/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Sequential CPU restatement of the main VM circuit, value level, for the opcode subset the engine builds
 * (nop, add, sub, jump, binop, mul, div, shifts, ptr, context + every addressing mode of src0 / dst0):
 *   main_vm_entry_point            /root/reference/src/main_vm/mod.rs:47-232
 *   initial_bootloader_state       /root/reference/src/main_vm/loading.rs:13-226
 *   vm_cycle                       /root/reference/src/main_vm/cycle.rs:28-795
 *   create_prestate                /root/reference/src/main_vm/pre_state.rs:71-519
 *   perform_initial_decoding       /root/reference/src/main_vm/decoded_opcode.rs:42-220, :395-527
 *   memory helpers                 /root/reference/src/main_vm/utils.rs:14-522, cycle.rs:799-935
 *   opcodes                        /root/reference/src/main_vm/opcodes/{nop,add_sub,jump,binop,mul_div,shifts,ptr,context}.rs
 *   ExecutionContextRecord::encode /root/reference/src/base_structures/vm_state/saved_context.rs:111-270
 * log / near_call / far_call / ret / uma are NOT restated yet: a cycle that decodes to one of them (this includes
 * every exception, which the circuit masks into ret.panic) reports ZKC_VM_CHK_UNSUPPORTED_OPCODE.
 * PARITY UNPINNED: the reference has no main_vm test and the ISA tables (zkevm_opcode_defs) are un-vendored; the
 * tables are input data (zkc_vm_isa) and the bit layout follows main_vm/opcode_bitmask.rs:83-127.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* ---- 256-bit helpers on little-endian u32 limbs ------------------------------------------------------ */
static int u256_is_zero(const uint32_t *a) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a[i]; return o == 0; }
static int u256_add(const uint32_t *a, const uint32_t *b, uint32_t *c) {
    uint64_t carry = 0;
    for (int i = 0; i < 8; i++) { const uint64_t t = (uint64_t)a[i] + b[i] + carry; c[i] = (uint32_t)t; carry = t >> 32; }
    return (int)carry;
}
static int u256_sub(const uint32_t *a, const uint32_t *b, uint32_t *c) {
    uint64_t borrow = 0;
    for (int i = 0; i < 8; i++) { const uint64_t t = (uint64_t)a[i] - b[i] - borrow; c[i] = (uint32_t)t; borrow = (t >> 32) & 1; }
    return (int)borrow;
}
static void u256_mul(const uint32_t *a, const uint32_t *b, uint32_t *lo, uint32_t *hi) {
    uint32_t r[16] = {0};
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 8; j++) {
            const uint64_t t = (uint64_t)a[i] * b[j] + r[i + j] + carry;
            r[i + j] = (uint32_t)t; carry = t >> 32;
        }
        r[i + 8] = (uint32_t)carry;
    }
    memcpy(lo, r, 32); memcpy(hi, r + 8, 32);
}
static int u256_ge(const uint32_t *a, const uint32_t *b) {
    for (int i = 7; i >= 0; i--) if (a[i] != b[i]) return a[i] > b[i];
    return 1;
}
/* q = a / b, r = a % b (b != 0); plain binary long division */
static void u256_divrem(const uint32_t *a, const uint32_t *b, uint32_t *q, uint32_t *r) {
    memset(q, 0, 32); memset(r, 0, 32);
    for (int bit = 255; bit >= 0; bit--) {
        uint32_t top = r[7] >> 31;
        for (int i = 7; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 31);
        r[0] = (r[0] << 1) | ((a[bit / 32] >> (bit % 32)) & 1);
        if (top || u256_ge(r, b)) { uint32_t t[8]; u256_sub(r, b, t); memcpy(r, t, 32); q[bit / 32] |= 1u << (bit % 32); }
    }
}

/* ---- state flattening (CSVarLengthEncodable order, vm_state/mod.rs:92-109) ----------------------------- */
size_t orc_vm_flatten_context_record(const zkc_vm_context *c, uint64_t *dst) {
    size_t n = 0;
    for (int i = 0; i < 5; i++) dst[n++] = c->this_address[i];
    for (int i = 0; i < 5; i++) dst[n++] = c->caller[i];
    for (int i = 0; i < 5; i++) dst[n++] = c->code_address[i];
    dst[n++] = c->code_page; dst[n++] = c->base_page; dst[n++] = c->heap_upper_bound; dst[n++] = c->aux_heap_upper_bound;
    for (int i = 0; i < 4; i++) dst[n++] = c->reverted_queue_head[i];
    for (int i = 0; i < 4; i++) dst[n++] = c->reverted_queue_tail[i];
    dst[n++] = c->reverted_queue_segment_len;
    dst[n++] = c->pc; dst[n++] = c->sp; dst[n++] = c->exception_handler_loc; dst[n++] = c->ergs_remaining;
    dst[n++] = c->is_static_execution; dst[n++] = c->is_kernel_mode;
    dst[n++] = c->this_shard_id; dst[n++] = c->caller_shard_id; dst[n++] = c->code_shard_id;
    for (int i = 0; i < 4; i++) dst[n++] = c->context_u128_value_composite[i];
    dst[n++] = c->is_local_call;
    return n; /* 42 */
}
size_t orc_vm_flatten_state(const zkc_vm_state *s, uint64_t *dst) {
    size_t n = 0;
    for (int i = 0; i < 8; i++) dst[n++] = s->previous_code_word[i];
    for (int r = 0; r < 15; r++) { dst[n++] = s->registers[r].is_pointer; for (int i = 0; i < 8; i++) dst[n++] = s->registers[r].value[i]; }
    for (int i = 0; i < 3; i++) dst[n++] = s->flags[i];
    dst[n++] = s->timestamp; dst[n++] = s->memory_page_counter; dst[n++] = s->tx_number_in_block; dst[n++] = s->previous_code_page;
    dst[n++] = s->previous_super_pc; dst[n++] = s->pending_exception; dst[n++] = s->ergs_per_pubdata_byte;
    n += orc_vm_flatten_context_record(&s->current_context, dst + n);
    for (int i = 0; i < 4; i++) dst[n++] = s->current_context.log_queue_forward_tail[i];
    dst[n++] = s->current_context.log_queue_forward_part_length;
    dst[n++] = s->context_stack_depth;
    for (int i = 0; i < 12; i++) dst[n++] = s->stack_sponge_state[i];
    for (int i = 0; i < 12; i++) dst[n++] = s->memory_queue_state[i];
    dst[n++] = s->memory_queue_length;
    for (int i = 0; i < 12; i++) dst[n++] = s->code_decommittment_queue_state[i];
    dst[n++] = s->code_decommittment_queue_length;
    for (int i = 0; i < 4; i++) dst[n++] = s->context_composite_u128[i];
    return n; /* 243 */
}

/* saved_context.rs:111-270 */
void orc_vm_context_encode(const zkc_vm_context *c, uint64_t e[32]) {
    for (int i = 0; i < 4; i++) { e[i] = c->reverted_queue_head[i]; e[4 + i] = c->reverted_queue_tail[i]; }
    for (int i = 0; i < 5; i++) { e[8 + i] = c->code_address[i]; e[13 + i] = c->this_address[i]; e[18 + i] = c->caller[i]; }
    for (int i = 0; i < 4; i++) e[23 + i] = c->context_u128_value_composite[i];
    e[27] = (uint64_t)c->code_page + ((uint64_t)c->pc << 32) + ((uint64_t)c->this_shard_id << 48) + ((uint64_t)c->is_static_execution << 56);
    e[28] = (uint64_t)c->base_page + ((uint64_t)c->sp << 32) + ((uint64_t)c->caller_shard_id << 48) + ((uint64_t)c->is_kernel_mode << 56);
    e[29] = (uint64_t)c->ergs_remaining + ((uint64_t)c->exception_handler_loc << 32) + ((uint64_t)c->code_shard_id << 48) + ((uint64_t)c->is_local_call << 56);
    const uint32_t sl = c->reverted_queue_segment_len;
    e[30] = (uint64_t)c->heap_upper_bound + ((uint64_t)(sl & 0xFF) << 32) + ((uint64_t)((sl >> 8) & 0xFF) << 40);
    e[31] = (uint64_t)c->aux_heap_upper_bound + ((uint64_t)((sl >> 16) & 0xFF) << 32) + ((uint64_t)(sl >> 24) << 40);
}

/* loading.rs:13-226 */
void orc_vm_initial_bootloader_state(const zkc_vm_closed_form *io, const zkc_vm_isa *isa, zkc_vm_state *st) {
    memset(st, 0, sizeof *st);
    zkc_vm_context *ctx = &st->current_context;
    ctx->base_page = isa->bootloader_base_page;
    ctx->code_page = isa->bootloader_code_page;
    ctx->exception_handler_loc = isa->initial_frame_formal_eh_location;
    ctx->ergs_remaining = isa->vm_initial_frame_ergs;
    ctx->code_address[0] = isa->bootloader_formal_address_low;
    ctx->this_address[0] = isa->bootloader_formal_address_low;
    memcpy(ctx->reverted_queue_tail, io->rollback_queue_tail_for_block, 32);
    memcpy(ctx->reverted_queue_head, io->rollback_queue_tail_for_block, 32);
    ctx->is_kernel_mode = 1;
    ctx->heap_upper_bound = isa->bootloader_max_memory;
    ctx->aux_heap_upper_bound = isa->bootloader_max_memory;
    zkc_vm_context empty;
    memset(&empty, 0, sizeof empty);
    memcpy(empty.reverted_queue_tail, io->rollback_queue_tail_for_block, 32);
    memcpy(empty.reverted_queue_head, io->rollback_queue_tail_for_block, 32);
    empty.is_kernel_mode = 1;
    uint64_t enc[32], s[12] = {0};
    orc_vm_context_encode(&empty, enc);
    for (int r = 0; r < 4; r++) { memcpy(s, enc + 8 * r, 64); orc_poseidon2_permutation(s); }
    memcpy(st->stack_sponge_state, s, 96);
    st->context_stack_depth = 1;
    st->memory_queue_length = io->memory_queue_initial_length;
    memcpy(st->memory_queue_state, io->memory_queue_initial_tail, 96);
    st->code_decommittment_queue_length = io->decommitment_queue_initial_length;
    memcpy(st->code_decommittment_queue_state, io->decommitment_queue_initial_tail, 96);
    st->timestamp = isa->starting_timestamp;
    st->memory_page_counter = isa->starting_base_page;
    /* r1 = formal fat pointer {offset 0, page CALLDATA, start 0, length 0} */
    st->registers[0].is_pointer = 1;
    st->registers[0].value[1] = isa->bootloader_calldata_page;
}

/* ---- memory model of the out-of-circuit run ------------------------------------------------------------- */
typedef struct orc_vm_memory {
    zkc_vm_register *code;  /* 2^16 words */
    zkc_vm_register *stack; /* 2^16 words */
    uint32_t code_page, stack_page;
} orc_vm_memory;

static void memq_push(uint64_t state[12], uint32_t *len, uint32_t ts, uint32_t page, uint32_t index, uint32_t rw,
                      const zkc_vm_register *v, int execute) {
    if (!execute) return;
    zkc_memory_query q;
    memset(&q, 0, sizeof q);
    q.timestamp = ts; q.memory_page = page; q.index = index; q.rw_flag = rw; q.is_ptr = v->is_pointer & 1;
    memcpy(q.value, v->value, 32);
    uint64_t enc[8];
    orc_memory_query_encode(&q, enc);
    memcpy(state, enc, 64);
    orc_poseidon2_permutation(state);
    (*len)++;
}

static int prop(uint64_t props, int bit) { return (int)((props >> bit) & 1); }

#define T(col) row[(size_t)(col) * stride]

/* one vm_cycle.  mem != NULL: out-of-circuit run (memory reads answered by the model and RECORDED into *w);
 * mem == NULL: witness-driven (reads answered from *w).  row/stride: trace row or NULL.  Returns check bits. */
static uint32_t vm_cycle(const zkc_vm_isa *isa, const zkc_vm_state *cur, zkc_vm_cycle_witness *w, orc_vm_memory *mem,
                         zkc_vm_state *out, uint64_t *row, size_t stride) {
    uint32_t checks = 0;
    zkc_vm_state s = *cur;
    zkc_vm_context *ctx = &s.current_context;
    /* ---------------- create_prestate, pre_state.rs:71-519 ---------------- */
    const int should_skip = s.context_stack_depth == 0;
    const int pending = (int)s.pending_exception;
    const int should_try_read = !should_skip && !pending;
    s.pending_exception = 0;
    const uint32_t pc = ctx->pc, pc_plus_one = (pc + 1) & 0xFFFF, super_pc = pc >> 2, sub_pc = pc & 3;
    const int should_read_new = !(s.previous_code_page == ctx->code_page && super_pc == s.previous_super_pc);
    const int should_read_opcode = should_try_read && should_read_new;
    const uint32_t ts0 = s.timestamp, ts_dst = ts0 + 3;
    const uint32_t next_ts = should_skip ? ts0 : ts0 + 4;
    zkc_vm_register code_val;
    memset(&code_val, 0, sizeof code_val);
    if (should_read_opcode) {
        if (mem) { code_val = mem->code[super_pc]; code_val.is_pointer = 0; memcpy(w->code_word, code_val.value, 32); }
        else memcpy(code_val.value, w->code_word, 32);
    } else if (mem) memset(w->code_word, 0, 32);
    memq_push(s.memory_queue_state, &s.memory_queue_length, ts0, ctx->code_page, super_pc, 0, &code_val, should_read_opcode);
    uint32_t code_word[8];
    memcpy(code_word, should_read_opcode ? code_val.value : s.previous_code_word, 32);
    uint32_t op_lo = code_word[6 - 2 * sub_pc], op_hi = code_word[7 - 2 * sub_pc]; /* :185-206 */
    if (should_skip) { op_lo = (uint32_t)isa->nop_opcode_encoding; op_hi = (uint32_t)(isa->nop_opcode_encoding >> 32); }
    if (pending) { op_lo = (uint32_t)isa->panic_opcode_encoding; op_hi = (uint32_t)(isa->panic_opcode_encoding >> 32); }
    if (row) {
        T(ZKC_VM_SHOULD_SKIP_CYCLE) = (uint64_t)should_skip; T(ZKC_VM_PENDING_EXCEPTION_IN) = (uint64_t)pending;
        T(ZKC_VM_SHOULD_READ_OPCODE) = (uint64_t)should_read_opcode; T(ZKC_VM_SUPER_PC) = super_pc; T(ZKC_VM_SUB_PC) = sub_pc;
        for (int i = 0; i < 8; i++) T(ZKC_VM_CODE_WORD + i) = code_word[i];
        for (int i = 0; i < 12; i++) T(ZKC_VM_MEMQ_AFTER_CODE + i) = s.memory_queue_state[i];
        T(ZKC_VM_MEMQ_AFTER_CODE + 12) = s.memory_queue_length;
        T(ZKC_VM_OPCODE) = op_lo; T(ZKC_VM_OPCODE + 1) = op_hi;
    }
    memcpy(s.previous_code_word, code_word, 32);
    s.previous_code_page = ctx->code_page;
    if (!should_skip) { ctx->pc = pc_plus_one; s.previous_super_pc = super_pc; }
    s.timestamp = next_ts;
    const int is_kernel = (int)ctx->is_kernel_mode, is_static = (int)ctx->is_static_execution;
    const int callstack_full = s.context_stack_depth == isa->vm_max_stack_depth;
    /* ---------------- perform_initial_decoding, decoded_opcode.rs:42-220 ---------------- */
    const uint32_t variant = op_lo & 0x7FF, cond_idx = (op_lo >> 13) & 7;
    uint32_t src_regs = (op_lo >> 16) & 0xFF, dst_regs = op_lo >> 24;
    const uint32_t imm0 = op_hi & 0xFFFF, imm1 = op_hi >> 16;
    const uint32_t price = isa->opcode_price[variant];
    const uint64_t props_full = isa->opcode_props[variant];
    uint64_t props = props_full & ((1ULL << ZKC_VM_DESCRIPTION_BITS_FLATTENED) - 1);
    const uint32_t aux = (uint32_t)(props_full >> ZKC_VM_DESCRIPTION_BITS_FLATTENED);
    const uint32_t encoded_flags = (s.flags[0] & 1) | ((s.flags[1] & 1) << 1) | ((s.flags[2] & 1) << 2);
    const int condition = isa->condition_table[cond_idx][encoded_flags];
    const uint32_t cost = should_skip ? 0 : price;
    const int out_of_ergs = ctx->ergs_remaining < cost;
    const uint32_t ergs_left = out_of_ergs ? 0 : ctx->ergs_remaining - cost;
    const int requires_kernel = (aux >> ZKC_VM_AUX_KERNEL_MODE) & 1, can_static = (aux >> ZKC_VM_AUX_CAN_BE_USED_IN_STATIC) & 1;
    const int explicit_panic = (aux >> ZKC_VM_AUX_EXPLICIT_PANIC) & 1;
    const int kernel_exc = requires_kernel && !is_kernel, static_exc = is_static && !can_static;
    const int mask_into_panic = explicit_panic || out_of_ergs || kernel_exc || static_exc || callstack_full;
    if (mask_into_panic) props = isa->panic_bitspread & ((1ULL << ZKC_VM_DESCRIPTION_BITS_FLATTENED) - 1);
    const int mask_into_nop = !mask_into_panic && !condition;
    if (mask_into_nop) props = isa->nop_bitspread & ((1ULL << ZKC_VM_DESCRIPTION_BITS_FLATTENED) - 1);
    if (mask_into_nop || mask_into_panic) { src_regs = 0; dst_regs = 0; }
    const uint32_t src0_r = src_regs & 15, src1_r = src_regs >> 4, dst0_r = dst_regs & 15, dst1_r = dst_regs >> 4;
    ctx->ergs_remaining = ergs_left;
    if (prop(props, ZKC_VM_BIT_TYPE(ZKC_OP_INVALID))) checks |= ZKC_VM_CHK_INVALID_OPCODE;
    if (row) {
        T(ZKC_VM_VARIANT) = variant; T(ZKC_VM_CONDITION_IDX) = cond_idx; T(ZKC_VM_CONDITION) = (uint64_t)condition;
        T(ZKC_VM_ERGS_COST) = cost; T(ZKC_VM_OUT_OF_ERGS) = (uint64_t)out_of_ergs; T(ZKC_VM_KERNEL_MODE_EXCEPTION) = (uint64_t)kernel_exc;
        T(ZKC_VM_STATIC_EXCEPTION) = (uint64_t)static_exc; T(ZKC_VM_CALLSTACK_IS_FULL) = (uint64_t)callstack_full;
        T(ZKC_VM_EXPLICIT_PANIC) = (uint64_t)explicit_panic; T(ZKC_VM_MASK_INTO_PANIC) = (uint64_t)mask_into_panic;
        T(ZKC_VM_MASK_INTO_NOP) = (uint64_t)mask_into_nop; T(ZKC_VM_PROPS) = props; T(ZKC_VM_DIRTY_ERGS_LEFT) = ergs_left;
        T(ZKC_VM_SRC0_REG) = src0_r; T(ZKC_VM_SRC1_REG) = src1_r; T(ZKC_VM_DST0_REG) = dst0_r; T(ZKC_VM_DST1_REG) = dst1_r;
        T(ZKC_VM_IMM0) = imm0; T(ZKC_VM_IMM1) = imm1;
    }
#define TYPE(t) prop(props, ZKC_VM_BIT_TYPE(t))
#define VAR(v) prop(props, ZKC_VM_BIT_VARIANT(v))
#define FLAG(f) prop(props, ZKC_VM_BIT_FLAG(f))
#define SRCM(m) prop(props, ZKC_VM_BIT_SRC_MODE(m))
#define DSTM(m) prop(props, ZKC_VM_BIT_DST_MODE(m))
    /* ---------------- operands, pre_state.rs:301-472 ---------------- */
    zkc_vm_register zero_reg;
    memset(&zero_reg, 0, sizeof zero_reg);
    const zkc_vm_register draft_src0 = src0_r ? s.registers[src0_r - 1] : zero_reg;
    const zkc_vm_register src1_register = src1_r ? s.registers[src1_r - 1] : zero_reg;
    const uint32_t src0_reg_lowest = draft_src0.value[0] & 0xFFFF;
    const uint32_t dst0_reg_lowest = (dst0_r ? s.registers[dst0_r - 1].value[0] : 0) & 0xFFFF;
    const uint32_t current_sp = ctx->sp, code_page = ctx->code_page;
    const uint32_t stack_page = ctx->base_page + 1, heap_page = ctx->base_page + 2, aux_heap_page = ctx->base_page + 3;
    (void)heap_page; (void)aux_heap_page;
    const int is_nop = TYPE(ZKC_OP_NOP);
    /* resolve_memory_region_and_index_for_source, utils.rs:237-305 */
    uint32_t src_page, src_index, sp_after_src0;
    int should_read_src0;
    {
        const int use_code = SRCM(ZKC_MODE_CODE_PAGE), abs_ = SRCM(ZKC_MODE_STACK_ABSOLUTE), rel = SRCM(ZKC_MODE_STACK_OFFSET), pp = SRCM(ZKC_MODE_STACK_PUSH_POP);
        const uint32_t idx_abs = (src0_reg_lowest + imm0) & 0xFFFF, idx_rel = (current_sp - idx_abs) & 0xFFFF;
        const int use_stack = abs_ || rel || pp;
        should_read_src0 = (use_stack || use_code) && !is_nop;
        src_page = use_stack ? stack_page : code_page;
        src_index = (use_code || abs_) ? idx_abs : idx_rel;
        sp_after_src0 = pp ? idx_rel : current_sp;
    }
    /* resolve_memory_region_and_index_for_dest, utils.rs:307-386 */
    uint32_t dst_page = stack_page, dst_index, new_sp;
    int dst0_mem;
    {
        const int abs_ = DSTM(ZKC_MODE_STACK_ABSOLUTE), rel = DSTM(ZKC_MODE_STACK_OFFSET), pp = DSTM(ZKC_MODE_STACK_PUSH_POP);
        const uint32_t idx_abs = (dst0_reg_lowest + imm1) & 0xFFFF;
        const uint32_t idx_rel_push = (sp_after_src0 + idx_abs) & 0xFFFF, idx_rel = (sp_after_src0 - idx_abs) & 0xFFFF;
        dst0_mem = (abs_ || rel || pp) && !is_nop;
        const uint32_t somewhat = pp ? sp_after_src0 : idx_rel;
        dst_index = abs_ ? idx_abs : somewhat;
        new_sp = pp ? idx_rel_push : sp_after_src0;
    }
    ctx->sp = new_sp;
    /* may_be_read_memory_for_source_operand, utils.rs:388-522 */
    zkc_vm_register src0_mem;
    memset(&src0_mem, 0, sizeof src0_mem);
    if (should_read_src0) {
        if (mem) {
            if (src_page == mem->code_page) { src0_mem = mem->code[src_index]; src0_mem.is_pointer = 0; }
            else if (src_page == mem->stack_page) src0_mem = mem->stack[src_index];
            w->src0_is_pointer = src0_mem.is_pointer; memcpy(w->src0_value, src0_mem.value, 32);
        } else { src0_mem.is_pointer = w->src0_is_pointer & 1; memcpy(src0_mem.value, w->src0_value, 32); }
    } else if (mem) { w->src0_is_pointer = 0; memset(w->src0_value, 0, 32); }
    memq_push(s.memory_queue_state, &s.memory_queue_length, ts0, src_page, src_index, 0, &src0_mem, should_read_src0);
    if (row) {
        T(ZKC_VM_SRC0_PAGE) = src_page; T(ZKC_VM_SRC0_INDEX) = src_index; T(ZKC_VM_SHOULD_READ_SRC0) = (uint64_t)should_read_src0;
        T(ZKC_VM_SP_AFTER_SRC0) = sp_after_src0; T(ZKC_VM_DST0_PAGE) = dst_page; T(ZKC_VM_DST0_INDEX) = dst_index;
        T(ZKC_VM_DST0_PERFORMS_MEMORY_ACCESS) = (uint64_t)dst0_mem; T(ZKC_VM_NEW_SP) = new_sp;
        T(ZKC_VM_SRC0_FROM_MEMORY) = src0_mem.is_pointer; for (int i = 0; i < 8; i++) T(ZKC_VM_SRC0_FROM_MEMORY + 1 + i) = src0_mem.value[i];
        for (int i = 0; i < 12; i++) T(ZKC_VM_MEMQ_AFTER_SRC0 + i) = s.memory_queue_state[i];
        T(ZKC_VM_MEMQ_AFTER_SRC0 + 12) = s.memory_queue_length;
    }
    zkc_vm_register src0 = SRCM(ZKC_MODE_REG_ONLY) ? draft_src0 : src0_mem;
    if (SRCM(ZKC_MODE_IMM16)) { src0 = zero_reg; src0.value[0] = imm0; }
    const int is_ptr_op = TYPE(ZKC_OP_PTR);
    const int swap = ((TYPE(ZKC_OP_SUB) || TYPE(ZKC_OP_DIV) || TYPE(ZKC_OP_SHIFT)) && FLAG(ZKC_VM_SWAP_OPERANDS_FLAG_IDX)) ||
                     (is_ptr_op && FLAG(ZKC_VM_SWAP_OPERANDS_PTR_FLAG_IDX));
    zkc_vm_register a = swap ? src1_register : src0, b = swap ? src0 : src1_register;
    {
        const int keep = TYPE(ZKC_OP_RET) || is_ptr_op || TYPE(ZKC_OP_UMA) || TYPE(ZKC_OP_FAR_CALL);
        const int erase0 = a.is_pointer && !keep && !is_kernel, erase1 = b.is_pointer && !is_kernel;
        if (erase0) { a.is_pointer = 0; a.value[1] = 0; a.value[2] = 0; }
        if (erase1) { b.is_pointer = 0; b.value[1] = 0; b.value[2] = 0; }
    }
    if (row) {
        T(ZKC_VM_SWAP_OPERANDS) = (uint64_t)swap;
        T(ZKC_VM_SRC0) = a.is_pointer; T(ZKC_VM_SRC1) = b.is_pointer;
        for (int i = 0; i < 8; i++) { T(ZKC_VM_SRC0 + 1 + i) = a.value[i]; T(ZKC_VM_SRC1 + 1 + i) = b.value[i]; }
    }
    /* ---------------- opcodes (cycle.rs:73-156): only the selected one matters at value level ---------------- */
    zkc_vm_register dst0 = zero_reg, dst1 = zero_reg;
    int dst0_to_mem_capable = 0, dst0_reg_only = 0, write_dst1 = 0;
    int set_flags = 0;
    uint32_t nf[3] = {0, 0, 0};
    int new_pending = 0;
    if (TYPE(ZKC_OP_NEAR_CALL) || TYPE(ZKC_OP_LOG) || TYPE(ZKC_OP_FAR_CALL) || TYPE(ZKC_OP_RET) || TYPE(ZKC_OP_UMA)) checks |= ZKC_VM_CHK_UNSUPPORTED_OPCODE;
    const int sf = FLAG(ZKC_VM_SET_FLAGS_FLAG_IDX);
    if (TYPE(ZKC_OP_ADD) || TYPE(ZKC_OP_SUB)) { /* add_sub.rs:8-166 */
        const int of = TYPE(ZKC_OP_ADD) ? u256_add(a.value, b.value, dst0.value) : u256_sub(a.value, b.value, dst0.value);
        const int z = u256_is_zero(dst0.value);
        nf[0] = (uint32_t)of; nf[1] = (uint32_t)z; nf[2] = (uint32_t)!(of || z);
        set_flags = sf; dst0_to_mem_capable = 1;
    }
    if (TYPE(ZKC_OP_JUMP)) ctx->pc = a.value[0] & 0xFFFF; /* jump.rs:3-38 */
    if (TYPE(ZKC_OP_BINOP)) { /* binop.rs:14-121 */
        for (int i = 0; i < 8; i++)
            dst0.value[i] = VAR(ZKC_VAR_BINOP_OR) ? (a.value[i] | b.value[i]) : VAR(ZKC_VAR_BINOP_AND) ? (a.value[i] & b.value[i]) : (a.value[i] ^ b.value[i]);
        nf[1] = (uint32_t)u256_is_zero(dst0.value);
        set_flags = sf; dst0_to_mem_capable = 1;
    }
    if (TYPE(ZKC_OP_MUL)) { /* mul_div.rs:199-417 */
        u256_mul(a.value, b.value, dst0.value, dst1.value);
        const int of = !u256_is_zero(dst1.value), eq = u256_is_zero(dst0.value);
        nf[0] = (uint32_t)of; nf[1] = (uint32_t)eq; nf[2] = (uint32_t)(!of && !eq);
        set_flags = sf; dst0_to_mem_capable = 1; write_dst1 = 1;
    }
    if (TYPE(ZKC_OP_DIV)) {
        const int dz = u256_is_zero(b.value);
        if (!dz) u256_divrem(a.value, b.value, dst0.value, dst1.value); /* divisor 0: quotient 0, remainder masked to 0 */
        nf[0] = (uint32_t)dz; nf[1] = (uint32_t)(!dz && u256_is_zero(dst0.value)); nf[2] = (uint32_t)(!dz && u256_is_zero(dst1.value));
        set_flags = sf; dst0_to_mem_capable = 1; write_dst1 = 1;
    }
    if (TYPE(ZKC_OP_SHIFT)) { /* shifts.rs:8-198 */
        const int is_rol = VAR(ZKC_VAR_SHIFT_ROL), is_ror = VAR(ZKC_VAR_SHIFT_ROR), is_shr = VAR(ZKC_VAR_SHIFT_SHR);
        const int cyclic = is_rol || is_ror, right = (is_ror || is_shr) && !cyclic;
        uint32_t shift = b.value[0] & 0xFF;
        if (is_ror && shift != 0) shift = 256 - shift;
        uint32_t pw[8] = {0}, lo[8], hi[8], q[8], r[8];
        pw[shift / 32] = 1u << (shift % 32);
        u256_mul(a.value, pw, lo, hi);
        u256_divrem(a.value, pw, q, r);
        for (int i = 0; i < 8; i++) dst0.value[i] = (right ? q[i] : lo[i]) + (cyclic ? hi[i] : 0);
        nf[1] = (uint32_t)u256_is_zero(dst0.value);
        set_flags = sf; dst0_to_mem_capable = 1;
    }
    if (is_ptr_op) { /* ptr.rs:6-183 */
        const int v_add = VAR(ZKC_VAR_PTR_ADD), v_sub = VAR(ZKC_VAR_PTR_SUB), v_pack = VAR(ZKC_VAR_PTR_PACK), v_shrink = VAR(ZKC_VAR_PTR_SHRINK);
        const int invalid_types = !(a.is_pointer && !b.is_pointer);
        int hi_nz = 0, lo_nz = 0;
        for (int i = 1; i < 8; i++) hi_nz |= b.value[i] != 0;
        for (int i = 0; i < 4; i++) lo_nz |= b.value[i] != 0;
        const int too_large = hi_nz && (v_add || v_sub), dirty_pack = lo_nz && v_pack;
        const uint64_t addr = (uint64_t)a.value[0] + b.value[0];
        const int of_add = v_add && (addr >> 32), uf_sub = v_sub && a.value[0] < b.value[0], uf_shrink = v_shrink && a.value[3] < b.value[0];
        const int panic = invalid_types || too_large || dirty_pack || of_add || uf_sub || uf_shrink;
        new_pending = panic;
        dst0 = a;
        if (v_add) dst0.value[0] = (uint32_t)addr;
        if (v_sub) dst0.value[0] = a.value[0] - b.value[0];
        if (v_shrink) dst0.value[3] = a.value[3] - b.value[0];
        if (v_pack) for (int i = 4; i < 8; i++) dst0.value[i] = b.value[i];
        dst0_to_mem_capable = !panic;
    }
    if (TYPE(ZKC_OP_CONTEXT)) { /* context.rs:7-307 */
        const int set_u128 = VAR(ZKC_VAR_CONTEXT_SET_U128), set_pubdata = VAR(ZKC_VAR_CONTEXT_SET_ERGS_PER_PUBDATA), inc_tx = VAR(ZKC_VAR_CONTEXT_INC_TX_NUMBER);
        dst0.value[0] = VAR(ZKC_VAR_CONTEXT_ERGS_LEFT) ? ergs_left : new_sp;
        if (VAR(ZKC_VAR_CONTEXT_GET_U128)) memcpy(dst0.value, ctx->context_u128_value_composite, 16);
        if (VAR(ZKC_VAR_CONTEXT_THIS)) memcpy(dst0.value, ctx->this_address, 20);
        if (VAR(ZKC_VAR_CONTEXT_CALLER)) memcpy(dst0.value, ctx->caller, 20);
        if (VAR(ZKC_VAR_CONTEXT_CODE_ADDRESS)) memcpy(dst0.value, ctx->code_address, 20);
        if (VAR(ZKC_VAR_CONTEXT_META)) {
            memset(dst0.value, 0, 32);
            dst0.value[0] = s.ergs_per_pubdata_byte; dst0.value[2] = ctx->heap_upper_bound; dst0.value[3] = ctx->aux_heap_upper_bound;
            dst0.value[7] = ctx->this_shard_id | (ctx->caller_shard_id << 8) | (ctx->code_shard_id << 16);
        }
        dst0_reg_only = !(set_u128 || set_pubdata || inc_tx);
        if (set_u128) memcpy(s.context_composite_u128, a.value, 16);
        if (set_pubdata) s.ergs_per_pubdata_byte = a.value[0];
        if (inc_tx) s.tx_number_in_block = s.tx_number_in_block + 1;
    }
    /* ---------------- state diffs, cycle.rs:158-616 ---------------- */
    const int perform_mem_write = dst0_mem && dst0_to_mem_capable;
    memq_push(s.memory_queue_state, &s.memory_queue_length, ts_dst, dst_page, dst_index, 1, &dst0, perform_mem_write);
    if (mem && perform_mem_write && dst_page == mem->stack_page) mem->stack[dst_index] = dst0;
    const int dst0_update_register = dst0_reg_only || (!dst0_mem && dst0_to_mem_capable);
    if (dst0_update_register && dst0_r) s.registers[dst0_r - 1] = dst0;
    if (write_dst1 && dst1_r) s.registers[dst1_r - 1] = dst1; /* dst1 applied after dst0, cycle.rs:421-433 */
    if (set_flags) memcpy(s.flags, nf, sizeof nf);
    s.pending_exception = (uint32_t)new_pending;
    s.memory_page_counter = cur->memory_page_counter; /* only far calls move it */
    if (row) {
        T(ZKC_VM_DST0) = dst0.is_pointer; T(ZKC_VM_DST1) = dst1.is_pointer;
        for (int i = 0; i < 8; i++) { T(ZKC_VM_DST0 + 1 + i) = dst0.value[i]; T(ZKC_VM_DST1 + 1 + i) = dst1.value[i]; }
        T(ZKC_VM_PERFORM_DST0_MEMORY_WRITE) = (uint64_t)perform_mem_write; T(ZKC_VM_DST0_UPDATE_REGISTER) = (uint64_t)dst0_update_register;
        for (int i = 0; i < 12; i++) T(ZKC_VM_MEMQ_AFTER_DST0 + i) = s.memory_queue_state[i];
        T(ZKC_VM_MEMQ_AFTER_DST0 + 12) = s.memory_queue_length;
        for (int i = 0; i < 3; i++) T(ZKC_VM_FLAGS_OUT + i) = s.flags[i];
        T(ZKC_VM_PENDING_EXCEPTION_OUT) = s.pending_exception; T(ZKC_VM_PC_OUT) = ctx->pc; T(ZKC_VM_ERGS_OUT) = ctx->ergs_remaining;
    }
    *out = s;
    return checks;
}

/* status code = most specific aggregate: a broken snapshot chain outranks an unsupported opcode outranks a failed
 * enforcement (an order-independent rule, so that the row-parallel engine reports the same code) */
static void fail(zkc_status *st, int64_t row, uint32_t bits) {
    st->failed_checks |= bits;
    st->code = (st->failed_checks & ZKC_VM_CHK_SNAPSHOT) ? ZKC_ERR_SNAPSHOT_MISMATCH
             : (st->failed_checks & ZKC_VM_CHK_UNSUPPORTED_OPCODE) ? ZKC_ERR_UNSUPPORTED : ZKC_ERR_UNSATISFIED;
    if (row >= 0 && (st->first_bad_row < 0 || row < st->first_bad_row)) st->first_bad_row = row;
}

/* out-of-circuit run: fills snapshots[cycles + 1] and witness[cycles]; code: [code_words][8] */
int orc_main_vm_run(const zkc_vm_isa *isa, const zkc_vm_state *initial, const uint32_t *code, size_t code_words,
                    size_t cycles, zkc_vm_state *snapshots, zkc_vm_cycle_witness *witness, zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    orc_vm_memory mem;
    mem.code = calloc(65536, sizeof(zkc_vm_register));
    mem.stack = calloc(65536, sizeof(zkc_vm_register));
    mem.code_page = initial->current_context.code_page;
    mem.stack_page = initial->current_context.base_page + 1;
    for (size_t i = 0; i < code_words && i < 65536; i++) memcpy(mem.code[i].value, code + 8 * i, 32);
    snapshots[0] = *initial;
    for (size_t c = 0; c < cycles; c++) {
        memset(&witness[c], 0, sizeof witness[c]);
        const uint32_t chk = vm_cycle(isa, &snapshots[c], &witness[c], &mem, &snapshots[c + 1], NULL, 0);
        if (chk) fail(&st, (int64_t)c, chk);
    }
    free(mem.code); free(mem.stack);
    if (status) *status = st;
    return st.code;
}

int orc_main_vm_entry_point(zkc_vm_closed_form *io, const zkc_vm_isa *isa, const zkc_vm_state *snapshots,
                            const zkc_vm_cycle_witness *witness, size_t limit, const zkc_vm_options *options,
                            uint64_t *trace, uint64_t commitment[4], zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    const int start = io->start_flag != 0;
    zkc_vm_state state;
    if (start) orc_vm_initial_bootloader_state(io, isa, &state);
    else state = io->hidden_fsm_input;
    uint64_t fa[243], fb[243];
    for (size_t c = 0; c < limit; c++) {
        /* the per-cycle snapshot is a hint: it must be the state the sequential run is in */
        orc_vm_flatten_state(&state, fa); orc_vm_flatten_state(&snapshots[c], fb);
        if (memcmp(fa, fb, sizeof fa)) fail(&st, (int64_t)c, ZKC_VM_CHK_SNAPSHOT);
        zkc_vm_cycle_witness w = witness[c];
        zkc_vm_state next;
        const uint32_t chk = vm_cycle(isa, &snapshots[c], &w, NULL, &next, trace ? trace + c : NULL, limit);
        if (chk) fail(&st, (int64_t)c, chk);
        state = next;
    }
    orc_vm_flatten_state(&state, fa); orc_vm_flatten_state(&snapshots[limit], fb);
    if (memcmp(fa, fb, sizeof fa)) fail(&st, (int64_t)limit - 1, ZKC_VM_CHK_SNAPSHOT);
    /* mod.rs:113-196 */
    const int done = state.context_stack_depth == 0;
    if (done && state.current_context.pc != 0) fail(&st, -1, ZKC_VM_CHK_BOOTLOADER_EXIT);
    zkc_queue_state4 log_out; zkc_queue_state12 mem_out, dec_out;
    memset(&log_out, 0, sizeof log_out); memset(&mem_out, 0, sizeof mem_out); memset(&dec_out, 0, sizeof dec_out);
    if (done) {
        memcpy(mem_out.tail, state.memory_queue_state, 96); mem_out.length = state.memory_queue_length;
        memcpy(dec_out.tail, state.code_decommittment_queue_state, 96); dec_out.length = state.code_decommittment_queue_length;
        memcpy(log_out.tail, state.current_context.log_queue_forward_tail, 32); log_out.length = state.current_context.log_queue_forward_part_length;
    }
    if (options && options->compare_expected) {
        orc_vm_flatten_state(&io->hidden_fsm_output, fb);
        if (memcmp(fa, fb, sizeof fa) || memcmp(&log_out, &io->log_queue_final_state, sizeof log_out) ||
            memcmp(&mem_out, &io->memory_queue_final_state, sizeof mem_out) || memcmp(&dec_out, &io->decommitment_queue_final_state, sizeof dec_out) ||
            (io->completion_flag != 0) != done)
            if (st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    /* commitments: VmInputData / VmOutputData / VmLocalState encodings, circuit_inputs/main_vm.rs:9-49 */
    uint64_t e_in[39], e_out[59], e_fin[243];
    size_t n = 0;
    for (int i = 0; i < 4; i++) e_in[n++] = io->rollback_queue_tail_for_block[i];
    for (int i = 0; i < 12; i++) e_in[n++] = io->memory_queue_initial_tail[i];
    e_in[n++] = io->memory_queue_initial_length;
    for (int i = 0; i < 12; i++) e_in[n++] = io->decommitment_queue_initial_tail[i];
    e_in[n++] = io->decommitment_queue_initial_length;
    e_in[n++] = io->zkporter_is_available;
    for (int i = 0; i < 8; i++) e_in[n++] = io->default_aa_code_hash[i];
    const size_t n_in = n; /* 39 */
    n = orc_put_queue_state4(e_out, &log_out);
    memcpy(e_out + n, mem_out.head, 96); n += 12; memcpy(e_out + n, mem_out.tail, 96); n += 12; e_out[n++] = mem_out.length;
    memcpy(e_out + n, dec_out.head, 96); n += 12; memcpy(e_out + n, dec_out.tail, 96); n += 12; e_out[n++] = dec_out.length;
    const size_t n_out = n; /* 59 */
    orc_vm_flatten_state(&io->hidden_fsm_input, e_fin);
    io->hidden_fsm_output = state;
    io->log_queue_final_state = log_out; io->memory_queue_final_state = mem_out; io->decommitment_queue_final_state = dec_out;
    io->completion_flag = (uint32_t)done;
    orc_closed_form_commitment(start, done, e_in, n_in, e_out, n_out, e_fin, 243, fa, 243, commitment);
    if (status) *status = st;
    return st.code;
}
