/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Sequential CPU restatement of the main VM circuit, value level (every opcode gadget: nop, add, sub, jump, binop, mul,
 * div, shifts, ptr, context, uma, log, near_call, far_call, ret + every addressing mode of src0 / dst0):
 *   main_vm_entry_point            /root/reference/src/main_vm/mod.rs:47-232
 *   initial_bootloader_state       /root/reference/src/main_vm/loading.rs:13-226
 *   vm_cycle                       /root/reference/src/main_vm/cycle.rs:28-795
 *   create_prestate                /root/reference/src/main_vm/pre_state.rs:71-519
 *   perform_initial_decoding       /root/reference/src/main_vm/decoded_opcode.rs:42-220, :395-527
 *   memory helpers                 /root/reference/src/main_vm/utils.rs:14-522, cycle.rs:799-935
 *   opcodes                        /root/reference/src/main_vm/opcodes/{nop,add_sub,jump,binop,mul_div,shifts,ptr,context}.rs
 *   uma                            /root/reference/src/main_vm/opcodes/uma.rs:18-1103
 *   log                            /root/reference/src/main_vm/opcodes/log.rs:16-671
 *   near_call / ret / callstack    /root/reference/src/main_vm/opcodes/call_ret.rs:24-512, call_ret_impl/{near_call.rs:34-184,
 *                                  ret.rs:29-479, mod.rs:38-86, far_call.rs:140-262 (FatPtrInABI)}
 *   ExecutionContextRecord::encode /root/reference/src/base_structures/vm_state/saved_context.rs:111-270
 *   far_call                       /root/reference/src/main_vm/opcodes/call_ret_impl/far_call.rs:268-1603
 *   DecommitQuery::encode          /root/reference/src/base_structures/decommit_query/mod.rs:31-107
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

/* ---- memory / storage model and bookkeeping of the out-of-circuit run ------------------------------------------ */
#define ORC_VM_PAGE_WORDS 65536
#define ORC_VM_STORAGE_SLOTS 4096
#define ORC_VM_MEM_CELLS (1u << 20)
typedef struct orc_vm_slot { uint32_t used, written, addr0; uint32_t key[8]; uint32_t value[8]; } orc_vm_slot;
typedef struct orc_vm_cell { uint32_t used, page, index; zkc_vm_register v; } orc_vm_cell;
/* one rollback-queue event of a frame: its own call marker or a revertable log */
typedef struct orc_vm_entry {
    int64_t prev;
    uint32_t kind;              /* 1 call marker, 2 log */
    int32_t slot;               /* storage slot a storage write touched, or -1 */
    uint32_t prev_value[8];     /* its value (and written marker) before the write */
    uint32_t prev_written;
    uint64_t enc16[4], cap[4];  /* rollback packing elements 16..19 and the sponge capacity after round 1 */
} orc_vm_entry;
typedef struct orc_vm_sim {
    orc_vm_cell *cells;         /* memory: (page, index) -> word, every page of every frame (index < 2^16) */
    orc_vm_slot *storage;
    const uint32_t *code; size_t code_words;  /* the one program of this world: every deployed address runs it */
    uint32_t code_hash[8];      /* its versioned hash */
    uint32_t decommitted_page;  /* page the program was first decommitted to (0: not yet) */
    zkc_vm_callstack_witness *stack; /* saved frames, [max_depth] */
    size_t max_depth;
    zkc_vm_callstack_witness *cw_out; size_t cw_cap, n_cw;
    /* what the cycle reports to the driver (rollback bookkeeping) */
    int ev_kind;                /* 0 none, 1 call, 2 ret ok, 3 ret revert / panic, 4 revertable log */
    orc_vm_entry ev;
    int overflow;               /* a model limit was hit (depth, storage slots, memory cells, callstack witness capacity) */
} orc_vm_sim;

static orc_vm_cell *sim_cell(orc_vm_sim *m, uint32_t page, uint32_t index, int create) {
    uint32_t h = (page * 0x9E3779B1u) ^ (index * 0x85EBCA77u);
    h ^= h >> 15;
    for (uint32_t probe = 0; probe < ORC_VM_MEM_CELLS; probe++) {
        orc_vm_cell *c = &m->cells[(h + probe) & (ORC_VM_MEM_CELLS - 1)];
        if (!c->used) {
            if (!create) return NULL;
            c->used = 1; c->page = page; c->index = index;
            return c;
        }
        if (c->page == page && c->index == index) return c;
    }
    m->overflow = 1;
    return NULL;
}
static zkc_vm_register sim_read(orc_vm_sim *m, uint32_t page, uint32_t index) {
    zkc_vm_register z;
    memset(&z, 0, sizeof z);
    if (index >= ORC_VM_PAGE_WORDS) return z;
    const orc_vm_cell *c = sim_cell(m, page, index, 0);
    return c ? c->v : z;
}
static void sim_write(orc_vm_sim *m, uint32_t page, uint32_t index, const zkc_vm_register *v) {
    if (index >= ORC_VM_PAGE_WORDS) return;
    orc_vm_cell *c = sim_cell(m, page, index, 1);
    if (c) c->v = *v;
}
static void sim_load_code(orc_vm_sim *m, uint32_t page) {
    zkc_vm_register r;
    memset(&r, 0, sizeof r);
    for (size_t i = 0; i < m->code_words && i < ORC_VM_PAGE_WORDS; i++) { memcpy(r.value, m->code + 8 * i, 32); sim_write(m, page, (uint32_t)i, &r); }
}
static int sim_slot(orc_vm_sim *m, uint32_t addr0, const uint32_t key[8]) {
    uint32_t h = 0x9E3779B9u ^ (addr0 * 0x27D4EB2Fu);
    for (int i = 0; i < 8; i++) h = (h ^ key[i]) * 0x85EBCA6Bu + (h >> 15);
    for (uint32_t probe = 0; probe < ORC_VM_STORAGE_SLOTS; probe++) {
        orc_vm_slot *s = &m->storage[(h + probe) % ORC_VM_STORAGE_SLOTS];
        if (!s->used) { s->used = 1; s->addr0 = addr0; memcpy(s->key, key, 32); return (int)((h + probe) % ORC_VM_STORAGE_SLOTS); }
        if (s->addr0 == addr0 && !memcmp(s->key, key, 32)) return (int)((h + probe) % ORC_VM_STORAGE_SLOTS);
    }
    m->overflow = 1;
    return 0;
}

static void mq_encode(uint32_t ts, uint32_t page, uint32_t index, uint32_t rw, const zkc_vm_register *v, uint64_t enc[8]) {
    zkc_memory_query q;
    memset(&q, 0, sizeof q);
    q.timestamp = ts; q.memory_page = page; q.index = index; q.rw_flag = rw; q.is_ptr = v->is_pointer & 1;
    memcpy(q.value, v->value, 32);
    orc_memory_query_encode(&q, enc);
}

static int prop(uint64_t props, int bit) { return (int)((props >> bit) & 1); }

/* the nine Poseidon2 relations of a cycle (cycle.rs:620-795): slot -> enforced flag + permutation output */
typedef struct vm_sponges { int enf[ZKC_VM_NUM_SPONGES]; uint64_t fin[ZKC_VM_NUM_SPONGES][12]; } vm_sponges;
/* absorb-with-replacement of 8 elements over the capacity cap[4] */
static void sponge_run(vm_sponges *sp, int slot, const uint64_t in8[8], const uint64_t cap[4]) {
    uint64_t s[12];
    memcpy(s, in8, 64); memcpy(s + 8, cap, 32);
    orc_poseidon2_permutation(s);
    sp->enf[slot] = 1; memcpy(sp->fin[slot], s, 96);
}
/* memory queue push (main_vm/utils.rs:194-230, :442-515, cycle.rs:845-905, uma.rs:362-520, :682-812) */
static void memq_push(vm_sponges *sp, int slot, uint64_t state[12], uint32_t *len, uint32_t ts, uint32_t page, uint32_t index,
                      uint32_t rw, const zkc_vm_register *v, int execute) {
    if (!execute) return;
    uint64_t enc[8];
    mq_encode(ts, page, index, rw, v, enc);
    sponge_run(sp, slot, enc, state + 8);
    memcpy(state, sp->fin[slot], 96);
    (*len)++;
}

/* FatPtrInABI::parse_and_validate, far_call.rs:140-196 */
typedef struct vm_fat_ptr { uint32_t offset, page, start, length; } vm_fat_ptr;
static vm_fat_ptr fat_ptr_parse(const uint32_t v[8], int as_fresh, uint32_t *upper_bound, int *generally_invalid, int *non_addressable) {
    vm_fat_ptr p = {v[0], v[1], v[2], v[3]};
    const uint64_t end = (uint64_t)p.start + p.length;
    const int range_of = (int)(end >> 32), invalid_slice = p.length < p.offset;
    const int invalid = (p.offset != 0 && as_fresh) || range_of || invalid_slice;
    if (invalid) { p.offset = 0; p.page = 0; p.start = 0; p.length = 0; }
    *upper_bound = (uint32_t)end; *generally_invalid = invalid; *non_addressable = range_of;
    return p;
}

static void flatten_record_cols(const zkc_vm_context *c, uint64_t *row, size_t stride, int col) {
    uint64_t f[42];
    orc_vm_flatten_context_record(c, f);
    for (int i = 0; i < 42; i++) row[(size_t)(col + i) * stride] = f[i];
}

#define T(col) row[(size_t)(col) * stride]

/* one vm_cycle.  sim != NULL: out-of-circuit run (oracle answers come from the model and are RECORDED into *w / the
 * callstack witness); sim == NULL: witness-driven (answers come from *w / cw).  row/stride: trace row or NULL.
 * Returns check bits. */
static uint32_t vm_cycle(const zkc_vm_isa *isa, const zkc_vm_closed_form *gc, const zkc_vm_state *cur, zkc_vm_cycle_witness *w,
                         const zkc_vm_callstack_witness *cw, size_t n_cw, orc_vm_sim *sim, zkc_vm_state *out, uint64_t *row, size_t stride) {
    uint32_t checks = 0;
    zkc_vm_state s = *cur;
    zkc_vm_context *ctx = &s.current_context;
    vm_sponges sp;
    memset(&sp, 0, sizeof sp);
    if (sim) sim->ev_kind = 0;
    if (row) for (int c = 0; c < ZKC_VM_NUM_COLS; c++) T(c) = 0;
    /* ---------------- create_prestate, pre_state.rs:71-519 ---------------- */
    const int should_skip = s.context_stack_depth == 0;
    const int pending = (int)s.pending_exception;
    const int should_try_read = !should_skip && !pending;
    s.pending_exception = 0;
    const uint32_t pc = ctx->pc, pc_plus_one = (pc + 1) & 0xFFFF, super_pc = pc >> 2, sub_pc = pc & 3;
    const int should_read_new = !(s.previous_code_page == ctx->code_page && super_pc == s.previous_super_pc);
    const int should_read_opcode = should_try_read && should_read_new;
    const uint32_t ts0 = s.timestamp, ts_log = ts0 + 1, ts_dst = ts0 + 3;
    const uint32_t next_ts = should_skip ? ts0 : ts0 + 4;
    zkc_vm_register code_val;
    memset(&code_val, 0, sizeof code_val);
    if (should_read_opcode) {
        if (sim) { code_val = sim_read(sim, ctx->code_page, super_pc); code_val.is_pointer = 0; memcpy(w->code_word, code_val.value, 32); }
        else memcpy(code_val.value, w->code_word, 32);
    } else if (sim) memset(w->code_word, 0, 32);
    memq_push(&sp, 0, s.memory_queue_state, &s.memory_queue_length, ts0, ctx->code_page, super_pc, 0, &code_val, should_read_opcode);
    uint32_t code_word[8];
    memcpy(code_word, should_read_opcode ? code_val.value : s.previous_code_word, 32);
    uint32_t op_lo = code_word[6 - 2 * sub_pc], op_hi = code_word[7 - 2 * sub_pc]; /* :185-206 */
    if (should_skip) { op_lo = (uint32_t)isa->nop_opcode_encoding; op_hi = (uint32_t)(isa->nop_opcode_encoding >> 32); }
    if (pending) { op_lo = (uint32_t)isa->panic_opcode_encoding; op_hi = (uint32_t)(isa->panic_opcode_encoding >> 32); }
    if (row) {
        T(ZKC_VM_SHOULD_SKIP_CYCLE) = (uint64_t)should_skip; T(ZKC_VM_PENDING_EXCEPTION_IN) = (uint64_t)pending;
        T(ZKC_VM_SHOULD_READ_OPCODE) = (uint64_t)should_read_opcode; T(ZKC_VM_SUPER_PC) = super_pc; T(ZKC_VM_SUB_PC) = sub_pc;
        for (int i = 0; i < 8; i++) T(ZKC_VM_CODE_WORD + i) = code_word[i];
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
        if (sim) {
            src0_mem = sim_read(sim, src_page, src_index);
            if (src_page == code_page) src0_mem.is_pointer = 0;
            w->src0_is_pointer = src0_mem.is_pointer; memcpy(w->src0_value, src0_mem.value, 32);
        } else { src0_mem.is_pointer = w->src0_is_pointer & 1; memcpy(src0_mem.value, w->src0_value, 32); }
    } else if (sim) { w->src0_is_pointer = 0; memset(w->src0_value, 0, 32); }
    memq_push(&sp, 1, s.memory_queue_state, &s.memory_queue_length, ts0, src_page, src_index, 0, &src0_mem, should_read_src0);
    if (row) {
        T(ZKC_VM_SRC0_PAGE) = src_page; T(ZKC_VM_SRC0_INDEX) = src_index; T(ZKC_VM_SHOULD_READ_SRC0) = (uint64_t)should_read_src0;
        T(ZKC_VM_SP_AFTER_SRC0) = sp_after_src0; T(ZKC_VM_DST0_PAGE) = dst_page; T(ZKC_VM_DST0_INDEX) = dst_index;
        T(ZKC_VM_DST0_PERFORMS_MEMORY_ACCESS) = (uint64_t)dst0_mem; T(ZKC_VM_NEW_SP) = new_sp;
        T(ZKC_VM_SRC0_FROM_MEMORY) = src0_mem.is_pointer; for (int i = 0; i < 8; i++) T(ZKC_VM_SRC0_FROM_MEMORY + 1 + i) = src0_mem.value[i];
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
    /* state the selected opcode may replace later (cycle.rs:435-610 applies them in this order) */
    uint32_t ergs_candidate = ergs_left;           /* new_ergs_left_candidates */
    int replace_callstack = 0;
    zkc_vm_context new_ctx;                        /* callstacks: the full new current context */
    uint64_t new_stack_sponge[12];
    uint32_t new_depth = s.context_stack_depth;
    int far_return_registers = 0;
    zkc_vm_register far_return_r1 = zero_reg;
    int far_call_registers = 0, far_call_system = 0;   /* far call: r1 / r2 set, ABI / reserved / implicit registers cleaned */
    zkc_vm_register far_call_r2 = zero_reg;
    uint32_t new_memory_page_counter = s.memory_page_counter;
    int decommit_applies = 0;
    uint64_t new_decommit_state[12]; uint32_t new_decommit_len = s.code_decommittment_queue_length;
    memcpy(new_decommit_state, s.code_decommittment_queue_state, 96);
    int reset_context_u128 = 0;
    uint64_t draft_memq[12];
    memcpy(draft_memq, s.memory_queue_state, 96);  /* memory queue after the prestate reads; UMA continues from here */
    uint32_t draft_memq_len = s.memory_queue_length;
    int uma_applies = 0;
    uint64_t uma_memq[12]; uint32_t uma_memq_len = 0;
    if (TYPE(ZKC_OP_UMA)) { /* uma.rs:18-1002 */
        const int heap_r = VAR(ZKC_VAR_UMA_HEAP_READ), heap_w = VAR(ZKC_VAR_UMA_HEAP_WRITE), aux_r = VAR(ZKC_VAR_UMA_AUX_HEAP_READ),
                  aux_w = VAR(ZKC_VAR_UMA_AUX_HEAP_WRITE), ptr_r = VAR(ZKC_VAR_UMA_FAT_PTR_READ);
        const int increment = FLAG(ZKC_VM_UMA_INCREMENT_FLAG_IDX);
        const int access_heap = heap_r || heap_w, access_aux = aux_r || aux_w;
        const int not_a_ptr = ptr_r && !a.is_pointer;
        /* QuasiFatPtrInUMA::parse_and_validate, :1004-1084 */
        const uint32_t offset = a.value[0], page = a.value[1], start = a.value[2], length = a.value[3];
        const int beyond = !(offset < length), skip_legit = beyond && ptr_r;
        const uint32_t abs_addr = (ptr_r ? start : 0) + offset;
        const uint64_t inc64 = (uint64_t)offset + 32;
        const uint32_t incremented = (uint32_t)inc64;
        const int non_addressable = (int)(inc64 >> 32) || incremented == 0xFFFFFFFFu;
        const int qp_panic = not_a_ptr || non_addressable;
        const int qp_skip = not_a_ptr || skip_legit || non_addressable;
        uint32_t bytes_oob = incremented - length;
        if (qp_skip || incremented < length) bytes_oob = 0;
        bytes_oob &= 31;
        /* heap growth, :110-142 */
        const uint32_t heap_bound = ctx->heap_upper_bound, aux_bound = ctx->aux_heap_upper_bound;
        const uint32_t heap_max = access_heap ? incremented : 0, aux_max = access_aux ? incremented : 0;
        const int heap_uf = heap_max < heap_bound, aux_uf = aux_max < aux_bound;
        const uint32_t heap_growth = heap_uf ? 0 : heap_max - heap_bound, aux_growth = aux_uf ? 0 : aux_max - aux_bound;
        const uint32_t new_heap_bound = heap_uf ? heap_bound : heap_max, new_aux_bound = aux_uf ? aux_bound : aux_max;
        uint32_t growth_cost = access_heap ? heap_growth : 0;
        if (access_aux) growth_cost = aux_growth;
        int top_nz = 0;
        for (int i = 1; i < 8; i++) top_nz |= a.value[i] != 0;
        const int exc_oob = (access_heap || access_aux) && (top_nz || non_addressable);
        if (exc_oob) growth_cost = 0xFFFFFFFFu;
        const int uf = ergs_left < growth_cost;
        const int set_panic = qp_panic || uf || exc_oob;
        const uint32_t ergs_after = uf ? 0 : ergs_left - growth_cost;
        const int skip_mem = qp_skip || set_panic;
        const int is_read = heap_r || aux_r || ptr_r, is_write = heap_w || aux_w;
        const uint32_t cell_idx = abs_addr / 32, unalignment = abs_addr % 32;
        const int unaligned = unalignment != 0;
        uint32_t mem_page = page;
        if (access_heap) mem_page = heap_page;
        if (access_aux) mem_page = aux_heap_page;
        const uint32_t b_idx = cell_idx + 1; /* wraps */
        const int read_a = !skip_mem, read_b = unaligned && !skip_mem;
        zkc_vm_register va = zero_reg, vb = zero_reg;
        if (sim) {
            if (read_a) va = sim_read(sim, mem_page, cell_idx);
            if (read_b) vb = sim_read(sim, mem_page, b_idx);
            va.is_pointer = 0; vb.is_pointer = 0;
            memcpy(w->value_a, va.value, 32); memcpy(w->value_b, vb.value, 32);
        } else {
            if (read_a) memcpy(va.value, w->value_a, 32); /* masked to zero when not read, :313, :357 */
            if (read_b) memcpy(vb.value, w->value_b, 32);
        }
        uma_applies = 1;
        memcpy(uma_memq, draft_memq, 96); uma_memq_len = draft_memq_len;
        /* the sponges run when the opcode applies without panic (apply_any, :955-985); skip_mem covers set_panic */
        memq_push(&sp, 1, uma_memq, &uma_memq_len, ts0, mem_page, cell_idx, 0, &va, read_a);
        memq_push(&sp, 2, uma_memq, &uma_memq_len, ts0, mem_page, b_idx, 0, &vb, read_b);
        /* 64-byte big-endian window over cells A, B (:533-560) */
        uint8_t bytes[64], word[32], written[64], wbytes[32];
        for (int i = 0; i < 32; i++) {
            bytes[i] = (uint8_t)(va.value[7 - i / 4] >> (8 * (3 - i % 4)));
            bytes[32 + i] = (uint8_t)(vb.value[7 - i / 4] >> (8 * (3 - i % 4)));
            wbytes[i] = (uint8_t)(b.value[7 - i / 4] >> (8 * (3 - i % 4)));
        }
        memcpy(word, bytes + unalignment, 32);
        /* fat-pointer reads beyond the slice end are zeroed (:562-585): the last bytes_oob bytes */
        const uint32_t cleanup = ptr_r ? bytes_oob : 0;
        for (uint32_t i = 0; i < cleanup; i++) word[31 - i] = 0;
        memcpy(written, bytes, 64);
        memcpy(written + unalignment, wbytes, 32);
        const int exec_write = is_write && !skip_mem, exec_write_b = exec_write && unaligned;
        zkc_vm_register wa = zero_reg, wb = zero_reg;
        for (int i = 0; i < 32; i++) {
            wa.value[7 - i / 4] |= (uint32_t)written[i] << (8 * (3 - i % 4));
            wb.value[7 - i / 4] |= (uint32_t)written[32 + i] << (8 * (3 - i % 4));
        }
        memq_push(&sp, 3, uma_memq, &uma_memq_len, ts_dst, mem_page, cell_idx, 1, &wa, exec_write);
        memq_push(&sp, 4, uma_memq, &uma_memq_len, ts_dst, mem_page, b_idx, 1, &wb, exec_write_b);
        if (sim) {
            if (exec_write) sim_write(sim, mem_page, cell_idx, &wa);
            if (exec_write_b) sim_write(sim, mem_page, b_idx, &wb);
        }
        zkc_vm_register read_reg = zero_reg, inc_reg = a;
        for (int i = 0; i < 32; i++) read_reg.value[7 - i / 4] |= (uint32_t)word[i] << (8 * (3 - i % 4));
        inc_reg.value[0] = incremented;
        const int w_inc = is_write && increment;
        const int no_panic = !set_panic;
        dst0 = w_inc ? inc_reg : read_reg;
        dst0_reg_only = no_panic && (is_read || w_inc);
        write_dst1 = no_panic && is_read && increment;
        dst1 = inc_reg;
        new_pending = set_panic;
        if (access_heap) ctx->heap_upper_bound = new_heap_bound;
        if (access_aux) ctx->aux_heap_upper_bound = new_aux_bound;
        ergs_candidate = ergs_after;
        if (row) {
            uint64_t *x = &T(ZKC_VM_OP_AUX);
            x[0] = abs_addr; x[stride] = cell_idx; x[2 * stride] = unalignment; x[3 * stride] = mem_page; x[4 * stride] = (uint64_t)skip_mem;
            x[5 * stride] = (uint64_t)set_panic; x[6 * stride] = growth_cost; x[7 * stride] = incremented;
            for (int i = 0; i < 8; i++) {
                x[(8 + i) * stride] = va.value[i]; x[(16 + i) * stride] = vb.value[i];
                x[(24 + i) * stride] = exec_write ? wa.value[i] : 0; x[(32 + i) * stride] = exec_write_b ? wb.value[i] : 0;
            }
        }
    }
    if (TYPE(ZKC_OP_LOG)) { /* log.rs:16-467 */
        const int st_read = VAR(ZKC_VAR_LOG_STORAGE_READ), st_write = VAR(ZKC_VAR_LOG_STORAGE_WRITE), is_event = VAR(ZKC_VAR_LOG_EVENT),
                  is_l1 = VAR(ZKC_VAR_LOG_TO_L1_MESSAGE), is_precompile = VAR(ZKC_VAR_LOG_PRECOMPILE_CALL);
        zkc_log_query q;
        memset(&q, 0, sizeof q);
        memcpy(q.address, ctx->this_address, 20);
        memcpy(q.key, a.value, 32);
        if (is_precompile && q.key[4] == 0) q.key[4] = heap_page;
        if (is_precompile && q.key[5] == 0) q.key[5] = heap_page;
        const int is_rollup = ctx->this_shard_id == 0, write_to_rollup = is_rollup && st_write;
        const int is_storage = st_read || st_write, nonrevertable = st_read || is_precompile, revertable = !nonrevertable;
        const uint32_t aux_byte = (is_storage ? isa->log_aux_bytes[0] : 0) + (is_event ? isa->log_aux_bytes[1] : 0) +
                                  (is_l1 ? isa->log_aux_bytes[2] : 0) + (is_precompile ? isa->log_aux_bytes[3] : 0);
        const int is_service = FLAG(ZKC_VM_FIRST_MESSAGE_FLAG_IDX);
        q.tx_number_in_block = s.tx_number_in_block; q.timestamp = ts_log;
        q.flags = ZKC_LQ_FLAGS(aux_byte, ctx->this_shard_id, revertable, 0, is_service);
        memcpy(q.written_value, b.value, 32);
        int slot = -1;
        if (sim) {
            w->refund = 0;
            if (st_write) { slot = sim_slot(sim, q.address[0], q.key); w->refund = sim->storage[slot].written ? isa->initial_storage_write_pubdata_bytes : 0; }
        }
        const uint32_t refund = w->refund;
        if (refund > isa->initial_storage_write_pubdata_bytes) checks |= ZKC_VM_CHK_LOG_REFUND; /* sub_no_overflow, :256 */
        const uint32_t net_cost = isa->initial_storage_write_pubdata_bytes - refund;
        uint32_t burn = write_to_rollup ? s.ergs_per_pubdata_byte * net_cost : 0;
        if (is_precompile) burn = b.value[0];
        if (is_l1) burn = s.ergs_per_pubdata_byte * isa->l1_message_pubdata_bytes;
        const int not_enough = ergs_left < burn;
        const uint32_t ergs_after = not_enough ? 0 : ergs_left - burn;
        const int execute = !not_enough, execute_rollback = execute && revertable;
        uint32_t read_value[8] = {0};
        if (sim) {
            memset(w->value_a, 0, 32);
            if (is_storage && execute) {
                if (slot < 0) slot = sim_slot(sim, q.address[0], q.key);
                memcpy(w->value_a, sim->storage[slot].value, 32);
            }
        }
        if (is_storage) memcpy(read_value, w->value_a, 32);
        memcpy(q.read_value, read_value, 32);
        if (!revertable) memcpy(q.written_value, read_value, 32); /* convention for reads, :328-330 */
        uint64_t enc[20], cap[4] = {0, 0, 0, 0};
        orc_log_query_encode(&q, enc);
        /* construct_hash_relations_for_log_and_new_queue_states, :469-671 */
        uint64_t in8[8];
        if (execute) {
            sponge_run(&sp, 1, enc, cap);
            sponge_run(&sp, 2, enc + 8, sp.fin[1] + 8);
            memcpy(in8, enc + 16, 32); memcpy(in8 + 4, ctx->log_queue_forward_tail, 32);
            sponge_run(&sp, 3, in8, sp.fin[2] + 8);
        }
        if (sim) {
            if (execute_rollback) {
                sim->ev_kind = 4; sim->ev.kind = 2; sim->ev.slot = -1;
                memcpy(sim->ev.enc16, enc + 16, 32); sim->ev.enc16[3] = 1; memcpy(sim->ev.cap, sp.fin[2] + 8, 32);
            }
            if (st_write && execute) {
                orc_vm_slot *sl = &sim->storage[slot];
                sim->ev.slot = slot; memcpy(sim->ev.prev_value, sl->value, 32); sim->ev.prev_written = sl->written;
                memcpy(sl->value, b.value, 32); sl->written = 1;
            }
        }
        if (execute_rollback) {
            memcpy(in8, enc + 16, 32); in8[3] = 1; /* update_packing_for_rollback, log_query/mod.rs:52-58 */
            memcpy(in8 + 4, w->rollback, 32);
            sponge_run(&sp, 4, in8, sp.fin[2] + 8);
            if (memcmp(sp.fin[4], ctx->reverted_queue_head, 32)) checks |= ZKC_VM_CHK_ROLLBACK_QUEUE; /* :627-632 */
            memcpy(ctx->reverted_queue_head, w->rollback, 32);
            ctx->reverted_queue_segment_len++;
        }
        if (execute) { memcpy(ctx->log_queue_forward_tail, sp.fin[3], 32); ctx->log_queue_forward_part_length++; }
        if (st_read) memcpy(dst0.value, read_value, 32);
        else dst0.value[0] = (uint32_t)execute; /* precompile call result; selected for every non-read variant, :396-412 */
        dst0_reg_only = st_read || is_precompile;
        ergs_candidate = ergs_after;
        if (row) {
            uint64_t *x = &T(ZKC_VM_OP_AUX);
            for (int i = 0; i < 20; i++) x[i * stride] = enc[i];
            for (int i = 0; i < 8; i++) x[(20 + i) * stride] = read_value[i];
            x[28 * stride] = (uint64_t)execute; x[29 * stride] = (uint64_t)execute_rollback; x[30 * stride] = burn;
        }
    }
    if (TYPE(ZKC_OP_NEAR_CALL) || TYPE(ZKC_OP_RET) || TYPE(ZKC_OP_FAR_CALL)) { /* call_ret.rs:24-512 */
        const int apply_near = TYPE(ZKC_OP_NEAR_CALL), apply_ret = TYPE(ZKC_OP_RET), apply_far = TYPE(ZKC_OP_FAR_CALL);
        int far_exception = 0;
        /* compute_shared_abi_parts, call_ret_impl/mod.rs:38-86 */
        const uint32_t fwd_byte = a.value[ZKC_VM_ABI_FORWARDING_MODE_BYTE_IDX / 4] >> (8 * (ZKC_VM_ABI_FORWARDING_MODE_BYTE_IDX % 4)) & 0xFF;
        const int use_aux = fwd_byte == ZKC_VM_FORWARD_USE_AUX_HEAP, fwd_ptr = fwd_byte == ZKC_VM_FORWARD_FAT_POINTER;
        const int use_heap = !(use_aux || fwd_ptr);
        uint32_t upper_bound; int generally_invalid, non_addressable;
        vm_fat_ptr fp = fat_ptr_parse(a.value, !fwd_ptr, &upper_bound, &generally_invalid, &non_addressable);
        (void)generally_invalid;
        zkc_vm_context old_entry, new_entry;
        uint64_t sponge_from[12];
        int is_panic_out = 0, perform_revert = 0;
        uint64_t new_fwd_tail[4]; uint32_t new_fwd_len = ctx->log_queue_forward_part_length;
        memcpy(new_fwd_tail, ctx->log_queue_forward_tail, 32);
        if (apply_near) { /* near_call.rs:34-184 */
            zkc_vm_context cur_e = *ctx;
            cur_e.pc = pc_plus_one;
            new_entry = cur_e;
            memcpy(new_entry.reverted_queue_tail, w->rollback, 32);
            memcpy(new_entry.reverted_queue_head, w->rollback, 32);
            new_entry.reverted_queue_segment_len = 0;
            const uint32_t ergs_passed = a.value[0];
            const uint32_t to_pass = ergs_passed == 0 ? ergs_left : ergs_passed;
            const int uf = ergs_left < to_pass;
            cur_e.ergs_remaining = uf ? 0 : ergs_left - to_pass;
            new_entry.ergs_remaining = uf ? ergs_left : to_pass;
            new_entry.pc = imm0; new_entry.exception_handler_loc = imm1; new_entry.is_local_call = 1;
            old_entry = cur_e;
            memcpy(sponge_from, s.stack_sponge_state, 96);
            new_depth = s.context_stack_depth + 1;
            if (sim) {
                if (s.context_stack_depth >= sim->max_depth) sim->overflow = 1;
                else { sim->stack[s.context_stack_depth].context = old_entry; memcpy(sim->stack[s.context_stack_depth].previous_sponge_state, s.stack_sponge_state, 96); }
                sim->ev_kind = 1; sim->ev.kind = 1; sim->ev.slot = -1;
            }
        } else if (apply_far) { /* far_call.rs:268-1098 */
            const int is_delegated = VAR(ZKC_VAR_FAR_CALL_DELEGATE), is_mimic = VAR(ZKC_VAR_FAR_CALL_MIMIC);
            zkc_vm_context cur_e = *ctx;
            cur_e.pc = pc_plus_one;
            memset(&new_entry, 0, sizeof new_entry);
            new_entry.heap_upper_bound = isa->new_frame_memory_stipend; new_entry.aux_heap_upper_bound = isa->new_frame_memory_stipend;
            const zkc_vm_register *mimic_reg = &s.registers[isa->call_implicit_parameter_reg_idx];
            const uint32_t *dest = b.value; /* src1: the target address */
            const int is_static_call = FLAG(ZKC_VM_FAR_CALL_STATIC_FLAG_IDX), is_call_shard = FLAG(ZKC_VM_FAR_CALL_SHARD_FLAG_IDX);
            /* FarCallPartialABI::from_register_view, :69-98 */
#define ABI_BYTE(k) ((a.value[(k) / 4] >> (8 * ((k) % 4))) & 0xFF)
            const uint32_t abi_ergs_passed = a.value[6], abi_shard = ABI_BYTE(ZKC_VM_ABI_SHARD_ID_BYTE_IDX);
            int abi_constructor = ABI_BYTE(ZKC_VM_ABI_CONSTRUCTOR_CALL_BYTE_IDX) != 0, abi_system = ABI_BYTE(ZKC_VM_ABI_SYSTEM_CALL_BYTE_IDX) != 0;
#undef ABI_BYTE
            const uint32_t caller_shard = cur_e.this_shard_id;
            const uint32_t dest_shard = is_call_shard ? abi_shard : caller_shard;
            const int target_is_zkporter = dest_shard != 0;
            const int target_is_kernel = (dest[0] >> 16) == 0 && dest[1] == 0 && dest[2] == 0 && dest[3] == 0 && dest[4] == 0; /* :396-423 */
            abi_constructor = abi_constructor && is_kernel; abi_system = abi_system && target_is_kernel;
            const uint32_t new_base_page = s.memory_page_counter;
            new_memory_page_counter = s.memory_page_counter + isa->new_memory_pages_per_far_call;
            /* may_be_read_code_hash, :1104-1272 */
            const int zkporter_available = gc->zkporter_is_available != 0;
            const int can_read = !target_is_zkporter || zkporter_available, should_read = can_read, needs_porter_mask = target_is_zkporter && !zkporter_available;
            zkc_log_query q;
            memset(&q, 0, sizeof q);
            q.address[0] = isa->deployer_system_contract_address_low;
            memcpy(q.key, dest, 20);
            q.tx_number_in_block = s.tx_number_in_block; q.timestamp = ts_log;
            q.flags = ZKC_LQ_FLAGS(isa->log_aux_bytes[0], dest_shard, 0, 0, 0);
            if (sim) { /* every address with an odd low word is deployed and runs the world's one program */
                memset(w->value_a, 0, 32);
                if (should_read && (dest[0] & 1)) memcpy(w->value_a, sim->code_hash, 32);
            }
            uint32_t hash[8];
            memcpy(hash, w->value_a, 32);
            memcpy(q.read_value, hash, 32); memcpy(q.written_value, hash, 32);
            const int empty = u256_is_zero(hash);
            const int mask_default_aa = should_read && empty && !target_is_kernel;
            if (mask_default_aa) memcpy(hash, gc->default_aa_code_hash, 32);
            if (needs_porter_mask) memset(hash, 0, 32);
            const int hash_is_trivial = (empty && !mask_default_aa) || needs_porter_mask || !should_read;
            if (should_read) { /* construct_hash_relations_code_hash_read, :1274-1411 */
                uint64_t enc[20], in8[8], cap0[4] = {0, 0, 0, 0};
                orc_log_query_encode(&q, enc);
                sponge_run(&sp, 5, enc, cap0);
                sponge_run(&sp, 6, enc + 8, sp.fin[5] + 8);
                memcpy(in8, enc + 16, 32); memcpy(in8 + 4, ctx->log_queue_forward_tail, 32);
                sponge_run(&sp, 7, in8, sp.fin[6] + 8);
                memcpy(new_fwd_tail, sp.fin[7], 32); new_fwd_len++;
            }
            uint32_t target_code_page = hash_is_trivial ? 0 : s.memory_page_counter;
            /* code hash format, :503-585 */
            const uint32_t top = hash[7], version_byte = top >> 24, marker = (top >> 16) & 0xFF;
            const int normal_marker = marker == 0, constructor_marker = marker == isa->code_hash_yet_constructed_marker;
            const int code_format_exception = version_byte != isa->code_hash_version_byte || !(normal_marker || constructor_marker);
            const int can_call_code = (normal_marker && !abi_constructor) || (constructor_marker && abi_constructor);
            uint32_t masked_hash[8];
            if (can_call_code) { memcpy(masked_hash, hash, 32); masked_hash[7] = (top & 0xFFFF) | (isa->code_hash_at_rest_marker << 16) | (isa->code_hash_version_byte << 24); }
            else if (target_is_kernel) memset(masked_hash, 0, 32);
            else memcpy(masked_hash, gc->default_aa_code_hash, 32);
            const uint32_t code_len_words = code_format_exception ? 0 : (masked_hash[7] & 0xFFFF);
            const int call_now_in_construction_kernel = !can_call_code && target_is_kernel;
            const int exceptions_collapsed = code_format_exception || call_now_in_construction_kernel || (fwd_ptr && !a.is_pointer) ||
                                             generally_invalid || non_addressable;
            vm_fat_ptr p = fp;
            vm_fat_ptr adjusted = {0, p.page, p.start + p.offset, p.length - p.offset};
            vm_fat_ptr for_heaps = {0, use_heap ? heap_page : aux_heap_page, p.start, p.length};
            p = fwd_ptr ? adjusted : for_heaps;
            if (exceptions_collapsed) memset(&p, 0, sizeof p);
            uint32_t ub = exceptions_collapsed ? 0 : upper_bound;
            if (non_addressable && !fwd_ptr) ub = 0xFFFFFFFFu;
            const uint32_t heap_max = use_heap ? ub : 0, aux_max = use_aux ? ub : 0;
            const int heap_uf = heap_max < cur_e.heap_upper_bound, aux_uf = aux_max < cur_e.aux_heap_upper_bound;
            uint32_t growth_cost = use_heap ? (heap_uf ? 0 : heap_max - cur_e.heap_upper_bound) : 0;
            if (use_aux) growth_cost = aux_uf ? 0 : aux_max - cur_e.aux_heap_upper_bound;
            const int growth_uf = ergs_left < growth_cost;
            const uint32_t ergs_after_growth = growth_uf ? 0 : ergs_left - growth_cost;
            if (use_heap) cur_e.heap_upper_bound = heap_uf ? cur_e.heap_upper_bound : heap_max;
            if (use_aux) cur_e.aux_heap_upper_bound = aux_uf ? cur_e.aux_heap_upper_bound : aux_max;
            int exception = exceptions_collapsed || growth_uf; /* callee stipend: FORCED_ERGS_FOR_MSG_VALUE_SIMUALTOR == false */
            int should_decommit = !exception;
            if (!should_decommit) target_code_page = 0;
            /* add_to_decommittment_queue, :1418-1603 */
            const uint32_t decommit_cost = isa->ergs_per_code_word_decommittment * code_len_words;
            const int not_enough_for_decommit = ergs_after_growth < decommit_cost;
            should_decommit = should_decommit && !not_enough_for_decommit;
            uint32_t ergs_after_decommit = should_decommit ? ergs_after_growth - decommit_cost : ergs_after_growth;
            if (sim) {
                w->suggested_page = 0;
                if (should_decommit) {
                    if (!memcmp(masked_hash, sim->code_hash, 32)) {
                        if (!sim->decommitted_page) { sim->decommitted_page = target_code_page; sim_load_code(sim, target_code_page); }
                        w->suggested_page = sim->decommitted_page;
                    } else w->suggested_page = target_code_page; /* unknown code: a fresh, empty page */
                }
            }
            const uint32_t suggested_page = w->suggested_page;
            const int is_first = target_code_page == suggested_page;
            if (should_decommit && !is_first) ergs_after_decommit = ergs_after_growth; /* refund: already decommitted */
            if (should_decommit) {
                uint64_t e8[8];
                e8[0] = (uint64_t)masked_hash[0] + ((uint64_t)(suggested_page & 0xFFFFFF) << 32);
                e8[1] = (uint64_t)masked_hash[1] + ((uint64_t)(suggested_page >> 24) << 32) + ((uint64_t)(ts_log & 0xFFFF) << 40);
                e8[2] = (uint64_t)masked_hash[2] + ((uint64_t)(ts_log >> 16) << 32) + ((uint64_t)is_first << 48);
                for (int i = 3; i < 8; i++) e8[i] = masked_hash[i];
                sponge_run(&sp, 8, e8, s.code_decommittment_queue_state + 8);
                decommit_applies = 1; memcpy(new_decommit_state, sp.fin[8], 96); new_decommit_len = s.code_decommittment_queue_length + 1;
            }
            const uint32_t code_memory_page = should_decommit ? suggested_page : 0; /* UNMAPPED_PAGE */
            exception = exception || not_enough_for_decommit;
            /* the 63 / 64 rule, :870-905 */
            const uint32_t max_passable = (ergs_after_decommit / 64) * 63, leftover = ergs_after_decommit - max_passable;
            const int pass_uf = max_passable < abi_ergs_passed;
            const uint32_t to_pass = pass_uf ? max_passable : abi_ergs_passed;
            cur_e.ergs_remaining = pass_uf ? leftover : leftover + (max_passable - abi_ergs_passed);
            memcpy(new_entry.reverted_queue_tail, w->rollback, 32); memcpy(new_entry.reverted_queue_head, w->rollback, 32);
            new_entry.ergs_remaining = to_pass; new_entry.pc = 0; new_entry.exception_handler_loc = imm0;
            new_entry.is_static_execution = is_static_call || cur_e.is_static_execution;
            new_entry.is_kernel_mode = is_delegated ? cur_e.is_kernel_mode : (uint32_t)target_is_kernel;
            new_entry.code_shard_id = dest_shard; memcpy(new_entry.code_address, dest, 20);
            new_entry.this_shard_id = is_delegated ? caller_shard : dest_shard;
            memcpy(new_entry.this_address, is_delegated ? cur_e.this_address : dest, 20);
            memcpy(new_entry.caller, is_delegated ? cur_e.caller : cur_e.this_address, 20);
            if (is_mimic) memcpy(new_entry.caller, mimic_reg->value, 20);
            new_entry.caller_shard_id = caller_shard;
            new_entry.code_page = code_memory_page; new_entry.base_page = new_base_page;
            memcpy(new_entry.context_u128_value_composite, is_delegated ? cur_e.context_u128_value_composite : s.context_composite_u128, 16);
            new_entry.is_local_call = 0;
            far_call_registers = 1; far_call_system = abi_system;
            far_return_r1 = zero_reg; far_return_r1.is_pointer = 1;
            far_return_r1.value[0] = p.offset; far_return_r1.value[1] = p.page; far_return_r1.value[2] = p.start; far_return_r1.value[3] = p.length;
            far_call_r2.value[0] = (uint32_t)abi_constructor + 2 * (uint32_t)abi_system;
            far_exception = exception;
            reset_context_u128 = 1;
            old_entry = cur_e;
            memcpy(sponge_from, s.stack_sponge_state, 96);
            new_depth = s.context_stack_depth + 1;
            if (sim) {
                if (s.context_stack_depth >= sim->max_depth) sim->overflow = 1;
                else { sim->stack[s.context_stack_depth].context = old_entry; memcpy(sim->stack[s.context_stack_depth].previous_sponge_state, s.stack_sponge_state, 96); }
                sim->ev_kind = 1; sim->ev.kind = 1; sim->ev.slot = -1;
            }
        } else { /* ret.rs:29-479 */
            const int is_ok = VAR(ZKC_VAR_RET_OK), is_revert = VAR(ZKC_VAR_RET_REVERT), is_ret_panic = VAR(ZKC_VAR_RET_PANIC);
            const int is_local = (int)ctx->is_local_call, is_far_return = !is_local;
            zkc_vm_register src0r = a;
            if (is_ret_panic) src0r = zero_reg;
            const int is_to_label = FLAG(ZKC_VM_RET_TO_LABEL_FLAG_IDX);
            zkc_vm_callstack_witness popped;
            memset(&popped, 0, sizeof popped);
            if (sim) {
                if (s.context_stack_depth >= 1 && s.context_stack_depth - 1 < sim->max_depth) popped = sim->stack[s.context_stack_depth - 1];
                if (sim->n_cw < sim->cw_cap) { w->callstack_index = (uint32_t)sim->n_cw; sim->cw_out[sim->n_cw++] = popped; }
                else sim->overflow = 1;
            } else if (w->callstack_index < n_cw) popped = cw[w->callstack_index];
            else checks |= ZKC_VM_CHK_CALLSTACK;
            new_entry = popped.context;
            const zkc_vm_context cur_e = *ctx;
            const int exc1 = fwd_ptr && !src0r.is_pointer && is_far_return;
            const int exc2 = fwd_ptr && fp.page < cur_e.base_page;
            const int exceptions_collapsed = exc1 || exc2 || is_ret_panic;
            vm_fat_ptr p = fp;
            if (exceptions_collapsed) memset(&p, 0, sizeof p);
            vm_fat_ptr adjusted = {0, p.page, p.start + p.offset, p.length - p.offset};
            vm_fat_ptr for_heaps = {0, use_heap ? heap_page : aux_heap_page, p.start, p.length};
            p = fwd_ptr ? adjusted : for_heaps;
            uint32_t ub = exceptions_collapsed ? 0 : upper_bound;
            if (non_addressable && !fwd_ptr) ub = 0xFFFFFFFFu;
            const uint32_t heap_max = use_heap ? ub : 0, aux_max = use_aux ? ub : 0;
            const uint32_t heap_growth = heap_max < cur_e.heap_upper_bound ? 0 : heap_max - cur_e.heap_upper_bound;
            const uint32_t aux_growth = aux_max < cur_e.aux_heap_upper_bound ? 0 : aux_max - cur_e.aux_heap_upper_bound;
            uint32_t growth_cost = (use_heap && is_far_return) ? heap_growth : 0;
            if (use_aux && is_far_return) growth_cost = aux_growth;
            const int uf = ergs_left < growth_cost;
            uint32_t ergs_after = uf ? 0 : ergs_left - growth_cost;
            if (is_local) ergs_after = ergs_left;
            const int non_local_panic = (exceptions_collapsed || uf || is_ret_panic) && is_far_return;
            if (non_local_panic) memset(&p, 0, sizeof p);
            const uint64_t ergs_sum = (uint64_t)ergs_after + new_entry.ergs_remaining;
            if (ergs_sum >> 32) checks |= ZKC_VM_CHK_CALLSTACK; /* add_no_overflow, :310-311 */
            new_entry.ergs_remaining = (uint32_t)ergs_sum;
            if (is_local) { new_entry.heap_upper_bound = cur_e.heap_upper_bound; new_entry.aux_heap_upper_bound = cur_e.aux_heap_upper_bound; }
            const int should_revert = is_revert || is_ret_panic || non_local_panic;
            perform_revert = should_revert;
            if (should_revert && memcmp(cur_e.reverted_queue_head, cur_e.log_queue_forward_tail, 32)) checks |= ZKC_VM_CHK_ROLLBACK_QUEUE; /* :373-383 */
            const int ret_ok = is_ok && !non_local_panic;
            if (ret_ok && memcmp(new_entry.reverted_queue_head, cur_e.reverted_queue_tail, 32)) checks |= ZKC_VM_CHK_ROLLBACK_QUEUE; /* :396-404 */
            if (should_revert) { memcpy(new_fwd_tail, cur_e.reverted_queue_tail, 32); new_fwd_len = cur_e.log_queue_forward_part_length + cur_e.reverted_queue_segment_len; }
            if (ret_ok) {
                memcpy(new_entry.reverted_queue_head, cur_e.reverted_queue_head, 32);
                new_entry.reverted_queue_segment_len = popped.context.reverted_queue_segment_len + cur_e.reverted_queue_segment_len;
            }
            const int use_label = is_to_label && is_local;
            const uint32_t ok_pc = use_label ? imm0 : new_entry.pc, eh_pc = use_label ? imm0 : cur_e.exception_handler_loc;
            new_entry.pc = should_revert ? eh_pc : ok_pc;
            far_return_registers = is_far_return;
            far_return_r1.is_pointer = 1;
            far_return_r1.value[0] = p.offset; far_return_r1.value[1] = p.page; far_return_r1.value[2] = p.start; far_return_r1.value[3] = p.length;
            is_panic_out = is_ret_panic || non_local_panic;
            reset_context_u128 = is_far_return;
            old_entry = popped.context;
            memcpy(sponge_from, popped.previous_sponge_state, 96);
            if (s.context_stack_depth == 0) checks |= ZKC_VM_CHK_CALLSTACK; /* :300 */
            new_depth = s.context_stack_depth - 1;
            if (sim) sim->ev_kind = should_revert ? 3 : 2;
        }
        /* the callstack sponge: 4 absorptions of the saved frame's encoding (call_ret.rs:176-284) */
        uint64_t enc[32], st12[12];
        orc_vm_context_encode(&old_entry, enc);
        memcpy(st12, sponge_from, 96);
        for (int r = 0; r < 4; r++) { sponge_run(&sp, 1 + r, enc + 8 * r, st12 + 8); memcpy(st12, sp.fin[1 + r], 96); }
        if (apply_ret) {
            if (memcmp(st12, s.stack_sponge_state, 96)) checks |= ZKC_VM_CHK_CALLSTACK;
            memcpy(new_stack_sponge, sponge_from, 96);
        } else memcpy(new_stack_sponge, st12, 96);
        replace_callstack = 1;
        new_ctx = new_entry;
        memcpy(new_ctx.log_queue_forward_tail, new_fwd_tail, 32);
        new_ctx.log_queue_forward_part_length = new_fwd_len;
        set_flags = 1; nf[0] = (uint32_t)(is_panic_out && apply_ret); nf[1] = 0; nf[2] = 0;
        new_pending = far_exception; /* pending_exception_if_far_call, call_ret.rs:131-132 */
        if (row) {
            flatten_record_cols(&new_entry, row, stride, ZKC_VM_OP_AUX);
            uint64_t *x = &T(ZKC_VM_OP_AUX);
            x[42 * stride] = (uint64_t)apply_near; x[43 * stride] = (uint64_t)apply_ret; x[44 * stride] = (uint64_t)is_panic_out;
            x[45 * stride] = (uint64_t)perform_revert; x[46 * stride] = (uint64_t)apply_far; x[47 * stride] = (uint64_t)far_exception;
        }
    }
    /* ---------------- state diffs, cycle.rs:158-616 ---------------- */
    /* dst0 / dst1 are dot products of (flag, candidate) pairs (:199-246): zero when no candidate's flag is set */
    if (!(dst0_to_mem_capable || dst0_reg_only)) dst0 = zero_reg;
    if (!write_dst1) dst1 = zero_reg;
    const int perform_mem_write = dst0_mem && dst0_to_mem_capable;
    memq_push(&sp, 2, s.memory_queue_state, &s.memory_queue_length, ts_dst, dst_page, dst_index, 1, &dst0, perform_mem_write);
    if (sim && perform_mem_write) sim_write(sim, dst_page, dst_index, &dst0);
    const int dst0_update_register = dst0_reg_only || (!dst0_mem && dst0_to_mem_capable);
    if (dst0_update_register && dst0_r) s.registers[dst0_r - 1] = dst0;
    if (far_return_registers) { /* specific updates, zeroing and pointer-marker removal ride on the dst0 slot, :352-384 */
        s.registers[0] = far_return_r1;
        for (int r = 1; r < 15; r++) s.registers[r] = zero_reg;
    }
    if (far_call_registers) { /* far_call.rs:1018-1066: r1 = calldata pointer, r2 = call flags, the rest per the calling convention */
        s.registers[0] = far_return_r1;
        s.registers[1] = far_call_r2;
        for (uint32_t r = isa->call_system_abi_registers[0]; r < isa->call_system_abi_registers[1] && r < 15; r++) {
            s.registers[r].is_pointer = 0;
            if (!far_call_system) memset(s.registers[r].value, 0, 32);
        }
        for (uint32_t r = isa->call_reserved_range[0]; r < isa->call_reserved_range[1] && r < 15; r++) s.registers[r] = zero_reg;
        if (isa->call_implicit_parameter_reg_idx < 15) s.registers[isa->call_implicit_parameter_reg_idx] = zero_reg;
    }
    /* dst1 is applied after dst0 and WHATEVER the gadgets flagged: write_as_dst1 is the decoded selector bit itself (cycle.rs:330,
     * :341-347; should_update_dst1 of :177-187 is collected and never read), so an opcode that encodes a dst1 register and has no
     * dst1 result -- a UMA write, a read without the increment flag, a panicking UMA -- leaves (not a pointer, 0) there */
    if (dst1_r) s.registers[dst1_r - 1] = dst1;
    ctx->ergs_remaining = ergs_candidate;
    if (reset_context_u128) memset(s.context_composite_u128, 0, 16);
    if (uma_applies) { memcpy(s.memory_queue_state, uma_memq, 96); s.memory_queue_length = uma_memq_len; }
    if (set_flags) memcpy(s.flags, nf, sizeof nf);
    if (replace_callstack) {
        s.current_context = new_ctx;
        s.context_stack_depth = new_depth;
        memcpy(s.stack_sponge_state, new_stack_sponge, 96);
    }
    s.pending_exception = (uint32_t)new_pending;
    s.memory_page_counter = new_memory_page_counter;
    if (decommit_applies) { memcpy(s.code_decommittment_queue_state, new_decommit_state, 96); s.code_decommittment_queue_length = new_decommit_len; }
    if (row) {
        T(ZKC_VM_DST0) = dst0.is_pointer; T(ZKC_VM_DST1) = dst1.is_pointer;
        for (int i = 0; i < 8; i++) { T(ZKC_VM_DST0 + 1 + i) = dst0.value[i]; T(ZKC_VM_DST1 + 1 + i) = dst1.value[i]; }
        T(ZKC_VM_PERFORM_DST0_MEMORY_WRITE) = (uint64_t)perform_mem_write; T(ZKC_VM_DST0_UPDATE_REGISTER) = (uint64_t)dst0_update_register;
        T(ZKC_VM_DST1_UPDATE_REGISTER) = (uint64_t)write_dst1;
        for (int i = 0; i < 3; i++) T(ZKC_VM_FLAGS_OUT + i) = s.flags[i];
        const zkc_vm_context *c = &s.current_context;
        T(ZKC_VM_PENDING_EXCEPTION_OUT) = s.pending_exception; T(ZKC_VM_PC_OUT) = c->pc; T(ZKC_VM_ERGS_OUT) = c->ergs_remaining;
        T(ZKC_VM_HEAP_BOUND_OUT) = c->heap_upper_bound; T(ZKC_VM_AUX_HEAP_BOUND_OUT) = c->aux_heap_upper_bound;
        T(ZKC_VM_MEMQ_LENGTH_OUT) = s.memory_queue_length; T(ZKC_VM_DEPTH_OUT) = s.context_stack_depth;
        for (int i = 0; i < 4; i++) { T(ZKC_VM_FORWARD_TAIL_OUT + i) = c->log_queue_forward_tail[i]; T(ZKC_VM_ROLLBACK_HEAD_OUT + i) = c->reverted_queue_head[i]; }
        T(ZKC_VM_FORWARD_TAIL_OUT + 4) = c->log_queue_forward_part_length; T(ZKC_VM_ROLLBACK_HEAD_OUT + 4) = c->reverted_queue_segment_len;
        for (int k = 0; k < ZKC_VM_NUM_SPONGES; k++) {
            T(ZKC_VM_SPONGE_ENFORCE + k) = (uint64_t)sp.enf[k];
            for (int i = 0; i < 12; i++) T(ZKC_VM_SPONGE_FINAL + 12 * k + i) = sp.enf[k] ? sp.fin[k][i] : 0;
        }
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

/* Rollback-queue resolution of the out-of-circuit run.  A frame's rollback segment is hash-chained BACKWARDS: every
 * revertable log claims a new head h' with H(rollback item, h') == current head (log.rs:351-371, :583-632), a frame
 * that returns ok hands its segment to its parent (the parent's saved head must be the child's tail, ret.rs:396-404)
 * and a frame that reverts must have its head where the forward queue ends (ret.rs:373-383), its tail becoming the
 * new forward tail.  So the claimed heads / frame tails are only known once a frame's fate is: walk its events
 * (own call marker, logs, merged children) from the most recent one back, starting from the required final head. */
typedef struct orc_vm_lists { orc_vm_entry *entries; int64_t *first, *last; } orc_vm_lists;
static void list_append(orc_vm_lists *L, size_t depth, int64_t e) {
    L->entries[e].prev = L->last[depth];
    L->last[depth] = e;
    if (L->first[depth] < 0) L->first[depth] = e;
}
static void list_merge_into_parent(orc_vm_lists *L, size_t child) {
    if (L->first[child] < 0) return;
    L->entries[L->first[child]].prev = L->last[child - 1];
    if (L->first[child - 1] < 0) L->first[child - 1] = L->first[child];
    L->last[child - 1] = L->last[child];
    L->first[child] = L->last[child] = -1;
}
/* walks the list of `depth` backwards from head value `cur`; resolve: writes the claimed heads / tails into the
 * witness; restore: undoes the storage writes (a reverted frame).  Returns the segment tail in cur. */
static void list_walk(orc_vm_lists *L, size_t depth, uint64_t cur[4], zkc_vm_cycle_witness *witness, int resolve, orc_vm_sim *sim, int restore) {
    for (int64_t e = L->last[depth]; e >= 0; e = L->entries[e].prev) {
        orc_vm_entry *en = &L->entries[e];
        if (resolve) memcpy(witness[e].rollback, cur, 32);
        if (en->kind == 2) {
            uint64_t st[12];
            memcpy(st, en->enc16, 32); memcpy(st + 4, cur, 32); memcpy(st + 8, en->cap, 32);
            orc_poseidon2_permutation(st);
            memcpy(cur, st, 32);
            if (restore && en->slot >= 0) { memcpy(sim->storage[en->slot].value, en->prev_value, 32); sim->storage[en->slot].written = en->prev_written; }
        }
    }
    L->first[depth] = L->last[depth] = -1;
}

/* one pass of the run.  resolve: pass 1 (rollback witness unknown: computed here and patched into `witness`, the
 * states it produces carry unresolved rollback heads and are discarded); otherwise witness[].rollback is input. */
static int vm_run_pass(const zkc_vm_isa *isa, const zkc_vm_closed_form *gc, const zkc_vm_state *initial, const uint32_t *code, size_t code_words, size_t cycles,
                       zkc_vm_state *snapshots, zkc_vm_cycle_witness *witness, zkc_vm_callstack_witness *cw_out, size_t cw_cap,
                       size_t *n_cw, uint64_t root_tail[4], int resolve, zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    orc_vm_sim sim;
    memset(&sim, 0, sizeof sim);
    sim.cells = calloc(ORC_VM_MEM_CELLS, sizeof(orc_vm_cell));
    sim.storage = calloc(ORC_VM_STORAGE_SLOTS, sizeof(orc_vm_slot));
    sim.max_depth = cycles + 2;
    sim.stack = calloc(sim.max_depth, sizeof(zkc_vm_callstack_witness));
    sim.cw_out = cw_out; sim.cw_cap = cw_cap;
    orc_vm_lists L;
    L.entries = calloc(cycles ? cycles : 1, sizeof(orc_vm_entry));
    L.first = malloc(sim.max_depth * sizeof(int64_t)); L.last = malloc(sim.max_depth * sizeof(int64_t));
    for (size_t i = 0; i < sim.max_depth; i++) L.first[i] = L.last[i] = -1;
    sim.code = code; sim.code_words = code_words;
    sim_load_code(&sim, initial->current_context.code_page);
    /* the versioned hash every deployed address answers with: [version byte | marker 0 | length in words] on top */
    for (int i = 0; i < 7; i++) sim.code_hash[i] = 0xC0DE0000u + (uint32_t)i;
    sim.code_hash[7] = (isa->code_hash_version_byte << 24) | (uint32_t)(code_words & 0xFFFF);
    /* the frame below the root: the empty context initial_bootloader_state hashes into the stack sponge */
    zkc_vm_state s0 = *initial;
    memcpy(s0.current_context.reverted_queue_head, root_tail, 32);
    memcpy(s0.current_context.reverted_queue_tail, root_tail, 32);
    {
        zkc_vm_context empty;
        memset(&empty, 0, sizeof empty);
        memcpy(empty.reverted_queue_tail, root_tail, 32); memcpy(empty.reverted_queue_head, root_tail, 32);
        empty.is_kernel_mode = 1;
        sim.stack[0].context = empty;
        uint64_t enc[32], sp12[12] = {0};
        orc_vm_context_encode(&empty, enc);
        for (int r = 0; r < 4; r++) { memcpy(sp12, enc + 8 * r, 64); orc_poseidon2_permutation(sp12); }
        memcpy(s0.stack_sponge_state, sp12, 96);
    }
    zkc_vm_state cur = s0, next;
    if (snapshots) snapshots[0] = cur;
    for (size_t c = 0; c < cycles; c++) {
        uint64_t keep[4];
        memcpy(keep, witness[c].rollback, 32);
        memset(&witness[c], 0, sizeof witness[c]);
        if (!resolve) memcpy(witness[c].rollback, keep, 32);
        const size_t depth = cur.context_stack_depth;
        const uint32_t chk = vm_cycle(isa, gc, &cur, &witness[c], NULL, 0, &sim, &next, NULL, 0);
        /* the joins of the rollback queue cannot hold before the witness is resolved */
        const uint32_t ignore = resolve ? ZKC_VM_CHK_ROLLBACK_QUEUE : 0;
        if (chk & ~ignore) fail(&st, (int64_t)c, chk & ~ignore);
        if (sim.ev_kind == 1) { list_append(&L, depth + 1, (int64_t)c); L.entries[c] = (orc_vm_entry){L.entries[c].prev, 1, -1, {0}, 0, {0}, {0}}; }
        else if (sim.ev_kind == 4) { const int64_t prev = L.last[depth]; L.entries[c] = sim.ev; L.entries[c].prev = prev; L.last[depth] = (int64_t)c; if (L.first[depth] < 0) L.first[depth] = (int64_t)c; }
        else if (sim.ev_kind == 2 && depth >= 2) list_merge_into_parent(&L, depth);
        else if (sim.ev_kind == 3) {
            /* the reverting frame's final head is the forward tail at this point; its tail becomes the forward tail */
            uint64_t h[4];
            memcpy(h, cur.current_context.log_queue_forward_tail, 32);
            list_walk(&L, depth, h, witness, resolve, &sim, 1);
            if (resolve) {
                memcpy(next.current_context.log_queue_forward_tail, h, 32);
                if (depth == 1) memcpy(root_tail, h, 32);
            }
        }
        cur = next;
        if (snapshots) snapshots[c + 1] = cur;
    }
    if (resolve) {
        /* frames still open (and the root after an ok exit): as if they all returned ok, chained from the given tail */
        size_t top = cur.context_stack_depth;
        for (size_t d = top; d >= 2; d--) list_merge_into_parent(&L, d);
        if (L.last[1] >= 0 || top >= 1) {
            uint64_t h[4];
            memcpy(h, root_tail, 32);
            if (L.last[1] >= 0) { list_walk(&L, 1, h, witness, 1, &sim, 0); memcpy(root_tail, h, 32); }
        }
    }
    if (sim.overflow) fail(&st, -1, ZKC_VM_CHK_UNSUPPORTED_OPCODE);
    if (n_cw) *n_cw = sim.n_cw;
    free(sim.cells);
    free(sim.storage); free(sim.stack); free(L.entries); free(L.first); free(L.last);
    if (status) *status = st;
    return st.code;
}

/* out-of-circuit run: fills snapshots[cycles + 1], witness[cycles] and the popped frames; code: [code_words][8].
 * rollback_tail_out: the resolved rollback_queue_tail_for_block (snapshots[0] carries it). */
int orc_main_vm_run(const zkc_vm_isa *isa, const zkc_vm_closed_form *gc, const zkc_vm_state *initial, const uint32_t *code, size_t code_words,
                    size_t cycles, zkc_vm_state *snapshots, zkc_vm_cycle_witness *witness, zkc_vm_callstack_witness *cw_out,
                    size_t cw_cap, size_t *n_cw, uint64_t rollback_tail_out[4], zkc_status *status) {
    uint64_t root_tail[4];
    memcpy(root_tail, initial->current_context.reverted_queue_tail, 32);
    memset(witness, 0, cycles * sizeof *witness);
    zkc_vm_closed_form none;
    if (!gc) { memset(&none, 0, sizeof none); gc = &none; }  /* GlobalContext: zkporter_is_available, default_aa_code_hash */
    vm_run_pass(isa, gc, initial, code, code_words, cycles, NULL, witness, cw_out, cw_cap, n_cw, root_tail, 1, status);
    const int rc = vm_run_pass(isa, gc, initial, code, code_words, cycles, snapshots, witness, cw_out, cw_cap, n_cw, root_tail, 0, status);
    if (rollback_tail_out) memcpy(rollback_tail_out, root_tail, 32);
    return rc;
}

int orc_main_vm_entry_point(zkc_vm_closed_form *io, const zkc_vm_isa *isa, const zkc_vm_state *snapshots,
                            const zkc_vm_cycle_witness *witness, const zkc_vm_callstack_witness *callstack_witness,
                            size_t n_callstack_witness, size_t limit, const zkc_vm_options *options,
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
        const uint32_t chk = vm_cycle(isa, io, &snapshots[c], &w, callstack_witness, n_callstack_witness, NULL, &next, trace ? trace + c : NULL, limit);
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
