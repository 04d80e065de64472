// Main VM circuit on sm_100a: main_vm_entry_point (/root/reference/src/main_vm/mod.rs:47-232) and vm_cycle
// (/root/reference/src/main_vm/cycle.rs:28-795, pre_state.rs:71-519, decoded_opcode.rs:42-527, utils.rs, opcodes/*).
//
// The reference runs `limit` cycles sequentially, each a function of the previous VmLocalState and of the witness
// oracle's answers.  With the per-cycle VmLocalState supplied by the host (the out-of-circuit VM run already has
// them) every cycle is independent: ONE THREAD PER CYCLE evaluates the cycle from its snapshot, verifies that the
// result is the next snapshot (the same "hint + verify every link" pattern as the queue heads of the sorters), runs
// the cycle's memory-queue sponges (code fetch, src0 read, dst0 write: up to 3 Poseidon2) and writes the trace row.
// zkc_main_vm_simulate is the out-of-circuit run itself (one thread per independent VM instance, same cycle
// function, memory reads answered by a per-instance memory model) -- the role of the external zk_evm crate.
//
// Built opcode subset: nop, add, sub, jump, binop, mul, div, shifts, ptr, context and every src0 / dst0 addressing
// mode.  log / near_call / far_call / ret / uma (and therefore executed exceptions, which the circuit masks into
// ret.panic) report ZKC_ERR_UNSUPPORTED.
#include "ctx.cuh"
#include "poseidon2.cuh"

namespace zkc {

struct VmDev {
    zkc_vm_closed_form io;
    zkc_vm_options opt;
    uint64_t limit;
    uint32_t start, pad0;
    zkc_vm_state s0;
    zkc_vm_state s_final;  // state after the last cycle, as computed (not the host's snapshot)
    unsigned long long first_bad;
    uint32_t failed_checks, pad1;
    uint64_t commitment[4];
    zkc_status status;
};

// ---- 256-bit helpers on little-endian u32 limbs -------------------------------------------------------------
struct U256 {
    uint32_t v[8];
};
__device__ __forceinline__ bool u256_is_zero(const U256 &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
__device__ __forceinline__ uint32_t u256_add(const U256 &a, const U256 &b, U256 &c) {
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { const uint64_t t = (uint64_t)a.v[i] + b.v[i] + carry; c.v[i] = (uint32_t)t; carry = t >> 32; }
    return (uint32_t)carry;
}
__device__ __forceinline__ uint32_t u256_sub(const U256 &a, const U256 &b, U256 &c) {
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { const uint64_t t = (uint64_t)a.v[i] - b.v[i] - borrow; c.v[i] = (uint32_t)t; borrow = (t >> 32) & 1; }
    return (uint32_t)borrow;
}
__device__ void u256_mul(const U256 &a, const U256 &b, U256 &lo, U256 &hi) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint64_t t = (uint64_t)a.v[i] * b.v[j] + r[i + j] + carry;
            r[i + j] = (uint32_t)t; carry = t >> 32;
        }
        r[i + 8] = (uint32_t)carry;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { lo.v[i] = r[i]; hi.v[i] = r[i + 8]; }
}
__device__ __forceinline__ bool u256_ge(const U256 &a, const U256 &b) {
    U256 t;
    return u256_sub(a, b, t) == 0;
}
// q = a / b, r = a % b for b != 0: Knuth's algorithm D on 32-bit limbs (at most 8 quotient digits, each one 64/32
// division + a multiply-subtract), instead of 256 shift-subtract steps that every lane of a warp would wait for
__device__ void u256_divrem(const U256 &a, const U256 &b, U256 &q, U256 &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) { q.v[i] = 0; r.v[i] = 0; }
    int n = 8;
    while (n > 1 && b.v[n - 1] == 0) n--;
    if (n == 1) {
        uint64_t rem = 0;
        const uint32_t d = b.v[0];
        for (int i = 7; i >= 0; i--) {
            const uint64_t cur = (rem << 32) | a.v[i];
            q.v[i] = (uint32_t)(cur / d);
            rem = cur % d;
        }
        r.v[0] = (uint32_t)rem;
        return;
    }
    const int sh = __clz(b.v[n - 1]);
    uint32_t v[8], u[9];
    for (int i = n - 1; i > 0; i--) v[i] = sh ? (b.v[i] << sh) | (b.v[i - 1] >> (32 - sh)) : b.v[i];
    v[0] = b.v[0] << sh;
    u[8] = sh ? a.v[7] >> (32 - sh) : 0;
    for (int i = 7; i > 0; i--) u[i] = sh ? (a.v[i] << sh) | (a.v[i - 1] >> (32 - sh)) : a.v[i];
    u[0] = a.v[0] << sh;
    for (int j = 8 - n; j >= 0; j--) {
        const uint64_t num = ((uint64_t)u[j + n] << 32) | u[j + n - 1];
        uint64_t qhat = num / v[n - 1], rhat = num % v[n - 1];
        while (qhat >= (1ull << 32) || qhat * v[n - 2] > ((rhat << 32) | u[j + n - 2])) {
            qhat--;
            rhat += v[n - 1];
            if (rhat >= (1ull << 32)) break;
        }
        // u[j .. j+n] -= qhat * v
        int64_t borrow = 0;
        uint64_t carry = 0;
        for (int i = 0; i < n; i++) {
            const uint64_t p = qhat * v[i] + carry;
            carry = p >> 32;
            const int64_t t = (int64_t)u[i + j] - (int64_t)(uint32_t)p + borrow;
            u[i + j] = (uint32_t)t;
            borrow = t >> 32;  // 0 or -1
        }
        const int64_t t = (int64_t)u[j + n] - (int64_t)carry + borrow;
        u[j + n] = (uint32_t)t;
        if (t < 0) {  // qhat was one too large: add the divisor back
            qhat--;
            uint64_t c = 0;
            for (int i = 0; i < n; i++) {
                const uint64_t x = (uint64_t)u[i + j] + v[i] + c;
                u[i + j] = (uint32_t)x;
                c = x >> 32;
            }
            u[j + n] += (uint32_t)c;
        }
        q.v[j] = (uint32_t)qhat;
    }
    for (int i = 0; i < n; i++) r.v[i] = sh ? (u[i] >> sh) | ((uint64_t)u[i + 1] << (32 - sh)) : u[i];
}
// (a << s) mod 2^256 and a >> (256 - s) for s in [0, 255]: the two halves of a * 2^s (shifts.rs:95-96)
__device__ void u256_shl_wide(const U256 &a, uint32_t s, U256 &lo, U256 &hi) {
    const uint32_t limbs = s >> 5, bits = s & 31;
    uint32_t w[17];
#pragma unroll
    for (int i = 0; i < 17; i++) w[i] = 0;
    for (int i = 0; i < 8; i++) {  // dynamic limb offset: small loop in local memory
        const uint64_t t = (uint64_t)a.v[i] << bits;
        w[i + limbs] |= (uint32_t)t;
        w[i + limbs + 1] |= (uint32_t)(t >> 32);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { lo.v[i] = w[i]; hi.v[i] = w[i + 8]; }
}
__device__ void u256_shr(const U256 &a, uint32_t s, U256 &q) {
    const uint32_t limbs = s >> 5, bits = s & 31;
    for (int i = 0; i < 8; i++) {
        const uint32_t lo = i + limbs < 8 ? a.v[i + limbs] : 0, hi = i + limbs + 1 < 8 ? a.v[i + limbs + 1] : 0;
        q.v[i] = bits ? (lo >> bits) | (hi << (32 - bits)) : lo;
    }
}

// ---- encodings / queue -------------------------------------------------------------------------------------------
__device__ __forceinline__ void vm_mq_encode(uint32_t ts, uint32_t page, uint32_t index, uint32_t rw, const zkc_vm_register &r, uint64_t (&e)[8]) {
    const uint32_t *v = r.value;  // MemoryQuery::encode, base_structures/memory_query/mod.rs:103-221
    e[0] = ts; e[1] = page;
    e[2] = (uint64_t)index | ((uint64_t)rw << 32) | ((uint64_t)(r.is_pointer & 1) << 33);
    e[3] = (uint64_t)v[0] | ((uint64_t)(v[5] & 0xFFFFFFu) << 32);
    e[4] = (uint64_t)v[1] | ((uint64_t)(v[5] >> 24) << 32) | ((uint64_t)(v[6] & 0xFFFFu) << 40);
    e[5] = (uint64_t)v[2] | ((uint64_t)(v[6] >> 16) << 32) | ((uint64_t)(v[7] & 0xFFu) << 48);
    e[6] = (uint64_t)v[3] | ((uint64_t)(v[7] >> 8) << 32);
    e[7] = v[4];
}
// tail' = P(enc || tail[8..12]) when `execute` (main_vm/utils.rs:194-230, :442-515, cycle.rs:845-905)
__device__ __forceinline__ void vm_memq_push(uint64_t (&q)[12], uint32_t &len, uint32_t ts, uint32_t page, uint32_t index, uint32_t rw,
                                             const zkc_vm_register &val, bool execute) {
    if (!execute) return;
    uint64_t e[8];
    vm_mq_encode(ts, page, index, rw, val, e);
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = e[i];
    poseidon2_permute(q);
    len++;
}

// saved_context.rs:111-270
__device__ void vm_context_encode(const zkc_vm_context &c, uint64_t (&e)[32]) {
    for (int i = 0; i < 4; i++) { e[i] = c.reverted_queue_head[i]; e[4 + i] = c.reverted_queue_tail[i]; }
    for (int i = 0; i < 5; i++) { e[8 + i] = c.code_address[i]; e[13 + i] = c.this_address[i]; e[18 + i] = c.caller[i]; }
    for (int i = 0; i < 4; i++) e[23 + i] = c.context_u128_value_composite[i];
    e[27] = (uint64_t)c.code_page + ((uint64_t)c.pc << 32) + ((uint64_t)c.this_shard_id << 48) + ((uint64_t)c.is_static_execution << 56);
    e[28] = (uint64_t)c.base_page + ((uint64_t)c.sp << 32) + ((uint64_t)c.caller_shard_id << 48) + ((uint64_t)c.is_kernel_mode << 56);
    e[29] = (uint64_t)c.ergs_remaining + ((uint64_t)c.exception_handler_loc << 32) + ((uint64_t)c.code_shard_id << 48) + ((uint64_t)c.is_local_call << 56);
    const uint32_t sl = c.reverted_queue_segment_len;
    e[30] = (uint64_t)c.heap_upper_bound + ((uint64_t)(sl & 0xFF) << 32) + ((uint64_t)((sl >> 8) & 0xFF) << 40);
    e[31] = (uint64_t)c.aux_heap_upper_bound + ((uint64_t)((sl >> 16) & 0xFF) << 32) + ((uint64_t)(sl >> 24) << 40);
}

// CSVarLengthEncodable order of VmLocalState (vm_state/mod.rs:92-109): 243 elements
__device__ int vm_flatten_state(const zkc_vm_state &s, uint64_t *dst) {
    int n = 0;
    for (int i = 0; i < 8; i++) dst[n++] = s.previous_code_word[i];
    for (int r = 0; r < 15; r++) { dst[n++] = s.registers[r].is_pointer; for (int i = 0; i < 8; i++) dst[n++] = s.registers[r].value[i]; }
    for (int i = 0; i < 3; i++) dst[n++] = s.flags[i];
    dst[n++] = s.timestamp; dst[n++] = s.memory_page_counter; dst[n++] = s.tx_number_in_block; dst[n++] = s.previous_code_page;
    dst[n++] = s.previous_super_pc; dst[n++] = s.pending_exception; dst[n++] = s.ergs_per_pubdata_byte;
    const zkc_vm_context &c = s.current_context;
    for (int i = 0; i < 5; i++) dst[n++] = c.this_address[i];
    for (int i = 0; i < 5; i++) dst[n++] = c.caller[i];
    for (int i = 0; i < 5; i++) dst[n++] = c.code_address[i];
    dst[n++] = c.code_page; dst[n++] = c.base_page; dst[n++] = c.heap_upper_bound; dst[n++] = c.aux_heap_upper_bound;
    for (int i = 0; i < 4; i++) dst[n++] = c.reverted_queue_head[i];
    for (int i = 0; i < 4; i++) dst[n++] = c.reverted_queue_tail[i];
    dst[n++] = c.reverted_queue_segment_len;
    dst[n++] = c.pc; dst[n++] = c.sp; dst[n++] = c.exception_handler_loc; dst[n++] = c.ergs_remaining;
    dst[n++] = c.is_static_execution; dst[n++] = c.is_kernel_mode;
    dst[n++] = c.this_shard_id; dst[n++] = c.caller_shard_id; dst[n++] = c.code_shard_id;
    for (int i = 0; i < 4; i++) dst[n++] = c.context_u128_value_composite[i];
    dst[n++] = c.is_local_call;
    for (int i = 0; i < 4; i++) dst[n++] = c.log_queue_forward_tail[i];
    dst[n++] = c.log_queue_forward_part_length;
    dst[n++] = s.context_stack_depth;
    for (int i = 0; i < 12; i++) dst[n++] = s.stack_sponge_state[i];
    for (int i = 0; i < 12; i++) dst[n++] = s.memory_queue_state[i];
    dst[n++] = s.memory_queue_length;
    for (int i = 0; i < 12; i++) dst[n++] = s.code_decommittment_queue_state[i];
    dst[n++] = s.code_decommittment_queue_length;
    for (int i = 0; i < 4; i++) dst[n++] = s.context_composite_u128[i];
    return n;
}

// field-wise equality (padding words are not state)
__device__ bool vm_state_equal(const zkc_vm_state &a, const zkc_vm_state &b) {
    bool eq = true;
    for (int i = 0; i < 8; i++) eq &= a.previous_code_word[i] == b.previous_code_word[i];
    for (int r = 0; r < 15; r++) {
        eq &= a.registers[r].is_pointer == b.registers[r].is_pointer;
        for (int i = 0; i < 8; i++) eq &= a.registers[r].value[i] == b.registers[r].value[i];
    }
    for (int i = 0; i < 3; i++) eq &= a.flags[i] == b.flags[i];
    eq &= a.timestamp == b.timestamp && a.memory_page_counter == b.memory_page_counter && a.tx_number_in_block == b.tx_number_in_block &&
          a.previous_code_page == b.previous_code_page && a.previous_super_pc == b.previous_super_pc &&
          a.pending_exception == b.pending_exception && a.ergs_per_pubdata_byte == b.ergs_per_pubdata_byte &&
          a.context_stack_depth == b.context_stack_depth && a.memory_queue_length == b.memory_queue_length &&
          a.code_decommittment_queue_length == b.code_decommittment_queue_length;
    for (int i = 0; i < 4; i++) eq &= a.context_composite_u128[i] == b.context_composite_u128[i];
    const zkc_vm_context &c = a.current_context, &e = b.current_context;
    for (int i = 0; i < 5; i++) eq &= c.this_address[i] == e.this_address[i] && c.caller[i] == e.caller[i] && c.code_address[i] == e.code_address[i];
    eq &= c.code_page == e.code_page && c.base_page == e.base_page && c.heap_upper_bound == e.heap_upper_bound &&
          c.aux_heap_upper_bound == e.aux_heap_upper_bound && c.reverted_queue_segment_len == e.reverted_queue_segment_len &&
          c.pc == e.pc && c.sp == e.sp && c.exception_handler_loc == e.exception_handler_loc && c.ergs_remaining == e.ergs_remaining &&
          c.is_static_execution == e.is_static_execution && c.is_kernel_mode == e.is_kernel_mode && c.this_shard_id == e.this_shard_id &&
          c.caller_shard_id == e.caller_shard_id && c.code_shard_id == e.code_shard_id && c.is_local_call == e.is_local_call &&
          c.log_queue_forward_part_length == e.log_queue_forward_part_length;
    for (int i = 0; i < 4; i++)
        eq &= c.reverted_queue_head[i] == e.reverted_queue_head[i] && c.reverted_queue_tail[i] == e.reverted_queue_tail[i] &&
              c.context_u128_value_composite[i] == e.context_u128_value_composite[i] && c.log_queue_forward_tail[i] == e.log_queue_forward_tail[i];
    for (int i = 0; i < 12; i++)
        eq &= a.stack_sponge_state[i] == b.stack_sponge_state[i] && a.memory_queue_state[i] == b.memory_queue_state[i] &&
              a.code_decommittment_queue_state[i] == b.code_decommittment_queue_state[i];
    return eq;
}

// loading.rs:13-226
__device__ void vm_initial_bootloader_state(const zkc_vm_closed_form &io, const zkc_vm_isa &isa, zkc_vm_state &st) {
    memset(&st, 0, sizeof st);
    zkc_vm_context &ctx = st.current_context;
    ctx.base_page = isa.bootloader_base_page;
    ctx.code_page = isa.bootloader_code_page;
    ctx.exception_handler_loc = isa.initial_frame_formal_eh_location;
    ctx.ergs_remaining = isa.vm_initial_frame_ergs;
    ctx.code_address[0] = isa.bootloader_formal_address_low;
    ctx.this_address[0] = isa.bootloader_formal_address_low;
    for (int i = 0; i < 4; i++) { ctx.reverted_queue_tail[i] = io.rollback_queue_tail_for_block[i]; ctx.reverted_queue_head[i] = io.rollback_queue_tail_for_block[i]; }
    ctx.is_kernel_mode = 1;
    ctx.heap_upper_bound = isa.bootloader_max_memory;
    ctx.aux_heap_upper_bound = isa.bootloader_max_memory;
    zkc_vm_context empty;
    memset(&empty, 0, sizeof empty);
    for (int i = 0; i < 4; i++) { empty.reverted_queue_tail[i] = io.rollback_queue_tail_for_block[i]; empty.reverted_queue_head[i] = io.rollback_queue_tail_for_block[i]; }
    empty.is_kernel_mode = 1;
    uint64_t enc[32], s[12];
    vm_context_encode(empty, enc);
    for (int i = 0; i < 12; i++) s[i] = 0;
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 8; i++) s[i] = enc[8 * r + i];
        poseidon2_permute(s);
    }
    for (int i = 0; i < 12; i++) st.stack_sponge_state[i] = s[i];
    st.context_stack_depth = 1;
    st.memory_queue_length = io.memory_queue_initial_length;
    st.code_decommittment_queue_length = io.decommitment_queue_initial_length;
    for (int i = 0; i < 12; i++) { st.memory_queue_state[i] = io.memory_queue_initial_tail[i]; st.code_decommittment_queue_state[i] = io.decommitment_queue_initial_tail[i]; }
    st.timestamp = isa.starting_timestamp;
    st.memory_page_counter = isa.starting_base_page;
    st.registers[0].is_pointer = 1;
    st.registers[0].value[1] = isa.bootloader_calldata_page;
}

// memory model of the out-of-circuit run: one code page and one stack page of 2^16 words
struct VmMemory {
    zkc_vm_register *code, *stack;
    uint32_t code_page, stack_page;
};

__device__ __forceinline__ bool prop(uint64_t props, int bit) { return (props >> bit) & 1; }
__device__ __forceinline__ zkc_vm_register reg_zero() {
    zkc_vm_register r;
    r.is_pointer = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r.value[i] = 0;
    return r;
}
__device__ __forceinline__ U256 as_u256(const zkc_vm_register &r) {
    U256 x;
#pragma unroll
    for (int i = 0; i < 8; i++) x.v[i] = r.value[i];
    return x;
}

// what one cycle changes in a VmLocalState; every other word must carry over unchanged
struct VmDelta {
    uint32_t pending, pc, sp, ergs, prev_code_page, prev_super_pc, timestamp, memq_len;
    uint32_t flags[3];
    uint32_t cw[8];       // previous_code_word after the opcode fetch
    uint32_t idx0, idx1;  // 1-based register written by dst0 / dst1, 0 = none (dst1 is applied after dst0)
    zkc_vm_register val0, val1;
    uint32_t set_u128, u128[4], set_pubdata, pubdata, inc_tx;
    uint32_t push_mask;   // bit k: memory queue push k happens (0 opcode fetch, 1 src0 read, 2 dst0 write)
};

__device__ void vm_apply_delta(zkc_vm_state &t, const VmDelta &d) {
    t.pending_exception = d.pending;
    t.current_context.pc = d.pc; t.current_context.sp = d.sp; t.current_context.ergs_remaining = d.ergs;
    t.previous_code_page = d.prev_code_page; t.previous_super_pc = d.prev_super_pc; t.timestamp = d.timestamp;
    t.memory_queue_length = d.memq_len;
    for (int i = 0; i < 3; i++) t.flags[i] = d.flags[i];
    for (int i = 0; i < 8; i++) t.previous_code_word[i] = d.cw[i];
    if (d.idx0) t.registers[d.idx0 - 1] = d.val0;
    if (d.idx1) t.registers[d.idx1 - 1] = d.val1;
    if (d.set_u128) for (int i = 0; i < 4; i++) t.context_composite_u128[i] = d.u128[i];
    if (d.set_pubdata) t.ergs_per_pubdata_byte = d.pubdata;
    if (d.inc_tx) t.tx_number_in_block += 1;
}

// memory queue push k of a cycle: tail' = P(enc || tail[8..12]) (main_vm/utils.rs:194-230, :442-515, cycle.rs:845-905).
// SIM hashes at once into the running state `q`; the batched circuit only records the encoding -- the sponges of
// all cycles run afterwards as dense launches (vm_memq_kernel), so that a warp never waits for a lane that hashes
template <bool SIM>
__device__ __forceinline__ void vm_push(int k, bool execute, VmDelta &d, uint64_t *q, uint64_t *penc, uint32_t ts, uint32_t page,
                                        uint32_t index, uint32_t rw, const zkc_vm_register &val) {
    if (!execute) return;
    uint64_t e[8];
    vm_mq_encode(ts, page, index, rw, val, e);
    d.push_mask |= 1u << k;
    d.memq_len++;
    if (SIM) {
        uint64_t t[12];
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = e[i];
#pragma unroll
        for (int i = 8; i < 12; i++) t[i] = q[i];
        poseidon2_permute(t);
#pragma unroll
        for (int i = 0; i < 12; i++) q[i] = t[i];
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) penc[8 * k + i] = e[i];
    }
}

// One vm_cycle from the state `s` (read only; only the words a cycle needs are touched): returns the check bits and
// what changes in `d`.  SIM: memory reads are answered by `mem` and recorded into `w`; otherwise they come from `w`.
// trace / limit / row: where to put the row (trace may be null; the MEMQ_AFTER_* columns are written by whoever runs
// the sponges).
template <bool SIM, typename W>
__device__ uint32_t vm_cycle_dev(const zkc_vm_isa *__restrict__ isa, const zkc_vm_state &s, VmDelta &d, W &w, VmMemory *mem,
                                 uint64_t *q, uint64_t *penc, uint64_t *__restrict__ trace, size_t limit, size_t row) {
#define TR(col) trace[(size_t)(col) * limit + row]
    const bool wr = trace != nullptr;
    uint32_t checks = 0;
    const zkc_vm_context &ctx = s.current_context;
    d.push_mask = 0; d.set_u128 = 0; d.set_pubdata = 0; d.inc_tx = 0; d.idx0 = 0; d.idx1 = 0;
    d.memq_len = s.memory_queue_length;
    // ---- create_prestate ---------------------------------------------------------------------------------------
    const bool should_skip = s.context_stack_depth == 0;
    const bool pending = s.pending_exception != 0;
    const bool should_try_read = !should_skip && !pending;
    const uint32_t pc = ctx.pc, super_pc = pc >> 2, sub_pc = pc & 3;
    const uint32_t code_page = ctx.code_page;
    const bool should_read_opcode = should_try_read && !(s.previous_code_page == code_page && super_pc == s.previous_super_pc);
    const uint32_t ts0 = s.timestamp;
    zkc_vm_register code_val = reg_zero();
    if (should_read_opcode) {
        if constexpr (SIM) {
            code_val = mem->code[super_pc]; code_val.is_pointer = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) w.code_word[i] = code_val.value[i];
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) code_val.value[i] = w.code_word[i];
        }
    } else if constexpr (SIM) {
#pragma unroll
        for (int i = 0; i < 8; i++) w.code_word[i] = 0;
    }
    vm_push<SIM>(0, should_read_opcode, d, q, penc, ts0, code_page, super_pc, 0, code_val);
#pragma unroll
    for (int i = 0; i < 8; i++) d.cw[i] = should_read_opcode ? code_val.value[i] : s.previous_code_word[i];
    uint32_t op_lo = 0, op_hi = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) if ((int)sub_pc == i) { op_lo = d.cw[6 - 2 * i]; op_hi = d.cw[7 - 2 * i]; }
    if (should_skip) { op_lo = (uint32_t)isa->nop_opcode_encoding; op_hi = (uint32_t)(isa->nop_opcode_encoding >> 32); }
    if (pending) { op_lo = (uint32_t)isa->panic_opcode_encoding; op_hi = (uint32_t)(isa->panic_opcode_encoding >> 32); }
    if (wr) {
        TR(ZKC_VM_SHOULD_SKIP_CYCLE) = should_skip; TR(ZKC_VM_PENDING_EXCEPTION_IN) = pending; TR(ZKC_VM_SHOULD_READ_OPCODE) = should_read_opcode;
        TR(ZKC_VM_SUPER_PC) = super_pc; TR(ZKC_VM_SUB_PC) = sub_pc;
#pragma unroll
        for (int i = 0; i < 8; i++) TR(ZKC_VM_CODE_WORD + i) = d.cw[i];
        TR(ZKC_VM_OPCODE) = op_lo; TR(ZKC_VM_OPCODE + 1) = op_hi;
    }
    d.prev_code_page = code_page;
    d.pc = should_skip ? pc : ((pc + 1) & 0xFFFF);
    d.prev_super_pc = should_skip ? s.previous_super_pc : super_pc;
    d.timestamp = should_skip ? ts0 : ts0 + 4;
    const bool is_kernel = ctx.is_kernel_mode != 0, is_static = ctx.is_static_execution != 0;
    const bool callstack_full = s.context_stack_depth == isa->vm_max_stack_depth;
    // ---- perform_initial_decoding ------------------------------------------------------------------------------
    const uint32_t variant = op_lo & 0x7FF, cond_idx = (op_lo >> 13) & 7;
    uint32_t src_regs = (op_lo >> 16) & 0xFF, dst_regs = op_lo >> 24;
    const uint32_t imm0 = op_hi & 0xFFFF, imm1 = op_hi >> 16;
    const uint64_t props_full = isa->opcode_props[variant];
    constexpr uint64_t MASK48 = (1ull << ZKC_VM_DESCRIPTION_BITS_FLATTENED) - 1;
    uint64_t props = props_full & MASK48;
    const uint32_t aux = (uint32_t)(props_full >> ZKC_VM_DESCRIPTION_BITS_FLATTENED);
    const uint32_t f0 = s.flags[0], f1 = s.flags[1], f2 = s.flags[2];
    const uint32_t encoded_flags = (f0 & 1) | ((f1 & 1) << 1) | ((f2 & 1) << 2);
    const bool condition = isa->condition_table[cond_idx][encoded_flags] != 0;
    const uint32_t cost = should_skip ? 0 : isa->opcode_price[variant];
    const uint32_t ergs_in = ctx.ergs_remaining;
    const bool out_of_ergs = ergs_in < cost;
    const uint32_t ergs_left = out_of_ergs ? 0 : ergs_in - cost;
    const bool explicit_panic = (aux >> ZKC_VM_AUX_EXPLICIT_PANIC) & 1;
    const bool kernel_exc = ((aux >> ZKC_VM_AUX_KERNEL_MODE) & 1) && !is_kernel;
    const bool static_exc = is_static && !((aux >> ZKC_VM_AUX_CAN_BE_USED_IN_STATIC) & 1);
    const bool mask_into_panic = explicit_panic || out_of_ergs || kernel_exc || static_exc || callstack_full;
    if (mask_into_panic) props = isa->panic_bitspread & MASK48;
    const bool mask_into_nop = !mask_into_panic && !condition;
    if (mask_into_nop) props = isa->nop_bitspread & MASK48;
    if (mask_into_nop || mask_into_panic) { src_regs = 0; dst_regs = 0; }
    const uint32_t src0_r = src_regs & 15, src1_r = src_regs >> 4, dst0_r = dst_regs & 15, dst1_r = dst_regs >> 4;
    d.ergs = ergs_left;
#define TYPE(t) prop(props, ZKC_VM_BIT_TYPE(t))
#define VAR(v) prop(props, ZKC_VM_BIT_VARIANT(v))
#define FLAG(f) prop(props, ZKC_VM_BIT_FLAG(f))
#define SRCM(m) prop(props, ZKC_VM_BIT_SRC_MODE(m))
#define DSTM(m) prop(props, ZKC_VM_BIT_DST_MODE(m))
    if (TYPE(ZKC_OP_INVALID)) checks |= ZKC_VM_CHK_INVALID_OPCODE;
    if (TYPE(ZKC_OP_NEAR_CALL) || TYPE(ZKC_OP_LOG) || TYPE(ZKC_OP_FAR_CALL) || TYPE(ZKC_OP_RET) || TYPE(ZKC_OP_UMA)) checks |= ZKC_VM_CHK_UNSUPPORTED_OPCODE;
    if (wr) {
        TR(ZKC_VM_VARIANT) = variant; TR(ZKC_VM_CONDITION_IDX) = cond_idx; TR(ZKC_VM_CONDITION) = condition; TR(ZKC_VM_ERGS_COST) = cost;
        TR(ZKC_VM_OUT_OF_ERGS) = out_of_ergs; TR(ZKC_VM_KERNEL_MODE_EXCEPTION) = kernel_exc; TR(ZKC_VM_STATIC_EXCEPTION) = static_exc;
        TR(ZKC_VM_CALLSTACK_IS_FULL) = callstack_full; TR(ZKC_VM_EXPLICIT_PANIC) = explicit_panic; TR(ZKC_VM_MASK_INTO_PANIC) = mask_into_panic;
        TR(ZKC_VM_MASK_INTO_NOP) = mask_into_nop; TR(ZKC_VM_PROPS) = props; TR(ZKC_VM_DIRTY_ERGS_LEFT) = ergs_left;
        TR(ZKC_VM_SRC0_REG) = src0_r; TR(ZKC_VM_SRC1_REG) = src1_r; TR(ZKC_VM_DST0_REG) = dst0_r; TR(ZKC_VM_DST1_REG) = dst1_r;
        TR(ZKC_VM_IMM0) = imm0; TR(ZKC_VM_IMM1) = imm1;
    }
    // ---- operands -------------------------------------------------------------------------------------------------
    const zkc_vm_register draft_src0 = src0_r ? s.registers[src0_r - 1] : reg_zero();
    const zkc_vm_register src1_register = src1_r ? s.registers[src1_r - 1] : reg_zero();
    const uint32_t src0_low = draft_src0.value[0] & 0xFFFF;
    const uint32_t dst0_low = (dst0_r ? s.registers[dst0_r - 1].value[0] : 0u) & 0xFFFF;
    const uint32_t current_sp = ctx.sp, stack_page = ctx.base_page + 1;
    const bool is_nop = TYPE(ZKC_OP_NOP);
    uint32_t src_page, src_index, sp_after_src0;
    bool should_read_src0;
    {
        const bool use_code = SRCM(ZKC_MODE_CODE_PAGE), abs_ = SRCM(ZKC_MODE_STACK_ABSOLUTE), rel = SRCM(ZKC_MODE_STACK_OFFSET), pp = SRCM(ZKC_MODE_STACK_PUSH_POP);
        const uint32_t idx_abs = (src0_low + imm0) & 0xFFFF, idx_rel = (current_sp - idx_abs) & 0xFFFF;
        const bool use_stack = abs_ || rel || pp;
        should_read_src0 = (use_stack || use_code) && !is_nop;
        src_page = use_stack ? stack_page : code_page;
        src_index = (use_code || abs_) ? idx_abs : idx_rel;
        sp_after_src0 = pp ? idx_rel : current_sp;
    }
    uint32_t dst_index, new_sp;
    bool dst0_mem;
    {
        const bool abs_ = DSTM(ZKC_MODE_STACK_ABSOLUTE), rel = DSTM(ZKC_MODE_STACK_OFFSET), pp = DSTM(ZKC_MODE_STACK_PUSH_POP);
        const uint32_t idx_abs = (dst0_low + imm1) & 0xFFFF;
        dst0_mem = (abs_ || rel || pp) && !is_nop;
        dst_index = abs_ ? idx_abs : (pp ? sp_after_src0 : ((sp_after_src0 - idx_abs) & 0xFFFF));
        new_sp = pp ? ((sp_after_src0 + idx_abs) & 0xFFFF) : sp_after_src0;
    }
    d.sp = new_sp;
    zkc_vm_register src0_mem = reg_zero();
    if (should_read_src0) {
        if constexpr (SIM) {
            if (src_page == mem->code_page) { src0_mem = mem->code[src_index]; src0_mem.is_pointer = 0; }
            else if (src_page == mem->stack_page) src0_mem = mem->stack[src_index];
            w.src0_is_pointer = src0_mem.is_pointer;
#pragma unroll
            for (int i = 0; i < 8; i++) w.src0_value[i] = src0_mem.value[i];
        } else {
            src0_mem.is_pointer = w.src0_is_pointer & 1;
#pragma unroll
            for (int i = 0; i < 8; i++) src0_mem.value[i] = w.src0_value[i];
        }
    } else if constexpr (SIM) {
        w.src0_is_pointer = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) w.src0_value[i] = 0;
    }
    vm_push<SIM>(1, should_read_src0, d, q, penc, ts0, src_page, src_index, 0, src0_mem);
    if (wr) {
        TR(ZKC_VM_SRC0_PAGE) = src_page; TR(ZKC_VM_SRC0_INDEX) = src_index; TR(ZKC_VM_SHOULD_READ_SRC0) = should_read_src0;
        TR(ZKC_VM_SP_AFTER_SRC0) = sp_after_src0; TR(ZKC_VM_DST0_PAGE) = stack_page; TR(ZKC_VM_DST0_INDEX) = dst_index;
        TR(ZKC_VM_DST0_PERFORMS_MEMORY_ACCESS) = dst0_mem; TR(ZKC_VM_NEW_SP) = new_sp;
        TR(ZKC_VM_SRC0_FROM_MEMORY) = src0_mem.is_pointer;
#pragma unroll
        for (int i = 0; i < 8; i++) TR(ZKC_VM_SRC0_FROM_MEMORY + 1 + i) = src0_mem.value[i];
    }
    zkc_vm_register src0 = SRCM(ZKC_MODE_REG_ONLY) ? draft_src0 : src0_mem;
    if (SRCM(ZKC_MODE_IMM16)) { src0 = reg_zero(); src0.value[0] = imm0; }
    const bool is_ptr_op = TYPE(ZKC_OP_PTR);
    const bool swap = ((TYPE(ZKC_OP_SUB) || TYPE(ZKC_OP_DIV) || TYPE(ZKC_OP_SHIFT)) && FLAG(ZKC_VM_SWAP_OPERANDS_FLAG_IDX)) ||
                      (is_ptr_op && FLAG(ZKC_VM_SWAP_OPERANDS_PTR_FLAG_IDX));
    zkc_vm_register ra = swap ? src1_register : src0, rb = swap ? src0 : src1_register;
    {
        const bool keep = TYPE(ZKC_OP_RET) || is_ptr_op || TYPE(ZKC_OP_UMA) || TYPE(ZKC_OP_FAR_CALL);
        if (ra.is_pointer && !keep && !is_kernel) { ra.is_pointer = 0; ra.value[1] = 0; ra.value[2] = 0; }
        if (rb.is_pointer && !is_kernel) { rb.is_pointer = 0; rb.value[1] = 0; rb.value[2] = 0; }
    }
    if (wr) {
        TR(ZKC_VM_SWAP_OPERANDS) = swap; TR(ZKC_VM_SRC0) = ra.is_pointer; TR(ZKC_VM_SRC1) = rb.is_pointer;
#pragma unroll
        for (int i = 0; i < 8; i++) { TR(ZKC_VM_SRC0 + 1 + i) = ra.value[i]; TR(ZKC_VM_SRC1 + 1 + i) = rb.value[i]; }
    }
    // ---- the selected opcode ----------------------------------------------------------------------------------------
    const U256 a = as_u256(ra), b = as_u256(rb);
    U256 d0, d1;
#pragma unroll
    for (int i = 0; i < 8; i++) { d0.v[i] = 0; d1.v[i] = 0; }
    uint32_t d0_is_ptr = 0;
    bool dst0_mem_capable = false, dst0_reg_only = false, write_dst1 = false, set_flags = false, new_pending = false;
    uint32_t nf0 = 0, nf1 = 0, nf2 = 0;
    const bool sf = FLAG(ZKC_VM_SET_FLAGS_FLAG_IDX);
    if (TYPE(ZKC_OP_ADD) || TYPE(ZKC_OP_SUB)) {
        const uint32_t of = TYPE(ZKC_OP_ADD) ? u256_add(a, b, d0) : u256_sub(a, b, d0);
        const bool z = u256_is_zero(d0);
        nf0 = of; nf1 = z; nf2 = !(of || z);
        set_flags = sf; dst0_mem_capable = true;
    } else if (TYPE(ZKC_OP_JUMP)) {
        d.pc = a.v[0] & 0xFFFF;
    } else if (TYPE(ZKC_OP_BINOP)) {
        const bool is_or = VAR(ZKC_VAR_BINOP_OR), is_and = VAR(ZKC_VAR_BINOP_AND);
#pragma unroll
        for (int i = 0; i < 8; i++) d0.v[i] = is_or ? (a.v[i] | b.v[i]) : (is_and ? (a.v[i] & b.v[i]) : (a.v[i] ^ b.v[i]));
        nf1 = u256_is_zero(d0);
        set_flags = sf; dst0_mem_capable = true;
    } else if (TYPE(ZKC_OP_MUL)) {
        u256_mul(a, b, d0, d1);
        const bool of = !u256_is_zero(d1), eq = u256_is_zero(d0);
        nf0 = of; nf1 = eq; nf2 = !of && !eq;
        set_flags = sf; dst0_mem_capable = true; write_dst1 = true;
    } else if (TYPE(ZKC_OP_DIV)) {
        const bool dz = u256_is_zero(b);
        if (!dz) u256_divrem(a, b, d0, d1);
        nf0 = dz; nf1 = !dz && u256_is_zero(d0); nf2 = !dz && u256_is_zero(d1);
        set_flags = sf; dst0_mem_capable = true; write_dst1 = true;
    } else if (TYPE(ZKC_OP_SHIFT)) {
        const bool is_rol = VAR(ZKC_VAR_SHIFT_ROL), is_ror = VAR(ZKC_VAR_SHIFT_ROR), is_shr = VAR(ZKC_VAR_SHIFT_SHR);
        const bool cyclic = is_rol || is_ror;
        uint32_t shift = b.v[0] & 0xFF;
        if (is_ror && shift != 0) shift = 256 - shift;
        if (is_shr) u256_shr(a, shift, d0);
        else {
            U256 lo, hi;
            u256_shl_wide(a, shift, lo, hi);
#pragma unroll
            for (int i = 0; i < 8; i++) d0.v[i] = lo.v[i] + (cyclic ? hi.v[i] : 0u);
        }
        nf1 = u256_is_zero(d0);
        set_flags = sf; dst0_mem_capable = true;
    } else if (is_ptr_op) {
        const bool v_add = VAR(ZKC_VAR_PTR_ADD), v_sub = VAR(ZKC_VAR_PTR_SUB), v_pack = VAR(ZKC_VAR_PTR_PACK), v_shrink = VAR(ZKC_VAR_PTR_SHRINK);
        const bool invalid_types = !(ra.is_pointer && !rb.is_pointer);
        const bool hi_nz = (b.v[1] | b.v[2] | b.v[3] | b.v[4] | b.v[5] | b.v[6] | b.v[7]) != 0, lo_nz = (b.v[0] | b.v[1] | b.v[2] | b.v[3]) != 0;
        const uint64_t addr = (uint64_t)a.v[0] + b.v[0];
        const bool panic = invalid_types || (hi_nz && (v_add || v_sub)) || (lo_nz && v_pack) || (v_add && (addr >> 32)) ||
                           (v_sub && a.v[0] < b.v[0]) || (v_shrink && a.v[3] < b.v[0]);
        new_pending = panic;
        d0 = a; d0_is_ptr = ra.is_pointer;
        if (v_add) d0.v[0] = (uint32_t)addr;
        if (v_sub) d0.v[0] = a.v[0] - b.v[0];
        if (v_shrink) d0.v[3] = a.v[3] - b.v[0];
        if (v_pack) { d0.v[4] = b.v[4]; d0.v[5] = b.v[5]; d0.v[6] = b.v[6]; d0.v[7] = b.v[7]; }
        dst0_mem_capable = !panic;
    } else if (TYPE(ZKC_OP_CONTEXT)) {
        d0.v[0] = VAR(ZKC_VAR_CONTEXT_ERGS_LEFT) ? ergs_left : new_sp;
        if (VAR(ZKC_VAR_CONTEXT_GET_U128)) for (int i = 0; i < 4; i++) d0.v[i] = ctx.context_u128_value_composite[i];
        if (VAR(ZKC_VAR_CONTEXT_THIS)) for (int i = 0; i < 5; i++) d0.v[i] = ctx.this_address[i];
        if (VAR(ZKC_VAR_CONTEXT_CALLER)) for (int i = 0; i < 5; i++) d0.v[i] = ctx.caller[i];
        if (VAR(ZKC_VAR_CONTEXT_CODE_ADDRESS)) for (int i = 0; i < 5; i++) d0.v[i] = ctx.code_address[i];
        if (VAR(ZKC_VAR_CONTEXT_META)) {
            for (int i = 0; i < 8; i++) d0.v[i] = 0;
            d0.v[0] = s.ergs_per_pubdata_byte; d0.v[2] = ctx.heap_upper_bound; d0.v[3] = ctx.aux_heap_upper_bound;
            d0.v[7] = ctx.this_shard_id | (ctx.caller_shard_id << 8) | (ctx.code_shard_id << 16);
        }
        const bool set_u128 = VAR(ZKC_VAR_CONTEXT_SET_U128), set_pubdata = VAR(ZKC_VAR_CONTEXT_SET_ERGS_PER_PUBDATA), inc_tx = VAR(ZKC_VAR_CONTEXT_INC_TX_NUMBER);
        dst0_reg_only = !(set_u128 || set_pubdata || inc_tx);
        d.set_u128 = set_u128; d.set_pubdata = set_pubdata; d.inc_tx = inc_tx;
#pragma unroll
        for (int i = 0; i < 4; i++) d.u128[i] = a.v[i];
        d.pubdata = a.v[0];
    }
    // ---- state diffs ---------------------------------------------------------------------------------------------------
    d.val0.is_pointer = d0_is_ptr; d.val1.is_pointer = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { d.val0.value[i] = d0.v[i]; d.val1.value[i] = d1.v[i]; }
    const bool perform_mem_write = dst0_mem && dst0_mem_capable;
    vm_push<SIM>(2, perform_mem_write, d, q, penc, ts0 + 3, stack_page, dst_index, 1, d.val0);
    if constexpr (SIM) { if (perform_mem_write && stack_page == mem->stack_page) mem->stack[dst_index] = d.val0; }
    const bool dst0_update_register = dst0_reg_only || (!dst0_mem && dst0_mem_capable);
    if (dst0_update_register) d.idx0 = dst0_r;
    if (write_dst1) d.idx1 = dst1_r;
    d.flags[0] = set_flags ? nf0 : f0; d.flags[1] = set_flags ? nf1 : f1; d.flags[2] = set_flags ? nf2 : f2;
    d.pending = new_pending;
    if (wr) {
        TR(ZKC_VM_DST0) = d.val0.is_pointer; TR(ZKC_VM_DST1) = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { TR(ZKC_VM_DST0 + 1 + i) = d.val0.value[i]; TR(ZKC_VM_DST1 + 1 + i) = d.val1.value[i]; }
        TR(ZKC_VM_PERFORM_DST0_MEMORY_WRITE) = perform_mem_write; TR(ZKC_VM_DST0_UPDATE_REGISTER) = dst0_update_register;
#pragma unroll
        for (int i = 0; i < 3; i++) TR(ZKC_VM_FLAGS_OUT + i) = d.flags[i];
        TR(ZKC_VM_PENDING_EXCEPTION_OUT) = d.pending; TR(ZKC_VM_PC_OUT) = d.pc; TR(ZKC_VM_ERGS_OUT) = d.ergs;
    }
#undef TR
#undef TYPE
#undef VAR
#undef FLAG
#undef SRCM
#undef DSTM
    return checks;
}

__device__ int vm_put_q12(uint64_t *dst, const zkc_queue_state12 &s) {
    for (int i = 0; i < 12; i++) dst[i] = s.head[i];
    for (int i = 0; i < 12; i++) dst[12 + i] = s.tail[i];
    dst[24] = s.length;
    return 25;
}

__global__ void vm_prologue_kernel(VmDev *devs, const zkc_vm_isa *isa, size_t n_instances) {
    const size_t inst = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= n_instances) return;
    VmDev *d = devs + inst;
    d->start = d->io.start_flag != 0;
    if (d->start) vm_initial_bootloader_state(d->io, *isa, d->s0);  // mod.rs:85-97
    else d->s0 = d->io.hidden_fsm_input;
    d->s_final = d->s0;
}

__device__ __forceinline__ void vm_report(VmDev *d, size_t row, uint32_t checks) {
    if (!checks) return;
    atomicOr(&d->failed_checks, checks);
    atomicMin(&d->first_bad, ((unsigned long long)row << 16) | checks);
}

// ---- one thread per cycle (of any instance of the batch) -----------------------------------------------------------------
constexpr int VM_WORDS = (int)(sizeof(zkc_vm_state) / 4);
constexpr int VM_DIFF_WORDS = (VM_WORDS + 31) / 32;
static_assert(sizeof(zkc_vm_state) == 1176 && VM_DIFF_WORDS == 10, "snapshot layout");
#define VW(f) ((int)(offsetof(zkc_vm_state, f) / 4))
#define VWC(f) ((int)((offsetof(zkc_vm_state, current_context) + offsetof(zkc_vm_context, f)) / 4))

struct VmMask { uint32_t w[VM_DIFF_WORDS]; };
__host__ __device__ constexpr VmMask vm_mask_clear(VmMask m, int lo, int n) {
    for (int i = lo; i < lo + n; i++) m.w[i >> 5] &= ~(1u << (i & 31));
    return m;
}
__host__ __device__ constexpr VmMask vm_mask_set(VmMask m, int lo, int n) {
    for (int i = lo; i < lo + n; i++) m.w[i >> 5] |= 1u << (i & 31);
    return m;
}
// words that may only change through an explicit flag of the delta: everything that is not padding, not a register,
// not one of the per-cycle scalars (those are compared with their expected value) and not the memory queue state
__host__ __device__ constexpr VmMask vm_keep_mask() {
    VmMask m{};
    m = vm_mask_set(m, 0, VM_WORDS);
    m = vm_mask_clear(m, VW(_pad), VW(current_context) - VW(_pad));  // _pad + the alignment hole behind it
    m = vm_mask_clear(m, VWC(aux_heap_upper_bound) + 1, 1);  // alignment hole in front of reverted_queue_head
    m = vm_mask_clear(m, VW(previous_code_word), 8);
    m = vm_mask_clear(m, VW(registers), 9 * ZKC_VM_REGISTERS);
    m = vm_mask_clear(m, VW(flags), 3);
    m = vm_mask_clear(m, VW(timestamp), 1);
    m = vm_mask_clear(m, VW(previous_code_page), 1);
    m = vm_mask_clear(m, VW(previous_super_pc), 1);
    m = vm_mask_clear(m, VW(pending_exception), 1);
    m = vm_mask_clear(m, VW(memory_queue_length), 1);
    m = vm_mask_clear(m, VWC(pc), 1);
    m = vm_mask_clear(m, VWC(sp), 1);
    m = vm_mask_clear(m, VWC(ergs_remaining), 1);
    m = vm_mask_clear(m, VW(memory_queue_state), 24);
    return m;
}
__host__ __device__ constexpr VmMask vm_memq_mask() {
    VmMask m{};
    return vm_mask_set(m, VW(memory_queue_state), 24);
}
static_assert(offsetof(zkc_vm_context, reverted_queue_head) == offsetof(zkc_vm_context, aux_heap_upper_bound) + 8, "context hole");
template <int K> struct VmKeepWord { static constexpr uint32_t value = vm_keep_mask().w[K]; };
template <int K> struct VmMemqWord { static constexpr uint32_t value = vm_memq_mask().w[K]; };
#define VM_FOR10(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9)

__device__ __forceinline__ bool reg_equal(const zkc_vm_register &a, const zkc_vm_register &b) {
    bool eq = a.is_pointer == b.is_pointer;
#pragma unroll
    for (int i = 0; i < 8; i++) eq &= a.value[i] == b.value[i];
    return eq;
}

// scratch of one batch: what the cycle launch leaves for the sponge launches
struct VmPushScratch {
    uint32_t *counts;  // [3]
    uint32_t *lists;   // [3][rows]: the rows whose push k happens, in no particular order
    uint8_t *mask;     // [rows]
    uint64_t *enc;     // [rows][3][8]
    uint64_t *state;   // [rows][3][12]: memory queue state after push k
};

// Every thread evaluates its cycle from snapshot `row` and checks that snapshot `row + 1` is the result.  The check
// has two parts.  (1) The warp walks its 32 consecutive snapshot pairs together: lane l compares words l, l+32, ...
// of snapshot c with snapshot c+1 -- fully coalesced, each snapshot is fetched once -- and the ballots of iteration c
// (a 294-bit "which words differ" mask) stay with lane c.  (2) The owner then demands that only words its cycle is
// allowed to change differ, and compares the changed ones (a few scalars, at most two registers) with the values the
// cycle produced.  The memory queue sponges are deferred: the cycle only emits their 8-word encodings.
__global__ void __launch_bounds__(128)
vm_cycles_kernel(VmDev *devs, const zkc_vm_isa *__restrict__ isa, const zkc_vm_state *__restrict__ snapshots,
                 const zkc_vm_cycle_witness *__restrict__ witness, uint64_t *__restrict__ trace, size_t limit, size_t n_instances,
                 VmPushScratch ps) {
    const size_t total = limit * n_instances;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = g < total;
    const unsigned lane = threadIdx.x & 31;
    const size_t inst = valid ? g / limit : 0, row = valid ? g - inst * limit : 0;
    const size_t idx = inst * (limit + 1) + row;
    // ---- (1) cooperative word diff --------------------------------------------------------------------------------
    uint32_t diff[VM_DIFF_WORDS];
#pragma unroll
    for (int k = 0; k < VM_DIFF_WORDS; k++) diff[k] = 0;
    {
        uint32_t cur[VM_DIFF_WORDS], nxt[VM_DIFF_WORDS];
#pragma unroll
        for (int k = 0; k < VM_DIFF_WORDS; k++) nxt[k] = 0;
        unsigned long long prev_idx = ~0ull;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
#pragma unroll 1
        for (int c = 0; c < 32; c++) {
            if (!((vmask >> c) & 1)) break;  // valid lanes are a prefix
            const unsigned long long ic = __shfl_sync(0xffffffffu, (unsigned long long)idx, c);
            const uint32_t *pc = reinterpret_cast<const uint32_t *>(snapshots + ic), *pn = pc + VM_WORDS;
            const bool chained = c > 0 && ic == prev_idx + 1;
#pragma unroll
            for (int k = 0; k < VM_DIFF_WORDS; k++) {
                const int j = (int)lane + 32 * k;
                const bool in = j < VM_WORDS;
                cur[k] = chained ? nxt[k] : (in ? __ldg(pc + j) : 0u);
                nxt[k] = in ? __ldg(pn + j) : 0u;
                const unsigned m = __ballot_sync(0xffffffffu, cur[k] != nxt[k]);
                if ((int)lane == c) diff[k] = m;
            }
            prev_idx = ic;
        }
    }
    // ---- the cycle ------------------------------------------------------------------------------------------------------
    VmDev *dev = devs + inst;
    const zkc_vm_state &s = snapshots[idx], &next = snapshots[idx + 1];
    uint32_t checks = 0, pmask = 0;
    if (valid) {
        if (row == 0 && !vm_state_equal(s, dev->s0)) checks |= ZKC_VM_CHK_SNAPSHOT;  // the hint chain starts at the circuit's own start state
        VmDelta d;
        checks |= vm_cycle_dev<false>(isa, s, d, witness[g], nullptr, nullptr, ps.enc + g * 24,
                                      trace ? trace + inst * (size_t)ZKC_VM_NUM_COLS * limit : nullptr, limit, row);
        pmask = d.push_mask;
        ps.mask[g] = (uint8_t)pmask;
        // ---- (2) is snapshot row + 1 what this cycle produces? ---------------------------------------------------------
        bool bad = false;
        if (d.set_u128) {
            diff[VW(context_composite_u128) >> 5] &= ~(15u << (VW(context_composite_u128) & 31));
            static_assert((VW(context_composite_u128) & 31) <= 28, "u128 words straddle a diff word");
            for (int i = 0; i < 4; i++) bad |= next.context_composite_u128[i] != d.u128[i];
        }
        if (d.set_pubdata) {
            diff[VW(ergs_per_pubdata_byte) >> 5] &= ~(1u << (VW(ergs_per_pubdata_byte) & 31));
            bad |= next.ergs_per_pubdata_byte != d.pubdata;
        }
        if (d.inc_tx) {
            diff[VW(tx_number_in_block) >> 5] &= ~(1u << (VW(tx_number_in_block) & 31));
            bad |= next.tx_number_in_block != s.tx_number_in_block + 1;
        }
        uint32_t stray = 0, memq_diff = 0;
#define X(K) stray |= diff[K] & VmKeepWord<K>::value; memq_diff |= diff[K] & VmMemqWord<K>::value;
        VM_FOR10(X)
#undef X
        bad |= stray != 0;
        if (!pmask) bad |= memq_diff != 0;  // otherwise the last sponge of the cycle compares (vm_memq_kernel)
        uint32_t regdiff = 0;
#pragma unroll
        for (int r = 0; r < ZKC_VM_REGISTERS; r++) {
            const int lo = VW(registers) + 9 * r;
            const uint32_t bits = __funnelshift_r(diff[lo >> 5], diff[(lo >> 5) + 1], lo & 31) & 0x1FFu;
            regdiff |= (bits != 0) << r;
        }
        const uint32_t may = (d.idx0 ? 1u << (d.idx0 - 1) : 0u) | (d.idx1 ? 1u << (d.idx1 - 1) : 0u);
        bad |= (regdiff & ~may) != 0;
        if (d.idx1) bad |= !reg_equal(next.registers[d.idx1 - 1], d.val1);
        if (d.idx0 && d.idx0 != d.idx1) bad |= !reg_equal(next.registers[d.idx0 - 1], d.val0);
        bad |= next.pending_exception != d.pending || next.current_context.pc != d.pc || next.current_context.sp != d.sp ||
               next.current_context.ergs_remaining != d.ergs || next.previous_code_page != d.prev_code_page ||
               next.previous_super_pc != d.prev_super_pc || next.timestamp != d.timestamp || next.memory_queue_length != d.memq_len;
#pragma unroll
        for (int i = 0; i < 3; i++) bad |= next.flags[i] != d.flags[i];
#pragma unroll
        for (int i = 0; i < 8; i++) bad |= next.previous_code_word[i] != d.cw[i];
        if (bad) {
            // attribute the broken link to the cycle that would consume the wrong snapshot (what a sequential run sees)
            if (row + 1 < limit) vm_report(dev, row + 1, ZKC_VM_CHK_SNAPSHOT);
            else checks |= ZKC_VM_CHK_SNAPSHOT;
        }
        if (row + 1 == limit) {  // the state the circuit ends in, as computed (its memory queue state: vm_memq_kernel)
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&s);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&dev->s_final);
            for (int i = 0; i < VM_WORDS; i++) dst[i] = src[i];
            vm_apply_delta(dev->s_final, d);
        }
        vm_report(dev, row, checks);
    }
    // ---- rows whose push k happens, for the dense sponge launches ---------------------------------------------------------
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const bool mine = (pmask >> k) & 1;
        const unsigned b = __ballot_sync(0xffffffffu, mine);
        if (!b) continue;
        const int leader = __ffs(b) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) base = atomicAdd(ps.counts + k, (uint32_t)__popc(b));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (mine) ps.lists[(size_t)k * total + base + __popc(b & ((1u << lane) - 1))] = (uint32_t)g;
    }
}

// push k of every cycle that has one: tail' = P(enc || tail[8..12]), one thread per push, all lanes busy.  The state it
// starts from is the cycle's previous push, or the snapshot; the last push of a cycle must land on the next snapshot.
__global__ void __launch_bounds__(128)
vm_memq_kernel(VmDev *devs, const zkc_vm_state *__restrict__ snapshots, VmPushScratch ps, int k, size_t limit, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ps.counts[k]) return;
    const size_t g = ps.lists[(size_t)k * total + i];
    const size_t inst = g / limit, row = g - inst * limit, idx = inst * (limit + 1) + row;
    const uint32_t m = ps.mask[g], before = m & ((1u << k) - 1);
    const uint64_t *from = before ? ps.state + (g * 3 + (31 - __clz(before))) * 12 : snapshots[idx].memory_queue_state;
    uint64_t q[12];
#pragma unroll
    for (int j = 0; j < 8; j++) q[j] = ps.enc[(g * 3 + k) * 8 + j];
#pragma unroll
    for (int j = 8; j < 12; j++) q[j] = from[j];
    poseidon2_permute(q);
    uint64_t *to = ps.state + (g * 3 + k) * 12;
#pragma unroll
    for (int j = 0; j < 12; j++) to[j] = q[j];
    if (m >> (k + 1)) return;
    VmDev *dev = devs + inst;
    const uint64_t *want = snapshots[idx + 1].memory_queue_state;
    bool same = true;
#pragma unroll
    for (int j = 0; j < 12; j++) same &= want[j] == q[j];
    if (!same) vm_report(dev, row + 1 < limit ? row + 1 : row, ZKC_VM_CHK_SNAPSHOT);
    if (row + 1 == limit)
        for (int j = 0; j < 12; j++) dev->s_final.memory_queue_state[j] = q[j];
}

// the 3 x 13 MEMQ_AFTER_* columns of the trace
__global__ void __launch_bounds__(256)
vm_memq_trace_kernel(const zkc_vm_state *__restrict__ snapshots, VmPushScratch ps, uint64_t *__restrict__ trace, size_t limit, size_t total) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const size_t inst = g / limit, row = g - inst * limit, idx = inst * (limit + 1) + row;
    uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit + row;
    const uint32_t m = ps.mask[g];
    const uint64_t *from = snapshots[idx].memory_queue_state;
    uint64_t len = snapshots[idx].memory_queue_length;
    constexpr int COL[3] = {ZKC_VM_MEMQ_AFTER_CODE, ZKC_VM_MEMQ_AFTER_SRC0, ZKC_VM_MEMQ_AFTER_DST0};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if ((m >> k) & 1) { from = ps.state + (g * 3 + k) * 12; len++; }
#pragma unroll
        for (int j = 0; j < 12; j++) t[(size_t)(COL[k] + j) * limit] = from[j];
        t[(size_t)(COL[k] + 12) * limit] = len;
    }
}

// 64 threads per instance: the four commitments of the closed form are independent sponges
// (ClosedFormInputCompactForm::from_full_form, fsm_input_output/mod.rs:178-255), each run by one 16-lane group with the
// 12-lane permutation (poseidon2.cuh) -- the 31 dependent permutations over a VM state are the latency of this launch.
// Each group flattens its encoding into its own global scratch row and absorbs from there, and skips the sponge whose
// result the start / completion flags mask to zero anyway; group 0 then commits the compact form.
constexpr int VM_FLAT_STRIDE = 248;
__global__ void __launch_bounds__(128)
vm_finalize_kernel(VmDev *devs, uint64_t *__restrict__ flat, size_t n_instances) {
    __shared__ uint64_t part[2][4][4];
    __shared__ uint64_t compact[2][24];
    const int slot = threadIdx.x >> 6, role = (threadIdx.x >> 4) & 3, i = threadIdx.x & 15;
    const unsigned gm = 0xFFFFu << (threadIdx.x & 16);
    const size_t inst = (size_t)blockIdx.x * 2 + slot;
    const bool active = inst < n_instances;
    VmDev *d = devs + (active ? inst : 0);
    zkc_vm_closed_form &io = d->io;
    const zkc_vm_state &state = d->s_final;
    const bool done = state.context_stack_depth == 0;  // mod.rs:113-122
    const bool start = d->start != 0;
    {
        const uint64_t *buf = flat + (inst * 4 + role) * VM_FLAT_STRIDE;
        int n = 0;
        const bool need = active && (role == 0 ? !done : role == 1 ? !start : role == 2 ? true : done);
        if (need) {
            if (i == 0) {
                uint64_t *w = flat + (inst * 4 + role) * VM_FLAT_STRIDE;
                if (role == 0) vm_flatten_state(state, w);                      // hidden FSM output
                else if (role == 1) vm_flatten_state(io.hidden_fsm_input, w);   // hidden FSM input
                else if (role == 2) {                                           // observable input (VmInputData)
                    int k = 0;
                    for (int j = 0; j < 4; j++) w[k++] = io.rollback_queue_tail_for_block[j];
                    for (int j = 0; j < 12; j++) w[k++] = io.memory_queue_initial_tail[j];
                    w[k++] = io.memory_queue_initial_length;
                    for (int j = 0; j < 12; j++) w[k++] = io.decommitment_queue_initial_tail[j];
                    w[k++] = io.decommitment_queue_initial_length;
                    w[k++] = io.zkporter_is_available;
                    for (int j = 0; j < 8; j++) w[k++] = io.default_aa_code_hash[j];
                } else {  // observable output (VmOutputData, mod.rs:124-196): log queue, memory queue, decommitment queue
                    int k = 0;
                    for (int j = 0; j < 4; j++) w[k++] = 0;
                    for (int j = 0; j < 4; j++) w[k++] = state.current_context.log_queue_forward_tail[j];
                    w[k++] = state.current_context.log_queue_forward_part_length;
                    for (int j = 0; j < 12; j++) w[k++] = 0;
                    for (int j = 0; j < 12; j++) w[k++] = state.memory_queue_state[j];
                    w[k++] = state.memory_queue_length;
                    for (int j = 0; j < 12; j++) w[k++] = 0;
                    for (int j = 0; j < 12; j++) w[k++] = state.code_decommittment_queue_state[j];
                    w[k++] = state.code_decommittment_queue_length;
                }
            }
            n = role < 2 ? ZKC_VM_STATE_FLAT : (role == 2 ? 39 : 59);
            __syncwarp(gm);
        }
        const uint64_t c = commit_encoding_coop(gm, buf, n, i);  // n == 0: no permutation, zeros
        if (i < 4) part[slot][role][i] = c;
    }
    __syncthreads();
    if (active && role == 0 && i == 0) {
        zkc_queue_state4 log_out;
        zkc_queue_state12 mem_out, dec_out;
        memset(&log_out, 0, sizeof log_out); memset(&mem_out, 0, sizeof mem_out); memset(&dec_out, 0, sizeof dec_out);
        if (done) {
            for (int j = 0; j < 12; j++) { mem_out.tail[j] = state.memory_queue_state[j]; dec_out.tail[j] = state.code_decommittment_queue_state[j]; }
            mem_out.length = state.memory_queue_length; dec_out.length = state.code_decommittment_queue_length;
            for (int j = 0; j < 4; j++) log_out.tail[j] = state.current_context.log_queue_forward_tail[j];
            log_out.length = state.current_context.log_queue_forward_part_length;
        }
        uint32_t checks = d->failed_checks;
        if (done && state.current_context.pc != 0) checks |= ZKC_VM_CHK_BOOTLOADER_EXIT;
        zkc_status st;
        st.code = ZKC_OK; st.cuda_error = 0; st.first_bad_row = -1; st.failed_checks = checks; st.reserved = 0;
        if (d->first_bad != ~0ull) st.first_bad_row = (int64_t)(d->first_bad >> 16);
        // most specific aggregate: broken snapshot chain > unsupported opcode > failed enforcement (order independent)
        if (checks) st.code = (checks & ZKC_VM_CHK_SNAPSHOT) ? ZKC_ERR_SNAPSHOT_MISMATCH
                            : (checks & ZKC_VM_CHK_UNSUPPORTED_OPCODE) ? ZKC_ERR_UNSUPPORTED : ZKC_ERR_UNSATISFIED;
        if (d->opt.compare_expected) {
            bool same = (io.completion_flag != 0) == done && vm_state_equal(io.hidden_fsm_output, state);
            for (int j = 0; j < 4; j++) same &= io.log_queue_final_state.head[j] == log_out.head[j] && io.log_queue_final_state.tail[j] == log_out.tail[j];
            same &= io.log_queue_final_state.length == log_out.length && io.memory_queue_final_state.length == mem_out.length &&
                    io.decommitment_queue_final_state.length == dec_out.length;
            for (int j = 0; j < 12; j++)
                same &= io.memory_queue_final_state.head[j] == mem_out.head[j] && io.memory_queue_final_state.tail[j] == mem_out.tail[j] &&
                        io.decommitment_queue_final_state.head[j] == dec_out.head[j] && io.decommitment_queue_final_state.tail[j] == dec_out.tail[j];
            if (!same && st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
        }
        io.log_queue_final_state = log_out; io.memory_queue_final_state = mem_out; io.decommitment_queue_final_state = dec_out;
        io.completion_flag = done;
        uint64_t *cf = compact[slot];
        cf[0] = start; cf[1] = done;
        for (int j = 0; j < 4; j++) {
            cf[2 + j] = part[slot][2][j];
            cf[6 + j] = part[slot][3][j];   // zero unless done
            cf[10 + j] = part[slot][1][j];  // zero if start
            cf[14 + j] = part[slot][0][j];  // zero if done
        }
        d->status = st;
    }
    __syncthreads();
    if (active) {  // the 64 threads of the instance publish the state the circuit ended in
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&state);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&io.hidden_fsm_output);
        for (int j = threadIdx.x & 63; j < (int)(sizeof(zkc_vm_state) / 4); j += 64) dst[j] = src[j];
    }
    if (role == 0) {
        const uint64_t c = commit_encoding_coop(gm, compact[slot], active ? 18 : 0, i);
        if (active && i < 4) d->commitment[i] = c;
    }
}

// ---- out-of-circuit run: one thread per independent VM instance ---------------------------------------------------------
__global__ void __launch_bounds__(32)
vm_simulate_kernel(const zkc_vm_isa *__restrict__ isa, const zkc_vm_state *__restrict__ initial, const uint32_t *__restrict__ code,
                   size_t code_words, size_t n_instances, size_t cycles, zkc_vm_register *__restrict__ pages,
                   zkc_vm_state *__restrict__ snapshots, zkc_vm_cycle_witness *__restrict__ witness, unsigned long long *first_bad,
                   uint32_t *failed_checks) {
    const size_t inst = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= n_instances) return;
    VmMemory mem;
    mem.code = pages + inst * 2 * 65536;
    mem.stack = mem.code + 65536;
    for (size_t i = 0; i < code_words && i < 65536; i++) {
        mem.code[i].is_pointer = 0;
        for (int j = 0; j < 8; j++) mem.code[i].value[j] = code[(inst * code_words + i) * 8 + j];
    }
    zkc_vm_state s = initial[inst];
    mem.code_page = s.current_context.code_page;
    mem.stack_page = s.current_context.base_page + 1;
    zkc_vm_state *snaps = snapshots + inst * (cycles + 1);
    zkc_vm_cycle_witness *wit = witness + inst * cycles;
    snaps[0] = s;
    for (size_t c = 0; c < cycles; c++) {
        zkc_vm_cycle_witness w;
        memset(&w, 0, sizeof w);
        VmDelta d;
        uint64_t q[12];
        for (int i = 0; i < 12; i++) q[i] = s.memory_queue_state[i];
        const uint32_t checks = vm_cycle_dev<true>(isa, s, d, w, &mem, q, nullptr, nullptr, 0, 0);
        vm_apply_delta(s, d);
        for (int i = 0; i < 12; i++) s.memory_queue_state[i] = q[i];
        wit[c] = w;
        snaps[c + 1] = s;
        if (checks) {
            atomicOr(failed_checks, checks);
            atomicMin(first_bad, ((unsigned long long)(inst * cycles + c) << 16) | checks);
        }
    }
}

__global__ void vm_initial_state_kernel(const zkc_vm_closed_form *io, const zkc_vm_isa *isa, zkc_vm_state *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) vm_initial_bootloader_state(*io, *isa, *out);
}

}  // namespace zkc

using namespace zkc;

extern "C" int zkc_main_vm_initial_state(zkc_ctx *ctx, const zkc_vm_closed_form *io, const zkc_vm_isa *isa, zkc_vm_state *out) {
    if (!ctx || !io || !isa || !out) return ZKC_ERR_INVALID_ARGUMENT;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    const size_t bytes = zkc_carver::bytes(1, sizeof(zkc_vm_closed_form)) + zkc_carver::bytes(1, sizeof(zkc_vm_isa)) + zkc_carver::bytes(1, sizeof(zkc_vm_state));
    void *blk = ctx->scratch(bytes);
    if (!blk) return ZKC_ERR_CUDA;
    zkc_carver cv(blk);
    zkc_vm_closed_form *dio = cv.take<zkc_vm_closed_form>(1);
    zkc_vm_isa *disa = cv.take<zkc_vm_isa>(1);
    zkc_vm_state *dout = cv.take<zkc_vm_state>(1);
    cudaStream_t s = ctx->stream;
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(dio, io, sizeof *io, cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(disa, isa, sizeof *isa, cudaMemcpyHostToDevice, s));
    ZKC_LAUNCH(ctx, "vm_initial_state", vm_initial_state_kernel, 1, 32, 0, dio, disa, dout);
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(out, dout, sizeof *out, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    return ZKC_OK;
}

extern "C" int zkc_main_vm_entry_point_batch(zkc_ctx *ctx, zkc_vm_closed_form *ios, size_t n_instances, const zkc_vm_isa *isa,
                                             const zkc_vm_state *snapshots, const zkc_vm_cycle_witness *witness, size_t limit,
                                             const zkc_vm_options *options, int on_device, uint64_t *trace, uint64_t *commitments,
                                             zkc_status *statuses) {
    if (!ctx || !ios || !isa || !commitments || !statuses || (limit && n_instances && (!snapshots || !witness)) ||
        limit > 0x0FFFFFFFull || n_instances > 0x00FFFFFFull || limit * n_instances > 0xFFFFFFFFull)
        return ZKC_ERR_INVALID_ARGUMENT;
    if (!n_instances) return ZKC_OK;
    zkc_status *status = statuses;
    for (size_t i = 0; i < n_instances; i++) statuses[i] = zkc_status{ZKC_OK, 0, -1, 0, 0};
    const bool in_dev = on_device & ZKC_INPUTS_ON_DEVICE, trace_dev = on_device & ZKC_TRACE_ON_DEVICE;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const size_t rows = limit * n_instances;
    size_t bytes = zkc_carver::bytes(n_instances, sizeof(VmDev)) + zkc_carver::bytes(1, sizeof(zkc_vm_isa));
    if (!in_dev) bytes += zkc_carver::bytes(rows + n_instances, sizeof(zkc_vm_state)) + zkc_carver::bytes(rows + 1, sizeof(zkc_vm_cycle_witness));
    if (trace && !trace_dev) bytes += zkc_carver::bytes((size_t)ZKC_VM_NUM_COLS * rows, 8);
    bytes += zkc_carver::bytes(n_instances * 4 * VM_FLAT_STRIDE, 8);
    bytes += zkc_carver::bytes(4, 4) + zkc_carver::bytes(3 * rows, 4) + zkc_carver::bytes(rows, 1) + zkc_carver::bytes(rows * 24, 8) +
             zkc_carver::bytes(rows * 36, 8);
    void *blk = ctx->scratch(bytes);
    VmDev *h = (VmDev *)ctx->pinned(n_instances * sizeof(VmDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    VmDev *d = cv.take<VmDev>(n_instances);
    zkc_vm_isa *disa = cv.take<zkc_vm_isa>(1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, n_instances * sizeof(VmDev));
    for (size_t i = 0; i < n_instances; i++) {
        h[i].io = ios[i];
        if (options) h[i].opt = *options;
        h[i].limit = limit;
        h[i].first_bad = ~0ull;
    }
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, n_instances * sizeof(VmDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(disa, isa, sizeof(zkc_vm_isa), cudaMemcpyHostToDevice, s));  // the ISA tables are always host data
    const zkc_vm_state *dsnap = snapshots;
    const zkc_vm_cycle_witness *dwit = witness;
    uint64_t *dtrace = trace;
    if (!in_dev && limit) {
        zkc_vm_state *bs = cv.take<zkc_vm_state>(rows + n_instances);
        zkc_vm_cycle_witness *bw = cv.take<zkc_vm_cycle_witness>(rows + 1);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(bs, snapshots, (rows + n_instances) * sizeof(zkc_vm_state), cudaMemcpyHostToDevice, s));
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(bw, witness, rows * sizeof(zkc_vm_cycle_witness), cudaMemcpyHostToDevice, s));
        dsnap = bs; dwit = bw;
    }
    if (trace && !trace_dev) dtrace = cv.take<uint64_t>((size_t)ZKC_VM_NUM_COLS * rows);
    uint64_t *flat = cv.take<uint64_t>(n_instances * 4 * VM_FLAT_STRIDE);
    VmPushScratch ps;
    ps.counts = cv.take<uint32_t>(4);
    ps.lists = cv.take<uint32_t>(3 * rows);
    ps.mask = cv.take<uint8_t>(rows);
    ps.enc = cv.take<uint64_t>(rows * 24);
    ps.state = cv.take<uint64_t>(rows * 36);
    ZKC_CUDA(ctx, status, cudaMemsetAsync(ps.counts, 0, 16, s));
    ZKC_LAUNCH(ctx, "vm_prologue", vm_prologue_kernel, (unsigned)((n_instances + 31) / 32), 32, 0, d, disa, n_instances);
    if (rows) {
        ZKC_LAUNCH(ctx, "vm_cycles", vm_cycles_kernel, (unsigned)((rows + 127) / 128), 128, 0, d, disa, dsnap, dwit, dtrace, limit, n_instances, ps);
        // the lists live on the device: size every sponge launch for the worst case, surplus threads leave at once
        for (int k = 0; k < 3; k++)
            ZKC_LAUNCH(ctx, "vm_memq", vm_memq_kernel, (unsigned)((rows + 127) / 128), 128, 0, d, dsnap, ps, k, limit, rows);
        if (dtrace) ZKC_LAUNCH(ctx, "vm_memq_trace", vm_memq_trace_kernel, (unsigned)((rows + 255) / 256), 256, 0, dsnap, ps, dtrace, limit, rows);
    }
    ZKC_LAUNCH(ctx, "vm_finalize", vm_finalize_kernel, (unsigned)((n_instances + 1) / 2), 128, 0, d, flat, n_instances);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, n_instances * sizeof(VmDev), cudaMemcpyDeviceToHost, s));
    if (!trace_dev && trace && rows)
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(trace, dtrace, (size_t)ZKC_VM_NUM_COLS * rows * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    int worst = ZKC_OK;
    for (size_t i = 0; i < n_instances; i++) {
        ios[i].hidden_fsm_output = h[i].io.hidden_fsm_output;
        ios[i].log_queue_final_state = h[i].io.log_queue_final_state;
        ios[i].memory_queue_final_state = h[i].io.memory_queue_final_state;
        ios[i].decommitment_queue_final_state = h[i].io.decommitment_queue_final_state;
        ios[i].completion_flag = h[i].io.completion_flag;
        memcpy(commitments + 4 * i, h[i].commitment, 32);
        statuses[i] = h[i].status;
        if (statuses[i].code != ZKC_OK && worst == ZKC_OK) worst = statuses[i].code;
    }
    return worst;
}

extern "C" int zkc_main_vm_entry_point(zkc_ctx *ctx, zkc_vm_closed_form *io, const zkc_vm_isa *isa, const zkc_vm_state *snapshots,
                                       const zkc_vm_cycle_witness *witness, size_t limit, const zkc_vm_options *options,
                                       int on_device, uint64_t *trace, uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!io || !commitment) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    const int rc = zkc_main_vm_entry_point_batch(ctx, io, 1, isa, snapshots, witness, limit, options, on_device, trace, commitment, status);
    if (rc == ZKC_ERR_INVALID_ARGUMENT) status->code = rc;
    return rc;
}

extern "C" int zkc_main_vm_simulate(zkc_ctx *ctx, const zkc_vm_isa *isa, const zkc_vm_state *initial_states, const uint32_t *code,
                                    size_t code_words, size_t n_instances, size_t cycles, zkc_vm_state *snapshots_out,
                                    zkc_vm_cycle_witness *witness_out, zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !isa || !initial_states || !code || !snapshots_out || !witness_out || code_words > 65536) {
        status->code = ZKC_ERR_INVALID_ARGUMENT;
        return ZKC_ERR_INVALID_ARGUMENT;
    }
    if (!n_instances) return ZKC_OK;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const size_t bytes = zkc_carver::bytes(1, sizeof(zkc_vm_isa)) + zkc_carver::bytes(4, 8) +
                         zkc_carver::bytes(n_instances * 2 * 65536, sizeof(zkc_vm_register));
    void *blk = ctx->scratch(bytes);
    unsigned long long *hres = (unsigned long long *)ctx->pinned(32);
    if (!blk || !hres) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    zkc_vm_isa *disa = cv.take<zkc_vm_isa>(1);
    unsigned long long *dres = cv.take<unsigned long long>(4);
    zkc_vm_register *pages = cv.take<zkc_vm_register>(n_instances * 2 * 65536);
    cudaStream_t s = ctx->stream;
    hres[0] = ~0ull; hres[1] = 0;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(disa, isa, sizeof(zkc_vm_isa), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(dres, hres, 16, cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(pages, 0, n_instances * 2 * 65536 * sizeof(zkc_vm_register), s));
    ZKC_LAUNCH(ctx, "vm_simulate", vm_simulate_kernel, (unsigned)((n_instances + 31) / 32), 32, 0, disa, initial_states, code, code_words,
               n_instances, cycles, pages, snapshots_out, witness_out, dres, (uint32_t *)(dres + 1));
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(hres, dres, 16, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    const uint32_t checks = (uint32_t)hres[1];
    if (checks) {
        status->failed_checks = checks;
        status->first_bad_row = (int64_t)(hres[0] >> 16);
        status->code = (checks & ZKC_VM_CHK_UNSUPPORTED_OPCODE) ? ZKC_ERR_UNSUPPORTED : ZKC_ERR_UNSATISFIED;
    }
    return status->code;
}
