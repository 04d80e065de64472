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
// Built opcodes: nop, add, sub, jump, binop, mul, div, shifts, ptr, context, uma, log, near_call, ret and every src0 / dst0
// addressing mode (exceptions included: the circuit masks them into ret.panic).  far_call reports ZKC_ERR_UNSUPPORTED.
//
// Poseidon2 relations of a cycle (1 opcode fetch + up to 8, cycle.rs:620-795) are NOT run by the cycle's thread: the
// cycle emits sponge JOBS (8 absorbed elements + where the capacity comes from + what the output must equal) and the
// jobs of every cycle run afterwards as dense launches, one slot at a time (vm_sponge_kernel), so a warp never waits
// for a lane that hashes.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <utility>
#include <cuda.h>  // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no link dependency on libcuda)
#include "ctx.cuh"
#include "poseidon2.cuh"
#include "u256.cuh"

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
    uint64_t hint_commitment[4];  // commitment computed ahead of time from the host's final snapshot (vm_finalize_kernel, mode 0)
    uint32_t hint_ok, pad2;
};

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

// ---- memory / storage model of the out-of-circuit run: code, stack, heap and aux heap pages of the root frame ---------
constexpr uint32_t VM_PAGE_WORDS = 65536, VM_STORAGE_SLOTS = 4096, VM_SIM_MAX_DEPTH = 4096;
struct VmSlot { uint32_t used, written, key[8], value[8]; };
struct VmEntry {  // one rollback-queue event of a frame: its own call marker or a revertable log
    long long prev;
    uint32_t kind;   // 1 call marker, 2 log
    int slot;        // storage slot a storage write touched, or -1
    uint32_t prev_value[8], prev_written;
    uint64_t enc16[4], cap[4];
};
struct VmSim {
    zkc_vm_register *pages[4];
    uint32_t page_ids[4];
    VmSlot *storage;
    zkc_vm_callstack_witness *stack;  // saved frames, [VM_SIM_MAX_DEPTH]
    zkc_vm_callstack_witness *cw_out;
    uint32_t cw_cap, n_cw;
    int ev_kind;  // 0 none, 1 call, 2 ret ok, 3 ret revert / panic, 4 revertable log
    VmEntry ev;
    int overflow;
};
__device__ zkc_vm_register reg_zero();
__device__ zkc_vm_register sim_read(const VmSim &m, uint32_t page, uint32_t index) {
    if (index < VM_PAGE_WORDS)
        for (int k = 0; k < 4; k++) if (page == m.page_ids[k]) return m.pages[k][index];
    return reg_zero();
}
__device__ void sim_write(VmSim &m, uint32_t page, uint32_t index, const zkc_vm_register &v) {
    if (index >= VM_PAGE_WORDS) return;
    for (int k = 1; k < 4; k++) if (page == m.page_ids[k]) { m.pages[k][index] = v; return; }
}
__device__ int sim_slot(VmSim &m, const uint32_t *key) {
    uint32_t h = 0x9E3779B9u;
    for (int i = 0; i < 8; i++) h = (h ^ key[i]) * 0x85EBCA6Bu + (h >> 15);
    for (uint32_t probe = 0; probe < VM_STORAGE_SLOTS; probe++) {
        VmSlot &s = m.storage[(h + probe) % VM_STORAGE_SLOTS];
        bool same = s.used != 0;
        if (same) for (int i = 0; i < 8; i++) same &= s.key[i] == key[i];
        if (!s.used) { s.used = 1; for (int i = 0; i < 8; i++) s.key[i] = key[i]; return (int)((h + probe) % VM_STORAGE_SLOTS); }
        if (same) return (int)((h + probe) % VM_STORAGE_SLOTS);
    }
    m.overflow = 1;
    return 0;
}

__device__ __forceinline__ bool prop(uint64_t props, int bit) { return (props >> bit) & 1; }
__device__ zkc_vm_register reg_zero() {
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

// sponge slots in use (far calls would add 5..8) and where a job's capacity comes from / what its output must equal
constexpr int VM_JOB_SLOTS = ZKC_VM_NUM_SPONGES;
// slots 0..4 are the common ones; the far call's 5..8 live in their own (sparsely touched) scratch arrays so that the common
// slots keep their dense per-row stride
constexpr int VM_JOB_SLOTS_LO = 5, VM_JOB_SLOTS_HI = VM_JOB_SLOTS - VM_JOB_SLOTS_LO;
enum : uint32_t { VM_CAP_ZERO = 9, VM_CAP_MEMQ = 10, VM_CAP_STACK = 11, VM_CAP_CALLSTACK_WITNESS = 12, VM_CAP_DECOMMIT = 13 };
enum : uint32_t { VM_CHK_NONE = 0, VM_CHK_NEXT_MEMQ, VM_CHK_NEXT_STACK, VM_CHK_CUR_STACK, VM_CHK_NEXT_FWD_TAIL, VM_CHK_CUR_RB_HEAD,
                  VM_CHK_NEXT_DECOMMIT };

// what one cycle changes in a VmLocalState; every other word must carry over unchanged
struct VmDelta {
    uint32_t pending, pc, sp, ergs, prev_code_page, prev_super_pc, timestamp, memq_len;
    uint32_t flags[3];
    uint32_t cw[8];       // previous_code_word after the opcode fetch
    uint32_t idx0, idx1;  // 1-based register written by dst0 / dst1, 0 = none (dst1 is applied after dst0)
    zkc_vm_register val0, val1;
    uint32_t set_u128, u128[4], set_pubdata, pubdata, inc_tx;
    uint32_t heap_bound, aux_bound, fwd_len, rb_len;
    uint64_t rb_head[4];
    uint32_t fwd_tail_kind;  // 0 unchanged, 1 output of the log's forward sponge (slot 3), 2 explicit (ret)
    uint64_t fwd_tail[4];
    uint32_t ctx_replaced;   // 0 no, 1 near call, 2 ret, 3 far call: the whole current context is nctx (forward tail / length apart)
    uint32_t far_ret;        // r1 = r1_val, r2..r15 zeroed
    uint32_t far_call;       // r1 = r1_val, r2 = r2_low, ABI / reserved / implicit registers cleaned (far_call_system: ABI values kept)
    uint32_t far_call_system, r2_low;
    zkc_vm_register r1_val;
    uint32_t depth;
    uint32_t page_counter, decommit_len;
    uint32_t cw_index;       // ret: the callstack witness used
    uint32_t job_mask;       // sponge jobs: bit k per slot
    uint64_t cap_from, chk;  // nibble k per slot
    uint64_t *penc_hi;       // circuit mode: where the encodings of slots 5..8 go (set by the caller)
};
// register r (0-based) after a far call (far_call.rs:1018-1066)
__device__ __forceinline__ zkc_vm_register vm_far_call_register(const zkc_vm_isa *isa, const VmDelta &d, int r, const zkc_vm_register &old) {
    if (r == 0) return d.r1_val;
    zkc_vm_register v = reg_zero();
    if (r == 1) { v.value[0] = d.r2_low; return v; }
    if ((uint32_t)r >= isa->call_system_abi_registers[0] && (uint32_t)r < isa->call_system_abi_registers[1]) {
        if (d.far_call_system) { v = old; v.is_pointer = 0; }
        return v;
    }
    if (((uint32_t)r >= isa->call_reserved_range[0] && (uint32_t)r < isa->call_reserved_range[1]) || (uint32_t)r == isa->call_implicit_parameter_reg_idx) return v;
    return old;
}
// sponge outputs of a cycle: known at once only to the out-of-circuit run
struct VmSimOut { uint64_t memq[12], stack[12], fwd_tail[4]; };

__device__ void vm_apply_delta(zkc_vm_state &t, const VmDelta &d, const zkc_vm_context &nctx, const zkc_vm_isa *isa) {
    for (int i = 0; i < 8; i++) t.previous_code_word[i] = d.cw[i];
    if (d.idx0) t.registers[d.idx0 - 1] = d.val0;
    if (d.far_ret) {
        t.registers[0] = d.r1_val;
        for (int r = 1; r < ZKC_VM_REGISTERS; r++) t.registers[r] = reg_zero();
    }
    if (d.far_call)
        for (int r = 0; r < ZKC_VM_REGISTERS; r++) t.registers[r] = vm_far_call_register(isa, d, r, t.registers[r]);
    t.memory_page_counter = d.page_counter;
    t.code_decommittment_queue_length = d.decommit_len;
    if (d.idx1) t.registers[d.idx1 - 1] = d.val1;
    if (d.set_u128) for (int i = 0; i < 4; i++) t.context_composite_u128[i] = d.u128[i];
    if (d.set_pubdata) t.ergs_per_pubdata_byte = d.pubdata;
    if (d.inc_tx) t.tx_number_in_block += 1;
    if (d.ctx_replaced) {
        const uint64_t t0 = t.current_context.log_queue_forward_tail[0], t1 = t.current_context.log_queue_forward_tail[1],
                       t2 = t.current_context.log_queue_forward_tail[2], t3 = t.current_context.log_queue_forward_tail[3];
        t.current_context = nctx;
        t.current_context.log_queue_forward_tail[0] = t0; t.current_context.log_queue_forward_tail[1] = t1;
        t.current_context.log_queue_forward_tail[2] = t2; t.current_context.log_queue_forward_tail[3] = t3;
    } else {
        t.current_context.pc = d.pc; t.current_context.sp = d.sp; t.current_context.ergs_remaining = d.ergs;
        t.current_context.heap_upper_bound = d.heap_bound; t.current_context.aux_heap_upper_bound = d.aux_bound;
        t.current_context.reverted_queue_segment_len = d.rb_len;
        for (int i = 0; i < 4; i++) t.current_context.reverted_queue_head[i] = d.rb_head[i];
    }
    t.current_context.log_queue_forward_part_length = d.fwd_len;
    if (d.fwd_tail_kind == 2) for (int i = 0; i < 4; i++) t.current_context.log_queue_forward_tail[i] = d.fwd_tail[i];
    t.context_stack_depth = d.depth;
    t.pending_exception = d.pending;
    t.previous_code_page = d.prev_code_page; t.previous_super_pc = d.prev_super_pc; t.timestamp = d.timestamp;
    t.memory_queue_length = d.memq_len;
    for (int i = 0; i < 3; i++) t.flags[i] = d.flags[i];
}

// One Poseidon2 relation of the cycle: 8 absorbed elements over a capacity.  SIM runs it at once (cap = the 4 capacity
// elements, out = the 12-element result); the circuit kernels record the job for the dense sponge launches.
template <bool SIM>
__device__ __forceinline__ void vm_job(int slot, VmDelta &d, uint64_t *penc, const uint64_t (&in8)[8], const uint64_t *cap,
                                       uint32_t cap_code, uint32_t chk, uint64_t *out) {
    d.job_mask |= 1u << slot;
    if constexpr (SIM) {
        uint64_t t[12];
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = in8[i];
#pragma unroll
        for (int i = 0; i < 4; i++) t[8 + i] = cap[i];
        poseidon2_permute(t);
#pragma unroll
        for (int i = 0; i < 12; i++) out[i] = t[i];
    } else {
        d.cap_from |= (uint64_t)cap_code << (4 * slot);
        d.chk |= (uint64_t)chk << (4 * slot);
        uint64_t *dst = slot < VM_JOB_SLOTS_LO ? penc + 8 * slot : d.penc_hi + 8 * (slot - VM_JOB_SLOTS_LO);
#pragma unroll
        for (int i = 0; i < 8; i++) dst[i] = in8[i];
    }
}

// memory queue push: tail' = P(enc || tail[8..12]) (main_vm/utils.rs:194-230, :442-515, cycle.rs:845-905, uma.rs:362-812).
// `last` = slot of the cycle's previous memory queue job (-1: none yet), updated.
template <bool SIM>
__device__ __forceinline__ void vm_push(int slot, bool execute, VmDelta &d, VmSimOut *so, uint64_t *penc, int &last, uint32_t ts,
                                        uint32_t page, uint32_t index, uint32_t rw, const zkc_vm_register &val) {
    if (!execute) return;
    uint64_t e[8];
    vm_mq_encode(ts, page, index, rw, val, e);
    d.memq_len++;
    if constexpr (SIM) vm_job<true>(slot, d, penc, e, so->memq + 8, 0, 0, so->memq);
    else vm_job<false>(slot, d, penc, e, nullptr, last < 0 ? VM_CAP_MEMQ : (uint32_t)last, VM_CHK_NONE, nullptr);
    last = slot;
}

// FatPtrInABI::parse_and_validate, call_ret_impl/far_call.rs:140-196
struct VmFatPtr { uint32_t offset, page, start, length; };
__device__ VmFatPtr vm_fat_ptr_parse(const U256 &v, bool as_fresh, uint32_t &upper_bound, bool &non_addressable) {
    VmFatPtr p{v.v[0], v.v[1], v.v[2], v.v[3]};
    const uint64_t end = (uint64_t)p.start + p.length;
    const bool range_of = (end >> 32) != 0;
    const bool invalid = (p.offset != 0 && as_fresh) || range_of || p.length < p.offset;
    if (invalid) p = VmFatPtr{0, 0, 0, 0};
    upper_bound = (uint32_t)end; non_addressable = range_of;
    return p;
}

// ExecutionContextRecord in flatten (allocation) order: 42 elements
__device__ void vm_flatten_record(const zkc_vm_context &c, uint64_t *dst) {
    int n = 0;
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
}

// LogQuery::encode, base_structures/log_query/mod.rs:121-517, from the pieces the log opcode has
__device__ void vm_log_encode(const uint32_t *address, const uint32_t *key, const uint32_t *read_value, const uint32_t *written_value,
                              uint32_t tx, uint32_t ts, uint32_t aux, uint32_t shard, uint32_t rw, uint32_t service, uint64_t *out) {
    uint8_t b[52];
    for (int l = 0; l < 8; l++)
        for (int j = 0; j < 4; j++) b[4 * l + j] = (uint8_t)(key[l] >> (8 * j));
    for (int l = 0; l < 5; l++)
        for (int j = 0; j < 4; j++) b[32 + 4 * l + j] = (uint8_t)(address[l] >> (8 * j));
    for (int i = 0; i < 16; i++) {
        const uint64_t w = i < 8 ? read_value[i] : written_value[i - 8];
        out[i] = w + ((uint64_t)b[3 * i] << 32) + ((uint64_t)b[3 * i + 1] << 40) + ((uint64_t)b[3 * i + 2] << 48);
    }
    out[16] = (uint64_t)ts + ((uint64_t)b[48] << 32) + ((uint64_t)b[49] << 40) + ((uint64_t)b[50] << 48);
    out[17] = (uint64_t)tx + ((uint64_t)b[51] << 32) + ((uint64_t)aux << 40) + ((uint64_t)shard << 48);
    out[18] = (uint64_t)rw + 2 * (uint64_t)service;
    out[19] = 0;
}

// One vm_cycle from the state `s` (read only; only the words a cycle needs are touched): returns the check bits, what
// changes in `d` (+ `nctx` when the callstack moves).  SIM: oracle answers come from `sim` and are recorded into `w` /
// the callstack witness, sponges run at once into `so`; otherwise answers come from `w` / `cw` and the sponges are
// emitted as jobs into `penc`.  trace / limit / row: where to put the row (trace may be null; the sponge columns are
// written by whoever runs the sponges).  next: the following snapshot (circuit mode; forward tail column only).
template <bool SIM, typename W, typename RS>
__device__ __forceinline__ uint32_t vm_cycle_dev(const zkc_vm_isa *__restrict__ isa, const zkc_vm_closed_form *gc, const zkc_vm_state &s, const RS &regs,
                                                 VmDelta &d, zkc_vm_context &nctx, W &w,
                                                 const zkc_vm_callstack_witness *__restrict__ cw, uint32_t n_cw, VmSim *sim, VmSimOut *so,
                                                 uint64_t *penc, const uint64_t *next_fwd_tail, uint64_t *__restrict__ trace, size_t limit, size_t row,
                                                 int aux_base = ZKC_VM_OP_AUX) {
#define TR(col) trace[(size_t)(col) * limit + row]
    const bool wr = trace != nullptr;
    uint32_t checks = 0;
    const zkc_vm_context &ctx = s.current_context;
    d.set_u128 = 0; d.set_pubdata = 0; d.inc_tx = 0; d.idx0 = 0; d.idx1 = 0;
    d.job_mask = 0; d.cap_from = 0; d.chk = 0; d.fwd_tail_kind = 0; d.ctx_replaced = 0; d.far_ret = 0; d.cw_index = 0;
    d.far_call = 0; d.far_call_system = 0; d.r2_low = 0;
    d.page_counter = s.memory_page_counter; d.decommit_len = s.code_decommittment_queue_length;
    d.memq_len = s.memory_queue_length;
    d.heap_bound = ctx.heap_upper_bound; d.aux_bound = ctx.aux_heap_upper_bound;
    d.fwd_len = ctx.log_queue_forward_part_length; d.rb_len = ctx.reverted_queue_segment_len;
#pragma unroll
    for (int i = 0; i < 4; i++) d.rb_head[i] = ctx.reverted_queue_head[i];
    d.depth = s.context_stack_depth;
    int last_memq = -1;
    if constexpr (SIM) sim->ev_kind = 0;
    // ---- create_prestate ---------------------------------------------------------------------------------------
    const bool should_skip = s.context_stack_depth == 0;
    const bool pending = s.pending_exception != 0;
    const bool should_try_read = !should_skip && !pending;
    const uint32_t pc = ctx.pc, super_pc = pc >> 2, sub_pc = pc & 3, pc_plus_one = (pc + 1) & 0xFFFF;
    const uint32_t code_page = ctx.code_page;
    const bool should_read_opcode = should_try_read && !(s.previous_code_page == code_page && super_pc == s.previous_super_pc);
    const uint32_t ts0 = s.timestamp;
    zkc_vm_register code_val = reg_zero();
    if (should_read_opcode) {
        if constexpr (SIM) {
            code_val = sim_read(*sim, code_page, super_pc); code_val.is_pointer = 0;
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
    vm_push<SIM>(0, should_read_opcode, d, so, penc, last_memq, ts0, code_page, super_pc, 0, code_val);
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
    d.pc = should_skip ? pc : pc_plus_one;
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
    if constexpr (SIM) { if (TYPE(ZKC_OP_FAR_CALL)) checks |= ZKC_VM_CHK_UNSUPPORTED_OPCODE; }  // the GPU run's memory model has one frame's pages
    if (wr) {
        TR(ZKC_VM_VARIANT) = variant; TR(ZKC_VM_CONDITION_IDX) = cond_idx; TR(ZKC_VM_CONDITION) = condition; TR(ZKC_VM_ERGS_COST) = cost;
        TR(ZKC_VM_OUT_OF_ERGS) = out_of_ergs; TR(ZKC_VM_KERNEL_MODE_EXCEPTION) = kernel_exc; TR(ZKC_VM_STATIC_EXCEPTION) = static_exc;
        TR(ZKC_VM_CALLSTACK_IS_FULL) = callstack_full; TR(ZKC_VM_EXPLICIT_PANIC) = explicit_panic; TR(ZKC_VM_MASK_INTO_PANIC) = mask_into_panic;
        TR(ZKC_VM_MASK_INTO_NOP) = mask_into_nop; TR(ZKC_VM_PROPS) = props; TR(ZKC_VM_DIRTY_ERGS_LEFT) = ergs_left;
        TR(ZKC_VM_SRC0_REG) = src0_r; TR(ZKC_VM_SRC1_REG) = src1_r; TR(ZKC_VM_DST0_REG) = dst0_r; TR(ZKC_VM_DST1_REG) = dst1_r;
        TR(ZKC_VM_IMM0) = imm0; TR(ZKC_VM_IMM1) = imm1;
    }
    // ---- operands -------------------------------------------------------------------------------------------------
    const zkc_vm_register draft_src0 = src0_r ? regs.get(src0_r - 1) : reg_zero();
    const zkc_vm_register src1_register = src1_r ? regs.get(src1_r - 1) : reg_zero();
    const uint32_t src0_low = draft_src0.value[0] & 0xFFFF;
    const uint32_t dst0_low = (dst0_r ? regs.low(dst0_r - 1) : 0u) & 0xFFFF;
    const uint32_t current_sp = ctx.sp, stack_page = ctx.base_page + 1, heap_page = ctx.base_page + 2, aux_heap_page = ctx.base_page + 3;
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
            src0_mem = sim_read(*sim, src_page, src_index);
            if (src_page == code_page) src0_mem.is_pointer = 0;
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
    vm_push<SIM>(1, should_read_src0, d, so, penc, last_memq, ts0, src_page, src_index, 0, src0_mem);
    if (wr) {
        TR(ZKC_VM_SRC0_PAGE) = src_page; TR(ZKC_VM_SRC0_INDEX) = src_index; TR(ZKC_VM_SHOULD_READ_SRC0) = should_read_src0;
        TR(ZKC_VM_SP_AFTER_SRC0) = sp_after_src0; TR(ZKC_VM_DST0_PAGE) = stack_page; TR(ZKC_VM_DST0_INDEX) = dst_index;
        TR(ZKC_VM_DST0_PERFORMS_MEMORY_ACCESS) = dst0_mem; TR(ZKC_VM_NEW_SP) = new_sp;
        TR(ZKC_VM_SRC0_FROM_MEMORY) = src0_mem.is_pointer;
#pragma unroll
        for (int i = 0; i < 8; i++) TR(ZKC_VM_SRC0_FROM_MEMORY + 1 + i) = src0_mem.value[i];
#pragma unroll 1
        for (int i = 0; i < ZKC_VM_OP_AUX_COLS; i++) TR(aux_base + i) = 0;  // the selected opcode family overwrites its part
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
    uint32_t d0_is_ptr = 0, d1_is_ptr = 0;
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
    } else if (TYPE(ZKC_OP_UMA)) {  // uma.rs:18-1084
        const bool heap_r = VAR(ZKC_VAR_UMA_HEAP_READ), heap_w = VAR(ZKC_VAR_UMA_HEAP_WRITE), aux_r = VAR(ZKC_VAR_UMA_AUX_HEAP_READ),
                   aux_w = VAR(ZKC_VAR_UMA_AUX_HEAP_WRITE), ptr_r = VAR(ZKC_VAR_UMA_FAT_PTR_READ);
        const bool increment = FLAG(ZKC_VM_UMA_INCREMENT_FLAG_IDX);
        const bool access_heap = heap_r || heap_w, access_aux = aux_r || aux_w;
        const bool not_a_ptr = ptr_r && !ra.is_pointer;
        const uint32_t offset = a.v[0], page = a.v[1], start = a.v[2], length = a.v[3];
        const bool skip_legit = !(offset < length) && ptr_r;
        const uint32_t abs_addr = (ptr_r ? start : 0u) + offset;
        const uint64_t inc64 = (uint64_t)offset + 32;
        const uint32_t incremented = (uint32_t)inc64;
        const bool non_addressable = (inc64 >> 32) != 0 || incremented == 0xFFFFFFFFu;
        const bool qp_panic = not_a_ptr || non_addressable, qp_skip = not_a_ptr || skip_legit || non_addressable;
        uint32_t bytes_oob = incremented - length;
        if (qp_skip || incremented < length) bytes_oob = 0;
        bytes_oob &= 31;
        const uint32_t heap_bound = ctx.heap_upper_bound, aux_bound = ctx.aux_heap_upper_bound;
        const uint32_t heap_max = access_heap ? incremented : 0u, aux_max = access_aux ? incremented : 0u;
        const bool heap_uf = heap_max < heap_bound, aux_uf = aux_max < aux_bound;
        uint32_t growth_cost = access_heap ? (heap_uf ? 0u : heap_max - heap_bound) : 0u;
        if (access_aux) growth_cost = aux_uf ? 0u : aux_max - aux_bound;
        const bool top_nz = (a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) != 0;
        const bool exc_oob = (access_heap || access_aux) && (top_nz || non_addressable);
        if (exc_oob) growth_cost = 0xFFFFFFFFu;
        const bool uf = ergs_left < growth_cost;
        const bool set_panic = qp_panic || uf || exc_oob;
        const bool skip_mem = qp_skip || set_panic;
        const bool is_read = heap_r || aux_r || ptr_r, is_write = heap_w || aux_w;
        const uint32_t cell_idx = abs_addr >> 5, unalignment = abs_addr & 31;
        const bool unaligned = unalignment != 0;
        const uint32_t mem_page = access_aux ? aux_heap_page : (access_heap ? heap_page : page);
        const uint32_t b_idx = cell_idx + 1;
        const bool read_a = !skip_mem, read_b = unaligned && !skip_mem;
        zkc_vm_register va = reg_zero(), vb = reg_zero();
        if constexpr (SIM) {
            if (read_a) va = sim_read(*sim, mem_page, cell_idx);
            if (read_b) vb = sim_read(*sim, mem_page, b_idx);
            va.is_pointer = 0; vb.is_pointer = 0;
            for (int i = 0; i < 8; i++) { w.value_a[i] = va.value[i]; w.value_b[i] = vb.value[i]; }
        } else {
            if (read_a) for (int i = 0; i < 8; i++) va.value[i] = w.value_a[i];
            if (read_b) for (int i = 0; i < 8; i++) vb.value[i] = w.value_b[i];
        }
        vm_push<SIM>(1, read_a, d, so, penc, last_memq, ts0, mem_page, cell_idx, 0, va);
        vm_push<SIM>(2, read_b, d, so, penc, last_memq, ts0, mem_page, b_idx, 0, vb);
        // 64-byte big-endian window over cells A, B (:533-560); byte i of a cell is bits of limb 7 - i / 4
        uint8_t bytes[64], written[64];
        for (int i = 0; i < 32; i++) {
            bytes[i] = (uint8_t)(va.value[7 - i / 4] >> (8 * (3 - i % 4)));
            bytes[32 + i] = (uint8_t)(vb.value[7 - i / 4] >> (8 * (3 - i % 4)));
        }
        for (int i = 0; i < 64; i++) written[i] = bytes[i];
        for (int i = 0; i < 32; i++) written[unalignment + i] = (uint8_t)(b.v[7 - i / 4] >> (8 * (3 - i % 4)));
        const uint32_t cleanup = ptr_r ? bytes_oob : 0u;
        U256 rd, wa, wb;
        for (int i = 0; i < 8; i++) { rd.v[i] = 0; wa.v[i] = 0; wb.v[i] = 0; }
        for (int i = 0; i < 32; i++) {
            const uint32_t byte = (uint32_t)i >= 32 - cleanup ? 0u : bytes[unalignment + i];
            rd.v[7 - i / 4] |= byte << (8 * (3 - i % 4));
            wa.v[7 - i / 4] |= (uint32_t)written[i] << (8 * (3 - i % 4));
            wb.v[7 - i / 4] |= (uint32_t)written[32 + i] << (8 * (3 - i % 4));
        }
        const bool exec_write = is_write && !skip_mem, exec_write_b = exec_write && unaligned;
        zkc_vm_register rwa = reg_zero(), rwb = reg_zero();
        for (int i = 0; i < 8; i++) { rwa.value[i] = wa.v[i]; rwb.value[i] = wb.v[i]; }
        vm_push<SIM>(3, exec_write, d, so, penc, last_memq, ts0 + 3, mem_page, cell_idx, 1, rwa);
        vm_push<SIM>(4, exec_write_b, d, so, penc, last_memq, ts0 + 3, mem_page, b_idx, 1, rwb);
        if constexpr (SIM) {
            if (exec_write) sim_write(*sim, mem_page, cell_idx, rwa);
            if (exec_write_b) sim_write(*sim, mem_page, b_idx, rwb);
        }
        const bool w_inc = is_write && increment, no_panic = !set_panic;
        U256 inc_reg = a;
        inc_reg.v[0] = incremented;
        d0 = w_inc ? inc_reg : rd; d0_is_ptr = w_inc ? ra.is_pointer : 0u;
        dst0_reg_only = no_panic && (is_read || w_inc);
        write_dst1 = no_panic && is_read && increment;
        d1 = inc_reg; d1_is_ptr = ra.is_pointer;
        new_pending = set_panic;
        if (access_heap) d.heap_bound = heap_uf ? heap_bound : heap_max;
        if (access_aux) d.aux_bound = aux_uf ? aux_bound : aux_max;
        d.ergs = uf ? 0u : ergs_left - growth_cost;
        if (wr) {
            TR(aux_base + 0) = abs_addr; TR(aux_base + 1) = cell_idx; TR(aux_base + 2) = unalignment; TR(aux_base + 3) = mem_page;
            TR(aux_base + 4) = skip_mem; TR(aux_base + 5) = set_panic; TR(aux_base + 6) = growth_cost; TR(aux_base + 7) = incremented;
            for (int i = 0; i < 8; i++) {
                TR(aux_base + 8 + i) = va.value[i]; TR(aux_base + 16 + i) = vb.value[i];
                TR(aux_base + 24 + i) = exec_write ? wa.v[i] : 0u; TR(aux_base + 32 + i) = exec_write_b ? wb.v[i] : 0u;
            }
        }
    } else if (TYPE(ZKC_OP_LOG)) {  // log.rs:16-671
        const bool st_read = VAR(ZKC_VAR_LOG_STORAGE_READ), st_write = VAR(ZKC_VAR_LOG_STORAGE_WRITE), is_event = VAR(ZKC_VAR_LOG_EVENT),
                   is_l1 = VAR(ZKC_VAR_LOG_TO_L1_MESSAGE), is_precompile = VAR(ZKC_VAR_LOG_PRECOMPILE_CALL);
        uint32_t key[8];
        for (int i = 0; i < 8; i++) key[i] = a.v[i];
        if (is_precompile && key[4] == 0) key[4] = heap_page;
        if (is_precompile && key[5] == 0) key[5] = heap_page;
        const bool write_to_rollup = ctx.this_shard_id == 0 && st_write;
        const bool is_storage = st_read || st_write, revertable = !(st_read || is_precompile);
        const uint32_t aux_byte = (is_storage ? isa->log_aux_bytes[0] : 0u) + (is_event ? isa->log_aux_bytes[1] : 0u) +
                                  (is_l1 ? isa->log_aux_bytes[2] : 0u) + (is_precompile ? isa->log_aux_bytes[3] : 0u);
        const uint32_t is_service = FLAG(ZKC_VM_FIRST_MESSAGE_FLAG_IDX);
        int slot = -1;
        if constexpr (SIM) {
            w.refund = 0;
            if (st_write) { slot = sim_slot(*sim, key); w.refund = sim->storage[slot].written ? isa->initial_storage_write_pubdata_bytes : 0u; }
        }
        const uint32_t refund = w.refund;
        if (refund > isa->initial_storage_write_pubdata_bytes) checks |= ZKC_VM_CHK_LOG_REFUND;
        const uint32_t net_cost = isa->initial_storage_write_pubdata_bytes - refund;
        uint32_t burn = write_to_rollup ? s.ergs_per_pubdata_byte * net_cost : 0u;
        if (is_precompile) burn = b.v[0];
        if (is_l1) burn = s.ergs_per_pubdata_byte * isa->l1_message_pubdata_bytes;
        const bool not_enough = ergs_left < burn;
        const bool execute = !not_enough, execute_rollback = execute && revertable;
        uint32_t read_value[8], written_value[8];
        if constexpr (SIM) {
            for (int i = 0; i < 8; i++) w.value_a[i] = 0;
            if (is_storage && execute) {
                if (slot < 0) slot = sim_slot(*sim, key);
                for (int i = 0; i < 8; i++) w.value_a[i] = sim->storage[slot].value[i];
            }
        }
        for (int i = 0; i < 8; i++) { read_value[i] = is_storage ? w.value_a[i] : 0u; written_value[i] = revertable ? b.v[i] : read_value[i]; }
        uint64_t enc[20];
        vm_log_encode(ctx.this_address, key, read_value, written_value, s.tx_number_in_block, ts0 + 1, aux_byte, ctx.this_shard_id, revertable,
                      is_service, enc);
        uint64_t fin[12], in8[8];
        const uint64_t zero4[4] = {0, 0, 0, 0};
        if (execute) {
            for (int i = 0; i < 8; i++) in8[i] = enc[i];
            vm_job<SIM>(1, d, penc, in8, zero4, VM_CAP_ZERO, VM_CHK_NONE, fin);
            for (int i = 0; i < 8; i++) in8[i] = enc[8 + i];
            vm_job<SIM>(2, d, penc, in8, fin + 8, 1, VM_CHK_NONE, fin);
            for (int i = 0; i < 4; i++) { in8[i] = enc[16 + i]; in8[4 + i] = ctx.log_queue_forward_tail[i]; }
            uint64_t fwd[12];
            vm_job<SIM>(3, d, penc, in8, fin + 8, 2, VM_CHK_NEXT_FWD_TAIL, fwd);
            d.fwd_tail_kind = 1; d.fwd_len++;
            if constexpr (SIM) for (int i = 0; i < 4; i++) so->fwd_tail[i] = fwd[i];
        }
        if constexpr (SIM) {
            if (execute_rollback) {
                sim->ev_kind = 4; sim->ev.kind = 2; sim->ev.slot = -1;
                for (int i = 0; i < 4; i++) { sim->ev.enc16[i] = enc[16 + i]; sim->ev.cap[i] = fin[8 + i]; }
                sim->ev.enc16[3] = 1;
            }
            if (st_write && execute) {
                VmSlot &sl = sim->storage[slot];
                sim->ev.slot = slot; sim->ev.prev_written = sl.written;
                for (int i = 0; i < 8; i++) { sim->ev.prev_value[i] = sl.value[i]; sl.value[i] = b.v[i]; }
                sl.written = 1;
            }
        }
        if (execute_rollback) {
            for (int i = 0; i < 4; i++) { in8[i] = enc[16 + i]; in8[4 + i] = w.rollback[i]; }
            in8[3] = 1;  // update_packing_for_rollback, log_query/mod.rs:52-58
            uint64_t rb[12];
            vm_job<SIM>(4, d, penc, in8, fin + 8, 2, VM_CHK_CUR_RB_HEAD, rb);
            if constexpr (SIM) {
                bool same = true;
                for (int i = 0; i < 4; i++) same &= rb[i] == ctx.reverted_queue_head[i];
                if (!same) checks |= ZKC_VM_CHK_ROLLBACK_QUEUE;
            }
            for (int i = 0; i < 4; i++) d.rb_head[i] = w.rollback[i];
            d.rb_len++;
        }
        if (st_read) for (int i = 0; i < 8; i++) d0.v[i] = read_value[i];
        else d0.v[0] = execute;
        dst0_reg_only = st_read || is_precompile;
        d.ergs = not_enough ? 0u : ergs_left - burn;
        if (wr) {
            for (int i = 0; i < 20; i++) TR(aux_base + i) = enc[i];
            for (int i = 0; i < 8; i++) TR(aux_base + 20 + i) = read_value[i];
            TR(aux_base + 28) = execute; TR(aux_base + 29) = execute_rollback; TR(aux_base + 30) = burn;
        }
    } else if (TYPE(ZKC_OP_NEAR_CALL) || TYPE(ZKC_OP_RET) || (!SIM && TYPE(ZKC_OP_FAR_CALL))) {  // call_ret.rs:24-512
        const bool apply_near = TYPE(ZKC_OP_NEAR_CALL), apply_far = TYPE(ZKC_OP_FAR_CALL), apply_ret = !apply_near && !apply_far;
        bool far_exception = false;
        const uint32_t fwd_byte = (a.v[ZKC_VM_ABI_FORWARDING_MODE_BYTE_IDX / 4] >> (8 * (ZKC_VM_ABI_FORWARDING_MODE_BYTE_IDX % 4))) & 0xFF;
        const bool use_aux = fwd_byte == ZKC_VM_FORWARD_USE_AUX_HEAP, fwd_ptr = fwd_byte == ZKC_VM_FORWARD_FAT_POINTER;
        const bool use_heap = !(use_aux || fwd_ptr);
        uint32_t upper_bound;
        bool non_addressable;
        const VmFatPtr fp = vm_fat_ptr_parse(a, !fwd_ptr, upper_bound, non_addressable);
        const bool generally_invalid = (a.v[0] != 0 && !fwd_ptr) || non_addressable || a.v[3] < a.v[0];
        zkc_vm_context old_entry;
        bool is_panic_out = false, perform_revert = false;
        // the draft context: what create_prestate leaves (pc, sp, ergs updated)
        zkc_vm_context cur_e = ctx;
        cur_e.pc = d.pc; cur_e.sp = new_sp; cur_e.ergs_remaining = ergs_left;
        uint32_t cap_code;
        const uint64_t *cap_sim = nullptr;
        uint64_t prev_sponge[12];
        if (apply_near) {  // near_call.rs:34-184
            cur_e.pc = pc_plus_one;
            nctx = cur_e;
            for (int i = 0; i < 4; i++) { nctx.reverted_queue_tail[i] = w.rollback[i]; nctx.reverted_queue_head[i] = w.rollback[i]; }
            nctx.reverted_queue_segment_len = 0;
            const uint32_t ergs_passed = a.v[0];
            const uint32_t to_pass = ergs_passed == 0 ? ergs_left : ergs_passed;
            const bool uf = ergs_left < to_pass;
            cur_e.ergs_remaining = uf ? 0u : ergs_left - to_pass;
            nctx.ergs_remaining = uf ? ergs_left : to_pass;
            nctx.pc = imm0; nctx.exception_handler_loc = imm1; nctx.is_local_call = 1;
            old_entry = cur_e;
            d.depth = s.context_stack_depth + 1;
            cap_code = VM_CAP_STACK;
            if constexpr (SIM) {
                cap_sim = so->stack + 8;
                if (s.context_stack_depth >= VM_SIM_MAX_DEPTH) sim->overflow = 1;
                else {
                    sim->stack[s.context_stack_depth].context = old_entry;
                    for (int i = 0; i < 12; i++) sim->stack[s.context_stack_depth].previous_sponge_state[i] = s.stack_sponge_state[i];
                }
                sim->ev_kind = 1; sim->ev.kind = 1; sim->ev.slot = -1;
            }
        } else if (apply_far) {  // far_call.rs:268-1098 (circuit side only: the GPU out-of-circuit run has no far calls)
            const bool is_delegated = VAR(ZKC_VAR_FAR_CALL_DELEGATE), is_mimic = VAR(ZKC_VAR_FAR_CALL_MIMIC);
            cur_e.pc = pc_plus_one;
            memset(&nctx, 0, sizeof nctx);
            nctx.heap_upper_bound = isa->new_frame_memory_stipend; nctx.aux_heap_upper_bound = isa->new_frame_memory_stipend;
            const bool is_static_call = FLAG(ZKC_VM_FAR_CALL_STATIC_FLAG_IDX), is_call_shard = FLAG(ZKC_VM_FAR_CALL_SHARD_FLAG_IDX);
#define ABI_BYTE(k) ((a.v[(k) / 4] >> (8 * ((k) % 4))) & 0xFF)
            const uint32_t abi_ergs_passed = a.v[6], abi_shard = ABI_BYTE(ZKC_VM_ABI_SHARD_ID_BYTE_IDX);
            bool abi_constructor = ABI_BYTE(ZKC_VM_ABI_CONSTRUCTOR_CALL_BYTE_IDX) != 0, abi_system = ABI_BYTE(ZKC_VM_ABI_SYSTEM_CALL_BYTE_IDX) != 0;
#undef ABI_BYTE
            const uint32_t caller_shard = cur_e.this_shard_id;
            const uint32_t dest_shard = is_call_shard ? abi_shard : caller_shard;
            const bool target_is_zkporter = dest_shard != 0;
            const bool target_is_kernel = (b.v[0] >> 16) == 0 && (b.v[1] | b.v[2] | b.v[3] | b.v[4]) == 0;
            abi_constructor = abi_constructor && is_kernel; abi_system = abi_system && target_is_kernel;
            const uint32_t new_base_page = s.memory_page_counter;
            d.page_counter = s.memory_page_counter + isa->new_memory_pages_per_far_call;
            // may_be_read_code_hash, :1104-1272
            const bool zkporter_available = gc->zkporter_is_available != 0;
            const bool should_read = !target_is_zkporter || zkporter_available, needs_porter_mask = target_is_zkporter && !zkporter_available;
            uint32_t hash[8], dep_addr[5] = {isa->deployer_system_contract_address_low, 0, 0, 0, 0}, key[8] = {b.v[0], b.v[1], b.v[2], b.v[3], b.v[4], 0, 0, 0};
            for (int i = 0; i < 8; i++) hash[i] = w.value_a[i];
            bool empty = true;
            for (int i = 0; i < 8; i++) empty &= hash[i] == 0;
            if (should_read) {  // construct_hash_relations_code_hash_read, :1274-1411
                uint64_t enc[20], in8[8], f[12];
                const uint64_t zero4[4] = {0, 0, 0, 0};
                vm_log_encode(dep_addr, key, hash, hash, s.tx_number_in_block, ts0 + 1, isa->log_aux_bytes[0], dest_shard, 0, 0, enc);
                for (int i = 0; i < 8; i++) in8[i] = enc[i];
                vm_job<SIM>(5, d, penc, in8, zero4, VM_CAP_ZERO, VM_CHK_NONE, f);
                for (int i = 0; i < 8; i++) in8[i] = enc[8 + i];
                vm_job<SIM>(6, d, penc, in8, f + 8, 5, VM_CHK_NONE, f);
                for (int i = 0; i < 4; i++) { in8[i] = enc[16 + i]; in8[4 + i] = ctx.log_queue_forward_tail[i]; }
                vm_job<SIM>(7, d, penc, in8, f + 8, 6, VM_CHK_NEXT_FWD_TAIL, f);
                d.fwd_tail_kind = 1; d.fwd_len++;
            }
            const bool mask_default_aa = should_read && empty && !target_is_kernel;
            if (mask_default_aa) for (int i = 0; i < 8; i++) hash[i] = gc->default_aa_code_hash[i];
            if (needs_porter_mask) for (int i = 0; i < 8; i++) hash[i] = 0;
            const bool hash_is_trivial = (empty && !mask_default_aa) || needs_porter_mask || !should_read;
            uint32_t target_code_page = hash_is_trivial ? 0u : s.memory_page_counter;
            const uint32_t top = hash[7], version_byte = top >> 24, marker = (top >> 16) & 0xFF;
            const bool normal_marker = marker == 0, constructor_marker = marker == isa->code_hash_yet_constructed_marker;
            const bool code_format_exception = version_byte != isa->code_hash_version_byte || !(normal_marker || constructor_marker);
            const bool can_call_code = (normal_marker && !abi_constructor) || (constructor_marker && abi_constructor);
            uint32_t masked_hash[8];
            for (int i = 0; i < 8; i++) masked_hash[i] = can_call_code ? hash[i] : (target_is_kernel ? 0u : gc->default_aa_code_hash[i]);
            if (can_call_code) masked_hash[7] = (top & 0xFFFF) | (isa->code_hash_at_rest_marker << 16) | (isa->code_hash_version_byte << 24);
            const uint32_t code_len_words = code_format_exception ? 0u : (masked_hash[7] & 0xFFFF);
            const bool exceptions_collapsed = code_format_exception || (!can_call_code && target_is_kernel) || (fwd_ptr && !ra.is_pointer) ||
                                              generally_invalid || non_addressable;
            VmFatPtr p = fwd_ptr ? VmFatPtr{0, fp.page, fp.start + fp.offset, fp.length - fp.offset} : VmFatPtr{0, use_heap ? heap_page : aux_heap_page, fp.start, fp.length};
            if (exceptions_collapsed) p = VmFatPtr{0, 0, 0, 0};
            uint32_t ub = exceptions_collapsed ? 0u : upper_bound;
            if (non_addressable && !fwd_ptr) ub = 0xFFFFFFFFu;
            const uint32_t heap_max = use_heap ? ub : 0u, aux_max = use_aux ? ub : 0u;
            const bool heap_uf = heap_max < cur_e.heap_upper_bound, aux_uf = aux_max < cur_e.aux_heap_upper_bound;
            uint32_t growth_cost = use_heap ? (heap_uf ? 0u : heap_max - cur_e.heap_upper_bound) : 0u;
            if (use_aux) growth_cost = aux_uf ? 0u : aux_max - cur_e.aux_heap_upper_bound;
            const bool growth_uf = ergs_left < growth_cost;
            const uint32_t ergs_after_growth = growth_uf ? 0u : ergs_left - growth_cost;
            if (use_heap && !heap_uf) cur_e.heap_upper_bound = heap_max;
            if (use_aux && !aux_uf) cur_e.aux_heap_upper_bound = aux_max;
            bool exception = exceptions_collapsed || growth_uf;
            bool should_decommit = !exception;
            if (!should_decommit) target_code_page = 0;
            // add_to_decommittment_queue, :1418-1603
            const uint32_t decommit_cost = isa->ergs_per_code_word_decommittment * code_len_words;
            const bool not_enough_for_decommit = ergs_after_growth < decommit_cost;
            should_decommit = should_decommit && !not_enough_for_decommit;
            uint32_t ergs_after_decommit = should_decommit ? ergs_after_growth - decommit_cost : ergs_after_growth;
            const uint32_t suggested_page = w.suggested_page;
            const bool is_first = target_code_page == suggested_page;
            if (should_decommit && !is_first) ergs_after_decommit = ergs_after_growth;
            if (should_decommit) {
                uint64_t e8[8], f[12];
                const uint32_t tsd = ts0 + 1;
                e8[0] = (uint64_t)masked_hash[0] + ((uint64_t)(suggested_page & 0xFFFFFF) << 32);
                e8[1] = (uint64_t)masked_hash[1] + ((uint64_t)(suggested_page >> 24) << 32) + ((uint64_t)(tsd & 0xFFFF) << 40);
                e8[2] = (uint64_t)masked_hash[2] + ((uint64_t)(tsd >> 16) << 32) + ((uint64_t)is_first << 48);
                for (int i = 3; i < 8; i++) e8[i] = masked_hash[i];
                vm_job<SIM>(8, d, penc, e8, nullptr, VM_CAP_DECOMMIT, VM_CHK_NEXT_DECOMMIT, f);
                d.decommit_len = s.code_decommittment_queue_length + 1;
            }
            const uint32_t code_memory_page = should_decommit ? suggested_page : 0u;
            exception = exception || not_enough_for_decommit;
            const uint32_t max_passable = (ergs_after_decommit / 64) * 63, leftover = ergs_after_decommit - max_passable;
            const bool pass_uf = max_passable < abi_ergs_passed;
            cur_e.ergs_remaining = pass_uf ? leftover : leftover + (max_passable - abi_ergs_passed);
            for (int i = 0; i < 4; i++) { nctx.reverted_queue_tail[i] = w.rollback[i]; nctx.reverted_queue_head[i] = w.rollback[i]; }
            nctx.ergs_remaining = pass_uf ? max_passable : abi_ergs_passed; nctx.pc = 0; nctx.exception_handler_loc = imm0;
            nctx.is_static_execution = is_static_call || cur_e.is_static_execution;
            nctx.is_kernel_mode = is_delegated ? cur_e.is_kernel_mode : (uint32_t)target_is_kernel;
            nctx.code_shard_id = dest_shard; nctx.this_shard_id = is_delegated ? caller_shard : dest_shard; nctx.caller_shard_id = caller_shard;
            const zkc_vm_register mimic_reg = regs.get(isa->call_implicit_parameter_reg_idx < ZKC_VM_REGISTERS ? isa->call_implicit_parameter_reg_idx : 0);
            for (int i = 0; i < 5; i++) {
                nctx.code_address[i] = b.v[i];
                nctx.this_address[i] = is_delegated ? cur_e.this_address[i] : b.v[i];
                nctx.caller[i] = is_mimic ? mimic_reg.value[i] : (is_delegated ? cur_e.caller[i] : cur_e.this_address[i]);
            }
            nctx.code_page = code_memory_page; nctx.base_page = new_base_page;
            for (int i = 0; i < 4; i++) nctx.context_u128_value_composite[i] = is_delegated ? cur_e.context_u128_value_composite[i] : s.context_composite_u128[i];
            d.far_call = 1; d.far_call_system = abi_system; d.r2_low = (uint32_t)abi_constructor + 2u * (uint32_t)abi_system;
            d.r1_val = reg_zero(); d.r1_val.is_pointer = 1;
            d.r1_val.value[0] = p.offset; d.r1_val.value[1] = p.page; d.r1_val.value[2] = p.start; d.r1_val.value[3] = p.length;
            far_exception = exception;
            d.set_u128 = 1;
            for (int i = 0; i < 4; i++) d.u128[i] = 0;
            old_entry = cur_e;
            d.depth = s.context_stack_depth + 1;
            cap_code = VM_CAP_STACK;
        } else {  // ret.rs:29-479
            const bool is_ok = VAR(ZKC_VAR_RET_OK), is_revert = VAR(ZKC_VAR_RET_REVERT), is_ret_panic = VAR(ZKC_VAR_RET_PANIC);
            const bool is_local = ctx.is_local_call != 0, is_far_return = !is_local;
            const bool src0_is_ptr = ra.is_pointer && !is_ret_panic;
            const bool is_to_label = FLAG(ZKC_VM_RET_TO_LABEL_FLAG_IDX);
            bool have = false;
            uint32_t cwi = 0;
            if constexpr (SIM) {
                if (s.context_stack_depth >= 1 && s.context_stack_depth - 1 < VM_SIM_MAX_DEPTH) {
                    const zkc_vm_callstack_witness &top = sim->stack[s.context_stack_depth - 1];
                    if (sim->n_cw < sim->cw_cap) { cwi = sim->n_cw; sim->cw_out[sim->n_cw++] = top; w.callstack_index = cwi; have = true; }
                    else sim->overflow = 1;
                    nctx = top.context;
                    for (int i = 0; i < 12; i++) prev_sponge[i] = top.previous_sponge_state[i];
                }
            } else {
                cwi = w.callstack_index;
                have = cwi < n_cw;
                if (have) {
                    nctx = cw[cwi].context;
                    for (int i = 0; i < 12; i++) prev_sponge[i] = cw[cwi].previous_sponge_state[i];
                } else checks |= ZKC_VM_CHK_CALLSTACK;
            }
            if (!have) {
                memset(&nctx, 0, sizeof nctx);
                for (int i = 0; i < 12; i++) prev_sponge[i] = 0;
            }
            d.cw_index = cwi;
            old_entry = nctx;
            const uint32_t popped_seg_len = nctx.reverted_queue_segment_len;
            const bool exc1 = fwd_ptr && !src0_is_ptr && is_far_return;
            const bool exc2 = fwd_ptr && fp.page < cur_e.base_page;
            const bool exceptions_collapsed = exc1 || exc2 || is_ret_panic;
            VmFatPtr p = exceptions_collapsed ? VmFatPtr{0, 0, 0, 0} : fp;
            p = fwd_ptr ? VmFatPtr{0, p.page, p.start + p.offset, p.length - p.offset} : VmFatPtr{0, use_heap ? heap_page : aux_heap_page, p.start, p.length};
            uint32_t ub = exceptions_collapsed ? 0u : upper_bound;
            if (non_addressable && !fwd_ptr) ub = 0xFFFFFFFFu;
            const uint32_t heap_max = use_heap ? ub : 0u, aux_max = use_aux ? ub : 0u;
            const uint32_t heap_growth = heap_max < cur_e.heap_upper_bound ? 0u : heap_max - cur_e.heap_upper_bound;
            const uint32_t aux_growth = aux_max < cur_e.aux_heap_upper_bound ? 0u : aux_max - cur_e.aux_heap_upper_bound;
            uint32_t growth_cost = (use_heap && is_far_return) ? heap_growth : 0u;
            if (use_aux && is_far_return) growth_cost = aux_growth;
            const bool uf = ergs_left < growth_cost;
            uint32_t ergs_after = uf ? 0u : ergs_left - growth_cost;
            if (is_local) ergs_after = ergs_left;
            const bool non_local_panic = (exceptions_collapsed || uf || is_ret_panic) && is_far_return;
            if (non_local_panic) p = VmFatPtr{0, 0, 0, 0};
            const uint64_t ergs_sum = (uint64_t)ergs_after + nctx.ergs_remaining;
            if (ergs_sum >> 32) checks |= ZKC_VM_CHK_CALLSTACK;
            nctx.ergs_remaining = (uint32_t)ergs_sum;
            if (is_local) { nctx.heap_upper_bound = cur_e.heap_upper_bound; nctx.aux_heap_upper_bound = cur_e.aux_heap_upper_bound; }
            const bool should_revert = is_revert || is_ret_panic || non_local_panic;
            perform_revert = should_revert;
            bool head_is_fwd_tail = true, head_is_tail = true;
            for (int i = 0; i < 4; i++) {
                head_is_fwd_tail &= cur_e.reverted_queue_head[i] == cur_e.log_queue_forward_tail[i];
                head_is_tail &= nctx.reverted_queue_head[i] == cur_e.reverted_queue_tail[i];
            }
            const bool ret_ok = is_ok && !non_local_panic;
            if ((should_revert && !head_is_fwd_tail) || (ret_ok && !head_is_tail)) checks |= ZKC_VM_CHK_ROLLBACK_QUEUE;
            if (should_revert) {
                d.fwd_tail_kind = 2;
                for (int i = 0; i < 4; i++) d.fwd_tail[i] = cur_e.reverted_queue_tail[i];
                d.fwd_len = cur_e.log_queue_forward_part_length + cur_e.reverted_queue_segment_len;
            }
            if (ret_ok) {
                for (int i = 0; i < 4; i++) nctx.reverted_queue_head[i] = cur_e.reverted_queue_head[i];
                nctx.reverted_queue_segment_len = popped_seg_len + cur_e.reverted_queue_segment_len;
            }
            const bool use_label = is_to_label && is_local;
            const uint32_t ok_pc = use_label ? imm0 : nctx.pc, eh_pc = use_label ? imm0 : cur_e.exception_handler_loc;
            nctx.pc = should_revert ? eh_pc : ok_pc;
            d.far_ret = is_far_return;
            d.r1_val = reg_zero();
            d.r1_val.is_pointer = 1;
            d.r1_val.value[0] = p.offset; d.r1_val.value[1] = p.page; d.r1_val.value[2] = p.start; d.r1_val.value[3] = p.length;
            is_panic_out = is_ret_panic || non_local_panic;
            if (is_far_return) { d.set_u128 = 1; for (int i = 0; i < 4; i++) d.u128[i] = 0; }
            if (s.context_stack_depth == 0) checks |= ZKC_VM_CHK_CALLSTACK;
            d.depth = s.context_stack_depth - 1;
            cap_code = VM_CAP_CALLSTACK_WITNESS;
            if constexpr (SIM) { cap_sim = prev_sponge + 8; sim->ev_kind = should_revert ? 3 : 2; }
        }
        // the callstack sponge: 4 absorptions of the saved frame's encoding (call_ret.rs:176-284)
        uint64_t enc[32], st12[12], in8[8];
        vm_context_encode(old_entry, enc);
        for (int r = 0; r < 4; r++) {
            for (int i = 0; i < 8; i++) in8[i] = enc[8 * r + i];
            vm_job<SIM>(1 + r, d, penc, in8, r == 0 ? cap_sim : st12 + 8, r == 0 ? cap_code : (uint32_t)r,
                        r == 3 ? (apply_ret ? VM_CHK_CUR_STACK : VM_CHK_NEXT_STACK) : VM_CHK_NONE, st12);
        }
        if constexpr (SIM) {
            if (apply_ret) {
                bool same = true;
                for (int i = 0; i < 12; i++) same &= st12[i] == s.stack_sponge_state[i];
                if (!same) checks |= ZKC_VM_CHK_CALLSTACK;
                for (int i = 0; i < 12; i++) so->stack[i] = prev_sponge[i];
            } else for (int i = 0; i < 12; i++) so->stack[i] = st12[i];
        }
        d.ctx_replaced = apply_near ? 1u : (apply_far ? 3u : 2u);
        set_flags = true; nf0 = is_panic_out && apply_ret; nf1 = 0; nf2 = 0;
        new_pending = far_exception;
        if (wr) {
            uint64_t f[42];
            vm_flatten_record(nctx, f);
            for (int i = 0; i < 42; i++) TR(aux_base + i) = f[i];
            TR(aux_base + 42) = apply_near; TR(aux_base + 43) = apply_ret; TR(aux_base + 44) = is_panic_out;
            TR(aux_base + 45) = perform_revert; TR(aux_base + 46) = apply_far; TR(aux_base + 47) = far_exception;
        }
    }
    // ---- state diffs ---------------------------------------------------------------------------------------------------
    // dst0 / dst1 are dot products of (flag, candidate) pairs (cycle.rs:199-246): zero when no candidate's flag is set
    const bool dst0_any = dst0_mem_capable || dst0_reg_only;
    d.val0.is_pointer = dst0_any ? d0_is_ptr : 0u; d.val1.is_pointer = write_dst1 ? d1_is_ptr : 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) { d.val0.value[i] = dst0_any ? d0.v[i] : 0u; d.val1.value[i] = write_dst1 ? d1.v[i] : 0u; }
    const bool perform_mem_write = dst0_mem && dst0_mem_capable;
    vm_push<SIM>(2, perform_mem_write, d, so, penc, last_memq, ts0 + 3, stack_page, dst_index, 1, d.val0);
    if constexpr (SIM) { if (perform_mem_write) sim_write(*sim, stack_page, dst_index, d.val0); }
    if constexpr (!SIM) { if (last_memq >= 0) d.chk |= (uint64_t)VM_CHK_NEXT_MEMQ << (4 * last_memq); }
    const bool dst0_update_register = dst0_reg_only || (!dst0_mem && dst0_mem_capable);
    if (dst0_update_register) d.idx0 = dst0_r;
    // write_as_dst1 is the decoded selector bit itself (cycle.rs:330, :341-347): an encoded dst1 register is overwritten whatever the
    // gadgets flagged -- with the zero dot product (val1 above) when none did
    d.idx1 = dst1_r;
    d.flags[0] = set_flags ? nf0 : f0; d.flags[1] = set_flags ? nf1 : f1; d.flags[2] = set_flags ? nf2 : f2;
    d.pending = new_pending;
    if (wr) {
        TR(ZKC_VM_DST0) = d.val0.is_pointer; TR(ZKC_VM_DST1) = d.val1.is_pointer;
#pragma unroll
        for (int i = 0; i < 8; i++) { TR(ZKC_VM_DST0 + 1 + i) = d.val0.value[i]; TR(ZKC_VM_DST1 + 1 + i) = d.val1.value[i]; }
        TR(ZKC_VM_PERFORM_DST0_MEMORY_WRITE) = perform_mem_write; TR(ZKC_VM_DST0_UPDATE_REGISTER) = dst0_update_register;
        TR(ZKC_VM_DST1_UPDATE_REGISTER) = write_dst1;
#pragma unroll
        for (int i = 0; i < 3; i++) TR(ZKC_VM_FLAGS_OUT + i) = d.flags[i];
        const bool rep = d.ctx_replaced != 0;
        TR(ZKC_VM_PENDING_EXCEPTION_OUT) = d.pending; TR(ZKC_VM_PC_OUT) = rep ? nctx.pc : d.pc; TR(ZKC_VM_ERGS_OUT) = rep ? nctx.ergs_remaining : d.ergs;
        TR(ZKC_VM_HEAP_BOUND_OUT) = rep ? nctx.heap_upper_bound : d.heap_bound; TR(ZKC_VM_AUX_HEAP_BOUND_OUT) = rep ? nctx.aux_heap_upper_bound : d.aux_bound;
        TR(ZKC_VM_MEMQ_LENGTH_OUT) = d.memq_len; TR(ZKC_VM_DEPTH_OUT) = d.depth;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // the forward tail a log produces is a sponge output: taken from the following snapshot (verified by the sponge launch)
            uint64_t ft = d.fwd_tail_kind == 2 ? d.fwd_tail[i] : ctx.log_queue_forward_tail[i];
            if (d.fwd_tail_kind == 1) {
                if constexpr (SIM) ft = so->fwd_tail[i];
                else ft = next_fwd_tail[i];
            }
            TR(ZKC_VM_FORWARD_TAIL_OUT + i) = ft;
            TR(ZKC_VM_ROLLBACK_HEAD_OUT + i) = rep ? nctx.reverted_queue_head[i] : d.rb_head[i];
        }
        TR(ZKC_VM_FORWARD_TAIL_OUT + 4) = d.fwd_len; TR(ZKC_VM_ROLLBACK_HEAD_OUT + 4) = rep ? nctx.reverted_queue_segment_len : d.rb_len;
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
    if (d->limit == 0) d->s_final = d->s0;  // with cycles, the last cycle's thread writes it (concurrently with this side-stream launch)
}

__device__ __forceinline__ void vm_report(VmDev *d, size_t row, uint32_t checks) {
    if (!checks) return;
    atomicOr(&d->failed_checks, checks);
    atomicMin(&d->first_bad, ((unsigned long long)row << 16) | checks);
}

__device__ __forceinline__ bool reg_equal(const zkc_vm_register &a, const zkc_vm_register &b) {
    bool eq = a.is_pointer == b.is_pointer;
#pragma unroll
    for (int i = 0; i < 8; i++) eq &= a.value[i] == b.value[i];
    return eq;
}
// ExecutionContextRecord equality (the forward-log fields are not part of the record)
__device__ bool vm_record_equal(const zkc_vm_context &c, const zkc_vm_context &e) {
    bool eq = true;
    for (int i = 0; i < 5; i++) eq &= c.this_address[i] == e.this_address[i] && c.caller[i] == e.caller[i] && c.code_address[i] == e.code_address[i];
    eq &= c.code_page == e.code_page && c.base_page == e.base_page && c.heap_upper_bound == e.heap_upper_bound &&
          c.aux_heap_upper_bound == e.aux_heap_upper_bound && c.reverted_queue_segment_len == e.reverted_queue_segment_len &&
          c.pc == e.pc && c.sp == e.sp && c.exception_handler_loc == e.exception_handler_loc && c.ergs_remaining == e.ergs_remaining &&
          c.is_static_execution == e.is_static_execution && c.is_kernel_mode == e.is_kernel_mode && c.this_shard_id == e.this_shard_id &&
          c.caller_shard_id == e.caller_shard_id && c.code_shard_id == e.code_shard_id && c.is_local_call == e.is_local_call;
    for (int i = 0; i < 4; i++)
        eq &= c.reverted_queue_head[i] == e.reverted_queue_head[i] && c.reverted_queue_tail[i] == e.reverted_queue_tail[i] &&
              c.context_u128_value_composite[i] == e.context_u128_value_composite[i];
    return eq;
}

// ---- one thread per cycle (of any instance of the batch) -----------------------------------------------------------------
// The per-cycle inputs live in HBM as COLUMNS (struct of arrays): word w of snapshot i at st[w * st_stride + i], word w of the
// oracle answers of cycle g at wt[w * wt_stride + g].  A warp's 32 consecutive cycles then read every word they need as ONE
// 128-byte line -- the snapshot is 294 lines per warp instead of 32 records 1 176 bytes apart -- and "snapshot i + 1" is the
// neighbouring element of the same lines.  Row-major (array of structs) inputs are transposed once per call
// (vm_rows_to_columns_kernel); the host-buffer path expands its transport stream straight into this layout.
constexpr int VM_WORDS = (int)(sizeof(zkc_vm_state) / 4);
constexpr int VM_WIT_WORDS = (int)(sizeof(zkc_vm_cycle_witness) / 4);
static_assert(sizeof(zkc_vm_state) == 1176 && sizeof(zkc_vm_cycle_witness) == 176, "snapshot / witness layout");
static_assert(VM_WORDS == ZKC_VM_STATE_WORDS && VM_WIT_WORDS == ZKC_VM_WITNESS_WORDS, "header constants");
#define VW(f) ((int)(offsetof(zkc_vm_state, f) / 4))
#define VWC(f) ((int)((offsetof(zkc_vm_state, current_context) + offsetof(zkc_vm_context, f)) / 4))
#define WW(f) ((int)(offsetof(zkc_vm_cycle_witness, f) / 4))
static_assert(offsetof(zkc_vm_context, reverted_queue_head) == offsetof(zkc_vm_context, aux_heap_upper_bound) + 8, "context hole");

struct VmCols {
    const uint32_t *st; size_t st_stride;
    const uint32_t *wt; size_t wt_stride;
    __device__ __forceinline__ uint32_t sw(int w, size_t idx) const { return __ldg(st + (size_t)w * st_stride + idx); }
    __device__ __forceinline__ uint64_t sw64(int w, size_t idx) const {
        return (uint64_t)__ldg(st + (size_t)w * st_stride + idx) | ((uint64_t)__ldg(st + (size_t)(w + 1) * st_stride + idx) << 32);
    }
    __device__ __forceinline__ uint32_t ww(int w, size_t g) const { return __ldg(wt + (size_t)w * wt_stride + g); }
};
// the registers of snapshot idx, fetched on demand (a cycle reads at most three of the fifteen)
struct VmRegsOfColumns {
    const uint32_t *base; size_t stride;  // st + idx
    __device__ __forceinline__ zkc_vm_register get(uint32_t r) const {
        const uint32_t *p = base + (size_t)(VW(registers) + 9 * (int)r) * stride;
        zkc_vm_register v;
        v.is_pointer = __ldg(p);
#pragma unroll
        for (int i = 0; i < 8; i++) v.value[i] = __ldg(p + (size_t)(1 + i) * stride);
        return v;
    }
    __device__ __forceinline__ uint32_t low(uint32_t r) const { return __ldg(base + (size_t)(VW(registers) + 9 * (int)r + 1) * stride); }
};
// the oracle answers of cycle g, fetched on demand: the member names of zkc_vm_cycle_witness
struct VmWitWord { const uint32_t *p; __device__ __forceinline__ operator uint32_t() const { return __ldg(p); } };
struct VmWitArr { const uint32_t *p; size_t stride; __device__ __forceinline__ uint32_t operator[](int i) const { return __ldg(p + (size_t)i * stride); } };
struct VmWitArr64 {
    const uint32_t *p; size_t stride;
    __device__ __forceinline__ uint64_t operator[](int i) const {
        return (uint64_t)__ldg(p + (size_t)(2 * i) * stride) | ((uint64_t)__ldg(p + (size_t)(2 * i + 1) * stride) << 32);
    }
};
struct VmWitOfColumns {
    VmWitArr code_word; VmWitWord src0_is_pointer; VmWitArr src0_value; VmWitWord callstack_index, refund, suggested_page;
    VmWitArr value_a, value_b; VmWitArr64 rollback;
    __device__ __forceinline__ VmWitOfColumns(const uint32_t *b, size_t s)
        : code_word{b + WW(code_word) * s, s}, src0_is_pointer{b + WW(src0_is_pointer) * s}, src0_value{b + WW(src0_value) * s, s},
          callstack_index{b + WW(callstack_index) * s}, refund{b + WW(refund) * s}, suggested_page{b + WW(suggested_page) * s},
          value_a{b + WW(value_a) * s, s}, value_b{b + WW(value_b) * s, s}, rollback{b + WW(rollback) * s, s} {}
};

__device__ __forceinline__ uint32_t reg_word(const zkc_vm_register &r, int i) { return i == 0 ? r.is_pointer : r.value[i - 1]; }

// scratch of one batch: what the cycle launch leaves for the sponge launches
struct VmPushScratch {
    uint32_t *counts;  // [16] (VM_JOB_SLOTS used)
    uint32_t *lists;   // [VM_JOB_SLOTS][rows]: the rows whose slot-k job runs, in no particular order
    uint64_t *meta;    // [rows][2]: job mask | capacity sources << 16 (nibble per slot), checks (nibble per slot)
    uint32_t *link;    // [rows]: which words of the NEXT snapshot the cycle may change (VM_LINK_*), for vm_link_kernel
    uint32_t *exp;     // [VM_EXP_WORDS][rows]: the values the cycle computed for the words it moves (columns, for vm_link_kernel)
    uint64_t *enc, *enc_hi;      // [rows][5][8], [rows][4][8]
    uint64_t *state, *state_hi;  // [rows][5][12], [rows][4][12]: permutation outputs
    __device__ __forceinline__ uint64_t *enc_of(size_t g, int k) const {
        return k < VM_JOB_SLOTS_LO ? enc + (g * VM_JOB_SLOTS_LO + k) * 8 : enc_hi + (g * VM_JOB_SLOTS_HI + (k - VM_JOB_SLOTS_LO)) * 8;
    }
    __device__ __forceinline__ uint64_t *state_of(size_t g, int k) const {
        return k < VM_JOB_SLOTS_LO ? state + (g * VM_JOB_SLOTS_LO + k) * 12 : state_hi + (g * VM_JOB_SLOTS_HI + (k - VM_JOB_SLOTS_LO)) * 12;
    }
    // COMPACT trace layout: every executed job also appends a record (null otherwise)
    zkc_vm_sponge_record *records;
    unsigned long long *n_records;
    unsigned long long records_capacity;
};

// Expected value of word w (>= VW(flags), outside the registers and the sponge-derived states) of the NEXT snapshot, given what
// the cycle produced.  w is a compile-time constant at every call site (unrolled loops): the chain below folds to one case.
#define CW(f) ((int)(offsetof(zkc_vm_context, f) / 4))
__device__ __forceinline__ uint32_t vm_half(uint64_t v, int hi) { return hi ? (uint32_t)(v >> 32) : (uint32_t)v; }
__device__ __forceinline__ uint32_t vm_expected_word(int w, const VmDelta &d, const zkc_vm_context &nctx, const uint64_t (&next_fwd_tail)[4],
                                                     const uint32_t *cur, size_t stride) {
#define KEEP __ldg(cur + (size_t)w * stride)
    if (w >= VW(flags) && w < VW(flags) + 3) return d.flags[w - VW(flags)];
    if (w == VW(timestamp)) return d.timestamp;
    if (w == VW(memory_page_counter)) return d.page_counter;
    if (w == VW(tx_number_in_block)) return KEEP + (d.inc_tx ? 1u : 0u);
    if (w == VW(previous_code_page)) return d.prev_code_page;
    if (w == VW(previous_super_pc)) return d.prev_super_pc;
    if (w == VW(pending_exception)) return d.pending;
    if (w == VW(ergs_per_pubdata_byte)) return d.set_pubdata ? d.pubdata : KEEP;
    if (w == VW(context_stack_depth)) return d.depth;
    if (w == VW(memory_queue_length)) return d.memq_len;
    if (w == VW(code_decommittment_queue_length)) return d.decommit_len;
    if (w >= VW(context_composite_u128) && w < VW(context_composite_u128) + 4) return d.set_u128 ? d.u128[(w - VW(context_composite_u128)) & 3] : KEEP;
    if (w >= VW(current_context)) {
        const int c = w - VW(current_context);
        if (c == CW(log_queue_forward_part_length)) return d.fwd_len;
        if (c >= CW(log_queue_forward_tail)) {  // kind 1: the log's forward sponge (slot 3 / 7) vouches for it
            const int i = (c - CW(log_queue_forward_tail)) & 7;
            return d.fwd_tail_kind == 2 ? vm_half(d.fwd_tail[i >> 1], i & 1) : (d.fwd_tail_kind == 1 ? vm_half(next_fwd_tail[i >> 1], i & 1) : KEEP);
        }
        if (d.ctx_replaced) return reinterpret_cast<const uint32_t *>(&nctx)[c];  // the whole record is the new frame's
        if (c == CW(pc)) return d.pc;
        if (c == CW(sp)) return d.sp;
        if (c == CW(ergs_remaining)) return d.ergs;
        if (c == CW(heap_upper_bound)) return d.heap_bound;
        if (c == CW(aux_heap_upper_bound)) return d.aux_bound;
        if (c == CW(reverted_queue_segment_len)) return d.rb_len;
        if (c >= CW(reverted_queue_head) && c < CW(reverted_queue_head) + 8) {
            const int i = (c - CW(reverted_queue_head)) & 7;
            return vm_half(d.rb_head[i >> 1], i & 1);
        }
    }
    return KEEP;
#undef KEEP
}

// ---- the snapshot link, split in two -----------------------------------------------------------------------------------------
// Words of the current context no ordinary cycle moves (only a near / far call or a ret replaces them, with the whole record)
__host__ __device__ constexpr bool vm_link_const_context_word(int w) {
    const int c = w - VW(current_context);
    return (c >= CW(this_address) && c < CW(heap_upper_bound)) ||                                   // addresses, code_page, base_page
           (c >= CW(reverted_queue_tail) && c < CW(reverted_queue_tail) + 8) || c == CW(exception_handler_loc) ||
           (c >= CW(is_static_execution) && c <= CW(is_local_call));                              // mode flags, shard ids, u128, is_local_call
}
// the words an ordinary cycle computes: previous_code_word, the scalars, the moving context fields
__host__ __device__ constexpr bool vm_link_dyn_word(int w) {
    if (w < VW(registers)) return true;
    if (w < VW(flags) || w >= VW(stack_sponge_state)) return false;
    if (w >= VW(_pad) && w < VW(current_context)) return false;  // padding is not state
    if (w == VWC(aux_heap_upper_bound) + 1) return false;        // alignment hole in front of reverted_queue_head
    return !(w >= VW(current_context) && vm_link_const_context_word(w));
}
__host__ __device__ constexpr int vm_exp_slot(int w) {
    int n = 0;
    for (int v = 0; v < w; v++) n += vm_link_dyn_word(v) ? 1 : 0;
    return n;
}
constexpr int VM_EXP_DYN = vm_exp_slot(VM_WORDS), VM_EXP_WORDS = VM_EXP_DYN + 18;  // + dst0 / dst1 register values
// one compile-time step of the word loops: W is a template argument, so slot numbers and word classes are constants of the code
template <int W, typename F>
__device__ __forceinline__ void vm_dyn_word_step(F &f) {
    if constexpr (vm_link_dyn_word(W)) f(std::integral_constant<int, W>{}, std::integral_constant<int, vm_exp_slot(W)>{});
}
template <typename F, int... W>
__device__ __forceinline__ void vm_for_dyn_words_impl(F &f, std::integer_sequence<int, W...>) { (vm_dyn_word_step<W>(f), ...); }
template <typename F>
__device__ __forceinline__ void vm_for_dyn_words(F &&f) { vm_for_dyn_words_impl(f, std::make_integer_sequence<int, VM_WORDS>{}); }
enum : uint32_t { VM_LINK_ALL_REGISTERS = 1u << 8, VM_LINK_MEMQ = 1u << 9, VM_LINK_STACK = 1u << 10, VM_LINK_DECOMMIT = 1u << 11, VM_LINK_CONTEXT = 1u << 12 };

// Carry-over half of the link check: every word group the cycle did NOT declare as changing (link[g]: dst0 / dst1 register
// numbers, "all registers", the three sponge-derived states, "context replaced") must be equal in snapshots row and row + 1.
// A pure stream over the state columns: lane i reads element i and i + 1 of each word's column (coalesced; the second an L1
// hit), 243 of the 294 words; the words a cycle computes are compared by vm_cycles_kernel itself.
__global__ void __launch_bounds__(256, 4)
vm_link_kernel(VmDev *devs, VmCols cols, const uint32_t *__restrict__ link, const uint32_t *__restrict__ exp, size_t limit, size_t n_instances,
               size_t row0, size_t row_count) {
    const size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= row_count * n_instances) return;
    const size_t inst = l / row_count, row = row0 + (l - inst * row_count);
    const size_t g = inst * limit + row, idx = inst * (limit + 1) + row;
    const uint32_t m = __ldg(link + g);
    const uint32_t *cur = cols.st + idx;
    const size_t stride = cols.st_stride;
#define DIFF(w) (__ldg(cur + (size_t)(w) * stride) ^ __ldg(cur + (size_t)(w) * stride + 1))
#define EXP(slot) __ldg(exp + (size_t)(slot) * total + g)
    const size_t total = limit * n_instances;
    bool bad = false;
#if VM_LINK_EXP
    {   // the words the cycle computed: snapshot row + 1 holds exactly those values
        uint32_t dd = 0;
        vm_for_dyn_words([&](auto w, auto slot) {
            dd |= __ldg(cur + (size_t)w.value * stride + 1) ^ EXP(slot.value);
            if (slot.value % 12 == 11) asm volatile("" ::: "memory");  // batches of 24 loads
        });
        bad |= dd != 0;
    }
    if (!(m & VM_LINK_ALL_REGISTERS)) {  // registers: dst1 / dst0 (dst1 is applied last) take the cycle's values, the others carry over
        const uint32_t idx0 = m & 15, idx1 = (m >> 4) & 15;
        uint32_t v0[9], v1[9];
#pragma unroll
        for (int i = 0; i < 9; i++) { v0[i] = EXP(VM_EXP_DYN + i); v1[i] = EXP(VM_EXP_DYN + 9 + i); }
        uint32_t dr = 0;
#pragma unroll 1
        for (int r = 0; r < ZKC_VM_REGISTERS; r++) {  // 18 independent loads in flight per thread and iteration
            const bool is0 = (uint32_t)(r + 1) == idx0, is1 = (uint32_t)(r + 1) == idx1;
#pragma unroll
            for (int i = 0; i < 9; i++) {
                const uint32_t c = __ldg(cur + (size_t)(VW(registers) + 9 * r + i) * stride), n = __ldg(cur + (size_t)(VW(registers) + 9 * r + i) * stride + 1);
                dr |= n ^ (is1 ? v1[i] : (is0 ? v0[i] : c));
            }
        }
        bad |= dr != 0;
    }
#else
    {
        const uint32_t idx0 = m & 15, idx1 = (m >> 4) & 15;
        uint32_t moved = 0;  // bit r: register r differs
#pragma unroll 1
        for (int r = 0; r < ZKC_VM_REGISTERS; r++) {  // 18 independent loads in flight per thread and iteration
            uint32_t dr = 0;
#pragma unroll
            for (int i = 0; i < 9; i++) dr |= DIFF(VW(registers) + 9 * r + i);
            moved |= (dr != 0 ? 1u : 0u) << r;
        }
        const uint32_t allowed = (m & VM_LINK_ALL_REGISTERS) ? 0x7FFFu : ((idx0 ? 1u << (idx0 - 1) : 0u) | (idx1 ? 1u << (idx1 - 1) : 0u));
        bad |= (moved & ~allowed) != 0;
    }
#endif
    {
        uint32_t dc = 0;
#pragma unroll
        for (int w = VW(current_context); w < VW(stack_sponge_state); w++) {
            if (vm_link_const_context_word(w)) dc |= DIFF(w);
            if (w % 12 == 11) asm volatile("" ::: "memory");  // batches of <= 24 loads: keeps the register count of a stream kernel
        }
        bad |= dc != 0 && !(m & VM_LINK_CONTEXT);
    }
    {
        uint32_t dm = 0, ds = 0, dd = 0;
#pragma unroll 1
        for (int i0 = 0; i0 < 24; i0 += 4) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                dm |= DIFF(VW(memory_queue_state) + i0 + i); ds |= DIFF(VW(stack_sponge_state) + i0 + i); dd |= DIFF(VW(code_decommittment_queue_state) + i0 + i);
            }
        }
        bad |= (dm != 0 && !(m & VM_LINK_MEMQ)) || (ds != 0 && !(m & VM_LINK_STACK)) || (dd != 0 && !(m & VM_LINK_DECOMMIT));
    }
#undef DIFF
#undef EXP
    // attribute the broken link to the cycle that would consume the wrong snapshot (what a sequential run sees)
    if (bad) vm_report(devs + inst, row + 1 < limit ? row + 1 : row, ZKC_VM_CHK_SNAPSHOT);
}

// Every thread evaluates its cycle from snapshot `row` and checks that snapshot `row + 1` is the result: the words a cycle
// can change are compared with what the cycle produced, every other word must carry over (the neighbouring element of the
// same column: an L1 hit for 31 of the 32 lanes).  The three sponge-derived states are vouched for by the cycle's sponge jobs
// when it has one (vm_sponge_kernel compares), and must not move otherwise.  The Poseidon2 relations are deferred: the cycle
// only emits their jobs.
#ifndef VM_CYCLES_MIN_BLOCKS
#define VM_CYCLES_MIN_BLOCKS 2
#endif
#ifndef VM_LINK_EXP
#define VM_LINK_EXP 0  // 1: the cycle kernel hands its computed words to vm_link_kernel as columns instead of comparing them itself (measured slower)
#endif
// The words every cycle reads -- previous_code_word (8) and the scalars + current context (79) -- are staged per warp by TMA:
// two tensor tiles [words x 32 consecutive snapshots] of the state columns land in shared memory (cp.async.bulk.tensor.2d,
// completion on the warp's mbarrier), so the ~87 field reads of a cycle are shared-memory loads at their use sites instead of
// global loads the compiler re-issues inside every divergent branch.  Registers of the snapshot, oracle answers and the rare
// whole-record reads stay global loads.
__device__ __forceinline__ uint32_t vm_smem_addr(const void *p);
constexpr int VM_TILE_A = VW(registers), VM_TILE_B = VW(stack_sponge_state) - VW(flags), VM_TILE_WORDS = VM_TILE_A + VM_TILE_B;
struct alignas(64) VmTmaps { CUtensorMap a, b; };  // boxes [32 x VM_TILE_A] and [32 x VM_TILE_B] over state_words [VM_WORDS][stride]
__device__ __forceinline__ int vm_tile_slot(int w) { return w < VW(registers) ? w : VM_TILE_A + (w - VW(flags)); }

// Two instantiations, selected per call: VM_USE_TMA = false reads the state words with plain (non-coherent) global loads the
// compiler places at their use sites -- measured 3 % FASTER on B200 (1.42 vs 1.46 ms per 2^20 cycles, profiles/README.md) because
// the column layout already makes every such load one full 128-byte line per warp and L1 holds the tile; the TMA variant
// (ZKC_VM_TMA=1) is kept for layouts / sizes where the tile does not stay in L1.
template <bool VM_USE_TMA>
__global__ void __launch_bounds__(128, VM_CYCLES_MIN_BLOCKS)
vm_cycles_kernel(VmDev *devs, const zkc_vm_isa *__restrict__ isa, VmCols cols,
                 const zkc_vm_callstack_witness *__restrict__ cws, uint32_t n_cw,
                 uint64_t *__restrict__ trace, size_t limit, size_t n_instances, size_t row0, size_t row_count, VmPushScratch ps,
                 int ncols, int aux_base, const __grid_constant__ VmTmaps tmaps, int use_tma) {
    __shared__ alignas(128) uint32_t vm_tile[VM_USE_TMA ? 4 : 1][VM_USE_TMA ? VM_TILE_WORDS * 32 : 32];
    __shared__ alignas(8) unsigned long long vm_bar[4];
    // this launch covers rows [row0, row0 + row_count) of every instance (one chunk of the pipelined host path, or all)
    const size_t total = limit * n_instances;
    const size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = l < row_count * n_instances;
    const unsigned lane = threadIdx.x & 31;
    const size_t inst = valid ? l / row_count : 0, row = valid ? row0 + (l - inst * row_count) : 0;
    const size_t g = inst * limit + row;
    const size_t idx = inst * (limit + 1) + row;
    VmDev *dev = devs + inst;
    uint32_t checks = 0, jmask = 0;
    // ---- TMA: the warp's tile, when its 32 cycles are 32 consecutive snapshots of one instance ------------------------------
    const int wib = VM_USE_TMA ? threadIdx.x >> 5 : 0;
    const uint32_t *tile = vm_tile[wib] + lane;  // word slot k of this lane's snapshot: tile[k * 32]
    bool staged = false;
    if (VM_USE_TMA && use_tma) {
        const unsigned long long idx0 = __shfl_sync(0xffffffffu, (unsigned long long)idx, 0);
        const bool lane0_valid = __shfl_sync(0xffffffffu, (int)valid, 0) != 0;
        staged = lane0_valid && (idx0 & 3) == 0 &&  // the tile's first element must sit on a 16-byte boundary of its column
                 __all_sync(0xffffffffu, !valid || (unsigned long long)idx == idx0 + lane);
        if (staged) {
            const uint32_t bar_a = vm_smem_addr(&vm_bar[wib]), tile_a = vm_smem_addr(vm_tile[wib]);
            if (lane == 0) {
                constexpr uint32_t BYTES = VM_TILE_WORDS * 32u * 4u;
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(BYTES) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(tile_a), "l"(reinterpret_cast<uint64_t>(&tmaps.a)), "r"(bar_a), "r"((int)idx0), "r"(0) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(tile_a + VM_TILE_A * 128u), "l"(reinterpret_cast<uint64_t>(&tmaps.b)), "r"(bar_a), "r"((int)idx0), "r"(VW(flags)) : "memory");
            }
            __syncwarp();
            uint32_t done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a) : "memory");
        }
    }
    if (valid) {
        const uint32_t *cur = cols.st + idx;  // word w of this snapshot: cur[w * stride]; of the next one: cur[w * stride + 1]
        const size_t stride = cols.st_stride;
#define CUR(w) __ldg(cur + (size_t)(w) * stride)
#define NXT(w) __ldg(cur + (size_t)(w) * stride + 1)
        // ---- the words a cycle reads outside the registers: scalars + the current context ---------------------------------
        zkc_vm_state s;
        uint32_t *sw = reinterpret_cast<uint32_t *>(&s);
        if constexpr (VM_USE_TMA) {
            // one source for every warp: a warp without a TMA tile (it straddles two instances, or its tile is not 16-byte aligned)
            // fills its lanes' slots of the shared tile with plain loads; the field reads below are then shared-memory loads the
            // compiler is free to place at their use sites
            if (!staged) {
                uint32_t *mine = vm_tile[wib] + lane;
#pragma unroll
                for (int w = 0; w < VW(registers); w++) mine[vm_tile_slot(w) * 32] = CUR(w);
#pragma unroll
                for (int w = VW(flags); w < VW(stack_sponge_state); w++) mine[vm_tile_slot(w) * 32] = CUR(w);
            }
#pragma unroll
            for (int w = 0; w < VW(registers); w++) sw[w] = tile[vm_tile_slot(w) * 32];
#pragma unroll
            for (int w = VW(flags); w < VW(stack_sponge_state); w++) sw[w] = tile[vm_tile_slot(w) * 32];
        } else {
#pragma unroll
            for (int w = 0; w < VW(registers); w++) sw[w] = CUR(w);
#pragma unroll
            for (int w = VW(flags); w < VW(stack_sponge_state); w++) sw[w] = CUR(w);
        }
        uint64_t next_fwd_tail[4];
#pragma unroll
        for (int i = 0; i < 4; i++) next_fwd_tail[i] = (uint64_t)NXT(VWC(log_queue_forward_tail) + 2 * i) | ((uint64_t)NXT(VWC(log_queue_forward_tail) + 2 * i + 1) << 32);
        VmDelta d;
        zkc_vm_context nctx;
        d.penc_hi = ps.enc_hi + g * (VM_JOB_SLOTS_HI * 8);
        const VmRegsOfColumns regs{cur, stride};
        VmWitOfColumns wit(cols.wt + g, cols.wt_stride);
        checks |= vm_cycle_dev<false>(isa, &dev->io, s, regs, d, nctx, wit, cws + inst * (size_t)n_cw, n_cw, nullptr, nullptr,
                                      ps.enc + g * (VM_JOB_SLOTS_LO * 8), next_fwd_tail,
                                      trace ? trace + inst * (size_t)ncols * limit : nullptr, limit, row, aux_base);
        jmask = d.job_mask;
        ps.meta[g * 2] = (uint64_t)jmask | (d.cap_from << 16); ps.meta[g * 2 + 1] = d.chk;
        // ---- is snapshot row + 1 what this cycle produces? -----------------------------------------------------------------
        // This thread compares the words its cycle CHANGES with what it computed; that every other word carries over unchanged
        // is checked by vm_link_kernel (a stream over the columns) from the mask written here.
        bool bad = false;
#if VM_LINK_EXP
        // (1) scalars + the context fields an ordinary cycle moves: the computed values go to vm_link_kernel as columns; the rest of
        // the context record only when the callstack moves (rare: compared here)
        {
            vm_for_dyn_words([&](auto w, auto slot) {
                uint32_t v;
                if constexpr (w.value < VW(registers)) v = d.cw[w.value];
                else v = vm_expected_word(w.value, d, nctx, next_fwd_tail, cur, stride);
                ps.exp[(size_t)slot.value * total + g] = v;
            });
            if (d.ctx_replaced) {
                uint32_t acc = 0;
#pragma unroll
                for (int w = VW(current_context); w < VW(stack_sponge_state); w++)
                    if (vm_link_const_context_word(w)) acc |= NXT(w) ^ reinterpret_cast<const uint32_t *>(&nctx)[w - VW(current_context)];
                bad |= acc != 0;
            }
        }
        // (2) registers: the values of dst0 / dst1 go to vm_link_kernel; a far call / far return rewrites all of them (compared here)
#pragma unroll
        for (int i = 0; i < 9; i++) {
            ps.exp[(size_t)(VM_EXP_DYN + i) * total + g] = reg_word(d.val0, i);
            ps.exp[(size_t)(VM_EXP_DYN + 9 + i) * total + g] = reg_word(d.val1, i);
        }
        if (d.far_ret || d.far_call) {
#pragma unroll 1
            for (int r = 0; r < ZKC_VM_REGISTERS; r++) {
                zkc_vm_register want = reg_zero();
                if (d.far_ret) { if (r == 0) want = d.r1_val; }
                else want = vm_far_call_register(isa, d, r, regs.get((uint32_t)r));
                if (d.idx1 == (uint32_t)(r + 1)) want = d.val1;  // the dst1 select comes after the specific updates / zeroing, cycle.rs:415-433
                for (int i = 0; i < 9; i++) bad |= NXT(VW(registers) + 9 * r + i) != reg_word(want, i);
            }
        }
#else
        // (1) scalars + the context fields an ordinary cycle moves; the whole record when the callstack moves
        {
            uint32_t acc = 0;
#pragma unroll
            for (int w = 0; w < VW(registers); w++) acc |= NXT(w) ^ d.cw[w];
#pragma unroll
            for (int w = VW(flags); w < VW(stack_sponge_state); w++) {
                if (w >= VW(_pad) && w < VW(current_context)) continue;  // padding is not state
                if (w == VWC(aux_heap_upper_bound) + 1) continue;        // alignment hole in front of reverted_queue_head
                if (vm_link_const_context_word(w)) continue;
                acc |= NXT(w) ^ vm_expected_word(w, d, nctx, next_fwd_tail, cur, stride);
            }
            if (d.ctx_replaced) {
#pragma unroll
                for (int w = VW(current_context); w < VW(stack_sponge_state); w++)
                    if (vm_link_const_context_word(w)) acc |= NXT(w) ^ reinterpret_cast<const uint32_t *>(&nctx)[w - VW(current_context)];
            }
            bad |= acc != 0;
        }
        // (2) registers: dst0 / dst1 (dst1 is applied last) hold the values the cycle produced; a far call / far return rewrites all
        if (!d.far_ret && !d.far_call) {
            uint32_t acc = 0;
            if (d.idx1) {
                const int base = VW(registers) + 9 * ((int)d.idx1 - 1);
#pragma unroll
                for (int i = 0; i < 9; i++) acc |= NXT(base + i) ^ reg_word(d.val1, i);
            }
            if (d.idx0 && d.idx0 != d.idx1) {
                const int base = VW(registers) + 9 * ((int)d.idx0 - 1);
#pragma unroll
                for (int i = 0; i < 9; i++) acc |= NXT(base + i) ^ reg_word(d.val0, i);
            }
            bad |= acc != 0;
        } else {
#pragma unroll 1
            for (int r = 0; r < ZKC_VM_REGISTERS; r++) {
                zkc_vm_register want = reg_zero();
                if (d.far_ret) { if (r == 0) want = d.r1_val; }
                else want = vm_far_call_register(isa, d, r, regs.get((uint32_t)r));
                if (d.idx1 == (uint32_t)(r + 1)) want = d.val1;  // the dst1 select comes after the specific updates / zeroing, cycle.rs:415-433
                for (int i = 0; i < 9; i++) bad |= NXT(VW(registers) + 9 * r + i) != reg_word(want, i);
            }
        }
#endif
        // (3) the sponge-derived states: vouched for by the cycle's last job on them (vm_sponge_kernel), else unchanged (vm_link_kernel)
        {
            bool memq_job = false, stack_job = false, decommit_job = false;
#pragma unroll
            for (int k = 0; k < VM_JOB_SLOTS; k++) {
                const uint32_t c = (uint32_t)(d.chk >> (4 * k)) & 15;
                memq_job |= c == VM_CHK_NEXT_MEMQ; stack_job |= c == VM_CHK_NEXT_STACK; decommit_job |= c == VM_CHK_NEXT_DECOMMIT;
            }
            if (d.ctx_replaced == 2) {  // ret: the stack state below the popped frame is the witness'
                const uint64_t *p = cws[inst * (size_t)n_cw + (d.cw_index < n_cw ? d.cw_index : 0)].previous_sponge_state;
                for (int i = 0; i < 12; i++) {
                    const uint64_t n = (uint64_t)NXT(VW(stack_sponge_state) + 2 * i) | ((uint64_t)NXT(VW(stack_sponge_state) + 2 * i + 1) << 32);
                    bad |= n != (d.cw_index < n_cw ? p[i] : 0ull);
                }
                stack_job = true;
            }
            ps.link[g] = d.idx0 | (d.idx1 << 4) | ((d.far_ret || d.far_call) ? VM_LINK_ALL_REGISTERS : 0u) | (memq_job ? VM_LINK_MEMQ : 0u) |
                         (stack_job ? VM_LINK_STACK : 0u) | (decommit_job ? VM_LINK_DECOMMIT : 0u) | (d.ctx_replaced ? VM_LINK_CONTEXT : 0u);
        }
        if (bad) {
            // attribute the broken link to the cycle that would consume the wrong snapshot (what a sequential run sees)
            if (row + 1 < limit) vm_report(dev, row + 1, ZKC_VM_CHK_SNAPSHOT);
            else checks |= ZKC_VM_CHK_SNAPSHOT;
        }
        if (row + 1 == limit) {  // the state the circuit ends in, as computed (its sponge-derived parts: vm_sponge_kernel)
            uint32_t *dst = reinterpret_cast<uint32_t *>(&dev->s_final);
#pragma unroll 1
            for (int w = 0; w < VM_WORDS; w++) dst[w] = CUR(w);
            vm_apply_delta(dev->s_final, d, nctx, isa);
            if (d.ctx_replaced == 2 && d.cw_index < n_cw)
                for (int i = 0; i < 12; i++) dev->s_final.stack_sponge_state[i] = cws[inst * (size_t)n_cw + d.cw_index].previous_sponge_state[i];
        }
        vm_report(dev, row, checks);
#undef CUR
#undef NXT
    }
    // ---- rows whose slot-k job runs, for the dense sponge launches ---------------------------------------------------------
#pragma unroll
    for (int k = 0; k <= VM_JOB_SLOTS_LO; k++) {  // slots 0..4 + 5 (a far call: its thread also runs 6, 7, 8)
        const bool mine = k < VM_JOB_SLOTS_LO ? (jmask >> k) & 1 : (jmask >> VM_JOB_SLOTS_LO) != 0;
        const unsigned b = __ballot_sync(0xffffffffu, mine);
        if (!b) continue;
        const int leader = __ffs(b) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) base = atomicAdd(ps.counts + k, (uint32_t)__popc(b));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (mine) ps.lists[(size_t)k * total + base + __popc(b & ((1u << lane) - 1))] = (uint32_t)g;
    }
}

// ---- row-major (array of structs) inputs -> columns: one warp per tile of 32 records, through shared memory ---------------
// records [n_inst][per_inst] of WORDS 32-bit words; rows [r0, r0 + cnt) of every instance.  The tile is CONTIGUOUS in the
// row-major input (32 x WORDS words = 37 632 bytes of snapshots): one elected lane brings it in with a single bulk
// asynchronous copy (TMA, cp.async.bulk global -> shared, completion on an mbarrier) -- 37 KB in flight per warp without a
// register or an instruction per word -- and the warp then writes one 128-byte line per word.  Tiles that are not 16-byte
// aligned (odd record index: 1 176 = 8 x 147) or not full take coalesced loads instead.
__device__ __forceinline__ uint32_t vm_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int WORDS>
__global__ void __launch_bounds__(32)
vm_rows_to_columns_kernel(const uint32_t *__restrict__ rows, uint32_t *__restrict__ cols, size_t stride, size_t per_inst, size_t n_inst,
                          size_t r0, size_t cnt) {
    __shared__ alignas(128) uint32_t tile[32 * WORDS];
    __shared__ alignas(8) unsigned long long bar;
    const int lane = threadIdx.x;
    const size_t tiles_per_inst = (cnt + 31) / 32;
    const size_t t = blockIdx.x;
    const size_t inst = t / tiles_per_inst, first = r0 + (t - inst * tiles_per_inst) * 32;
    const int n = (int)min((size_t)32, r0 + cnt - first);
    const size_t base = inst * per_inst + first;
    const uint32_t *src = rows + base * WORDS;
    constexpr uint32_t BYTES = 32u * WORDS * 4u;
    if (n == 32 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const uint32_t bar_a = vm_smem_addr(&bar), tile_a = vm_smem_addr(tile);
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tile_a), "l"(src), "r"(BYTES), "r"(bar_a) : "memory");
        }
        __syncwarp();
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a) : "memory");
    } else {
        for (int j = lane; j < n * WORDS; j += 32) tile[j] = __ldg(src + j);
        __syncwarp();
    }
    if (lane < n) {
        uint32_t *dst = cols + base + lane;
#pragma unroll 6
        for (int w = 0; w < WORDS; w++) dst[(size_t)w * stride] = tile[lane * WORDS + w];  // 2-way bank conflict at most (WORDS = 6, 12 mod 32)
    }
}

// first and final snapshot of every instance as records: ends[2 * inst], ends[2 * inst + 1] (what the closed-form
// commitments hash; vm_finalize_kernel).  One warp per record.
__global__ void vm_gather_ends_kernel(VmCols cols, zkc_vm_state *__restrict__ ends, size_t limit, size_t n_instances) {
    const size_t wi = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wi >= 2 * n_instances) return;
    const size_t idx = (wi >> 1) * (limit + 1) + ((wi & 1) ? limit : 0);
    uint32_t *dst = reinterpret_cast<uint32_t *>(ends + wi);
    for (int w = lane; w < VM_WORDS; w += 32) dst[w] = cols.sw(w, idx);
}

// slot k of every cycle that has one: out = P(enc || capacity), one thread per job, all lanes busy.  The capacity is a
// previous job's output of the same cycle, zeros, or a queue state of the snapshot / the callstack witness; the output
// of the last job of a chain must land on the next snapshot (or, for the joins the circuit enforces, on the current one).
// One job: slot k of cycle g.  `flags` (persistent launch): per-job completion flags -- a job whose capacity is another
// job's output waits for it, and publishes its own output when done.
__device__ __forceinline__ void vm_sponge_job(VmDev *devs, const VmCols &cols,
                                              const zkc_vm_callstack_witness *__restrict__ cws, uint32_t n_cw, const VmPushScratch &ps, int k, size_t g,
                                              size_t limit, uint32_t *flags) {
    const size_t inst = g / limit, row = g - inst * limit, idx = inst * (limit + 1) + row;
    const uint32_t cap_from = (uint32_t)(ps.meta[g * 2] >> (16 + 4 * k)) & 15, chk = (uint32_t)(ps.meta[g * 2 + 1] >> (4 * k)) & 15;
    uint64_t q[12];
#pragma unroll
    for (int j = 0; j < 8; j++) q[j] = ps.enc_of(g, k)[j];
    if (flags && cap_from < VM_JOB_SLOTS) {  // acquire the producer's output
        const uint32_t *f = flags + g * VM_JOB_SLOTS + cap_from;
        uint32_t v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v) break;
            __nanosleep(100);
        }
    }
    if (flags) __syncwarp(__activemask());  // lanes that had to wait rejoin before the permutation
    if (cap_from == VM_CAP_ZERO) {
#pragma unroll
        for (int j = 8; j < 12; j++) q[j] = 0;
    } else if (cap_from < VM_JOB_SLOTS) {
        const uint64_t *from = ps.state_of(g, (int)cap_from);
#pragma unroll
        for (int j = 8; j < 12; j++) q[j] = from[j];
    } else if (cap_from == VM_CAP_CALLSTACK_WITNESS) {
        // an index outside the table was already reported by the cycle kernel (ZKC_VM_CHK_CALLSTACK);
        // the job then runs from the all-zero capacity like the oracle does
        const uint32_t cwi = cols.ww(WW(callstack_index), g);
        const uint64_t *from = cwi < n_cw ? cws[inst * (size_t)n_cw + cwi].previous_sponge_state : nullptr;
#pragma unroll
        for (int j = 8; j < 12; j++) q[j] = from ? from[j] : 0ull;
    } else {
        const int w0 = cap_from == VM_CAP_MEMQ ? VW(memory_queue_state) : (cap_from == VM_CAP_STACK ? VW(stack_sponge_state) : VW(code_decommittment_queue_state));
#pragma unroll
        for (int j = 8; j < 12; j++) q[j] = cols.sw64(w0 + 2 * j, idx);
    }
    poseidon2_permute(q);
    uint64_t *to = ps.state_of(g, k);
#pragma unroll
    for (int j = 0; j < 12; j++) to[j] = q[j];
    if (flags) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + g * VM_JOB_SLOTS + k), "r"(1u) : "memory");
    if (ps.records) {  // one allocation per warp
        const unsigned act = __activemask();
        const int leader = __ffs(act) - 1;
        unsigned long long base = 0;
        if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(ps.n_records, (unsigned long long)__popc(act));
        base = __shfl_sync(act, base, leader);
        const unsigned long long pos = base + __popc(act & ((1u << (threadIdx.x & 31)) - 1));
        if (pos < ps.records_capacity) {
            zkc_vm_sponge_record &r = ps.records[pos];
            r.row = (uint32_t)g; r.slot = (uint32_t)k;
#pragma unroll
            for (int j = 0; j < 12; j++) r.out[j] = q[j];
        }
    }
    if (chk == VM_CHK_NONE) return;
    VmDev *dev = devs + inst;
    const bool last_row = row + 1 == limit;
    int w0, n = 12;
    size_t at = idx + 1;
    if (chk == VM_CHK_NEXT_MEMQ) w0 = VW(memory_queue_state);
    else if (chk == VM_CHK_NEXT_STACK) w0 = VW(stack_sponge_state);
    else if (chk == VM_CHK_CUR_STACK) { w0 = VW(stack_sponge_state); at = idx; }
    else if (chk == VM_CHK_NEXT_FWD_TAIL) { w0 = VWC(log_queue_forward_tail); n = 4; }
    else if (chk == VM_CHK_NEXT_DECOMMIT) w0 = VW(code_decommittment_queue_state);
    else { w0 = VWC(reverted_queue_head); n = 4; at = idx; }
    bool same = true;
#pragma unroll
    for (int j = 0; j < 12; j++) same &= j >= n || cols.sw64(w0 + 2 * j, at) == q[j];
    if (chk == VM_CHK_CUR_STACK) { if (!same) vm_report(dev, row, ZKC_VM_CHK_CALLSTACK); return; }
    if (chk == VM_CHK_CUR_RB_HEAD) { if (!same) vm_report(dev, row, ZKC_VM_CHK_ROLLBACK_QUEUE); return; }
    if (!same) vm_report(dev, last_row ? row : row + 1, ZKC_VM_CHK_SNAPSHOT);
    if (last_row) {
        uint64_t *dst = chk == VM_CHK_NEXT_MEMQ ? dev->s_final.memory_queue_state
                      : chk == VM_CHK_NEXT_STACK ? dev->s_final.stack_sponge_state
                      : chk == VM_CHK_NEXT_DECOMMIT ? dev->s_final.code_decommittment_queue_state : dev->s_final.current_context.log_queue_forward_tail;
        for (int j = 0; j < n; j++) dst[j] = q[j];
    }
}

// slot k of every cycle that has one, one thread per job (one launch per slot)
__global__ void __launch_bounds__(128)
vm_sponge_kernel(VmDev *devs, VmCols cols,
                 const zkc_vm_callstack_witness *__restrict__ cws, uint32_t n_cw, VmPushScratch ps, int k, size_t limit, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ps.counts[k]) return;
    vm_sponge_job(devs, cols, cws, n_cw, ps, k, ps.lists[(size_t)k * total + i], limit, nullptr);
}

// the far calls of the launch: slots 5, 6, 7 (code-hash read: a chain) and 8 (decommitment queue) of a cycle by one thread
__global__ void __launch_bounds__(128)
vm_sponge_far_kernel(VmDev *devs, VmCols cols,
                     const zkc_vm_callstack_witness *__restrict__ cws, uint32_t n_cw, VmPushScratch ps, size_t limit, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ps.counts[VM_JOB_SLOTS_LO]) return;
    const size_t g = ps.lists[(size_t)VM_JOB_SLOTS_LO * total + i];
    const uint32_t m = (uint32_t)ps.meta[g * 2] & 0xFFFFu;
    for (int k = VM_JOB_SLOTS_LO; k < VM_JOB_SLOTS; k++)
        if ((m >> k) & 1) vm_sponge_job(devs, cols, cws, n_cw, ps, k, g, limit, nullptr);
}

// All slots in ONE persistent launch (grid = what is resident at once).  The jobs are numbered slot by slot (every
// slot's range padded to a multiple of 32) and handed out in that order, a warp at a time, by an atomic ticket; a job
// that continues another slot's output spins on that job's flag.  Its dependency has a smaller number, so it was handed
// out earlier to a warp that is resident and never waits on a larger number: no deadlock, and no launch boundary --
// the tail of slot k overlaps the head of slot k + 1 instead of draining the machine five times.
__global__ void __launch_bounds__(128)
vm_sponge_persistent_kernel(VmDev *devs, VmCols cols,
                            const zkc_vm_callstack_witness *__restrict__ cws, uint32_t n_cw, VmPushScratch ps, unsigned long long *ticket,
                            uint32_t *flags, size_t limit, size_t total) {
    const unsigned lane = threadIdx.x & 31;
    unsigned long long start[VM_JOB_SLOTS_LO + 1];
    start[0] = 0;
#pragma unroll
    for (int k = 0; k < VM_JOB_SLOTS_LO; k++) start[k + 1] = start[k] + (((unsigned long long)ps.counts[k] + 31) & ~31ull);
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(ticket, 32ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= start[VM_JOB_SLOTS_LO]) return;
        int k = 0;
#pragma unroll
        for (int j = 1; j < VM_JOB_SLOTS_LO; j++) k += t >= start[j];
        const unsigned long long i = t - start[k] + lane;
        if (i < ps.counts[k]) vm_sponge_job(devs, cols, cws, n_cw, ps, k, ps.lists[(size_t)k * total + i], limit, flags);
    }
}

// the sponge columns of the trace: 9 enforce flags + 9 x 12 permutation outputs (zeros where a relation is not enforced)
__global__ void __launch_bounds__(256)
vm_sponge_trace_kernel(VmPushScratch ps, uint64_t *__restrict__ trace, size_t limit, size_t n_instances, size_t row0, size_t row_count) {
    const size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= row_count * n_instances) return;
    const size_t inst = l / row_count, row = row0 + (l - inst * row_count), g = inst * limit + row;
    uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit + row;
    const uint32_t m = (uint32_t)ps.meta[g * 2] & 0xFFFFu;
#pragma unroll
    for (int k = 0; k < ZKC_VM_NUM_SPONGES; k++) {
        const bool on = (m >> k) & 1;
        t[(size_t)(ZKC_VM_SPONGE_ENFORCE + k) * limit] = on;
        const uint64_t *from = ps.state_of(g, k);
#pragma unroll
        for (int j = 0; j < 12; j++) t[(size_t)(ZKC_VM_SPONGE_FINAL + 12 * k + j) * limit] = on ? from[j] : 0ull;
    }
}

// 64 threads per instance: the four commitments of the closed form are independent sponges
// (ClosedFormInputCompactForm::from_full_form, fsm_input_output/mod.rs:178-255), each run by one 16-lane group with the
// 12-lane permutation (poseidon2.cuh) -- the 31 dependent permutations over a VM state are the latency of this launch.
// Each group flattens its encoding into its own global scratch row and absorbs from there, and skips the sponge whose
// result the start / completion flags mask to zero anyway; group 0 then commits the compact form.
// Two modes.  HINT (mode 0) runs on a side stream concurrently with the cycle launches and commits the closed form from
// the HOST's final snapshot; FINAL (mode 1) runs last: when every snapshot link verified, the state the circuit ends in
// IS that snapshot and the hinted commitment is taken (the 31 dependent permutations are off the critical path),
// otherwise everything is recomputed from the computed final state.  FINAL also owns the check that the hint chain
// starts at the circuit's own start state (snapshot 0 == s0), so that the cycle launch does not wait for the prologue.
constexpr int VM_FLAT_STRIDE = 248;
__global__ void __launch_bounds__(128)
vm_finalize_kernel(VmDev *devs, uint64_t *__restrict__ flat, size_t n_instances, const zkc_vm_state *__restrict__ ends, size_t limit, int mode) {
    __shared__ uint64_t part[2][4][4];
    __shared__ uint64_t compact[2][24];
    const int slot = threadIdx.x >> 6, role = (threadIdx.x >> 4) & 3, i = threadIdx.x & 15;
    const unsigned gm = 0xFFFFu << (threadIdx.x & 16);
    const size_t inst = (size_t)blockIdx.x * 2 + slot;
    const bool active = inst < n_instances;
    VmDev *d = devs + (active ? inst : 0);
    zkc_vm_closed_form &io = d->io;
    const bool hint_mode = mode == 0;
    const zkc_vm_state &state = hint_mode ? ends[2 * (active ? inst : 0) + 1] : d->s_final;
    const bool done = state.context_stack_depth == 0;  // mod.rs:113-122
    const bool start = hint_mode ? io.start_flag != 0 : d->start != 0;
    bool use_hint = false;
    if (!hint_mode && active) {
        const bool chain_starts_right = limit == 0 || vm_state_equal(ends[2 * inst], d->s0);
        if (!chain_starts_right && threadIdx.x % 64 == 0) vm_report(d, 0, ZKC_VM_CHK_SNAPSHOT);
        use_hint = limit != 0 && chain_starts_right && d->hint_ok && !(d->failed_checks & ZKC_VM_CHK_SNAPSHOT);
    }
    __syncthreads();  // the report above is read below
    {
        const uint64_t *buf = flat + (inst * 4 + role) * VM_FLAT_STRIDE;
        int n = 0;
        const bool need = active && !use_hint && (role == 0 ? !done : role == 1 ? !start : role == 2 ? true : done);
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
        uint64_t *cf = compact[slot];
        cf[0] = start; cf[1] = done;
        for (int j = 0; j < 4; j++) {
            cf[2 + j] = part[slot][2][j];
            cf[6 + j] = part[slot][3][j];   // zero unless done
            cf[10 + j] = part[slot][1][j];  // zero if start
            cf[14 + j] = part[slot][0][j];  // zero if done
        }
    }
    if (active && role == 0 && i == 0 && !hint_mode) {
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
        d->status = st;
    }
    __syncthreads();
    if (active && !hint_mode) {  // the 64 threads of the instance publish the state the circuit ended in
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&state);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&io.hidden_fsm_output);
        for (int j = threadIdx.x & 63; j < (int)(sizeof(zkc_vm_state) / 4); j += 64) dst[j] = src[j];
    }
    if (role == 0) {
        const uint64_t c = commit_encoding_coop(gm, compact[slot], active && !use_hint ? 18 : 0, i);
        if (active && i < 4) {
            if (hint_mode) d->hint_commitment[i] = c;
            else d->commitment[i] = use_hint ? d->hint_commitment[i] : c;
        }
        if (active && hint_mode && i == 0) d->hint_ok = 1;
    }
}

// ---- out-of-circuit run: one thread per independent VM instance ---------------------------------------------------------
// Rollback-queue resolution.  A frame's rollback segment is hash-chained BACKWARDS: every revertable log claims a new head
// h' with H(rollback item, h') == current head (log.rs:351-371, :583-632), a frame that returns ok hands its segment to
// its parent (the parent's saved head must be the child's tail, ret.rs:396-404) and a frame that reverts must have its
// head where the forward queue ends (ret.rs:373-383), its tail becoming the new forward tail.  The claimed heads / frame
// tails are therefore only known once a frame's fate is: its events (own call marker, logs, merged children) are walked
// from the most recent one back, starting from the required final head.  Pass 1 (resolve) does that and patches the
// witness; pass 2 replays with the witness given and records the snapshots.
struct VmLists { VmEntry *entries; long long *first, *last; };
__device__ void vm_list_merge_into_parent(VmLists &L, size_t child) {
    if (L.first[child] < 0) return;
    L.entries[L.first[child]].prev = L.last[child - 1];
    if (L.first[child - 1] < 0) L.first[child - 1] = L.first[child];
    L.last[child - 1] = L.last[child];
    L.first[child] = L.last[child] = -1;
}
__device__ void vm_list_walk(VmLists &L, size_t depth, uint64_t (&cur)[4], zkc_vm_cycle_witness *witness, bool resolve, VmSim &sim, bool restore) {
    for (long long e = L.last[depth]; e >= 0; e = L.entries[e].prev) {
        VmEntry &en = L.entries[e];
        if (resolve) for (int i = 0; i < 4; i++) witness[e].rollback[i] = cur[i];
        if (en.kind == 2) {
            uint64_t st[12];
            for (int i = 0; i < 4; i++) { st[i] = en.enc16[i]; st[4 + i] = cur[i]; st[8 + i] = en.cap[i]; }
            poseidon2_permute(st);
            for (int i = 0; i < 4; i++) cur[i] = st[i];
            if (restore && en.slot >= 0) {
                for (int i = 0; i < 8; i++) sim.storage[en.slot].value[i] = en.prev_value[i];
                sim.storage[en.slot].written = en.prev_written;
            }
        }
    }
    L.first[depth] = L.last[depth] = -1;
}

struct VmRegsOfState {
    const zkc_vm_state &s;
    __device__ __forceinline__ zkc_vm_register get(uint32_t r) const { return s.registers[r]; }
    __device__ __forceinline__ uint32_t low(uint32_t r) const { return s.registers[r].value[0]; }
};

struct VmSimScratch {
    zkc_vm_register *pages;            // [n][4][VM_PAGE_WORDS]
    VmSlot *storage;                   // [n][VM_STORAGE_SLOTS]
    zkc_vm_callstack_witness *stack;   // [n][VM_SIM_MAX_DEPTH]
    VmEntry *entries;                  // [n][cycles]
    long long *first, *last;           // [n][VM_SIM_MAX_DEPTH + 2]
    uint64_t *root_tails;              // [n][4] in / out
    uint32_t *n_cw;                    // [n] out
};

__global__ void __launch_bounds__(32)
vm_simulate_kernel(const zkc_vm_isa *__restrict__ isa, const zkc_vm_state *__restrict__ initial, const uint32_t *__restrict__ code,
                   size_t code_words, size_t n_instances, size_t cycles, VmSimScratch sc, zkc_vm_state *__restrict__ snapshots,
                   zkc_vm_cycle_witness *__restrict__ witness, zkc_vm_callstack_witness *__restrict__ cw_out, uint32_t cw_cap, int resolve,
                   unsigned long long *first_bad, uint32_t *failed_checks) {
    const size_t inst = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= n_instances) return;
    VmSim sim;
    for (int k = 0; k < 4; k++) sim.pages[k] = sc.pages + (inst * 4 + k) * VM_PAGE_WORDS;
    sim.storage = sc.storage + inst * VM_STORAGE_SLOTS;
    sim.stack = sc.stack + inst * VM_SIM_MAX_DEPTH;
    sim.cw_out = cw_out + inst * (size_t)cw_cap; sim.cw_cap = cw_cap; sim.n_cw = 0;
    sim.ev_kind = 0; sim.overflow = 0;
    VmLists L;
    L.entries = sc.entries + inst * cycles;
    L.first = sc.first + inst * (VM_SIM_MAX_DEPTH + 2); L.last = sc.last + inst * (VM_SIM_MAX_DEPTH + 2);
    for (uint32_t i = 0; i < VM_SIM_MAX_DEPTH + 2; i++) { L.first[i] = -1; L.last[i] = -1; }
    for (size_t i = 0; i < (size_t)4 * VM_PAGE_WORDS; i++) sim.pages[0][i] = reg_zero();
    for (uint32_t i = 0; i < VM_STORAGE_SLOTS; i++) { sim.storage[i].used = 0; sim.storage[i].written = 0; for (int j = 0; j < 8; j++) sim.storage[i].value[j] = 0; }
    for (size_t i = 0; i < code_words && i < VM_PAGE_WORDS; i++)
        for (int j = 0; j < 8; j++) sim.pages[0][i].value[j] = code[(inst * code_words + i) * 8 + j];
    zkc_vm_state s = initial[inst];
    sim.page_ids[0] = s.current_context.code_page;
    sim.page_ids[1] = s.current_context.base_page + 1; sim.page_ids[2] = s.current_context.base_page + 2; sim.page_ids[3] = s.current_context.base_page + 3;
    uint64_t root_tail[4];
    for (int i = 0; i < 4; i++) root_tail[i] = sc.root_tails[inst * 4 + i];
    {   // the frame below the root: the empty context initial_bootloader_state hashes into the stack sponge (loading.rs:96-186)
        for (int i = 0; i < 4; i++) { s.current_context.reverted_queue_head[i] = root_tail[i]; s.current_context.reverted_queue_tail[i] = root_tail[i]; }
        zkc_vm_context empty;
        memset(&empty, 0, sizeof empty);
        for (int i = 0; i < 4; i++) { empty.reverted_queue_tail[i] = root_tail[i]; empty.reverted_queue_head[i] = root_tail[i]; }
        empty.is_kernel_mode = 1;
        sim.stack[0].context = empty;
        uint64_t enc[32], sp12[12];
        vm_context_encode(empty, enc);
        for (int i = 0; i < 12; i++) { sp12[i] = 0; sim.stack[0].previous_sponge_state[i] = 0; }
        for (int r = 0; r < 4; r++) {
            for (int i = 0; i < 8; i++) sp12[i] = enc[8 * r + i];
            poseidon2_permute(sp12);
        }
        for (int i = 0; i < 12; i++) s.stack_sponge_state[i] = sp12[i];
    }
    zkc_vm_state *snaps = snapshots + inst * (cycles + 1);
    zkc_vm_cycle_witness *wit = witness + inst * cycles;
    if (!resolve) snaps[0] = s;
    const uint32_t ignore = resolve ? ZKC_VM_CHK_ROLLBACK_QUEUE : 0u;  // the joins cannot hold before the witness is resolved
    for (size_t c = 0; c < cycles; c++) {
        zkc_vm_cycle_witness w;
        memset(&w, 0, sizeof w);
        if (!resolve) for (int i = 0; i < 4; i++) w.rollback[i] = wit[c].rollback[i];
        VmDelta d;
        zkc_vm_context nctx;
        VmSimOut so;
        for (int i = 0; i < 12; i++) { so.memq[i] = s.memory_queue_state[i]; so.stack[i] = s.stack_sponge_state[i]; }
        const size_t depth = s.context_stack_depth;
        uint64_t fwd_before[4];
        for (int i = 0; i < 4; i++) fwd_before[i] = s.current_context.log_queue_forward_tail[i];
        const uint32_t checks = vm_cycle_dev<true>(isa, nullptr, s, VmRegsOfState{s}, d, nctx, w, nullptr, 0, &sim, &so, nullptr, nullptr, nullptr, 0, 0) & ~ignore;
        vm_apply_delta(s, d, nctx, isa);
        for (int i = 0; i < 12; i++) s.memory_queue_state[i] = so.memq[i];
        if (d.ctx_replaced) for (int i = 0; i < 12; i++) s.stack_sponge_state[i] = so.stack[i];
        if (d.fwd_tail_kind == 1) for (int i = 0; i < 4; i++) s.current_context.log_queue_forward_tail[i] = so.fwd_tail[i];
        if (sim.ev_kind == 1 && depth + 1 < VM_SIM_MAX_DEPTH + 2) {
            VmEntry &en = L.entries[c];
            en.kind = 1; en.slot = -1; en.prev = L.last[depth + 1];
            L.last[depth + 1] = (long long)c;
            if (L.first[depth + 1] < 0) L.first[depth + 1] = (long long)c;
        } else if (sim.ev_kind == 4) {
            L.entries[c] = sim.ev;
            L.entries[c].prev = L.last[depth];
            L.last[depth] = (long long)c;
            if (L.first[depth] < 0) L.first[depth] = (long long)c;
        } else if (sim.ev_kind == 2 && depth >= 2) vm_list_merge_into_parent(L, depth);
        else if (sim.ev_kind == 3) {
            // the reverting frame's final head is the forward tail at this point; its tail becomes the forward tail
            vm_list_walk(L, depth, fwd_before, wit, resolve != 0, sim, true);
            if (resolve) {
                for (int i = 0; i < 4; i++) s.current_context.log_queue_forward_tail[i] = fwd_before[i];
                if (depth == 1) for (int i = 0; i < 4; i++) root_tail[i] = fwd_before[i];
            }
        }
        if (!resolve) { wit[c] = w; snaps[c + 1] = s; }  // pass 1 writes the witness only through the walks (rollback fields)
        if (checks) {
            atomicOr(failed_checks, checks);
            atomicMin(first_bad, ((unsigned long long)(inst * cycles + c) << 16) | checks);
        }
    }
    if (resolve) {
        // frames still open (and the root after an ok exit): as if they all returned ok, chained from the given tail
        size_t top = s.context_stack_depth;
        if (top > VM_SIM_MAX_DEPTH) top = VM_SIM_MAX_DEPTH;
        for (size_t dd = top; dd >= 2; dd--) vm_list_merge_into_parent(L, dd);
        if (L.last[1] >= 0) vm_list_walk(L, 1, root_tail, wit, true, sim, false);
        for (int i = 0; i < 4; i++) sc.root_tails[inst * 4 + i] = root_tail[i];
    }
    sc.n_cw[inst] = sim.n_cw;
    if (sim.overflow) {
        atomicOr(failed_checks, (uint32_t)ZKC_VM_CHK_UNSUPPORTED_OPCODE);
        atomicMin(first_bad, ((unsigned long long)(inst * cycles) << 16) | ZKC_VM_CHK_UNSUPPORTED_OPCODE);
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

#include "main_vm_stream.cuh"

namespace zkc {

// what the per-cycle inputs of a call are: records (array of structs, host or device), columns (device) or segmented
// streams (host); and where the witness goes: a DENSE / COMPACT trace (host or device) or the PACKED transport form (host)
struct VmInput {
    const zkc_vm_state *snapshots = nullptr;
    const zkc_vm_cycle_witness *witness = nullptr;
    bool rows_on_device = false;
    const zkc_vm_columns *columns = nullptr;
    const zkc_vm_input_stream *const *streams = nullptr;
    const zkc_vm_callstack_witness *callstack_witness = nullptr;
    size_t n_callstack_witness = 0;
    bool callstack_on_device = false;
    zkc_vm_packed_trace *packed = nullptr;
};

static inline size_t vm_round32(size_t n) { return (std::max<size_t>(n, 1) + 31) & ~(size_t)31; }

// [lines] pieces of `width` bytes, src_pitch / dst_pitch apart
static cudaError_t vm_copy_lines(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width, size_t lines, cudaMemcpyKind kind,
                                 cudaStream_t st) {
    if (!width || !lines) return cudaSuccess;
    if (lines == 1 || (width == src_pitch && width == dst_pitch)) return cudaMemcpyAsync(dst, src, width * lines, kind, st);
    if (src_pitch < (1ull << 31) && dst_pitch < (1ull << 31)) return cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width, lines, kind, st);
    for (size_t i = 0; i < lines; i++) {
        const cudaError_t e = cudaMemcpyAsync((char *)dst + i * dst_pitch, (const char *)src + i * src_pitch, width, kind, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// CUtensorMap of state_words [VM_WORDS][stride] (u32, row pitch stride * 4 bytes) with a [32 x rows] box; false when the driver
// entry point is missing or the layout does not meet the 16-byte rules (the kernel then takes its plain-load path)
static bool vm_make_tmaps(const uint32_t *st, size_t stride, VmTmaps *out) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = []() -> encode_fn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return (encode_fn)p;
    }();
    if (!encode || (reinterpret_cast<uintptr_t>(st) & 15) || (stride & 3) || stride >= (1ull << 31)) return false;
    const cuuint64_t dims[2] = {stride, (cuuint64_t)VM_WORDS}, strides[1] = {stride * 4};
    const cuuint32_t elem[2] = {1, 1};
    const cuuint32_t box_a[2] = {32, (cuuint32_t)VM_TILE_A}, box_b[2] = {32, (cuuint32_t)VM_TILE_B};
    void *base = const_cast<uint32_t *>(st);
    return encode(&out->a, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box_a, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
           encode(&out->b, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box_b, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int vm_entry_batch(zkc_ctx *ctx, zkc_vm_closed_form *ios, size_t n_instances, const zkc_vm_isa *isa, const VmInput &in, size_t limit,
                          const zkc_vm_options *options, bool trace_dev, uint64_t *trace, uint64_t *commitments, zkc_status *statuses) {
    const bool have_stream = in.streams != nullptr;
    const bool have_rows = in.columns == nullptr && !have_stream;
    if (!ctx || !ios || !isa || !commitments || !statuses ||
        (limit && n_instances && have_rows && (!in.snapshots || !in.witness)) ||
        (limit && n_instances && in.columns && (!in.columns->state_words || !in.columns->witness_words ||
                                                in.columns->state_stride < n_instances * (limit + 1) || in.columns->witness_stride < n_instances * limit)) ||
        (in.n_callstack_witness && !in.callstack_witness) || in.n_callstack_witness > 0xFFFFFFFFull ||
        limit > 0x0FFFFFFFull || n_instances > 0x00FFFFFFull || limit * n_instances > 0xFFFFFFFFull)
        return ZKC_ERR_INVALID_ARGUMENT;
    if (!n_instances) return ZKC_OK;
    size_t seg_cycles = 0, n_segments = 0, blob_total = 0;
    if (have_stream) {
        for (size_t i = 0; i < n_instances; i++) {
            const zkc_vm_input_stream *st = in.streams[i];
            if (!st || st->limit != limit || !st->segment_cycles || !st->n_segments || !st->segments) return ZKC_ERR_INVALID_ARGUMENT;
            if (i == 0) { seg_cycles = st->segment_cycles; n_segments = st->n_segments; }
            if (st->segment_cycles != seg_cycles || st->n_segments != n_segments || n_segments != std::max<size_t>(1, (limit + seg_cycles - 1) / seg_cycles))
                return ZKC_ERR_INVALID_ARGUMENT;
            for (size_t k = 0; k < n_segments; k++) {
                const zkc_vm_segment_header *h = (const zkc_vm_segment_header *)st->segments[k].blob;
                if (!h || h->magic != ZKC_VM_SEGMENT_MAGIC || h->blob_bytes != st->segments[k].blob_bytes || h->first_cycle != k * seg_cycles ||
                    h->n_cycles != std::min(seg_cycles, limit - k * seg_cycles))
                    return ZKC_ERR_INVALID_ARGUMENT;
                blob_total += (h->blob_bytes + 255) & ~(size_t)255;
            }
        }
    }
    zkc_status *status = statuses;
    for (size_t i = 0; i < n_instances; i++) statuses[i] = zkc_status{ZKC_OK, 0, -1, 0, 0};
    const bool rows_host = have_rows && !in.rows_on_device;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const size_t rows = limit * n_instances, n_cw = in.n_callstack_witness * n_instances;
    const bool own_cols = have_rows || have_stream;
    const size_t st_stride = own_cols ? vm_round32(rows + n_instances) : in.columns->state_stride;
    const size_t wt_stride = own_cols ? vm_round32(rows) : in.columns->witness_stride;
    zkc_vm_packed_trace *packed = in.packed;
    if (packed && (trace || (rows && (!packed->cols8 || !packed->cols16 || !packed->cols32 || !packed->cols64)) ||
                   (packed->aux_capacity && !packed->aux_records) || (packed->sponge_capacity && !packed->sponge_records) ||
                   (packed->limb_capacity && !packed->limb_records)))
        return ZKC_ERR_INVALID_ARGUMENT;
    const bool compact = packed || (options && options->trace_layout == ZKC_VM_TRACE_COMPACT && trace);
    if (!packed && options && options->trace_layout > ZKC_VM_TRACE_COMPACT) return ZKC_ERR_INVALID_ARGUMENT;
    if (!packed && compact && options->sponge_records_capacity && !options->sponge_records) return ZKC_ERR_INVALID_ARGUMENT;
    const int ncols = compact ? ZKC_VM_COMPACT_COLS : ZKC_VM_NUM_COLS, aux_base = compact ? ZKC_VM_COMPACT_OP_AUX : ZKC_VM_OP_AUX;
    const size_t rec_cap = packed ? 0 : (compact ? (size_t)options->sponge_records_capacity : 0);
    // Host buffers: the rows are cut into chunks and pipelined over three streams -- H2D of chunk i+1 | kernels of chunk i |
    // D2H of chunk i-1 -- so that a step costs max(H2D, D2H) instead of their sum (PCIe is full duplex).  A stream's chunks
    // are its segments.
    size_t n_chunks = 1;
    if ((rows_host || (trace && !trace_dev)) && limit >= 8192 && rows >= (1u << 16)) n_chunks = limit >= (1u << 18) ? 16 : 4;
    if (have_stream) n_chunks = n_segments;
    const bool side_streams = n_chunks > 1 || have_stream || packed;
    if (side_streams && !ctx->copy_streams()) { if (!have_stream) n_chunks = 1; }
    const size_t chunk_rows = have_stream ? seg_cycles : (limit + n_chunks - 1) / std::max<size_t>(n_chunks, 1);
    const bool piped = side_streams && ctx->copy_in && ctx->copy_out;
    size_t bytes = zkc_carver::bytes(n_instances, sizeof(VmDev)) + zkc_carver::bytes(1, sizeof(zkc_vm_isa)) + zkc_carver::bytes(2 * n_instances, sizeof(zkc_vm_state));
    if (rows_host) bytes += zkc_carver::bytes(rows + n_instances, sizeof(zkc_vm_state)) + zkc_carver::bytes(rows + 1, sizeof(zkc_vm_cycle_witness));
    if (!in.callstack_on_device) bytes += zkc_carver::bytes(n_cw + 1, sizeof(zkc_vm_callstack_witness));
    if (own_cols) bytes += zkc_carver::bytes(st_stride * VM_WORDS, 4) + zkc_carver::bytes(wt_stride * VM_WIT_WORDS, 4);
    if (have_stream) bytes += blob_total + 256;
    if ((trace && !trace_dev) || packed) bytes += zkc_carver::bytes((size_t)ncols * rows, 8);
    if (trace && !trace_dev) bytes += zkc_carver::bytes(rec_cap + 1, sizeof(zkc_vm_sponge_record));
    const size_t chunk_cells = chunk_rows * n_instances;  // rows of one chunk over the batch
    if (packed)
        bytes += zkc_carver::bytes(rows * VM_PK_N8, 1) + zkc_carver::bytes(rows * VM_PK_N16, 2) + zkc_carver::bytes(rows * VM_PK_N32, 4) +
                 zkc_carver::bytes(rows * VM_PK_N64, 8) + zkc_carver::bytes(n_chunks * chunk_cells + 1, sizeof(zkc_vm_aux_record)) +
                 zkc_carver::bytes(n_chunks * chunk_cells * VM_JOB_SLOTS + 1, sizeof(zkc_vm_sponge_record)) +
                 zkc_carver::bytes(n_chunks * chunk_cells * 3 + 1, sizeof(zkc_vm_limb_record)) + zkc_carver::bytes(4 * n_chunks, 8);
    bytes += zkc_carver::bytes(n_instances * 4 * VM_FLAT_STRIDE, 8);
    bytes += zkc_carver::bytes(16 * n_chunks, 4) + zkc_carver::bytes(VM_JOB_SLOTS * rows, 4) + zkc_carver::bytes(rows * 2, 8) + zkc_carver::bytes(rows, 4) + zkc_carver::bytes(rows * VM_EXP_WORDS, 4) +
             zkc_carver::bytes(rows * VM_JOB_SLOTS * 8, 8) + zkc_carver::bytes(rows * VM_JOB_SLOTS * 12, 8) +
             zkc_carver::bytes(rows * VM_JOB_SLOTS, 4) + zkc_carver::bytes(n_chunks + 32, 8) + 4096;  // + slack: the slot arrays are carved in two parts
    void *blk = ctx->scratch(bytes);
    const size_t h_counts_off = (n_instances * sizeof(VmDev) + 63) & ~(size_t)63;
    VmDev *h = (VmDev *)ctx->pinned(h_counts_off + 32 * n_chunks + 64);
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    unsigned long long *h_counts = (unsigned long long *)((char *)h + h_counts_off);  // [n_chunks][4]: aux, sponge, limb records of a chunk
    zkc_carver cv(blk);
    VmDev *d = cv.take<VmDev>(n_instances);
    zkc_vm_isa *disa = cv.take<zkc_vm_isa>(1);
    zkc_vm_state *ends = cv.take<zkc_vm_state>(2 * n_instances);
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
    const zkc_vm_callstack_witness *dcw = in.callstack_witness;
    uint64_t *dtrace = trace;
    uint64_t *flat = cv.take<uint64_t>(n_instances * 4 * VM_FLAT_STRIDE);
    cudaStream_t s_in = piped ? ctx->copy_in : s, s_out = piped ? ctx->copy_out : s;
    uint32_t *counts = cv.take<uint32_t>(16 * n_chunks);
    uint32_t *lists = cv.take<uint32_t>(VM_JOB_SLOTS * rows);
    VmPushScratch ps;
    ps.meta = cv.take<uint64_t>(rows * 2);
    ps.link = cv.take<uint32_t>(rows);
    ps.exp = cv.take<uint32_t>(rows * VM_EXP_WORDS);
    ps.enc = cv.take<uint64_t>(rows * VM_JOB_SLOTS_LO * 8);
    ps.state = cv.take<uint64_t>(rows * VM_JOB_SLOTS_LO * 12);
    ps.enc_hi = cv.take<uint64_t>(rows * VM_JOB_SLOTS_HI * 8);
    ps.state_hi = cv.take<uint64_t>(rows * VM_JOB_SLOTS_HI * 12);
    uint32_t *job_flags = cv.take<uint32_t>(rows * VM_JOB_SLOTS);
    unsigned long long *tickets = cv.take<unsigned long long>(n_chunks + 32);  // [0 .. n_chunks) chunk tickets, [n_chunks + 31] record count
    ZKC_CUDA(ctx, status, cudaMemsetAsync(counts, 0, 64 * n_chunks, s));
    static const int sponge_mode = getenv("ZKC_VM_SPONGE_MODE") ? atoi(getenv("ZKC_VM_SPONGE_MODE")) : 0;
    if (sponge_mode) ZKC_CUDA(ctx, status, cudaMemsetAsync(job_flags, 0, rows * VM_JOB_SLOTS * 4, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(tickets, 0, (n_chunks + 32) * 8, s));
    static int sponge_blocks_per_sm = 0;
    if (!sponge_blocks_per_sm) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sponge_blocks_per_sm, vm_sponge_persistent_kernel, 128, 0) != cudaSuccess || sponge_blocks_per_sm < 1)
            sponge_blocks_per_sm = 1;
    }
    std::vector<cudaEvent_t> used_events;
    auto event = [&]() { cudaEvent_t e = ctx->get_event(); used_events.push_back(e); return e; };
    if (piped) {  // the copy streams start after what the main stream has queued so far (scratch reuse, VmDev upload)
        cudaEvent_t e0 = event();
        ZKC_CUDA(ctx, status, cudaEventRecord(e0, s));
        ZKC_CUDA(ctx, status, cudaStreamWaitEvent(s_in, e0, 0));
        ZKC_CUDA(ctx, status, cudaStreamWaitEvent(s_out, e0, 0));
    }
    // the records the transposition reads: the caller's (device) or a staging copy of the host's
    const zkc_vm_state *rsnap = in.snapshots;
    const zkc_vm_cycle_witness *rwit = in.witness;
    zkc_vm_state *bs = nullptr;
    zkc_vm_cycle_witness *bw = nullptr;
    if (rows_host && limit) {
        bs = cv.take<zkc_vm_state>(rows + n_instances);
        bw = cv.take<zkc_vm_cycle_witness>(rows + 1);
        rsnap = bs; rwit = bw;
    }
    if (!in.callstack_on_device) {
        zkc_vm_callstack_witness *bc = cv.take<zkc_vm_callstack_witness>(n_cw + 1);
        if (n_cw) ZKC_CUDA(ctx, status, cudaMemcpyAsync(bc, in.callstack_witness, n_cw * sizeof(zkc_vm_callstack_witness), cudaMemcpyHostToDevice, s));
        dcw = bc;
    }
    VmCols cols;
    if (own_cols) {
        uint32_t *stc = cv.take<uint32_t>(st_stride * VM_WORDS), *wtc = cv.take<uint32_t>(wt_stride * VM_WIT_WORDS);
        cols = VmCols{stc, st_stride, wtc, wt_stride};
    } else cols = VmCols{in.columns->state_words, st_stride, in.columns->witness_words, wt_stride};
    // tensor maps of the state columns for the cycle kernel's TMA tiles (ZKC_VM_TMA=1 selects that variant; default: plain loads, measured faster)
    VmTmaps tmaps;
    memset(&tmaps, 0, sizeof tmaps);
    static const bool tma_wanted = getenv("ZKC_VM_TMA") && atoi(getenv("ZKC_VM_TMA"));
    const int use_tma = tma_wanted && vm_make_tmaps(cols.st, st_stride, &tmaps);
    uint32_t *stc_w = const_cast<uint32_t *>(cols.st), *wtc_w = const_cast<uint32_t *>(cols.wt);  // written only when they are this call's scratch
    char *dblobs = have_stream ? cv.take<char>(blob_total + 256) : nullptr;
    if ((trace && !trace_dev) || packed) dtrace = cv.take<uint64_t>((size_t)ncols * rows);
    ps.records = nullptr; ps.n_records = tickets + n_chunks + 31; ps.records_capacity = rec_cap;
    if (compact && !packed) ps.records = (trace_dev || !rec_cap) ? options->sponge_records : cv.take<zkc_vm_sponge_record>(rec_cap + 1);
    if (compact && !packed && !ps.records) ps.records = (zkc_vm_sponge_record *)(tickets + n_chunks + 30);  // capacity 0: count only (never written)
    // PACKED: typed column blocks + per-chunk record regions (worst-case capacity: nothing is dropped on the device)
    VmPackOut po{};
    zkc_vm_sponge_record *sp_regions = nullptr;
    unsigned long long *rec_counts = nullptr;  // [n_chunks][4]: aux, sponge, limb, -
    if (packed) {
        po.c8 = cv.take<uint8_t>(rows * VM_PK_N8); po.c16 = cv.take<uint16_t>(rows * VM_PK_N16);
        po.c32 = cv.take<uint32_t>(rows * VM_PK_N32); po.c64 = cv.take<uint64_t>(rows * VM_PK_N64);
        po.rows = rows;
        po.aux = cv.take<zkc_vm_aux_record>(n_chunks * chunk_cells + 1);
        sp_regions = cv.take<zkc_vm_sponge_record>(n_chunks * chunk_cells * VM_JOB_SLOTS + 1);
        po.limb = cv.take<zkc_vm_limb_record>(n_chunks * chunk_cells * 3 + 1);
        rec_counts = cv.take<unsigned long long>(4 * n_chunks);
        ZKC_CUDA(ctx, status, cudaMemsetAsync(rec_counts, 0, 32 * n_chunks, s));
        packed->n_aux_records = 0; packed->n_sponge_records = 0; packed->n_limb_records = 0;
    }
    // Side stream: the start state (4 dependent permutations) and the closed-form commitments from the host's final
    // snapshot (31 dependent permutations) run beside the cycle launches; the FINAL pass below takes them when every link
    // verified.  `ends` = first and final snapshot of every instance as records (a stream's arrive with its last segment).
    cudaStream_t s_aux = ctx->aux_stream();
    if (!s_aux) s_aux = s;
    cudaEvent_t e_hint = nullptr, e_link = nullptr;
    auto launch_hint = [&](bool ends_from_columns) -> int {
        if (s_aux != s) {
            cudaEvent_t e = event();
            ZKC_CUDA(ctx, status, cudaEventRecord(e, s));
            ZKC_CUDA(ctx, status, cudaStreamWaitEvent(s_aux, e, 0));
        }
        ctx->launches++;
        vm_prologue_kernel<<<(unsigned)((n_instances + 31) / 32), 32, 0, s_aux>>>(d, disa, n_instances);
        if (limit) {
            if (!ends_from_columns) {
                const cudaMemcpyKind kind = rows_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
                ZKC_CUDA(ctx, status, vm_copy_lines(ends, 2 * sizeof(zkc_vm_state), in.snapshots, (limit + 1) * sizeof(zkc_vm_state), sizeof(zkc_vm_state),
                                                    n_instances, kind, s_aux));
                ZKC_CUDA(ctx, status, vm_copy_lines(ends + 1, 2 * sizeof(zkc_vm_state), in.snapshots + limit, (limit + 1) * sizeof(zkc_vm_state),
                                                    sizeof(zkc_vm_state), n_instances, kind, s_aux));
            } else {
                ctx->launches++;
                vm_gather_ends_kernel<<<(unsigned)((2 * n_instances * 32 + 127) / 128), 128, 0, s_aux>>>(cols, ends, limit, n_instances);
            }
            ctx->launches++;
            vm_finalize_kernel<<<(unsigned)((n_instances + 1) / 2), 128, 0, s_aux>>>(d, flat, n_instances, ends, limit, 0);
        }
        if (s_aux != s) {
            e_hint = event();
            ZKC_CUDA(ctx, status, cudaEventRecord(e_hint, s_aux));
        }
        return ZKC_OK;
    };
    if (!have_stream || !limit) { const int rc = launch_hint(!have_rows && limit); if (rc) return rc; }
    // PACKED: the records of a chunk go home two chunks later, when their count is known, without stalling the queue
    std::vector<cudaEvent_t> count_events(n_chunks, nullptr);
    size_t next_records = 0, aux_home = 0, sp_home = 0, limb_home = 0;
    auto send_records_home = [&](size_t c) -> int {
        ZKC_CUDA(ctx, status, cudaEventSynchronize(count_events[c]));
        const unsigned long long na = h_counts[4 * c], nsp = h_counts[4 * c + 1], nl = h_counts[4 * c + 2];
        const size_t ca = aux_home < packed->aux_capacity ? std::min<size_t>(na, packed->aux_capacity - aux_home) : 0;
        const size_t cs = sp_home < packed->sponge_capacity ? std::min<size_t>(nsp, packed->sponge_capacity - sp_home) : 0;
        if (ca) ZKC_CUDA(ctx, status, cudaMemcpyAsync(packed->aux_records + aux_home, po.aux + c * chunk_cells, ca * sizeof(zkc_vm_aux_record), cudaMemcpyDeviceToHost, s_out));
        if (cs) ZKC_CUDA(ctx, status, cudaMemcpyAsync(packed->sponge_records + sp_home, sp_regions + c * chunk_cells * VM_JOB_SLOTS, cs * sizeof(zkc_vm_sponge_record),
                                                      cudaMemcpyDeviceToHost, s_out));
        const size_t cl = limb_home < packed->limb_capacity ? std::min<size_t>(nl, packed->limb_capacity - limb_home) : 0;
        if (cl) ZKC_CUDA(ctx, status, cudaMemcpyAsync(packed->limb_records + limb_home, po.limb + c * chunk_cells * 3, cl * sizeof(zkc_vm_limb_record),
                                                      cudaMemcpyDeviceToHost, s_out));
        aux_home += na; sp_home += nsp; limb_home += nl;
        packed->n_aux_records += na; packed->n_sponge_records += nsp; packed->n_limb_records += nl;
        return ZKC_OK;
    };
    size_t blob_off = 0;
    for (size_t c = 0; c < n_chunks && limit; c++) {
        const size_t r0 = c * chunk_rows;
        if (r0 >= limit) break;
        const size_t cnt = std::min(chunk_rows, limit - r0), n_thr = cnt * n_instances;
        if (rows_host) {
            ZKC_CUDA(ctx, status, vm_copy_lines(bs + r0, (limit + 1) * sizeof(zkc_vm_state), in.snapshots + r0, (limit + 1) * sizeof(zkc_vm_state),
                                                (cnt + 1) * sizeof(zkc_vm_state), n_instances, cudaMemcpyHostToDevice, s_in));
            ZKC_CUDA(ctx, status, vm_copy_lines(bw + r0, limit * sizeof(zkc_vm_cycle_witness), in.witness + r0, limit * sizeof(zkc_vm_cycle_witness),
                                                cnt * sizeof(zkc_vm_cycle_witness), n_instances, cudaMemcpyHostToDevice, s_in));
            if (piped) {
                cudaEvent_t e = event();
                ZKC_CUDA(ctx, status, cudaEventRecord(e, s_in));
                ZKC_CUDA(ctx, status, cudaStreamWaitEvent(s, e, 0));
            }
        }
        if (have_rows) {  // records -> columns (snapshot rows r0 .. r0 + cnt inclusive: the last one is the next chunk's first)
            const size_t st_tiles = ((cnt + 1 + 31) / 32) * n_instances, wt_tiles = ((cnt + 31) / 32) * n_instances;
            ZKC_LAUNCH(ctx, "vm_rows_to_columns", vm_rows_to_columns_kernel<VM_WORDS>, (unsigned)st_tiles, 32, 0,
                       reinterpret_cast<const uint32_t *>(rsnap), stc_w, st_stride, limit + 1, n_instances, r0, cnt + 1);
            ZKC_LAUNCH(ctx, "vm_rows_to_columns", vm_rows_to_columns_kernel<VM_WIT_WORDS>, (unsigned)wt_tiles, 32, 0,
                       reinterpret_cast<const uint32_t *>(rwit), wtc_w, wt_stride, limit, n_instances, r0, cnt);
        }
        if (have_stream) {  // segment c of every instance: one copy per blob, then expansion into the columns
            const size_t blob_off0 = blob_off;
            for (size_t i = 0; i < n_instances; i++) {
                const zkc_vm_input_segment &sg = in.streams[i]->segments[c];
                ZKC_CUDA(ctx, status, cudaMemcpyAsync(dblobs + blob_off, sg.blob, sg.blob_bytes, cudaMemcpyHostToDevice, s_in));
                blob_off += (sg.blob_bytes + 255) & ~(size_t)255;
            }
            if (piped) {
                cudaEvent_t e = event();
                ZKC_CUDA(ctx, status, cudaEventRecord(e, s_in));
                ZKC_CUDA(ctx, status, cudaStreamWaitEvent(s, e, 0));
            }
            size_t off = blob_off0;
            for (size_t i = 0; i < n_instances; i++) {
                const zkc_vm_input_segment &sg = in.streams[i]->segments[c];
                const zkc_vm_segment_header &hd = *(const zkc_vm_segment_header *)sg.blob;
                const char *db = dblobs + off;
                off += (sg.blob_bytes + 255) & ~(size_t)255;
                const size_t sbase = i * (limit + 1) + r0, wbase = i * limit + r0, n1 = cnt + 1;
                if (hd.n_dense_state)
                    ZKC_LAUNCH(ctx, "vm_expand", vm_expand_dense_kernel, dim3((unsigned)((n1 + 255) / 256), hd.n_dense_state), 256, 0,
                               (const uint16_t *)(db + hd.off_dense_state_word), (const uint32_t *)(db + hd.off_dense_state), n1, stc_w, st_stride, sbase);
                if (hd.n_sparse_state) {
                    const uint32_t *offs = (const uint32_t *)((const char *)sg.blob + hd.off_sparse_state_offsets);
                    uint32_t most = 0;
                    for (int w = 0; w < ZKC_VM_STATE_WORDS; w++) most = std::max(most, offs[w + 1] - offs[w]);
                    ZKC_LAUNCH(ctx, "vm_expand", vm_expand_runs_kernel, dim3((most + 127) / 128, ZKC_VM_STATE_WORDS), 128, 0,
                               (const uint32_t *)(db + hd.off_sparse_state_offsets), (const uint32_t *)(db + hd.off_sparse_state_index),
                               (const uint32_t *)(db + hd.off_sparse_state_value), (uint32_t)n1, stc_w, st_stride, sbase);
                }
                ZKC_CUDA(ctx, status, cudaMemset2DAsync(wtc_w + wbase, wt_stride * 4, 0, cnt * 4, VM_WIT_WORDS, s));
                if (hd.n_dense_witness)
                    ZKC_LAUNCH(ctx, "vm_expand", vm_expand_dense_kernel, dim3((unsigned)((cnt + 255) / 256), hd.n_dense_witness), 256, 0,
                               (const uint16_t *)(db + hd.off_dense_witness_word), (const uint32_t *)(db + hd.off_dense_witness), cnt, wtc_w, wt_stride, wbase);
                if (hd.n_sparse_witness)
                    ZKC_LAUNCH(ctx, "vm_expand", vm_expand_scatter_kernel, (hd.n_sparse_witness + 255) / 256, 256, 0,
                               (const uint32_t *)(db + hd.off_sparse_witness_offsets), (int)ZKC_VM_WITNESS_WORDS,
                               (const uint32_t *)(db + hd.off_sparse_witness_index), (const uint32_t *)(db + hd.off_sparse_witness_value),
                               hd.n_sparse_witness, wtc_w, wt_stride, wbase);
            }
            if (r0 + cnt == limit) { const int rc = launch_hint(true); if (rc) return rc; }  // the final snapshots are in the columns now
        }
        ps.counts = counts + 16 * c;
        ps.lists = lists + n_instances * r0;
        if (packed) {
            ps.records = sp_regions + c * chunk_cells * VM_JOB_SLOTS;
            ps.n_records = rec_counts + 4 * c + 1;
            ps.records_capacity = chunk_cells * VM_JOB_SLOTS;
        }
        if (use_tma)
            ZKC_LAUNCH(ctx, "vm_cycles", vm_cycles_kernel<true>, (unsigned)((n_thr + 127) / 128), 128, 0, d, disa, cols, dcw,
                       (uint32_t)in.n_callstack_witness, dtrace, limit, n_instances, r0, cnt, ps, ncols, aux_base, tmaps, use_tma);
        else
            ZKC_LAUNCH(ctx, "vm_cycles", vm_cycles_kernel<false>, (unsigned)((n_thr + 127) / 128), 128, 0, d, disa, cols, dcw,
                       (uint32_t)in.n_callstack_witness, dtrace, limit, n_instances, r0, cnt, ps, ncols, aux_base, tmaps, use_tma);
        // The link check streams the state columns (HBM-bound) while the sponge kernels below are integer-bound: it runs beside
        // them on the side stream (in line when per-kernel profiling wants one kernel at a time).
        static const bool link_in_line = getenv("ZKC_VM_LINK_IN_LINE") != nullptr;  // for A/B timing
        if (s_aux != s && !ctx->profiling && !link_in_line) {
            cudaEvent_t e = event();
            ZKC_CUDA(ctx, status, cudaEventRecord(e, s));
            ZKC_CUDA(ctx, status, cudaStreamWaitEvent(s_aux, e, 0));
            ctx->launches++;
            vm_link_kernel<<<(unsigned)((n_thr + 255) / 256), 256, 0, s_aux>>>(d, cols, (const uint32_t *)ps.link, (const uint32_t *)ps.exp, limit, n_instances, r0, cnt);
            e_link = event();
            ZKC_CUDA(ctx, status, cudaEventRecord(e_link, s_aux));
        } else {
            ZKC_LAUNCH(ctx, "vm_link", vm_link_kernel, (unsigned)((n_thr + 255) / 256), 256, 0, d, cols, (const uint32_t *)ps.link, (const uint32_t *)ps.exp, limit,
                       n_instances, r0, cnt);
        }
        // every Poseidon2 relation of the chunk: one persistent launch over the per-slot job lists, or one launch per slot
        if (sponge_mode == 0) {
            for (int k = 0; k < VM_JOB_SLOTS_LO; k++)
                ZKC_LAUNCH(ctx, "vm_sponge", vm_sponge_kernel, (unsigned)((n_thr + 127) / 128), 128, 0, d, cols, dcw,
                           (uint32_t)in.n_callstack_witness, ps, k, limit, rows);
        } else {
            const size_t max_blocks = (size_t)ctx->sm_count * sponge_blocks_per_sm, need_blocks = (n_thr * VM_JOB_SLOTS + 127) / 128;
            ZKC_LAUNCH(ctx, "vm_sponge", vm_sponge_persistent_kernel, (unsigned)std::min(max_blocks, std::max<size_t>(need_blocks, 1)), 128, 0, d, cols,
                       dcw, (uint32_t)in.n_callstack_witness, ps, tickets + c, job_flags, limit, rows);
        }
        ZKC_LAUNCH(ctx, "vm_sponge_far", vm_sponge_far_kernel, (unsigned)((n_thr + 127) / 128), 128, 0, d, cols, dcw,
                   (uint32_t)in.n_callstack_witness, ps, limit, rows);
        if (dtrace && !compact)
            ZKC_LAUNCH(ctx, "vm_sponge_trace", vm_sponge_trace_kernel, (unsigned)((n_thr + 255) / 256), 256, 0, ps, dtrace, limit, n_instances, r0, cnt);
        if (packed) {
            VmPackOut pc = po;
            pc.aux = po.aux + c * chunk_cells; pc.n_aux = rec_counts + 4 * c; pc.aux_cap = chunk_cells;
            pc.limb = po.limb + c * chunk_cells * 3; pc.n_limb = rec_counts + 4 * c + 2;
            ZKC_LAUNCH(ctx, "vm_pack", vm_pack_kernel, (unsigned)((n_thr + 255) / 256), 256, 0, dtrace, limit, n_instances, r0, cnt, pc);
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(h_counts + 4 * c, rec_counts + 4 * c, 32, cudaMemcpyDeviceToHost, s));
            count_events[c] = event();
            ZKC_CUDA(ctx, status, cudaEventRecord(count_events[c], s));
            if (piped) ZKC_CUDA(ctx, status, cudaStreamWaitEvent(s_out, count_events[c], 0));
            ZKC_CUDA(ctx, status, vm_copy_lines(packed->cols8 + r0, limit, po.c8 + r0, limit, cnt, n_instances * (size_t)VM_PK_N8, cudaMemcpyDeviceToHost, s_out));
            ZKC_CUDA(ctx, status, vm_copy_lines(packed->cols16 + r0, limit * 2, po.c16 + r0, limit * 2, cnt * 2, n_instances * (size_t)VM_PK_N16, cudaMemcpyDeviceToHost, s_out));
            ZKC_CUDA(ctx, status, vm_copy_lines(packed->cols32 + r0, limit * 4, po.c32 + r0, limit * 4, cnt * 4, n_instances * (size_t)VM_PK_N32, cudaMemcpyDeviceToHost, s_out));
            ZKC_CUDA(ctx, status, vm_copy_lines(packed->cols64 + r0, limit * 8, po.c64 + r0, limit * 8, cnt * 8, n_instances * (size_t)VM_PK_N64, cudaMemcpyDeviceToHost, s_out));
            while (next_records + 2 <= c) { const int rc = send_records_home(next_records++); if (rc) return rc; }
        } else if (trace && !trace_dev) {
            if (piped) {
                cudaEvent_t e = event();
                ZKC_CUDA(ctx, status, cudaEventRecord(e, s));
                ZKC_CUDA(ctx, status, cudaStreamWaitEvent(s_out, e, 0));
            }
            ZKC_CUDA(ctx, status, vm_copy_lines(trace + r0, limit * 8, dtrace + r0, limit * 8, cnt * 8, n_instances * (size_t)ncols, cudaMemcpyDeviceToHost, s_out));
        }
    }
    if (e_hint) ZKC_CUDA(ctx, status, cudaStreamWaitEvent(s, e_hint, 0));
    if (e_link) ZKC_CUDA(ctx, status, cudaStreamWaitEvent(s, e_link, 0));  // the side stream is in order: the last link covers all
    ZKC_LAUNCH(ctx, "vm_finalize", vm_finalize_kernel, (unsigned)((n_instances + 1) / 2), 128, 0, d, flat, n_instances, ends, limit, 1);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, n_instances * sizeof(VmDev), cudaMemcpyDeviceToHost, s));
    if (packed)
        while (next_records < n_chunks && count_events[next_records]) { const int rc = send_records_home(next_records++); if (rc) return rc; }
    unsigned long long n_records = 0;
    if (compact && !packed) ZKC_CUDA(ctx, status, cudaMemcpyAsync(&n_records, ps.n_records, 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    if (compact && !packed && !trace_dev && rec_cap && n_records)  // the records of the whole call, one copy (they are ~4 % of the dense sponge columns)
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(options->sponge_records, ps.records, std::min<size_t>(n_records, rec_cap) * sizeof(zkc_vm_sponge_record),
                                              cudaMemcpyDeviceToHost, s));
    if (compact && !packed) ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    if (piped) {
        ZKC_CUDA(ctx, status, cudaStreamSynchronize(s_in));
        ZKC_CUDA(ctx, status, cudaStreamSynchronize(s_out));
    }
    for (cudaEvent_t e : used_events) ctx->event_pool.push_back(e);
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
    if (compact && !packed) statuses[0].reserved = (uint32_t)std::min<unsigned long long>(n_records, 0xFFFFFFFFull);
    return worst;
}

}  // namespace zkc

extern "C" int zkc_main_vm_entry_point_batch(zkc_ctx *ctx, zkc_vm_closed_form *ios, size_t n_instances, const zkc_vm_isa *isa,
                                             const zkc_vm_state *snapshots, const zkc_vm_cycle_witness *witness,
                                             const zkc_vm_callstack_witness *callstack_witness, size_t n_callstack_witness, size_t limit,
                                             const zkc_vm_options *options, int on_device, uint64_t *trace, uint64_t *commitments,
                                             zkc_status *statuses) {
    VmInput in;
    in.snapshots = snapshots; in.witness = witness; in.rows_on_device = (on_device & ZKC_INPUTS_ON_DEVICE) != 0;
    in.callstack_witness = callstack_witness; in.n_callstack_witness = n_callstack_witness; in.callstack_on_device = in.rows_on_device;
    return vm_entry_batch(ctx, ios, n_instances, isa, in, limit, options, (on_device & ZKC_TRACE_ON_DEVICE) != 0, trace, commitments, statuses);
}

extern "C" int zkc_main_vm_entry_point_columns(zkc_ctx *ctx, zkc_vm_closed_form *ios, size_t n_instances, const zkc_vm_isa *isa,
                                               const zkc_vm_columns *columns, const zkc_vm_callstack_witness *callstack_witness,
                                               size_t n_callstack_witness, size_t limit, const zkc_vm_options *options, int trace_on_device,
                                               uint64_t *trace, uint64_t *commitments, zkc_status *statuses) {
    if (!columns) return ZKC_ERR_INVALID_ARGUMENT;
    VmInput in;
    in.columns = columns;
    in.callstack_witness = callstack_witness; in.n_callstack_witness = n_callstack_witness; in.callstack_on_device = true;
    return vm_entry_batch(ctx, ios, n_instances, isa, in, limit, options, trace_on_device != 0, trace, commitments, statuses);
}

extern "C" int zkc_main_vm_entry_point_stream(zkc_ctx *ctx, zkc_vm_closed_form *ios, size_t n_instances, const zkc_vm_isa *isa,
                                              const zkc_vm_input_stream *const *streams, const zkc_vm_callstack_witness *callstack_witness,
                                              size_t n_callstack_witness, size_t limit, const zkc_vm_options *options, zkc_vm_packed_trace *out,
                                              uint64_t *commitments, zkc_status *statuses) {
    if (!streams) return ZKC_ERR_INVALID_ARGUMENT;
    VmInput in;
    in.streams = streams;
    in.callstack_witness = callstack_witness; in.n_callstack_witness = n_callstack_witness; in.callstack_on_device = false;
    in.packed = out;
    return vm_entry_batch(ctx, ios, n_instances, isa, in, limit, options, false, nullptr, commitments, statuses);
}

extern "C" int zkc_main_vm_rows_to_columns(zkc_ctx *ctx, const zkc_vm_state *snapshots, const zkc_vm_cycle_witness *witness, size_t n_instances,
                                           size_t limit, uint32_t *state_words, size_t state_stride, uint32_t *witness_words, size_t witness_stride) {
    if (!ctx || !snapshots || !witness || !state_words || !witness_words || state_stride < n_instances * (limit + 1) ||
        witness_stride < n_instances * limit)
        return ZKC_ERR_INVALID_ARGUMENT;
    if (!n_instances) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    ZKC_LAUNCH(ctx, "vm_rows_to_columns", vm_rows_to_columns_kernel<VM_WORDS>, (unsigned)(((limit + 1 + 31) / 32) * n_instances), 32, 0,
               reinterpret_cast<const uint32_t *>(snapshots), state_words, state_stride, limit + 1, n_instances, 0, limit + 1);
    if (limit)
        ZKC_LAUNCH(ctx, "vm_rows_to_columns", vm_rows_to_columns_kernel<VM_WIT_WORDS>, (unsigned)(((limit + 31) / 32) * n_instances), 32, 0,
                   reinterpret_cast<const uint32_t *>(witness), witness_words, witness_stride, limit, n_instances, 0, limit);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    return ZKC_OK;
}

extern "C" int zkc_main_vm_entry_point(zkc_ctx *ctx, zkc_vm_closed_form *io, const zkc_vm_isa *isa, const zkc_vm_state *snapshots,
                                       const zkc_vm_cycle_witness *witness, const zkc_vm_callstack_witness *callstack_witness,
                                       size_t n_callstack_witness, size_t limit, const zkc_vm_options *options,
                                       int on_device, uint64_t *trace, uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!io || !commitment) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    const int rc = zkc_main_vm_entry_point_batch(ctx, io, 1, isa, snapshots, witness, callstack_witness, n_callstack_witness, limit, options,
                                                 on_device, trace, commitment, status);
    if (rc == ZKC_ERR_INVALID_ARGUMENT) status->code = rc;
    return rc;
}

extern "C" int zkc_main_vm_simulate(zkc_ctx *ctx, const zkc_vm_isa *isa, const zkc_vm_state *initial_states, const uint32_t *code,
                                    size_t code_words, size_t n_instances, size_t cycles, zkc_vm_state *snapshots_out,
                                    zkc_vm_cycle_witness *witness_out, zkc_vm_callstack_witness *callstack_witness_out,
                                    size_t callstack_capacity, uint32_t *n_callstack_out, uint64_t *rollback_tails_out,
                                    zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !isa || !initial_states || !code || !snapshots_out || !witness_out || code_words > 65536 ||
        (callstack_capacity && !callstack_witness_out) || callstack_capacity > 0xFFFFFFFFull) {
        status->code = ZKC_ERR_INVALID_ARGUMENT;
        return ZKC_ERR_INVALID_ARGUMENT;
    }
    if (!n_instances) return ZKC_OK;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const size_t n = n_instances;
    const size_t bytes = zkc_carver::bytes(1, sizeof(zkc_vm_isa)) + zkc_carver::bytes(4, 8) +
                         zkc_carver::bytes(n * 4 * VM_PAGE_WORDS, sizeof(zkc_vm_register)) + zkc_carver::bytes(n * VM_STORAGE_SLOTS, sizeof(VmSlot)) +
                         zkc_carver::bytes(n * VM_SIM_MAX_DEPTH, sizeof(zkc_vm_callstack_witness)) + zkc_carver::bytes(n * (cycles + 1), sizeof(VmEntry)) +
                         2 * zkc_carver::bytes(n * (VM_SIM_MAX_DEPTH + 2), 8) + zkc_carver::bytes(n * 4, 8) + zkc_carver::bytes(n, 4) +
                         zkc_carver::bytes(1, sizeof(zkc_vm_callstack_witness));
    void *blk = ctx->scratch(bytes);
    const size_t hbytes = 32 + n * 32 + n * 4;
    unsigned long long *hres = (unsigned long long *)ctx->pinned(hbytes);
    if (!blk || !hres) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    uint64_t *htails = (uint64_t *)(hres + 4);
    uint32_t *hncw = (uint32_t *)(htails + 4 * n);
    zkc_carver cv(blk);
    zkc_vm_isa *disa = cv.take<zkc_vm_isa>(1);
    unsigned long long *dres = cv.take<unsigned long long>(4);
    VmSimScratch sc;
    sc.pages = cv.take<zkc_vm_register>(n * 4 * VM_PAGE_WORDS);
    sc.storage = cv.take<VmSlot>(n * VM_STORAGE_SLOTS);
    sc.stack = cv.take<zkc_vm_callstack_witness>(n * VM_SIM_MAX_DEPTH);
    sc.entries = cv.take<VmEntry>(n * (cycles + 1));
    sc.first = cv.take<long long>(n * (VM_SIM_MAX_DEPTH + 2));
    sc.last = cv.take<long long>(n * (VM_SIM_MAX_DEPTH + 2));
    sc.root_tails = cv.take<uint64_t>(n * 4);
    sc.n_cw = cv.take<uint32_t>(n);
    zkc_vm_callstack_witness *dummy_cw = cv.take<zkc_vm_callstack_witness>(1);
    cudaStream_t s = ctx->stream;
    hres[0] = ~0ull; hres[1] = 0;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(disa, isa, sizeof(zkc_vm_isa), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(dres, hres, 16, cudaMemcpyHostToDevice, s));
    // the block's rollback tail each run starts from: the start state's own (resolved by pass 1)
    ZKC_CUDA(ctx, status, cudaMemcpy2DAsync(sc.root_tails, 32, (const char *)initial_states + offsetof(zkc_vm_state, current_context) +
                                            offsetof(zkc_vm_context, reverted_queue_tail), sizeof(zkc_vm_state), 32, n, cudaMemcpyDeviceToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(witness_out, 0, n * cycles * sizeof(zkc_vm_cycle_witness), s));
    zkc_vm_callstack_witness *cwo = callstack_capacity ? callstack_witness_out : dummy_cw;
    for (int resolve = 1; resolve >= 0; resolve--)
        ZKC_LAUNCH(ctx, "vm_simulate", vm_simulate_kernel, (unsigned)((n + 31) / 32), 32, 0, disa, initial_states, code, code_words, n, cycles,
                   sc, snapshots_out, witness_out, cwo, (uint32_t)callstack_capacity, resolve, dres, (uint32_t *)(dres + 1));
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(hres, dres, 16, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(htails, sc.root_tails, n * 32, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(hncw, sc.n_cw, n * 4, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    if (rollback_tails_out) memcpy(rollback_tails_out, htails, n * 32);
    if (n_callstack_out) memcpy(n_callstack_out, hncw, n * 4);
    const uint32_t checks = (uint32_t)hres[1];
    if (checks) {
        status->failed_checks = checks;
        status->first_bad_row = (int64_t)(hres[0] >> 16);
        status->code = (checks & ZKC_VM_CHK_UNSUPPORTED_OPCODE) ? ZKC_ERR_UNSUPPORTED : ZKC_ERR_UNSATISFIED;
    }
    return status->code;
}
