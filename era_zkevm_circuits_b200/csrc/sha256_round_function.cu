// sha256 precompile circuit on sm_100a: sha256_round_function_entry_point
// (/root/reference/src/sha256_round_function/mod.rs:343-470) and its work cycle sha256_precompile_inner (:88-340).
// Same decomposition as keccak256_round_function.cu: a call of `num_rounds` blocks takes exactly `num_rounds` cycles
// (2 memory reads + one compression each, digest written at the last one), so the plan is a prefix sum of the
// calls' round counts; one thread per call chains the compressions; the memory queue is handled by
// precompile_common.cuh.
#include "ctx.cuh"
#include "log_query.cuh"
#include "precompile_common.cuh"
#include "scan.cuh"
#include "sha256.cuh"

namespace zkc {

struct ShDev {
    zkc_sha256_closed_form io;
    zkc_precompile_options opt;
    uint64_t n_requests, n_reads, n_memory_states, limit;
    uint32_t start, prologue_checks, n_units, unit0_fresh;
    zkc_sha256_fsm s0;
    zkc_queue_state4 rq0;
    zkc_queue_state12 mq0;
    uint64_t commit_obs_in[4], commit_fsm_in[4];
    zkc_sha256_fsm s_last, s_final;
    uint32_t popped_requests, pad0;
    uint64_t req_head_final[4];
    unsigned long long first_bad;
    uint32_t failed_checks, hint_bad;
    uint64_t commitment[4];
    zkc_status status;
};

struct ShPlan {
    uint32_t cycles, reads, pushes, pad;
};
struct ShPlanOp {
    static __device__ __forceinline__ ShPlan identity() { return ShPlan{0, 0, 0, 0}; }
    static __device__ __forceinline__ ShPlan combine(const ShPlan &a, const ShPlan &b) {
        return ShPlan{a.cycles + b.cycles, a.reads + b.reads, a.pushes + b.pushes, 0};
    }
};

__device__ int sh_encode_fsm(const zkc_sha256_fsm &f, uint64_t *dst) {
    int n = 0;
    dst[n++] = f.read_precompile_call; dst[n++] = f.read_words_for_round; dst[n++] = f.completed;
    for (int i = 0; i < 8; i++) dst[n++] = f.sha256_inner_state[i];
    dst[n++] = f.timestamp_to_use_for_read; dst[n++] = f.timestamp_to_use_for_write;
    dst[n++] = f.input_page; dst[n++] = f.input_offset; dst[n++] = f.output_page; dst[n++] = f.output_offset; dst[n++] = f.num_rounds;
    n += put_queue_state4(dst + n, f.log_queue_state);
    for (int i = 0; i < 12; i++) dst[n++] = f.memory_queue_state.head[i];
    for (int i = 0; i < 12; i++) dst[n++] = f.memory_queue_state.tail[i];
    dst[n++] = f.memory_queue_state.length;
    return n;  // 52
}
__device__ int sh_put_q12(uint64_t *dst, const zkc_queue_state12 &s) {
    for (int i = 0; i < 12; i++) dst[i] = s.head[i];
    for (int i = 0; i < 12; i++) dst[12 + i] = s.tail[i];
    dst[24] = s.length;
    return 25;
}

__global__ void sh_prologue_kernel(ShDev *d) {
    // warp 0: scalar start selection; warps 1 / 2: the two input commitments, every permutation on 12 cooperating lanes
    __shared__ uint64_t buf[2][53];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane >= 16 || (warp == 0 && lane != 0)) return;
    const unsigned gm = 0xFFFFu;
    const zkc_sha256_closed_form &io = d->io;
    if (warp == 0) {
        const bool start = io.start_flag != 0;
        d->start = start;
        d->rq0 = start ? io.initial_log_queue_state : io.hidden_fsm_input.log_queue_state;
        d->mq0 = start ? io.initial_memory_queue_state : io.hidden_fsm_input.memory_queue_state;
        zkc_sha256_fsm s;
        if (start) {  // mod.rs:411-419; the placeholder FSM carries the IV (input.rs:36-50)
            memset(&s, 0, sizeof s);
            s.read_precompile_call = 1;
            for (int i = 0; i < 8; i++) s.sha256_inner_state[i] = SHA_IV[i];
        } else s = io.hidden_fsm_input;
        const bool cfi = s.read_precompile_call && d->rq0.length == 0;  // :121-135
        if (cfi) { s.read_precompile_call = 0; s.read_words_for_round = 0; s.completed = 1; }
        d->s0 = s; d->s_last = s; d->s_final = s;
        d->popped_requests = 0;
        uint32_t checks = 0;
        for (int i = 0; i < 4; i++) if (io.initial_log_queue_state.head[i]) checks |= ZKC_KC_CHK_TRIVIAL_HEAD;
        for (int i = 0; i < 12; i++) if (io.initial_memory_queue_state.head[i]) checks |= ZKC_KC_CHK_TRIVIAL_HEAD;
        d->prologue_checks = checks;
        d->unit0_fresh = (s.read_precompile_call || s.completed) ? 1 : 0;
        const uint64_t avail = d->rq0.length < d->n_requests ? d->rq0.length : d->n_requests;
        d->n_units = 1 + (s.completed ? 0 : (uint32_t)avail);
    } else {
        uint64_t *b = buf[warp - 1];
        int n = 0;
        if (lane == 0) {
            if (warp == 1) {
                n = put_queue_state4(b, io.initial_log_queue_state);
                n += sh_put_q12(b + n, io.initial_memory_queue_state);
            } else {
                n = sh_encode_fsm(io.hidden_fsm_input, b);
            }
        }
        __syncwarp(gm);
        n = __shfl_sync(gm, n, 0, 16);
        const uint64_t c = commit_encoding_coop(gm, b, n, lane);
        if (lane < 4) (warp == 1 ? d->commit_obs_in : d->commit_fsm_in)[lane] = c;
    }
}

// cycles a call occupies: its round count (an illegal 0 never terminates: runs to the end of the instance)
__device__ __forceinline__ uint32_t sh_cycles_of(uint32_t num_rounds, size_t limit) {
    return num_rounds ? (num_rounds < limit ? num_rounds : (uint32_t)limit) : (uint32_t)limit;
}

__global__ void __launch_bounds__(SCAN_THREADS)
sh_plan_kernel(ShDev *d, const zkc_log_query *__restrict__ requests, ShPlan *__restrict__ starts, ScanGlobal *sg,
               TileStateT<ShPlan> *tiles) {
    __shared__ ScanSharedT<ShPlan> sh;
    const unsigned int tile = scan_take_ticket(sg, sh);
    const size_t u = (size_t)tile * SCAN_THREADS + threadIdx.x;
    const uint32_t n_units = d->n_units;
    ShPlan v = ShPlanOp::identity();
    if (u < n_units) {
        uint32_t rounds = 0;
        bool run = true;
        if (u == 0) { run = !d->unit0_fresh; rounds = d->s0.num_rounds; }
        else rounds = __ldg(&requests[u - 1].key[6]);
        if (run) {
            v.cycles = sh_cycles_of(rounds, d->limit);
            // reads happen while num_rounds != 0 at the start of the cycle
            const uint32_t reading = rounds ? v.cycles : 0;
            v.reads = 2 * reading;
            v.pushes = 2 * reading + ((rounds && rounds == v.cycles) ? 1 : 0);
        }
    }
    ShPlan incl;
    const ShPlan excl = scan_tile_generic<ShPlan, ShPlanOp>(v, tile, ShPlanOp::identity(), tiles, sh, incl);
    if (u < n_units) starts[u] = excl;
    if (u + 1 == n_units) starts[n_units] = incl;
}

__device__ __forceinline__ void sh_report(ShDev *d, size_t row, uint32_t checks) {
    if (!checks) return;
    atomicOr(&d->failed_checks, checks);
    atomicMin(&d->first_bad, ((unsigned long long)row << 16) | checks);
}

// one iteration of the main work cycle, mod.rs:146-330
__device__ __forceinline__ void sh_cycle(zkc_sha256_fsm &s, const zkc_log_query &call, bool queue_empty_after,
                                         const uint32_t *__restrict__ reads, size_t n_reads, size_t &read_cursor,
                                         uint64_t *__restrict__ push_enc, uint32_t *__restrict__ slot_meta, uint32_t &push_ordinal,
                                         uint64_t *__restrict__ trace, size_t limit, size_t row, uint32_t &checks) {
#define TR(col) trace[(size_t)(col) * limit + row]
    const bool wr = trace != nullptr;
    const bool read_call = s.read_precompile_call;
    if (wr) { TR(ZKC_SH_FLAGS_IN + 0) = s.read_precompile_call; TR(ZKC_SH_FLAGS_IN + 1) = s.read_words_for_round; TR(ZKC_SH_FLAGS_IN + 2) = s.completed; }
    if (read_call) {
        s.input_offset = call.key[0]; s.output_offset = call.key[2]; s.input_page = call.key[4]; s.output_page = call.key[5];
        s.num_rounds = call.key[6];
        s.timestamp_to_use_for_read = call.timestamp; s.timestamp_to_use_for_write = call.timestamp + 1;
        if (s.num_rounds == 0) checks |= ZKC_SH_CHK_ZERO_ROUNDS;
    }
    const bool reset_buffer = read_call || s.completed;
    s.read_words_for_round = read_call || s.read_words_for_round;
    s.read_precompile_call = 0;
    const bool should_read = s.num_rounds != 0;
    if (wr) {
        TR(ZKC_SH_PARAMS + 0) = s.input_page; TR(ZKC_SH_PARAMS + 1) = s.input_offset; TR(ZKC_SH_PARAMS + 2) = s.output_page;
        TR(ZKC_SH_PARAMS + 3) = s.output_offset; TR(ZKC_SH_PARAMS + 4) = s.num_rounds;
        TR(ZKC_SH_TS_READ) = s.timestamp_to_use_for_read; TR(ZKC_SH_TS_WRITE) = s.timestamp_to_use_for_write;
        TR(ZKC_SH_RESET_BUFFER) = reset_buffer; TR(ZKC_SH_SHOULD_READ) = should_read;
    }
    uint32_t m[16];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        uint32_t value[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (should_read) {
            if (read_cursor < n_reads) {
#pragma unroll
                for (int i = 0; i < 8; i++) value[i] = __ldg(reads + 8 * read_cursor + i);
            } else checks |= ZKC_KC_CHK_WITNESS_EXHAUSTED;
            uint64_t e[8];
            mq_encode(s.timestamp_to_use_for_read, s.input_page, s.input_offset, 0, value, e);
#pragma unroll
            for (int i = 0; i < 8; i++) push_enc[8 * (size_t)push_ordinal + i] = e[i];
            read_cursor++; push_ordinal++;
        }
        if (s.read_words_for_round) s.input_offset++;
        slot_meta[3 * row + q] = push_ordinal | (should_read ? 0x80000000u : 0u);
#pragma unroll
        for (int i = 0; i < 8; i++) m[8 * q + i] = value[7 - i];
        if (wr) {
            const int b = ZKC_SH_QUERY + q * ZKC_SH_QUERY_STRIDE;
#pragma unroll
            for (int i = 0; i < 8; i++) TR(b + i) = value[i];
            TR(b + 21) = s.input_offset;
        }
    }
    if (s.read_words_for_round) s.num_rounds--;
    uint32_t cur[8];
#pragma unroll
    for (int i = 0; i < 8; i++) cur[i] = reset_buffer ? SHA_IV[i] : s.sha256_inner_state[i];
    if (wr) {
#pragma unroll
        for (int i = 0; i < 8; i++) TR(ZKC_SH_STATE_IN + i) = cur[i];
#pragma unroll
        for (int i = 0; i < 16; i++) TR(ZKC_SH_MESSAGE + i) = m[i];
    }
    sha256_compress(cur, m);
#pragma unroll
    for (int i = 0; i < 8; i++) s.sha256_inner_state[i] = cur[i];
    const bool write_result = s.read_words_for_round && s.num_rounds == 0;
    uint32_t result[8];
#pragma unroll
    for (int k = 0; k < 8; k++) result[7 - k] = cur[k];
    if (write_result) {
        uint64_t e[8];
        mq_encode(s.timestamp_to_use_for_write, s.output_page, s.output_offset, 1, result, e);
#pragma unroll
        for (int i = 0; i < 8; i++) push_enc[8 * (size_t)push_ordinal + i] = e[i];
        push_ordinal++;
    }
    slot_meta[3 * row + 2] = push_ordinal | (write_result ? 0x80000000u : 0u);
    const bool nothing_left = write_result && queue_empty_after, process_next = write_result && !queue_empty_after;
    s.read_precompile_call = process_next;
    s.completed = s.completed || nothing_left;
    s.read_words_for_round = !(s.read_precompile_call || s.completed);
    if (wr) {
        TR(ZKC_SH_NUM_ROUNDS) = s.num_rounds;
#pragma unroll
        for (int i = 0; i < 8; i++) { TR(ZKC_SH_STATE_OUT + i) = cur[i]; TR(ZKC_SH_RESULT + i) = result[i]; }
        TR(ZKC_SH_WRITE_RESULT) = write_result;
        TR(ZKC_SH_FLAGS_OUT + 0) = s.read_precompile_call; TR(ZKC_SH_FLAGS_OUT + 1) = s.read_words_for_round; TR(ZKC_SH_FLAGS_OUT + 2) = s.completed;
    }
#undef TR
}

__global__ void __launch_bounds__(128)
sh_calls_kernel(ShDev *d, const zkc_log_query *__restrict__ requests, const uint64_t *__restrict__ req_prev,
                const uint32_t *__restrict__ reads, const ShPlan *__restrict__ starts, uint64_t *__restrict__ push_enc,
                uint32_t *__restrict__ slot_meta, uint64_t *__restrict__ trace) {
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_units = d->n_units;
    if (u >= n_units) return;
    const size_t limit = d->limit;
    const ShPlan st = starts[u], en = starts[u + 1];
    size_t row = st.cycles;
    if (row >= limit || en.cycles == st.cycles) return;
    zkc_sha256_fsm s = d->s0;
    zkc_log_query call = lq_zero();
    const uint32_t aux_byte = d->opt.aux_byte ? d->opt.aux_byte : ZKC_PRECOMPILE_AUX_BYTE_DEFAULT;
    const uint32_t formal = d->opt.precompile_address ? d->opt.precompile_address : ZKC_SHA256_PRECOMPILE_ADDRESS_DEFAULT;
    const uint32_t rq_len0 = d->rq0.length;
    uint64_t head[4];
    uint32_t len_after, checks = 0;
    if (u == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) head[i] = d->rq0.head[i];
        len_after = rq_len0;
    } else {
        call = lq_load(requests + (u - 1));
        s.read_precompile_call = 1; s.read_words_for_round = 0; s.completed = 0;
        uint64_t e[20], sp[12], chain[4];
        bool hint_ok = true;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            chain[i] = __ldg(req_prev + 4 * (u - 1) + i);
            if (u == 1 && chain[i] != d->rq0.head[i]) hint_ok = false;
        }
        lq_encode(call, e);
        lq_absorb_head(e, sp);
        lq_absorb_tail(e, chain, sp);
#pragma unroll
        for (int i = 0; i < 4; i++) head[i] = sp[i];
        if (u + 1 < n_units && starts[u + 1].cycles < limit) {
#pragma unroll
            for (int i = 0; i < 4; i++) hint_ok &= __ldg(req_prev + 4 * u + i) == head[i];
        }
        if (!hint_ok) { checks |= ZKC_KC_CHK_QUEUE_HINT; d->hint_bad = 1; }
        len_after = rq_len0 - (uint32_t)u;
        if (ZKC_LQ_AUX(call.flags) != aux_byte) checks |= ZKC_KC_CHK_AUX_BYTE;
        if (call.address[0] != formal || call.address[1] || call.address[2] || call.address[3] || call.address[4]) checks |= ZKC_KC_CHK_ADDRESS;
    }
    size_t read_cursor = st.reads;
    uint32_t push_ordinal = st.pushes;
    bool first_cycle = true;
    while (row < limit && row < en.cycles) {
        uint32_t cyc_checks = first_cycle ? checks : 0;
        if (trace) {
            for (int i = 0; i < 36; i++) trace[(size_t)(ZKC_SH_CALL_ITEM + i) * limit + row] = first_cycle ? lq_flat(call, i) : 0;
            for (int i = 0; i < 4; i++) trace[(size_t)(ZKC_SH_REQ_HEAD + i) * limit + row] = head[i];
            trace[(size_t)ZKC_SH_REQ_LEN * limit + row] = len_after;
        }
        sh_cycle(s, first_cycle ? call : lq_zero(), len_after == 0, reads, d->n_reads, read_cursor, push_enc, slot_meta, push_ordinal,
                 trace, limit, row, cyc_checks);
        sh_report(d, row, cyc_checks);
        first_cycle = false;
        row++;
    }
    const size_t total = starts[n_units].cycles;
    if (row == limit) d->s_final = s;
    if (en.cycles == total && row == en.cycles) d->s_last = s;
    if (u >= 1 && (u + 1 == n_units || starts[u + 1].cycles >= limit)) {
#pragma unroll
        for (int i = 0; i < 4; i++) d->req_head_final[i] = head[i];
        d->popped_requests = (uint32_t)u;
    }
}

__global__ void __launch_bounds__(128)
sh_tail_kernel(ShDev *d, const ShPlan *__restrict__ starts, uint32_t *__restrict__ slot_meta, uint64_t *__restrict__ trace,
               uint64_t *__restrict__ push_enc) {
    const size_t limit = d->limit;
    const size_t first = starts[d->n_units].cycles;
    const size_t row = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    zkc_sha256_fsm s = d->s_last;
    uint32_t checks = 0;
    if (row > first) {
        // state left by the previous tail row: compression of the zero block from the IV (the cycle resets first)
        uint32_t st[8], w[16];
        for (int i = 0; i < 8; i++) st[i] = SHA_IV[i];
        for (int i = 0; i < 16; i++) w[i] = 0;
        sha256_compress(st, w);
        for (int i = 0; i < 8; i++) s.sha256_inner_state[i] = st[i];
    }
    const uint32_t popped = d->popped_requests;
    const uint32_t len_now = d->rq0.length - popped;
    if (s.read_precompile_call) checks |= ZKC_KC_CHK_WITNESS_EXHAUSTED;
    if (trace) {
        for (int i = 0; i < 36; i++) trace[(size_t)(ZKC_SH_CALL_ITEM + i) * limit + row] = 0;
        for (int i = 0; i < 4; i++) trace[(size_t)(ZKC_SH_REQ_HEAD + i) * limit + row] = popped ? d->req_head_final[i] : d->rq0.head[i];
        trace[(size_t)ZKC_SH_REQ_LEN * limit + row] = len_now;
    }
    size_t rc = 0;
    uint32_t po = starts[d->n_units].pushes;
    sh_cycle(s, lq_zero(), len_now == 0, nullptr, 0, rc, push_enc, slot_meta, po, trace, limit, row, checks);
    sh_report(d, row, checks);
    if (row == limit - 1) d->s_final = s;
}

__global__ void sh_finalize_kernel(ShDev *d, const uint32_t *__restrict__ slot_meta, const uint64_t *__restrict__ states, size_t n_states) {
    // lane 0 does the scalar bookkeeping; the commitments' permutations run on the two 16-lane groups, 12 lanes each
    __shared__ uint64_t e_out[53], o_out[32], compact[24];
    __shared__ uint32_t sh_done, sh_n_out;
    const int lane = threadIdx.x & 31, li = lane & 15;
    const unsigned gm = lane < 16 ? 0xFFFFu : 0xFFFF0000u;
    if (lane == 0) {
    zkc_sha256_closed_form &io = d->io;
    const size_t limit = d->limit;
    zkc_sha256_fsm out = limit ? d->s_final : d->s0;
    zkc_queue_state4 rq = d->rq0;
    const uint32_t popped = limit ? d->popped_requests : 0;
    if (popped) for (int i = 0; i < 4; i++) rq.head[i] = d->req_head_final[i];
    rq.length = d->rq0.length - popped;
    zkc_queue_state12 mq = d->mq0;
    bool hint_bad = d->hint_bad;
    if (limit) {
        const uint32_t pushes = slot_meta[3 * (limit - 1) + 2] & 0x7FFFFFFFu;
        if (pushes) {
            if (pushes - 1 < n_states) for (int i = 0; i < 12; i++) mq.tail[i] = states[12 * (size_t)(pushes - 1) + i];
            else hint_bad = true;
        }
        mq.length += pushes;
    }
    out.log_queue_state = rq;
    out.memory_queue_state = mq;
    out._pad = 0;
    uint32_t checks = d->failed_checks | d->prologue_checks;
    if (rq.length == 0)
        for (int i = 0; i < 4; i++) if (rq.head[i] != rq.tail[i]) checks |= ZKC_KC_CHK_QUEUE_CONSISTENCY;
    const bool done = out.completed;
    zkc_queue_state12 obs_out;
    memset(&obs_out, 0, sizeof obs_out);
    if (done) obs_out = mq;
    uint64_t e_exp[52], o_exp[25];
    const int n_out = sh_encode_fsm(out, e_out);
    sh_put_q12(o_out, obs_out);
    zkc_status st;
    st.code = ZKC_OK; st.cuda_error = 0; st.first_bad_row = -1; st.failed_checks = checks; st.reserved = 0;
    if (d->first_bad != ~0ull) st.first_bad_row = (int64_t)(d->first_bad >> 16);
    if (checks) st.code = ZKC_ERR_UNSATISFIED;
    if (hint_bad) { st.code = ZKC_ERR_QUEUE_WITNESS_INCONSISTENT; st.failed_checks |= ZKC_KC_CHK_QUEUE_HINT; }
    if (d->opt.compare_expected) {
        sh_encode_fsm(io.hidden_fsm_output, e_exp);
        sh_put_q12(o_exp, io.final_memory_state);
        bool same = (io.completion_flag != 0) == done;
        for (int i = 0; i < n_out; i++) same &= e_out[i] == e_exp[i];
        for (int i = 0; i < 25; i++) same &= o_out[i] == o_exp[i];
        if (!same && st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    io.hidden_fsm_output = out;
    io.final_memory_state = obs_out;
    io.completion_flag = done;
    compact[0] = d->start; compact[1] = done;
    for (int i = 0; i < 4; i++) {
        compact[2 + i] = d->commit_obs_in[i];
        compact[10 + i] = d->start ? 0 : d->commit_fsm_in[i];
    }
    d->status = st;
    sh_done = done; sh_n_out = n_out;
    }
    __syncwarp();
    const bool done = sh_done;
    const uint64_t c = commit_encoding_coop(gm, lane < 16 ? e_out : o_out, lane < 16 ? (int)sh_n_out : 25, li);
    if (lane < 4) compact[14 + lane] = done ? 0 : c;
    if (lane >= 16 && lane < 20) compact[6 + lane - 16] = done ? c : 0;
    __syncwarp();
    if (lane < 16) {
        const uint64_t f = commit_encoding_coop(gm, compact, 18, li);
        if (li < 4) d->commitment[li] = f;
    }
}


// ---- constraint evaluation of a finished trace ------------------------------------------------------------------------------
// One thread per cycle re-evaluates every relation sha256_precompile_inner (mod.rs:146-330) places that is local to a cycle or to a
// cycle and its predecessor: the FSM flags carried from the previous cycle, the conditional pop (ranges, aux byte / formal address
// enforcement, queue length / head), Sha256PrecompileCallParams::from_encoding and the selects on the call parameters and
// timestamps, reset / should_read, the two read queries (offset increments, big-endian message words), the round counter, the
// state the compression starts from and the SHA-256 compression itself, the result word, write_result and the next FSM flags, the
// memory queue's length / tail bookkeeping over the three conditional pushes.  With ZKC_GATES_ROUND_FUNCTION also the Poseidon2
// permutations (3 of the pop, 1 per executed memory push).
template <bool ROUND_FUNCTION>
__global__ void __launch_bounds__(128)
sh_check_kernel(ShDev *d, unsigned long long *violations, const uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const bool first = row == 0;
    const zkc_sha256_fsm &s0 = d->s0;
#define TR(col) __ldg(trace + (size_t)(col) * limit + row)
#define TP(col) __ldg(trace + (size_t)(col) * limit + row - 1)
    uint32_t bad = 0;
    const uint32_t aux_byte = d->opt.aux_byte ? d->opt.aux_byte : ZKC_PRECOMPILE_AUX_BYTE_DEFAULT;
    const uint32_t formal = d->opt.precompile_address ? d->opt.precompile_address : ZKC_SHA256_PRECOMPILE_ADDRESS_DEFAULT;
    // FSM flags on entry = the flags the previous cycle left
    const uint64_t rpc = TR(ZKC_SH_FLAGS_IN + 0), rwr_in = TR(ZKC_SH_FLAGS_IN + 1), completed_in = TR(ZKC_SH_FLAGS_IN + 2);
    if ((rpc | rwr_in | completed_in) > 1 || rpc != (first ? (uint64_t)s0.read_precompile_call : TP(ZKC_SH_FLAGS_OUT + 0)) ||
        rwr_in != (first ? (uint64_t)s0.read_words_for_round : TP(ZKC_SH_FLAGS_OUT + 1)) || completed_in != (first ? (uint64_t)s0.completed : TP(ZKC_SH_FLAGS_OUT + 2)))
        bad |= ZKC_SHV_FSM;
    // the conditional pop
    uint64_t f[36], limbs = 0;
#pragma unroll
    for (int i = 0; i < 36; i++) f[i] = TR(ZKC_SH_CALL_ITEM + i);
#pragma unroll
    for (int i = 0; i < 29; i++) limbs |= f[i];
    if ((limbs | f[34] | f[35]) >> 32 || (f[29] | f[33]) >> 8 || (f[30] | f[31] | f[32]) > 1) bad |= ZKC_SHV_BOOLEAN;
    if (rpc && (f[29] != aux_byte || f[0] != formal || (f[1] | f[2] | f[3] | f[4]))) bad |= ZKC_SHV_ENFORCE;
    const uint64_t len_prev = first ? d->rq0.length : TP(ZKC_SH_REQ_LEN), len = TR(ZKC_SH_REQ_LEN);
    if (len + rpc != len_prev || (len >> 32)) bad |= ZKC_SHV_QUEUE;
    uint64_t head[4], head_prev[4];
    bool same = true;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        head[i] = TR(ZKC_SH_REQ_HEAD + i);
        head_prev[i] = first ? d->rq0.head[i] : TP(ZKC_SH_REQ_HEAD + i);
        same &= head[i] == head_prev[i];
        if (head[i] >= GL_P) bad |= ZKC_SHV_BOOLEAN;
    }
    if (!rpc && !same) bad |= ZKC_SHV_QUEUE;
    if (ROUND_FUNCTION && rpc) {
        zkc_log_query q = lq_zero();
#pragma unroll
        for (int i = 0; i < 5; i++) q.address[i] = (uint32_t)f[i];
#pragma unroll
        for (int i = 0; i < 8; i++) { q.key[i] = (uint32_t)f[5 + i]; q.read_value[i] = (uint32_t)f[13 + i]; q.written_value[i] = (uint32_t)f[21 + i]; }
        q.flags = ZKC_LQ_FLAGS((uint32_t)f[29], (uint32_t)f[33], (uint32_t)f[30], (uint32_t)f[31], (uint32_t)f[32]);
        q.tx_number_in_block = (uint32_t)f[34]; q.timestamp = (uint32_t)f[35];
        uint64_t e[20], st[12];
        lq_encode(q, e);
        lq_absorb_head(e, st);
        lq_absorb_tail(e, head_prev, st);
#pragma unroll
        for (int i = 0; i < 4; i++) if (st[i] != head[i]) bad |= ZKC_SHV_ROUND_FUNCTION;
    }
    // call parameters and timestamps after the selects (:175-200); carried: page / offsets / rounds as the previous cycle left them
    uint64_t p[5];
#pragma unroll
    for (int i = 0; i < 5; i++) p[i] = TR(ZKC_SH_PARAMS + i);
    const uint64_t ts_read = TR(ZKC_SH_TS_READ), ts_write = TR(ZKC_SH_TS_WRITE);
    {
        const uint64_t prev_p[5] = {first ? (uint64_t)s0.input_page : TP(ZKC_SH_PARAMS + 0),
                                    first ? (uint64_t)s0.input_offset : TP(ZKC_SH_QUERY + ZKC_SH_QUERY_STRIDE + 21),
                                    first ? (uint64_t)s0.output_page : TP(ZKC_SH_PARAMS + 2), first ? (uint64_t)s0.output_offset : TP(ZKC_SH_PARAMS + 3),
                                    first ? (uint64_t)s0.num_rounds : TP(ZKC_SH_NUM_ROUNDS)};
        const uint64_t from_call[5] = {f[5 + 4], f[5 + 0], f[5 + 5], f[5 + 2], f[5 + 6]};  // from_encoding: key limbs 4, 0, 5, 2, 6
        uint64_t range = ts_read | ts_write;
#pragma unroll
        for (int i = 0; i < 5; i++) { range |= p[i]; if (p[i] != (rpc ? from_call[i] : prev_p[i])) bad |= ZKC_SHV_PARAMS; }
        const uint64_t tr_prev = first ? (uint64_t)s0.timestamp_to_use_for_read : TP(ZKC_SH_TS_READ), tw_prev = first ? (uint64_t)s0.timestamp_to_use_for_write : TP(ZKC_SH_TS_WRITE);
        if (ts_read != (rpc ? f[35] : tr_prev) || ts_write != (rpc ? (uint64_t)(uint32_t)(ts_read + 1) : tw_prev) || (range >> 32)) bad |= ZKC_SHV_PARAMS;
    }
    const uint64_t reset = TR(ZKC_SH_RESET_BUFFER), should_read = TR(ZKC_SH_SHOULD_READ);
    const uint64_t rwr = rpc | rwr_in;  // read_words_for_round after :205-208
    if ((reset | should_read) > 1 || reset != (rpc | completed_in) || should_read != (uint64_t)(p[4] != 0)) bad |= ZKC_SHV_FSM;
    // the two reads: offsets, message words, the memory queue (tail / length chain: previous write -> read 0 -> read 1 -> write)
    uint32_t m[16];
    uint64_t mt_prev[12], ml_prev;
#pragma unroll
    for (int i = 0; i < 12; i++) mt_prev[i] = first ? d->mq0.tail[i] : TP(ZKC_SH_WRITE_TAIL + i);
    ml_prev = first ? d->mq0.length : TP(ZKC_SH_WRITE_LEN);
    uint64_t offset = p[1];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int b = ZKC_SH_QUERY + q * ZKC_SH_QUERY_STRIDE;
        uint32_t value[8];
        uint64_t vr = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { const uint64_t v = TR(b + i); vr |= v; value[i] = (uint32_t)v; }
        if ((vr >> 32) || (!should_read && vr)) bad |= ZKC_SHV_BOOLEAN;  // conditionally_allocate: zero when nothing is read
        const uint64_t off_after = TR(b + 21);
        if (off_after != (uint64_t)(uint32_t)(offset + rwr)) bad |= ZKC_SHV_PARAMS;
#pragma unroll
        for (int i = 0; i < 8; i++) m[8 * q + i] = value[7 - i];
        uint64_t mt[12], st[12];
        bool msame = true;
#pragma unroll
        for (int i = 0; i < 12; i++) { mt[i] = TR(b + 8 + i); msame &= mt[i] == mt_prev[i]; if (mt[i] >= GL_P) bad |= ZKC_SHV_BOOLEAN; }
        const uint64_t ml = TR(b + 20);
        if (ml != ml_prev + should_read || (!should_read && !msame)) bad |= ZKC_SHV_MEMORY_QUEUE;
        if (ROUND_FUNCTION && should_read) {
            mq_encode((uint32_t)ts_read, (uint32_t)p[0], (uint32_t)offset, 0, value, st);
#pragma unroll
            for (int i = 8; i < 12; i++) st[i] = mt_prev[i];
            poseidon2_permute(st);
#pragma unroll
            for (int i = 0; i < 12; i++) if (st[i] != mt[i]) bad |= ZKC_SHV_ROUND_FUNCTION;
        }
#pragma unroll
        for (int i = 0; i < 12; i++) mt_prev[i] = mt[i];
        ml_prev = ml;
        offset = off_after;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) if (TR(ZKC_SH_MESSAGE + i) != m[i]) bad |= ZKC_SHV_COMPRESSION;
    const uint64_t rounds_after = TR(ZKC_SH_NUM_ROUNDS);
    if (rounds_after != (uint64_t)(uint32_t)(p[4] - rwr)) bad |= ZKC_SHV_PARAMS;
    // the compression
    uint32_t cur[8], result[8];
    {
        uint64_t range = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint64_t si = TR(ZKC_SH_STATE_IN + i);
            range |= si;
            if (si != (reset ? (uint64_t)SHA_IV[i] : (first ? (uint64_t)s0.sha256_inner_state[i] : TP(ZKC_SH_STATE_OUT + i)))) bad |= ZKC_SHV_COMPRESSION;
            cur[i] = (uint32_t)si;
        }
        if (range >> 32) bad |= ZKC_SHV_BOOLEAN;
        sha256_compress(cur, m);
#pragma unroll
        for (int i = 0; i < 8; i++) if (TR(ZKC_SH_STATE_OUT + i) != cur[i]) bad |= ZKC_SHV_COMPRESSION;
#pragma unroll
        for (int k = 0; k < 8; k++) result[7 - k] = cur[k];
#pragma unroll
        for (int i = 0; i < 8; i++) if (TR(ZKC_SH_RESULT + i) != result[i]) bad |= ZKC_SHV_COMPRESSION;
    }
    // the write and the next FSM flags
    const uint64_t write_result = TR(ZKC_SH_WRITE_RESULT);
    if (write_result != (rwr & (uint64_t)(rounds_after == 0))) bad |= ZKC_SHV_FSM;
    {
        uint64_t mt[12], st[12];
        bool msame = true;
#pragma unroll
        for (int i = 0; i < 12; i++) { mt[i] = TR(ZKC_SH_WRITE_TAIL + i); msame &= mt[i] == mt_prev[i]; if (mt[i] >= GL_P) bad |= ZKC_SHV_BOOLEAN; }
        if (TR(ZKC_SH_WRITE_LEN) != ml_prev + write_result || (!write_result && !msame)) bad |= ZKC_SHV_MEMORY_QUEUE;
        if (ROUND_FUNCTION && write_result) {
            mq_encode((uint32_t)ts_write, (uint32_t)p[2], (uint32_t)p[3], 1, result, st);
#pragma unroll
            for (int i = 8; i < 12; i++) st[i] = mt_prev[i];
            poseidon2_permute(st);
#pragma unroll
            for (int i = 0; i < 12; i++) if (st[i] != mt[i]) bad |= ZKC_SHV_ROUND_FUNCTION;
        }
    }
    {
        const uint64_t empty = len == 0;
        const uint64_t o_rpc = TR(ZKC_SH_FLAGS_OUT + 0), o_rwr = TR(ZKC_SH_FLAGS_OUT + 1), o_completed = TR(ZKC_SH_FLAGS_OUT + 2);
        if (o_rpc != (write_result & (1 - empty)) || o_completed != ((write_result & empty) | completed_in) || o_rwr != 1 - ((o_rpc | o_completed) & 1) ||
            (o_rpc | o_rwr | o_completed) > 1)
            bad |= ZKC_SHV_FSM;
    }
#undef TR
#undef TP
    if (bad) {
        atomicAdd(violations, 1ull);
        atomicOr(&d->failed_checks, bad);
        atomicMin(&d->first_bad, ((unsigned long long)row << 16) | bad);
    }
}

}  // namespace zkc

using namespace zkc;

extern "C" int zkc_sha256_round_function_entry_point(zkc_ctx *ctx, zkc_sha256_closed_form *io, const zkc_log_query *requests,
                                                     const uint64_t *requests_prev_tails, size_t n_requests,
                                                     const uint32_t *memory_reads, size_t n_reads, const uint64_t *memory_states,
                                                     size_t n_memory_states, size_t limit, const zkc_precompile_options *options,
                                                     int on_device, uint64_t *trace, uint64_t commitment[ZKC_COMMITMENT_LEN],
                                                     zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !commitment || (n_requests && (!requests || !requests_prev_tails)) || (n_reads && !memory_reads) ||
        limit > 0x0FFFFFFFull) {
        status->code = ZKC_ERR_INVALID_ARGUMENT;
        return ZKC_ERR_INVALID_ARGUMENT;
    }
    const bool in_dev = on_device & ZKC_INPUTS_ON_DEVICE, trace_dev = on_device & ZKC_TRACE_ON_DEVICE;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const size_t max_units = n_requests + 1;
    const size_t tiles = (max_units + SCAN_THREADS - 1) / SCAN_THREADS;
    const bool have_states = memory_states != nullptr;
    const size_t max_pushes = 3 * limit + 1;
    if (!have_states) n_memory_states = max_pushes;
    size_t bytes = zkc_carver::bytes(1, sizeof(ShDev)) + zkc_carver::bytes(1, sizeof(ScanGlobal)) +
                   zkc_carver::bytes(tiles + 1, sizeof(TileStateT<ShPlan>)) + zkc_carver::bytes(max_units + 2, sizeof(ShPlan)) +
                   zkc_carver::bytes(max_pushes * 8, 8) + zkc_carver::bytes(3 * limit + 8, 4);
    if (!in_dev) bytes += zkc_carver::bytes(n_requests + 1, sizeof(zkc_log_query)) + zkc_carver::bytes(n_requests * 4 + 4, 8) +
                          zkc_carver::bytes(n_reads * 8 + 8, 4);
    if (!in_dev || !have_states) bytes += zkc_carver::bytes(n_memory_states * 12 + 12, 8);
    if (trace && !trace_dev) bytes += zkc_carver::bytes((size_t)ZKC_SH_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    ShDev *h = (ShDev *)ctx->pinned(sizeof(ShDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    ShDev *d = cv.take<ShDev>(1);
    char *zero_begin = cv.base + cv.off;
    ScanGlobal *sg = cv.take<ScanGlobal>(1);
    TileStateT<ShPlan> *ts = cv.take<TileStateT<ShPlan>>(tiles + 1);
    char *zero_end = cv.base + cv.off;
    ShPlan *starts = cv.take<ShPlan>(max_units + 2);
    uint64_t *push_enc = cv.take<uint64_t>(max_pushes * 8);
    uint32_t *slot_meta = cv.take<uint32_t>(3 * limit + 8);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(ShDev));
    h->io = *io;
    if (options) h->opt = *options;
    h->n_requests = n_requests; h->n_reads = n_reads; h->n_memory_states = n_memory_states; h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(ShDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(zero_begin, 0, zero_end - zero_begin, s));
    const zkc_log_query *dreq = requests;
    const uint64_t *dprev = requests_prev_tails, *dstates = memory_states;
    const uint32_t *dreads = memory_reads;
    uint64_t *dtrace = trace;
    if (!in_dev) {
        zkc_log_query *br = cv.take<zkc_log_query>(n_requests + 1);
        uint64_t *bp = cv.take<uint64_t>(n_requests * 4 + 4);
        uint32_t *bm = cv.take<uint32_t>(n_reads * 8 + 8);
        if (n_requests) {
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(br, requests, n_requests * sizeof(zkc_log_query), cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bp, requests_prev_tails, n_requests * 32, cudaMemcpyHostToDevice, s));
        }
        if (n_reads) ZKC_CUDA(ctx, status, cudaMemcpyAsync(bm, memory_reads, n_reads * 32, cudaMemcpyHostToDevice, s));
        dreq = br; dprev = bp; dreads = bm;
    }
    if (!in_dev || !have_states) {
        uint64_t *bs = cv.take<uint64_t>(n_memory_states * 12 + 12);
        if (have_states && n_memory_states)
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bs, memory_states, n_memory_states * 96, cudaMemcpyHostToDevice, s));
        dstates = bs;
    }
    if (trace && !trace_dev) dtrace = cv.take<uint64_t>((size_t)ZKC_SH_NUM_COLS * limit);

    ZKC_LAUNCH(ctx, "sh_prologue", sh_prologue_kernel, 1, 96, 0, d);
    ZKC_LAUNCH(ctx, "sh_plan", sh_plan_kernel, (unsigned)tiles, SCAN_THREADS, 0, d, dreq, starts, sg, ts);
    if (limit) {
        ZKC_LAUNCH(ctx, "sh_calls", sh_calls_kernel, (unsigned)((max_units + 127) / 128), 128, 0, d, dreq, dprev, dreads, starts,
                   push_enc, slot_meta, dtrace);
        ZKC_LAUNCH(ctx, "sh_tail", sh_tail_kernel, (unsigned)((limit + 127) / 128), 128, 0, d, starts, slot_meta, dtrace, push_enc);
        if (!have_states) ZKC_LAUNCH(ctx, "sh_mem_chain", (pc_mem_chain_kernel<ShDev, 3>), 1, 32, 0, d, push_enc, slot_meta, (uint64_t *)dstates);
        ZKC_LAUNCH(ctx, "sh_memq", (pc_memq_kernel<ShDev, 3, ZKC_SH_QUERY + 8, ZKC_SH_QUERY_STRIDE, ZKC_SH_WRITE_TAIL, ZKC_KC_CHK_QUEUE_HINT>),
                   (unsigned)((3 * limit + 255) / 256), 256, 0, d, push_enc, slot_meta, dstates, n_memory_states, have_states, dtrace);
    }
    ZKC_LAUNCH(ctx, "sh_finalize", sh_finalize_kernel, 1, 32, 0, d, slot_meta, dstates, n_memory_states);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(ShDev), cudaMemcpyDeviceToHost, s));
    if (!trace_dev && trace && limit)
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(trace, dtrace, (size_t)ZKC_SH_NUM_COLS * limit * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    io->hidden_fsm_output = h->io.hidden_fsm_output;
    io->final_memory_state = h->io.final_memory_state;
    io->completion_flag = h->io.completion_flag;
    memcpy(commitment, h->commitment, 32);
    *status = h->status;
    return status->code;
}

extern "C" int zkc_sha256_round_function_check_trace(zkc_ctx *ctx, const zkc_sha256_closed_form *io, const zkc_precompile_options *options,
                                                     const uint64_t *trace, size_t limit, uint32_t gates, int on_device, uint64_t *violations,
                                                     zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !violations || (limit && !trace)) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    size_t bytes = zkc_carver::bytes(1, sizeof(ShDev)) + zkc_carver::bytes(1, 8);
    if (!on_device) bytes += zkc_carver::bytes((size_t)ZKC_SH_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    ShDev *h = (ShDev *)ctx->pinned(sizeof(ShDev) + 8);
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    ShDev *d = cv.take<ShDev>(1);
    unsigned long long *dviol = cv.take<unsigned long long>(1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(ShDev));
    h->io = *io;
    if (options) h->opt = *options;
    h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(ShDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(dviol, 0, 8, s));
    const uint64_t *dt = trace;
    if (!on_device && limit) {
        uint64_t *b = cv.take<uint64_t>((size_t)ZKC_SH_NUM_COLS * limit);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(b, trace, (size_t)ZKC_SH_NUM_COLS * limit * 8, cudaMemcpyHostToDevice, s));
        dt = b;
    }
    ZKC_LAUNCH(ctx, "sh_prologue", sh_prologue_kernel, 1, 96, 0, d);
    if (limit) {
        const unsigned grid = (unsigned)((limit + 127) / 128);
        if (gates == 0 || (gates & ZKC_GATES_ROUND_FUNCTION)) ZKC_LAUNCH(ctx, "sh_check_rf", sh_check_kernel<true>, grid, 128, 0, d, dviol, dt);
        else ZKC_LAUNCH(ctx, "sh_check", sh_check_kernel<false>, grid, 128, 0, d, dviol, dt);
    }
    ZKC_CUDA(ctx, status, cudaGetLastError());
    unsigned long long *hviol = (unsigned long long *)(h + 1);
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(ShDev), cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(hviol, dviol, 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    *violations = *hviol;
    status->failed_checks = h->failed_checks;
    if (*hviol) {
        status->code = ZKC_ERR_UNSATISFIED;
        status->first_bad_row = (int64_t)(h->first_bad >> 16);
    }
    return status->code;
}
