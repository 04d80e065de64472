// keccak256 precompile circuit on sm_100a: keccak256_round_function_entry_point
// (/root/reference/src/keccak256_round_function/mod.rs:673-794) and its work cycle
// keccak256_precompile_inner (:155-670).
//
// The reference runs one FSM cycle after another.  A precompile CALL, however, starts from a reset
// buffer / sponge state (:321, :361-375), so the cycles partition into independent calls:
//   plan    one thread per call dry-runs the control part of the FSM (how many cycles, memory reads and
//           memory-queue pushes the call takes: a function of its length and unalignment only);
//   scan    exclusive prefix sums place every call in the cycle / witness / push order;
//   calls   one thread per call walks its cycles: unaligned reads into the 192-byte shift buffer,
//           padding, keccak-f[1600] in registers, digest write; rows beyond the last call are the
//           constant "completed" cycle (tail kernel);
//   memq    the memory queue (<= 7 conditional pushes per cycle, a Poseidon2 hash chain) is verified
//           per push against host-supplied tails or rebuilt by a sequential chain kernel;
//   finalize  FSM output, observable output, commitment.
#include "ctx.cuh"
#include "log_query.cuh"
#include "precompile_common.cuh"
#include "keccak_f1600.cuh"
#include "scan.cuh"

namespace zkc {

struct KcDev {
    zkc_keccak_closed_form io;
    zkc_precompile_options opt;
    uint64_t n_requests, n_reads, n_memory_states, limit;
    // prologue
    uint32_t start, prologue_checks, n_units, unit0_fresh;
    zkc_keccak_fsm s0;   // FSM state on entry to cycle 0 (after the start select and the can_finish_immediately masking)
    zkc_queue_state4 rq0;
    zkc_queue_state12 mq0;
    uint64_t commit_obs_in[4], commit_fsm_in[4];
    // calls / tail
    zkc_keccak_fsm s_last;   // FSM state after the last executed call cycle (input of the tail rows)
    zkc_keccak_fsm s_final;  // FSM state after row limit - 1
    uint32_t total_cycles, total_reads, total_pushes, popped_requests;
    uint64_t req_head_final[4];
    // status
    unsigned long long first_bad;
    uint32_t failed_checks, hint_bad;
    uint64_t commitment[4];
    zkc_status status;
};

struct KcPlan {
    uint32_t cycles, reads, pushes, pad;
};
struct KcPlanOp {
    static __device__ __forceinline__ KcPlan identity() { return KcPlan{0, 0, 0, 0}; }
    static __device__ __forceinline__ KcPlan combine(const KcPlan &a, const KcPlan &b) {
        return KcPlan{a.cycles + b.cycles, a.reads + b.reads, a.pushes + b.pushes, 0};
    }
};

// per-cycle working state of one call
struct KcState {
    uint32_t read_precompile_call, read_unaligned, padding_round, completed;
    uint32_t ts_read, ts_write;
    uint32_t input_page, byte_offset, byte_length, output_page, output_word_offset, needs_full_padding;
    uint32_t filled;
    uint8_t buffer[ZKC_KECCAK_BUFFER_SIZE];
    uint8_t sponge[200];  // [i][j][byte]
};

__device__ void kc_load_state(const zkc_keccak_fsm &f, KcState &s) {
    s.read_precompile_call = f.read_precompile_call; s.read_unaligned = f.read_unaligned_words_for_round;
    s.padding_round = f.padding_round; s.completed = f.completed;
    s.ts_read = f.timestamp_to_use_for_read; s.ts_write = f.timestamp_to_use_for_write;
    s.input_page = f.input_page; s.byte_offset = f.input_memory_byte_offset; s.byte_length = f.input_memory_byte_length;
    s.output_page = f.output_page; s.output_word_offset = f.output_word_offset; s.needs_full_padding = f.needs_full_padding_round;
    s.filled = f.buffer_filled;
    for (int i = 0; i < ZKC_KECCAK_BUFFER_SIZE; i++) s.buffer[i] = f.buffer_bytes[i];
    for (int i = 0; i < 200; i++) s.sponge[i] = f.keccak_internal_state[i];
}
__device__ void kc_store_state(const KcState &s, zkc_keccak_fsm &f) {
    f.read_precompile_call = s.read_precompile_call; f.read_unaligned_words_for_round = s.read_unaligned;
    f.padding_round = s.padding_round; f.completed = s.completed;
    f.timestamp_to_use_for_read = s.ts_read; f.timestamp_to_use_for_write = s.ts_write;
    f.input_page = s.input_page; f.input_memory_byte_offset = s.byte_offset; f.input_memory_byte_length = s.byte_length;
    f.output_page = s.output_page; f.output_word_offset = s.output_word_offset; f.needs_full_padding_round = s.needs_full_padding;
    f.buffer_filled = s.filled; f._pad = 0;
    for (int i = 0; i < ZKC_KECCAK_BUFFER_SIZE; i++) f.buffer_bytes[i] = s.buffer[i];
    for (int i = 0; i < 200; i++) f.keccak_internal_state[i] = s.sponge[i];
}

// what one cycle did to the outside world
struct KcCycleOut {
    uint32_t reads, pushes;
    bool write_result;
};

// relation family of a trace column, for the constraint evaluator (ZKC_KCV_* bits)
__device__ __forceinline__ uint32_t kc_col_family(int col) {
    if (col < ZKC_KC_CALL_ITEM || (col >= ZKC_KC_RESET_BUFFER && col < ZKC_KC_QUERY) || (col >= ZKC_KC_ZERO_BYTES_LEFT && col < ZKC_KC_INPUT) ||
        col == ZKC_KC_WRITE_RESULT || (col >= ZKC_KC_FLAGS_OUT && col < ZKC_KC_BUFFER_OUT))
        return ZKC_KCV_FSM;
    if (col >= ZKC_KC_PARAMS && col < ZKC_KC_RESET_BUFFER) return ZKC_KCV_PARAMS;
    if (col >= ZKC_KC_QUERY && col < ZKC_KC_ZERO_BYTES_LEFT) return ZKC_KCV_QUERIES;
    if (col >= ZKC_KC_BUFFER_OUT) return ZKC_KCV_BUFFER;
    return ZKC_KCV_SPONGE;  // INPUT, STATE_OUT, RESULT
}

// One iteration of the main work cycle (mod.rs:230-665).  DRY: control only (no witness, no sponge, no trace).
// `call` is the precompile call popped this cycle (zero item when none is popped); cursors advance.
// CHECK: constraint evaluation -- `trace` is a finished trace: the words read from memory are taken from its QUERY columns and
// every cell the cycle would write is COMPARED with the trace instead (a mismatch sets the cell's ZKC_KCV_* family bit in `checks`).
template <bool DRY, bool CHECK = false>
__device__ __forceinline__ KcCycleOut kc_cycle(KcState &s, const zkc_log_query &call, bool input_queue_empty_after_pop,
                                               const uint32_t *__restrict__ reads, size_t n_reads, size_t &read_cursor,
                                               uint64_t *__restrict__ push_enc, uint32_t *__restrict__ slot_meta, uint32_t &push_ordinal,
                                               uint64_t *__restrict__ trace, size_t limit, size_t row, uint32_t &checks) {
    KcCycleOut out{0, 0, false};
    const bool wr = !DRY && trace != nullptr;
    auto put = [&](int col, uint64_t v) {
        if constexpr (CHECK) { if (__ldg(trace + (size_t)col * limit + row) != v) checks |= kc_col_family(col); }
        else trace[(size_t)col * limit + row] = v;
    };
    const bool read_call = s.read_precompile_call;
    if (wr) {
        put(ZKC_KC_FLAGS_IN + 0, s.read_precompile_call); put(ZKC_KC_FLAGS_IN + 1, s.read_unaligned);
        put(ZKC_KC_FLAGS_IN + 2, s.padding_round); put(ZKC_KC_FLAGS_IN + 3, s.completed);
    }
    const uint32_t new_len = call.key[1];
    if (read_call) {
        s.byte_offset = call.key[0]; s.byte_length = call.key[1]; s.output_word_offset = call.key[2];
        s.input_page = call.key[4]; s.output_page = call.key[5];
        s.needs_full_padding = (call.key[1] % ZKC_KECCAK_RATE_BYTES) == 0;
        s.ts_read = call.timestamp; s.ts_write = call.timestamp + 1;
    }
    const bool reset_buffer = read_call || s.completed;
    const bool read_zero = read_call && new_len == 0, read_nonzero = read_call && new_len != 0;
    s.read_precompile_call = 0;
    s.read_unaligned = s.read_unaligned || read_nonzero;
    s.padding_round = s.padding_round || read_zero;
    if (reset_buffer) {
        s.filled = 0;
        if (!DRY) {
            for (int i = 0; i < ZKC_KECCAK_BUFFER_SIZE; i++) s.buffer[i] = 0;
            for (int i = 0; i < 200; i++) s.sponge[i] = 0;
        }
    }
    if (wr) {
        put(ZKC_KC_PARAMS + 0, s.input_page); put(ZKC_KC_PARAMS + 1, s.byte_offset); put(ZKC_KC_PARAMS + 2, s.byte_length);
        put(ZKC_KC_PARAMS + 3, s.output_page); put(ZKC_KC_PARAMS + 4, s.output_word_offset); put(ZKC_KC_PARAMS + 5, s.needs_full_padding);
        put(ZKC_KC_TS_READ, s.ts_read); put(ZKC_KC_TS_WRITE, s.ts_write);
        put(ZKC_KC_RESET_BUFFER, reset_buffer); put(ZKC_KC_READ_ZERO_LENGTH, read_zero); put(ZKC_KC_READ_NON_ZERO_LENGTH, read_nonzero);
    }
#pragma unroll 1
    for (int q = 0; q < ZKC_KECCAK_MEMORY_QUERIES_PER_CYCLE; q++) {
        const uint32_t aligned = s.byte_offset / 32, unal = s.byte_offset % 32, at_most = 32 - unal;
        const uint32_t meaningful = s.byte_length < at_most ? s.byte_length : at_most;
        const uint32_t next_filled = s.filled + meaningful;
        if (next_filled > 255) checks |= CHECK ? ZKC_KCV_ENFORCE : ZKC_KC_CHK_BUFFER_OVERFLOW;
        const bool should_read = meaningful != 0 && next_filled <= ZKC_KECCAK_BUFFER_SIZE && s.read_unaligned;
        uint32_t value[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (should_read) {
            if (!DRY) {
                if constexpr (CHECK) {
#pragma unroll
                    for (int i = 0; i < 8; i++) value[i] = (uint32_t)__ldg(trace + (size_t)(ZKC_KC_QUERY + q * ZKC_KC_QUERY_STRIDE + 4 + i) * limit + row);
                } else {
                    if (read_cursor < n_reads) {
#pragma unroll
                        for (int i = 0; i < 8; i++) value[i] = __ldg(reads + 8 * read_cursor + i);
                    } else checks |= ZKC_KC_CHK_WITNESS_EXHAUSTED;
                    uint64_t e[8];
                    mq_encode(s.ts_read, s.input_page, aligned, 0, value, e);
#pragma unroll
                    for (int i = 0; i < 8; i++) push_enc[8 * (size_t)push_ordinal + i] = e[i];
                }
                // fill_with_bytes(be_bytes, offset = unal, meaningful), buffer/mod.rs:73-136
                for (uint32_t idx = 0; idx < 32; idx++) {
                    const uint32_t pos = s.filled + idx;
                    if (pos >= ZKC_KECCAK_BUFFER_SIZE) break;
                    const uint32_t k = unal + idx;  // big-endian byte k of the word
                    s.buffer[pos] = (idx < meaningful && k < 32) ? (uint8_t)(value[7 - k / 4] >> (8 * (3 - k % 4))) : 0;
                }
            }
            read_cursor++; push_ordinal++; out.reads++; out.pushes++;
            s.byte_offset += meaningful; s.byte_length -= meaningful; s.filled += meaningful;
        }
        if (!DRY && !CHECK) slot_meta[7 * row + q] = push_ordinal | (should_read ? 0x80000000u : 0u);
        if (wr) {
            const int b = ZKC_KC_QUERY + q * ZKC_KC_QUERY_STRIDE;
            put(b + 0, aligned); put(b + 1, unal); put(b + 2, meaningful); put(b + 3, should_read);
#pragma unroll
            for (int i = 0; i < 8; i++) put(b + 4 + i, value[i]);
            put(b + 25, s.byte_offset); put(b + 26, s.byte_length); put(b + 27, s.filled);
        }
    }
    const bool zero_bytes_left = s.byte_length == 0;
    const uint32_t currently_filled = s.filled;
    const bool do_one_byte = currently_filled == ZKC_KECCAK_RATE_BYTES - 1;
    s.filled = s.filled < ZKC_KECCAK_RATE_BYTES ? 0 : s.filled - ZKC_KECCAK_RATE_BYTES;  // consume::<136>(allow_partial)
    const bool buffer_now_empty = s.filled == 0;
    const bool apply_padding = zero_bytes_left && buffer_now_empty && s.read_unaligned && !s.needs_full_padding;
    const bool write_result = apply_padding || s.padding_round;
    uint32_t result[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!DRY) {
        // block = first 136 buffer bytes with padding, XORed into the sponge lanes (mod.rs:802-817)
        uint64_t A[25];
#pragma unroll 1
        for (int idx = 0; idx < 25; idx++) {
            const int i = idx % 5, j = idx / 5;
            uint64_t lane = 0;
            for (int b = 0; b < 8; b++) lane |= (uint64_t)s.sponge[(i * 5 + j) * 8 + b] << (8 * b);
            if (idx < 17) {
                for (int b = 0; b < 8; b++) {
                    const int k = 8 * idx + b;
                    uint8_t byte = s.buffer[k];
                    if (apply_padding) {
                        if (k < 135 && (uint32_t)k == currently_filled) byte = 0x01;
                        if (k == 135) byte = do_one_byte ? 0x81 : 0x80;
                    }
                    if (s.padding_round) byte = k == 0 ? 0x01 : (k == 135 ? 0x80 : 0);
                    if (wr) put(ZKC_KC_INPUT + k, byte);
                    lane ^= (uint64_t)byte << (8 * b);
                }
            }
            A[idx] = lane;
        }
        for (int i = 0; i < ZKC_KECCAK_BUFFER_SIZE - ZKC_KECCAK_RATE_BYTES; i++) s.buffer[i] = s.buffer[i + ZKC_KECCAK_RATE_BYTES];
        for (int i = ZKC_KECCAK_BUFFER_SIZE - ZKC_KECCAK_RATE_BYTES; i < ZKC_KECCAK_BUFFER_SIZE; i++) s.buffer[i] = 0;
        keccak_f1600(A);
#pragma unroll 1
        for (int idx = 0; idx < 25; idx++) {
            const int i = idx % 5, j = idx / 5;
            for (int b = 0; b < 8; b++) {
                const uint8_t byte = (uint8_t)(A[idx] >> (8 * b));
                s.sponge[(i * 5 + j) * 8 + b] = byte;
                if (wr) put(ZKC_KC_STATE_OUT + (i * 5 + j) * 8 + b, byte);
            }
        }
        // UInt256::from_be_bytes(state[0..4][0]): digest byte d = lane d/8, byte d%8; limb l = BE bytes 28-4l..31-4l
#pragma unroll
        for (int l = 0; l < 8; l++) {
            uint32_t w = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int dgt = 28 - 4 * l + t;
                w = (w << 8) | (uint32_t)((A[dgt / 8] >> (8 * (dgt % 8))) & 0xFF);
            }
            result[l] = w;
        }
        if (write_result && !CHECK) {
            uint64_t e[8];
            mq_encode(s.ts_write, s.output_page, s.output_word_offset, 1, result, e);
#pragma unroll
            for (int i = 0; i < 8; i++) push_enc[8 * (size_t)push_ordinal + i] = e[i];
        }
    }
    if (write_result) { push_ordinal++; out.pushes++; }
    if (!DRY && !CHECK) slot_meta[7 * row + 6] = push_ordinal | (write_result ? 0x80000000u : 0u);
    out.write_result = write_result;
    const bool nothing_left = write_result && input_queue_empty_after_pop, process_next = write_result && !input_queue_empty_after_pop;
    s.read_precompile_call = process_next;
    s.completed = s.completed || nothing_left;
    s.padding_round = s.read_unaligned && zero_bytes_left && buffer_now_empty && s.needs_full_padding;
    s.read_unaligned = !(s.read_precompile_call || s.padding_round || s.completed);
    if (wr) {
        put(ZKC_KC_ZERO_BYTES_LEFT, zero_bytes_left); put(ZKC_KC_CURRENTLY_FILLED, currently_filled);
        put(ZKC_KC_DO_ONE_BYTE_OF_PADDING, do_one_byte); put(ZKC_KC_BUFFER_NOW_EMPTY, buffer_now_empty);
        put(ZKC_KC_APPLY_PADDING, apply_padding); put(ZKC_KC_WRITE_RESULT, write_result);
#pragma unroll
        for (int i = 0; i < 8; i++) put(ZKC_KC_RESULT + i, result[i]);
        put(ZKC_KC_FLAGS_OUT + 0, s.read_precompile_call); put(ZKC_KC_FLAGS_OUT + 1, s.read_unaligned);
        put(ZKC_KC_FLAGS_OUT + 2, s.padding_round); put(ZKC_KC_FLAGS_OUT + 3, s.completed);
        for (int i = 0; i < ZKC_KECCAK_BUFFER_SIZE; i++) put(ZKC_KC_BUFFER_OUT + i, s.buffer[i]);
    }
    return out;
}

__device__ int kc_encode_fsm(const zkc_keccak_fsm &f, uint64_t *dst) {
    int n = 0;
    dst[n++] = f.read_precompile_call; dst[n++] = f.read_unaligned_words_for_round; dst[n++] = f.padding_round; dst[n++] = f.completed;
    for (int i = 0; i < 200; i++) dst[n++] = f.keccak_internal_state[i];
    dst[n++] = f.timestamp_to_use_for_read; dst[n++] = f.timestamp_to_use_for_write;
    dst[n++] = f.input_page; dst[n++] = f.input_memory_byte_offset; dst[n++] = f.input_memory_byte_length;
    dst[n++] = f.output_page; dst[n++] = f.output_word_offset; dst[n++] = f.needs_full_padding_round;
    for (int i = 0; i < 192; i++) dst[n++] = f.buffer_bytes[i];
    dst[n++] = f.buffer_filled;
    n += put_queue_state4(dst + n, f.log_queue_state);
    for (int i = 0; i < 12; i++) dst[n++] = f.memory_queue_state.head[i];
    for (int i = 0; i < 12; i++) dst[n++] = f.memory_queue_state.tail[i];
    dst[n++] = f.memory_queue_state.length;
    return n;  // 439
}
__device__ int put_q12(uint64_t *dst, const zkc_queue_state12 &s) {
    for (int i = 0; i < 12; i++) dst[i] = s.head[i];
    for (int i = 0; i < 12; i++) dst[12 + i] = s.tail[i];
    dst[24] = s.length;
    return 25;
}

__global__ void kc_prologue_kernel(KcDev *d) {
    // warp 0: scalar start selection; warps 1 / 2: the two input commitments, every permutation on 12 cooperating lanes
    __shared__ uint64_t buf[2][440];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane >= 16 || (warp == 0 && lane != 0)) return;
    const unsigned gm = 0xFFFFu;
    const zkc_keccak_closed_form &io = d->io;
    if (warp == 0) {
        const bool start = io.start_flag != 0;
        d->start = start;
        d->rq0 = start ? io.initial_log_queue_state : io.hidden_fsm_input.log_queue_state;
        d->mq0 = start ? io.initial_memory_queue_state : io.hidden_fsm_input.memory_queue_state;
        zkc_keccak_fsm s;
        if (start) { memset(&s, 0, sizeof s); s.read_precompile_call = 1; }  // :733-741
        else s = io.hidden_fsm_input;
        const bool cfi = s.read_precompile_call && d->rq0.length == 0;       // :196-213
        if (cfi) { s.read_precompile_call = 0; s.read_unaligned_words_for_round = 0; s.completed = 1; }
        d->s0 = s;
        d->s_last = s;
        d->s_final = s;
        d->popped_requests = 0;
        uint32_t checks = 0;
        for (int i = 0; i < 4; i++) if (io.initial_log_queue_state.head[i]) checks |= ZKC_KC_CHK_TRIVIAL_HEAD;
        for (int i = 0; i < 12; i++) if (io.initial_memory_queue_state.head[i]) checks |= ZKC_KC_CHK_TRIVIAL_HEAD;
        d->prologue_checks = checks;
        // unit 0 = the call in progress on entry (empty if the FSM is about to pop, or already completed)
        d->unit0_fresh = (s.read_precompile_call || s.completed) ? 1 : 0;
        const uint64_t avail = d->rq0.length < d->n_requests ? d->rq0.length : d->n_requests;
        d->n_units = 1 + (s.completed ? 0 : (uint32_t)avail);
    } else {
        uint64_t *b = buf[warp - 1];
        int n = 0;
        if (lane == 0) {
            if (warp == 1) {
                n = put_queue_state4(b, io.initial_log_queue_state);
                n += put_q12(b + n, io.initial_memory_queue_state);
            } else {
                n = kc_encode_fsm(io.hidden_fsm_input, b);
            }
        }
        __syncwarp(gm);
        n = __shfl_sync(gm, n, 0, 16);
        const uint64_t c = commit_encoding_coop(gm, b, n, lane);
        if (lane < 4) (warp == 1 ? d->commit_obs_in : d->commit_fsm_in)[lane] = c;
    }
}

// ---- plan: dry run of every call ------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS)
kc_plan_kernel(KcDev *d, const zkc_log_query *__restrict__ requests, KcPlan *__restrict__ starts, ScanGlobal *sg,
               TileStateT<KcPlan> *tiles) {
    __shared__ ScanSharedT<KcPlan> sh;
    const unsigned int tile = scan_take_ticket(sg, sh);
    const size_t u = (size_t)tile * SCAN_THREADS + threadIdx.x;
    const uint32_t n_units = d->n_units;
    KcPlan v = KcPlanOp::identity();
    if (u < n_units) {
        KcState s;
        kc_load_state(d->s0, s);
        zkc_log_query call = lq_zero();
        bool run = true;
        if (u == 0) {
            run = !d->unit0_fresh;
        } else {
            call = lq_load(requests + (u - 1));
            s.read_precompile_call = 1; s.read_unaligned = 0; s.padding_round = 0; s.completed = 0;
        }
        size_t rc = 0;
        uint32_t po = 0, checks = 0;
        const size_t limit = d->limit;
        while (run && v.cycles < limit) {
            const KcCycleOut o = kc_cycle<true>(s, call, false, nullptr, 0, rc, nullptr, nullptr, po, nullptr, 0, 0, checks);
            v.cycles++; v.reads += o.reads; v.pushes += o.pushes;
            call = lq_zero();
            if (o.write_result) break;
        }
    }
    KcPlan incl;
    const KcPlan excl = scan_tile_generic<KcPlan, KcPlanOp>(v, tile, KcPlanOp::identity(), tiles, sh, incl);
    if (u < n_units) starts[u] = excl;
    if (u + 1 == n_units) starts[n_units] = incl;
}

__device__ __forceinline__ void kc_report(KcDev *d, size_t row, uint32_t checks) {
    if (!checks) return;
    atomicOr(&d->failed_checks, checks);
    atomicMin(&d->first_bad, ((unsigned long long)row << 16) | checks);
}

// ---- calls: one thread per call ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
kc_calls_kernel(KcDev *d, const zkc_log_query *__restrict__ requests, const uint64_t *__restrict__ req_prev,
                const uint32_t *__restrict__ reads, const KcPlan *__restrict__ starts, uint64_t *__restrict__ push_enc,
                uint32_t *__restrict__ slot_meta, uint64_t *__restrict__ trace) {
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_units = d->n_units;
    if (u >= n_units) return;
    const size_t limit = d->limit;
    const KcPlan st = starts[u], en = starts[u + 1];
    size_t row = st.cycles;
    if (row >= limit || en.cycles == st.cycles) return;  // nothing of this call falls inside [0, limit)
    KcState s;
    kc_load_state(d->s0, s);
    zkc_log_query call = lq_zero();
    const uint32_t aux_byte = d->opt.aux_byte ? d->opt.aux_byte : ZKC_PRECOMPILE_AUX_BYTE_DEFAULT;
    const uint32_t formal = d->opt.precompile_address ? d->opt.precompile_address : ZKC_KECCAK256_PRECOMPILE_ADDRESS_DEFAULT;
    const uint32_t rq_len0 = d->rq0.length;
    uint64_t head[4];
    uint32_t len_after;
    uint32_t checks = 0;
    if (u == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) head[i] = d->rq0.head[i];
        len_after = rq_len0;
    } else {
        call = lq_load(requests + (u - 1));
        s.read_precompile_call = 1; s.read_unaligned = 0; s.padding_round = 0; s.completed = 0;
        // carried FSM fields a fresh call overwrites at its first cycle anyway; keep the reference's values for the
        // ones it does not: none (params, timestamps are all selected on read_precompile_call)
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
        const bool next_exists = u + 1 < n_units && starts[u + 1].cycles < limit;
        if (next_exists) {
#pragma unroll
            for (int i = 0; i < 4; i++) hint_ok &= __ldg(req_prev + 4 * u + i) == head[i];
        }
        if (!hint_ok) { checks |= ZKC_KC_CHK_QUEUE_HINT; d->hint_bad = 1; }
        len_after = rq_len0 - (uint32_t)u;
        if (ZKC_LQ_AUX(call.flags) != aux_byte) checks |= ZKC_KC_CHK_AUX_BYTE;
        if (call.address[0] != formal || call.address[1] || call.address[2] || call.address[3] || call.address[4]) checks |= ZKC_KC_CHK_ADDRESS;
    }
    const bool queue_empty_after = len_after == 0;
    size_t read_cursor = st.reads;
    uint32_t push_ordinal = st.pushes;
    bool first_cycle = true;
    while (row < limit && row < en.cycles) {
        uint32_t cyc_checks = first_cycle ? checks : 0;
        if (trace) {
            for (int i = 0; i < 36; i++) trace[(size_t)(ZKC_KC_CALL_ITEM + i) * limit + row] = first_cycle ? lq_flat(call, i) : 0;
            for (int i = 0; i < 4; i++) trace[(size_t)(ZKC_KC_REQ_HEAD + i) * limit + row] = head[i];
            trace[(size_t)ZKC_KC_REQ_LEN * limit + row] = len_after;
        }
        kc_cycle<false>(s, first_cycle ? call : lq_zero(), queue_empty_after, reads, d->n_reads, read_cursor, push_enc, slot_meta,
                        push_ordinal, trace, limit, row, cyc_checks);
        kc_report(d, row, cyc_checks);
        first_cycle = false;
        row++;
    }
    // hand-over states
    const size_t total = starts[n_units].cycles;
    if (row == limit) kc_store_state(s, d->s_final);                 // this call executed row limit - 1
    if (en.cycles == total && row == en.cycles) kc_store_state(s, d->s_last);  // last call, ran to its end: input of the tail rows
    if (u >= 1 && (u + 1 == n_units || starts[u + 1].cycles >= limit)) {      // last call that was popped inside [0, limit)
#pragma unroll
        for (int i = 0; i < 4; i++) d->req_head_final[i] = head[i];
        d->popped_requests = (uint32_t)u;
    }
}

// ---- tail: rows after the last call (FSM completed): every row is the same constant cycle ----------------------------
__global__ void __launch_bounds__(128)
kc_tail_kernel(KcDev *d, const KcPlan *__restrict__ starts, uint32_t *__restrict__ slot_meta, uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t first = starts[d->n_units].cycles;
    const size_t row = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    KcState s;
    kc_load_state(d->s_last, s);
    // after the first tail row the carried sponge / buffer are the constant post-reset values; rows > first start
    // from exactly that state because the cycle resets before using them
    const uint32_t total_pushes = starts[d->n_units].pushes;
    size_t rc = 0;
    uint32_t po = total_pushes, checks = 0;
    const uint32_t popped = d->popped_requests;
    const uint32_t len_now = d->rq0.length - popped;
    if (s.read_precompile_call) checks |= ZKC_KC_CHK_WITNESS_EXHAUSTED;  // the FSM wants a call the witness does not hold
    if (trace) {
        for (int i = 0; i < 36; i++) trace[(size_t)(ZKC_KC_CALL_ITEM + i) * limit + row] = 0;
        for (int i = 0; i < 4; i++) trace[(size_t)(ZKC_KC_REQ_HEAD + i) * limit + row] = popped ? d->req_head_final[i] : d->rq0.head[i];
        trace[(size_t)ZKC_KC_REQ_LEN * limit + row] = len_now;
    }
    kc_cycle<false>(s, lq_zero(), len_now == 0, nullptr, 0, rc, nullptr, slot_meta, po, trace, limit, row, checks);
    kc_report(d, row, checks);
    if (row == limit - 1) kc_store_state(s, d->s_final);
}

// ---- finalize ----------------------------------------------------------------------------------------------------------------
__global__ void kc_finalize_kernel(KcDev *d, const KcPlan *__restrict__ starts, const uint32_t *__restrict__ slot_meta,
                                   const uint64_t *__restrict__ states, size_t n_states) {
    // lane 0 does the scalar bookkeeping; the commitments' permutations run on the two 16-lane groups, 12 lanes each
    __shared__ uint64_t e_out[440], o_out[32], compact[24];
    __shared__ uint32_t sh_done, sh_n_out;
    const int lane = threadIdx.x & 31, li = lane & 15;
    const unsigned gm = lane < 16 ? 0xFFFFu : 0xFFFF0000u;
    if (lane == 0) {
    zkc_keccak_closed_form &io = d->io;
    const size_t limit = d->limit;
    zkc_keccak_fsm out = limit ? d->s_final : d->s0;
    // requests queue after the executed rows
    zkc_queue_state4 rq = d->rq0;
    const uint32_t popped = limit ? d->popped_requests : 0;
    if (popped) for (int i = 0; i < 4; i++) rq.head[i] = d->req_head_final[i];
    rq.length = d->rq0.length - popped;
    zkc_queue_state12 mq = d->mq0;
    bool hint_bad = d->hint_bad;
    if (limit) {
        const uint32_t pushes = slot_meta[7 * (limit - 1) + 6] & 0x7FFFFFFFu;
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
        for (int i = 0; i < 4; i++) if (rq.head[i] != rq.tail[i]) checks |= ZKC_KC_CHK_QUEUE_CONSISTENCY;  // :667
    const bool done = out.completed;
    zkc_queue_state12 obs_out;
    memset(&obs_out, 0, sizeof obs_out);
    if (done) obs_out = mq;
    uint64_t e_exp[439], o_exp[25];
    const int n_out = kc_encode_fsm(out, e_out);
    put_q12(o_out, obs_out);
    zkc_status st;
    st.code = ZKC_OK; st.cuda_error = 0; st.first_bad_row = -1; st.failed_checks = checks; st.reserved = 0;
    if (d->first_bad != ~0ull) st.first_bad_row = (int64_t)(d->first_bad >> 16);
    if (checks) st.code = ZKC_ERR_UNSATISFIED;
    if (hint_bad) { st.code = ZKC_ERR_QUEUE_WITNESS_INCONSISTENT; st.failed_checks |= ZKC_KC_CHK_QUEUE_HINT; }
    if (d->opt.compare_expected) {
        kc_encode_fsm(io.hidden_fsm_output, e_exp);
        put_q12(o_exp, io.final_memory_state);
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
// One thread per cycle: the FSM state on entry is rebuilt from the PREVIOUS cycle's cells (flags, call parameters, offsets, buffer
// bytes and fill count, keccak state; cycle 0: the start state), the free inputs of the cycle are taken from its own cells (the
// popped call, the words read from memory), and the cycle function itself (kc_cycle, the code that generates the trace) runs in CHECK
// mode: every cell it would write -- selects, the six unaligned reads and the buffer fills, padding decisions, the absorbed block,
// keccak-f[1600], the digest word, the next flags, the buffer after the cycle -- is compared with the trace.  Around it: the
// conditional pop (ranges, aux byte / formal address, queue length / head) and the memory queue's length / tail chain over the
// seven conditional pushes.  With ZKC_GATES_ROUND_FUNCTION also the Poseidon2 permutations (the pop, every executed push).
template <bool ROUND_FUNCTION>
__global__ void __launch_bounds__(64)
kc_check_kernel(KcDev *d, unsigned long long *violations, const uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const bool first = row == 0;
#define TR(col) __ldg(trace + (size_t)(col) * limit + row)
#define TP(col) __ldg(trace + (size_t)(col) * limit + row - 1)
    uint32_t bad = 0;
    const uint32_t aux_byte = d->opt.aux_byte ? d->opt.aux_byte : ZKC_PRECOMPILE_AUX_BYTE_DEFAULT;
    const uint32_t formal = d->opt.precompile_address ? d->opt.precompile_address : ZKC_KECCAK256_PRECOMPILE_ADDRESS_DEFAULT;
    KcState s;
    if (first) kc_load_state(d->s0, s);
    else {
        const int lastq = ZKC_KC_QUERY + 5 * ZKC_KC_QUERY_STRIDE;
        s.read_precompile_call = (uint32_t)TP(ZKC_KC_FLAGS_OUT + 0); s.read_unaligned = (uint32_t)TP(ZKC_KC_FLAGS_OUT + 1);
        s.padding_round = (uint32_t)TP(ZKC_KC_FLAGS_OUT + 2); s.completed = (uint32_t)TP(ZKC_KC_FLAGS_OUT + 3);
        s.ts_read = (uint32_t)TP(ZKC_KC_TS_READ); s.ts_write = (uint32_t)TP(ZKC_KC_TS_WRITE);
        s.input_page = (uint32_t)TP(ZKC_KC_PARAMS + 0); s.byte_offset = (uint32_t)TP(lastq + 25); s.byte_length = (uint32_t)TP(lastq + 26);
        s.output_page = (uint32_t)TP(ZKC_KC_PARAMS + 3); s.output_word_offset = (uint32_t)TP(ZKC_KC_PARAMS + 4); s.needs_full_padding = (uint32_t)TP(ZKC_KC_PARAMS + 5);
        const uint32_t cf = (uint32_t)TP(ZKC_KC_CURRENTLY_FILLED);
        s.filled = cf < ZKC_KECCAK_RATE_BYTES ? 0 : cf - ZKC_KECCAK_RATE_BYTES;  // consume::<136>(allow_partial)
        for (int i = 0; i < ZKC_KECCAK_BUFFER_SIZE; i++) s.buffer[i] = (uint8_t)TP(ZKC_KC_BUFFER_OUT + i);
        for (int i = 0; i < 200; i++) s.sponge[i] = (uint8_t)TP(ZKC_KC_STATE_OUT + i);
    }
    const bool read_call = s.read_precompile_call;
    // the conditional pop
    uint64_t f[36], limbs = 0, any = 0;
#pragma unroll
    for (int i = 0; i < 36; i++) { f[i] = TR(ZKC_KC_CALL_ITEM + i); any |= f[i]; }
#pragma unroll
    for (int i = 0; i < 29; i++) limbs |= f[i];
    if ((limbs | f[34] | f[35]) >> 32 || (f[29] | f[33]) >> 8 || (f[30] | f[31] | f[32]) > 1 || (!read_call && any)) bad |= ZKC_KCV_BOOLEAN;
    if (read_call && (f[29] != aux_byte || f[0] != formal || (f[1] | f[2] | f[3] | f[4]))) bad |= ZKC_KCV_ENFORCE;
    zkc_log_query call = lq_zero();
#pragma unroll
    for (int i = 0; i < 5; i++) call.address[i] = (uint32_t)f[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { call.key[i] = (uint32_t)f[5 + i]; call.read_value[i] = (uint32_t)f[13 + i]; call.written_value[i] = (uint32_t)f[21 + i]; }
    call.flags = ZKC_LQ_FLAGS((uint32_t)f[29], (uint32_t)f[33], (uint32_t)f[30], (uint32_t)f[31], (uint32_t)f[32]);
    call.tx_number_in_block = (uint32_t)f[34]; call.timestamp = (uint32_t)f[35];
    const uint64_t len_prev = first ? d->rq0.length : TP(ZKC_KC_REQ_LEN), len = TR(ZKC_KC_REQ_LEN);
    if (len + (read_call ? 1u : 0u) != len_prev || (len >> 32)) bad |= ZKC_KCV_QUEUE;
    {
        uint64_t head[4], head_prev[4];
        bool same = true;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            head[i] = TR(ZKC_KC_REQ_HEAD + i);
            head_prev[i] = first ? d->rq0.head[i] : TP(ZKC_KC_REQ_HEAD + i);
            same &= head[i] == head_prev[i];
            if (head[i] >= GL_P) bad |= ZKC_KCV_BOOLEAN;
        }
        if (!read_call && !same) bad |= ZKC_KCV_QUEUE;
        if (ROUND_FUNCTION && read_call) {
            uint64_t e[20], st[12];
            lq_encode(call, e);
            lq_absorb_head(e, st);
            lq_absorb_tail(e, head_prev, st);
#pragma unroll
            for (int i = 0; i < 4; i++) if (st[i] != head[i]) bad |= ZKC_KCV_ROUND_FUNCTION;
        }
    }
    // the cycle function in CHECK mode: compares every cell it computes with the trace
    {
        size_t read_cursor = 0;
        uint32_t push_ordinal = 0, checks = 0;
        kc_cycle<false, true>(s, call, len == 0, nullptr, 0, read_cursor, nullptr, nullptr, push_ordinal, const_cast<uint64_t *>(trace), limit, row, checks);
        bad |= checks;
    }
    // the memory queue: previous write -> six conditional reads -> conditional digest write
    {
        uint64_t mt_prev[12], ml_prev = first ? d->mq0.length : TP(ZKC_KC_WRITE_LEN);
#pragma unroll
        for (int i = 0; i < 12; i++) mt_prev[i] = first ? d->mq0.tail[i] : TP(ZKC_KC_WRITE_TAIL + i);
        const uint32_t ts_read = (uint32_t)TR(ZKC_KC_TS_READ), ts_write = (uint32_t)TR(ZKC_KC_TS_WRITE);
        const uint32_t in_page = (uint32_t)TR(ZKC_KC_PARAMS + 0), out_page = (uint32_t)TR(ZKC_KC_PARAMS + 3), out_offset = (uint32_t)TR(ZKC_KC_PARAMS + 4);
#pragma unroll 1
        for (int q = 0; q < 7; q++) {
            const int b = ZKC_KC_QUERY + q * ZKC_KC_QUERY_STRIDE;
            const int tail_col = q < 6 ? b + 12 : ZKC_KC_WRITE_TAIL, len_col = q < 6 ? b + 24 : ZKC_KC_WRITE_LEN;
            const uint64_t pushed = q < 6 ? TR(b + 3) : TR(ZKC_KC_WRITE_RESULT);
            uint64_t mt[12], st[12];
            bool msame = true;
#pragma unroll
            for (int i = 0; i < 12; i++) { mt[i] = TR(tail_col + i); msame &= mt[i] == mt_prev[i]; if (mt[i] >= GL_P) bad |= ZKC_KCV_BOOLEAN; }
            const uint64_t ml = TR(len_col);
            if (pushed > 1 || ml != ml_prev + pushed || (!pushed && !msame)) bad |= ZKC_KCV_MEMORY_QUEUE;
            if (ROUND_FUNCTION && pushed) {
                uint32_t value[8];
#pragma unroll
                for (int i = 0; i < 8; i++) value[i] = (uint32_t)(q < 6 ? TR(b + 4 + i) : TR(ZKC_KC_RESULT + i));
                if (q < 6) mq_encode(ts_read, in_page, (uint32_t)TR(b + 0), 0, value, st);
                else mq_encode(ts_write, out_page, out_offset, 1, value, st);
#pragma unroll
                for (int i = 8; i < 12; i++) st[i] = mt_prev[i];
                poseidon2_permute(st);
#pragma unroll
                for (int i = 0; i < 12; i++) if (st[i] != mt[i]) bad |= ZKC_KCV_ROUND_FUNCTION;
            }
#pragma unroll
            for (int i = 0; i < 12; i++) mt_prev[i] = mt[i];
            ml_prev = ml;
        }
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

extern "C" int zkc_keccak256_round_function_entry_point(zkc_ctx *ctx, zkc_keccak_closed_form *io, const zkc_log_query *requests,
                                                        const uint64_t *requests_prev_tails, size_t n_requests,
                                                        const uint32_t *memory_reads, size_t n_reads,
                                                        const uint64_t *memory_states, size_t n_memory_states, size_t limit,
                                                        const zkc_precompile_options *options, int on_device, uint64_t *trace,
                                                        uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status) {
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
    const size_t max_pushes = 7 * limit + 1;
    if (!have_states) n_memory_states = max_pushes;
    size_t bytes = zkc_carver::bytes(1, sizeof(KcDev)) + zkc_carver::bytes(1, sizeof(ScanGlobal)) +
                   zkc_carver::bytes(tiles + 1, sizeof(TileStateT<KcPlan>)) + zkc_carver::bytes(max_units + 2, sizeof(KcPlan)) +
                   zkc_carver::bytes(max_pushes * 8, 8) + zkc_carver::bytes(7 * limit + 8, 4);
    if (!in_dev) bytes += zkc_carver::bytes(n_requests + 1, sizeof(zkc_log_query)) + zkc_carver::bytes(n_requests * 4 + 4, 8) +
                          zkc_carver::bytes(n_reads * 8 + 8, 4);
    if (!in_dev || !have_states) bytes += zkc_carver::bytes(n_memory_states * 12 + 12, 8);
    if (trace && !trace_dev) bytes += zkc_carver::bytes((size_t)ZKC_KC_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    KcDev *h = (KcDev *)ctx->pinned(sizeof(KcDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    KcDev *d = cv.take<KcDev>(1);
    char *zero_begin = cv.base + cv.off;
    ScanGlobal *sg = cv.take<ScanGlobal>(1);
    TileStateT<KcPlan> *ts = cv.take<TileStateT<KcPlan>>(tiles + 1);
    char *zero_end = cv.base + cv.off;
    KcPlan *starts = cv.take<KcPlan>(max_units + 2);
    uint64_t *push_enc = cv.take<uint64_t>(max_pushes * 8);
    uint32_t *slot_meta = cv.take<uint32_t>(7 * limit + 8);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(KcDev));
    h->io = *io;
    if (options) h->opt = *options;
    h->n_requests = n_requests; h->n_reads = n_reads; h->n_memory_states = n_memory_states; h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(KcDev), cudaMemcpyHostToDevice, s));
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
    if (trace && !trace_dev) dtrace = cv.take<uint64_t>((size_t)ZKC_KC_NUM_COLS * limit);

    ZKC_LAUNCH(ctx, "kc_prologue", kc_prologue_kernel, 1, 96, 0, d);
    ZKC_LAUNCH(ctx, "kc_plan", kc_plan_kernel, (unsigned)tiles, SCAN_THREADS, 0, d, dreq, starts, sg, ts);
    if (limit) {
        ZKC_LAUNCH(ctx, "kc_calls", kc_calls_kernel, (unsigned)((max_units + 127) / 128), 128, 0, d, dreq, dprev, dreads, starts,
                   push_enc, slot_meta, dtrace);
        ZKC_LAUNCH(ctx, "kc_tail", kc_tail_kernel, (unsigned)((limit + 127) / 128), 128, 0, d, starts, slot_meta, dtrace);
        if (!have_states) ZKC_LAUNCH(ctx, "kc_mem_chain", (pc_mem_chain_kernel<KcDev, 7>), 1, 32, 0, d, push_enc, slot_meta, (uint64_t *)dstates);
        ZKC_LAUNCH(ctx, "kc_memq", (pc_memq_kernel<KcDev, 7, ZKC_KC_QUERY + 12, ZKC_KC_QUERY_STRIDE, ZKC_KC_WRITE_TAIL, ZKC_KC_CHK_QUEUE_HINT>),
                   (unsigned)((7 * limit + 255) / 256), 256, 0, d, push_enc, slot_meta, dstates, n_memory_states, have_states, dtrace);
    }
    ZKC_LAUNCH(ctx, "kc_finalize", kc_finalize_kernel, 1, 32, 0, d, starts, slot_meta, dstates, n_memory_states);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(KcDev), cudaMemcpyDeviceToHost, s));
    if (!trace_dev && trace && limit)
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(trace, dtrace, (size_t)ZKC_KC_NUM_COLS * limit * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    io->hidden_fsm_output = h->io.hidden_fsm_output;
    io->final_memory_state = h->io.final_memory_state;
    io->completion_flag = h->io.completion_flag;
    memcpy(commitment, h->commitment, 32);
    *status = h->status;
    return status->code;
}

extern "C" int zkc_keccak256_round_function_check_trace(zkc_ctx *ctx, const zkc_keccak_closed_form *io, const zkc_precompile_options *options,
                                                        const uint64_t *trace, size_t limit, uint32_t gates, int on_device, uint64_t *violations,
                                                        zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !violations || (limit && !trace)) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    size_t bytes = zkc_carver::bytes(1, sizeof(KcDev)) + zkc_carver::bytes(1, 8);
    if (!on_device) bytes += zkc_carver::bytes((size_t)ZKC_KC_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    KcDev *h = (KcDev *)ctx->pinned(sizeof(KcDev) + 8);
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    KcDev *d = cv.take<KcDev>(1);
    unsigned long long *dviol = cv.take<unsigned long long>(1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(KcDev));
    h->io = *io;
    if (options) h->opt = *options;
    h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(KcDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(dviol, 0, 8, s));
    const uint64_t *dt = trace;
    if (!on_device && limit) {
        uint64_t *b = cv.take<uint64_t>((size_t)ZKC_KC_NUM_COLS * limit);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(b, trace, (size_t)ZKC_KC_NUM_COLS * limit * 8, cudaMemcpyHostToDevice, s));
        dt = b;
    }
    ZKC_LAUNCH(ctx, "kc_prologue", kc_prologue_kernel, 1, 96, 0, d);
    if (limit) {
        const unsigned grid = (unsigned)((limit + 63) / 64);
        if (gates == 0 || (gates & ZKC_GATES_ROUND_FUNCTION)) ZKC_LAUNCH(ctx, "kc_check_rf", kc_check_kernel<true>, grid, 64, 0, d, dviol, dt);
        else ZKC_LAUNCH(ctx, "kc_check", kc_check_kernel<false>, grid, 64, 0, d, dviol, dt);
    }
    ZKC_CUDA(ctx, status, cudaGetLastError());
    unsigned long long *hviol = (unsigned long long *)(h + 1);
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(KcDev), cudaMemcpyDeviceToHost, s));
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
