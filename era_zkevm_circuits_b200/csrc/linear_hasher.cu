// L2 -> L1 message hasher on sm_100a: linear_hasher_entry_point (/root/reference/src/linear_hasher/mod.rs:35-214), one thread
// per loop iteration.  The loop's sequential state and how each row recovers it:
//   - the popped queue's head: previous-tail column of the raw queue witness, verified link by link (as in the demultiplexer);
//   - the byte buffer (a Vec whose length is a compile-time function of the cycle, :114-125): before cycle c it holds the last
//     (88 c) mod 136 bytes of the serialisation stream, i.e. bytes of items c - 1 and c - 2 -- the row re-serialises its two
//     predecessors' records instead of carrying bytes;
//   - the keccak state: a chain no witness of the reference breaks.  The out-of-circuit hasher holds the state after every
//     cycle; with those (`keccak_states`, verified row by row) every cycle is independent and runs its (at most two)
//     keccak-f[1600] on its own thread.  Without them one thread rebuilds the chain first (lh_chain_kernel), exactly as
//     sequential as the reference.
#include "ctx.cuh"
#include "keccak_f1600.cuh"
#include "log_query.cuh"

namespace zkc {

constexpr int LH_MSG = ZKC_LH_MESSAGE_BYTES, LH_RATE = ZKC_KECCAK_RATE_BYTES;

struct LhDev {
    zkc_linear_hasher_closed_form io;
    zkc_sorter_options opt;
    uint64_t n_records, limit;
    uint32_t prologue_checks, pad0;
    uint64_t commit_obs_in[4];
    uint64_t head_final[4];
    unsigned long long first_bad;
    uint32_t failed_checks, hint_bad;
    uint64_t commitment[4];
    zkc_status status;
};

// LogQuery::into_bytes, base_structures/log_query/mod.rs:647-686 (big-endian fields)
__device__ __forceinline__ void lh_into_bytes(const zkc_log_query &q, uint8_t *out) {
    int n = 0;
    out[n++] = (uint8_t)ZKC_LQ_SHARD(q.flags);
    out[n++] = (uint8_t)ZKC_LQ_SERVICE(q.flags);
    out[n++] = (uint8_t)(q.tx_number_in_block >> 8);
    out[n++] = (uint8_t)q.tx_number_in_block;
#pragma unroll
    for (int l = 4; l >= 0; l--)
#pragma unroll
        for (int b = 3; b >= 0; b--) out[n++] = (uint8_t)(q.address[l] >> (8 * b));
#pragma unroll
    for (int l = 7; l >= 0; l--)
#pragma unroll
        for (int b = 3; b >= 0; b--) out[n++] = (uint8_t)(q.key[l] >> (8 * b));
#pragma unroll
    for (int l = 7; l >= 0; l--)
#pragma unroll
        for (int b = 3; b >= 0; b--) out[n++] = (uint8_t)(q.written_value[l] >> (8 * b));
}

// state ^= block (17 little-endian lanes), keccak-f (storage_application/mod.rs:66-82)
__device__ __forceinline__ void lh_absorb(uint64_t (&A)[25], const uint8_t *block) {
#pragma unroll
    for (int i = 0; i < LH_RATE / 8; i++) {
        uint64_t w = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) w |= (uint64_t)block[8 * i + b] << (8 * b);
        A[i] ^= w;
    }
    keccak_f1600(A);
}
// the padded remainder (:146-159)
__device__ __forceinline__ void lh_pad(const uint8_t *rest, int len, uint8_t *last) {
    for (int i = 0; i < LH_RATE; i++) last[i] = i < len ? rest[i] : 0;
    if (len == LH_RATE - 1) last[len] = 0x81;
    else { last[len] = 0x01; last[LH_RATE - 1] = 0x80; }
}

// lane 0: checks; lanes of the first 16-lane group: commitment to the observable input
__global__ void lh_prologue_kernel(LhDev *d) {
    __shared__ uint64_t buf[16];
    const int i = threadIdx.x & 31;
    if (i >= 16) return;
    const unsigned gm = 0xFFFFu;
    const zkc_linear_hasher_closed_form &io = d->io;
    if (i == 0) {
        uint32_t checks = 0;
        if (io.start_flag == 0) checks |= ZKC_LH_CHK_START_FLAG;  // :66
        for (int k = 0; k < 4; k++)
            if (io.queue_state.head[k]) checks |= ZKC_LH_CHK_TRIVIAL_HEAD;  // :71
        d->prologue_checks = checks;
        put_queue_state4(buf, io.queue_state);
    }
    __syncwarp(gm);
    const uint64_t c = commit_encoding_coop(gm, buf, 9, i);
    if (i < 4) d->commit_obs_in[i] = c;
}

// the keccak chain when the caller supplies no states: one thread, the loop of the reference (:103-171) on the hash alone
__global__ void lh_chain_kernel(const LhDev *d, const zkc_log_query *__restrict__ recs, uint64_t *__restrict__ states) {
    if (threadIdx.x || blockIdx.x) return;
    const size_t limit = d->limit, len0 = d->io.queue_state.length;
    uint64_t A[25];
    for (int i = 0; i < 25; i++) A[i] = 0;
    uint8_t buffer[2 * LH_RATE], last[LH_RATE];
    int len = 0;
    for (size_t c = 0; c < limit; c++) {
        if (c < len0) {
            const zkc_log_query it = c < d->n_records ? lq_load(recs + c) : lq_zero();
            lh_into_bytes(it, buffer + len);
            len += LH_MSG;
            if (len >= LH_RATE) {
                lh_absorb(A, buffer);
                for (int i = LH_RATE; i < len; i++) buffer[i - LH_RATE] = buffer[i];
                len -= LH_RATE;
            }
            if (c + 1 == len0) {
                lh_pad(buffer, len, last);
                lh_absorb(A, last);
            }
        }
        for (int i = 0; i < 25; i++) states[25 * c + i] = A[i];
    }
}

__global__ void __launch_bounds__(128)
lh_rows_kernel(LhDev *d, const zkc_log_query *__restrict__ recs, const uint64_t *__restrict__ prev, const uint64_t *__restrict__ states,
               uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const uint32_t len0 = d->io.queue_state.length;
    const bool queue_is_empty = row >= len0, should_pop = !queue_is_empty;
    const size_t active_rows = limit < len0 ? limit : len0;
    uint32_t checks = 0;
#define TR(col) trace[(size_t)(col) * limit + row]
    const bool wr = trace != nullptr;
    zkc_log_query it = lq_zero();
    if (should_pop && row < d->n_records) it = lq_load(recs + row);
    uint64_t e[20], s[12], head[4];
    lq_encode(it, e);
    if (should_pop) {
        uint64_t chain[4];
        bool hint_ok = true;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            chain[i] = __ldg(prev + 4 * row + i);
            if (row == 0 && chain[i] != d->io.queue_state.head[i]) hint_ok = false;
        }
        lq_absorb_head(e, s);
        lq_absorb_tail(e, chain, s);
#pragma unroll
        for (int i = 0; i < 4; i++) head[i] = s[i];
        if (row + 1 < active_rows) {
#pragma unroll
            for (int i = 0; i < 4; i++) hint_ok &= __ldg(prev + 4 * (row + 1) + i) == head[i];
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) d->head_final[i] = head[i];
        }
        if (!hint_ok) { checks |= ZKC_LH_CHK_QUEUE_HINT; d->hint_bad = 1; }
    } else {
        // an empty queue keeps its head: the initial one when nothing was ever popped, else the tail (checked by enforce_consistency)
#pragma unroll
        for (int i = 0; i < 4; i++) head[i] = len0 == 0 ? d->io.queue_state.head[i] : d->io.queue_state.tail[i];
    }
    if (it.tx_number_in_block >> 16) checks |= ZKC_LH_CHK_TX_NUMBER_RANGE;
    const bool now_empty = row + 1 >= len0, is_last = should_pop && now_empty;
    const bool continue_to_absorb = row < len0;  // done before this cycle <=> row >= len0
    // ---- the byte window: items row - 2, row - 1, row ----------------------------------------------------------------------
    uint8_t win[3 * LH_MSG], last[LH_RATE];
    {
        const zkc_log_query a = (row >= 2 && row - 2 < len0 && row - 2 < d->n_records) ? lq_load(recs + row - 2) : lq_zero();
        const zkc_log_query b = (row >= 1 && row - 1 < len0 && row - 1 < d->n_records) ? lq_load(recs + row - 1) : lq_zero();
        lh_into_bytes(a, win);
        lh_into_bytes(b, win + LH_MSG);
        lh_into_bytes(it, win + 2 * LH_MSG);
    }
    const int lb = (int)((row * LH_MSG) % LH_RATE);     // buffer length before the cycle
    const bool has_full = lb + LH_MSG >= LH_RATE;       // :120
    const bool absorb_full = has_full && continue_to_absorb, absorb_last = continue_to_absorb && is_last;
    uint64_t A[25];
#pragma unroll
    for (int i = 0; i < 25; i++) A[i] = row ? __ldg(states + 25 * (row - 1) + i) : 0ull;
    const uint8_t *buf = win + 2 * LH_MSG - lb;
    if (absorb_full) lh_absorb(A, buf);
    if (wr) {
#pragma unroll
        for (int i = 0; i < 25; i++) { TR(ZKC_LH_STATE_MID + 2 * i) = (uint32_t)A[i]; TR(ZKC_LH_STATE_MID + 2 * i + 1) = A[i] >> 32; }
    }
    if (absorb_last) {
        const int rest = has_full ? lb + LH_MSG - LH_RATE : lb + LH_MSG;
        lh_pad(has_full ? buf + LH_RATE : buf, rest, last);
        lh_absorb(A, last);
    }
    {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 25; i++) ok &= __ldg(states + 25 * row + i) == A[i];
        if (!ok) { checks |= ZKC_LH_CHK_STATE_HINT; d->hint_bad = 1; }
    }
    if (wr) {
        TR(ZKC_LH_QUEUE_IS_EMPTY) = queue_is_empty; TR(ZKC_LH_SHOULD_POP) = should_pop;
#pragma unroll
        for (int i = 0; i < 36; i++) TR(ZKC_LH_ITEM + i) = lq_flat(it, i);
#pragma unroll
        for (int i = 0; i < 20; i++) TR(ZKC_LH_ENC + i) = e[i];
#pragma unroll
        for (int i = 0; i < 4; i++) TR(ZKC_LH_HEAD + i) = head[i];
        const size_t popped_now = row + 1 < active_rows ? row + 1 : active_rows;
        TR(ZKC_LH_LEN) = len0 - (uint32_t)popped_now;
        TR(ZKC_LH_NOW_EMPTY) = now_empty; TR(ZKC_LH_IS_LAST_SERIALIZATION) = is_last;
        for (int i = 0; i < LH_MSG; i++) TR(ZKC_LH_BYTES + i) = win[2 * LH_MSG + i];
        TR(ZKC_LH_CONTINUE_TO_ABSORB) = continue_to_absorb; TR(ZKC_LH_ABSORB_FULL) = absorb_full; TR(ZKC_LH_ABSORB_LAST) = absorb_last;
#pragma unroll
        for (int i = 0; i < 25; i++) { TR(ZKC_LH_STATE_OUT + 2 * i) = (uint32_t)A[i]; TR(ZKC_LH_STATE_OUT + 2 * i + 1) = A[i] >> 32; }
        TR(ZKC_LH_DONE) = row + 1 >= len0;
    }
    if (checks) {
        atomicOr(&d->failed_checks, checks);
        atomicMin(&d->first_bad, ((unsigned long long)row << 16) | checks);
    }
#undef TR
}

// 16 lanes: lane 0 does the bookkeeping, the two commitments run on 12 lanes
__global__ void lh_finalize_kernel(LhDev *d, const uint64_t *__restrict__ states) {
    __shared__ uint64_t e_out[32], compact[24];
    __shared__ uint32_t sh_completed;
    const int i = threadIdx.x & 31;
    if (i >= 16) return;
    const unsigned gm = 0xFFFFu;
    zkc_linear_hasher_closed_form &io = d->io;
    if (i == 0) {
        const size_t limit = d->limit;
        const uint32_t len0 = io.queue_state.length;
        const size_t popped = limit < len0 ? limit : len0;
        uint64_t head[4];
        for (int k = 0; k < 4; k++) head[k] = popped ? d->head_final[k] : io.queue_state.head[k];
        const uint32_t len = len0 - (uint32_t)popped;
        uint32_t checks = d->failed_checks | d->prologue_checks;
        const bool completed = len == 0;
        if (completed)
            for (int k = 0; k < 4; k++)
                if (head[k] != io.queue_state.tail[k]) checks |= ZKC_LH_CHK_QUEUE_CONSISTENCY;  // :173
        if (!completed) checks |= ZKC_LH_CHK_NOT_COMPLETED;                                      // :176
        uint32_t digest[32];
        if (len0 == 0) {  // no_work: Keccak-256 of the empty string (:87-96, :195-196)
            uint64_t A[25];
            for (int k = 0; k < 25; k++) A[k] = 0;
            A[0] = 0x01; A[16] = 0x8000000000000000ull;
            keccak_f1600(A);
            for (int k = 0; k < 32; k++) digest[k] = (uint32_t)(A[k >> 3] >> (8 * (k & 7))) & 0xFF;
        } else {
            for (int k = 0; k < 32; k++) digest[k] = limit ? (uint32_t)(states[25 * (limit - 1) + (k >> 3)] >> (8 * (k & 7))) & 0xFF : 0u;
        }
        zkc_status st;
        st.code = ZKC_OK; st.cuda_error = 0; st.first_bad_row = -1; st.failed_checks = checks; st.reserved = 0;
        if (d->first_bad != ~0ull) st.first_bad_row = (int64_t)(d->first_bad >> 16);
        if (checks) st.code = ZKC_ERR_UNSATISFIED;
        if (d->hint_bad) st.code = ZKC_ERR_QUEUE_WITNESS_INCONSISTENT;
        if (d->opt.compare_expected) {  // hook_compare_witness, :200
            bool same = (io.completion_flag != 0) == completed;
            for (int k = 0; k < 32; k++) same &= io.keccak256_hash[k] == digest[k];
            if (!same && st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
        }
        for (int k = 0; k < 32; k++) { io.keccak256_hash[k] = digest[k]; e_out[k] = digest[k]; }
        io.completion_flag = completed;
        d->status = st;
        sh_completed = completed;
    }
    __syncwarp(gm);
    const uint64_t c = commit_encoding_coop(gm, e_out, 32, i);
    const bool completed = sh_completed;
    if (i == 0) {
        compact[0] = io.start_flag != 0; compact[1] = completed;
        for (int k = 0; k < 4; k++) {
            compact[2 + k] = d->commit_obs_in[k];
            compact[10 + k] = 0;  // the hidden FSM input / output are `()`: the commitment of an empty encoding is zero
            compact[14 + k] = 0;
        }
    }
    if (i < 4) compact[6 + i] = completed ? c : 0;
    __syncwarp(gm);
    const uint64_t f = commit_encoding_coop(gm, compact, 18, i);
    if (i < 4) d->commitment[i] = f;
}


// ---- constraint evaluation of a finished trace ------------------------------------------------------------------------------
// One thread per cycle re-evaluates every relation of the loop of linear_hasher_entry_point (mod.rs:103-171) that is local to a
// cycle or to a cycle and its predecessors: the conditional pop (ranges, LogQuery::encode, queue length / head), into_bytes and its
// tx-number range, now_empty / is_last_serialization / done / continue_to_absorb, the two absorption conditions, and the keccak
// sponge itself: the buffer before the cycle is the last (88 c) mod 136 bytes of the serialisation stream, i.e. of the BYTES columns
// of cycles c - 2 and c - 1; STATE_MID / STATE_OUT are the previous cycle's state after the conditional full-block round and the
// conditional padded last round (keccak-f[1600] recomputed).  With ZKC_GATES_ROUND_FUNCTION also the three Poseidon2 permutations
// of the pop.
template <bool ROUND_FUNCTION>
__global__ void __launch_bounds__(128)
lh_check_kernel(LhDev *d, unsigned long long *violations, const uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const bool first = row == 0;
    const zkc_queue_state4 &q0 = d->io.queue_state;
#define TR(col) __ldg(trace + (size_t)(col) * limit + row)
#define TP(col) __ldg(trace + (size_t)(col) * limit + row - 1)
    uint32_t bad = 0;
    const uint64_t is_empty = TR(ZKC_LH_QUEUE_IS_EMPTY), should_pop = TR(ZKC_LH_SHOULD_POP);
    if ((is_empty | should_pop) > 1 || should_pop != 1 - is_empty) bad |= ZKC_LHV_BOOLEAN;
    uint64_t f[36], limbs = 0;
#pragma unroll
    for (int i = 0; i < 36; i++) f[i] = TR(ZKC_LH_ITEM + i);
#pragma unroll
    for (int i = 0; i < 29; i++) limbs |= f[i];
    if ((limbs | f[34] | f[35]) >> 32 || (f[29] | f[33]) >> 8 || (f[30] | f[31] | f[32]) > 1) bad |= ZKC_LHV_BOOLEAN;
    if (f[34] >> 16) bad |= ZKC_LHV_ENFORCE;  // into_bytes: tx_number_in_block is two bytes on the wire (log_query/mod.rs:666-668)
    zkc_log_query q = lq_zero();
#pragma unroll
    for (int i = 0; i < 5; i++) q.address[i] = (uint32_t)f[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { q.key[i] = (uint32_t)f[5 + i]; q.read_value[i] = (uint32_t)f[13 + i]; q.written_value[i] = (uint32_t)f[21 + i]; }
    q.flags = ZKC_LQ_FLAGS((uint32_t)f[29], (uint32_t)f[33], (uint32_t)f[30], (uint32_t)f[31], (uint32_t)f[32]);
    q.tx_number_in_block = (uint32_t)f[34]; q.timestamp = (uint32_t)f[35];
    uint64_t e[20], enc[20];
    lq_encode(q, e);
#pragma unroll
    for (int i = 0; i < 20; i++) { enc[i] = TR(ZKC_LH_ENC + i); if (enc[i] != e[i]) bad |= ZKC_LHV_ENCODING; }
    const uint64_t len_prev = first ? q0.length : TP(ZKC_LH_LEN), len = TR(ZKC_LH_LEN);
    if (is_empty != (uint64_t)(len_prev == 0) || len + should_pop != len_prev) bad |= ZKC_LHV_QUEUE;
    {
        uint64_t head[4], head_prev[4];
        bool same = true;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            head[i] = TR(ZKC_LH_HEAD + i);
            head_prev[i] = first ? q0.head[i] : TP(ZKC_LH_HEAD + i);
            same &= head[i] == head_prev[i];
            if (head[i] >= GL_P) bad |= ZKC_LHV_BOOLEAN;
        }
        if (!should_pop && !same) bad |= ZKC_LHV_QUEUE;
        if (ROUND_FUNCTION && should_pop) {
            uint64_t st[12];
            lq_absorb_head(enc, st);
            lq_absorb_tail(enc, head_prev, st);
#pragma unroll
            for (int i = 0; i < 4; i++) if (st[i] != head[i]) bad |= ZKC_LHV_ROUND_FUNCTION;
        }
    }
    // the flags of :104-118 and :170
    const uint64_t now_empty = TR(ZKC_LH_NOW_EMPTY), is_last = TR(ZKC_LH_IS_LAST_SERIALIZATION), cont = TR(ZKC_LH_CONTINUE_TO_ABSORB),
                   absorb_full = TR(ZKC_LH_ABSORB_FULL), absorb_last = TR(ZKC_LH_ABSORB_LAST), done = TR(ZKC_LH_DONE);
    const uint64_t done_prev = first ? (uint64_t)(q0.length == 0) : TP(ZKC_LH_DONE);
    const int lb = (int)((row * LH_MSG) % LH_RATE);  // buffer length before the cycle: a function of the cycle index (:114-125)
    const bool has_full = lb + LH_MSG >= LH_RATE;
    if ((now_empty | is_last | cont | absorb_full | absorb_last | done) > 1 || now_empty != (uint64_t)(len == 0) || is_last != (should_pop & now_empty) ||
        cont != 1 - (done_prev & 1) || done != (done_prev | is_last) || absorb_full != ((uint64_t)has_full & cont) || absorb_last != (cont & is_last))
        bad |= ZKC_LHV_FLAGS;
    // the byte window: BYTES of cycles row - 2, row - 1, row; this cycle's bytes are into_bytes of the popped item
    uint8_t win[3 * LH_MSG], last[LH_RATE];
    {
        uint8_t want[LH_MSG];
        lh_into_bytes(q, want);
        uint64_t range = 0;
        for (int i = 0; i < LH_MSG; i++) {
            const uint64_t b2 = row >= 2 ? __ldg(trace + (size_t)(ZKC_LH_BYTES + i) * limit + row - 2) : 0ull;
            const uint64_t b1 = row >= 1 ? TP(ZKC_LH_BYTES + i) : 0ull;
            const uint64_t b0 = TR(ZKC_LH_BYTES + i);
            range |= b0;
            if (b0 != want[i]) bad |= ZKC_LHV_ENCODING;
            win[i] = (uint8_t)b2; win[LH_MSG + i] = (uint8_t)b1; win[2 * LH_MSG + i] = (uint8_t)b0;
        }
        if (range >> 8) bad |= ZKC_LHV_BOOLEAN;
    }
    // the sponge: previous state -> (full block?) -> STATE_MID -> (padded last block?) -> STATE_OUT
    uint64_t A[25];
    {
        uint64_t range = 0;
#pragma unroll
        for (int i = 0; i < 25; i++) {
            const uint64_t lo = first ? 0ull : TP(ZKC_LH_STATE_OUT + 2 * i), hi = first ? 0ull : TP(ZKC_LH_STATE_OUT + 2 * i + 1);
            A[i] = lo | (hi << 32);
        }
        const uint8_t *buf = win + 2 * LH_MSG - lb;
        if (absorb_full) lh_absorb(A, buf);
#pragma unroll
        for (int i = 0; i < 25; i++) {
            const uint64_t lo = TR(ZKC_LH_STATE_MID + 2 * i), hi = TR(ZKC_LH_STATE_MID + 2 * i + 1);
            range |= lo | hi;
            if ((lo | (hi << 32)) != A[i]) bad |= ZKC_LHV_SPONGE;
        }
        if (absorb_last) {
            const int rest = has_full ? lb + LH_MSG - LH_RATE : lb + LH_MSG;
            lh_pad(has_full ? buf + LH_RATE : buf, rest, last);
            lh_absorb(A, last);
        }
#pragma unroll
        for (int i = 0; i < 25; i++) {
            const uint64_t lo = TR(ZKC_LH_STATE_OUT + 2 * i), hi = TR(ZKC_LH_STATE_OUT + 2 * i + 1);
            range |= lo | hi;
            if ((lo | (hi << 32)) != A[i]) bad |= ZKC_LHV_SPONGE;
        }
        if (range >> 32) bad |= ZKC_LHV_BOOLEAN;
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

extern "C" int zkc_linear_hasher_entry_point(zkc_ctx *ctx, zkc_linear_hasher_closed_form *io, const zkc_log_query *records,
                                             const uint64_t *prev_tails, size_t n_records, const uint64_t *keccak_states, size_t limit,
                                             const zkc_sorter_options *options, int on_device, uint64_t *trace,
                                             uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    auto invalid = [&]() { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; };
    if (!ctx || !io || !commitment || (n_records && !records) || limit > 0x7FFFFFFFull) return invalid();
    const size_t need = limit < io->queue_state.length ? limit : io->queue_state.length;
    if (n_records < need || (need && !prev_tails)) return invalid();
    const bool in_dev = on_device & ZKC_INPUTS_ON_DEVICE, trace_dev = on_device & ZKC_TRACE_ON_DEVICE;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const bool have_states = keccak_states != nullptr;
    size_t bytes = zkc_carver::bytes(1, sizeof(LhDev));
    if (!in_dev) bytes += zkc_carver::bytes(need + 1, sizeof(zkc_log_query)) + zkc_carver::bytes(need * 4 + 4, 8);
    if (!in_dev || !have_states) bytes += zkc_carver::bytes(limit * 25 + 25, 8);
    if (trace && !trace_dev) bytes += zkc_carver::bytes((size_t)ZKC_LH_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    LhDev *h = (LhDev *)ctx->pinned(sizeof(LhDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    LhDev *d = cv.take<LhDev>(1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(LhDev));
    h->io = *io;
    if (options) h->opt = *options;
    h->n_records = n_records; h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(LhDev), cudaMemcpyHostToDevice, s));
    const zkc_log_query *dr = records;
    const uint64_t *dp = prev_tails, *dstates = keccak_states;
    uint64_t *dtrace = trace;
    if (!in_dev) {
        zkc_log_query *br = cv.take<zkc_log_query>(need + 1);
        uint64_t *bp = cv.take<uint64_t>(need * 4 + 4);
        if (need) {
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(br, records, need * sizeof(zkc_log_query), cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bp, prev_tails, need * 32, cudaMemcpyHostToDevice, s));
        }
        dr = br; dp = bp;
    }
    if (!in_dev || !have_states) {
        uint64_t *bs = cv.take<uint64_t>(limit * 25 + 25);
        if (have_states && limit) ZKC_CUDA(ctx, status, cudaMemcpyAsync(bs, keccak_states, limit * 200, cudaMemcpyHostToDevice, s));
        dstates = bs;
    }
    if (trace && !trace_dev) dtrace = cv.take<uint64_t>((size_t)ZKC_LH_NUM_COLS * limit);

    ZKC_LAUNCH(ctx, "lh_prologue", lh_prologue_kernel, 1, 32, 0, d);
    if (limit) {
        if (!have_states) ZKC_LAUNCH(ctx, "lh_chain", lh_chain_kernel, 1, 32, 0, d, dr, (uint64_t *)dstates);
        ZKC_LAUNCH(ctx, "lh_rows", lh_rows_kernel, (unsigned)((limit + 127) / 128), 128, 0, d, dr, dp, dstates, dtrace);
    }
    ZKC_LAUNCH(ctx, "lh_finalize", lh_finalize_kernel, 1, 32, 0, d, dstates);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(LhDev), cudaMemcpyDeviceToHost, s));
    if (!trace_dev && trace && limit)
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(trace, dtrace, (size_t)ZKC_LH_NUM_COLS * limit * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    memcpy(io->keccak256_hash, h->io.keccak256_hash, sizeof io->keccak256_hash);
    io->completion_flag = h->io.completion_flag;
    memcpy(commitment, h->commitment, 32);
    *status = h->status;
    return status->code;
}

extern "C" int zkc_linear_hasher_check_trace(zkc_ctx *ctx, const zkc_linear_hasher_closed_form *io, const uint64_t *trace, size_t limit, uint32_t gates,
                                             int on_device, uint64_t *violations, zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !violations || (limit && !trace)) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    size_t bytes = zkc_carver::bytes(1, sizeof(LhDev)) + zkc_carver::bytes(1, 8);
    if (!on_device) bytes += zkc_carver::bytes((size_t)ZKC_LH_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    LhDev *h = (LhDev *)ctx->pinned(sizeof(LhDev) + 8);
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    LhDev *d = cv.take<LhDev>(1);
    unsigned long long *dviol = cv.take<unsigned long long>(1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(LhDev));
    h->io = *io;
    h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(LhDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(dviol, 0, 8, s));
    const uint64_t *dt = trace;
    if (!on_device && limit) {
        uint64_t *b = cv.take<uint64_t>((size_t)ZKC_LH_NUM_COLS * limit);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(b, trace, (size_t)ZKC_LH_NUM_COLS * limit * 8, cudaMemcpyHostToDevice, s));
        dt = b;
    }
    if (limit) {
        const unsigned grid = (unsigned)((limit + 127) / 128);
        if (gates == 0 || (gates & ZKC_GATES_ROUND_FUNCTION)) ZKC_LAUNCH(ctx, "lh_check_rf", lh_check_kernel<true>, grid, 128, 0, d, dviol, dt);
        else ZKC_LAUNCH(ctx, "lh_check", lh_check_kernel<false>, grid, 128, 0, d, dviol, dt);
    }
    ZKC_CUDA(ctx, status, cudaGetLastError());
    unsigned long long *hviol = (unsigned long long *)(h + 1);
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(LhDev), cudaMemcpyDeviceToHost, s));
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
