// Code decommitter on sm_100a: unpack_code_into_memory_entry_point (/root/reference/src/code_unpacker_sha256/mod.rs:33-148)
// and its work cycle unpack_code_into_memory_inner (:150-453): every deduplicated decommitment request is popped, its
// bytecode written to memory two words per cycle, and the SHA-256 of the code compared with the versioned hash.
// A request of w words takes exactly (w + 1) / 2 cycles, so the plan is a prefix sum over the requests' round counts.  The only
// state a cycle inherits that no witness supplies is the SHA-256 state (a hash chain): pass 1 runs ONE THREAD PER REQUEST
// over nothing but that chain (pop of the request, 2 code words in, one compression, 32 bytes of state out per round);
// pass 2 is ONE THREAD PER CYCLE: it rebuilds the FSM state of its cycle in closed form from (request, round index, chained
// SHA-256 state), evaluates the cycle and writes its 134 witness cells, coalesced.  The rows after the last request idle and
// are row-parallel too; the memory queue's conditional pushes (2 slots per cycle) are handled by precompile_common.cuh
// against host-supplied states or rebuilt.
#include "ctx.cuh"
#include "poseidon2.cuh"
#include "precompile_common.cuh"
#include "scan.cuh"
#include "sha256.cuh"

namespace zkc {

struct CuDev {
    zkc_code_unpacker_closed_form io;
    zkc_sorter_options opt;
    uint64_t n_requests, n_code_words, n_memory_states, limit;
    uint32_t start, n_units, unit0_fresh, pad0;
    zkc_code_decommittment_fsm s0;
    zkc_queue_state12 rq0, mq0;
    uint64_t commit_obs_in[4], commit_fsm_in[4];
    zkc_code_decommittment_fsm s_last, s_final;
    uint32_t popped_requests, pad1;
    uint64_t req_head_final[12];
    unsigned long long first_bad;
    uint32_t failed_checks, hint_bad;
    uint64_t commitment[4];
    zkc_status status;
};

struct CuPlan {
    uint32_t cycles, words;
};
struct CuPlanOp {
    static __device__ __forceinline__ CuPlan identity() { return CuPlan{0, 0}; }
    static __device__ __forceinline__ CuPlan combine(const CuPlan &a, const CuPlan &b) { return CuPlan{a.cycles + b.cycles, a.words + b.words}; }
};

__device__ __forceinline__ zkc_decommit_query cu_load_request(const zkc_decommit_query *p) {
    zkc_decommit_query q;
    const uint4 *s = reinterpret_cast<const uint4 *>(p);
    uint4 *d = reinterpret_cast<uint4 *>(&q);
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = __ldg(s + i);
    q.is_first &= 1u;
    q._pad = 0;
    return q;
}
__device__ __forceinline__ zkc_decommit_query cu_zero_request() {
    zkc_decommit_query q;
    uint4 *d = reinterpret_cast<uint4 *>(&q);
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = make_uint4(0, 0, 0, 0);
    return q;
}
// DecommitQuery::encode, decommit_query/mod.rs:31-107
__device__ __forceinline__ void cu_encode_request(const zkc_decommit_query &q, uint64_t (&e)[8]) {
    e[0] = (uint64_t)q.code_hash[0] | ((uint64_t)(q.page & 0xFFFFFFu) << 32);
    e[1] = (uint64_t)q.code_hash[1] | ((uint64_t)(q.page >> 24) << 32) | ((uint64_t)(q.timestamp & 0xFFFFu) << 40);
    e[2] = (uint64_t)q.code_hash[2] | ((uint64_t)(q.timestamp >> 16) << 32) | ((uint64_t)(q.is_first & 1u) << 48);
#pragma unroll
    for (int i = 3; i < 8; i++) e[i] = q.code_hash[i];
}
__device__ __forceinline__ uint64_t cu_flat_request(const zkc_decommit_query &q, int i) {
    return i < 8 ? q.code_hash[i] : i == 8 ? q.page : i == 9 ? (q.is_first & 1u) : q.timestamp;
}

static __device__ int cu_put_q12(uint64_t *dst, const zkc_queue_state12 &s) {
    for (int i = 0; i < 12; i++) dst[i] = s.head[i];
    for (int i = 0; i < 12; i++) dst[12 + i] = s.tail[i];
    dst[24] = s.length;
    return 25;
}
// CSVarLengthEncodable order of CodeDecommitterFSMInputOutput (input.rs:70-74) over CodeDecommittmentFSM (:27-38)
static __device__ int cu_encode_fsm(const zkc_code_unpacker_fsm &f, uint64_t *dst) {
    const zkc_code_decommittment_fsm &s = f.internal_fsm;
    int n = 0;
    for (int i = 0; i < 8; i++) dst[n++] = s.sha256_inner_state[i];
    for (int i = 0; i < 8; i++) dst[n++] = s.hash_to_compare_against[i];
    dst[n++] = s.current_index; dst[n++] = s.current_page; dst[n++] = s.timestamp;
    dst[n++] = s.num_rounds_left; dst[n++] = s.length_in_bits;
    dst[n++] = s.state_get_from_queue; dst[n++] = s.state_decommit; dst[n++] = s.finished;
    n += cu_put_q12(dst + n, f.decommittment_requests_queue_state);
    n += cu_put_q12(dst + n, f.memory_queue_state);
    return n;  // 74
}

// cycles a request with `rounds` rounds left occupies (UInt16 counter: 0 wraps and takes 65536), capped by the instance
__device__ __forceinline__ uint32_t cu_cycles_of(uint32_t rounds, size_t limit) {
    const uint32_t c = rounds ? rounds : 65536u;
    return c < limit ? c : (uint32_t)limit;
}

// warp 0: start selection and the unit layout; warps 1 / 2: the two input commitments on 12 cooperating lanes
__global__ void cu_prologue_kernel(CuDev *d) {
    __shared__ uint64_t buf[2][80];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane >= 16 || (warp == 0 && lane != 0)) return;
    const unsigned gm = 0xFFFFu;
    const zkc_code_unpacker_closed_form &io = d->io;
    if (warp == 0) {
        const bool start = io.start_flag != 0;
        d->start = start;
        d->rq0 = start ? io.sorted_requests_queue_initial_state : io.hidden_fsm_input.decommittment_requests_queue_state;  // :59-69
        d->mq0 = start ? io.memory_queue_initial_state : io.hidden_fsm_input.memory_queue_state;                            // :77-84
        zkc_code_decommittment_fsm s;
        if (start) { memset(&s, 0, sizeof s); s.state_get_from_queue = 1; }  // :86-95
        else s = io.hidden_fsm_input.internal_fsm;
        s.state_get_from_queue &= 1; s.state_decommit &= 1; s.finished &= 1; s.num_rounds_left &= 0xFFFF;
        d->s0 = s; d->s_last = s; d->s_final = s;
        d->popped_requests = 0;
        // unit 0 = the request in progress on entry (empty if the FSM is about to pop, or idle); then one unit per request
        d->unit0_fresh = (s.state_get_from_queue || !s.state_decommit) ? 1 : 0;
        const uint64_t avail = d->rq0.length < d->n_requests ? d->rq0.length : d->n_requests;
        d->n_units = 1 + ((s.state_get_from_queue || s.state_decommit) ? (uint32_t)avail : 0);
    } else {
        uint64_t *b = buf[warp - 1];
        int n = 0;
        if (lane == 0) {
            if (warp == 1) {
                n = cu_put_q12(b, io.memory_queue_initial_state);
                n += cu_put_q12(b + n, io.sorted_requests_queue_initial_state);
            } else {
                n = cu_encode_fsm(io.hidden_fsm_input, b);
            }
        }
        __syncwarp(gm);
        n = __shfl_sync(gm, n, 0, 16);
        const uint64_t c = commit_encoding_coop(gm, b, n, lane);
        if (lane < 4) (warp == 1 ? d->commit_obs_in : d->commit_fsm_in)[lane] = c;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
cu_plan_kernel(CuDev *d, const zkc_decommit_query *__restrict__ requests, CuPlan *__restrict__ starts, ScanGlobal *sg,
               TileStateT<CuPlan> *tiles) {
    __shared__ ScanSharedT<CuPlan> sh;
    const unsigned int tile = scan_take_ticket(sg, sh);
    const size_t u = (size_t)tile * SCAN_THREADS + threadIdx.x;
    const uint32_t n_units = d->n_units;
    CuPlan v = CuPlanOp::identity();
    if (u < n_units) {
        bool run = true;
        uint32_t rounds = 0;
        if (u == 0) { run = !d->unit0_fresh; rounds = d->s0.num_rounds_left; }
        else rounds = ((__ldg(&requests[u - 1].code_hash[7]) & 0xFFFFu) + 1) >> 1;  // :207-221
        if (run) {
            v.cycles = cu_cycles_of(rounds, d->limit);
            v.words = 2 * v.cycles - 1;  // two words per cycle, one in the finalizing one
        }
    }
    CuPlan incl;
    const CuPlan excl = scan_tile_generic<CuPlan, CuPlanOp>(v, tile, CuPlanOp::identity(), tiles, sh, incl);
    if (u < n_units) starts[u] = excl;
    if (u + 1 == n_units) starts[n_units] = incl;
}

__device__ __forceinline__ void cu_report(CuDev *d, size_t row, uint32_t checks) {
    if (!checks) return;
    atomicOr(&d->failed_checks, checks);
    atomicMin(&d->first_bad, ((unsigned long long)row << 16) | checks);
}

// one iteration of the work cycle, mod.rs:191-447.  `req`: the request popped in this cycle (zero if none); `can_pop`:
// a request is available when the FSM asks for one (otherwise the pop is reported and the FSM idles from here on)
__device__ __forceinline__ void cu_cycle(zkc_code_decommittment_fsm &s, const zkc_decommit_query &req, bool can_pop, bool queue_empty_after,
                                         const uint32_t *__restrict__ words, size_t n_words, size_t &word_cursor,
                                         uint64_t *__restrict__ push_enc, uint32_t *__restrict__ slot_meta, uint32_t &push_ordinal,
                                         uint64_t *__restrict__ trace, size_t limit, size_t row, uint32_t &checks) {
#define TR(col) trace[(size_t)(col) * limit + row]
    const bool wr = trace != nullptr;
    if (wr) { TR(ZKC_CU_FLAGS_IN + 0) = s.state_get_from_queue; TR(ZKC_CU_FLAGS_IN + 1) = s.state_decommit; TR(ZKC_CU_FLAGS_IN + 2) = s.finished; }
    if (s.state_get_from_queue && !can_pop) { checks |= ZKC_CU_CHK_WITNESS_EXHAUSTED; s.state_get_from_queue = 0; }
    const bool get = s.state_get_from_queue;
    const uint32_t top = req.code_hash[7];
    const bool version_matches = (top >> 16) == ZKC_CODE_HASH_VERSION_TOP16;
    if (get && !version_matches) checks |= ZKC_CU_CHK_VERSION;  // :202-204
    const uint32_t length_in_words = get ? (top & 0xFFFFu) : 1u;
    if ((length_in_words + 1) & 1u) checks |= ZKC_CU_CHK_LENGTH;  // :215-221
    const uint32_t length_in_rounds = (length_in_words + 1) >> 1;
    if (get) {  // :233-275
        s.num_rounds_left = length_in_rounds;
        s.length_in_bits = length_in_words * 256u;
        s.timestamp = req.timestamp;
        s.current_page = req.page;
#pragma unroll
        for (int i = 0; i < 7; i++) s.hash_to_compare_against[i] = req.code_hash[i];
        s.hash_to_compare_against[7] = 0;
        s.current_index = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s.sha256_inner_state[i] = SHA_IV[i];
    }
    s.state_decommit = s.state_decommit || get;
    s.state_get_from_queue = 0;
    const bool decommit = s.state_decommit;
    if (decommit) s.num_rounds_left = (s.num_rounds_left - 1) & 0xFFFFu;  // :281-287
    const bool last_round = s.num_rounds_left == 0;
    const bool finalize = last_round && decommit, process_second_word = !last_round && decommit;
    if (wr) {
        TR(ZKC_CU_VERSION_MATCHES) = version_matches; TR(ZKC_CU_LENGTH_IN_WORDS) = length_in_words;
        TR(ZKC_CU_LENGTH_IN_ROUNDS) = length_in_rounds; TR(ZKC_CU_LENGTH_IN_BITS) = s.length_in_bits;
        TR(ZKC_CU_TIMESTAMP) = s.timestamp; TR(ZKC_CU_PAGE) = s.current_page;
#pragma unroll
        for (int i = 0; i < 8; i++) TR(ZKC_CU_HASH_TO_COMPARE + i) = s.hash_to_compare_against[i];
        TR(ZKC_CU_DECOMMIT) = decommit; TR(ZKC_CU_NUM_ROUNDS_LEFT) = s.num_rounds_left; TR(ZKC_CU_LAST_ROUND) = last_round;
        TR(ZKC_CU_FINALIZE) = finalize; TR(ZKC_CU_PROCESS_SECOND_WORD) = process_second_word;
    }
    uint32_t m[16];
#pragma unroll
    for (int q = 0; q < 2; q++) {  // :295-352: the two conditional code words and their memory writes
        const bool take = q == 0 ? decommit : process_second_word;
        uint32_t value[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const uint32_t index = s.current_index;
        if (take) {
            if (word_cursor < n_words) {
#pragma unroll
                for (int i = 0; i < 8; i++) value[i] = __ldg(words + 8 * word_cursor + i);
            } else checks |= ZKC_CU_CHK_WITNESS_EXHAUSTED;
            word_cursor++;
            uint64_t e[8];
            mq_encode(s.timestamp, s.current_page, index, 1, value, e);
#pragma unroll
            for (int i = 0; i < 8; i++) push_enc[8 * (size_t)push_ordinal + i] = e[i];
            push_ordinal++;
            s.current_index++;
        }
        slot_meta[2 * row + q] = push_ordinal | (take ? 0x80000000u : 0u);
#pragma unroll
        for (int i = 0; i < 8; i++) m[8 * q + i] = value[7 - i];
        if (wr) {
#pragma unroll
            for (int i = 0; i < 8; i++) TR((q ? ZKC_CU_WORD1 : ZKC_CU_WORD0) + i) = value[i];
            TR(q ? ZKC_CU_INDEX1 : ZKC_CU_INDEX0) = index;
        }
    }
    if (finalize) {  // :366-377
        m[8] = 0x80000000u;
#pragma unroll
        for (int i = 9; i < 15; i++) m[i] = 0;
        m[15] = s.length_in_bits;
    }
    uint32_t ns[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ns[i] = s.sha256_inner_state[i];
    if (wr) {
        TR(ZKC_CU_INDEX_OUT) = s.current_index;
#pragma unroll
        for (int i = 0; i < 16; i++) TR(ZKC_CU_MESSAGE + i) = m[i];
#pragma unroll
        for (int i = 0; i < 8; i++) TR(ZKC_CU_STATE_IN + i) = ns[i];
    }
    sha256_compress(ns, m);
    if (decommit) {
#pragma unroll
        for (int i = 0; i < 8; i++) s.sha256_inner_state[i] = ns[i];
    }
    if (finalize) {  // :393-420: hash = [ns7 .. ns1, 0] as little-endian limbs
        bool same = s.hash_to_compare_against[7] == 0;
#pragma unroll
        for (int i = 0; i < 7; i++) same &= ns[7 - i] == s.hash_to_compare_against[i];
        if (!same) checks |= ZKC_CU_CHK_HASH;
    }
    s.finished = s.finished || (queue_empty_after && finalize);
    s.state_get_from_queue = !queue_empty_after && finalize;
    s.state_decommit = process_second_word;
    if (wr) {
#pragma unroll
        for (int i = 0; i < 8; i++) { TR(ZKC_CU_STATE_NEW + i) = ns[i]; TR(ZKC_CU_STATE_OUT + i) = s.sha256_inner_state[i]; }
        TR(ZKC_CU_FLAGS_OUT + 0) = s.state_get_from_queue; TR(ZKC_CU_FLAGS_OUT + 1) = s.state_decommit; TR(ZKC_CU_FLAGS_OUT + 2) = s.finished;
    }
#undef TR
}

// pass 1, one thread per request: pop (verified against the queue witness), then ONLY the SHA-256 chain of its rounds:
// sha_in[row] = the state before the round of that row.  The next round's code words are in flight while this one compresses.
__global__ void __launch_bounds__(128)
cu_chain_kernel(CuDev *d, const zkc_decommit_query *__restrict__ requests, const uint64_t *__restrict__ req_prev,
                const uint32_t *__restrict__ words, const CuPlan *__restrict__ starts, uint32_t *__restrict__ sha_in,
                uint64_t *__restrict__ heads) {
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_units = d->n_units;
    if (u >= n_units) return;
    const size_t limit = d->limit, n_words = d->n_code_words;
    const CuPlan st = starts[u], en = starts[u + 1];
    size_t row = st.cycles;
    if (row >= limit || en.cycles == st.cycles) return;
    uint32_t sha[8], rounds_left, length_in_bits;
    if (u == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) sha[i] = d->s0.sha256_inner_state[i];
        rounds_left = d->s0.num_rounds_left;
        length_in_bits = d->s0.length_in_bits;
#pragma unroll
        for (int i = 0; i < 12; i++) heads[i] = d->rq0.head[i];
    } else {
        const zkc_decommit_query req = cu_load_request(requests + (u - 1));
        uint64_t e[8], sp[12];
        bool hint_ok = true;
        cu_encode_request(req, e);
#pragma unroll
        for (int i = 0; i < 12; i++) {
            const uint64_t h = __ldg(req_prev + 12 * (u - 1) + i);
            sp[i] = i < 8 ? e[i] : h;
            if (u == 1 && h != d->rq0.head[i]) hint_ok = false;
        }
        poseidon2_permute(sp);
#pragma unroll
        for (int i = 0; i < 12; i++) heads[12 * u + i] = sp[i];
        if (u + 1 < n_units && starts[u + 1].cycles < limit) {
#pragma unroll
            for (int i = 0; i < 12; i++) hint_ok &= __ldg(req_prev + 12 * u + i) == sp[i];
        }
        if (!hint_ok) { d->hint_bad = 1; cu_report(d, row, ZKC_CU_CHK_QUEUE_HINT); }
        if (u + 1 == n_units || starts[u + 1].cycles >= limit) {
#pragma unroll
            for (int i = 0; i < 12; i++) d->req_head_final[i] = sp[i];
            d->popped_requests = (uint32_t)u;
        }
        const uint32_t length_in_words = req.code_hash[7] & 0xFFFFu;
        rounds_left = (length_in_words + 1) >> 1;
        length_in_bits = length_in_words * 256u;
#pragma unroll
        for (int i = 0; i < 8; i++) sha[i] = SHA_IV[i];
    }
    size_t wc = st.words;
    uint32_t nxt[16];
    auto fetch = [&](size_t cursor, bool second) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const bool take = (q == 0 || second) && cursor + q < n_words;
#pragma unroll
            for (int i = 0; i < 8; i++) nxt[8 * q + i] = take ? __ldg(words + 8 * (cursor + q) + 7 - i) : 0u;
        }
    };
    fetch(wc, ((rounds_left - 1) & 0xFFFFu) != 0);
    while (row < limit && row < en.cycles) {
        uint4 *o = reinterpret_cast<uint4 *>(sha_in + 8 * row);
        o[0] = make_uint4(sha[0], sha[1], sha[2], sha[3]);
        o[1] = make_uint4(sha[4], sha[5], sha[6], sha[7]);
        rounds_left = (rounds_left - 1) & 0xFFFFu;
        const bool last_round = rounds_left == 0;
        uint32_t m[16];
#pragma unroll
        for (int i = 0; i < 16; i++) m[i] = nxt[i];
        if (last_round) {  // :366-377
            m[8] = 0x80000000u;
#pragma unroll
            for (int i = 9; i < 15; i++) m[i] = 0;
            m[15] = length_in_bits;
        }
        wc += last_round ? 1 : 2;
        row++;
        if (row < limit && row < en.cycles) fetch(wc, ((rounds_left - 1) & 0xFFFFu) != 0);
        sha256_compress(sha, m);
    }
}

// pass 2, one thread per cycle of a request: the FSM state on entry in closed form, then the cycle itself
__global__ void __launch_bounds__(128)
cu_rows_kernel(CuDev *d, const zkc_decommit_query *__restrict__ requests, const uint32_t *__restrict__ words,
               const CuPlan *__restrict__ starts, const uint32_t *__restrict__ sha_in, const uint64_t *__restrict__ heads,
               uint64_t *__restrict__ push_enc, uint32_t *__restrict__ slot_meta, uint64_t *__restrict__ trace) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t limit = d->limit;
    const uint32_t n_units = d->n_units;
    const size_t total = starts[n_units].cycles;
    if (row >= limit || row >= total) return;
    // the request this cycle belongs to: the last unit that starts at or before it
    uint32_t lo = 0, hi = n_units - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (starts[mid].cycles <= row) lo = mid; else hi = mid - 1;
    }
    const uint32_t u = lo;
    const CuPlan st = starts[u];
    const uint32_t idx = (uint32_t)(row - st.cycles);
    zkc_code_decommittment_fsm s = d->s0;
    zkc_decommit_query req = cu_zero_request();
    if (u >= 1) {
        req = cu_load_request(requests + (u - 1));
        if (idx == 0) {
            s.state_get_from_queue = 1;
            if (u > 1 || !d->unit0_fresh) s.state_decommit = 0;  // the previous request ended with its finalizing round
        } else {
            const uint32_t length_in_words = req.code_hash[7] & 0xFFFFu;
            s.state_get_from_queue = 0; s.state_decommit = 1;
            s.num_rounds_left = (((length_in_words + 1) >> 1) - idx) & 0xFFFFu;
            s.length_in_bits = length_in_words * 256u;
            s.timestamp = req.timestamp;
            s.current_page = req.page;
#pragma unroll
            for (int i = 0; i < 7; i++) s.hash_to_compare_against[i] = req.code_hash[i];
            s.hash_to_compare_against[7] = 0;
            s.current_index = 2 * idx;
        }
    } else if (idx > 0) {
        s.state_get_from_queue = 0; s.state_decommit = 1;
        s.num_rounds_left = (d->s0.num_rounds_left - idx) & 0xFFFFu;
        s.current_index = d->s0.current_index + 2 * idx;
    }
    if (idx > 0 || u == 0) {
        const uint4 *in = reinterpret_cast<const uint4 *>(sha_in + 8 * row);
        const uint4 a = in[0], b = in[1];
        s.sha256_inner_state[0] = a.x; s.sha256_inner_state[1] = a.y; s.sha256_inner_state[2] = a.z; s.sha256_inner_state[3] = a.w;
        s.sha256_inner_state[4] = b.x; s.sha256_inner_state[5] = b.y; s.sha256_inner_state[6] = b.z; s.sha256_inner_state[7] = b.w;
    }
    const uint32_t len_after = d->rq0.length - u;
    if (trace) {
        for (int i = 0; i < 11; i++) trace[(size_t)(ZKC_CU_REQUEST + i) * limit + row] = (idx == 0 && u >= 1) ? cu_flat_request(req, i) : 0;
        for (int i = 0; i < 12; i++) trace[(size_t)(ZKC_CU_REQ_HEAD + i) * limit + row] = heads[12 * (size_t)u + i];
        trace[(size_t)ZKC_CU_REQ_LEN * limit + row] = len_after;
    }
    size_t word_cursor = (size_t)st.words + 2 * (size_t)idx;
    uint32_t push_ordinal = st.words + 2 * idx;
    uint32_t checks = 0;
    cu_cycle(s, (idx == 0 && u >= 1) ? req : cu_zero_request(), true, len_after == 0, words, d->n_code_words, word_cursor, push_enc,
             slot_meta, push_ordinal, trace, limit, row, checks);
    cu_report(d, row, checks);
    if (row == limit - 1) d->s_final = s;
    if (row == total - 1) d->s_last = s;
}

// the rows after the last request: the FSM idles (or reports that it wanted a request nobody supplied), row-parallel
__global__ void __launch_bounds__(128)
cu_tail_kernel(CuDev *d, const CuPlan *__restrict__ starts, uint32_t *__restrict__ slot_meta, uint64_t *__restrict__ trace,
               uint64_t *__restrict__ push_enc) {
    const size_t limit = d->limit;
    const size_t first = starts[d->n_units].cycles;
    const size_t row = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    zkc_code_decommittment_fsm s = d->s_last;
    if (row > first) s.state_get_from_queue = 0;  // the first tail row reported the missing request; idle afterwards
    uint32_t checks = 0;
    const uint32_t popped = d->popped_requests;
    const uint32_t len_now = d->rq0.length - popped;
    if (trace) {
        for (int i = 0; i < 11; i++) trace[(size_t)(ZKC_CU_REQUEST + i) * limit + row] = 0;
        for (int i = 0; i < 12; i++) trace[(size_t)(ZKC_CU_REQ_HEAD + i) * limit + row] = popped ? d->req_head_final[i] : d->rq0.head[i];
        trace[(size_t)ZKC_CU_REQ_LEN * limit + row] = len_now;
    }
    size_t wc = starts[d->n_units].words;
    uint32_t po = starts[d->n_units].words;
    cu_cycle(s, cu_zero_request(), false, len_now == 0, nullptr, 0, wc, push_enc, slot_meta, po, trace, limit, row, checks);
    cu_report(d, row, checks);
    if (row == limit - 1) d->s_final = s;
}

__global__ void cu_finalize_kernel(CuDev *d, const uint32_t *__restrict__ slot_meta, const uint64_t *__restrict__ states, size_t n_states) {
    // lane 0 does the scalar bookkeeping; the commitments' permutations run on the two 16-lane groups, 12 lanes each
    __shared__ zkc_code_unpacker_fsm out;
    __shared__ uint64_t e_out[80], o_out[32], compact[24];
    __shared__ uint32_t sh_done, sh_n_out;
    const int lane = threadIdx.x & 31, li = lane & 15;
    const unsigned gm = lane < 16 ? 0xFFFFu : 0xFFFF0000u;
    if (lane == 0) {
        zkc_code_unpacker_closed_form &io = d->io;
        const size_t limit = d->limit;
        memset(&out, 0, sizeof out);
        out.internal_fsm = limit ? d->s_final : d->s0;
        zkc_queue_state12 rq = d->rq0;
        const uint32_t popped = limit ? d->popped_requests : 0;
        if (popped) for (int i = 0; i < 12; i++) rq.head[i] = d->req_head_final[i];
        rq.length = d->rq0.length - popped;
        zkc_queue_state12 mq = d->mq0;
        bool hint_bad = d->hint_bad;
        if (limit) {
            const uint32_t pushes = slot_meta[2 * (limit - 1) + 1] & 0x7FFFFFFFu;
            if (pushes) {
                if (pushes - 1 < n_states) for (int i = 0; i < 12; i++) mq.tail[i] = states[12 * (size_t)(pushes - 1) + i];
                else hint_bad = true;
            }
            mq.length += pushes;
        }
        out.decommittment_requests_queue_state = rq;
        out.memory_queue_state = mq;
        uint32_t checks = d->failed_checks;
        if (rq.length == 0)
            for (int i = 0; i < 12; i++) if (rq.head[i] != rq.tail[i]) checks |= ZKC_CU_CHK_QUEUE_CONSISTENCY;  // :449
        const bool done = out.internal_fsm.finished;
        zkc_queue_state12 obs_out;
        memset(&obs_out, 0, sizeof obs_out);
        if (done) obs_out = mq;
        const int n_out = cu_encode_fsm(out, e_out);
        cu_put_q12(o_out, obs_out);
        zkc_status st;
        st.code = ZKC_OK; st.cuda_error = 0; st.first_bad_row = -1; st.failed_checks = checks; st.reserved = 0;
        if (d->first_bad != ~0ull) st.first_bad_row = (int64_t)(d->first_bad >> 16);
        if (checks) st.code = ZKC_ERR_UNSATISFIED;
        if (hint_bad) { st.code = ZKC_ERR_QUEUE_WITNESS_INCONSISTENT; st.failed_checks |= ZKC_CU_CHK_QUEUE_HINT; }
        if (d->opt.compare_expected) {
            uint64_t e_exp[74], o_exp[25];
            cu_encode_fsm(io.hidden_fsm_output, e_exp);
            cu_put_q12(o_exp, io.memory_queue_final_state);
            bool same = (io.completion_flag != 0) == done;
            for (int i = 0; i < n_out; i++) same &= e_out[i] == e_exp[i];
            for (int i = 0; i < 25; i++) same &= o_out[i] == o_exp[i];
            if (!same && st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
        }
        io.hidden_fsm_output = out;
        io.memory_queue_final_state = obs_out;
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
// One thread per cycle re-evaluates every relation unpack_code_into_memory_inner (mod.rs:191-447) places that is local to a cycle
// or to a cycle and its predecessor: the FSM flags carried from the previous cycle, the conditional pop (ranges, queue length /
// head), the versioned-hash decomposition (version match, length in words / rounds / bits), the selects on timestamp / page / hash /
// index / SHA-256 state, the round counter, last_round / finalize / process_second_word, the two conditional code words and their
// indices, the SHA-256 block (padding selected in on finalize) and the compression, the state select, the hash comparison on
// finalize, the next FSM flags, the memory queue's length / tail over the two conditional writes.  With
// ZKC_GATES_ROUND_FUNCTION also the Poseidon2 permutations (the pop, every executed memory write).
template <bool ROUND_FUNCTION>
__global__ void __launch_bounds__(128)
cu_check_kernel(CuDev *d, unsigned long long *violations, const uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const bool first = row == 0;
    const zkc_code_decommittment_fsm &s0 = d->s0;
#define TR(col) __ldg(trace + (size_t)(col) * limit + row)
#define TP(col) __ldg(trace + (size_t)(col) * limit + row - 1)
    uint32_t bad = 0;
    const uint64_t get = TR(ZKC_CU_FLAGS_IN + 0), decommit_in = TR(ZKC_CU_FLAGS_IN + 1), finished_in = TR(ZKC_CU_FLAGS_IN + 2);
    if ((get | decommit_in | finished_in) > 1 || get != (first ? (uint64_t)s0.state_get_from_queue : TP(ZKC_CU_FLAGS_OUT + 0)) ||
        decommit_in != (first ? (uint64_t)s0.state_decommit : TP(ZKC_CU_FLAGS_OUT + 1)) || finished_in != (first ? (uint64_t)s0.finished : TP(ZKC_CU_FLAGS_OUT + 2)))
        bad |= ZKC_CUV_FSM;
    // the conditional pop
    uint64_t f[11], limbs = 0;
#pragma unroll
    for (int i = 0; i < 11; i++) { f[i] = TR(ZKC_CU_REQUEST + i); limbs |= f[i]; }
    if ((limbs >> 32) || f[9] > 1) bad |= ZKC_CUV_BOOLEAN;
    const uint64_t len_prev = first ? d->rq0.length : TP(ZKC_CU_REQ_LEN), len = TR(ZKC_CU_REQ_LEN);
    if (len + get != len_prev || (len >> 32)) bad |= ZKC_CUV_QUEUE;
    {
        uint64_t head[12], st[12];
        bool same = true;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            head[i] = TR(ZKC_CU_REQ_HEAD + i);
            const uint64_t hp = first ? d->rq0.head[i] : TP(ZKC_CU_REQ_HEAD + i);
            same &= head[i] == hp;
            st[i] = hp;
            if (head[i] >= GL_P) bad |= ZKC_CUV_BOOLEAN;
        }
        if (!get && !same) bad |= ZKC_CUV_QUEUE;
        if (ROUND_FUNCTION && get) {
            zkc_decommit_query q;
#pragma unroll
            for (int i = 0; i < 8; i++) q.code_hash[i] = (uint32_t)f[i];
            q.page = (uint32_t)f[8]; q.is_first = (uint32_t)f[9] & 1u; q.timestamp = (uint32_t)f[10]; q._pad = 0;
            uint64_t e[8];
            cu_encode_request(q, e);
#pragma unroll
            for (int i = 0; i < 8; i++) st[i] = e[i];
            poseidon2_permute(st);
#pragma unroll
            for (int i = 0; i < 12; i++) if (st[i] != head[i]) bad |= ZKC_CUV_ROUND_FUNCTION;
        }
    }
    // :198-221 the versioned hash: top 16 bits = version, low 16 bits of limb 7 = length in words (odd), rounds = (words + 1) / 2
    const uint64_t version_matches = TR(ZKC_CU_VERSION_MATCHES), words = TR(ZKC_CU_LENGTH_IN_WORDS), rounds = TR(ZKC_CU_LENGTH_IN_ROUNDS);
    if (version_matches != (uint64_t)((f[7] >> 16) == ZKC_CODE_HASH_VERSION_TOP16) || words != (get ? (f[7] & 0xFFFFu) : 1ull) || 2 * rounds != words + 1 ||
        (rounds >> 16))
        bad |= ZKC_CUV_LENGTH;
    if (get & (1 - (version_matches & 1))) bad |= ZKC_CUV_ENFORCE;
    // :226-275 the selects
    const uint64_t bits = TR(ZKC_CU_LENGTH_IN_BITS), ts = TR(ZKC_CU_TIMESTAMP), page = TR(ZKC_CU_PAGE), index0 = TR(ZKC_CU_INDEX0);
    uint64_t hash_cmp[8];
    {
        bool ok = bits == (get ? (uint64_t)(uint32_t)(words * 256) : (first ? (uint64_t)s0.length_in_bits : TP(ZKC_CU_LENGTH_IN_BITS)));
        ok &= ts == (get ? f[10] : (first ? (uint64_t)s0.timestamp : TP(ZKC_CU_TIMESTAMP)));
        ok &= page == (get ? f[8] : (first ? (uint64_t)s0.current_page : TP(ZKC_CU_PAGE)));
        ok &= index0 == (get ? 0ull : (first ? (uint64_t)s0.current_index : TP(ZKC_CU_INDEX_OUT)));
        uint64_t range = bits | ts | page | index0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            hash_cmp[i] = TR(ZKC_CU_HASH_TO_COMPARE + i);
            range |= hash_cmp[i];
            ok &= hash_cmp[i] == (get ? (i < 7 ? f[i] : 0ull) : (first ? (uint64_t)s0.hash_to_compare_against[i] : TP(ZKC_CU_HASH_TO_COMPARE + i)));
        }
        if (!ok || (range >> 32)) bad |= ZKC_CUV_SELECTS;
    }
    // :277-291 decommit flag, round counter, the three phase flags
    const uint64_t decommit = TR(ZKC_CU_DECOMMIT), rounds_left = TR(ZKC_CU_NUM_ROUNDS_LEFT), last = TR(ZKC_CU_LAST_ROUND), finalize = TR(ZKC_CU_FINALIZE),
                   second = TR(ZKC_CU_PROCESS_SECOND_WORD);
    {
        const uint64_t selected = get ? rounds : (first ? (uint64_t)s0.num_rounds_left : TP(ZKC_CU_NUM_ROUNDS_LEFT));
        if ((decommit | last | finalize | second) > 1 || decommit != (decommit_in | get) || rounds_left != (decommit ? ((selected - 1) & 0xFFFFu) : selected) ||
            last != (uint64_t)(rounds_left == 0) || finalize != (last & decommit) || second != ((1 - last) & decommit))
            bad |= ZKC_CUV_FSM;
    }
    // :293-352 the two code words, their indices, the two conditional memory writes
    uint32_t w0[8], w1[8], m[16];
    {
        uint64_t r0 = 0, r1 = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { const uint64_t a = TR(ZKC_CU_WORD0 + i), b = TR(ZKC_CU_WORD1 + i); r0 |= a; r1 |= b; w0[i] = (uint32_t)a; w1[i] = (uint32_t)b; }
        if (((r0 | r1) >> 32) || (!decommit && r0) || (!second && r1)) bad |= ZKC_CUV_BOOLEAN;  // conditionally_allocate: zero when not taken
    }
    const uint64_t index1 = TR(ZKC_CU_INDEX1), index_out = TR(ZKC_CU_INDEX_OUT);
    if (index1 != (uint64_t)(uint32_t)(index0 + decommit) || index_out != (uint64_t)(uint32_t)(index1 + second)) bad |= ZKC_CUV_SELECTS;
    {
        uint64_t mt_prev[12], ml_prev = first ? d->mq0.length : TP(ZKC_CU_MEM_TAIL1 + 12);
#pragma unroll
        for (int i = 0; i < 12; i++) mt_prev[i] = first ? d->mq0.tail[i] : TP(ZKC_CU_MEM_TAIL1 + i);
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int b = q ? ZKC_CU_MEM_TAIL1 : ZKC_CU_MEM_TAIL0;
            const uint64_t pushed = q ? second : decommit;
            uint64_t mt[12], st[12];
            bool msame = true;
#pragma unroll
            for (int i = 0; i < 12; i++) { mt[i] = TR(b + i); msame &= mt[i] == mt_prev[i]; if (mt[i] >= GL_P) bad |= ZKC_CUV_BOOLEAN; }
            const uint64_t ml = TR(b + 12);
            if (ml != ml_prev + pushed || (!pushed && !msame)) bad |= ZKC_CUV_MEMORY_QUEUE;
            if (ROUND_FUNCTION && pushed) {
                mq_encode((uint32_t)ts, (uint32_t)page, (uint32_t)(q ? index1 : index0), 1, q ? w1 : w0, st);
#pragma unroll
                for (int i = 8; i < 12; i++) st[i] = mt_prev[i];
                poseidon2_permute(st);
#pragma unroll
                for (int i = 0; i < 12; i++) if (st[i] != mt[i]) bad |= ZKC_CUV_ROUND_FUNCTION;
            }
#pragma unroll
            for (int i = 0; i < 12; i++) mt_prev[i] = mt[i];
            ml_prev = ml;
        }
    }
    // :354-391 the SHA-256 block (big-endian words; the padding block selected in on finalize), the compression, the state select
#pragma unroll
    for (int i = 0; i < 8; i++) {
        m[i] = w0[7 - i];
        const uint32_t pad = i == 0 ? 0x80000000u : (i == 7 ? (uint32_t)bits : 0u);
        m[8 + i] = finalize ? pad : w1[7 - i];
    }
#pragma unroll
    for (int i = 0; i < 16; i++) if (TR(ZKC_CU_MESSAGE + i) != m[i]) bad |= ZKC_CUV_COMPRESSION;
    uint32_t cur[8];
    {
        uint64_t range = 0;
        uint64_t st_in[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            st_in[i] = TR(ZKC_CU_STATE_IN + i);
            range |= st_in[i];
            if (st_in[i] != (get ? (uint64_t)SHA_IV[i] : (first ? (uint64_t)s0.sha256_inner_state[i] : TP(ZKC_CU_STATE_OUT + i)))) bad |= ZKC_CUV_COMPRESSION;
            cur[i] = (uint32_t)st_in[i];
        }
        if (range >> 32) bad |= ZKC_CUV_BOOLEAN;
        sha256_compress(cur, m);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (TR(ZKC_CU_STATE_NEW + i) != cur[i]) bad |= ZKC_CUV_COMPRESSION;
            if (TR(ZKC_CU_STATE_OUT + i) != (decommit ? (uint64_t)cur[i] : st_in[i])) bad |= ZKC_CUV_COMPRESSION;
        }
    }
    // :393-420 on finalize the digest (top 4 bytes ignored) is the hash of the request
    if (finalize) {
#pragma unroll
        for (int i = 0; i < 8; i++) if (hash_cmp[i] != (i < 7 ? (uint64_t)cur[7 - i] : 0ull)) bad |= ZKC_CUV_ENFORCE;
    }
    // :422-430 the next FSM flags
    {
        const uint64_t empty = len == 0;
        const uint64_t o_get = TR(ZKC_CU_FLAGS_OUT + 0), o_decommit = TR(ZKC_CU_FLAGS_OUT + 1), o_finished = TR(ZKC_CU_FLAGS_OUT + 2);
        if ((o_get | o_decommit | o_finished) > 1 || o_get != ((1 - empty) & finalize) || o_decommit != second || o_finished != (finished_in | (empty & finalize)))
            bad |= ZKC_CUV_FSM;
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

extern "C" int zkc_code_unpacker_entry_point(zkc_ctx *ctx, zkc_code_unpacker_closed_form *io, const zkc_decommit_query *requests,
                                             const uint64_t *requests_prev_states, size_t n_requests, const uint32_t *code_words,
                                             size_t n_code_words, const uint64_t *memory_states, size_t n_memory_states, size_t limit,
                                             const zkc_sorter_options *options, int on_device, uint64_t *trace,
                                             uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !commitment || (n_requests && (!requests || !requests_prev_states)) || (n_code_words && !code_words) ||
        limit > 0x0FFFFFFFull) {
        status->code = ZKC_ERR_INVALID_ARGUMENT;
        return ZKC_ERR_INVALID_ARGUMENT;
    }
    const bool in_dev = on_device & ZKC_INPUTS_ON_DEVICE, trace_dev = on_device & ZKC_TRACE_ON_DEVICE;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const size_t max_units = n_requests + 1;
    const size_t tiles = (max_units + SCAN_THREADS - 1) / SCAN_THREADS;
    const bool have_states = memory_states != nullptr;
    const size_t max_pushes = 2 * limit + 1;
    if (!have_states) n_memory_states = max_pushes;
    size_t bytes = zkc_carver::bytes(1, sizeof(CuDev)) + zkc_carver::bytes(1, sizeof(ScanGlobal)) +
                   zkc_carver::bytes(tiles + 1, sizeof(TileStateT<CuPlan>)) + zkc_carver::bytes(max_units + 2, sizeof(CuPlan)) +
                   zkc_carver::bytes(max_pushes * 8, 8) + zkc_carver::bytes(2 * limit + 8, 4) + zkc_carver::bytes(8 * limit + 8, 4) +
                   zkc_carver::bytes(12 * (max_units + 1), 8);
    if (!in_dev) bytes += zkc_carver::bytes(n_requests + 1, sizeof(zkc_decommit_query)) + zkc_carver::bytes(n_requests * 12 + 12, 8) +
                          zkc_carver::bytes(n_code_words * 8 + 8, 4);
    if (!in_dev || !have_states) bytes += zkc_carver::bytes(n_memory_states * 12 + 12, 8);
    if (trace && !trace_dev) bytes += zkc_carver::bytes((size_t)ZKC_CU_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    CuDev *h = (CuDev *)ctx->pinned(sizeof(CuDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    CuDev *d = cv.take<CuDev>(1);
    char *zero_begin = cv.base + cv.off;
    ScanGlobal *sg = cv.take<ScanGlobal>(1);
    TileStateT<CuPlan> *ts = cv.take<TileStateT<CuPlan>>(tiles + 1);
    char *zero_end = cv.base + cv.off;
    CuPlan *starts = cv.take<CuPlan>(max_units + 2);
    uint64_t *push_enc = cv.take<uint64_t>(max_pushes * 8);
    uint32_t *slot_meta = cv.take<uint32_t>(2 * limit + 8);
    uint32_t *sha_in = cv.take<uint32_t>(8 * limit + 8);
    uint64_t *heads = cv.take<uint64_t>(12 * (max_units + 1));
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(CuDev));
    h->io = *io;
    if (options) h->opt = *options;
    h->n_requests = n_requests; h->n_code_words = n_code_words; h->n_memory_states = n_memory_states; h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(CuDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(zero_begin, 0, zero_end - zero_begin, s));
    const zkc_decommit_query *dreq = requests;
    const uint64_t *dprev = requests_prev_states, *dstates = memory_states;
    const uint32_t *dwords = code_words;
    uint64_t *dtrace = trace;
    if (!in_dev) {
        zkc_decommit_query *br = cv.take<zkc_decommit_query>(n_requests + 1);
        uint64_t *bp = cv.take<uint64_t>(n_requests * 12 + 12);
        uint32_t *bw = cv.take<uint32_t>(n_code_words * 8 + 8);
        if (n_requests) {
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(br, requests, n_requests * sizeof(zkc_decommit_query), cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bp, requests_prev_states, n_requests * 96, cudaMemcpyHostToDevice, s));
        }
        if (n_code_words) ZKC_CUDA(ctx, status, cudaMemcpyAsync(bw, code_words, n_code_words * 32, cudaMemcpyHostToDevice, s));
        dreq = br; dprev = bp; dwords = bw;
    }
    if (!in_dev || !have_states) {
        uint64_t *bs = cv.take<uint64_t>(n_memory_states * 12 + 12);
        if (have_states && n_memory_states)
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bs, memory_states, n_memory_states * 96, cudaMemcpyHostToDevice, s));
        dstates = bs;
    }
    if (trace && !trace_dev) dtrace = cv.take<uint64_t>((size_t)ZKC_CU_NUM_COLS * limit);

    ZKC_LAUNCH(ctx, "cu_prologue", cu_prologue_kernel, 1, 96, 0, d);
    ZKC_LAUNCH(ctx, "cu_plan", cu_plan_kernel, (unsigned)tiles, SCAN_THREADS, 0, d, dreq, starts, sg, ts);
    if (limit) {
        ZKC_LAUNCH(ctx, "cu_chain", cu_chain_kernel, (unsigned)((max_units + 127) / 128), 128, 0, d, dreq, dprev, dwords, starts, sha_in,
                   heads);
        ZKC_LAUNCH(ctx, "cu_rows", cu_rows_kernel, (unsigned)((limit + 127) / 128), 128, 0, d, dreq, dwords, starts, sha_in, heads,
                   push_enc, slot_meta, dtrace);
        ZKC_LAUNCH(ctx, "cu_tail", cu_tail_kernel, (unsigned)((limit + 127) / 128), 128, 0, d, starts, slot_meta, dtrace, push_enc);
        if (!have_states) ZKC_LAUNCH(ctx, "cu_mem_chain", (pc_mem_chain_kernel<CuDev, 2>), 1, 32, 0, d, push_enc, slot_meta, (uint64_t *)dstates);
        ZKC_LAUNCH(ctx, "cu_memq", (pc_memq_kernel<CuDev, 2, ZKC_CU_MEM_TAIL0, 13, ZKC_CU_MEM_TAIL1, ZKC_CU_CHK_QUEUE_HINT>),
                   (unsigned)((2 * limit + 255) / 256), 256, 0, d, push_enc, slot_meta, dstates, n_memory_states, have_states, dtrace);
    }
    ZKC_LAUNCH(ctx, "cu_finalize", cu_finalize_kernel, 1, 32, 0, d, slot_meta, dstates, n_memory_states);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(CuDev), cudaMemcpyDeviceToHost, s));
    if (!trace_dev && trace && limit)
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(trace, dtrace, (size_t)ZKC_CU_NUM_COLS * limit * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    io->hidden_fsm_output = h->io.hidden_fsm_output;
    io->memory_queue_final_state = h->io.memory_queue_final_state;
    io->completion_flag = h->io.completion_flag;
    memcpy(commitment, h->commitment, 32);
    *status = h->status;
    return status->code;
}

extern "C" int zkc_code_unpacker_check_trace(zkc_ctx *ctx, const zkc_code_unpacker_closed_form *io, const uint64_t *trace, size_t limit, uint32_t gates,
                                             int on_device, uint64_t *violations, zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !violations || (limit && !trace)) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    size_t bytes = zkc_carver::bytes(1, sizeof(CuDev)) + zkc_carver::bytes(1, 8);
    if (!on_device) bytes += zkc_carver::bytes((size_t)ZKC_CU_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    CuDev *h = (CuDev *)ctx->pinned(sizeof(CuDev) + 8);
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    CuDev *d = cv.take<CuDev>(1);
    unsigned long long *dviol = cv.take<unsigned long long>(1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(CuDev));
    h->io = *io;
    h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(CuDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(dviol, 0, 8, s));
    const uint64_t *dt = trace;
    if (!on_device && limit) {
        uint64_t *b = cv.take<uint64_t>((size_t)ZKC_CU_NUM_COLS * limit);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(b, trace, (size_t)ZKC_CU_NUM_COLS * limit * 8, cudaMemcpyHostToDevice, s));
        dt = b;
    }
    ZKC_LAUNCH(ctx, "cu_prologue", cu_prologue_kernel, 1, 96, 0, d);
    if (limit) {
        const unsigned grid = (unsigned)((limit + 127) / 128);
        if (gates == 0 || (gates & ZKC_GATES_ROUND_FUNCTION)) ZKC_LAUNCH(ctx, "cu_check_rf", cu_check_kernel<true>, grid, 128, 0, d, dviol, dt);
        else ZKC_LAUNCH(ctx, "cu_check", cu_check_kernel<false>, grid, 128, 0, d, dviol, dt);
    }
    ZKC_CUDA(ctx, status, cudaGetLastError());
    unsigned long long *hviol = (unsigned long long *)(h + 1);
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(CuDev), cudaMemcpyDeviceToHost, s));
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
