// Code-decommitment request sorter / deduplicator on sm_100a:
// sort_and_deduplicate_code_decommittments_entry_point (/root/reference/src/sort_decommittment_requests/mod.rs:40-233)
// and its loop sort_and_deduplicate_code_decommittments_inner (:235-381), one thread per loop iteration.
// The sequential state of the loop is recovered row-parallel:
//   - queue heads (full-state queues, 12 elements): from the previous-state column of the raw queue witness
//     (input.rs:114-131), verified link by link, exactly as in ram_permutation.cu;
//   - previous record / packed key: the neighbouring row's sorted item;
//   - running grand products, number of executed result pushes and `first_encountered_timestamp`: ONE decoupled
//     look-back scan over {4 products, push counter, "last row that started a new hash"} -- the timestamp the
//     reference carries from row to row (:345-350) is the timestamp of the row that opened the current run of
//     equal hashes, i.e. a "last setter" index, which is associative;
//   - the RESULT queue is a full-state hash chain over the executed pushes (tail' = P(enc || tail[8..12])): verified
//     against host-supplied states (`result_states`) row-parallel, or rebuilt by a sequential chain kernel.
#include "ctx.cuh"
#include "poseidon2.cuh"
#include "scan.cuh"

namespace zkc {

struct DqDev {
    zkc_decommit_sorter_closed_form io;
    zkc_sorter_options opt;
    uint64_t n_unsorted, n_sorted, n_result_states, limit;
    // prologue
    uint64_t ch[2][9];
    uint64_t acc0[4];  // rep*2 + side
    uint32_t start, prev_trivial0, first_ts0, prologue_checks;
    zkc_queue_state12 uq0, sq0, rq0;
    zkc_decommit_query previous_record0;
    uint32_t previous_packed_key0[ZKC_DQ_PACKED_KEY_LENGTH], pad0;
    uint64_t commit_obs_in[4], commit_fsm_in[4];
    // rows
    uint64_t acc_final[4];
    uint32_t pushes_in_loop, last_final;  // last_final: 1 + last row whose hash differs from its predecessor's (0: none)
    uint64_t head_final[2][12];
    // status
    unsigned long long first_bad;
    uint32_t failed_checks, hint_bad;
    // finalize
    uint64_t commitment[4];
    zkc_status status;
};

// scan element: grand products, executed pushes, last run start
struct DqVal {
    uint64_t p[4];
    uint32_t c;
    uint32_t last;
};
struct DqValOp {
    static __device__ __forceinline__ DqVal identity() { return DqVal{{1, 1, 1, 1}, 0, 0}; }
    static __device__ __forceinline__ DqVal combine(const DqVal &a, const DqVal &b) {
        DqVal r;
#pragma unroll
        for (int i = 0; i < 4; i++) r.p[i] = gl_mul(a.p[i], b.p[i]);
        r.c = a.c + b.c;
        r.last = b.last ? b.last : a.last;
        return r;
    }
};
using DqTile = TileStateT<DqVal>;
using DqShared = ScanSharedT<DqVal>;

__device__ __forceinline__ zkc_decommit_query dq_load(const zkc_decommit_query *p) {
    zkc_decommit_query q;
    const uint4 *s = reinterpret_cast<const uint4 *>(p);
    uint4 *d = reinterpret_cast<uint4 *>(&q);
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = __ldg(s + i);
    q.is_first &= 1u;
    q._pad = 0;
    return q;
}
__device__ __forceinline__ zkc_decommit_query dq_zero() {
    zkc_decommit_query q;
    uint4 *d = reinterpret_cast<uint4 *>(&q);
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = make_uint4(0, 0, 0, 0);
    return q;
}
// DecommitQuery::encode, decommit_query/mod.rs:31-107
__device__ __forceinline__ void dq_encode(const zkc_decommit_query &q, uint64_t (&e)[8]) {
    e[0] = (uint64_t)q.code_hash[0] | ((uint64_t)(q.page & 0xFFFFFFu) << 32);
    e[1] = (uint64_t)q.code_hash[1] | ((uint64_t)(q.page >> 24) << 32) | ((uint64_t)(q.timestamp & 0xFFFFu) << 40);
    e[2] = (uint64_t)q.code_hash[2] | ((uint64_t)(q.timestamp >> 16) << 32) | ((uint64_t)(q.is_first & 1u) << 48);
#pragma unroll
    for (int i = 3; i < 8; i++) e[i] = q.code_hash[i];
}
// flatten_as_variables, decommit_query/mod.rs:133-150
__device__ __forceinline__ uint64_t dq_flat(const zkc_decommit_query &q, int i) {
    return i < 8 ? q.code_hash[i] : i == 8 ? q.page : i == 9 ? (q.is_first & 1u) : q.timestamp;
}

static __device__ int dq_put_queue_state12(uint64_t *dst, const zkc_queue_state12 &s) {
    for (int i = 0; i < 12; i++) dst[i] = s.head[i];
    for (int i = 0; i < 12; i++) dst[12 + i] = s.tail[i];
    dst[24] = s.length;
    return 25;
}
// CSVarLengthEncodable order of CodeDecommittmentsDeduplicatorFSMInputOutput, input.rs:26-38
static __device__ int dq_encode_fsm(const zkc_decommit_sorter_fsm &f, uint64_t *dst) {
    int n = dq_put_queue_state12(dst, f.initial_queue_state);
    n += dq_put_queue_state12(dst + n, f.sorted_queue_state);
    n += dq_put_queue_state12(dst + n, f.final_queue_state);
    dst[n++] = f.lhs_accumulator[0]; dst[n++] = f.lhs_accumulator[1];
    dst[n++] = f.rhs_accumulator[0]; dst[n++] = f.rhs_accumulator[1];
    for (int i = 0; i < ZKC_DQ_PACKED_KEY_LENGTH; i++) dst[n++] = f.previous_packed_key[i];
    dst[n++] = f.first_encountered_timestamp;
    for (int i = 0; i < ZKC_DECOMMIT_QUERY_FLAT; i++) dst[n++] = dq_flat(f.previous_record, i);
    return n;  // 100
}

// three warps, one 16-lane group each, every permutation spread over 12 lanes (poseidon2_permute_coop):
// warp 0: start selection + Fiat-Shamir challenges, warp 1 / 2: commitments to the observable input / FSM input
__global__ void dq_prologue_kernel(DqDev *d) {
    __shared__ uint64_t buf[3][104];
    const int warp = threadIdx.x >> 5, i = threadIdx.x & 31;
    if (i >= 16) return;
    const unsigned gm = 0xFFFFu;
    const zkc_decommit_sorter_closed_form &io = d->io;
    if (warp == 0) {
        if (i == 0) {
            const bool start = io.start_flag != 0;
            const zkc_decommit_sorter_fsm &f = io.hidden_fsm_input;
            d->start = start;
            d->uq0 = start ? io.initial_queue_state : f.initial_queue_state;
            d->sq0 = start ? io.sorted_queue_initial_state : f.sorted_queue_state;
            zkc_queue_state12 empty;
            memset(&empty, 0, sizeof empty);
            d->rq0 = start ? empty : f.final_queue_state;  // :104-114
            for (int k = 0; k < 2; k++) {
                d->acc0[k * 2 + 0] = start ? 1 : f.lhs_accumulator[k];
                d->acc0[k * 2 + 1] = start ? 1 : f.rhs_accumulator[k];
            }
            zkc_decommit_query pr = start ? dq_zero() : f.previous_record;  // :150-156
            pr.is_first &= 1u; pr._pad = 0;
            d->previous_record0 = pr;
            for (int k = 0; k < ZKC_DQ_PACKED_KEY_LENGTH; k++) d->previous_packed_key0[k] = start ? 0 : f.previous_packed_key[k];
            d->first_ts0 = start ? 0 : f.first_encountered_timestamp;
            d->prev_trivial0 = (d->uq0.length == 0) || start;  // :271-273
            uint32_t checks = 0;
            for (int k = 0; k < 12; k++)
                if (io.initial_queue_state.head[k] | io.sorted_queue_initial_state.head[k]) checks |= ZKC_DQ_CHK_TRIVIAL_HEAD;
            if (d->uq0.length != d->sq0.length) checks |= ZKC_DQ_CHK_LENGTHS_EQUAL;
            d->prologue_checks = checks;
            // produce_fs_challenges<_, 12, 9, 2>, utils.rs:12-78, over tail || len || tail || len (26 elements)
            uint64_t *in = buf[0];
            for (int k = 0; k < 12; k++) in[k] = io.initial_queue_state.tail[k];
            in[12] = io.initial_queue_state.length;
            for (int k = 0; k < 12; k++) in[13 + k] = io.sorted_queue_initial_state.tail[k];
            in[25] = io.sorted_queue_initial_state.length;
        }
        __syncwarp(gm);
        fs_challenges_coop(gm, buf[0], 26, 9, &d->ch[0][0], i);
    } else if (warp == 1) {
        int n = 0;
        if (i == 0) {
            n = dq_put_queue_state12(buf[1], io.initial_queue_state);
            n += dq_put_queue_state12(buf[1] + n, io.sorted_queue_initial_state);
        }
        __syncwarp(gm);
        n = __shfl_sync(gm, n, 0, 16);
        const uint64_t c = commit_encoding_coop(gm, buf[1], n, i);
        if (i < 4) d->commit_obs_in[i] = c;
    } else if (warp == 2) {
        int n = 0;
        if (i == 0) n = dq_encode_fsm(io.hidden_fsm_input, buf[2]);
        __syncwarp(gm);
        n = __shfl_sync(gm, n, 0, 16);
        const uint64_t c = commit_encoding_coop(gm, buf[2], n, i);
        if (i < 4) d->commit_fsm_in[i] = c;
    }
}

__device__ __forceinline__ void dq_report(DqDev *d, size_t row, uint32_t checks) {
    if (!checks) return;
    atomicOr(&d->failed_checks, checks);
    atomicMin(&d->first_bad, ((unsigned long long)row << 16) | checks);
}

// first_encountered_timestamp for a "last run start" value of the scan (0: still the FSM input's run)
__device__ __forceinline__ uint32_t dq_first_ts(const DqDev *d, const zkc_decommit_query *sorted, uint32_t last) {
    if (last == 0) return d->first_ts0;
    const size_t r = last - 1;
    // a run can only start on a row that popped, or on the first trivial row after a non-zero hash (timestamp 0)
    return (r < d->uq0.length && r < d->n_sorted) ? __ldg(&sorted[r].timestamp) : 0u;
}

// ---- the row kernel: pops, grand product, ordering, deduplication, what to push -------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS)
dq_rows_kernel(DqDev *d, const zkc_decommit_query *__restrict__ unsorted, const uint64_t *__restrict__ uprev,
               const zkc_decommit_query *__restrict__ sorted, const uint64_t *__restrict__ sprev,
               uint64_t *__restrict__ trace, uint64_t *__restrict__ penc_out, uint32_t *__restrict__ meta,
               uint32_t *__restrict__ push_row, ScanGlobal *sg, DqTile *tiles) {
    __shared__ DqShared sh;
    __shared__ uint64_t ch[2][9];
    if (threadIdx.x < 18) ch[threadIdx.x / 9][threadIdx.x % 9] = d->ch[threadIdx.x / 9][threadIdx.x % 9];
    const unsigned int tile = scan_take_ticket(sg, sh);
    const size_t limit = d->limit;
    const size_t row = (size_t)tile * SCAN_THREADS + threadIdx.x;
    const bool in_range = row < limit;
    const uint32_t ulen0 = d->uq0.length, slen0 = d->sq0.length;
    const bool o_empty = row >= ulen0, s_empty = row >= slen0;
    const bool should_pop = in_range && !o_empty;
    const size_t active_rows = limit < ulen0 ? limit : ulen0;
    uint32_t checks = 0;
    if (in_range && o_empty != s_empty) checks |= ZKC_DQ_CHK_EMPTY_SYNC;
#define TR(col) trace[(size_t)(col) * limit + row]
    const bool wr = in_range && trace != nullptr;
    zkc_decommit_query si = dq_zero();
    uint64_t contrib[4];
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        const zkc_decommit_query *recs = k ? sorted : unsorted;
        const uint64_t *prev = k ? sprev : uprev;
        const size_t n_rec = k ? d->n_sorted : d->n_unsorted;
        const zkc_queue_state12 &q0 = k ? d->sq0 : d->uq0;
        zkc_decommit_query it = dq_zero();
        if (should_pop && row < n_rec) it = dq_load(recs + row);
        uint64_t e[8], s[12];
        dq_encode(it, e);
        if (should_pop) {
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = e[i];
            bool hint_ok = true;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const uint64_t h = __ldg(prev + 12 * row + i);
                if (i >= 8) s[i] = h;
                if (row == 0 && h != q0.head[i]) hint_ok = false;
            }
            poseidon2_permute(s);
            if (row + 1 < active_rows) {
#pragma unroll
                for (int i = 0; i < 12; i++) hint_ok &= __ldg(prev + 12 * (row + 1) + i) == s[i];
            } else {
#pragma unroll
                for (int i = 0; i < 12; i++) d->head_final[k][i] = s[i];
            }
            if (!hint_ok) { checks |= ZKC_DQ_CHK_QUEUE_HINT; d->hint_bad = 1; }
        } else {
            // nothing popped: rows past the end of the queue see the drained queue, whose head equals its tail
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = ulen0 == 0 ? q0.head[i] : q0.tail[i];
        }
        if (wr) {
            const int base = k ? ZKC_DQ_SORTED_ITEM : ZKC_DQ_UNSORTED_ITEM;
#pragma unroll
            for (int i = 0; i < 11; i++) TR(base + i) = dq_flat(it, i);
#pragma unroll
            for (int i = 0; i < 8; i++) TR(base + 11 + i) = e[i];
#pragma unroll
            for (int i = 0; i < 12; i++) TR(base + 19 + i) = s[i];
            const uint32_t len0 = k ? slen0 : ulen0;
            const size_t popped_now = row + 1 < active_rows ? row + 1 : active_rows;
            TR(base + 31) = len0 >= popped_now ? len0 - (uint32_t)popped_now : 0;
        }
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            uint64_t c = ch[rep][8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                c = gl_fma(e[i], ch[rep][i], c);
                if (wr) TR(ZKC_DQ_GP_CHAIN + (rep * 2 + k) * 8 + i) = c;
            }
            contrib[rep * 2 + k] = c;
        }
        if (k == 1) si = it;
    }

    // ---- :306-341 ordering by (hash, timestamp), first-marker and page rules, what to push ------------------------
    zkc_decommit_query pq;
    uint32_t prev_key[ZKC_DQ_PACKED_KEY_LENGTH];
    bool previous_is_trivial;
    if (row == 0) {
        pq = d->previous_record0;
#pragma unroll
        for (int i = 0; i < ZKC_DQ_PACKED_KEY_LENGTH; i++) prev_key[i] = d->previous_packed_key0[i];
        previous_is_trivial = d->prev_trivial0;
    } else {
        pq = dq_zero();
        if (in_range && row - 1 < active_rows && row - 1 < d->n_sorted) pq = dq_load(sorted + row - 1);
        prev_key[0] = pq.timestamp;
#pragma unroll
        for (int i = 0; i < 8; i++) prev_key[1 + i] = pq.code_hash[i];
        previous_is_trivial = row - 1 >= ulen0;
    }
    uint32_t borrow = 0;
    bool keys_equal = true;
#pragma unroll
    for (int i = 0; i < ZKC_DQ_PACKED_KEY_LENGTH; i++) {  // previous - current, least significant limb first
        const uint32_t cur = i == 0 ? si.timestamp : si.code_hash[i - 1];
        const uint64_t dd = (uint64_t)prev_key[i] - cur - borrow;
        const uint32_t diff = (uint32_t)dd;
        borrow = (uint32_t)(dd >> 32) & 1u;
        keys_equal &= diff == 0;
        if (wr) { TR(ZKC_DQ_CMP_DIFF + i) = diff; TR(ZKC_DQ_CMP_BORROW + i) = borrow; TR(ZKC_DQ_CMP_LIMB_EQ + i) = diff == 0; }
    }
    const bool new_key_is_greater = borrow;
    if (should_pop && !new_key_is_greater) checks |= ZKC_DQ_CHK_ORDER;
    bool same_hash = true;
#pragma unroll
    for (int i = 0; i < 8; i++) same_hash &= pq.code_hash[i] == si.code_hash[i];
    const bool different_hash = !same_hash;
    const bool enforce_must_be_first = different_hash && should_pop;
    if (enforce_must_be_first && !si.is_first) checks |= ZKC_DQ_CHK_MUST_BE_FIRST;
    const bool previous_is_non_trivial = !previous_is_trivial;
    const bool enforce_same_memory_page = same_hash && previous_is_non_trivial;
    if (in_range && enforce_same_memory_page && si.page != pq.page) checks |= ZKC_DQ_CHK_SAME_MEMORY_PAGE;
    const bool add = in_range && previous_is_non_trivial && different_hash;

    DqVal v = DqValOp::identity();
    if (should_pop) {
#pragma unroll
        for (int i = 0; i < 4; i++) v.p[i] = contrib[i];
    }
    v.c = add;
    v.last = (in_range && different_hash) ? (uint32_t)row + 1 : 0;
    DqVal init;
#pragma unroll
    for (int i = 0; i < 4; i++) init.p[i] = d->acc0[i];
    init.c = 0; init.last = 0;
    DqVal incl;
    const DqVal excl = scan_tile_generic<DqVal, DqValOp>(v, tile, init, tiles, sh, incl);

    // record_to_add = the PREVIOUS record with the timestamp of the first request of its hash, :338-340
    zkc_decommit_query to_add = pq;
    to_add.is_first = 1;
    to_add.timestamp = in_range ? dq_first_ts(d, sorted, excl.last) : 0;
    uint64_t pe[8];
    dq_encode(to_add, pe);
    if (in_range) {
        meta[row] = (excl.c << 1) | (uint32_t)add;
        if (add) push_row[excl.c] = (uint32_t)row;
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(penc_out + 8 * row);
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = make_ulonglong2(pe[2 * i], pe[2 * i + 1]);
    }

    if (wr) {
        TR(ZKC_DQ_ORIGINAL_IS_EMPTY) = o_empty; TR(ZKC_DQ_SORTED_IS_EMPTY) = s_empty; TR(ZKC_DQ_SHOULD_POP) = should_pop;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            TR(ZKC_DQ_GP_NEW + i) = should_pop ? incl.p[i] : gl_mul(excl.p[i], contrib[i]);
            TR(ZKC_DQ_GP_ACC + i) = incl.p[i];
        }
        TR(ZKC_DQ_KEYS_ARE_EQUAL) = keys_equal; TR(ZKC_DQ_SAME_HASH) = same_hash;
        TR(ZKC_DQ_ENFORCE_MUST_BE_FIRST) = enforce_must_be_first; TR(ZKC_DQ_PREVIOUS_IS_TRIVIAL) = previous_is_trivial;
        TR(ZKC_DQ_ENFORCE_SAME_MEMORY_PAGE) = enforce_same_memory_page; TR(ZKC_DQ_ADD_TO_QUEUE) = add;
#pragma unroll
        for (int i = 0; i < 11; i++) TR(ZKC_DQ_PUSH_ITEM + i) = dq_flat(to_add, i);
#pragma unroll
        for (int i = 0; i < 8; i++) TR(ZKC_DQ_PUSH_ENC + i) = pe[i];
        TR(ZKC_DQ_RESULT_LEN) = d->rq0.length + incl.c;
        TR(ZKC_DQ_FIRST_TIMESTAMP) = different_hash ? si.timestamp : to_add.timestamp;
    }
    if (in_range && row == limit - 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) d->acc_final[i] = incl.p[i];
        d->pushes_in_loop = incl.c;
        d->last_final = incl.last;
    }
    if (in_range) dq_report(d, row, checks);
#undef TR
}

// sequential reconstruction of the result-queue states when the caller does not supply them (1 permutation per push)
__global__ void dq_chain_kernel(const DqDev *d, const uint64_t *__restrict__ penc, const uint32_t *__restrict__ push_row,
                                uint64_t *__restrict__ states) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t s[12];
    for (int i = 0; i < 12; i++) s[i] = d->rq0.tail[i];
    const size_t pushes = d->pushes_in_loop;
    for (size_t k = 0; k < pushes; k++) {
        const size_t row = push_row[k];
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = penc[8 * row + i];
        poseidon2_permute(s);
#pragma unroll
        for (int i = 0; i < 12; i++) states[12 * k + i] = s[i];
    }
}

// the conditional FullStateCircuitQueue::push of every row against the supplied / rebuilt states.  Pushes are sparse
// (one per distinct hash), so the permutations run DENSE over the push index (thread k = k-th executed push, its row
// from push_row), and the per-row trace columns are a plain copy of the state the row leaves behind.
__global__ void __launch_bounds__(256)
dq_push_kernel(DqDev *d, const uint64_t *__restrict__ penc, const uint32_t *__restrict__ meta,
               const uint32_t *__restrict__ push_row, const uint64_t *__restrict__ states, size_t n_states,
               uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= limit) return;
    if (trace) {
        const uint32_t m = meta[t];
        const size_t after = (size_t)(m >> 1) + (m & 1u);  // pushes executed up to and including this row
        const bool have = after == 0 || after - 1 < n_states;
#pragma unroll
        for (int i = 0; i < 12; i++)
            trace[(size_t)(ZKC_DQ_RESULT_TAIL + i) * limit + t] =
                after == 0 ? d->rq0.tail[i] : (have ? __ldg(states + 12 * (after - 1) + i) : 0ull);
    }
    if (t >= d->pushes_in_loop) return;
    const size_t k = t, row = push_row[k];
    uint64_t s[12];
    bool ok = k < n_states;
    const ulonglong2 *in = reinterpret_cast<const ulonglong2 *>(penc + 8 * row);
#pragma unroll
    for (int i = 0; i < 4; i++) { const ulonglong2 a = in[i]; s[2 * i] = a.x; s[2 * i + 1] = a.y; }
    if (k == 0) {
#pragma unroll
        for (int i = 8; i < 12; i++) s[i] = d->rq0.tail[i];
    } else if (k - 1 < n_states) {
#pragma unroll
        for (int i = 8; i < 12; i++) s[i] = __ldg(states + 12 * (k - 1) + i);
    } else {
        ok = false;
#pragma unroll
        for (int i = 8; i < 12; i++) s[i] = 0;
    }
    poseidon2_permute(s);
    if (ok) {
#pragma unroll
        for (int i = 0; i < 12; i++) ok &= __ldg(states + 12 * k + i) == s[i];
    }
    if (!ok) {
        d->hint_bad = 1;
        atomicOr(&d->failed_checks, ZKC_DQ_CHK_QUEUE_HINT);
        atomicMin(&d->first_bad, ((unsigned long long)row << 16) | ZKC_DQ_CHK_QUEUE_HINT);
    }
}

// ---- finalize: last push, consistency, FSM output, commitment -------------------------------------------------------
// one warp: lane 0 does the scalar bookkeeping, the permutations run cooperatively on the two 16-lane groups
__global__ void dq_finalize_kernel(DqDev *d, const zkc_decommit_query *__restrict__ sorted, const uint64_t *__restrict__ states,
                                   size_t n_states) {
    __shared__ zkc_decommit_sorter_fsm out;
    __shared__ zkc_queue_state12 rq, obs_out;
    __shared__ uint64_t e_out[104], o_out[32], compact[24], push_in[16];
    __shared__ uint32_t sh_checks, sh_completed, sh_push, sh_hint_bad;
    const int lane = threadIdx.x & 31, i = lane & 15;
    const unsigned gm = lane < 16 ? 0xFFFFu : 0xFFFF0000u;
    zkc_decommit_sorter_closed_form &io = d->io;
    const size_t limit = d->limit;
    if (lane == 0) {
        const uint32_t len0 = d->uq0.length;
        const size_t popped = limit < len0 ? limit : len0;
        memset(&out, 0, sizeof out);
        out.initial_queue_state = d->uq0;
        out.sorted_queue_state = d->sq0;
        if (popped > 0)
            for (int k = 0; k < 12; k++) {
                out.initial_queue_state.head[k] = d->head_final[0][k];
                out.sorted_queue_state.head[k] = d->head_final[1][k];
            }
        out.initial_queue_state.length = len0 - (uint32_t)popped;
        const size_t spopped = d->sq0.length < popped ? d->sq0.length : popped;
        out.sorted_queue_state.length = d->sq0.length - (uint32_t)spopped;
        zkc_decommit_query previous_record = d->previous_record0;
        uint32_t first_ts = d->first_ts0;
        bool previous_is_trivial = d->prev_trivial0;
        rq = d->rq0;
        bool hint_bad = d->hint_bad;
        for (int k = 0; k < ZKC_DQ_PACKED_KEY_LENGTH; k++) out.previous_packed_key[k] = d->previous_packed_key0[k];
        if (limit > 0) {
            for (int k = 0; k < 2; k++) { out.lhs_accumulator[k] = d->acc_final[2 * k]; out.rhs_accumulator[k] = d->acc_final[2 * k + 1]; }
            previous_record = (limit - 1 < popped && limit - 1 < d->n_sorted) ? dq_load(sorted + limit - 1) : dq_zero();
            out.previous_packed_key[0] = previous_record.timestamp;
            for (int k = 0; k < 8; k++) out.previous_packed_key[1 + k] = previous_record.code_hash[k];
            previous_is_trivial = limit - 1 >= len0;
            first_ts = dq_first_ts(d, sorted, d->last_final);
            const uint32_t pushes = d->pushes_in_loop;
            if (pushes) {
                if (pushes - 1 < n_states) for (int k = 0; k < 12; k++) rq.tail[k] = states[12 * (size_t)(pushes - 1) + k];
                else hint_bad = true;
            }
            rq.length += pushes;
        } else {
            for (int k = 0; k < 2; k++) { out.lhs_accumulator[k] = d->acc0[2 * k]; out.rhs_accumulator[k] = d->acc0[2 * k + 1]; }
        }
        const bool completed = out.initial_queue_state.length == 0;
        uint32_t checks = d->failed_checks | d->prologue_checks;
        if (completed != (out.sorted_queue_state.length == 0)) checks |= ZKC_DQ_CHK_EMPTY_SYNC;  // :360-362
        // finalisation push, :364-375
        const bool push = !previous_is_trivial && completed;
        if (push) {
            zkc_decommit_query to_add = previous_record;
            to_add.is_first = 1;
            to_add.timestamp = first_ts;
            uint64_t pe[8];
            dq_encode(to_add, pe);
            for (int k = 0; k < 12; k++) push_in[k] = k < 8 ? pe[k] : rq.tail[k];
        }
        out.first_encountered_timestamp = first_ts;
        out.previous_record = previous_record;
        sh_checks = checks; sh_completed = completed; sh_push = push; sh_hint_bad = hint_bad;
    }
    __syncwarp();
    if (sh_push && lane < 16) {
        const uint64_t x = poseidon2_permute_coop(gm, i < 12 ? push_in[i] : 0ull, i);
        if (i < 12) rq.tail[i] = x;
    }
    __syncwarp();
    if (lane == 0) {
        const bool completed = sh_completed;
        uint32_t checks = sh_checks;
        if (sh_push) rq.length++;
        out.final_queue_state = rq;
        const zkc_queue_state12 *qs[2] = {&out.initial_queue_state, &out.sorted_queue_state};
        for (int k = 0; k < 2; k++)
            if (qs[k]->length == 0)
                for (int j = 0; j < 12; j++)
                    if (qs[k]->head[j] != qs[k]->tail[j]) checks |= ZKC_DQ_CHK_QUEUE_CONSISTENCY;
        if (completed)
            for (int k = 0; k < 2; k++)
                if (out.lhs_accumulator[k] != out.rhs_accumulator[k]) checks |= ZKC_DQ_CHK_GRAND_PRODUCT;
        memset(&obs_out, 0, sizeof obs_out);
        if (completed) obs_out = rq;
        const int n_out = dq_encode_fsm(out, e_out);
        dq_put_queue_state12(o_out, obs_out);
        zkc_status st;
        st.code = ZKC_OK; st.cuda_error = 0; st.first_bad_row = -1; st.failed_checks = checks; st.reserved = 0;
        if (d->first_bad != ~0ull) st.first_bad_row = (int64_t)(d->first_bad >> 16);
        if (checks) st.code = ZKC_ERR_UNSATISFIED;
        if (sh_hint_bad) { st.code = ZKC_ERR_QUEUE_WITNESS_INCONSISTENT; st.failed_checks |= ZKC_DQ_CHK_QUEUE_HINT; }
        if (d->opt.compare_expected) {
            bool same = (io.completion_flag != 0) == completed;
            uint64_t e_exp[100];
            dq_encode_fsm(io.hidden_fsm_output, e_exp);
            for (int k = 0; k < n_out; k++) same &= e_out[k] == e_exp[k];
            uint64_t o_exp[25];
            dq_put_queue_state12(o_exp, io.final_queue_state);
            for (int k = 0; k < 25; k++) same &= o_out[k] == o_exp[k];
            if (!same && st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
        }
        io.hidden_fsm_output = out;
        io.final_queue_state = obs_out;
        io.completion_flag = completed;
        d->status = st;
    }
    __syncwarp();
    // commitments to the FSM output (group 0) and the observable output (group 1), side by side
    const uint64_t c = commit_encoding_coop(gm, lane < 16 ? e_out : o_out, lane < 16 ? 100 : 25, i);
    const bool completed = sh_completed;
    if (lane < 4) compact[14 + lane] = completed ? 0 : c;
    if (lane >= 16 && lane < 20) compact[6 + lane - 16] = completed ? c : 0;
    if (lane == 0) {
        compact[0] = d->start; compact[1] = completed;
        for (int k = 0; k < 4; k++) {
            compact[2 + k] = d->commit_obs_in[k];
            compact[10 + k] = d->start ? 0 : d->commit_fsm_in[k];
        }
    }
    __syncwarp();
    if (lane < 16) {
        const uint64_t f = commit_encoding_coop(gm, compact, 18, i);
        if (i < 4) d->commitment[i] = f;
    }
}

// FullStateCircuitQueue::push of whole queues: one thread per independent queue
__global__ void decommit_queue_simulate_kernel(const zkc_decommit_query *__restrict__ recs, size_t n_per_queue, size_t n_queues,
                                               uint64_t *__restrict__ prev_states, zkc_queue_state12 *__restrict__ final_states) {
    const size_t qi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= n_queues) return;
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    for (size_t r = 0; r < n_per_queue; r++) {
        const size_t g = qi * n_per_queue + r;
        if (prev_states) {
#pragma unroll
            for (int i = 0; i < 12; i++) prev_states[12 * g + i] = s[i];
        }
        const zkc_decommit_query it = dq_load(recs + g);
        uint64_t e[8];
        dq_encode(it, e);
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = e[i];
        poseidon2_permute(s);
    }
    zkc_queue_state12 &o = final_states[qi];
#pragma unroll
    for (int i = 0; i < 12; i++) { o.head[i] = 0; o.tail[i] = s[i]; }
    o.length = (uint32_t)n_per_queue;
    o._pad = 0;
}


// ---- constraint evaluation of a finished trace ------------------------------------------------------------------------------
// One thread per row re-evaluates every relation the loop body of sort_and_deduplicate_code_decommittments_inner (mod.rs:235-381)
// places that is local to a row or to a row and its predecessor: booleans / ranges of the allocated items, DecommitQuery::encode of
// both pops and of the pushed record, queue-length / head bookkeeping, the 4 x 8 Num::fma chains and the accumulator update, the
// 9-limb key comparison, same-hash / first-marker / same-page flags and their enforcements, the record to add (previous record with
// the first-encountered timestamp), the carried first-encountered timestamp, the result queue's length / tail selection.  Streams
// all ZKC_DQ_NUM_COLS columns once; with ZKC_GATES_ROUND_FUNCTION also the three permutations (two pops, one push).
template <bool ROUND_FUNCTION>
__global__ void __launch_bounds__(128)
dq_check_kernel(DqDev *d, unsigned long long *violations, const uint64_t *__restrict__ trace) {
    __shared__ uint64_t ch[2][9];
    if (threadIdx.x < 18) ch[threadIdx.x / 9][threadIdx.x % 9] = d->ch[threadIdx.x / 9][threadIdx.x % 9];
    __syncthreads();
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const bool first = row == 0;
#define TR(col) __ldg(trace + (size_t)(col) * limit + row)
#define TP(col) __ldg(trace + (size_t)(col) * limit + row - 1)
    uint32_t bad = 0;
    const uint64_t o_empty = TR(ZKC_DQ_ORIGINAL_IS_EMPTY), s_empty = TR(ZKC_DQ_SORTED_IS_EMPTY), should_pop = TR(ZKC_DQ_SHOULD_POP);
    if ((o_empty | s_empty | should_pop) > 1 || o_empty != s_empty || should_pop != 1 - o_empty) bad |= ZKC_DQV_BOOLEAN;
    zkc_decommit_query si = dq_zero();
    uint64_t enc[2][8];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int base = k ? ZKC_DQ_SORTED_ITEM : ZKC_DQ_UNSORTED_ITEM;
        const zkc_queue_state12 &q0 = k ? d->sq0 : d->uq0;
        uint64_t f[11], limbs = 0;
#pragma unroll
        for (int i = 0; i < 11; i++) { f[i] = TR(base + i); limbs |= f[i]; }
        if ((limbs >> 32) || f[9] > 1) bad |= ZKC_DQV_BOOLEAN;
        zkc_decommit_query q = dq_zero();
#pragma unroll
        for (int i = 0; i < 8; i++) q.code_hash[i] = (uint32_t)f[i];
        q.page = (uint32_t)f[8]; q.is_first = (uint32_t)f[9] & 1u; q.timestamp = (uint32_t)f[10];
        uint64_t e[8];
        dq_encode(q, e);
#pragma unroll
        for (int i = 0; i < 8; i++) { enc[k][i] = TR(base + 11 + i); if (enc[k][i] != e[i]) bad |= ZKC_DQV_ENCODING; }
        const uint64_t len_prev = first ? q0.length : TP(base + 31), len = TR(base + 31);
        if ((k ? s_empty : o_empty) != (uint64_t)(len_prev == 0) || len + should_pop != len_prev) bad |= ZKC_DQV_QUEUE_LEN;
        bool same = true;
        uint64_t st[12], head[12];
#pragma unroll
        for (int i = 0; i < 12; i++) {
            head[i] = TR(base + 19 + i);
            const uint64_t hp = first ? q0.head[i] : TP(base + 19 + i);
            same &= head[i] == hp;
            st[i] = i < 8 ? enc[k][i] : hp;
            if (head[i] >= GL_P) bad |= ZKC_DQV_BOOLEAN;
        }
        if (!should_pop && !same) bad |= ZKC_DQV_QUEUE_LEN;
        if (ROUND_FUNCTION && should_pop) {  // head' = P(enc || head[8..12])
            poseidon2_permute(st);
#pragma unroll
            for (int i = 0; i < 12; i++) if (st[i] != head[i]) bad |= ZKC_DQV_ROUND_FUNCTION;
        }
        if (k == 1) si = q;
    }
    // utils.rs:104-135
#pragma unroll
    for (int rep = 0; rep < 2; rep++) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int g = rep * 2 + k;
            uint64_t c = ch[rep][8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint64_t cell = TR(ZKC_DQ_GP_CHAIN + g * 8 + i);
                if (cell != gl_fma(enc[k][i], ch[rep][i], c)) bad |= ZKC_DQV_GP_CHAIN;
                c = cell;
            }
            const uint64_t acc_prev = first ? d->acc0[g] : TP(ZKC_DQ_GP_ACC + g);
            const uint64_t nw = TR(ZKC_DQ_GP_NEW + g), acc = TR(ZKC_DQ_GP_ACC + g);
            if (nw != gl_mul(acc_prev, c) || acc != (should_pop ? nw : acc_prev)) bad |= ZKC_DQV_GP_ACC;
        }
    }
    // the previous record / key / triviality / first-encountered timestamp: the neighbouring row (row 0: the FSM input)
    zkc_decommit_query pq = dq_zero();
    uint32_t prev_key[ZKC_DQ_PACKED_KEY_LENGTH];
    uint64_t prev_trivial, prev_first_ts;
    if (first) {
        pq = d->previous_record0;
#pragma unroll
        for (int i = 0; i < ZKC_DQ_PACKED_KEY_LENGTH; i++) prev_key[i] = d->previous_packed_key0[i];
        prev_trivial = d->prev_trivial0;
        prev_first_ts = d->first_ts0;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) pq.code_hash[i] = (uint32_t)TP(ZKC_DQ_SORTED_ITEM + i);
        pq.page = (uint32_t)TP(ZKC_DQ_SORTED_ITEM + 8); pq.is_first = (uint32_t)TP(ZKC_DQ_SORTED_ITEM + 9) & 1u; pq.timestamp = (uint32_t)TP(ZKC_DQ_SORTED_ITEM + 10);
        prev_key[0] = pq.timestamp;
#pragma unroll
        for (int i = 0; i < 8; i++) prev_key[1 + i] = pq.code_hash[i];
        prev_trivial = TP(ZKC_DQ_ORIGINAL_IS_EMPTY);
        prev_first_ts = TP(ZKC_DQ_FIRST_TIMESTAMP);
    }
    // :309-310 borrow chain, previous - current, least significant limb first: prev + 2^32 * borrow_out = diff + cur + borrow_in
    uint64_t borrow = 0, all_eq = 1;
#pragma unroll
    for (int i = 0; i < ZKC_DQ_PACKED_KEY_LENGTH; i++) {
        const uint64_t cur = i == 0 ? si.timestamp : si.code_hash[i - 1];
        const uint64_t diff = TR(ZKC_DQ_CMP_DIFF + i), bo = TR(ZKC_DQ_CMP_BORROW + i), leq = TR(ZKC_DQ_CMP_LIMB_EQ + i);
        if ((diff >> 32) || bo > 1 || leq != (uint64_t)(diff == 0) || (uint64_t)prev_key[i] + (bo << 32) != diff + cur + borrow) bad |= ZKC_DQV_COMPARISON;
        borrow = bo;
        all_eq &= leq;
    }
    const uint64_t new_key_is_greater = borrow;
    // :314-341 flags
    const uint64_t keys_equal = TR(ZKC_DQ_KEYS_ARE_EQUAL), same_hash = TR(ZKC_DQ_SAME_HASH), must_be_first = TR(ZKC_DQ_ENFORCE_MUST_BE_FIRST),
                   pit = TR(ZKC_DQ_PREVIOUS_IS_TRIVIAL), same_page = TR(ZKC_DQ_ENFORCE_SAME_MEMORY_PAGE), add = TR(ZKC_DQ_ADD_TO_QUEUE);
    bool heq = true;
#pragma unroll
    for (int i = 0; i < 8; i++) heq &= pq.code_hash[i] == si.code_hash[i];
    if ((keys_equal | same_hash | must_be_first | pit | same_page | add) > 1 || keys_equal != all_eq || same_hash != (uint64_t)heq ||
        must_be_first != ((1 - same_hash) & should_pop) || pit != prev_trivial || same_page != (same_hash & (1 - pit)) || add != ((1 - pit) & (1 - same_hash)))
        bad |= ZKC_DQV_FLAGS;
    // conditional enforcements :312, :319-321, :328-333
    if ((should_pop & (1 - (new_key_is_greater & 1))) | (must_be_first & (1 - (uint64_t)(si.is_first & 1u))) | (same_page & (uint64_t)(si.page != pq.page))) bad |= ZKC_DQV_ENFORCE;
    // :338-350 record_to_add = the previous record with is_first = 1 and the first-encountered timestamp of its hash; the carried timestamp
    const uint64_t first_ts = TR(ZKC_DQ_FIRST_TIMESTAMP);
    if (first_ts != (same_hash ? prev_first_ts : (uint64_t)si.timestamp) || (first_ts >> 32)) bad |= ZKC_DQV_FLAGS;
    uint64_t penc[8];
    {
        zkc_decommit_query to_add = pq;
        to_add.is_first = 1;
        to_add.timestamp = (uint32_t)prev_first_ts;
#pragma unroll
        for (int i = 0; i < 11; i++) if (TR(ZKC_DQ_PUSH_ITEM + i) != dq_flat(to_add, i)) bad |= ZKC_DQV_RESULT_QUEUE;
        uint64_t pe[8];
        dq_encode(to_add, pe);
#pragma unroll
        for (int i = 0; i < 8; i++) { penc[i] = TR(ZKC_DQ_PUSH_ENC + i); if (penc[i] != pe[i]) bad |= ZKC_DQV_ENCODING; }
    }
    // the result queue: length counts the pushes; the tail moves only on a push (to P(enc || tail[8..12]))
    {
        const uint64_t len_prev = first ? d->rq0.length : TP(ZKC_DQ_RESULT_LEN);
        if (TR(ZKC_DQ_RESULT_LEN) != len_prev + add) bad |= ZKC_DQV_RESULT_QUEUE;
        uint64_t st[12], tail[12];
        bool same = true;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            tail[i] = TR(ZKC_DQ_RESULT_TAIL + i);
            const uint64_t tp = first ? d->rq0.tail[i] : TP(ZKC_DQ_RESULT_TAIL + i);
            same &= tail[i] == tp;
            st[i] = i < 8 ? penc[i] : tp;
            if (tail[i] >= GL_P) bad |= ZKC_DQV_BOOLEAN;
        }
        if (!add && !same) bad |= ZKC_DQV_RESULT_QUEUE;
        if (ROUND_FUNCTION && add) {
            poseidon2_permute(st);
#pragma unroll
            for (int i = 0; i < 12; i++) if (st[i] != tail[i]) bad |= ZKC_DQV_ROUND_FUNCTION;
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

extern "C" int zkc_decommit_queue_simulate(zkc_ctx *ctx, const zkc_decommit_query *records, size_t n_per_queue, size_t n_queues,
                                           uint64_t *prev_states, zkc_queue_state12 *final_states, int on_device) {
    if (!ctx || !final_states || (n_per_queue && n_queues && !records)) return ZKC_ERR_INVALID_ARGUMENT;
    if (!n_queues) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    const size_t n = n_per_queue * n_queues;
    const zkc_decommit_query *dr = records;
    uint64_t *dp = prev_states;
    zkc_queue_state12 *df = final_states;
    cudaStream_t s = ctx->stream;
    if (!on_device) {
        const size_t bytes = zkc_carver::bytes(n + 1, sizeof(zkc_decommit_query)) + zkc_carver::bytes(n * 12 + 12, 8) +
                             zkc_carver::bytes(n_queues, sizeof(zkc_queue_state12));
        void *blk = ctx->scratch(bytes);
        if (!blk) return ZKC_ERR_CUDA;
        zkc_carver cv(blk);
        zkc_decommit_query *br = cv.take<zkc_decommit_query>(n + 1);
        uint64_t *bp = cv.take<uint64_t>(n * 12 + 12);
        df = cv.take<zkc_queue_state12>(n_queues);
        dp = prev_states ? bp : nullptr;
        if (n) ZKC_CUDA(ctx, st, cudaMemcpyAsync(br, records, n * sizeof(zkc_decommit_query), cudaMemcpyHostToDevice, s));
        dr = br;
    }
    ZKC_LAUNCH(ctx, "decommit_queue_simulate", decommit_queue_simulate_kernel, (unsigned)((n_queues + 31) / 32), 32, 0, dr,
               n_per_queue, n_queues, dp, df);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    if (!on_device) {
        if (prev_states && n) ZKC_CUDA(ctx, st, cudaMemcpyAsync(prev_states, dp, n * 96, cudaMemcpyDeviceToHost, s));
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(final_states, df, n_queues * sizeof(zkc_queue_state12), cudaMemcpyDeviceToHost, s));
        ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    }
    return ZKC_OK;
}

extern "C" int zkc_sort_decommittments_entry_point(zkc_ctx *ctx, zkc_decommit_sorter_closed_form *io,
                                                   const zkc_decommit_query *unsorted, const uint64_t *unsorted_prev_states,
                                                   size_t n_unsorted, const zkc_decommit_query *sorted,
                                                   const uint64_t *sorted_prev_states, size_t n_sorted,
                                                   const uint64_t *result_states, size_t n_result_states, size_t limit,
                                                   const zkc_sorter_options *options, int on_device, uint64_t *trace,
                                                   uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !commitment || (n_unsorted && !unsorted) || (n_sorted && !sorted) || limit > 0x7FFFFFFFull) {
        status->code = ZKC_ERR_INVALID_ARGUMENT;
        return ZKC_ERR_INVALID_ARGUMENT;
    }
    const zkc_queue_state12 &uq = io->start_flag ? io->initial_queue_state : io->hidden_fsm_input.initial_queue_state;
    const size_t need = limit < uq.length ? limit : uq.length;
    if (n_unsorted < need || n_sorted < need || (need && (!unsorted_prev_states || !sorted_prev_states))) {
        status->code = ZKC_ERR_INVALID_ARGUMENT;
        return ZKC_ERR_INVALID_ARGUMENT;
    }
    const bool in_dev = on_device & ZKC_INPUTS_ON_DEVICE, trace_dev = on_device & ZKC_TRACE_ON_DEVICE;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const size_t tiles = (limit + SCAN_THREADS - 1) / SCAN_THREADS;
    const bool have_states = result_states != nullptr;
    if (!have_states) n_result_states = limit + 1;
    size_t bytes = zkc_carver::bytes(1, sizeof(DqDev)) + zkc_carver::bytes(1, sizeof(ScanGlobal)) +
                   zkc_carver::bytes(tiles + 1, sizeof(DqTile)) + zkc_carver::bytes(limit * 8 + 8, 8) +
                   2 * zkc_carver::bytes(limit + 1, 4);
    if (!in_dev) bytes += 2 * zkc_carver::bytes(need + 1, sizeof(zkc_decommit_query)) + 2 * zkc_carver::bytes(need * 12 + 12, 8);
    if (!in_dev || !have_states) bytes += zkc_carver::bytes(n_result_states * 12 + 12, 8);
    if (trace && !trace_dev) bytes += zkc_carver::bytes((size_t)ZKC_DQ_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    DqDev *h = (DqDev *)ctx->pinned(sizeof(DqDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    DqDev *d = cv.take<DqDev>(1);
    ScanGlobal *sg = cv.take<ScanGlobal>(1);
    DqTile *ts = cv.take<DqTile>(tiles + 1);
    uint64_t *penc = cv.take<uint64_t>(limit * 8 + 8);
    uint32_t *meta = cv.take<uint32_t>(limit + 1), *push_row = cv.take<uint32_t>(limit + 1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(DqDev));
    h->io = *io;
    if (options) h->opt = *options;
    h->n_unsorted = n_unsorted; h->n_sorted = n_sorted; h->n_result_states = n_result_states; h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(DqDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(sg, 0, (char *)(ts + tiles + 1) - (char *)sg, s));
    const zkc_decommit_query *du = unsorted, *dsq = sorted;
    const uint64_t *dup = unsorted_prev_states, *dsp = sorted_prev_states, *dstates = result_states;
    uint64_t *dtrace = trace;
    if (!in_dev) {
        zkc_decommit_query *bu = cv.take<zkc_decommit_query>(need + 1), *bs = cv.take<zkc_decommit_query>(need + 1);
        uint64_t *bup = cv.take<uint64_t>(need * 12 + 12), *bsp = cv.take<uint64_t>(need * 12 + 12);
        if (need) {
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bu, unsorted, need * sizeof(zkc_decommit_query), cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bs, sorted, need * sizeof(zkc_decommit_query), cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bup, unsorted_prev_states, need * 96, cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bsp, sorted_prev_states, need * 96, cudaMemcpyHostToDevice, s));
        }
        du = bu; dsq = bs; dup = bup; dsp = bsp;
    }
    if (!in_dev || !have_states) {
        uint64_t *bt = cv.take<uint64_t>(n_result_states * 12 + 12);
        if (have_states && n_result_states)
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bt, result_states, n_result_states * 96, cudaMemcpyHostToDevice, s));
        dstates = bt;
    }
    if (trace && !trace_dev) dtrace = cv.take<uint64_t>((size_t)ZKC_DQ_NUM_COLS * limit);

    ZKC_LAUNCH(ctx, "dq_prologue", dq_prologue_kernel, 1, 96, 0, d);
    if (tiles) {
        ZKC_LAUNCH(ctx, "dq_rows", dq_rows_kernel, (unsigned)tiles, SCAN_THREADS, 0, d, du, dup, dsq, dsp, dtrace, penc, meta, push_row, sg, ts);
        if (!have_states) ZKC_LAUNCH(ctx, "dq_chain", dq_chain_kernel, 1, 32, 0, d, penc, push_row, (uint64_t *)dstates);
        ZKC_LAUNCH(ctx, "dq_push", dq_push_kernel, (unsigned)((limit + 255) / 256), 256, 0, d, penc, meta, push_row, dstates, n_result_states, dtrace);
    }
    ZKC_LAUNCH(ctx, "dq_finalize", dq_finalize_kernel, 1, 32, 0, d, dsq, dstates, n_result_states);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(DqDev), cudaMemcpyDeviceToHost, s));
    if (!trace_dev && trace && limit)
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(trace, dtrace, (size_t)ZKC_DQ_NUM_COLS * limit * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    io->hidden_fsm_output = h->io.hidden_fsm_output;
    io->final_queue_state = h->io.final_queue_state;
    io->completion_flag = h->io.completion_flag;
    memcpy(commitment, h->commitment, 32);
    *status = h->status;
    return status->code;
}

extern "C" int zkc_sort_decommittments_check_trace(zkc_ctx *ctx, const zkc_decommit_sorter_closed_form *io, const uint64_t *trace, size_t limit,
                                                   uint32_t gates, int on_device, uint64_t *violations, zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !violations || (limit && !trace)) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    size_t bytes = zkc_carver::bytes(1, sizeof(DqDev)) + zkc_carver::bytes(1, 8);
    if (!on_device) bytes += zkc_carver::bytes((size_t)ZKC_DQ_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    DqDev *h = (DqDev *)ctx->pinned(sizeof(DqDev) + 8);
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    DqDev *d = cv.take<DqDev>(1);
    unsigned long long *dviol = cv.take<unsigned long long>(1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(DqDev));
    h->io = *io;
    h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(DqDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(dviol, 0, 8, s));
    const uint64_t *dt = trace;
    if (!on_device && limit) {
        uint64_t *b = cv.take<uint64_t>((size_t)ZKC_DQ_NUM_COLS * limit);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(b, trace, (size_t)ZKC_DQ_NUM_COLS * limit * 8, cudaMemcpyHostToDevice, s));
        dt = b;
    }
    ZKC_LAUNCH(ctx, "dq_prologue", dq_prologue_kernel, 1, 96, 0, d);
    if (limit) {
        const unsigned grid = (unsigned)((limit + 127) / 128);
        if (gates == 0 || (gates & ZKC_GATES_ROUND_FUNCTION)) ZKC_LAUNCH(ctx, "dq_check_rf", dq_check_kernel<true>, grid, 128, 0, d, dviol, dt);
        else ZKC_LAUNCH(ctx, "dq_check", dq_check_kernel<false>, grid, 128, 0, d, dviol, dt);
    }
    ZKC_CUDA(ctx, status, cudaGetLastError());
    unsigned long long *hviol = (unsigned long long *)(h + 1);
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(DqDev), cudaMemcpyDeviceToHost, s));
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
