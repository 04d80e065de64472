// Storage access sorter / deduplicator on sm_100a:
// sort_and_deduplicate_storage_access_entry_point (/root/reference/src/storage_validity_by_grand_product/
// mod.rs:166-507) and its loop sort_and_deduplicate_storage_access_inner (:510-897).
//
// The reference threads a per-cell state machine (base value, current value, rollback depth,
// "explicit read at depth 0" flag) through the sorted queue.  Here it is recovered row-parallel:
//   pass 1  pops, permutation contributions, key / timestamp ordering; scan #1 carries the running
//           products, the rollback depth (segmented sum, reset on a new cell) and the row index of
//           the last row that SET the current value (new cell, write, rollback);
//   pass 2  scan #2 carries the "read at depth 0" flag (segmented or) and the index of the last row
//           that set the base value (new cell, read at depth 0) -- it needs pass 1's depth;
//   pass 3  every row rebuilds the cell state before and after itself by gathering the setter rows,
//           decides the push of the finished cell, hashes rounds 0-1 of it; scan #3 counts pushes;
//   result_queue.cuh  chains the result-queue tail (round 2);
//   finalize  flush of the last cell, entry-point enforcements, FSM output, commitment.
#include "ctx.cuh"
#include "log_query.cuh"
#include "result_queue.cuh"
#include "scan.cuh"

namespace zkc {

struct StDev {
    zkc_storage_closed_form io;
    zkc_sorter_options opt;
    uint64_t n_unsorted, n_sorted, n_result_tails, limit;
    // prologue
    uint64_t ch[2][21];
    uint64_t acc0[4];
    uint32_t start, prev_trivial0, cycle0, prologue_checks;
    uint32_t packed_key0[13];
    uint32_t shard, pad1, pad2;
    zkc_queue_state4 uq0, sq0, rq0;
    uint64_t commit_obs_in[4], commit_fsm_in[4];
    // rows
    uint64_t acc_final[4];
    uint32_t pushes_in_loop, pad0;
    uint64_t head_final[2][4];
    // status
    unsigned long long first_bad;
    uint32_t failed_checks, hint_bad;
    // finalize
    uint64_t commitment[4];
    zkc_status status;
};

// per-row record left by pass 1 for passes 2 / 3
struct alignas(16) StMeta1 {
    uint32_t depth;      // this_cell_current_depth at the end of the iteration
    int32_t cur_setter;  // last row <= this one that set this_cell_current_value, -1 = none in this call
    uint32_t bits;
    uint32_t ts;         // TimestampedStorageLogRecord.timestamp of the sorted item
};
enum : uint32_t { B_NEW_CELL = 1, B_READ_SAME = 2, B_WNR = 4, B_WRB = 8, B_RW = 16, B_TRIVIAL = 32, B_KEYS_EQ = 64, B_ROLLBACK = 128 };
struct StMeta2 {
    int32_t base_setter;
    uint32_t flag;
};

struct V1 {
    uint64_t p[4];
    uint32_t seg, depth;
    int32_t setter;
    uint32_t pad;
};
struct V1Op {
    static __device__ __forceinline__ V1 identity() { return V1{{1, 1, 1, 1}, 0, 0, -1, 0}; }
    static __device__ __forceinline__ V1 combine(const V1 &a, const V1 &b) {
        V1 r;
#pragma unroll
        for (int i = 0; i < 4; i++) r.p[i] = gl_mul(a.p[i], b.p[i]);
        r.seg = a.seg | b.seg;
        r.depth = b.seg ? b.depth : a.depth + b.depth;
        r.setter = a.setter > b.setter ? a.setter : b.setter;
        r.pad = 0;
        return r;
    }
};
struct V2 {
    uint32_t seg, flag;
    int32_t setter;
    uint32_t pad;
};
struct V2Op {
    static __device__ __forceinline__ V2 identity() { return V2{0, 0, -1, 0}; }
    static __device__ __forceinline__ V2 combine(const V2 &a, const V2 &b) {
        return V2{a.seg | b.seg, b.seg ? b.flag : (a.flag | b.flag), a.setter > b.setter ? a.setter : b.setter, 0};
    }
};
struct V3 {
    uint32_t c, pad;
};
struct V3Op {
    static __device__ __forceinline__ V3 identity() { return V3{0, 0}; }
    static __device__ __forceinline__ V3 combine(const V3 &a, const V3 &b) { return V3{a.c + b.c, 0}; }
};

__device__ int st_encode_fsm(const zkc_storage_fsm &f, uint64_t *dst) {
    int n = 0;
    dst[n++] = f.lhs_accumulator[0]; dst[n++] = f.lhs_accumulator[1];
    dst[n++] = f.rhs_accumulator[0]; dst[n++] = f.rhs_accumulator[1];
    n += put_queue_state4(dst + n, f.current_unsorted_queue_state);
    n += put_queue_state4(dst + n, f.current_intermediate_sorted_queue_state);
    n += put_queue_state4(dst + n, f.current_final_sorted_queue_state);
    dst[n++] = f.cycle_idx;
    for (int i = 0; i < 13; i++) dst[n++] = f.previous_packed_key[i];
    for (int i = 0; i < 8; i++) dst[n++] = f.previous_key[i];
    for (int i = 0; i < 5; i++) dst[n++] = f.previous_address[i];
    dst[n++] = f.previous_timestamp;
    dst[n++] = f.this_cell_has_explicit_read_and_rollback_depth_zero;
    for (int i = 0; i < 8; i++) dst[n++] = f.this_cell_base_value[i];
    for (int i = 0; i < 8; i++) dst[n++] = f.this_cell_current_value[i];
    dst[n++] = f.this_cell_current_depth;
    return n;  // 77
}

// three warps, one 16-lane group each, every permutation spread over 12 lanes (poseidon2_permute_coop)
__global__ void st_prologue_kernel(StDev *d) {
    __shared__ uint64_t buf[3][80];
    const int warp = threadIdx.x >> 5, i = threadIdx.x & 31;
    if (i >= 16) return;
    const unsigned gm = 0xFFFFu;
    const zkc_storage_closed_form &io = d->io;
    if (warp == 0) {
        if (i == 0) {
            const bool start = io.start_flag != 0;
            d->start = start;
            d->shard = io.shard_id_to_process & 0xFF;
            d->uq0 = start ? io.unsorted_log_queue_state : io.hidden_fsm_input.current_unsorted_queue_state;
            d->sq0 = start ? io.intermediate_sorted_queue_state : io.hidden_fsm_input.current_intermediate_sorted_queue_state;
            zkc_queue_state4 empty;
            for (int i = 0; i < 4; i++) empty.head[i] = empty.tail[i] = 0;
            empty.length = 0; empty._pad = 0;
            d->rq0 = start ? empty : io.hidden_fsm_input.current_final_sorted_queue_state;
            for (int i = 0; i < 2; i++) {
                d->acc0[i * 2 + 0] = start ? 1 : io.hidden_fsm_input.lhs_accumulator[i];
                d->acc0[i * 2 + 1] = start ? 1 : io.hidden_fsm_input.rhs_accumulator[i];
            }
            for (int i = 0; i < 13; i++) d->packed_key0[i] = start ? 0 : io.hidden_fsm_input.previous_packed_key[i];  // :382-387
            d->cycle0 = start ? 0 : io.hidden_fsm_input.cycle_idx;                                                       // :389-394
            d->prev_trivial0 = (d->uq0.length == 0) || start;                                                            // :574-575
            uint32_t checks = 0;
            for (int i = 0; i < 4; i++)
                if (io.unsorted_log_queue_state.head[i] | io.intermediate_sorted_queue_state.head[i]) checks |= ZKC_ST_CHK_TRIVIAL_HEAD;
            if (d->uq0.length != d->sq0.length) checks |= ZKC_ST_CHK_LENGTHS_EQUAL;
            d->prologue_checks = checks;
            // produce_fs_challenges over tail || len || tail || len (10 elements), 2 x 20 challenges
            uint64_t *in = buf[0];
            for (int k = 0; k < 4; k++) { in[k] = io.unsorted_log_queue_state.tail[k]; in[5 + k] = io.intermediate_sorted_queue_state.tail[k]; }
            in[4] = io.unsorted_log_queue_state.length; in[9] = io.intermediate_sorted_queue_state.length;
        }
        __syncwarp(gm);
        fs_challenges_coop(gm, buf[0], 10, 21, &d->ch[0][0], i);
    } else {
        uint64_t *b = buf[warp];
        int n = 0;
        if (i == 0) {
            if (warp == 1) {
                b[n++] = io.shard_id_to_process & 0xFF;
                n += put_queue_state4(b + n, io.unsorted_log_queue_state);
                n += put_queue_state4(b + n, io.intermediate_sorted_queue_state);
            } else {
                n = st_encode_fsm(io.hidden_fsm_input, b);
            }
        }
        __syncwarp(gm);
        n = __shfl_sync(gm, n, 0, 16);
        const uint64_t c = commit_encoding_coop(gm, b, n, i);
        if (i < 4) (warp == 1 ? d->commit_obs_in : d->commit_fsm_in)[i] = c;
    }
}

__device__ __forceinline__ void st_report(StDev *d, size_t row, uint32_t checks) {
    if (!checks) return;
    atomicOr(&d->failed_checks, checks);
    atomicMin(&d->first_bad, ((unsigned long long)row << 16) | checks);
}

// ---- pass 1 -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS)
st_rows_kernel(StDev *d, const zkc_log_query *__restrict__ unsorted, const uint64_t *__restrict__ uprev,
               const zkc_log_query *__restrict__ sorted, const uint32_t *__restrict__ sorted_ts,
               const uint64_t *__restrict__ sprev, uint64_t *__restrict__ trace, StMeta1 *__restrict__ meta1,
               ScanGlobal *sg, TileStateT<V1> *tiles) {
    __shared__ ScanSharedT<V1> sh;
    __shared__ uint64_t ch[2][21];
    if (threadIdx.x < 42) ch[threadIdx.x / 21][threadIdx.x % 21] = d->ch[threadIdx.x / 21][threadIdx.x % 21];
    const unsigned int tile = scan_take_ticket(sg, sh);
    const size_t limit = d->limit;
    const size_t row = (size_t)tile * SCAN_THREADS + threadIdx.x;
    const bool in_range = row < limit;
    const uint32_t ulen0 = d->uq0.length, slen0 = d->sq0.length;
    const bool o_empty = row >= ulen0, s_empty = row >= slen0;
    const bool should_pop = in_range && !o_empty && !s_empty;
    const size_t minlen = ulen0 < slen0 ? ulen0 : slen0;
    const size_t active_rows = limit < minlen ? limit : minlen;
    const uint32_t original_timestamp = d->cycle0 + (uint32_t)row;
    uint32_t checks = 0;
    if (in_range && o_empty != s_empty) checks |= ZKC_ST_CHK_EMPTY_SYNC;
#define TR(col) trace[(size_t)(col) * limit + row]
    const bool wr = in_range && trace != nullptr;
    zkc_log_query si = lq_zero();
    uint32_t ts = 0;
    uint64_t contrib[4];
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        const zkc_log_query *recs = k ? sorted : unsorted;
        const uint64_t *prev = k ? sprev : uprev;
        const size_t n_rec = k ? d->n_sorted : d->n_unsorted;
        const zkc_queue_state4 &q0 = k ? d->sq0 : d->uq0;
        zkc_log_query it = lq_zero();
        uint32_t its = 0;
        if (should_pop && row < n_rec) {
            it = lq_load(recs + row);
            if (k && sorted_ts) its = __ldg(sorted_ts + row);
        }
        uint64_t e[20];
        lq_encode(it, e);
        const uint64_t raw19 = e[19];
        if (k) e[19] += (uint64_t)its << 8;  // TimestampedStorageLogRecord::encode, :98-109
        uint64_t head[4];
        if (should_pop) {
            uint64_t s[12], chain[4];
            bool hint_ok = true;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                chain[i] = __ldg(prev + 4 * row + i);
                if (row == 0 && chain[i] != q0.head[i]) hint_ok = false;
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
                for (int i = 0; i < 4; i++) d->head_final[k][i] = head[i];
            }
            if (!hint_ok) { checks |= ZKC_ST_CHK_QUEUE_HINT; d->hint_bad = 1; }
        } else {
            const uint32_t len0 = k ? slen0 : ulen0;
#pragma unroll
            for (int i = 0; i < 4; i++) head[i] = (len0 == 0 || active_rows == 0) ? q0.head[i] : q0.tail[i];
        }
        if (wr) {
            const int base = k ? ZKC_ST_SORTED_ITEM : ZKC_ST_UNSORTED_ITEM;
#pragma unroll
            for (int i = 0; i < 36; i++) TR(base + i) = lq_flat(it, i);
            if (k) {
                TR(ZKC_ST_SORTED_ITEM + 36) = its;
#pragma unroll
                for (int i = 0; i < 20; i++) TR(ZKC_ST_SORTED_ENC + i) = e[i];
            } else {
#pragma unroll
                for (int i = 0; i < 19; i++) TR(ZKC_ST_UNSORTED_ENC + i) = e[i];
                TR(ZKC_ST_UNSORTED_ENC + 19) = raw19;
            }
            const int hb = k ? ZKC_ST_SORTED_HEAD : ZKC_ST_UNSORTED_HEAD;
#pragma unroll
            for (int i = 0; i < 4; i++) TR(hb + i) = head[i];
            const uint32_t len0 = k ? slen0 : ulen0;
            const size_t popped_now = row + 1 < active_rows ? row + 1 : active_rows;
            TR(hb + 4) = len0 >= popped_now ? len0 - (uint32_t)popped_now : 0;
        }
        if (!k) {
            e[19] += (uint64_t)original_timestamp << 8;  // append_timestamp_to_raw_query_encoding, :605-610
            if (wr) TR(ZKC_ST_UNSORTED_EXT19) = e[19];
        }
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            uint64_t c = ch[rep][20];
#pragma unroll
            for (int i = 0; i < 20; i++) {
                c = gl_fma(e[i], ch[rep][i], c);
                if (wr) TR(ZKC_ST_GP_CHAIN + (rep * 2 + k) * 20 + i) = c;
            }
            contrib[rep * 2 + k] = c;
        }
        if (k == 1) { si = it; ts = its; }
    }
    const bool shard_ok = ZKC_LQ_SHARD(si.flags) == d->shard;
    if (should_pop && !shard_ok) checks |= ZKC_ST_CHK_SHARD_ID;

    // ---- :630-648 ordering against the previous row -------------------------------------------------
    uint32_t prev_pk[13], prev_ts;
    if (row == 0) {
#pragma unroll
        for (int i = 0; i < 13; i++) prev_pk[i] = d->packed_key0[i];
        prev_ts = d->io.hidden_fsm_input.previous_timestamp;
    } else {
        zkc_log_query pq = lq_zero();
        prev_ts = 0;
        if (in_range && row - 1 < active_rows && row - 1 < d->n_sorted) {
            pq = lq_load(sorted + row - 1);
            if (sorted_ts) prev_ts = __ldg(sorted_ts + row - 1);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) prev_pk[i] = pq.key[i];
#pragma unroll
        for (int i = 0; i < 5; i++) prev_pk[8 + i] = pq.address[i];
    }
    uint32_t borrow = 0;
    bool keys_equal = true;
#pragma unroll
    for (int i = 0; i < 13; i++) {  // packed_key - previous_packed_key, least significant limb first
        const uint32_t cur = i < 8 ? si.key[i] : si.address[i - 8];
        const uint64_t dd = (uint64_t)cur - prev_pk[i] - borrow;
        const uint32_t diff = (uint32_t)dd;
        borrow = (uint32_t)(dd >> 32) & 1u;
        keys_equal &= diff == 0;
        if (wr) { TR(ZKC_ST_CMP_DIFF + i) = diff; TR(ZKC_ST_CMP_BORROW + i) = borrow; TR(ZKC_ST_CMP_LIMB_EQ + i) = diff == 0; }
    }
    const bool previous_key_is_greater = borrow;
    const bool item_is_trivial = o_empty, not_trivial = !item_is_trivial;
    if (in_range && not_trivial && previous_key_is_greater) checks |= ZKC_ST_CHK_KEY_ORDER;
    const uint64_t td = (uint64_t)prev_ts - ts;
    const bool previous_ts_is_less = (td >> 32) & 1;
    const bool must_enforce = keys_equal && not_trivial;
    if (in_range && must_enforce && !previous_ts_is_less) checks |= ZKC_ST_CHK_TIMESTAMP_ORDER;
    if (row == 0 && d->start && should_pop && keys_equal) checks |= ZKC_ST_CHK_FIRST_KEY_NONZERO;  // :657-661

    const bool rw = ZKC_LQ_RW(si.flags), rollback = ZKC_LQ_ROLLBACK(si.flags);
    const bool new_cell = in_range && not_trivial && !keys_equal;
    const bool nt_same = in_range && not_trivial && keys_equal;
    const bool read_same = nt_same && !rw, write_same = nt_same && rw;
    const bool wnr = write_same && !rollback, wrb = write_same && rollback;

    V1 v = V1Op::identity();
    if (should_pop) {
#pragma unroll
        for (int i = 0; i < 4; i++) v.p[i] = contrib[i];
    }
    v.seg = new_cell;
    v.depth = new_cell ? (rw ? 1u : 0u) : (wnr ? 1u : (wrb ? 0xFFFFFFFFu : 0u));
    v.setter = (new_cell || wnr || wrb) ? (int32_t)row : -1;
    V1 init = V1Op::identity();
#pragma unroll
    for (int i = 0; i < 4; i++) init.p[i] = d->acc0[i];
    init.depth = d->io.hidden_fsm_input.this_cell_current_depth;
    V1 incl;
    const V1 excl = scan_tile_generic<V1, V1Op>(v, tile, init, tiles, sh, incl);
    if (wrb && excl.depth == 0) checks |= ZKC_ST_CHK_DEPTH_UNDERFLOW;

    if (in_range) {
        alignas(16) StMeta1 m;
        m.depth = incl.depth;
        m.cur_setter = incl.setter;
        m.bits = (new_cell ? B_NEW_CELL : 0) | (read_same ? B_READ_SAME : 0) | (wnr ? B_WNR : 0) | (wrb ? B_WRB : 0) |
                 (rw ? B_RW : 0) | (item_is_trivial ? B_TRIVIAL : 0) | (keys_equal ? B_KEYS_EQ : 0) | (rollback ? B_ROLLBACK : 0);
        m.ts = ts;
        *reinterpret_cast<uint4 *>(meta1 + row) = *reinterpret_cast<uint4 *>(&m);
    }
    if (wr) {
        TR(ZKC_ST_ORIGINAL_IS_EMPTY) = o_empty; TR(ZKC_ST_SORTED_IS_EMPTY) = s_empty; TR(ZKC_ST_SHOULD_POP) = should_pop;
        TR(ZKC_ST_ORIGINAL_TIMESTAMP) = original_timestamp; TR(ZKC_ST_SHARD_ID_IS_VALID) = shard_ok;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            TR(ZKC_ST_GP_NEW + i) = should_pop ? incl.p[i] : gl_mul(excl.p[i], contrib[i]);
            TR(ZKC_ST_GP_ACC + i) = incl.p[i];
        }
        TR(ZKC_ST_KEYS_ARE_EQUAL) = keys_equal; TR(ZKC_ST_PREVIOUS_KEY_IS_GREATER) = previous_key_is_greater;
        TR(ZKC_ST_TS_DIFF) = (uint32_t)td; TR(ZKC_ST_PREVIOUS_TIMESTAMP_IS_LESS) = previous_ts_is_less;
        TR(ZKC_ST_MUST_ENFORCE) = must_enforce; TR(ZKC_ST_NEW_NON_TRIVIAL_CELL) = new_cell;
        TR(ZKC_ST_NON_TRIVIAL_AND_SAME_CELL) = nt_same; TR(ZKC_ST_READ_OF_SAME_CELL) = read_same;
        TR(ZKC_ST_WRITE_OF_SAME_CELL) = write_same; TR(ZKC_ST_WRITE_NO_ROLLBACK) = wnr; TR(ZKC_ST_WRITE_ROLLBACK) = wrb;
        TR(ZKC_ST_CELL_CURRENT_DEPTH) = incl.depth; TR(ZKC_ST_ROLLBACK_DEPTH_IS_ZERO) = incl.depth == 0;
        TR(ZKC_ST_READ_AT_DEPTH_ZERO_OF_SAME_CELL) = incl.depth == 0 && read_same;
        TR(ZKC_ST_CHECK_READ_CONSISTENCY) = read_same || wnr;
    }
    if (in_range && row == limit - 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) d->acc_final[i] = incl.p[i];
    }
    if (in_range) st_report(d, row, checks);
#undef TR
}

// ---- pass 2: base-value setter and the depth-0 read flag ---------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS)
st_cell_kernel(StDev *d, const StMeta1 *__restrict__ meta1, StMeta2 *__restrict__ meta2, ScanGlobal *sg,
               TileStateT<V2> *tiles) {
    __shared__ ScanSharedT<V2> sh;
    const unsigned int tile = scan_take_ticket(sg, sh);
    const size_t limit = d->limit;
    const size_t row = (size_t)tile * SCAN_THREADS + threadIdx.x;
    const bool in_range = row < limit;
    V2 v = V2Op::identity();
    if (in_range) {
        const StMeta1 m = meta1[row];
        const bool new_cell = m.bits & B_NEW_CELL;
        const bool r0 = (m.bits & B_READ_SAME) && m.depth == 0;
        v.seg = new_cell;
        v.flag = new_cell ? !(m.bits & B_RW) : r0;
        v.setter = (new_cell || r0) ? (int32_t)row : -1;
    }
    V2 init = V2Op::identity();
    init.flag = d->io.hidden_fsm_input.this_cell_has_explicit_read_and_rollback_depth_zero & 1;
    V2 incl;
    scan_tile_generic<V2, V2Op>(v, tile, init, tiles, sh, incl);
    if (in_range) meta2[row] = StMeta2{incl.setter, incl.flag};
}

// current value established by setter row r: new cell -> rw ? written : read; write -> written; rollback -> read
__device__ __forceinline__ void st_value_set_by(const zkc_log_query *sorted, const StMeta1 *meta1, int32_t r, uint32_t (&out)[8]) {
    const uint32_t bits = meta1[r].bits;
    const bool take_written = (bits & B_WNR) || ((bits & B_NEW_CELL) && (bits & B_RW));
    const uint32_t *w = take_written ? sorted[r].written_value : sorted[r].read_value;
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = __ldg(w + i);
}

struct StCell {
    uint32_t base[8], cur[8], depth, flag;
};
// cell state at the END of row r (r = -1: the FSM input)
__device__ __forceinline__ StCell st_state_after(const StDev *d, const zkc_log_query *sorted, const StMeta1 *meta1,
                                                 const StMeta2 *meta2, long long r) {
    StCell c;
    const zkc_storage_fsm &f = d->io.hidden_fsm_input;
    int32_t cs = -1, bs = -1;
    if (r >= 0) {
        const StMeta1 m1 = meta1[r];
        const StMeta2 m2 = meta2[r];
        cs = m1.cur_setter; bs = m2.base_setter;
        c.depth = m1.depth; c.flag = m2.flag;
    } else {
        c.depth = f.this_cell_current_depth;
        c.flag = f.this_cell_has_explicit_read_and_rollback_depth_zero & 1;
    }
    if (cs >= 0) st_value_set_by(sorted, meta1, cs, c.cur);
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) c.cur[i] = f.this_cell_current_value[i];
    }
    if (bs >= 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) c.base[i] = __ldg(&sorted[bs].read_value[i]);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) c.base[i] = f.this_cell_base_value[i];
    }
    return c;
}

struct StDecision {
    bool value_is_unchanged, depth_is_zero, unchanged_not_by_rollback, issue_protective_read, should_write, should_update;
};
__device__ __forceinline__ StDecision st_decide(const StCell &c) {
    StDecision r;
    r.value_is_unchanged = true;
#pragma unroll
    for (int i = 0; i < 8; i++) r.value_is_unchanged &= c.cur[i] == c.base[i];
    r.depth_is_zero = c.depth == 0;
    r.unchanged_not_by_rollback = r.value_is_unchanged && !r.depth_is_zero;
    r.issue_protective_read = c.flag || r.unchanged_not_by_rollback;
    r.should_write = !r.value_is_unchanged;
    r.should_update = r.issue_protective_read || r.should_write;
    return r;
}
// net query of a finished cell, :676-688
__device__ __forceinline__ zkc_log_query st_net_query(const uint32_t *address, const uint32_t *key, const StCell &c,
                                                      bool should_write, uint32_t shard) {
    zkc_log_query q = lq_zero();
#pragma unroll
    for (int i = 0; i < 5; i++) q.address[i] = address[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { q.key[i] = key[i]; q.read_value[i] = c.base[i]; q.written_value[i] = c.cur[i]; }
    q.flags = ZKC_LQ_FLAGS(0, shard, should_write, 0, 0);
    return q;
}

// ---- pass 3: cell state around every row, the push decision, rounds 0-1 of the push ---------------------
__global__ void __launch_bounds__(SCAN_THREADS)
st_push_rows_kernel(StDev *d, const zkc_log_query *__restrict__ sorted, const StMeta1 *__restrict__ meta1,
                    const StMeta2 *__restrict__ meta2, uint64_t *__restrict__ trace, uint64_t *__restrict__ r2in,
                    uint32_t *__restrict__ meta, ScanGlobal *sg, TileStateT<V3> *tiles) {
    __shared__ ScanSharedT<V3> sh;
    const unsigned int tile = scan_take_ticket(sg, sh);
    const size_t limit = d->limit;
    const size_t row = (size_t)tile * SCAN_THREADS + threadIdx.x;
    const bool in_range = row < limit;
    const size_t r = in_range ? row : 0;
    const uint32_t ulen0 = d->uq0.length, slen0 = d->sq0.length;
    const size_t minlen = ulen0 < slen0 ? ulen0 : slen0;
    const size_t active_rows = limit < minlen ? limit : minlen;
#define TR(col) trace[(size_t)(col) * limit + row]
    const bool wr = in_range && trace != nullptr;
    const StMeta1 m1 = meta1[r];
    const bool keys_equal = m1.bits & B_KEYS_EQ;
    // state left by the previous iteration
    const StCell before = st_state_after(d, sorted, meta1, meta2, (long long)r - 1);
    uint32_t prev_address[5], prev_key[8];
    bool previous_item_is_trivial;
    if (r == 0) {
        const zkc_storage_fsm &f = d->io.hidden_fsm_input;
#pragma unroll
        for (int i = 0; i < 5; i++) prev_address[i] = f.previous_address[i];
#pragma unroll
        for (int i = 0; i < 8; i++) prev_key[i] = f.previous_key[i];
        previous_item_is_trivial = d->prev_trivial0;
    } else {
        const bool popped = r - 1 < active_rows && r - 1 < d->n_sorted;
#pragma unroll
        for (int i = 0; i < 5; i++) prev_address[i] = popped ? __ldg(&sorted[r - 1].address[i]) : 0;
#pragma unroll
        for (int i = 0; i < 8; i++) prev_key[i] = popped ? __ldg(&sorted[r - 1].key[i]) : 0;
        previous_item_is_trivial = r - 1 >= ulen0;
    }
    const StDecision dec = st_decide(before);
    const bool should_push = in_range && !previous_item_is_trivial && !keys_equal && dec.should_update;
    {
        const zkc_log_query q = st_net_query(prev_address, prev_key, before, dec.should_write, d->shard);
        uint64_t pe[20], s[12];
        lq_encode(q, pe);
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = i < 8 ? pe[i] : 0;
        poseidon2_permute(s);
        if (wr) {
#pragma unroll
            for (int i = 0; i < 20; i++) TR(ZKC_ST_PUSH_ENC + i) = pe[i];
#pragma unroll
            for (int i = 0; i < 12; i++) TR(ZKC_ST_PUSH_ROUND0 + i) = s[i];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = pe[8 + i];
        poseidon2_permute(s);
        if (wr) {
#pragma unroll
            for (int i = 0; i < 12; i++) TR(ZKC_ST_PUSH_ROUND1 + i) = s[i];
        }
        if (in_range) {
            ulonglong2 *o = reinterpret_cast<ulonglong2 *>(r2in + 8 * row);
            o[0] = make_ulonglong2(pe[16], pe[17]); o[1] = make_ulonglong2(pe[18], pe[19]);
            o[2] = make_ulonglong2(s[8], s[9]); o[3] = make_ulonglong2(s[10], s[11]);
        }
    }
    // state at the end of this iteration + the read-consistency check of the same-cell branch
    const StCell after = st_state_after(d, sorted, meta1, meta2, (long long)r);
    uint32_t checks = 0;
    bool read_is_equal = true;
    {
        const bool popped = r < active_rows && r < d->n_sorted;
        const bool new_cell = m1.bits & B_NEW_CELL;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t rv = popped ? __ldg(&sorted[r].read_value[i]) : 0;
            const uint32_t wv = popped ? __ldg(&sorted[r].written_value[i]) : 0;
            const uint32_t cur1 = new_cell ? ((m1.bits & B_RW) ? wv : rv) : before.cur[i];  // after the new-cell select
            read_is_equal &= cur1 == rv;
        }
        const bool check_read = (m1.bits & B_READ_SAME) || (m1.bits & B_WNR);
        if (in_range && check_read && !read_is_equal) checks |= ZKC_ST_CHK_READ_CONSISTENCY;
    }
    V3 v{should_push ? 1u : 0u, 0};
    V3 incl;
    const V3 excl = scan_tile_generic<V3, V3Op>(v, tile, V3Op::identity(), tiles, sh, incl);
    if (in_range) meta[row] = (excl.c << 1) | (uint32_t)should_push;
    if (wr) {
        TR(ZKC_ST_VALUE_IS_UNCHANGED) = dec.value_is_unchanged; TR(ZKC_ST_CURRENT_DEPTH_IS_ZERO) = dec.depth_is_zero;
        TR(ZKC_ST_UNCHANGED_BUT_NOT_BY_ROLLBACK) = dec.unchanged_not_by_rollback;
        TR(ZKC_ST_ISSUE_PROTECTIVE_READ) = dec.issue_protective_read; TR(ZKC_ST_SHOULD_WRITE) = dec.should_write;
        TR(ZKC_ST_SHOULD_UPDATE) = dec.should_update; TR(ZKC_ST_SHOULD_PUSH) = should_push;
        TR(ZKC_ST_RESULT_LEN) = d->rq0.length + incl.c;
#pragma unroll
        for (int i = 0; i < 8; i++) { TR(ZKC_ST_CELL_BASE_VALUE + i) = after.base[i]; TR(ZKC_ST_CELL_CURRENT_VALUE + i) = after.cur[i]; }
        TR(ZKC_ST_CELL_HAS_READ_AT_DEPTH_ZERO) = after.flag;
        TR(ZKC_ST_READ_IS_EQUAL_TO_CURRENT) = read_is_equal;
    }
    if (in_range && row == limit - 1) d->pushes_in_loop = incl.c;
    if (in_range) st_report(d, row, checks);
#undef TR
}

// ---- finalize ---------------------------------------------------------------------------------------------
__global__ void st_finalize_kernel(StDev *d, const zkc_log_query *__restrict__ sorted, const StMeta1 *__restrict__ meta1,
                                   const StMeta2 *__restrict__ meta2, const uint64_t *__restrict__ tails, size_t n_tails) {
    // lane 0 does the scalar bookkeeping; the commitments' permutations run on the two 16-lane groups, 12 lanes each
    __shared__ uint64_t e_out[80], o_out[16], compact[24];
    __shared__ uint32_t sh_completed, sh_n_out;
    const int lane = threadIdx.x & 31, li = lane & 15;
    const unsigned gm = lane < 16 ? 0xFFFFu : 0xFFFF0000u;
    if (lane == 0) {
    zkc_storage_closed_form &io = d->io;
    const zkc_storage_fsm &fin = io.hidden_fsm_input;
    const size_t limit = d->limit;
    const uint32_t ulen0 = d->uq0.length, slen0 = d->sq0.length;
    const size_t minlen = ulen0 < slen0 ? ulen0 : slen0;
    const size_t popped = limit < minlen ? limit : minlen;
    zkc_storage_fsm out;
    memset(&out, 0, sizeof out);
    out.current_unsorted_queue_state = d->uq0;
    out.current_intermediate_sorted_queue_state = d->sq0;
    if (popped > 0)
        for (int i = 0; i < 4; i++) {
            out.current_unsorted_queue_state.head[i] = d->head_final[0][i];
            out.current_intermediate_sorted_queue_state.head[i] = d->head_final[1][i];
        }
    out.current_unsorted_queue_state.length = ulen0 - (uint32_t)popped;
    out.current_intermediate_sorted_queue_state.length = slen0 - (uint32_t)popped;
    out.cycle_idx = d->cycle0 + (uint32_t)limit;
    zkc_queue_state4 rq = d->rq0;
    bool hint_bad = d->hint_bad;
    bool previous_item_is_trivial = d->prev_trivial0;
    StCell c = st_state_after(d, sorted, meta1, meta2, (long long)limit - 1);
    if (limit > 0) {
        for (int i = 0; i < 2; i++) { out.lhs_accumulator[i] = d->acc_final[2 * i]; out.rhs_accumulator[i] = d->acc_final[2 * i + 1]; }
        if (limit - 1 < popped && limit - 1 < d->n_sorted) {
            const zkc_log_query q = sorted[limit - 1];
            for (int i = 0; i < 8; i++) { out.previous_key[i] = q.key[i]; out.previous_packed_key[i] = q.key[i]; }
            for (int i = 0; i < 5; i++) { out.previous_address[i] = q.address[i]; out.previous_packed_key[8 + i] = q.address[i]; }
            out.previous_timestamp = meta1[limit - 1].ts;
        }
        previous_item_is_trivial = limit - 1 >= ulen0;
        const uint32_t pushes = d->pushes_in_loop;
        if (pushes) {
            if (pushes - 1 < n_tails) for (int i = 0; i < 4; i++) rq.tail[i] = tails[4 * (size_t)(pushes - 1) + i];
            else hint_bad = true;
        }
        rq.length += pushes;
    } else {
        for (int i = 0; i < 2; i++) { out.lhs_accumulator[i] = d->acc0[2 * i]; out.rhs_accumulator[i] = d->acc0[2 * i + 1]; }
        for (int i = 0; i < 13; i++) out.previous_packed_key[i] = d->packed_key0[i];
        for (int i = 0; i < 8; i++) out.previous_key[i] = fin.previous_key[i];
        for (int i = 0; i < 5; i++) out.previous_address[i] = fin.previous_address[i];
        out.previous_timestamp = fin.previous_timestamp;
    }
    // finalisation, :836-880
    {
        const bool queues_exhausted = out.current_unsorted_queue_state.length == 0;
        const StDecision dec = st_decide(c);
        const bool should_push = !previous_item_is_trivial && dec.should_update && queues_exhausted;
        if (should_push) {
            const zkc_log_query q = st_net_query(out.previous_address, out.previous_key, c, dec.should_write, d->shard);
            uint64_t pe[20], s[12], chain[4];
            lq_encode(q, pe);
            for (int i = 0; i < 4; i++) chain[i] = rq.tail[i];
            lq_absorb_head(pe, s);
            lq_absorb_tail(pe, chain, s);
            for (int i = 0; i < 4; i++) rq.tail[i] = s[i];
            rq.length++;
        }
        if (queues_exhausted) c.flag = 0;
    }
    out.this_cell_has_explicit_read_and_rollback_depth_zero = c.flag;
    for (int i = 0; i < 8; i++) { out.this_cell_base_value[i] = c.base[i]; out.this_cell_current_value[i] = c.cur[i]; }
    out.this_cell_current_depth = c.depth;
    out.current_final_sorted_queue_state = rq;
    uint32_t checks = d->failed_checks | d->prologue_checks;
    const zkc_queue_state4 *qs[2] = {&out.current_unsorted_queue_state, &out.current_intermediate_sorted_queue_state};
    for (int k = 0; k < 2; k++)
        if (qs[k]->length == 0)
            for (int i = 0; i < 4; i++)
                if (qs[k]->head[i] != qs[k]->tail[i]) checks |= ZKC_ST_CHK_QUEUE_CONSISTENCY;
    if ((qs[0]->length == 0) != (qs[1]->length == 0)) checks |= ZKC_ST_CHK_EMPTY_SYNC;
    const bool completed = qs[0]->length == 0 && qs[1]->length == 0;
    if (completed)
        for (int i = 0; i < 2; i++)
            if (out.lhs_accumulator[i] != out.rhs_accumulator[i]) checks |= ZKC_ST_CHK_GRAND_PRODUCT;
    zkc_queue_state4 obs_out;
    memset(&obs_out, 0, sizeof obs_out);
    if (completed) obs_out = rq;
    uint64_t e_exp[77], o_exp[9];
    const int n_out = st_encode_fsm(out, e_out);
    put_queue_state4(o_out, obs_out);
    zkc_status st;
    st.code = ZKC_OK; st.cuda_error = 0; st.first_bad_row = -1; st.failed_checks = checks; st.reserved = 0;
    if (d->first_bad != ~0ull) st.first_bad_row = (int64_t)(d->first_bad >> 16);
    if (checks) st.code = ZKC_ERR_UNSATISFIED;
    if (hint_bad) { st.code = ZKC_ERR_QUEUE_WITNESS_INCONSISTENT; st.failed_checks |= ZKC_ST_CHK_QUEUE_HINT; }
    if (d->opt.compare_expected) {
        st_encode_fsm(io.hidden_fsm_output, e_exp);
        put_queue_state4(o_exp, io.final_sorted_queue_state);
        bool same = (io.completion_flag != 0) == completed;
        for (int i = 0; i < n_out; i++) same &= e_out[i] == e_exp[i];
        for (int i = 0; i < 9; i++) same &= o_out[i] == o_exp[i];
        if (!same && st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    io.hidden_fsm_output = out;
    io.final_sorted_queue_state = obs_out;
    io.completion_flag = completed;
    compact[0] = d->start; compact[1] = completed;
    for (int i = 0; i < 4; i++) {
        compact[2 + i] = d->commit_obs_in[i];
        compact[10 + i] = d->start ? 0 : d->commit_fsm_in[i];
    }
    d->status = st;
    sh_completed = completed; sh_n_out = n_out;
    }
    __syncwarp();
    const bool completed = sh_completed;
    const uint64_t c = commit_encoding_coop(gm, lane < 16 ? e_out : o_out, lane < 16 ? (int)sh_n_out : 9, li);
    if (lane < 4) compact[14 + lane] = completed ? 0 : c;
    if (lane >= 16 && lane < 20) compact[6 + lane - 16] = completed ? c : 0;
    __syncwarp();
    if (lane < 16) {
        const uint64_t f = commit_encoding_coop(gm, compact, 18, li);
        if (li < 4) d->commitment[li] = f;
    }
}


// ---- constraint evaluation of a finished trace ------------------------------------------------------------------------------
// One thread per row re-evaluates every relation the loop body of sort_and_deduplicate_storage_access_inner places that is local
// to a row or to a row and its predecessor (mod.rs:560-800; the role of `check_if_satisfied` over these cells): booleans / ranges
// of the allocated items, LogQuery::encode of both pops (the timestamped forms of :98-109 and :605-610) and of the pushed net
// query, queue-length / head bookkeeping, the 4 x 20 Num::fma chains and the accumulator update, the 13-limb key comparison and
// the timestamp comparison, the flag algebra, the per-cell state machine (base / current value, rollback depth, the explicit-read
// flag) from the previous row's state, the push decision, the result queue's length / tail selection, the conditional
// enforcements.  Streams all ZKC_ST_NUM_COLS columns once (+ the previous row of the carried ones); with
// ZKC_GATES_ROUND_FUNCTION also the permutations: 3 per popped item and queue, 3 of the push.
template <bool ROUND_FUNCTION>
__global__ void __launch_bounds__(128)
st_check_kernel(StDev *d, unsigned long long *violations, const uint64_t *__restrict__ trace) {
    __shared__ uint64_t ch[2][21];
    if (threadIdx.x < 42) ch[threadIdx.x / 21][threadIdx.x % 21] = d->ch[threadIdx.x / 21][threadIdx.x % 21];
    __syncthreads();
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const bool first = row == 0;
    const zkc_storage_fsm &fsm = d->io.hidden_fsm_input;
#define TR(col) __ldg(trace + (size_t)(col) * limit + row)
#define TP(col) __ldg(trace + (size_t)(col) * limit + row - 1)
    uint32_t bad = 0;
    const uint64_t o_empty = TR(ZKC_ST_ORIGINAL_IS_EMPTY), s_empty = TR(ZKC_ST_SORTED_IS_EMPTY), should_pop = TR(ZKC_ST_SHOULD_POP);
    if ((o_empty | s_empty | should_pop) > 1 || o_empty != s_empty || should_pop != 1 - o_empty) bad |= ZKC_STV_BOOLEAN;
    const uint64_t original_ts = TR(ZKC_ST_ORIGINAL_TIMESTAMP);
    if (original_ts != (first ? (uint64_t)d->cycle0 : TP(ZKC_ST_ORIGINAL_TIMESTAMP) + 1) || original_ts >> 32) bad |= ZKC_STV_BOOLEAN;  // cycle_idx, :585-589
    zkc_log_query si = lq_zero();
    uint64_t sorted_ts = 0;
    uint64_t enc[2][20];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int base = k ? ZKC_ST_SORTED_ITEM : ZKC_ST_UNSORTED_ITEM, enc_base = k ? ZKC_ST_SORTED_ENC : ZKC_ST_UNSORTED_ENC;
        const int head_base = k ? ZKC_ST_SORTED_HEAD : ZKC_ST_UNSORTED_HEAD;
        const zkc_queue_state4 &q0 = k ? d->sq0 : d->uq0;
        uint64_t f[36], limbs = 0;
#pragma unroll
        for (int i = 0; i < 36; i++) f[i] = TR(base + i);
#pragma unroll
        for (int i = 0; i < 29; i++) limbs |= f[i];
        if ((limbs | f[34] | f[35]) >> 32 || (f[29] | f[33]) >> 8 || (f[30] | f[31] | f[32]) > 1) bad |= ZKC_STV_BOOLEAN;
        zkc_log_query q = lq_zero();
#pragma unroll
        for (int i = 0; i < 5; i++) q.address[i] = (uint32_t)f[i];
#pragma unroll
        for (int i = 0; i < 8; i++) { q.key[i] = (uint32_t)f[5 + i]; q.read_value[i] = (uint32_t)f[13 + i]; q.written_value[i] = (uint32_t)f[21 + i]; }
        q.flags = ZKC_LQ_FLAGS((uint32_t)f[29], (uint32_t)f[33], (uint32_t)f[30], (uint32_t)f[31], (uint32_t)f[32]);
        q.tx_number_in_block = (uint32_t)f[34]; q.timestamp = (uint32_t)f[35];
        uint64_t e[20];
        lq_encode(q, e);
        if (k) {  // TimestampedStorageLogRecord::encode, :98-109
            sorted_ts = TR(ZKC_ST_SORTED_ITEM + 36);
            if (sorted_ts >> 32) bad |= ZKC_STV_BOOLEAN;
            e[19] += sorted_ts << 8;
        }
#pragma unroll
        for (int i = 0; i < 20; i++) { enc[k][i] = TR(enc_base + i); if (enc[k][i] != e[i]) bad |= ZKC_STV_ENCODING; }
        // queue: is_empty <=> previous length == 0, length decrements on a pop, the head only moves on a pop
        const uint64_t len_prev = first ? q0.length : TP(head_base + 4), len = TR(head_base + 4);
        if ((k ? s_empty : o_empty) != (uint64_t)(len_prev == 0) || len + should_pop != len_prev) bad |= ZKC_STV_QUEUE_LEN;
        bool same = true;
        uint64_t head_prev[4], head[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            head[i] = TR(head_base + i);
            head_prev[i] = first ? q0.head[i] : TP(head_base + i);
            same &= head[i] == head_prev[i];
            if (head[i] >= GL_P) bad |= ZKC_STV_BOOLEAN;
        }
        if (!should_pop && !same) bad |= ZKC_STV_QUEUE_LEN;
        if (ROUND_FUNCTION && should_pop) {  // the popped item's absorption from the head before the pop
            uint64_t st[12];
            lq_absorb_head(enc[k], st);
            lq_absorb_tail(enc[k], head_prev, st);
#pragma unroll
            for (int i = 0; i < 4; i++) if (st[i] != head[i]) bad |= ZKC_STV_ROUND_FUNCTION;
        }
        if (!k) {  // append_timestamp_to_raw_query_encoding, :605-610: the lhs of the grand product carries the pop index
            const uint64_t ext = TR(ZKC_ST_UNSORTED_EXT19);
            if (ext != enc[0][19] + (original_ts << 8)) bad |= ZKC_STV_ENCODING;
            enc[0][19] = ext;
        }
        if (k == 1) si = q;
    }
    // utils.rs:104-135
#pragma unroll
    for (int rep = 0; rep < 2; rep++) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int g = rep * 2 + k;
            uint64_t c = ch[rep][20];
#pragma unroll
            for (int i = 0; i < 20; i++) {
                const uint64_t cell = TR(ZKC_ST_GP_CHAIN + g * 20 + i);
                if (cell != gl_fma(enc[k][i], ch[rep][i], c)) bad |= ZKC_STV_GP_CHAIN;
                c = cell;
            }
            const uint64_t acc_prev = first ? d->acc0[g] : TP(ZKC_ST_GP_ACC + g);
            const uint64_t nw = TR(ZKC_ST_GP_NEW + g), acc = TR(ZKC_ST_GP_ACC + g);
            if (nw != gl_mul(acc_prev, c) || acc != (should_pop ? nw : acc_prev)) bad |= ZKC_STV_GP_ACC;
        }
    }
    // ---- :612-661 shard id, ordering against the previous row (row 0: the FSM input) ---------------------------------------------
    const uint64_t shard_ok = TR(ZKC_ST_SHARD_ID_IS_VALID);
    if (shard_ok != (uint64_t)(ZKC_LQ_SHARD(si.flags) == d->shard)) bad |= ZKC_STV_FLAGS;
    uint32_t prev_pk[13], prev_address[5], prev_key[8];
    uint64_t prev_ts, prev_trivial;
    if (first) {
#pragma unroll
        for (int i = 0; i < 13; i++) prev_pk[i] = d->packed_key0[i];
#pragma unroll
        for (int i = 0; i < 5; i++) prev_address[i] = fsm.previous_address[i];
#pragma unroll
        for (int i = 0; i < 8; i++) prev_key[i] = fsm.previous_key[i];
        prev_ts = fsm.previous_timestamp;
        prev_trivial = d->prev_trivial0;
    } else {
#pragma unroll
        for (int i = 0; i < 5; i++) { prev_address[i] = (uint32_t)TP(ZKC_ST_SORTED_ITEM + i); prev_pk[8 + i] = prev_address[i]; }
#pragma unroll
        for (int i = 0; i < 8; i++) { prev_key[i] = (uint32_t)TP(ZKC_ST_SORTED_ITEM + 5 + i); prev_pk[i] = prev_key[i]; }
        prev_ts = TP(ZKC_ST_SORTED_ITEM + 36);
        prev_trivial = TP(ZKC_ST_ORIGINAL_IS_EMPTY);
    }
    uint64_t borrow = 0, all_eq = 1;
#pragma unroll
    for (int i = 0; i < 13; i++) {  // packed_key - previous_packed_key, least significant limb first: cur + 2^32 * borrow_out = diff + prev + borrow_in
        const uint64_t cur = i < 8 ? si.key[i] : si.address[i - 8];
        const uint64_t diff = TR(ZKC_ST_CMP_DIFF + i), bo = TR(ZKC_ST_CMP_BORROW + i), leq = TR(ZKC_ST_CMP_LIMB_EQ + i);
        if ((diff >> 32) || bo > 1 || leq != (uint64_t)(diff == 0) || cur + (bo << 32) != diff + prev_pk[i] + borrow) bad |= ZKC_STV_COMPARISON;
        borrow = bo;
        all_eq &= leq;
    }
    const uint64_t keys_equal = TR(ZKC_ST_KEYS_ARE_EQUAL), prev_greater = TR(ZKC_ST_PREVIOUS_KEY_IS_GREATER);
    const uint64_t ts_diff = TR(ZKC_ST_TS_DIFF), prev_ts_less = TR(ZKC_ST_PREVIOUS_TIMESTAMP_IS_LESS);
    if (keys_equal != all_eq || prev_greater != borrow || (ts_diff >> 32) || prev_ts_less > 1 || prev_ts + (prev_ts_less << 32) != ts_diff + sorted_ts)
        bad |= ZKC_STV_COMPARISON;
    // ---- flags ---------------------------------------------------------------------------------------------------------------------
    const uint64_t trivial = o_empty, not_trivial = 1 - (o_empty & 1);
    const uint64_t rw = ZKC_LQ_RW(si.flags), rollback = ZKC_LQ_ROLLBACK(si.flags);
    const uint64_t must_enforce = TR(ZKC_ST_MUST_ENFORCE), new_cell = TR(ZKC_ST_NEW_NON_TRIVIAL_CELL), nt_same = TR(ZKC_ST_NON_TRIVIAL_AND_SAME_CELL);
    const uint64_t read_same = TR(ZKC_ST_READ_OF_SAME_CELL), write_same = TR(ZKC_ST_WRITE_OF_SAME_CELL), wnr = TR(ZKC_ST_WRITE_NO_ROLLBACK), wrb = TR(ZKC_ST_WRITE_ROLLBACK);
    if (must_enforce != (keys_equal & not_trivial) || new_cell != (not_trivial & (1 - (keys_equal & 1))) || nt_same != (not_trivial & keys_equal) ||
        read_same != (nt_same & (1 - rw)) || write_same != (nt_same & rw) || wnr != (write_same & (1 - rollback)) || wrb != (write_same & rollback))
        bad |= ZKC_STV_FLAGS;
    (void)trivial;
    // ---- the cell state machine: state left by the previous row -> state after this one ------------------------------------------
    uint32_t b_base[8], b_cur[8], a_base[8], a_cur[8];
    uint64_t b_depth, b_flag;
    {
        uint64_t range = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint64_t ab = TR(ZKC_ST_CELL_BASE_VALUE + i), ac = TR(ZKC_ST_CELL_CURRENT_VALUE + i);
            range |= ab | ac;
            a_base[i] = (uint32_t)ab; a_cur[i] = (uint32_t)ac;
            b_base[i] = first ? fsm.this_cell_base_value[i] : (uint32_t)TP(ZKC_ST_CELL_BASE_VALUE + i);
            b_cur[i] = first ? fsm.this_cell_current_value[i] : (uint32_t)TP(ZKC_ST_CELL_CURRENT_VALUE + i);
        }
        if (range >> 32) bad |= ZKC_STV_BOOLEAN;
        b_depth = first ? (uint64_t)fsm.this_cell_current_depth : TP(ZKC_ST_CELL_CURRENT_DEPTH);
        b_flag = first ? (uint64_t)(fsm.this_cell_has_explicit_read_and_rollback_depth_zero & 1) : TP(ZKC_ST_CELL_HAS_READ_AT_DEPTH_ZERO);
    }
    // the finished cell's net query: decided on the state BEFORE this row's update (:663-705)
    const uint64_t viu = TR(ZKC_ST_VALUE_IS_UNCHANGED), dz = TR(ZKC_ST_CURRENT_DEPTH_IS_ZERO), ubnr = TR(ZKC_ST_UNCHANGED_BUT_NOT_BY_ROLLBACK);
    const uint64_t ipr = TR(ZKC_ST_ISSUE_PROTECTIVE_READ), should_write = TR(ZKC_ST_SHOULD_WRITE), should_update = TR(ZKC_ST_SHOULD_UPDATE);
    const uint64_t should_push = TR(ZKC_ST_SHOULD_PUSH);
    {
        bool unchanged = true;
#pragma unroll
        for (int i = 0; i < 8; i++) unchanged &= b_cur[i] == b_base[i];
        if (viu != (uint64_t)unchanged || dz != (uint64_t)(b_depth == 0) || ubnr != (viu & (1 - (dz & 1))) || ipr != (b_flag | ubnr) || should_write != 1 - (viu & 1) ||
            should_update != (ipr | should_write) || should_push != ((1 - (prev_trivial & 1)) & (1 - (keys_equal & 1)) & should_update) || prev_trivial > 1 || b_flag > 1)
            bad |= ZKC_STV_FLAGS;
    }
    // this row's update of the state (:707-800)
    const uint64_t a_depth = TR(ZKC_ST_CELL_CURRENT_DEPTH), a_flag = TR(ZKC_ST_CELL_HAS_READ_AT_DEPTH_ZERO);
    const uint64_t depth_zero_after = TR(ZKC_ST_ROLLBACK_DEPTH_IS_ZERO), r0 = TR(ZKC_ST_READ_AT_DEPTH_ZERO_OF_SAME_CELL);
    const uint64_t read_is_equal = TR(ZKC_ST_READ_IS_EQUAL_TO_CURRENT), check_read = TR(ZKC_ST_CHECK_READ_CONSISTENCY);
    {
        const uint64_t want_depth = new_cell ? rw : (uint64_t)(uint32_t)((uint32_t)b_depth + (uint32_t)wnr - (uint32_t)wrb);
        bool ok = a_depth == want_depth && (a_depth >> 32) == 0 && depth_zero_after == (uint64_t)(a_depth == 0) && r0 == (depth_zero_after & read_same) &&
                  a_flag == (new_cell ? 1 - rw : (b_flag | r0)) && check_read == (read_same | wnr);
        bool req = true;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t rv = si.read_value[i], wv = si.written_value[i];
            const uint32_t cur1 = new_cell ? (rw ? wv : rv) : b_cur[i];  // after the new-cell select
            req &= cur1 == rv;
            ok &= a_cur[i] == (new_cell ? cur1 : (wnr ? wv : (wrb ? rv : b_cur[i])));
            ok &= a_base[i] == ((new_cell | r0) ? rv : b_base[i]);
        }
        ok &= read_is_equal == (uint64_t)req;
        if (!ok) bad |= ZKC_STV_CELL_STATE;
    }
    // conditional enforcements: :612-614 shard, :638-639 key order, :645-648 timestamp order, :657-661 first key, :787-792 read
    // consistency; a rollback below depth 0 (decrement_unchecked, :774)
    if ((should_pop & (1 - (shard_ok & 1))) | (not_trivial & prev_greater) | (must_enforce & (1 - (prev_ts_less & 1))) | (check_read & (1 - (read_is_equal & 1))) |
        (uint64_t)(first && d->start && should_pop && keys_equal) | (wrb & (uint64_t)(b_depth == 0)))
        bad |= ZKC_STV_ENFORCE;
    // ---- the pushed net query and the result queue (:676-705) ----------------------------------------------------------------------
    {
        zkc_log_query q = lq_zero();
#pragma unroll
        for (int i = 0; i < 5; i++) q.address[i] = prev_address[i];
#pragma unroll
        for (int i = 0; i < 8; i++) { q.key[i] = prev_key[i]; q.read_value[i] = b_base[i]; q.written_value[i] = b_cur[i]; }
        q.flags = ZKC_LQ_FLAGS(0, d->shard, (uint32_t)(should_write & 1), 0, 0);
        uint64_t pe[20], penc[20], s[12];
        lq_encode(q, pe);
#pragma unroll
        for (int i = 0; i < 20; i++) { penc[i] = TR(ZKC_ST_PUSH_ENC + i); if (penc[i] != pe[i]) bad |= ZKC_STV_ENCODING; }
        uint64_t r0s[12], r1s[12], r2s[12], tail_prev[4];
#pragma unroll
        for (int i = 0; i < 12; i++) { r0s[i] = TR(ZKC_ST_PUSH_ROUND0 + i); r1s[i] = TR(ZKC_ST_PUSH_ROUND1 + i); r2s[i] = TR(ZKC_ST_PUSH_ROUND2 + i); }
#pragma unroll
        for (int i = 0; i < 4; i++) tail_prev[i] = first ? d->rq0.tail[i] : TP(ZKC_ST_RESULT_TAIL + i);
        const uint64_t len_prev = first ? d->rq0.length : TP(ZKC_ST_RESULT_LEN);
        if (TR(ZKC_ST_RESULT_LEN) != len_prev + should_push) bad |= ZKC_STV_RESULT_QUEUE;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (TR(ZKC_ST_RESULT_TAIL + i) != (should_push ? r2s[i] : tail_prev[i])) bad |= ZKC_STV_RESULT_QUEUE;
        if (ROUND_FUNCTION) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = i < 8 ? penc[i] : 0;
            poseidon2_permute(s);
#pragma unroll
            for (int i = 0; i < 12; i++) if (s[i] != r0s[i]) bad |= ZKC_STV_ROUND_FUNCTION;
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = penc[8 + i];
            poseidon2_permute(s);
#pragma unroll
            for (int i = 0; i < 12; i++) if (s[i] != r1s[i]) bad |= ZKC_STV_ROUND_FUNCTION;
#pragma unroll
            for (int i = 0; i < 4; i++) { s[i] = penc[16 + i]; s[4 + i] = tail_prev[i]; }
            poseidon2_permute(s);
#pragma unroll
            for (int i = 0; i < 12; i++) if (s[i] != r2s[i]) bad |= ZKC_STV_ROUND_FUNCTION;
        } else {
            uint64_t big = 0;
#pragma unroll
            for (int i = 0; i < 12; i++) big |= (uint64_t)(r0s[i] >= GL_P) | (uint64_t)(r1s[i] >= GL_P) | (uint64_t)(r2s[i] >= GL_P);
            if (big) bad |= ZKC_STV_BOOLEAN;
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

extern "C" int zkc_storage_validity_entry_point(zkc_ctx *ctx, zkc_storage_closed_form *io, const zkc_log_query *unsorted,
                                                const uint64_t *unsorted_prev_tails, size_t n_unsorted,
                                                const zkc_log_query *sorted, const uint32_t *sorted_timestamps,
                                                const uint64_t *sorted_prev_tails, size_t n_sorted,
                                                const uint64_t *result_tails, size_t n_result_tails, size_t limit,
                                                const zkc_sorter_options *options, int on_device, uint64_t *trace,
                                                uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !commitment || (n_unsorted && !unsorted) || (n_sorted && !sorted) || limit > 0x7FFFFFFFull) {
        status->code = ZKC_ERR_INVALID_ARGUMENT;
        return ZKC_ERR_INVALID_ARGUMENT;
    }
    const zkc_queue_state4 &uq = io->start_flag ? io->unsorted_log_queue_state : io->hidden_fsm_input.current_unsorted_queue_state;
    const zkc_queue_state4 &sq = io->start_flag ? io->intermediate_sorted_queue_state
                                                : io->hidden_fsm_input.current_intermediate_sorted_queue_state;
    const size_t minlen = uq.length < sq.length ? uq.length : sq.length;
    const size_t need = limit < minlen ? limit : minlen;
    if (n_unsorted < need || n_sorted < need || (need && (!unsorted_prev_tails || !sorted_prev_tails || !sorted_timestamps))) {
        status->code = ZKC_ERR_INVALID_ARGUMENT;
        return ZKC_ERR_INVALID_ARGUMENT;
    }
    const bool in_dev = on_device & ZKC_INPUTS_ON_DEVICE, trace_dev = on_device & ZKC_TRACE_ON_DEVICE;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const size_t tiles = (limit + SCAN_THREADS - 1) / SCAN_THREADS;
    const bool have_tails = result_tails != nullptr;
    if (!have_tails) n_result_tails = limit + 1;
    size_t bytes = zkc_carver::bytes(1, sizeof(StDev)) + 3 * zkc_carver::bytes(1, sizeof(ScanGlobal)) +
                   zkc_carver::bytes(tiles + 1, sizeof(TileStateT<V1>)) + zkc_carver::bytes(tiles + 1, sizeof(TileStateT<V2>)) +
                   zkc_carver::bytes(tiles + 1, sizeof(TileStateT<V3>)) + zkc_carver::bytes(limit * 8 + 8, 8) +
                   zkc_carver::bytes(limit + 1, 4) + zkc_carver::bytes(limit + 1, sizeof(StMeta1)) +
                   zkc_carver::bytes(limit + 1, sizeof(StMeta2));
    if (!in_dev) bytes += 2 * zkc_carver::bytes(need + 1, sizeof(zkc_log_query)) + 2 * zkc_carver::bytes(need * 4 + 4, 8) +
                          zkc_carver::bytes(need + 1, 4);
    if (!in_dev || !have_tails) bytes += zkc_carver::bytes(n_result_tails * 4 + 4, 8);
    if (trace && !trace_dev) bytes += zkc_carver::bytes((size_t)ZKC_ST_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    StDev *h = (StDev *)ctx->pinned(sizeof(StDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    StDev *d = cv.take<StDev>(1);
    // scan bookkeeping, zeroed in one memset
    char *zero_begin = cv.base + cv.off;
    ScanGlobal *sg1 = cv.take<ScanGlobal>(1), *sg2 = cv.take<ScanGlobal>(1), *sg3 = cv.take<ScanGlobal>(1);
    TileStateT<V1> *ts1 = cv.take<TileStateT<V1>>(tiles + 1);
    TileStateT<V2> *ts2 = cv.take<TileStateT<V2>>(tiles + 1);
    TileStateT<V3> *ts3 = cv.take<TileStateT<V3>>(tiles + 1);
    char *zero_end = cv.base + cv.off;
    uint64_t *r2in = cv.take<uint64_t>(limit * 8 + 8);
    uint32_t *meta = cv.take<uint32_t>(limit + 1);
    StMeta1 *meta1 = cv.take<StMeta1>(limit + 1);
    StMeta2 *meta2 = cv.take<StMeta2>(limit + 1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(StDev));
    h->io = *io;
    if (options) h->opt = *options;
    h->n_unsorted = n_unsorted; h->n_sorted = n_sorted; h->n_result_tails = n_result_tails; h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(StDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(zero_begin, 0, zero_end - zero_begin, s));
    const zkc_log_query *du = unsorted, *dsq = sorted;
    const uint32_t *dts = sorted_timestamps;
    const uint64_t *dup = unsorted_prev_tails, *dsp = sorted_prev_tails, *dtails = result_tails;
    uint64_t *dtrace = trace;
    if (!in_dev) {
        zkc_log_query *bu = cv.take<zkc_log_query>(need + 1), *bs = cv.take<zkc_log_query>(need + 1);
        uint64_t *bup = cv.take<uint64_t>(need * 4 + 4), *bsp = cv.take<uint64_t>(need * 4 + 4);
        uint32_t *bts = cv.take<uint32_t>(need + 1);
        if (need) {
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bu, unsorted, need * sizeof(zkc_log_query), cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bs, sorted, need * sizeof(zkc_log_query), cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bup, unsorted_prev_tails, need * 32, cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bsp, sorted_prev_tails, need * 32, cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bts, sorted_timestamps, need * 4, cudaMemcpyHostToDevice, s));
        }
        du = bu; dsq = bs; dup = bup; dsp = bsp; dts = bts;
    }
    if (!in_dev || !have_tails) {
        uint64_t *bt = cv.take<uint64_t>(n_result_tails * 4 + 4);
        if (have_tails && n_result_tails)
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bt, result_tails, n_result_tails * 32, cudaMemcpyHostToDevice, s));
        dtails = bt;
    }
    if (trace && !trace_dev) dtrace = cv.take<uint64_t>((size_t)ZKC_ST_NUM_COLS * limit);

    ZKC_LAUNCH(ctx, "st_prologue", st_prologue_kernel, 1, 96, 0, d);
    if (tiles) {
        ZKC_LAUNCH(ctx, "st_rows", st_rows_kernel, (unsigned)tiles, SCAN_THREADS, 0, d, du, dup, dsq, dts, dsp, dtrace, meta1, sg1, ts1);
        ZKC_LAUNCH(ctx, "st_cell", st_cell_kernel, (unsigned)tiles, SCAN_THREADS, 0, d, meta1, meta2, sg2, ts2);
        ZKC_LAUNCH(ctx, "st_push_rows", st_push_rows_kernel, (unsigned)tiles, SCAN_THREADS, 0, d, dsq, meta1, meta2, dtrace, r2in, meta, sg3, ts3);
        if (!have_tails) ZKC_LAUNCH(ctx, "st_chain", rq_chain_kernel<StDev>, 1, 32, 0, d, r2in, meta, (uint64_t *)dtails);
        ZKC_LAUNCH(ctx, "st_push", (rq_push_kernel<StDev, ZKC_ST_PUSH_ROUND2, ZKC_ST_RESULT_TAIL, ZKC_ST_CHK_QUEUE_HINT>),
                   (unsigned)((limit + 255) / 256), 256, 0, d, r2in, meta, dtails, n_result_tails, dtrace);
    }
    ZKC_LAUNCH(ctx, "st_finalize", st_finalize_kernel, 1, 32, 0, d, dsq, meta1, meta2, dtails, n_result_tails);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(StDev), cudaMemcpyDeviceToHost, s));
    if (!trace_dev && trace && limit)
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(trace, dtrace, (size_t)ZKC_ST_NUM_COLS * limit * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    io->hidden_fsm_output = h->io.hidden_fsm_output;
    io->final_sorted_queue_state = h->io.final_sorted_queue_state;
    io->completion_flag = h->io.completion_flag;
    memcpy(commitment, h->commitment, 32);
    *status = h->status;
    return status->code;
}

extern "C" int zkc_storage_validity_check_trace(zkc_ctx *ctx, const zkc_storage_closed_form *io, const uint64_t *trace, size_t limit, uint32_t gates,
                                                int on_device, uint64_t *violations, zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !violations || (limit && !trace)) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    size_t bytes = zkc_carver::bytes(1, sizeof(StDev)) + zkc_carver::bytes(1, 8);
    if (!on_device) bytes += zkc_carver::bytes((size_t)ZKC_ST_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    StDev *h = (StDev *)ctx->pinned(sizeof(StDev) + 8);
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    StDev *d = cv.take<StDev>(1);
    unsigned long long *dviol = cv.take<unsigned long long>(1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(StDev));
    h->io = *io;
    h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(StDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(dviol, 0, 8, s));
    const uint64_t *dt = trace;
    if (!on_device && limit) {
        uint64_t *b = cv.take<uint64_t>((size_t)ZKC_ST_NUM_COLS * limit);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(b, trace, (size_t)ZKC_ST_NUM_COLS * limit * 8, cudaMemcpyHostToDevice, s));
        dt = b;
    }
    ZKC_LAUNCH(ctx, "st_prologue", st_prologue_kernel, 1, 96, 0, d);
    if (limit) {
        const unsigned grid = (unsigned)((limit + 127) / 128);
        if (gates == 0 || (gates & ZKC_GATES_ROUND_FUNCTION)) ZKC_LAUNCH(ctx, "st_check_rf", st_check_kernel<true>, grid, 128, 0, d, dviol, dt);
        else ZKC_LAUNCH(ctx, "st_check", st_check_kernel<false>, grid, 128, 0, d, dviol, dt);
    }
    ZKC_CUDA(ctx, status, cudaGetLastError());
    unsigned long long *hviol = (unsigned long long *)(h + 1);
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(StDev), cudaMemcpyDeviceToHost, s));
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
