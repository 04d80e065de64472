// Events / L2->L1 message sorter on sm_100a: sort_and_deduplicate_events_entry_point
// (/root/reference/src/log_sorter/mod.rs:34-232) and its loop
// repack_and_prove_events_rollbacks_inner (:234-441), one thread per loop iteration.
// Sequential state is recovered row-parallel exactly as in ram_permutation.cu; the additional
// piece is the RESULT queue, whose tail is a hash chain over the executed pushes: rounds 0 and 1
// of every push depend on the pushed item only and are computed per row, round 2 consumes the
// previous tail and is either verified against host-supplied tails (`result_tails`) or rebuilt
// by a sequential chain kernel (1 permutation per push).
#include "ctx.cuh"
#include "log_query.cuh"
#include "result_queue.cuh"
#include "scan.cuh"

namespace zkc {

struct EvDev {
    zkc_events_closed_form io;
    zkc_sorter_options opt;
    uint64_t n_unsorted, n_sorted, n_result_tails, limit;
    // prologue
    uint64_t ch[2][21];
    uint64_t acc0[4];  // rep*2 + side
    uint32_t start, prev_trivial0, previous_key0, prologue_checks;
    zkc_queue_state4 uq0, sq0, rq0;
    zkc_log_query previous_item0;
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
    unsigned long long violations;  // zkc_log_sorter_check_trace
};

__device__ int ev_encode_fsm(const zkc_events_fsm &f, uint64_t *dst) {
    int n = 0;
    dst[n++] = f.lhs_accumulator[0]; dst[n++] = f.lhs_accumulator[1];
    dst[n++] = f.rhs_accumulator[0]; dst[n++] = f.rhs_accumulator[1];
    n += put_queue_state4(dst + n, f.initial_unsorted_queue_state);
    n += put_queue_state4(dst + n, f.intermediate_sorted_queue_state);
    n += put_queue_state4(dst + n, f.final_result_queue_state);
    dst[n++] = f.previous_key;
    for (int i = 0; i < 36; i++) dst[n++] = lq_flat(f.previous_item, i);
    return n;  // 68
}

// query_to_add, log_sorter/mod.rs:381-393
__device__ __forceinline__ zkc_log_query ev_cleaned_up(const zkc_log_query &p) {
    zkc_log_query q = lq_zero();
#pragma unroll
    for (int i = 0; i < 5; i++) q.address[i] = p.address[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { q.key[i] = p.key[i]; q.written_value[i] = p.written_value[i]; }
    q.tx_number_in_block = p.tx_number_in_block;
    q.flags = ZKC_LQ_FLAGS(0, ZKC_LQ_SHARD(p.flags), 0, 0, ZKC_LQ_SERVICE(p.flags));
    return q;
}

// three warps, one 16-lane group each, every permutation spread over 12 lanes (poseidon2_permute_coop)
__global__ void ev_prologue_kernel(EvDev *d) {
    __shared__ uint64_t buf[3][80];
    const int warp = threadIdx.x >> 5, i = threadIdx.x & 31;
    if (i >= 16) return;
    const unsigned gm = 0xFFFFu;
    const zkc_events_closed_form &io = d->io;
    if (warp == 0) {
        if (i == 0) {
            const bool start = io.start_flag != 0;
            d->start = start;
            d->uq0 = start ? io.initial_log_queue_state : io.hidden_fsm_input.initial_unsorted_queue_state;
            d->sq0 = start ? io.intermediate_sorted_queue_state : io.hidden_fsm_input.intermediate_sorted_queue_state;
            zkc_queue_state4 empty;
            for (int i = 0; i < 4; i++) empty.head[i] = empty.tail[i] = 0;
            empty.length = 0; empty._pad = 0;
            d->rq0 = start ? empty : io.hidden_fsm_input.final_result_queue_state;
            for (int i = 0; i < 2; i++) {
                d->acc0[i * 2 + 0] = start ? 1 : io.hidden_fsm_input.lhs_accumulator[i];
                d->acc0[i * 2 + 1] = start ? 1 : io.hidden_fsm_input.rhs_accumulator[i];
            }
            d->previous_key0 = start ? 0 : io.hidden_fsm_input.previous_key;
            d->previous_item0 = start ? lq_zero() : io.hidden_fsm_input.previous_item;
            d->prev_trivial0 = (d->uq0.length == 0) || start;  // :266-267
            uint32_t checks = 0;
            for (int i = 0; i < 4; i++)
                if (io.initial_log_queue_state.head[i] | io.intermediate_sorted_queue_state.head[i]) checks |= ZKC_EV_CHK_TRIVIAL_HEAD;
            if (d->uq0.length != d->sq0.length) checks |= ZKC_EV_CHK_LENGTHS_EQUAL;
            d->prologue_checks = checks;
            // produce_fs_challenges over tail || len || tail || len (10 elements), 2 x 20 challenges
            uint64_t *in = buf[0];
            for (int k = 0; k < 4; k++) { in[k] = io.initial_log_queue_state.tail[k]; in[5 + k] = io.intermediate_sorted_queue_state.tail[k]; }
            in[4] = io.initial_log_queue_state.length; in[9] = io.intermediate_sorted_queue_state.length;
        }
        __syncwarp(gm);
        fs_challenges_coop(gm, buf[0], 10, 21, &d->ch[0][0], i);
    } else {
        uint64_t *b = buf[warp];
        int n = 0;
        if (i == 0) {
            if (warp == 1) {
                n = put_queue_state4(b, io.initial_log_queue_state);
                n += put_queue_state4(b + n, io.intermediate_sorted_queue_state);
            } else {
                n = ev_encode_fsm(io.hidden_fsm_input, b);
            }
        }
        __syncwarp(gm);
        n = __shfl_sync(gm, n, 0, 16);
        const uint64_t c = commit_encoding_coop(gm, b, n, i);
        if (i < 4) (warp == 1 ? d->commit_obs_in : d->commit_fsm_in)[i] = c;
    }
}

__device__ __forceinline__ void ev_report(EvDev *d, size_t row, uint32_t checks) {
    if (!checks) return;
    atomicOr(&d->failed_checks, checks);
    atomicMin(&d->first_bad, ((unsigned long long)row << 16) | checks);
}

// ---- pass A: pops, grand product, ordering / rollback logic, rounds 0-1 of the result push ------
__global__ void __launch_bounds__(SCAN_THREADS)
ev_rows_kernel(EvDev *d, const zkc_log_query *__restrict__ unsorted, const uint64_t *__restrict__ uprev,
               const zkc_log_query *__restrict__ sorted, const uint64_t *__restrict__ sprev,
               uint64_t *__restrict__ trace, uint64_t *__restrict__ r2in, uint32_t *__restrict__ meta,
               ScanGlobal *sg, TileState *tiles) {
    __shared__ ScanShared sh;
    __shared__ uint64_t ch[2][21];
    if (threadIdx.x < 42) ch[threadIdx.x / 21][threadIdx.x % 21] = d->ch[threadIdx.x / 21][threadIdx.x % 21];
    const unsigned int tile = scan_take_ticket(sg, sh);
    const size_t limit = d->limit;
    const size_t row = (size_t)tile * SCAN_THREADS + threadIdx.x;
    const bool in_range = row < limit;
    const uint32_t ulen0 = d->uq0.length, slen0 = d->sq0.length;
    const bool o_empty = row >= ulen0, s_empty = row >= slen0;
    const bool should_pop = in_range && !o_empty;
    const size_t active_rows = limit < ulen0 ? limit : ulen0;
    uint32_t checks = 0;
    if (in_range && o_empty != s_empty) checks |= ZKC_EV_CHK_EMPTY_SYNC;
#define TR(col) trace[(size_t)(col) * limit + row]
    const bool wr = in_range && trace != nullptr;
    zkc_log_query si = lq_zero();
    uint64_t contrib[4];
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        const zkc_log_query *recs = k ? sorted : unsorted;
        const uint64_t *prev = k ? sprev : uprev;
        const size_t n_rec = k ? d->n_sorted : d->n_unsorted;
        const zkc_queue_state4 &q0 = k ? d->sq0 : d->uq0;
        zkc_log_query it = lq_zero();
        if (should_pop && row < n_rec) it = lq_load(recs + row);
        uint64_t e[20];
        lq_encode(it, e);
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
            if (!hint_ok) { checks |= ZKC_EV_CHK_QUEUE_HINT; d->hint_bad = 1; }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) head[i] = ulen0 == 0 ? q0.head[i] : q0.tail[i];
        }
        if (should_pop && !ZKC_LQ_RW(it.flags)) checks |= k ? ZKC_EV_CHK_SORTED_IS_WRITE : ZKC_EV_CHK_UNSORTED_IS_WRITE;
        if (wr) {
            const int base = k ? ZKC_EV_SORTED_ITEM : ZKC_EV_UNSORTED_ITEM;
#pragma unroll
            for (int i = 0; i < 36; i++) TR(base + i) = lq_flat(it, i);
#pragma unroll
            for (int i = 0; i < 20; i++) TR(base + 36 + i) = e[i];
#pragma unroll
            for (int i = 0; i < 4; i++) TR(base + 56 + i) = head[i];
            const uint32_t len0 = k ? slen0 : ulen0;
            const size_t popped_now = row + 1 < active_rows ? row + 1 : active_rows;
            TR(base + 60) = len0 >= popped_now ? len0 - (uint32_t)popped_now : 0;
        }
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            uint64_t c = ch[rep][20];
#pragma unroll
            for (int i = 0; i < 20; i++) {
                c = gl_fma(e[i], ch[rep][i], c);
                if (wr) TR(ZKC_EV_GP_CHAIN + (rep * 2 + k) * 20 + i) = c;
            }
            contrib[rep * 2 + k] = c;
        }
        if (k == 1) si = it;
    }

    // ---- :315-400 ordering, rollback pairing, what to push ---------------------------------------------
    zkc_log_query pq;
    uint32_t previous_key;
    bool previous_is_trivial;
    if (row == 0) {
        pq = d->previous_item0;
        previous_key = d->previous_key0;
        previous_is_trivial = d->prev_trivial0;
    } else {
        pq = lq_zero();
        if (in_range && row - 1 < active_rows && row - 1 < d->n_sorted) pq = lq_load(sorted + row - 1);
        previous_key = pq.timestamp;
        previous_is_trivial = row - 1 >= ulen0;
    }
    const bool is_trivial = o_empty;
    const uint32_t sorting_key = si.timestamp;
    const uint64_t dd = (uint64_t)sorting_key - previous_key;  // b - a with a = previous, b = current
    const uint32_t diff = (uint32_t)dd;
    const bool new_key_is_smaller = (dd >> 32) & 1, keys_equal = diff == 0;
    if (should_pop && new_key_is_smaller) checks |= ZKC_EV_CHK_ORDER;
    const bool same_log = keys_equal;
    const bool same_nontrivial = should_pop && same_log;
    const bool maybe_different = !same_log;
    const bool different_nontrivial = should_pop && maybe_different;
    const bool rollback = ZKC_LQ_ROLLBACK(si.flags);
    if (different_nontrivial && rollback) checks |= ZKC_EV_CHK_NOT_ROLLBACK;
    if (same_nontrivial && !rollback) checks |= ZKC_EV_CHK_IS_ROLLBACK;
    bool item_keys_equal = true, values_equal = true;
#pragma unroll
    for (int i = 0; i < 8; i++) { item_keys_equal &= si.key[i] == pq.key[i]; values_equal &= si.written_value[i] == pq.written_value[i]; }
    const bool same_body = item_keys_equal && values_equal;
    const bool previous_non_trivial = !previous_is_trivial;
    const bool should_enforce = same_log && previous_non_trivial;
    if (in_range && should_enforce && !same_body) checks |= ZKC_EV_CHK_SAME_BODY;
    const bool maybe_add = maybe_different || is_trivial;
    const bool add = in_range && previous_non_trivial && maybe_add && !ZKC_LQ_ROLLBACK(pq.flags);

    // rounds 0 and 1 of result_queue.push(query_to_add)
    {
        const zkc_log_query to_add = ev_cleaned_up(pq);
        uint64_t pe[20], s[12];
        lq_encode(to_add, pe);
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = i < 8 ? pe[i] : 0;
        poseidon2_permute(s);
        if (wr) {
#pragma unroll
            for (int i = 0; i < 20; i++) TR(ZKC_EV_PUSH_ENC + i) = pe[i];
#pragma unroll
            for (int i = 0; i < 12; i++) TR(ZKC_EV_PUSH_ROUND0 + i) = s[i];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = pe[8 + i];
        poseidon2_permute(s);
        if (wr) {
#pragma unroll
            for (int i = 0; i < 12; i++) TR(ZKC_EV_PUSH_ROUND1 + i) = s[i];
        }
        if (in_range) {
            ulonglong2 *o = reinterpret_cast<ulonglong2 *>(r2in + 8 * row);
            o[0] = make_ulonglong2(pe[16], pe[17]); o[1] = make_ulonglong2(pe[18], pe[19]);
            o[2] = make_ulonglong2(s[8], s[9]); o[3] = make_ulonglong2(s[10], s[11]);
        }
    }

    ScanVal v = scan_identity();
    if (should_pop) {
#pragma unroll
        for (int i = 0; i < 4; i++) v.p[i] = contrib[i];
    }
    v.c = add;
    ScanVal init;
#pragma unroll
    for (int i = 0; i < 4; i++) init.p[i] = d->acc0[i];
    init.c = 0;
    ScanVal incl;
    const ScanVal excl = scan_tile(v, tile, init, tiles, sh, incl);
    if (in_range) meta[row] = (excl.c << 1) | (uint32_t)add;

    if (wr) {
        TR(ZKC_EV_ORIGINAL_IS_EMPTY) = o_empty; TR(ZKC_EV_SORTED_IS_EMPTY) = s_empty; TR(ZKC_EV_SHOULD_POP) = should_pop;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            TR(ZKC_EV_GP_NEW + i) = should_pop ? incl.p[i] : gl_mul(excl.p[i], contrib[i]);
            TR(ZKC_EV_GP_ACC + i) = incl.p[i];
        }
        TR(ZKC_EV_CMP_DIFF) = diff; TR(ZKC_EV_CMP_BORROW) = new_key_is_smaller; TR(ZKC_EV_KEYS_EQUAL) = keys_equal;
        TR(ZKC_EV_SAME_NONTRIVIAL_LOG) = same_nontrivial; TR(ZKC_EV_DIFFERENT_NONTRIVIAL_LOG) = different_nontrivial;
        TR(ZKC_EV_ITEM_KEYS_EQUAL) = item_keys_equal; TR(ZKC_EV_VALUES_EQUAL) = values_equal; TR(ZKC_EV_SAME_BODY) = same_body;
        TR(ZKC_EV_PREVIOUS_IS_TRIVIAL) = previous_is_trivial; TR(ZKC_EV_SHOULD_ENFORCE) = should_enforce;
        TR(ZKC_EV_MAYBE_ADD) = maybe_add; TR(ZKC_EV_ADD_TO_QUEUE) = add;
        TR(ZKC_EV_RESULT_LEN) = d->rq0.length + incl.c;
    }
    if (in_range && row == limit - 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) d->acc_final[i] = incl.p[i];
        d->pushes_in_loop = incl.c;
    }
    if (in_range) ev_report(d, row, checks);
#undef TR
}

// ---- finalize -----------------------------------------------------------------------------------------
__global__ void ev_finalize_kernel(EvDev *d, const zkc_log_query *__restrict__ sorted, const uint64_t *__restrict__ tails,
                                   size_t n_tails) {
    // lane 0 does the scalar bookkeeping; the commitments' permutations run on the two 16-lane groups, 12 lanes each
    __shared__ uint64_t e_out[80], o_out[16], compact[24];
    __shared__ uint32_t sh_completed, sh_n_out;
    const int lane = threadIdx.x & 31, li = lane & 15;
    const unsigned gm = lane < 16 ? 0xFFFFu : 0xFFFF0000u;
    if (lane == 0) {
    zkc_events_closed_form &io = d->io;
    const size_t limit = d->limit;
    const uint32_t len0 = d->uq0.length;
    const size_t popped = limit < len0 ? limit : len0;
    zkc_events_fsm out;
    memset(&out, 0, sizeof out);
    out.initial_unsorted_queue_state = d->uq0;
    out.intermediate_sorted_queue_state = d->sq0;
    if (popped > 0)
        for (int i = 0; i < 4; i++) {
            out.initial_unsorted_queue_state.head[i] = d->head_final[0][i];
            out.intermediate_sorted_queue_state.head[i] = d->head_final[1][i];
        }
    out.initial_unsorted_queue_state.length = len0 - (uint32_t)popped;
    const size_t spopped = d->sq0.length < popped ? d->sq0.length : popped;
    out.intermediate_sorted_queue_state.length = d->sq0.length - (uint32_t)spopped;
    zkc_log_query previous_item = d->previous_item0;
    uint32_t previous_key = d->previous_key0;
    bool previous_is_trivial = d->prev_trivial0;
    zkc_queue_state4 rq = d->rq0;
    bool hint_bad = d->hint_bad;
    if (limit > 0) {
        for (int i = 0; i < 2; i++) { out.lhs_accumulator[i] = d->acc_final[2 * i]; out.rhs_accumulator[i] = d->acc_final[2 * i + 1]; }
        previous_item = (limit - 1 < popped && limit - 1 < d->n_sorted) ? sorted[limit - 1] : lq_zero();
        previous_key = previous_item.timestamp;
        previous_is_trivial = limit - 1 >= len0;
        const uint32_t pushes = d->pushes_in_loop;
        if (pushes) {
            if (pushes - 1 < n_tails) for (int i = 0; i < 4; i++) rq.tail[i] = tails[4 * (size_t)(pushes - 1) + i];
            else hint_bad = true;
        }
        rq.length += pushes;
    } else {
        for (int i = 0; i < 2; i++) { out.lhs_accumulator[i] = d->acc0[2 * i]; out.rhs_accumulator[i] = d->acc0[2 * i + 1]; }
    }
    // finalisation push, :406-435
    {
        const bool now_empty = out.initial_unsorted_queue_state.length == 0;
        const bool add = !previous_is_trivial && !ZKC_LQ_ROLLBACK(previous_item.flags) && now_empty;
        if (add) {
            const zkc_log_query to_add = ev_cleaned_up(previous_item);
            uint64_t pe[20], s[12], chain[4];
            lq_encode(to_add, pe);
            for (int i = 0; i < 4; i++) chain[i] = rq.tail[i];
            lq_absorb_head(pe, s);
            lq_absorb_tail(pe, chain, s);
            for (int i = 0; i < 4; i++) rq.tail[i] = s[i];
            rq.length++;
        }
    }
    out.previous_key = previous_key;
    out.previous_item = previous_item;
    out.final_result_queue_state = rq;
    uint32_t checks = d->failed_checks | d->prologue_checks;
    const zkc_queue_state4 *qs[2] = {&out.initial_unsorted_queue_state, &out.intermediate_sorted_queue_state};
    for (int k = 0; k < 2; k++)
        if (qs[k]->length == 0)
            for (int i = 0; i < 4; i++)
                if (qs[k]->head[i] != qs[k]->tail[i]) checks |= ZKC_EV_CHK_QUEUE_CONSISTENCY;
    if ((qs[0]->length == 0) != (qs[1]->length == 0)) checks |= ZKC_EV_CHK_EMPTY_SYNC;
    const bool completed = qs[0]->length == 0;
    if (completed)
        for (int i = 0; i < 2; i++)
            if (out.lhs_accumulator[i] != out.rhs_accumulator[i]) checks |= ZKC_EV_CHK_GRAND_PRODUCT;
    zkc_queue_state4 obs_out;
    memset(&obs_out, 0, sizeof obs_out);
    if (completed) obs_out = rq;

    uint64_t e_exp[68], o_exp[9];
    const int n_out = ev_encode_fsm(out, e_out);
    put_queue_state4(o_out, obs_out);
    zkc_status st;
    st.code = ZKC_OK; st.cuda_error = 0; st.first_bad_row = -1; st.failed_checks = checks; st.reserved = 0;
    if (d->first_bad != ~0ull) st.first_bad_row = (int64_t)(d->first_bad >> 16);
    if (checks) st.code = ZKC_ERR_UNSATISFIED;
    if (hint_bad) { st.code = ZKC_ERR_QUEUE_WITNESS_INCONSISTENT; st.failed_checks |= ZKC_EV_CHK_QUEUE_HINT; }
    if (d->opt.compare_expected) {
        ev_encode_fsm(io.hidden_fsm_output, e_exp);
        put_queue_state4(o_exp, io.final_queue_state);
        bool same = (io.completion_flag != 0) == completed;
        for (int i = 0; i < n_out; i++) same &= e_out[i] == e_exp[i];
        for (int i = 0; i < 9; i++) same &= o_out[i] == o_exp[i];
        if (!same && st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    io.hidden_fsm_output = out;
    io.final_queue_state = obs_out;
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

// CircuitQueue::push of whole queues: one thread per independent queue
__global__ void log_queue_simulate_kernel(const zkc_log_query *__restrict__ recs, const uint32_t *__restrict__ extra_ts,
                                          size_t n_per_queue, size_t n_queues, uint64_t *__restrict__ prev_tails,
                                          zkc_queue_state4 *__restrict__ final_states) {
    const size_t qi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= n_queues) return;
    uint64_t tail[4] = {0, 0, 0, 0};
    for (size_t r = 0; r < n_per_queue; r++) {
        const size_t g = qi * n_per_queue + r;
        if (prev_tails) {
#pragma unroll
            for (int i = 0; i < 4; i++) prev_tails[4 * g + i] = tail[i];
        }
        const zkc_log_query it = lq_load(recs + g);
        uint64_t e[20], s[12];
        lq_encode(it, e);
        if (extra_ts) e[19] += (uint64_t)extra_ts[g] << 8;  // storage_validity_by_grand_product/mod.rs:72-96
        lq_absorb_head(e, s);
        lq_absorb_tail(e, tail, s);
#pragma unroll
        for (int i = 0; i < 4; i++) tail[i] = s[i];
    }
    zkc_queue_state4 &o = final_states[qi];
#pragma unroll
    for (int i = 0; i < 4; i++) { o.head[i] = 0; o.tail[i] = tail[i]; }
    o.length = (uint32_t)n_per_queue;
    o._pad = 0;
}

// ---- constraint evaluation of a finished trace ------------------------------------------------------------------------------
// One thread per row re-evaluates every relation the loop body of repack_and_prove_events_rollbacks_inner places that is local
// to a row or to a row and its predecessor (the role of `check_if_satisfied` over these cells, log_sorter/mod.rs:626-634):
// booleans / ranges of the allocated items, LogQuery::encode of both pops and of the pushed record, queue-length bookkeeping,
// the 4 x 20 Num::fma chains and the accumulator update, the timestamp borrow chain, the flag algebra of :327-372, the
// conditional enforcements, the result queue's length / tail selection.  Streams all ZKC_EV_NUM_COLS columns once (+ the
// previous row of the carried ones, an L1 / L2 hit); with ZKC_GATES_ROUND_FUNCTION also the three permutations of the push.
template <bool ROUND_FUNCTION>
__global__ void __launch_bounds__(128)
ev_check_kernel(EvDev *d, const uint64_t *__restrict__ trace) {
    __shared__ uint64_t ch[2][21];
    if (threadIdx.x < 42) ch[threadIdx.x / 21][threadIdx.x % 21] = d->ch[threadIdx.x / 21][threadIdx.x % 21];
    __syncthreads();
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const bool first = row == 0;
#define TR(col) __ldg(trace + (size_t)(col) * limit + row)
#define TP(col) __ldg(trace + (size_t)(col) * limit + row - 1)
    uint32_t bad = 0;
    const uint64_t o_empty = TR(ZKC_EV_ORIGINAL_IS_EMPTY), s_empty = TR(ZKC_EV_SORTED_IS_EMPTY), should_pop = TR(ZKC_EV_SHOULD_POP);
    if ((o_empty | s_empty | should_pop) > 1 || o_empty != s_empty || should_pop != 1 - o_empty) bad |= ZKC_EVV_BOOLEAN;
    zkc_log_query si = lq_zero(), pq = lq_zero();
    uint64_t enc[2][20];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int base = k ? ZKC_EV_SORTED_ITEM : ZKC_EV_UNSORTED_ITEM;
        const zkc_queue_state4 &q0 = k ? d->sq0 : d->uq0;
        uint64_t f[36], limbs = 0;
#pragma unroll
        for (int i = 0; i < 36; i++) f[i] = TR(base + i);
#pragma unroll
        for (int i = 0; i < 29; i++) limbs |= f[i];
        if ((limbs | f[34] | f[35]) >> 32 || (f[29] | f[33]) >> 8 || (f[30] | f[31] | f[32]) > 1) bad |= ZKC_EVV_BOOLEAN;
        zkc_log_query q = lq_zero();
#pragma unroll
        for (int i = 0; i < 5; i++) q.address[i] = (uint32_t)f[i];
#pragma unroll
        for (int i = 0; i < 8; i++) { q.key[i] = (uint32_t)f[5 + i]; q.read_value[i] = (uint32_t)f[13 + i]; q.written_value[i] = (uint32_t)f[21 + i]; }
        q.flags = ZKC_LQ_FLAGS((uint32_t)f[29], (uint32_t)f[33], (uint32_t)f[30], (uint32_t)f[31], (uint32_t)f[32]);
        q.tx_number_in_block = (uint32_t)f[34]; q.timestamp = (uint32_t)f[35];
        uint64_t e[20];
        lq_encode(q, e);
#pragma unroll
        for (int i = 0; i < 20; i++) { enc[k][i] = TR(base + 36 + i); if (enc[k][i] != e[i]) bad |= ZKC_EVV_ENCODING; }
        // queue: is_empty <=> previous length == 0, length decrements on a pop, the head only moves on a pop
        const uint64_t len_prev = first ? q0.length : TP(base + 60), len = TR(base + 60);
        if ((k ? s_empty : o_empty) != (uint64_t)(len_prev == 0) || len + should_pop != len_prev) bad |= ZKC_EVV_QUEUE_LEN;
        bool same = true;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint64_t h = TR(base + 56 + i);
            same &= h == (first ? q0.head[i] : TP(base + 56 + i));
            if (h >= GL_P) bad |= ZKC_EVV_BOOLEAN;
        }
        if (!should_pop && !same) bad |= ZKC_EVV_QUEUE_LEN;
        if (should_pop && !ZKC_LQ_RW(q.flags)) bad |= ZKC_EVV_ENFORCE;  // :295-297, :318-320
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
                const uint64_t cell = TR(ZKC_EV_GP_CHAIN + g * 20 + i);
                if (cell != gl_fma(enc[k][i], ch[rep][i], c)) bad |= ZKC_EVV_GP_CHAIN;
                c = cell;
            }
            const uint64_t acc_prev = first ? d->acc0[g] : TP(ZKC_EV_GP_ACC + g);
            const uint64_t nw = TR(ZKC_EV_GP_NEW + g), acc = TR(ZKC_EV_GP_ACC + g);
            if (nw != gl_mul(acc_prev, c) || acc != (should_pop ? nw : acc_prev)) bad |= ZKC_EVV_GP_ACC;
        }
    }
    // the previous item / key / triviality: the neighbouring row (row 0: the FSM input)
    uint64_t previous_key, prev_trivial;
    if (first) { pq = d->previous_item0; previous_key = d->previous_key0; prev_trivial = d->prev_trivial0; }
    else {
#pragma unroll
        for (int i = 0; i < 5; i++) pq.address[i] = (uint32_t)TP(ZKC_EV_SORTED_ITEM + i);
#pragma unroll
        for (int i = 0; i < 8; i++) { pq.key[i] = (uint32_t)TP(ZKC_EV_SORTED_ITEM + 5 + i); pq.written_value[i] = (uint32_t)TP(ZKC_EV_SORTED_ITEM + 21 + i); }
        pq.flags = ZKC_LQ_FLAGS(0, (uint32_t)TP(ZKC_EV_SORTED_ITEM + 33), 0, (uint32_t)TP(ZKC_EV_SORTED_ITEM + 31), (uint32_t)TP(ZKC_EV_SORTED_ITEM + 32));
        pq.tx_number_in_block = (uint32_t)TP(ZKC_EV_SORTED_ITEM + 34);
        previous_key = TP(ZKC_EV_SORTED_ITEM + 35);
        prev_trivial = TP(ZKC_EV_ORIGINAL_IS_EMPTY);
    }
    // :327 borrow chain: current - previous = diff - 2^32 * borrow
    const uint64_t diff = TR(ZKC_EV_CMP_DIFF), borrow = TR(ZKC_EV_CMP_BORROW), keys_equal = TR(ZKC_EV_KEYS_EQUAL);
    if ((diff >> 32) || borrow > 1 || keys_equal != (uint64_t)(diff == 0) || (uint64_t)si.timestamp + (borrow << 32) != diff + previous_key) bad |= ZKC_EVV_COMPARISON;
    // :335-372 flags
    const uint64_t same_nt = TR(ZKC_EV_SAME_NONTRIVIAL_LOG), diff_nt = TR(ZKC_EV_DIFFERENT_NONTRIVIAL_LOG), ike = TR(ZKC_EV_ITEM_KEYS_EQUAL),
                   ve = TR(ZKC_EV_VALUES_EQUAL), same_body = TR(ZKC_EV_SAME_BODY), pit = TR(ZKC_EV_PREVIOUS_IS_TRIVIAL),
                   should_enforce = TR(ZKC_EV_SHOULD_ENFORCE), maybe_add = TR(ZKC_EV_MAYBE_ADD), add = TR(ZKC_EV_ADD_TO_QUEUE);
    bool keq = true, veq = true;
#pragma unroll
    for (int i = 0; i < 8; i++) { keq &= si.key[i] == pq.key[i]; veq &= si.written_value[i] == pq.written_value[i]; }
    const uint64_t rollback = ZKC_LQ_ROLLBACK(si.flags);
    if ((same_nt | diff_nt | ike | ve | same_body | pit | should_enforce | maybe_add | add) > 1 || same_nt != (should_pop & keys_equal) ||
        diff_nt != (should_pop & (1 - keys_equal)) || ike != (uint64_t)keq || ve != (uint64_t)veq || same_body != (ike & ve) || pit != prev_trivial ||
        should_enforce != (keys_equal & (1 - pit)) || maybe_add != ((1 - keys_equal) | o_empty) ||
        add != ((1 - pit) & maybe_add & (1 - (uint64_t)ZKC_LQ_ROLLBACK(pq.flags))))
        bad |= ZKC_EVV_FLAGS;
    // conditional enforcements :331, :342-343, :347-349, :362
    if ((should_pop & borrow) | (diff_nt & rollback) | (same_nt & (1 - rollback)) | (should_enforce & (1 - same_body))) bad |= ZKC_EVV_ENFORCE;
    // :381-397 the pushed record and the result queue
    {
        const zkc_log_query to_add = ev_cleaned_up(pq);
        uint64_t pe[20], penc[20], s[12];
        lq_encode(to_add, pe);
#pragma unroll
        for (int i = 0; i < 20; i++) { penc[i] = TR(ZKC_EV_PUSH_ENC + i); if (penc[i] != pe[i]) bad |= ZKC_EVV_ENCODING; }
        uint64_t r0[12], r1[12], r2[12], tail_prev[4];
#pragma unroll
        for (int i = 0; i < 12; i++) { r0[i] = TR(ZKC_EV_PUSH_ROUND0 + i); r1[i] = TR(ZKC_EV_PUSH_ROUND1 + i); r2[i] = TR(ZKC_EV_PUSH_ROUND2 + i); }
#pragma unroll
        for (int i = 0; i < 4; i++) tail_prev[i] = first ? d->rq0.tail[i] : TP(ZKC_EV_RESULT_TAIL + i);
        const uint64_t len_prev = first ? d->rq0.length : TP(ZKC_EV_RESULT_LEN);
        if (TR(ZKC_EV_RESULT_LEN) != len_prev + add) bad |= ZKC_EVV_RESULT_QUEUE;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (TR(ZKC_EV_RESULT_TAIL + i) != (add ? r2[i] : tail_prev[i])) bad |= ZKC_EVV_RESULT_QUEUE;
        if (ROUND_FUNCTION) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = i < 8 ? penc[i] : 0;
            poseidon2_permute(s);
#pragma unroll
            for (int i = 0; i < 12; i++) if (s[i] != r0[i]) bad |= ZKC_EVV_ROUND_FUNCTION;
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = penc[8 + i];
            poseidon2_permute(s);
#pragma unroll
            for (int i = 0; i < 12; i++) if (s[i] != r1[i]) bad |= ZKC_EVV_ROUND_FUNCTION;
#pragma unroll
            for (int i = 0; i < 4; i++) { s[i] = penc[16 + i]; s[4 + i] = tail_prev[i]; }
            poseidon2_permute(s);
#pragma unroll
            for (int i = 0; i < 12; i++) if (s[i] != r2[i]) bad |= ZKC_EVV_ROUND_FUNCTION;
        } else {
            uint64_t big = 0;
#pragma unroll
            for (int i = 0; i < 12; i++) big |= (uint64_t)(r0[i] >= GL_P) | (uint64_t)(r1[i] >= GL_P) | (uint64_t)(r2[i] >= GL_P);
            if (big) bad |= ZKC_EVV_BOOLEAN;
        }
    }
#undef TR
#undef TP
    if (bad) {
        atomicAdd(&d->violations, 1ull);
        atomicOr(&d->failed_checks, bad);
        atomicMin(&d->first_bad, ((unsigned long long)row << 16) | bad);
    }
}

}  // namespace zkc

using namespace zkc;

extern "C" int zkc_log_queue_simulate(zkc_ctx *ctx, const zkc_log_query *records, const uint32_t *extra_timestamps,
                                      size_t n_per_queue, size_t n_queues, uint64_t *prev_tails,
                                      zkc_queue_state4 *final_states, int on_device) {
    if (!ctx || !final_states || (n_per_queue && n_queues && !records)) return ZKC_ERR_INVALID_ARGUMENT;
    if (!n_queues) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    const size_t n = n_per_queue * n_queues;
    const zkc_log_query *dr = records;
    const uint32_t *dt = extra_timestamps;
    uint64_t *dp = prev_tails;
    zkc_queue_state4 *df = final_states;
    cudaStream_t s = ctx->stream;
    if (!on_device) {
        size_t bytes = zkc_carver::bytes(n, sizeof(zkc_log_query)) + zkc_carver::bytes(n, 4) + zkc_carver::bytes(n * 4, 8) +
                       zkc_carver::bytes(n_queues, sizeof(zkc_queue_state4));
        void *blk = ctx->scratch(bytes);
        if (!blk) return ZKC_ERR_CUDA;
        zkc_carver cv(blk);
        zkc_log_query *br = cv.take<zkc_log_query>(n);
        uint32_t *bt = cv.take<uint32_t>(n);
        dp = prev_tails ? cv.take<uint64_t>(n * 4) : nullptr;
        df = cv.take<zkc_queue_state4>(n_queues);
        if (n) ZKC_CUDA(ctx, st, cudaMemcpyAsync(br, records, n * sizeof(zkc_log_query), cudaMemcpyHostToDevice, s));
        if (n && extra_timestamps) ZKC_CUDA(ctx, st, cudaMemcpyAsync(bt, extra_timestamps, n * 4, cudaMemcpyHostToDevice, s));
        dr = br; dt = extra_timestamps ? bt : nullptr;
    }
    ZKC_LAUNCH(ctx, "log_queue_simulate", log_queue_simulate_kernel, (unsigned)((n_queues + 31) / 32), 32, 0, dr, dt,
               n_per_queue, n_queues, dp, df);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    if (!on_device) {
        if (prev_tails && n) ZKC_CUDA(ctx, st, cudaMemcpyAsync(prev_tails, dp, n * 32, cudaMemcpyDeviceToHost, s));
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(final_states, df, n_queues * sizeof(zkc_queue_state4), cudaMemcpyDeviceToHost, s));
        ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    }
    return ZKC_OK;
}

extern "C" int zkc_log_sorter_entry_point(zkc_ctx *ctx, zkc_events_closed_form *io, const zkc_log_query *unsorted,
                                          const uint64_t *unsorted_prev_tails, size_t n_unsorted,
                                          const zkc_log_query *sorted, const uint64_t *sorted_prev_tails, size_t n_sorted,
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
    const zkc_queue_state4 &uq = io->start_flag ? io->initial_log_queue_state : io->hidden_fsm_input.initial_unsorted_queue_state;
    const size_t need = limit < uq.length ? limit : uq.length;
    if (n_unsorted < need || n_sorted < need || (need && (!unsorted_prev_tails || !sorted_prev_tails))) {
        status->code = ZKC_ERR_INVALID_ARGUMENT;
        return ZKC_ERR_INVALID_ARGUMENT;
    }
    const bool in_dev = on_device & ZKC_INPUTS_ON_DEVICE, trace_dev = on_device & ZKC_TRACE_ON_DEVICE;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const size_t tiles = (limit + SCAN_THREADS - 1) / SCAN_THREADS;
    const bool have_tails = result_tails != nullptr;
    if (!have_tails) n_result_tails = limit + 1;
    size_t bytes = zkc_carver::bytes(1, sizeof(EvDev)) + zkc_carver::bytes(1, sizeof(ScanGlobal)) +
                   zkc_carver::bytes(tiles + 1, sizeof(TileState)) + zkc_carver::bytes(limit * 8 + 8, 8) +
                   zkc_carver::bytes(limit + 1, 4);
    if (!in_dev) bytes += 2 * zkc_carver::bytes(need + 1, sizeof(zkc_log_query)) + 2 * zkc_carver::bytes(need * 4 + 4, 8);
    if (!in_dev || !have_tails) bytes += zkc_carver::bytes(n_result_tails * 4 + 4, 8);
    if (trace && !trace_dev) bytes += zkc_carver::bytes((size_t)ZKC_EV_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    EvDev *h = (EvDev *)ctx->pinned(sizeof(EvDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    EvDev *d = cv.take<EvDev>(1);
    ScanGlobal *sg = cv.take<ScanGlobal>(1);
    TileState *ts = cv.take<TileState>(tiles + 1);
    uint64_t *r2in = cv.take<uint64_t>(limit * 8 + 8);
    uint32_t *meta = cv.take<uint32_t>(limit + 1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(EvDev));
    h->io = *io;
    if (options) h->opt = *options;
    h->n_unsorted = n_unsorted; h->n_sorted = n_sorted; h->n_result_tails = n_result_tails; h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(EvDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(sg, 0, (char *)(ts + tiles + 1) - (char *)sg, s));
    const zkc_log_query *du = unsorted, *dsq = sorted;
    const uint64_t *dup = unsorted_prev_tails, *dsp = sorted_prev_tails, *dtails = result_tails;
    uint64_t *dtrace = trace;
    if (!in_dev) {
        zkc_log_query *bu = cv.take<zkc_log_query>(need + 1), *bs = cv.take<zkc_log_query>(need + 1);
        uint64_t *bup = cv.take<uint64_t>(need * 4 + 4), *bsp = cv.take<uint64_t>(need * 4 + 4);
        if (need) {
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bu, unsorted, need * sizeof(zkc_log_query), cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bs, sorted, need * sizeof(zkc_log_query), cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bup, unsorted_prev_tails, need * 32, cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bsp, sorted_prev_tails, need * 32, cudaMemcpyHostToDevice, s));
        }
        du = bu; dsq = bs; dup = bup; dsp = bsp;
    }
    if (!in_dev || !have_tails) {
        uint64_t *bt = cv.take<uint64_t>(n_result_tails * 4 + 4);
        if (have_tails && n_result_tails)
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bt, result_tails, n_result_tails * 32, cudaMemcpyHostToDevice, s));
        dtails = bt;
    }
    if (trace && !trace_dev) dtrace = cv.take<uint64_t>((size_t)ZKC_EV_NUM_COLS * limit);

    ZKC_LAUNCH(ctx, "ev_prologue", ev_prologue_kernel, 1, 96, 0, d);
    if (tiles) {
        ZKC_LAUNCH(ctx, "ev_rows", ev_rows_kernel, (unsigned)tiles, SCAN_THREADS, 0, d, du, dup, dsq, dsp, dtrace, r2in, meta, sg, ts);
        if (!have_tails) ZKC_LAUNCH(ctx, "ev_chain", rq_chain_kernel<EvDev>, 1, 32, 0, d, r2in, meta, (uint64_t *)dtails);
        ZKC_LAUNCH(ctx, "ev_push", (rq_push_kernel<EvDev, ZKC_EV_PUSH_ROUND2, ZKC_EV_RESULT_TAIL, ZKC_EV_CHK_QUEUE_HINT>), (unsigned)((limit + 255) / 256), 256, 0, d, r2in, meta, dtails, n_result_tails, dtrace);
    }
    ZKC_LAUNCH(ctx, "ev_finalize", ev_finalize_kernel, 1, 32, 0, d, dsq, dtails, n_result_tails);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(EvDev), cudaMemcpyDeviceToHost, s));
    if (!trace_dev && trace && limit)
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(trace, dtrace, (size_t)ZKC_EV_NUM_COLS * limit * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    io->hidden_fsm_output = h->io.hidden_fsm_output;
    io->final_queue_state = h->io.final_queue_state;
    io->completion_flag = h->io.completion_flag;
    memcpy(commitment, h->commitment, 32);
    *status = h->status;
    return status->code;
}

extern "C" int zkc_log_sorter_check_trace(zkc_ctx *ctx, const zkc_events_closed_form *io, const uint64_t *trace, size_t limit, uint32_t gates,
                                          int on_device, uint64_t *violations, zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !violations || (limit && !trace)) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    size_t bytes = zkc_carver::bytes(1, sizeof(EvDev));
    if (!on_device) bytes += zkc_carver::bytes((size_t)ZKC_EV_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    EvDev *h = (EvDev *)ctx->pinned(sizeof(EvDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    EvDev *d = cv.take<EvDev>(1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(EvDev));
    h->io = *io;
    h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(EvDev), cudaMemcpyHostToDevice, s));
    const uint64_t *dt = trace;
    if (!on_device && limit) {
        uint64_t *b = cv.take<uint64_t>((size_t)ZKC_EV_NUM_COLS * limit);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(b, trace, (size_t)ZKC_EV_NUM_COLS * limit * 8, cudaMemcpyHostToDevice, s));
        dt = b;
    }
    ZKC_LAUNCH(ctx, "ev_prologue", ev_prologue_kernel, 1, 96, 0, d);
    if (limit) {
        const unsigned grid = (unsigned)((limit + 127) / 128);
        if (gates == 0 || (gates & ZKC_GATES_ROUND_FUNCTION)) ZKC_LAUNCH(ctx, "ev_check_rf", ev_check_kernel<true>, grid, 128, 0, d, dt);
        else ZKC_LAUNCH(ctx, "ev_check", ev_check_kernel<false>, grid, 128, 0, d, dt);
    }
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(EvDev), cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    *violations = h->violations;
    status->failed_checks = h->failed_checks;
    if (h->violations) {
        status->code = ZKC_ERR_UNSATISFIED;
        status->first_bad_row = (int64_t)(h->first_bad >> 16);
    }
    return status->code;
}
