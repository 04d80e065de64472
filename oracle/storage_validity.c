/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Sequential CPU restatement of the storage access sorter / deduplicator:
 *   sort_and_deduplicate_storage_access_entry_point  /root/reference/src/storage_validity_by_grand_product/mod.rs:166-507
 *   sort_and_deduplicate_storage_access_inner        /root/reference/src/storage_validity_by_grand_product/mod.rs:510-897
 *   concatenate_key / unpacked_long_comparison       /root/reference/src/storage_validity_by_grand_product/mod.rs:899-944
 * Pinning: loop logic pinned by the reference's vectors (test_input.rs, test mod.rs:1035: every in-loop
 * enforcement holds; the vector is NOT a permutation, which only the entry point would notice);
 * hash-dependent values PARITY UNPINNED (Poseidon2).
 */
#include "oracle.h"
#include <string.h>

static void fail(zkc_status *st, int64_t row, uint32_t bit) {
    st->code = ZKC_ERR_UNSATISFIED;
    st->failed_checks |= bit;
    if (row >= 0 && (st->first_bad_row < 0 || row < st->first_bad_row)) st->first_bad_row = row;
}

size_t orc_storage_encode_fsm(const zkc_storage_fsm *f, uint64_t *dst) {
    size_t n = 0;
    dst[n++] = f->lhs_accumulator[0]; dst[n++] = f->lhs_accumulator[1];
    dst[n++] = f->rhs_accumulator[0]; dst[n++] = f->rhs_accumulator[1];
    n += orc_put_queue_state4(dst + n, &f->current_unsorted_queue_state);
    n += orc_put_queue_state4(dst + n, &f->current_intermediate_sorted_queue_state);
    n += orc_put_queue_state4(dst + n, &f->current_final_sorted_queue_state);
    dst[n++] = f->cycle_idx;
    for (int i = 0; i < 13; i++) dst[n++] = f->previous_packed_key[i];
    for (int i = 0; i < 8; i++) dst[n++] = f->previous_key[i];
    for (int i = 0; i < 5; i++) dst[n++] = f->previous_address[i];
    dst[n++] = f->previous_timestamp;
    dst[n++] = f->this_cell_has_explicit_read_and_rollback_depth_zero;
    for (int i = 0; i < 8; i++) dst[n++] = f->this_cell_base_value[i];
    for (int i = 0; i < 8; i++) dst[n++] = f->this_cell_current_value[i];
    dst[n++] = f->this_cell_current_depth;
    return n; /* 77 */
}

typedef struct {
    uint32_t packed_key[13], key[8], address[5], timestamp;
    uint32_t flag, base[8], cur[8], depth;
    int item_is_trivial;
} cell_state;

/* the net query of the finished cell, :676-688 / :849-861 */
static zkc_log_query net_query(const cell_state *c, int should_write, uint32_t shard) {
    zkc_log_query q;
    memset(&q, 0, sizeof q);
    memcpy(q.address, c->address, sizeof q.address);
    memcpy(q.key, c->key, sizeof q.key);
    memcpy(q.read_value, c->base, sizeof q.read_value);
    memcpy(q.written_value, c->cur, sizeof q.written_value);
    q.flags = ZKC_LQ_FLAGS(0, shard, should_write, 0, 0);
    return q;
}

#define T(col, r) trace[(size_t)(col) * limit + (r)]

int orc_storage_validity_entry_point(zkc_storage_closed_form *io, const zkc_log_query *unsorted, size_t n_unsorted,
                                     const zkc_log_query *sorted, const uint32_t *sorted_ts, size_t n_sorted, size_t limit,
                                     const zkc_sorter_options *options, uint64_t *trace, uint64_t *result_tails,
                                     size_t *n_result_tails, uint64_t commitment[4], zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    const int start = io->start_flag != 0;
    const zkc_storage_fsm *fin = &io->hidden_fsm_input;
    static const uint64_t zero4[4] = {0, 0, 0, 0};
    if (memcmp(io->unsorted_log_queue_state.head, zero4, 32) || memcmp(io->intermediate_sorted_queue_state.head, zero4, 32))
        fail(&st, -1, ZKC_ST_CHK_TRIVIAL_HEAD);
    zkc_queue_state4 uq = start ? io->unsorted_log_queue_state : fin->current_unsorted_queue_state;
    zkc_queue_state4 sq = start ? io->intermediate_sorted_queue_state : fin->current_intermediate_sorted_queue_state;
    zkc_queue_state4 rq;
    memset(&rq, 0, sizeof rq);
    if (!start) rq = fin->current_final_sorted_queue_state;
    uint64_t ch[2][21];
    orc_produce_fs_challenges(io->unsorted_log_queue_state.tail, io->unsorted_log_queue_state.length,
                              io->intermediate_sorted_queue_state.tail, io->intermediate_sorted_queue_state.length, 4, 21,
                              &ch[0][0]);
    uint64_t lhs[2], rhs[2];
    for (int i = 0; i < 2; i++) {
        lhs[i] = start ? 1 : fin->lhs_accumulator[i];
        rhs[i] = start ? 1 : fin->rhs_accumulator[i];
    }
    cell_state c;
    memset(&c, 0, sizeof c);
    if (!start) memcpy(c.packed_key, fin->previous_packed_key, sizeof c.packed_key); /* :382-387 */
    uint32_t cycle_idx = start ? 0 : fin->cycle_idx;                                  /* :389-394 */
    /* the remaining FSM fields are NOT masked by start_flag, :419-427 */
    memcpy(c.key, fin->previous_key, sizeof c.key);
    memcpy(c.address, fin->previous_address, sizeof c.address);
    c.timestamp = fin->previous_timestamp;
    c.flag = fin->this_cell_has_explicit_read_and_rollback_depth_zero & 1;
    memcpy(c.base, fin->this_cell_base_value, sizeof c.base);
    memcpy(c.cur, fin->this_cell_current_value, sizeof c.cur);
    c.depth = fin->this_cell_current_depth;
    const uint32_t shard = io->shard_id_to_process & 0xFF;

    if (uq.length != sq.length) fail(&st, -1, ZKC_ST_CHK_LENGTHS_EQUAL); /* :565-569 */
    const int no_work = uq.length == 0;
    int previous_item_is_trivial = no_work || start; /* :574-575 */

    size_t upos = 0, spos = 0, pushes = 0;
    for (size_t cyc = 0; cyc < limit; cyc++) {
        const uint32_t original_timestamp = cycle_idx;
        cycle_idx = cycle_idx + 1; /* :585-590 */
        const int o_empty = uq.length == 0, s_empty = sq.length == 0;
        if (o_empty != s_empty) fail(&st, (int64_t)cyc, ZKC_ST_CHK_EMPTY_SYNC);
        const int should_pop = !o_empty && !s_empty;
        const int item_is_trivial = o_empty;
        zkc_log_query ui, si;
        uint32_t ts = 0;
        memset(&ui, 0, sizeof ui); memset(&si, 0, sizeof si);
        if (should_pop) {
            if (upos < n_unsorted) ui = unsorted[upos++];
            if (spos < n_sorted) { si = sorted[spos]; ts = sorted_ts ? sorted_ts[spos] : 0; spos++; }
        }
        uint64_t uenc[20], senc[20], uext[20];
        orc_log_query_encode(&ui, uenc);
        orc_log_query_encode(&si, senc);
        senc[19] += (uint64_t)ts << 8; /* TimestampedStorageLogRecord::encode, :98-109 */
        if (should_pop) {
            orc_log_queue_absorb(uq.head, uenc, NULL); uq.length--;
            orc_log_queue_absorb(sq.head, senc, NULL); sq.length--;
        }
        memcpy(uext, uenc, sizeof uext);
        uext[19] += (uint64_t)original_timestamp << 8; /* :605-610 */

        const int shard_ok = ZKC_LQ_SHARD(si.flags) == shard;
        if (should_pop && !shard_ok) fail(&st, (int64_t)cyc, ZKC_ST_CHK_SHARD_ID);

        uint64_t chain[4][20], gp_new[4];
        for (int rep = 0; rep < 2; rep++) {
            uint64_t lc = ch[rep][20], rc = ch[rep][20];
            for (int i = 0; i < 20; i++) {
                lc = gl_fma(uext[i], ch[rep][i], lc); chain[rep * 2][i] = lc;
                rc = gl_fma(senc[i], ch[rep][i], rc); chain[rep * 2 + 1][i] = rc;
            }
            gp_new[rep * 2] = gl_mul(lhs[rep], lc);
            gp_new[rep * 2 + 1] = gl_mul(rhs[rep], rc);
            if (should_pop) { lhs[rep] = gp_new[rep * 2]; rhs[rep] = gp_new[rep * 2 + 1]; }
        }

        /* :630-648 */
        uint32_t packed_key[13];
        memcpy(packed_key, si.key, 32);
        memcpy(packed_key + 8, si.address, 20);
        /* unpacked_long_comparison(a = previous_packed_key, b = packed_key): b - a */
        uint32_t diff[13]; int bor[13], leq[13], borrow = 0, keys_equal = 1;
        for (int i = 0; i < 13; i++) {
            const uint64_t d = (uint64_t)packed_key[i] - c.packed_key[i] - (uint64_t)borrow;
            diff[i] = (uint32_t)d; borrow = (int)((d >> 32) & 1); bor[i] = borrow; leq[i] = diff[i] == 0;
            keys_equal &= leq[i];
        }
        const int previous_key_is_greater = borrow;
        const int not_trivial = !item_is_trivial;
        if (not_trivial && previous_key_is_greater) fail(&st, (int64_t)cyc, ZKC_ST_CHK_KEY_ORDER);
        const uint64_t td = (uint64_t)c.timestamp - ts;
        const uint32_t ts_diff = (uint32_t)td;
        const int previous_ts_is_less = (int)((td >> 32) & 1);
        const int must_enforce = keys_equal && not_trivial;
        if (must_enforce && !previous_ts_is_less) fail(&st, (int64_t)cyc, ZKC_ST_CHK_TIMESTAMP_ORDER);

        /* new cell, :654-752 */
        const int not_keys_equal = !keys_equal;
        if (cyc == 0 && start && should_pop && !not_keys_equal) fail(&st, (int64_t)cyc, ZKC_ST_CHK_FIRST_KEY_NONZERO);
        const int value_is_unchanged = memcmp(c.cur, c.base, 32) == 0;
        const int depth_is_zero = c.depth == 0;
        const int unchanged_not_by_rollback = value_is_unchanged && !depth_is_zero;
        const int issue_protective_read = (int)c.flag || unchanged_not_by_rollback;
        const int should_write = !value_is_unchanged;
        const zkc_log_query query = net_query(&c, should_write, shard);
        const int should_update = issue_protective_read || should_write;
        const int should_push = !previous_item_is_trivial && not_keys_equal && should_update;
        uint64_t penc[20], rounds[36], newtail[4];
        orc_log_query_encode(&query, penc);
        memcpy(newtail, rq.tail, 32);
        orc_log_queue_absorb(newtail, penc, rounds);
        if (should_push) {
            memcpy(rq.tail, newtail, 32); rq.length++;
            if (result_tails) memcpy(result_tails + 4 * pushes, newtail, 32);
            pushes++;
        }
        const int rw = ZKC_LQ_RW(si.flags), rollback = ZKC_LQ_ROLLBACK(si.flags);
        const int new_cell = not_trivial && not_keys_equal;
        if (new_cell) {
            memcpy(c.base, si.read_value, 32);
            memcpy(c.cur, rw ? si.written_value : si.read_value, 32);
            c.depth = rw ? 1 : 0;
            c.flag = !rw;
        }
        /* same cell, :756-825 */
        const int nt_same = not_trivial && keys_equal;
        const int read_same = nt_same && !rw, write_same = nt_same && rw;
        const int wnr = write_same && !rollback, wrb = write_same && rollback;
        if (wnr) c.depth = c.depth + 1;
        if (wrb) {
            if (c.depth == 0) fail(&st, (int64_t)cyc, ZKC_ST_CHK_DEPTH_UNDERFLOW);
            c.depth = c.depth - 1;
        }
        const int read_is_equal = memcmp(c.cur, si.read_value, 32) == 0;
        const int check_read = read_same || wnr;
        if (check_read && !read_is_equal) fail(&st, (int64_t)cyc, ZKC_ST_CHK_READ_CONSISTENCY);
        if (wnr) memcpy(c.cur, si.written_value, 32);
        if (wrb) memcpy(c.cur, si.read_value, 32);
        const int rollback_depth_is_zero = c.depth == 0;
        const int read_at_zero = rollback_depth_is_zero && read_same;
        if (read_at_zero) { memcpy(c.base, si.read_value, 32); c.flag = 1; }

        if (trace) {
            T(ZKC_ST_ORIGINAL_IS_EMPTY, cyc) = (uint64_t)o_empty; T(ZKC_ST_SORTED_IS_EMPTY, cyc) = (uint64_t)s_empty;
            T(ZKC_ST_SHOULD_POP, cyc) = (uint64_t)should_pop; T(ZKC_ST_ORIGINAL_TIMESTAMP, cyc) = original_timestamp;
            uint64_t flat[36];
            orc_log_query_flatten(&ui, flat);
            for (int i = 0; i < 36; i++) T(ZKC_ST_UNSORTED_ITEM + i, cyc) = flat[i];
            orc_log_query_flatten(&si, flat);
            for (int i = 0; i < 36; i++) T(ZKC_ST_SORTED_ITEM + i, cyc) = flat[i];
            T(ZKC_ST_SORTED_ITEM + 36, cyc) = ts;
            for (int i = 0; i < 20; i++) { T(ZKC_ST_UNSORTED_ENC + i, cyc) = uenc[i]; T(ZKC_ST_SORTED_ENC + i, cyc) = senc[i]; }
            T(ZKC_ST_UNSORTED_EXT19, cyc) = uext[19];
            for (int i = 0; i < 4; i++) { T(ZKC_ST_UNSORTED_HEAD + i, cyc) = uq.head[i]; T(ZKC_ST_SORTED_HEAD + i, cyc) = sq.head[i]; }
            T(ZKC_ST_UNSORTED_LEN, cyc) = uq.length; T(ZKC_ST_SORTED_LEN, cyc) = sq.length;
            T(ZKC_ST_SHARD_ID_IS_VALID, cyc) = (uint64_t)shard_ok;
            for (int k = 0; k < 4; k++) {
                for (int i = 0; i < 20; i++) T(ZKC_ST_GP_CHAIN + k * 20 + i, cyc) = chain[k][i];
                T(ZKC_ST_GP_NEW + k, cyc) = gp_new[k];
            }
            T(ZKC_ST_GP_ACC + 0, cyc) = lhs[0]; T(ZKC_ST_GP_ACC + 1, cyc) = rhs[0];
            T(ZKC_ST_GP_ACC + 2, cyc) = lhs[1]; T(ZKC_ST_GP_ACC + 3, cyc) = rhs[1];
            for (int i = 0; i < 13; i++) {
                T(ZKC_ST_CMP_DIFF + i, cyc) = diff[i]; T(ZKC_ST_CMP_BORROW + i, cyc) = (uint64_t)bor[i];
                T(ZKC_ST_CMP_LIMB_EQ + i, cyc) = (uint64_t)leq[i];
            }
            T(ZKC_ST_KEYS_ARE_EQUAL, cyc) = (uint64_t)keys_equal; T(ZKC_ST_PREVIOUS_KEY_IS_GREATER, cyc) = (uint64_t)previous_key_is_greater;
            T(ZKC_ST_TS_DIFF, cyc) = ts_diff; T(ZKC_ST_PREVIOUS_TIMESTAMP_IS_LESS, cyc) = (uint64_t)previous_ts_is_less;
            T(ZKC_ST_MUST_ENFORCE, cyc) = (uint64_t)must_enforce;
            T(ZKC_ST_VALUE_IS_UNCHANGED, cyc) = (uint64_t)value_is_unchanged; T(ZKC_ST_CURRENT_DEPTH_IS_ZERO, cyc) = (uint64_t)depth_is_zero;
            T(ZKC_ST_UNCHANGED_BUT_NOT_BY_ROLLBACK, cyc) = (uint64_t)unchanged_not_by_rollback;
            T(ZKC_ST_ISSUE_PROTECTIVE_READ, cyc) = (uint64_t)issue_protective_read; T(ZKC_ST_SHOULD_WRITE, cyc) = (uint64_t)should_write;
            T(ZKC_ST_SHOULD_UPDATE, cyc) = (uint64_t)should_update; T(ZKC_ST_SHOULD_PUSH, cyc) = (uint64_t)should_push;
            T(ZKC_ST_NEW_NON_TRIVIAL_CELL, cyc) = (uint64_t)new_cell;
            for (int i = 0; i < 20; i++) T(ZKC_ST_PUSH_ENC + i, cyc) = penc[i];
            for (int i = 0; i < 36; i++) T(ZKC_ST_PUSH_ROUND0 + i, cyc) = rounds[i];
            for (int i = 0; i < 4; i++) T(ZKC_ST_RESULT_TAIL + i, cyc) = rq.tail[i];
            T(ZKC_ST_RESULT_LEN, cyc) = rq.length;
            for (int i = 0; i < 8; i++) { T(ZKC_ST_CELL_BASE_VALUE + i, cyc) = c.base[i]; T(ZKC_ST_CELL_CURRENT_VALUE + i, cyc) = c.cur[i]; }
            T(ZKC_ST_CELL_CURRENT_DEPTH, cyc) = c.depth; T(ZKC_ST_CELL_HAS_READ_AT_DEPTH_ZERO, cyc) = c.flag;
            T(ZKC_ST_NON_TRIVIAL_AND_SAME_CELL, cyc) = (uint64_t)nt_same; T(ZKC_ST_READ_OF_SAME_CELL, cyc) = (uint64_t)read_same;
            T(ZKC_ST_WRITE_OF_SAME_CELL, cyc) = (uint64_t)write_same; T(ZKC_ST_WRITE_NO_ROLLBACK, cyc) = (uint64_t)wnr;
            T(ZKC_ST_WRITE_ROLLBACK, cyc) = (uint64_t)wrb; T(ZKC_ST_READ_IS_EQUAL_TO_CURRENT, cyc) = (uint64_t)read_is_equal;
            T(ZKC_ST_CHECK_READ_CONSISTENCY, cyc) = (uint64_t)check_read; T(ZKC_ST_ROLLBACK_DEPTH_IS_ZERO, cyc) = (uint64_t)rollback_depth_is_zero;
            T(ZKC_ST_READ_AT_DEPTH_ZERO_OF_SAME_CELL, cyc) = (uint64_t)read_at_zero;
        }
        /* :827-832 */
        memcpy(c.address, si.address, sizeof c.address);
        memcpy(c.key, si.key, sizeof c.key);
        previous_item_is_trivial = item_is_trivial;
        c.timestamp = ts;
        memcpy(c.packed_key, packed_key, sizeof packed_key);
    }
    /* finalisation, :836-880 */
    {
        const int queues_exhausted = uq.length == 0;
        const int value_is_unchanged = memcmp(c.cur, c.base, 32) == 0;
        const int unchanged_not_by_rollback = value_is_unchanged && c.depth != 0;
        const int issue_protective_read = (int)c.flag || unchanged_not_by_rollback;
        const int should_write = !value_is_unchanged;
        const int should_push = !previous_item_is_trivial && (issue_protective_read || should_write) && queues_exhausted;
        if (should_push) {
            const zkc_log_query query = net_query(&c, should_write, shard);
            uint64_t penc[20];
            orc_log_query_encode(&query, penc);
            orc_log_queue_absorb(rq.tail, penc, NULL);
            rq.length++;
            if (result_tails) memcpy(result_tails + 4 * pushes, rq.tail, 32);
            pushes++;
        }
        if (queues_exhausted) c.flag = 0; /* :872-879 */
    }
    if (n_result_tails) *n_result_tails = pushes;
    if (uq.length == 0 && memcmp(uq.head, uq.tail, 32)) fail(&st, -1, ZKC_ST_CHK_QUEUE_CONSISTENCY);
    if (sq.length == 0 && memcmp(sq.head, sq.tail, 32)) fail(&st, -1, ZKC_ST_CHK_QUEUE_CONSISTENCY);
    if ((uq.length == 0) != (sq.length == 0)) fail(&st, -1, ZKC_ST_CHK_EMPTY_SYNC);
    const int completed = uq.length == 0 && sq.length == 0;
    if (completed && (lhs[0] != rhs[0] || lhs[1] != rhs[1])) fail(&st, -1, ZKC_ST_CHK_GRAND_PRODUCT);

    zkc_storage_fsm out;
    memset(&out, 0, sizeof out);
    out.cycle_idx = cycle_idx;
    memcpy(out.previous_packed_key, c.packed_key, sizeof c.packed_key);
    memcpy(out.previous_key, c.key, sizeof c.key);
    memcpy(out.previous_address, c.address, sizeof c.address);
    out.previous_timestamp = c.timestamp;
    out.this_cell_has_explicit_read_and_rollback_depth_zero = c.flag;
    memcpy(out.this_cell_base_value, c.base, 32);
    memcpy(out.this_cell_current_value, c.cur, 32);
    out.this_cell_current_depth = c.depth;
    for (int i = 0; i < 2; i++) { out.lhs_accumulator[i] = lhs[i]; out.rhs_accumulator[i] = rhs[i]; }
    out.current_unsorted_queue_state = uq;
    out.current_intermediate_sorted_queue_state = sq;
    out.current_final_sorted_queue_state = rq;
    zkc_queue_state4 obs_out;
    memset(&obs_out, 0, sizeof obs_out);
    if (completed) obs_out = rq;

    if (options && options->compare_expected) {
        uint64_t a[77], b[77], c9[9], d9[9];
        orc_storage_encode_fsm(&out, a); orc_storage_encode_fsm(&io->hidden_fsm_output, b);
        orc_put_queue_state4(c9, &obs_out); orc_put_queue_state4(d9, &io->final_sorted_queue_state);
        if (memcmp(a, b, sizeof a) || memcmp(c9, d9, sizeof c9) || (io->completion_flag != 0) != completed)
            if (st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    io->hidden_fsm_output = out;
    io->final_sorted_queue_state = obs_out;
    io->completion_flag = (uint32_t)completed;

    uint64_t e_in[19], e_out[9], e_fin[77], e_fout[77];
    size_t n_in = 0;
    e_in[n_in++] = shard;
    n_in += orc_put_queue_state4(e_in + n_in, &io->unsorted_log_queue_state);
    n_in += orc_put_queue_state4(e_in + n_in, &io->intermediate_sorted_queue_state);
    const size_t n_out = orc_put_queue_state4(e_out, &obs_out);
    const size_t n_fin = orc_storage_encode_fsm(fin, e_fin);
    const size_t n_fout = orc_storage_encode_fsm(&out, e_fout);
    orc_closed_form_commitment(start, completed, e_in, n_in, e_out, n_out, e_fin, n_fin, e_fout, n_fout, commitment);
    if (status) *status = st;
    return st.code;
}
