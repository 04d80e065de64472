/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Sequential CPU restatement of the events / L2->L1 message sorter:
 *   sort_and_deduplicate_events_entry_point   /root/reference/src/log_sorter/mod.rs:34-232
 *   repack_and_prove_events_rollbacks_inner   /root/reference/src/log_sorter/mod.rs:234-441
 * Pinning: loop logic pinned by the reference's test vector (mod.rs:637-816, every enforcement
 * holds); hash-dependent values PARITY UNPINNED (Poseidon2, see poseidon2.c).
 */
#include "oracle.h"
#include <string.h>

static void fail(zkc_status *st, int64_t row, uint32_t bit) {
    st->code = ZKC_ERR_UNSATISFIED;
    st->failed_checks |= bit;
    if (row >= 0 && (st->first_bad_row < 0 || row < st->first_bad_row)) st->first_bad_row = row;
}

size_t orc_events_encode_fsm(const zkc_events_fsm *f, uint64_t *dst) {
    size_t n = 0;
    dst[n++] = f->lhs_accumulator[0]; dst[n++] = f->lhs_accumulator[1];
    dst[n++] = f->rhs_accumulator[0]; dst[n++] = f->rhs_accumulator[1];
    n += orc_put_queue_state4(dst + n, &f->initial_unsorted_queue_state);
    n += orc_put_queue_state4(dst + n, &f->intermediate_sorted_queue_state);
    n += orc_put_queue_state4(dst + n, &f->final_result_queue_state);
    dst[n++] = f->previous_key;
    orc_log_query_flatten(&f->previous_item, dst + n);
    return n + 36; /* 68 */
}

/* query_to_add, mod.rs:381-393 */
static zkc_log_query cleaned_up(const zkc_log_query *p) {
    zkc_log_query q;
    memset(&q, 0, sizeof q);
    memcpy(q.address, p->address, sizeof q.address);
    memcpy(q.key, p->key, sizeof q.key);
    memcpy(q.written_value, p->written_value, sizeof q.written_value);
    q.tx_number_in_block = p->tx_number_in_block;
    q.flags = ZKC_LQ_FLAGS(0, ZKC_LQ_SHARD(p->flags), 0, 0, ZKC_LQ_SERVICE(p->flags));
    return q;
}

#define T(col, r) trace[(size_t)(col) * limit + (r)]

int orc_log_sorter_entry_point(zkc_events_closed_form *io, const zkc_log_query *unsorted, size_t n_unsorted,
                               const zkc_log_query *sorted, size_t n_sorted, size_t limit,
                               const zkc_sorter_options *options, uint64_t *trace, uint64_t *result_tails,
                               size_t *n_result_tails, uint64_t commitment[4], zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    const int start = io->start_flag != 0;
    const zkc_events_fsm *fin = &io->hidden_fsm_input;
    static const uint64_t zero4[4] = {0, 0, 0, 0};
    if (memcmp(io->initial_log_queue_state.head, zero4, 32) || memcmp(io->intermediate_sorted_queue_state.head, zero4, 32))
        fail(&st, -1, ZKC_EV_CHK_TRIVIAL_HEAD); /* :61, :87 */
    zkc_queue_state4 uq = start ? io->initial_log_queue_state : fin->initial_unsorted_queue_state;
    zkc_queue_state4 sq = start ? io->intermediate_sorted_queue_state : fin->intermediate_sorted_queue_state;
    zkc_queue_state4 rq;
    memset(&rq, 0, sizeof rq);
    if (!start) rq = fin->final_result_queue_state; /* :100-109 */

    uint64_t ch[2][21];
    orc_produce_fs_challenges(io->initial_log_queue_state.tail, io->initial_log_queue_state.length,
                              io->intermediate_sorted_queue_state.tail, io->intermediate_sorted_queue_state.length, 4, 21,
                              &ch[0][0]);
    uint64_t lhs[2], rhs[2];
    for (int i = 0; i < 2; i++) {
        lhs[i] = start ? 1 : fin->lhs_accumulator[i];
        rhs[i] = start ? 1 : fin->rhs_accumulator[i];
    }
    uint32_t previous_key = start ? 0 : fin->previous_key; /* :148-154 */
    zkc_log_query previous_item;
    memset(&previous_item, 0, sizeof previous_item);
    if (!start) previous_item = fin->previous_item; /* :157-164 */

    /* repack_and_prove_events_rollbacks_inner */
    const int no_work = uq.length == 0;
    int previous_is_trivial = no_work || start; /* :266-267 */
    if (uq.length != sq.length) fail(&st, -1, ZKC_EV_CHK_LENGTHS_EQUAL);

    size_t upos = 0, spos = 0, pushes = 0;
    for (size_t cyc = 0; cyc < limit; cyc++) {
        const int o_empty = uq.length == 0, s_empty = sq.length == 0;
        if (o_empty != s_empty) fail(&st, (int64_t)cyc, ZKC_EV_CHK_EMPTY_SYNC);
        const int should_pop = !o_empty, is_trivial = o_empty;
        zkc_log_query ui, si;
        memset(&ui, 0, sizeof ui); memset(&si, 0, sizeof si);
        if (should_pop) {
            if (upos < n_unsorted) ui = unsorted[upos++];
            if (spos < n_sorted) si = sorted[spos++];
        }
        uint64_t uenc[20], senc[20];
        orc_log_query_encode(&ui, uenc);
        orc_log_query_encode(&si, senc);
        if (should_pop) {
            orc_log_queue_absorb(uq.head, uenc, NULL); uq.length--;
            orc_log_queue_absorb(sq.head, senc, NULL); sq.length--;
        }
        if (should_pop && !ZKC_LQ_RW(ui.flags)) fail(&st, (int64_t)cyc, ZKC_EV_CHK_UNSORTED_IS_WRITE); /* :295-297 */

        uint64_t chain[4][20], gp_new[4]; /* :299-313 */
        for (int rep = 0; rep < 2; rep++) {
            uint64_t lc = ch[rep][20], rc = ch[rep][20];
            for (int i = 0; i < 20; i++) {
                lc = gl_fma(uenc[i], ch[rep][i], lc); chain[rep * 2][i] = lc;
                rc = gl_fma(senc[i], ch[rep][i], rc); chain[rep * 2 + 1][i] = rc;
            }
            gp_new[rep * 2] = gl_mul(lhs[rep], lc);
            gp_new[rep * 2 + 1] = gl_mul(rhs[rep], rc);
            if (should_pop) { lhs[rep] = gp_new[rep * 2]; rhs[rep] = gp_new[rep * 2 + 1]; }
        }

        if (should_pop && !ZKC_LQ_RW(si.flags)) fail(&st, (int64_t)cyc, ZKC_EV_CHK_SORTED_IS_WRITE); /* :318-320 */
        const uint32_t sorting_key = si.timestamp;
        /* unpacked_long_comparison(a = [previous_key], b = [sorting_key]): b - a */
        const uint64_t d = (uint64_t)sorting_key - previous_key;
        const uint32_t diff = (uint32_t)d;
        const int new_key_is_smaller = (int)((d >> 32) & 1), keys_equal = diff == 0;
        if (should_pop && new_key_is_smaller) fail(&st, (int64_t)cyc, ZKC_EV_CHK_ORDER); /* :331 */
        const int same_log = keys_equal;
        const int same_nontrivial = should_pop && same_log;
        const int maybe_different = !same_log;
        const int different_nontrivial = should_pop && maybe_different;
        const int rollback = ZKC_LQ_ROLLBACK(si.flags);
        if (different_nontrivial && rollback) fail(&st, (int64_t)cyc, ZKC_EV_CHK_NOT_ROLLBACK); /* :342-343 */
        if (same_nontrivial && !rollback) fail(&st, (int64_t)cyc, ZKC_EV_CHK_IS_ROLLBACK);     /* :347-349 */
        const int item_keys_equal = memcmp(si.key, previous_item.key, 32) == 0;
        const int values_equal = memcmp(si.written_value, previous_item.written_value, 32) == 0;
        const int same_body = item_keys_equal && values_equal;
        const int previous_non_trivial = !previous_is_trivial;
        const int should_enforce = same_log && previous_non_trivial;
        if (should_enforce && !same_body) fail(&st, (int64_t)cyc, ZKC_EV_CHK_SAME_BODY); /* :362 */
        const int prev_not_rollback = !ZKC_LQ_ROLLBACK(previous_item.flags);
        const int maybe_add = maybe_different || is_trivial;
        const int add = previous_non_trivial && maybe_add && prev_not_rollback;
        const zkc_log_query to_add = cleaned_up(&previous_item);
        uint64_t penc[20], rounds[36], newtail[4];
        orc_log_query_encode(&to_add, penc);
        memcpy(newtail, rq.tail, 32);
        orc_log_queue_absorb(newtail, penc, rounds);
        if (add) {
            memcpy(rq.tail, newtail, 32); rq.length++;
            if (result_tails) memcpy(result_tails + 4 * pushes, newtail, 32);
            pushes++;
        }

        if (trace) {
            T(ZKC_EV_ORIGINAL_IS_EMPTY, cyc) = (uint64_t)o_empty; T(ZKC_EV_SORTED_IS_EMPTY, cyc) = (uint64_t)s_empty;
            T(ZKC_EV_SHOULD_POP, cyc) = (uint64_t)should_pop;
            uint64_t flat[36];
            orc_log_query_flatten(&ui, flat);
            for (int i = 0; i < 36; i++) T(ZKC_EV_UNSORTED_ITEM + i, cyc) = flat[i];
            orc_log_query_flatten(&si, flat);
            for (int i = 0; i < 36; i++) T(ZKC_EV_SORTED_ITEM + i, cyc) = flat[i];
            for (int i = 0; i < 20; i++) { T(ZKC_EV_UNSORTED_ENC + i, cyc) = uenc[i]; T(ZKC_EV_SORTED_ENC + i, cyc) = senc[i]; }
            for (int i = 0; i < 4; i++) { T(ZKC_EV_UNSORTED_HEAD + i, cyc) = uq.head[i]; T(ZKC_EV_SORTED_HEAD + i, cyc) = sq.head[i]; }
            T(ZKC_EV_UNSORTED_LEN, cyc) = uq.length; T(ZKC_EV_SORTED_LEN, cyc) = sq.length;
            for (int k = 0; k < 4; k++) {
                for (int i = 0; i < 20; i++) T(ZKC_EV_GP_CHAIN + k * 20 + i, cyc) = chain[k][i];
                T(ZKC_EV_GP_NEW + k, cyc) = gp_new[k];
            }
            T(ZKC_EV_GP_ACC + 0, cyc) = lhs[0]; T(ZKC_EV_GP_ACC + 1, cyc) = rhs[0];
            T(ZKC_EV_GP_ACC + 2, cyc) = lhs[1]; T(ZKC_EV_GP_ACC + 3, cyc) = rhs[1];
            T(ZKC_EV_CMP_DIFF, cyc) = diff; T(ZKC_EV_CMP_BORROW, cyc) = (uint64_t)new_key_is_smaller;
            T(ZKC_EV_KEYS_EQUAL, cyc) = (uint64_t)keys_equal;
            T(ZKC_EV_SAME_NONTRIVIAL_LOG, cyc) = (uint64_t)same_nontrivial;
            T(ZKC_EV_DIFFERENT_NONTRIVIAL_LOG, cyc) = (uint64_t)different_nontrivial;
            T(ZKC_EV_ITEM_KEYS_EQUAL, cyc) = (uint64_t)item_keys_equal; T(ZKC_EV_VALUES_EQUAL, cyc) = (uint64_t)values_equal;
            T(ZKC_EV_SAME_BODY, cyc) = (uint64_t)same_body; T(ZKC_EV_PREVIOUS_IS_TRIVIAL, cyc) = (uint64_t)previous_is_trivial;
            T(ZKC_EV_SHOULD_ENFORCE, cyc) = (uint64_t)should_enforce; T(ZKC_EV_MAYBE_ADD, cyc) = (uint64_t)maybe_add;
            T(ZKC_EV_ADD_TO_QUEUE, cyc) = (uint64_t)add;
            for (int i = 0; i < 20; i++) T(ZKC_EV_PUSH_ENC + i, cyc) = penc[i];
            for (int i = 0; i < 36; i++) T(ZKC_EV_PUSH_ROUND0 + i, cyc) = rounds[i];
            for (int i = 0; i < 4; i++) T(ZKC_EV_RESULT_TAIL + i, cyc) = rq.tail[i];
            T(ZKC_EV_RESULT_LEN, cyc) = rq.length;
        }
        previous_is_trivial = is_trivial;
        previous_item = si;
        previous_key = sorting_key;
    }
    /* finalisation, :406-435 */
    {
        const int now_empty = uq.length == 0;
        const int add = !previous_is_trivial && !ZKC_LQ_ROLLBACK(previous_item.flags) && now_empty;
        if (add) {
            const zkc_log_query to_add = cleaned_up(&previous_item);
            uint64_t penc[20];
            orc_log_query_encode(&to_add, penc);
            orc_log_queue_absorb(rq.tail, penc, NULL);
            rq.length++;
            if (result_tails) memcpy(result_tails + 4 * pushes, rq.tail, 32);
            pushes++;
        }
    }
    if (n_result_tails) *n_result_tails = pushes;
    /* :437-438 enforce_consistency */
    if (uq.length == 0 && memcmp(uq.head, uq.tail, 32)) fail(&st, -1, ZKC_EV_CHK_QUEUE_CONSISTENCY);
    if (sq.length == 0 && memcmp(sq.head, sq.tail, 32)) fail(&st, -1, ZKC_EV_CHK_QUEUE_CONSISTENCY);
    /* entry point :183-191 */
    if ((uq.length == 0) != (sq.length == 0)) fail(&st, -1, ZKC_EV_CHK_EMPTY_SYNC);
    const int completed = uq.length == 0;
    if (completed && (lhs[0] != rhs[0] || lhs[1] != rhs[1])) fail(&st, -1, ZKC_EV_CHK_GRAND_PRODUCT);

    zkc_events_fsm out;
    memset(&out, 0, sizeof out);
    out.previous_key = previous_key;
    out.previous_item = previous_item;
    for (int i = 0; i < 2; i++) { out.lhs_accumulator[i] = lhs[i]; out.rhs_accumulator[i] = rhs[i]; }
    out.initial_unsorted_queue_state = uq;
    out.intermediate_sorted_queue_state = sq;
    out.final_result_queue_state = rq;
    zkc_queue_state4 obs_out;
    memset(&obs_out, 0, sizeof obs_out);
    if (completed) obs_out = rq; /* :207-215 */

    if (options && options->compare_expected) {
        uint64_t a[68], b[68];
        orc_events_encode_fsm(&out, a); orc_events_encode_fsm(&io->hidden_fsm_output, b);
        uint64_t c[9], d9[9];
        orc_put_queue_state4(c, &obs_out); orc_put_queue_state4(d9, &io->final_queue_state);
        if (memcmp(a, b, sizeof a) || memcmp(c, d9, sizeof c) || (io->completion_flag != 0) != completed)
            if (st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    io->hidden_fsm_output = out;
    io->final_queue_state = obs_out;
    io->completion_flag = (uint32_t)completed;

    uint64_t e_in[18], e_out[9], e_fin[68], e_fout[68];
    size_t n_in = orc_put_queue_state4(e_in, &io->initial_log_queue_state);
    n_in += orc_put_queue_state4(e_in + n_in, &io->intermediate_sorted_queue_state);
    const size_t n_out = orc_put_queue_state4(e_out, &obs_out);
    const size_t n_fin = orc_events_encode_fsm(fin, e_fin);
    const size_t n_fout = orc_events_encode_fsm(&out, e_fout);
    orc_closed_form_commitment(start, completed, e_in, n_in, e_out, n_out, e_fin, n_fin, e_fout, n_fout, commitment);
    if (status) *status = st;
    return st.code;
}
