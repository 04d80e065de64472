/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Sequential CPU restatement of the code-decommitment request sorter / deduplicator:
 *   DecommitQuery::encode                                  /root/reference/src/base_structures/decommit_query/mod.rs:31-107
 *   sort_and_deduplicate_code_decommittments_entry_point   /root/reference/src/sort_decommittment_requests/mod.rs:40-233
 *   sort_and_deduplicate_code_decommittments_inner         /root/reference/src/sort_decommittment_requests/mod.rs:235-381
 *   concatenate_key                                        /root/reference/src/sort_decommittment_requests/mod.rs:383-401
 *   unpacked_long_comparison                               /root/reference/src/storage_validity_by_grand_product/mod.rs:925-944
 * Pinning: loop logic pinned by the reference's test vector (mod.rs:565-1390, limit 16: every enforcement
 * holds); hash-dependent values PARITY UNPINNED (Poseidon2, see poseidon2.c).
 */
#include "oracle.h"
#include <string.h>

static void fail(zkc_status *st, int64_t row, uint32_t bit) {
    st->code = ZKC_ERR_UNSATISFIED;
    st->failed_checks |= bit;
    if (row >= 0 && (st->first_bad_row < 0 || row < st->first_bad_row)) st->first_bad_row = row;
}

/* decommit_query/mod.rs:31-107 */
void orc_decommit_query_encode(const zkc_decommit_query *q, uint64_t out[8]) {
    uint8_t pb[4], tb[4];
    for (int i = 0; i < 4; i++) { pb[i] = (uint8_t)(q->page >> (8 * i)); tb[i] = (uint8_t)(q->timestamp >> (8 * i)); }
    out[0] = (uint64_t)q->code_hash[0] + ((uint64_t)pb[0] << 32) + ((uint64_t)pb[1] << 40) + ((uint64_t)pb[2] << 48);
    out[1] = (uint64_t)q->code_hash[1] + ((uint64_t)pb[3] << 32) + ((uint64_t)tb[0] << 40) + ((uint64_t)tb[1] << 48);
    out[2] = (uint64_t)q->code_hash[2] + ((uint64_t)tb[2] << 32) + ((uint64_t)tb[3] << 40) + ((uint64_t)(q->is_first & 1) << 48);
    for (int i = 3; i < 8; i++) out[i] = q->code_hash[i];
}

/* decommit_query/mod.rs:133-150 */
void orc_decommit_query_flatten(const zkc_decommit_query *q, uint64_t out[11]) {
    for (int i = 0; i < 8; i++) out[i] = q->code_hash[i];
    out[8] = q->page; out[9] = q->is_first & 1; out[10] = q->timestamp;
}

/* FullStateCircuitQueue::push / pop_front: state' = P(enc[0..8] || state[8..12]) */
static void full_state_absorb(uint64_t state[12], const uint64_t enc[8]) {
    memcpy(state, enc, 8 * sizeof(uint64_t));
    orc_poseidon2_permutation(state);
}

void orc_decommit_queue_simulate(const zkc_decommit_query *q, size_t n, uint64_t *prev_states, zkc_queue_state12 *final_state) {
    uint64_t tail[12] = {0}, enc[8];
    for (size_t i = 0; i < n; i++) {
        if (prev_states) memcpy(prev_states + 12 * i, tail, sizeof tail);
        orc_decommit_query_encode(&q[i], enc);
        full_state_absorb(tail, enc);
    }
    memset(final_state, 0, sizeof *final_state);
    memcpy(final_state->tail, tail, sizeof tail);
    final_state->length = (uint32_t)n;
}

static size_t put_queue_state12(uint64_t *dst, const zkc_queue_state12 *s) {
    memcpy(dst, s->head, 96);
    memcpy(dst + 12, s->tail, 96);
    dst[24] = s->length;
    return 25;
}

/* CSVarLengthEncodable order of CodeDecommittmentsDeduplicatorFSMInputOutput, input.rs:26-38 */
size_t orc_decommit_sorter_encode_fsm(const zkc_decommit_sorter_fsm *f, uint64_t *dst) {
    size_t n = put_queue_state12(dst, &f->initial_queue_state);
    n += put_queue_state12(dst + n, &f->sorted_queue_state);
    n += put_queue_state12(dst + n, &f->final_queue_state);
    dst[n++] = f->lhs_accumulator[0]; dst[n++] = f->lhs_accumulator[1];
    dst[n++] = f->rhs_accumulator[0]; dst[n++] = f->rhs_accumulator[1];
    for (int i = 0; i < 9; i++) dst[n++] = f->previous_packed_key[i];
    dst[n++] = f->first_encountered_timestamp;
    orc_decommit_query_flatten(&f->previous_record, dst + n);
    return n + 11; /* 100 */
}

#define T(col, r) trace[(size_t)(col) * limit + (r)]

int orc_sort_decommittments_entry_point(zkc_decommit_sorter_closed_form *io, const zkc_decommit_query *unsorted, size_t n_unsorted,
                                        const zkc_decommit_query *sorted, size_t n_sorted, size_t limit,
                                        const zkc_sorter_options *options, uint64_t *trace, uint64_t *result_states,
                                        size_t *n_result_states, uint64_t commitment[4], zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    const int start = io->start_flag != 0;
    const zkc_decommit_sorter_fsm *fin = &io->hidden_fsm_input;
    static const uint64_t zero12[12] = {0};
    if (memcmp(io->initial_queue_state.head, zero12, 96) || memcmp(io->sorted_queue_initial_state.head, zero12, 96))
        fail(&st, -1, ZKC_DQ_CHK_TRIVIAL_HEAD); /* :78, :93 */
    zkc_queue_state12 uq = start ? io->initial_queue_state : fin->initial_queue_state;
    zkc_queue_state12 sq = start ? io->sorted_queue_initial_state : fin->sorted_queue_state;
    zkc_queue_state12 rq;
    memset(&rq, 0, sizeof rq);
    if (!start) rq = fin->final_queue_state; /* :104-114 */

    uint64_t ch[2][9]; /* :116-132 */
    orc_produce_fs_challenges(io->initial_queue_state.tail, io->initial_queue_state.length, io->sorted_queue_initial_state.tail,
                              io->sorted_queue_initial_state.length, 12, 9, &ch[0][0]);
    uint64_t lhs[2], rhs[2];
    for (int i = 0; i < 2; i++) {
        lhs[i] = start ? 1 : fin->lhs_accumulator[i];
        rhs[i] = start ? 1 : fin->rhs_accumulator[i];
    }
    zkc_decommit_query previous_record; /* :150-156 */
    memset(&previous_record, 0, sizeof previous_record);
    if (!start) previous_record = fin->previous_record;
    previous_record._pad = 0;
    uint32_t previous_packed_key[9] = {0}; /* :158-164 */
    if (!start) memcpy(previous_packed_key, fin->previous_packed_key, sizeof previous_packed_key);
    uint32_t first_ts = start ? 0 : fin->first_encountered_timestamp; /* :166-173 */

    /* sort_and_deduplicate_code_decommittments_inner */
    if (uq.length != sq.length) fail(&st, -1, ZKC_DQ_CHK_LENGTHS_EQUAL); /* :262-269 */
    const int no_work = uq.length == 0;
    int previous_is_trivial = no_work || start; /* :271-273 */

    size_t upos = 0, spos = 0, pushes = 0;
    for (size_t cyc = 0; cyc < limit; cyc++) {
        const int o_empty = uq.length == 0, s_empty = sq.length == 0;
        if (o_empty != s_empty) fail(&st, (int64_t)cyc, ZKC_DQ_CHK_EMPTY_SYNC);
        const int should_pop = !o_empty, is_trivial = o_empty;
        zkc_decommit_query ui, si;
        memset(&ui, 0, sizeof ui); memset(&si, 0, sizeof si);
        if (should_pop) {
            if (upos < n_unsorted) ui = unsorted[upos++];
            if (spos < n_sorted) si = sorted[spos++];
            ui._pad = si._pad = 0; ui.is_first &= 1; si.is_first &= 1;
        }
        uint64_t uenc[8], senc[8];
        orc_decommit_query_encode(&ui, uenc);
        orc_decommit_query_encode(&si, senc);
        if (should_pop) {
            full_state_absorb(uq.head, uenc); uq.length--;
            /* a sorted queue shorter than the original one: its pop does not execute meaningfully; the
             * length mismatch is already reported above */
            full_state_absorb(sq.head, senc); if (sq.length) sq.length--;
        }

        uint64_t chain[4][8], gp_new[4]; /* :289-304 */
        for (int rep = 0; rep < 2; rep++) {
            uint64_t lc = ch[rep][8], rc = ch[rep][8];
            for (int i = 0; i < 8; i++) {
                lc = gl_fma(uenc[i], ch[rep][i], lc); chain[rep * 2][i] = lc;
                rc = gl_fma(senc[i], ch[rep][i], rc); chain[rep * 2 + 1][i] = rc;
            }
            gp_new[rep * 2] = gl_mul(lhs[rep], lc);
            gp_new[rep * 2 + 1] = gl_mul(rhs[rep], rc);
            if (should_pop) { lhs[rep] = gp_new[rep * 2]; rhs[rep] = gp_new[rep * 2 + 1]; }
        }

        /* :306-312: packed_key = [timestamp, hash limbs]; unpacked_long_comparison(a = packed_key, b = previous): b - a */
        uint32_t packed_key[9];
        packed_key[0] = si.timestamp;
        for (int i = 0; i < 8; i++) packed_key[1 + i] = si.code_hash[i];
        uint32_t diff[9]; int bor[9], leq[9], borrow = 0, keys_equal = 1;
        for (int i = 0; i < 9; i++) {
            const uint64_t d = (uint64_t)previous_packed_key[i] - packed_key[i] - (uint64_t)borrow;
            diff[i] = (uint32_t)d; borrow = (int)((d >> 32) & 1); bor[i] = borrow; leq[i] = diff[i] == 0;
            keys_equal &= leq[i];
        }
        const int new_key_is_greater = borrow;
        if (should_pop && !new_key_is_greater) fail(&st, (int64_t)cyc, ZKC_DQ_CHK_ORDER);

        const int same_hash = memcmp(previous_record.code_hash, si.code_hash, 32) == 0; /* :314 */
        const int different_hash = !same_hash;
        const int enforce_must_be_first = different_hash && should_pop;
        if (enforce_must_be_first && !si.is_first) fail(&st, (int64_t)cyc, ZKC_DQ_CHK_MUST_BE_FIRST);
        const int previous_is_non_trivial = !previous_is_trivial;
        const int enforce_same_memory_page = same_hash && previous_is_non_trivial;
        if (enforce_same_memory_page && si.page != previous_record.page) fail(&st, (int64_t)cyc, ZKC_DQ_CHK_SAME_MEMORY_PAGE);

        /* :335-341: maybe add the PREVIOUS record to the result queue */
        const int add = previous_is_non_trivial && different_hash;
        zkc_decommit_query to_add = previous_record;
        to_add.is_first = 1;
        to_add.timestamp = first_ts;
        uint64_t penc[8], newtail[12];
        orc_decommit_query_encode(&to_add, penc);
        memcpy(newtail, rq.tail, 96);
        full_state_absorb(newtail, penc);
        if (add) {
            memcpy(rq.tail, newtail, 96); rq.length++;
            if (result_states) memcpy(result_states + 12 * pushes, newtail, 96);
            pushes++;
        }

        previous_is_trivial = is_trivial;
        if (!same_hash) first_ts = si.timestamp; /* :345-350 */
        previous_record = si;
        memcpy(previous_packed_key, packed_key, sizeof packed_key);

        if (trace) {
            T(ZKC_DQ_ORIGINAL_IS_EMPTY, cyc) = (uint64_t)o_empty; T(ZKC_DQ_SORTED_IS_EMPTY, cyc) = (uint64_t)s_empty;
            T(ZKC_DQ_SHOULD_POP, cyc) = (uint64_t)should_pop;
            uint64_t flat[11];
            orc_decommit_query_flatten(&ui, flat);
            for (int i = 0; i < 11; i++) T(ZKC_DQ_UNSORTED_ITEM + i, cyc) = flat[i];
            orc_decommit_query_flatten(&si, flat);
            for (int i = 0; i < 11; i++) T(ZKC_DQ_SORTED_ITEM + i, cyc) = flat[i];
            for (int i = 0; i < 8; i++) { T(ZKC_DQ_UNSORTED_ENC + i, cyc) = uenc[i]; T(ZKC_DQ_SORTED_ENC + i, cyc) = senc[i]; }
            for (int i = 0; i < 12; i++) { T(ZKC_DQ_UNSORTED_HEAD + i, cyc) = uq.head[i]; T(ZKC_DQ_SORTED_HEAD + i, cyc) = sq.head[i]; }
            T(ZKC_DQ_UNSORTED_LEN, cyc) = uq.length; T(ZKC_DQ_SORTED_LEN, cyc) = sq.length;
            for (int k = 0; k < 4; k++) {
                for (int i = 0; i < 8; i++) T(ZKC_DQ_GP_CHAIN + k * 8 + i, cyc) = chain[k][i];
                T(ZKC_DQ_GP_NEW + k, cyc) = gp_new[k];
            }
            T(ZKC_DQ_GP_ACC + 0, cyc) = lhs[0]; T(ZKC_DQ_GP_ACC + 1, cyc) = rhs[0];
            T(ZKC_DQ_GP_ACC + 2, cyc) = lhs[1]; T(ZKC_DQ_GP_ACC + 3, cyc) = rhs[1];
            for (int i = 0; i < 9; i++) {
                T(ZKC_DQ_CMP_DIFF + i, cyc) = diff[i]; T(ZKC_DQ_CMP_BORROW + i, cyc) = (uint64_t)bor[i];
                T(ZKC_DQ_CMP_LIMB_EQ + i, cyc) = (uint64_t)leq[i];
            }
            T(ZKC_DQ_KEYS_ARE_EQUAL, cyc) = (uint64_t)keys_equal;
            T(ZKC_DQ_SAME_HASH, cyc) = (uint64_t)same_hash;
            T(ZKC_DQ_ENFORCE_MUST_BE_FIRST, cyc) = (uint64_t)enforce_must_be_first;
            T(ZKC_DQ_PREVIOUS_IS_TRIVIAL, cyc) = (uint64_t)!previous_is_non_trivial;
            T(ZKC_DQ_ENFORCE_SAME_MEMORY_PAGE, cyc) = (uint64_t)enforce_same_memory_page;
            T(ZKC_DQ_ADD_TO_QUEUE, cyc) = (uint64_t)add;
            orc_decommit_query_flatten(&to_add, flat);
            for (int i = 0; i < 11; i++) T(ZKC_DQ_PUSH_ITEM + i, cyc) = flat[i];
            for (int i = 0; i < 8; i++) T(ZKC_DQ_PUSH_ENC + i, cyc) = penc[i];
            for (int i = 0; i < 12; i++) T(ZKC_DQ_RESULT_TAIL + i, cyc) = rq.tail[i];
            T(ZKC_DQ_RESULT_LEN, cyc) = rq.length;
            T(ZKC_DQ_FIRST_TIMESTAMP, cyc) = first_ts;
        }
    }
    /* :357-362 */
    const int completed = uq.length == 0;
    if (completed != (sq.length == 0)) fail(&st, -1, ZKC_DQ_CHK_EMPTY_SYNC);
    /* finalisation, :364-375 */
    if (!previous_is_trivial && completed) {
        zkc_decommit_query to_add = previous_record;
        to_add.is_first = 1;
        to_add.timestamp = first_ts;
        uint64_t penc[8];
        orc_decommit_query_encode(&to_add, penc);
        full_state_absorb(rq.tail, penc);
        rq.length++;
        if (result_states) memcpy(result_states + 12 * pushes, rq.tail, 96);
        pushes++;
    }
    if (n_result_states) *n_result_states = pushes;
    /* :377-378 enforce_consistency */
    if (uq.length == 0 && memcmp(uq.head, uq.tail, 96)) fail(&st, -1, ZKC_DQ_CHK_QUEUE_CONSISTENCY);
    if (sq.length == 0 && memcmp(sq.head, sq.tail, 96)) fail(&st, -1, ZKC_DQ_CHK_QUEUE_CONSISTENCY);
    /* entry point :183-185 */
    if (completed && (lhs[0] != rhs[0] || lhs[1] != rhs[1])) fail(&st, -1, ZKC_DQ_CHK_GRAND_PRODUCT);

    zkc_decommit_sorter_fsm out;
    memset(&out, 0, sizeof out);
    out.initial_queue_state = uq;
    out.sorted_queue_state = sq;
    out.final_queue_state = rq;
    for (int i = 0; i < 2; i++) { out.lhs_accumulator[i] = lhs[i]; out.rhs_accumulator[i] = rhs[i]; }
    memcpy(out.previous_packed_key, previous_packed_key, sizeof previous_packed_key);
    out.first_encountered_timestamp = first_ts;
    out.previous_record = previous_record;
    zkc_queue_state12 obs_out;
    memset(&obs_out, 0, sizeof obs_out);
    if (completed) obs_out = rq; /* :204-209 */

    if (options && options->compare_expected) {
        uint64_t a[100], b[100], c[25], d25[25];
        orc_decommit_sorter_encode_fsm(&out, a); orc_decommit_sorter_encode_fsm(&io->hidden_fsm_output, b);
        put_queue_state12(c, &obs_out); put_queue_state12(d25, &io->final_queue_state);
        if (memcmp(a, b, sizeof a) || memcmp(c, d25, sizeof c) || (io->completion_flag != 0) != completed)
            if (st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    io->hidden_fsm_output = out;
    io->final_queue_state = obs_out;
    io->completion_flag = (uint32_t)completed;

    uint64_t e_in[50], e_out[25], e_fin[100], e_fout[100];
    size_t n_in = put_queue_state12(e_in, &io->initial_queue_state);
    n_in += put_queue_state12(e_in + n_in, &io->sorted_queue_initial_state);
    const size_t n_out = put_queue_state12(e_out, &obs_out);
    const size_t n_fin = orc_decommit_sorter_encode_fsm(fin, e_fin);
    const size_t n_fout = orc_decommit_sorter_encode_fsm(&out, e_fout);
    orc_closed_form_commitment(start, completed, e_in, n_in, e_out, n_out, e_fin, n_fin, e_fout, n_fout, commitment);
    if (status) *status = st;
    return st.code;
}
