/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Sequential CPU restatement of the RAM permutation circuit:
 *   MemoryQuery::encode            /root/reference/src/base_structures/memory_query/mod.rs:103-221
 *   accumulate_grand_products      /root/reference/src/utils.rs:81-137
 *   unpacked_long_comparison       /root/reference/src/storage_validity_by_grand_product/mod.rs:925-944
 *   long_equals                    /root/reference/src/ram_permutation/mod.rs:384-392
 *   partial_accumulate_inner       /root/reference/src/ram_permutation/mod.rs:212-382
 *   ram_permutation_entry_point    /root/reference/src/ram_permutation/mod.rs:31-210
 * One loop iteration at a time, carrying state in scalars exactly like the reference does.
 * Pinning: the loop logic is pinned by the reference's own test vector (mod.rs:559-634: every
 * enforcement holds); accumulator / queue-state / commitment values depend on Poseidon2 and are
 * therefore PARITY UNPINNED (see poseidon2.c).
 */
#include "oracle.h"
#include <string.h>

/* memory_query/mod.rs:103-221 */
void orc_memory_query_encode(const zkc_memory_query *q, uint64_t out[8]) {
    const uint32_t *v = q->value;
    uint8_t d5[4], d6[4], d7[4];
    for (int i = 0; i < 4; i++) {
        d5[i] = (uint8_t)(v[5] >> (8 * i));
        d6[i] = (uint8_t)(v[6] >> (8 * i));
        d7[i] = (uint8_t)(v[7] >> (8 * i));
    }
    out[0] = q->timestamp;
    out[1] = q->memory_page;
    out[2] = (uint64_t)q->index + ((uint64_t)(q->rw_flag & 1) << 32) + ((uint64_t)(q->is_ptr & 1) << 33);
    out[3] = (uint64_t)v[0] + ((uint64_t)d5[0] << 32) + ((uint64_t)d5[1] << 40) + ((uint64_t)d5[2] << 48);
    out[4] = (uint64_t)v[1] + ((uint64_t)d5[3] << 32) + ((uint64_t)d6[0] << 40) + ((uint64_t)d6[1] << 48);
    out[5] = (uint64_t)v[2] + ((uint64_t)d6[2] << 32) + ((uint64_t)d6[3] << 40) + ((uint64_t)d7[0] << 48);
    out[6] = (uint64_t)v[3] + ((uint64_t)d7[1] << 32) + ((uint64_t)d7[2] << 40) + ((uint64_t)d7[3] << 48);
    out[7] = v[4];
}

/* FullStateCircuitQueue::push (un-vendored boojum; the in-repo restatement of the same rule is
 * /root/reference/src/main_vm/utils.rs:194-212): tail' = P(enc[0..8] || tail[8..12]) */
static void full_state_absorb(uint64_t state[12], const uint64_t enc[8]) {
    memcpy(state, enc, 8 * sizeof(uint64_t));
    orc_poseidon2_permutation(state);
}

void orc_memory_queue_simulate(const zkc_memory_query *q, size_t n, uint64_t *prev_states,
                               zkc_queue_state12 *final_state) {
    uint64_t tail[12] = {0}, enc[8];
    for (size_t i = 0; i < n; i++) {
        if (prev_states) memcpy(prev_states + 12 * i, tail, sizeof tail);
        orc_memory_query_encode(&q[i], enc);
        full_state_absorb(tail, enc);
    }
    memset(final_state, 0, sizeof *final_state);
    memcpy(final_state->tail, tail, sizeof tail);
    final_state->length = (uint32_t)n;
}

/* utils.rs:81-137 on column-major inputs */
void orc_accumulate_grand_products(const uint64_t *lhs_enc, const uint64_t *rhs_enc, const uint8_t *should_acc,
                                   size_t enc_len, size_t rows, const uint64_t *challenges,
                                   const uint64_t acc_in[4], uint64_t *acc_out, uint64_t *chain_out,
                                   uint64_t acc_final[4]) {
    uint64_t lhs[2] = {acc_in[0], acc_in[1]}, rhs[2] = {acc_in[2], acc_in[3]};
    for (size_t r = 0; r < rows; r++) {
        for (int rep = 0; rep < 2; rep++) {
            const uint64_t *ch = challenges + rep * (enc_len + 1);
            uint64_t lc = ch[enc_len], rc = ch[enc_len];
            for (size_t i = 0; i < enc_len; i++) {
                lc = gl_fma(lhs_enc[i * rows + r], ch[i], lc);
                rc = gl_fma(rhs_enc[i * rows + r], ch[i], rc);
                if (chain_out) {
                    chain_out[((rep * 2 + 0) * enc_len + i) * rows + r] = lc;
                    chain_out[((rep * 2 + 1) * enc_len + i) * rows + r] = rc;
                }
            }
            uint64_t nl = gl_mul(lhs[rep], lc), nr = gl_mul(rhs[rep], rc);
            int f = should_acc ? should_acc[r] : 1;
            if (f) { lhs[rep] = nl; rhs[rep] = nr; }
        }
        if (acc_out) {
            acc_out[0 * rows + r] = lhs[0]; acc_out[1 * rows + r] = lhs[1];
            acc_out[2 * rows + r] = rhs[0]; acc_out[3 * rows + r] = rhs[1];
        }
    }
    acc_final[0] = lhs[0]; acc_final[1] = lhs[1]; acc_final[2] = rhs[0]; acc_final[3] = rhs[1];
}

/* CSVarLengthEncodable of QueueState<F,12>: head, tail.tail, tail.length */
static size_t put_queue_state12(uint64_t *dst, const zkc_queue_state12 *s) {
    memcpy(dst, s->head, 96);
    memcpy(dst + 12, s->tail, 96);
    dst[24] = s->length;
    return 25;
}
/* RamPermutationInputData, ram_permutation/input.rs:27-31 */
size_t orc_ram_encode_input_data(const zkc_ram_input_data *d, uint64_t *dst) {
    size_t n = put_queue_state12(dst, &d->unsorted_queue_initial_state);
    n += put_queue_state12(dst + n, &d->sorted_queue_initial_state);
    dst[n++] = d->non_deterministic_bootloader_memory_snapshot_length;
    return n; /* 51 */
}
/* RamPermutationFSMInputOutput, ram_permutation/input.rs:52-62 */
size_t orc_ram_encode_fsm(const zkc_ram_fsm *f, uint64_t *dst) {
    size_t n = 0;
    dst[n++] = f->lhs_accumulator[0]; dst[n++] = f->lhs_accumulator[1];
    dst[n++] = f->rhs_accumulator[0]; dst[n++] = f->rhs_accumulator[1];
    n += put_queue_state12(dst + n, &f->current_unsorted_queue_state);
    n += put_queue_state12(dst + n, &f->current_sorted_queue_state);
    for (int i = 0; i < 3; i++) dst[n++] = f->previous_sorting_key[i];
    for (int i = 0; i < 2; i++) dst[n++] = f->previous_full_key[i];
    for (int i = 0; i < 8; i++) dst[n++] = f->previous_value[i];
    dst[n++] = f->previous_is_ptr;
    dst[n++] = f->num_nondeterministic_writes;
    return n; /* 69 */
}

static int all_zero64(const uint64_t *p, int n) {
    for (int i = 0; i < n; i++) if (p[i]) return 0;
    return 1;
}

static void fail(zkc_status *st, int64_t row, uint32_t bit) {
    st->code = ZKC_ERR_UNSATISFIED;
    st->failed_checks |= bit;
    if (row >= 0 && (st->first_bad_row < 0 || row < st->first_bad_row)) st->first_bad_row = row;
}

#define T(col, r) trace[(size_t)(col) * limit + (r)]

int orc_ram_permutation_entry_point(zkc_ram_closed_form *io, const zkc_memory_query *unsorted, size_t n_unsorted,
                                    const zkc_memory_query *sorted, size_t n_sorted, size_t limit,
                                    const zkc_ram_options *options, uint64_t *trace, uint64_t commitment[4],
                                    zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    const uint32_t heap_page = options && options->bootloader_heap_page ? options->bootloader_heap_page
                                                                         : ZKC_BOOTLOADER_HEAP_PAGE_DEFAULT;
    const int start = io->start_flag != 0;
    const zkc_ram_input_data *obs = &io->observable_input;
    const zkc_ram_fsm *fin = &io->hidden_fsm_input;

    /* mod.rs:58-60, 85-87 enforce_trivial_head */
    if (!all_zero64(obs->unsorted_queue_initial_state.head, 12) || !all_zero64(obs->sorted_queue_initial_state.head, 12))
        fail(&st, -1, ZKC_RAM_CHK_TRIVIAL_HEAD);

    zkc_queue_state12 uq = start ? obs->unsorted_queue_initial_state : fin->current_unsorted_queue_state;
    zkc_queue_state12 sq = start ? obs->sorted_queue_initial_state : fin->current_sorted_queue_state;

    /* mod.rs:111-116 */
    uint64_t ch[2][9];
    orc_produce_fs_challenges(obs->unsorted_queue_initial_state.tail, obs->unsorted_queue_initial_state.length,
                              obs->sorted_queue_initial_state.tail, obs->sorted_queue_initial_state.length, 12, 9,
                              &ch[0][0]);

    uint64_t lhs[2], rhs[2];
    for (int i = 0; i < 2; i++) {
        lhs[i] = start ? 1 : fin->lhs_accumulator[i];
        rhs[i] = start ? 1 : fin->rhs_accumulator[i];
    }
    uint32_t nnw = start ? 0 : fin->num_nondeterministic_writes;
    uint32_t prev_sk[3], prev_fk[2], prev_val[8], prev_is_ptr = fin->previous_is_ptr;
    memcpy(prev_sk, fin->previous_sorting_key, sizeof prev_sk);
    memcpy(prev_fk, fin->previous_full_key, sizeof prev_fk);
    memcpy(prev_val, fin->previous_value, sizeof prev_val);

    /* partial_accumulate_inner, mod.rs:212-382 */
    const int not_start = !start;
    if (uq.length != sq.length) fail(&st, -1, ZKC_RAM_CHK_LENGTHS_EQUAL); /* :233-237 */

    size_t upos = 0, spos = 0;
    for (size_t cyc = 0; cyc < limit; cyc++) {
        const int u_empty = uq.length == 0, s_empty = sq.length == 0; /* :247-248 */
        const uint64_t ulen_before = uq.length, slen_before = sq.length;
        if (u_empty != s_empty) fail(&st, (int64_t)cyc, ZKC_RAM_CHK_EMPTY_SYNC); /* :252 */
        const int can_pop = !u_empty; /* :253 */

        /* :256-257 pop_front x2; with a false flag boojum hands out the placeholder (all-zero)
         * witness and leaves head/length untouched */
        zkc_memory_query ui, si;
        uint64_t uenc[8], senc[8];
        memset(&ui, 0, sizeof ui); memset(&si, 0, sizeof si);
        if (can_pop) {
            if (upos < n_unsorted) ui = unsorted[upos++];
            if (spos < n_sorted) si = sorted[spos++];
        }
        orc_memory_query_encode(&ui, uenc);
        orc_memory_query_encode(&si, senc);
        if (can_pop) {
            full_state_absorb(uq.head, uenc); uq.length--;
            full_state_absorb(sq.head, senc); sq.length--;
        }

        /* :260-290 */
        const int ts_is_zero = si.timestamp == 0;
        const int page_is_heap = si.memory_page == heap_page;
        const int is_write = si.rw_flag & 1, is_ptr = si.is_ptr & 1, not_ptr = !is_ptr;
        const int is_nondet = can_pop && ts_is_zero && page_is_heap && is_write && not_ptr;
        if (is_nondet) nnw = nnw + 1;

        /* :296-304 unpacked_long_comparison(a = sorting_key, b = previous): b - a, LSW first */
        const uint32_t sk[3] = {si.timestamp, si.index, si.memory_page};
        const uint32_t fk[2] = {si.index, si.memory_page};
        uint32_t diff[3]; int bor[3], leq[3], borrow = 0, keys_equal = 1;
        for (int i = 0; i < 3; i++) {
            uint64_t d = (uint64_t)prev_sk[i] - sk[i] - (uint64_t)borrow;
            diff[i] = (uint32_t)d;
            borrow = (int)((d >> 32) & 1);
            bor[i] = borrow;
            leq[i] = diff[i] == 0;
            keys_equal &= leq[i];
        }
        const int prev_smaller = borrow;
        /* :308-316 */
        if (cyc != 0) { if (can_pop && !prev_smaller) fail(&st, (int64_t)cyc, ZKC_RAM_CHK_ASCENDING); }
        else { if (can_pop && not_start && !prev_smaller) fail(&st, (int64_t)cyc, ZKC_RAM_CHK_ASCENDING); }

        /* :318-331 */
        const int same_cell = fk[0] == prev_fk[0] && fk[1] == prev_fk[1];
        const int value_equal = memcmp(si.value, prev_val, 32) == 0;
        int value_is_zero = 1;
        for (int i = 0; i < 8; i++) value_is_zero &= si.value[i] == 0;
        const int not_same_cell = !same_cell, not_rw = !is_write;
        const int is_zero = value_is_zero && not_ptr;
        const int ptr_equality = (int)prev_is_ptr == is_ptr;
        const int value_and_ptr_equal = value_equal && ptr_equality;
        int read_uninit, check_equality;
        if (cyc != 0) { /* :334-340 */
            read_uninit = not_same_cell && not_rw;
            check_equality = same_cell && not_rw;
        } else { /* :341-357 */
            const int a = not_start && not_same_cell && not_rw;
            const int b = start && not_rw;
            read_uninit = a || b;
            check_equality = same_cell && not_rw && not_start;
        }
        if (read_uninit && !is_zero) fail(&st, (int64_t)cyc, ZKC_RAM_CHK_UNINIT_READ_ZERO);
        if (check_equality && !value_and_ptr_equal) fail(&st, (int64_t)cyc, ZKC_RAM_CHK_READ_CONSISTENT);

        /* :359-362 */
        uint32_t old_fk[2], old_val[8];
        const uint64_t old_is_ptr = prev_is_ptr;
        memcpy(old_fk, prev_fk, sizeof old_fk); memcpy(old_val, prev_val, sizeof old_val);
        memcpy(prev_sk, sk, sizeof sk); memcpy(prev_fk, fk, sizeof fk);
        memcpy(prev_val, si.value, 32); prev_is_ptr = (uint32_t)is_ptr;

        /* :366-380 accumulate_grand_products */
        uint64_t chain[4][8], gp_new[4];
        for (int rep = 0; rep < 2; rep++) {
            uint64_t lc = ch[rep][8], rc = ch[rep][8];
            for (int i = 0; i < 8; i++) {
                lc = gl_fma(uenc[i], ch[rep][i], lc); chain[rep * 2 + 0][i] = lc;
                rc = gl_fma(senc[i], ch[rep][i], rc); chain[rep * 2 + 1][i] = rc;
            }
            gp_new[rep * 2 + 0] = gl_mul(lhs[rep], lc);
            gp_new[rep * 2 + 1] = gl_mul(rhs[rep], rc);
            if (can_pop) { lhs[rep] = gp_new[rep * 2 + 0]; rhs[rep] = gp_new[rep * 2 + 1]; }
        }

        if (trace) {
            T(ZKC_RAM_UNSORTED_IS_EMPTY, cyc) = (uint64_t)u_empty;
            T(ZKC_RAM_SORTED_IS_EMPTY, cyc) = (uint64_t)s_empty;
            T(ZKC_RAM_CAN_POP, cyc) = (uint64_t)can_pop;
            const zkc_memory_query *items[2] = {&ui, &si};
            const uint64_t *encs[2] = {uenc, senc};
            const zkc_queue_state12 *qs[2] = {&uq, &sq};
            const int base[2] = {ZKC_RAM_UNSORTED_ITEM, ZKC_RAM_SORTED_ITEM};
            for (int k = 0; k < 2; k++) {
                int c = base[k];
                T(c++, cyc) = items[k]->timestamp; T(c++, cyc) = items[k]->memory_page;
                T(c++, cyc) = items[k]->index; T(c++, cyc) = items[k]->rw_flag & 1;
                T(c++, cyc) = items[k]->is_ptr & 1;
                for (int i = 0; i < 8; i++) T(c++, cyc) = items[k]->value[i];
                for (int i = 0; i < 8; i++) T(c++, cyc) = encs[k][i];
                for (int i = 0; i < 12; i++) T(c++, cyc) = qs[k]->head[i];
                T(c++, cyc) = qs[k]->length;
            }
            T(ZKC_RAM_TS_IS_ZERO, cyc) = (uint64_t)ts_is_zero;
            T(ZKC_RAM_PAGE_IS_BOOTLOADER_HEAP, cyc) = (uint64_t)page_is_heap;
            T(ZKC_RAM_IS_NONDET_WRITE, cyc) = (uint64_t)is_nondet;
            T(ZKC_RAM_NUM_NONDET_WRITES, cyc) = nnw;
            for (int i = 0; i < 3; i++) {
                T(ZKC_RAM_CMP_DIFF + i, cyc) = diff[i];
                T(ZKC_RAM_CMP_BORROW + i, cyc) = (uint64_t)bor[i];
                T(ZKC_RAM_CMP_LIMB_EQ + i, cyc) = (uint64_t)leq[i];
            }
            T(ZKC_RAM_KEYS_EQUAL, cyc) = (uint64_t)keys_equal;
            T(ZKC_RAM_PREV_KEY_SMALLER, cyc) = (uint64_t)prev_smaller;
            T(ZKC_RAM_SAME_CELL, cyc) = (uint64_t)same_cell;
            T(ZKC_RAM_VALUE_EQUAL, cyc) = (uint64_t)value_equal;
            T(ZKC_RAM_VALUE_IS_ZERO, cyc) = (uint64_t)value_is_zero;
            T(ZKC_RAM_IS_ZERO, cyc) = (uint64_t)is_zero;
            T(ZKC_RAM_PTR_EQUALITY, cyc) = (uint64_t)ptr_equality;
            T(ZKC_RAM_VALUE_AND_PTR_EQUAL, cyc) = (uint64_t)value_and_ptr_equal;
            T(ZKC_RAM_READ_UNINIT, cyc) = (uint64_t)read_uninit;
            T(ZKC_RAM_CHECK_EQUALITY, cyc) = (uint64_t)check_equality;
            for (int k = 0; k < 4; k++) {
                for (int i = 0; i < 8; i++) T(ZKC_RAM_GP_CHAIN + k * 8 + i, cyc) = chain[k][i];
                T(ZKC_RAM_GP_NEW + k, cyc) = gp_new[k];
            }
            T(ZKC_RAM_GP_ACC + 0, cyc) = lhs[0]; T(ZKC_RAM_GP_ACC + 1, cyc) = rhs[0];
            T(ZKC_RAM_GP_ACC + 2, cyc) = lhs[1]; T(ZKC_RAM_GP_ACC + 3, cyc) = rhs[1];
            /* cells of the gadgets the loop body calls (un-vendored boojum, from the published constructions):
             * decompose_into_bytes (memory_query/mod.rs:133-135), Num::is_zero = ZeroCheckGate (flag, inverse witness),
             * Num::equals = is_zero(a - b), UInt32 / UInt256::equals per limb */
            for (int k = 0; k < 2; k++)
                for (int l = 0; l < 3; l++)
                    for (int b = 0; b < 4; b++)
                        T((k ? ZKC_RAM_SORTED_ENC_BYTES : ZKC_RAM_UNSORTED_ENC_BYTES) + 4 * l + b, cyc) = (items[k]->value[5 + l] >> (8 * b)) & 0xFF;
            T(ZKC_RAM_UNSORTED_LEN_INV, cyc) = orc_gl_inv(ulen_before); T(ZKC_RAM_SORTED_LEN_INV, cyc) = orc_gl_inv(slen_before); /* :247-248 */
            T(ZKC_RAM_TS_INV, cyc) = orc_gl_inv(si.timestamp);                                           /* :261 */
            const uint64_t page_diff = gl_sub(si.memory_page, heap_page);                                /* :263 */
            T(ZKC_RAM_PAGE_DIFF, cyc) = page_diff; T(ZKC_RAM_PAGE_DIFF_INV, cyc) = orc_gl_inv(page_diff);
            for (int i = 0; i < 3; i++) T(ZKC_RAM_CMP_DIFF_INV + i, cyc) = orc_gl_inv(diff[i]);
            for (int i = 0; i < 2; i++) {                                                                /* :318 */
                const uint64_t df = gl_sub(fk[i], old_fk[i]);
                T(ZKC_RAM_CELL_DIFF + i, cyc) = df; T(ZKC_RAM_CELL_DIFF_INV + i, cyc) = orc_gl_inv(df); T(ZKC_RAM_CELL_LIMB_EQ + i, cyc) = df == 0;
            }
            for (int i = 0; i < 8; i++) {                                                                /* :319, :326 */
                const uint64_t df = gl_sub(si.value[i], old_val[i]);
                T(ZKC_RAM_VALUE_DIFF + i, cyc) = df; T(ZKC_RAM_VALUE_DIFF_INV + i, cyc) = orc_gl_inv(df); T(ZKC_RAM_VALUE_LIMB_EQ + i, cyc) = df == 0;
                T(ZKC_RAM_VALUE_ZERO_DIFF + i, cyc) = si.value[i]; T(ZKC_RAM_VALUE_ZERO_DIFF_INV + i, cyc) = orc_gl_inv(si.value[i]);
                T(ZKC_RAM_VALUE_ZERO_LIMB_EQ + i, cyc) = si.value[i] == 0;
            }
            const uint64_t ptr_diff = gl_sub(old_is_ptr, (uint64_t)is_ptr);                              /* :330 */
            T(ZKC_RAM_PTR_DIFF, cyc) = ptr_diff; T(ZKC_RAM_PTR_DIFF_INV, cyc) = orc_gl_inv(ptr_diff);
        }
    }

    /* mod.rs:161-162 enforce_consistency: an empty queue must have head == tail */
    if (uq.length == 0 && memcmp(uq.head, uq.tail, 96) != 0) fail(&st, -1, ZKC_RAM_CHK_QUEUE_CONSISTENCY);
    if (sq.length == 0 && memcmp(sq.head, sq.tail, 96) != 0) fail(&st, -1, ZKC_RAM_CHK_QUEUE_CONSISTENCY);

    const int completed = uq.length == 0; /* :164 */
    if (completed) {
        if (lhs[0] != rhs[0] || lhs[1] != rhs[1]) fail(&st, -1, ZKC_RAM_CHK_GRAND_PRODUCT); /* :166-168 */
        if (nnw != obs->non_deterministic_bootloader_memory_snapshot_length)
            fail(&st, -1, ZKC_RAM_CHK_NONDET_COUNT); /* :170-175 */
    }

    zkc_ram_fsm out;
    memset(&out, 0, sizeof out);
    out.num_nondeterministic_writes = nnw;
    out.current_unsorted_queue_state = uq;
    out.current_sorted_queue_state = sq;
    for (int i = 0; i < 2; i++) { out.lhs_accumulator[i] = lhs[i]; out.rhs_accumulator[i] = rhs[i]; }
    memcpy(out.previous_sorting_key, prev_sk, sizeof prev_sk);
    memcpy(out.previous_full_key, prev_fk, sizeof prev_fk);
    memcpy(out.previous_value, prev_val, sizeof prev_val);
    out.previous_is_ptr = prev_is_ptr;

    /* hook_compare_witness, fsm_input_output/mod.rs:102-133 */
    if (options && options->compare_expected) {
        uint64_t a[69], b[69];
        orc_ram_encode_fsm(&out, a); orc_ram_encode_fsm(&io->hidden_fsm_output, b);
        if (memcmp(a, b, sizeof a) != 0 || (io->completion_flag != 0) != completed) {
            if (st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
        }
    }
    io->hidden_fsm_output = out;
    io->completion_flag = (uint32_t)completed;

    /* mod.rs:200-209 */
    uint64_t e_in[51], e_fin[69], e_fout[69];
    size_t n_in = orc_ram_encode_input_data(obs, e_in);
    size_t n_fin = orc_ram_encode_fsm(fin, e_fin);
    size_t n_fout = orc_ram_encode_fsm(&out, e_fout);
    orc_closed_form_commitment(start, completed, e_in, n_in, NULL, 0, e_fin, n_fin, e_fout, n_fout, commitment);

    if (status) *status = st;
    return st.code;
}
