/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * Sequential CPU restatement of the L2 -> L1 message hasher:
 *   linear_hasher_entry_point                              /root/reference/src/linear_hasher/mod.rs:35-214
 *   LogQuery as ByteSerializable<88> (into_bytes)          /root/reference/src/base_structures/log_query/mod.rs:645-686
 *   keccak256_conditionally_absorb_and_run_permutation     /root/reference/src/storage_application/mod.rs:55-93
 * The byte buffer of the reference is a Rust Vec whose length is a compile-time function of the cycle; it is kept here
 * exactly like that (extend by 88, cut 136 off the front when it holds that many).
 * Pinning: the digest is Keccak-256 of the concatenated serialisations (standard; orc_keccak_f1600 is pinned on the
 * Keccak-256 KATs / hashlib, tests/test_oracle_keccak.py); queue-state values PARITY UNPINNED (Poseidon2, see poseidon2.c).
 * The reference holds no test vector for this circuit.
 */
#include "oracle.h"
#include <string.h>

static void fail(zkc_status *st, int64_t row, uint32_t bit) {
    st->code = ZKC_ERR_UNSATISFIED;
    st->failed_checks |= bit;
    if (row >= 0 && (st->first_bad_row < 0 || row < st->first_bad_row)) st->first_bad_row = row;
}

/* log_query/mod.rs:647-686; returns 0 when the two truncated bytes of tx_number_in_block are not zero (:666-668) */
int orc_log_query_into_bytes(const zkc_log_query *q, uint8_t out[ZKC_LH_MESSAGE_BYTES]) {
    int n = 0;
    out[n++] = (uint8_t)ZKC_LQ_SHARD(q->flags);
    out[n++] = (uint8_t)ZKC_LQ_SERVICE(q->flags);
    out[n++] = (uint8_t)(q->tx_number_in_block >> 8);
    out[n++] = (uint8_t)q->tx_number_in_block;
    for (int l = 4; l >= 0; l--)
        for (int b = 3; b >= 0; b--) out[n++] = (uint8_t)(q->address[l] >> (8 * b));
    for (int l = 7; l >= 0; l--)
        for (int b = 3; b >= 0; b--) out[n++] = (uint8_t)(q->key[l] >> (8 * b));
    for (int l = 7; l >= 0; l--)
        for (int b = 3; b >= 0; b--) out[n++] = (uint8_t)(q->written_value[l] >> (8 * b));
    return (q->tx_number_in_block >> 16) == 0;
}

/* storage_application/mod.rs:55-93 on lanes A[x + 5y] (state[i][j] of the reference is lane i + 5j) */
static void conditionally_absorb(int condition, uint64_t A[25], const uint8_t block[ZKC_KECCAK_RATE_BYTES]) {
    uint64_t N[25];
    memcpy(N, A, sizeof N);
    for (int idx = 0; idx < ZKC_KECCAK_RATE_BYTES / 8; idx++) {
        uint64_t w = 0;
        for (int b = 0; b < 8; b++) w |= (uint64_t)block[8 * idx + b] << (8 * b);
        N[idx] ^= w;
    }
    orc_keccak_f1600(N);
    if (condition) memcpy(A, N, sizeof N);
}

#define T(col, r) trace[(size_t)(col) * limit + (r)]

/* keccak_states (optional out): [limit][25], the keccak state after every cycle */
int orc_linear_hasher_entry_point(zkc_linear_hasher_closed_form *io, const zkc_log_query *records, size_t n_records, size_t limit,
                                  const zkc_sorter_options *options, uint64_t *trace, uint64_t *keccak_states, uint64_t commitment[4],
                                  zkc_status *status) {
    zkc_status st = {ZKC_OK, 0, -1, 0, 0};
    static const uint64_t zero4[4] = {0, 0, 0, 0};
    if (io->start_flag == 0) fail(&st, -1, ZKC_LH_CHK_START_FLAG);                                 /* :66 */
    if (memcmp(io->queue_state.head, zero4, 32)) fail(&st, -1, ZKC_LH_CHK_TRIVIAL_HEAD);           /* :71 */
    zkc_queue_state4 q = io->queue_state;
    uint64_t A[25];
    memset(A, 0, sizeof A);
    uint8_t buffer[2 * ZKC_KECCAK_RATE_BYTES];
    size_t buffer_len = 0;
    int done = q.length == 0;                                                                      /* :99 */
    const int no_work = done;
    size_t pos = 0;
    for (size_t cyc = 0; cyc < limit; cyc++) {
        const int queue_is_empty = q.length == 0, should_pop = !queue_is_empty;
        zkc_log_query it;
        memset(&it, 0, sizeof it);
        if (should_pop && pos < n_records) it = records[pos++];
        uint64_t enc[20];
        orc_log_query_encode(&it, enc);
        if (should_pop) { orc_log_queue_absorb(q.head, enc, NULL); q.length--; }
        const int now_empty = q.length == 0;
        const int is_last_serialization = should_pop && now_empty;
        uint8_t bytes[ZKC_LH_MESSAGE_BYTES];
        if (!orc_log_query_into_bytes(&it, bytes)) fail(&st, (int64_t)cyc, ZKC_LH_CHK_TX_NUMBER_RANGE);
        memcpy(buffer + buffer_len, bytes, ZKC_LH_MESSAGE_BYTES);
        buffer_len += ZKC_LH_MESSAGE_BYTES;
        const int continue_to_absorb = !done;
        int absorb_full = 0;
        if (buffer_len >= ZKC_KECCAK_RATE_BYTES) {                                                 /* :120-137 */
            absorb_full = continue_to_absorb;
            conditionally_absorb(continue_to_absorb, A, buffer);
            memmove(buffer, buffer + ZKC_KECCAK_RATE_BYTES, buffer_len - ZKC_KECCAK_RATE_BYTES);
            buffer_len -= ZKC_KECCAK_RATE_BYTES;
        }
        uint64_t mid[25];
        memcpy(mid, A, sizeof mid);
        const int absorb_last = continue_to_absorb && is_last_serialization;                       /* :144-145 */
        {
            uint8_t last[ZKC_KECCAK_RATE_BYTES];
            memset(last, 0, sizeof last);
            memcpy(last, buffer, buffer_len);
            if (buffer_len == ZKC_KECCAK_RATE_BYTES - 1) last[buffer_len] = 0x81;
            else { last[buffer_len] = 0x01; last[ZKC_KECCAK_RATE_BYTES - 1] = 0x80; }
            conditionally_absorb(absorb_last, A, last);
        }
        done = done || is_last_serialization;                                                      /* :170 */
        if (keccak_states) memcpy(keccak_states + 25 * cyc, A, sizeof A);
        if (trace) {
            T(ZKC_LH_QUEUE_IS_EMPTY, cyc) = (uint64_t)queue_is_empty; T(ZKC_LH_SHOULD_POP, cyc) = (uint64_t)should_pop;
            uint64_t flat[36];
            orc_log_query_flatten(&it, flat);
            for (int i = 0; i < 36; i++) T(ZKC_LH_ITEM + i, cyc) = flat[i];
            for (int i = 0; i < 20; i++) T(ZKC_LH_ENC + i, cyc) = enc[i];
            for (int i = 0; i < 4; i++) T(ZKC_LH_HEAD + i, cyc) = q.head[i];
            T(ZKC_LH_LEN, cyc) = q.length;
            T(ZKC_LH_NOW_EMPTY, cyc) = (uint64_t)now_empty; T(ZKC_LH_IS_LAST_SERIALIZATION, cyc) = (uint64_t)is_last_serialization;
            for (int i = 0; i < ZKC_LH_MESSAGE_BYTES; i++) T(ZKC_LH_BYTES + i, cyc) = bytes[i];
            T(ZKC_LH_CONTINUE_TO_ABSORB, cyc) = (uint64_t)continue_to_absorb;
            T(ZKC_LH_ABSORB_FULL, cyc) = (uint64_t)absorb_full; T(ZKC_LH_ABSORB_LAST, cyc) = (uint64_t)absorb_last;
            for (int i = 0; i < 25; i++) {
                T(ZKC_LH_STATE_MID + 2 * i, cyc) = (uint32_t)mid[i]; T(ZKC_LH_STATE_MID + 2 * i + 1, cyc) = mid[i] >> 32;
                T(ZKC_LH_STATE_OUT + 2 * i, cyc) = (uint32_t)A[i]; T(ZKC_LH_STATE_OUT + 2 * i + 1, cyc) = A[i] >> 32;
            }
            T(ZKC_LH_DONE, cyc) = (uint64_t)done;
        }
    }
    if (q.length == 0 && memcmp(q.head, q.tail, 32)) fail(&st, -1, ZKC_LH_CHK_QUEUE_CONSISTENCY);  /* :173 */
    const int completed = q.length == 0;
    if (!completed) fail(&st, -1, ZKC_LH_CHK_NOT_COMPLETED);                                       /* :176 */
    uint8_t digest[32];
    if (no_work) orc_keccak256((const uint8_t *)"", 0, digest);                                    /* :87-96, :195-196 */
    else
        for (int i = 0; i < 4; i++)
            for (int b = 0; b < 8; b++) digest[8 * i + b] = (uint8_t)(A[i] >> (8 * b));
    uint64_t e_in[9], e_out[32];
    const size_t n_in = orc_put_queue_state4(e_in, &io->queue_state);
    for (int i = 0; i < 32; i++) e_out[i] = digest[i];
    if (options && options->compare_expected) {
        int same = (io->completion_flag != 0) == completed;
        for (int i = 0; i < 32; i++) same &= io->keccak256_hash[i] == digest[i];
        if (!same && st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    for (int i = 0; i < 32; i++) io->keccak256_hash[i] = digest[i];
    io->completion_flag = (uint32_t)completed;
    /* the hidden FSM input / output are `()`: empty encodings */
    orc_closed_form_commitment(io->start_flag != 0, completed, e_in, n_in, e_out, 32, NULL, 0, NULL, 0, commitment);
    if (status) *status = st;
    return st.code;
}
