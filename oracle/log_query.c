/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * LogQuery encoding and the 4-wide hash-chained queue the log circuits use:
 *   LogQuery::encode                       /root/reference/src/base_structures/log_query/mod.rs:121-517
 *   flatten order                          /root/reference/src/base_structures/log_query/mod.rs:60-101
 *   append_timestamp_to_raw_query_encoding /root/reference/src/storage_validity_by_grand_product/mod.rs:72-96
 *   CircuitQueue push / pop_front          un-vendored boojum; the same absorption schedule is restated
 *                                          in-repo at /root/reference/src/main_vm/opcodes/log.rs:505-600
 *                                          (empty state, no length specialisation; enc[0..8], enc[8..16],
 *                                          enc[16..20] || previous 4-element tail; new tail = state[0..4])
 */
#include "oracle.h"
#include <string.h>

void orc_log_query_encode(const zkc_log_query *q, uint64_t out[20]) {
    uint8_t b[52];
    for (int l = 0; l < 8; l++)
        for (int j = 0; j < 4; j++) b[4 * l + j] = (uint8_t)(q->key[l] >> (8 * j));
    for (int l = 0; l < 5; l++)
        for (int j = 0; j < 4; j++) b[32 + 4 * l + j] = (uint8_t)(q->address[l] >> (8 * j));
    for (int i = 0; i < 16; i++) {
        const uint64_t w = i < 8 ? q->read_value[i] : q->written_value[i - 8];
        out[i] = w + ((uint64_t)b[3 * i] << 32) + ((uint64_t)b[3 * i + 1] << 40) + ((uint64_t)b[3 * i + 2] << 48);
    }
    out[16] = (uint64_t)q->timestamp + ((uint64_t)b[48] << 32) + ((uint64_t)b[49] << 40) + ((uint64_t)b[50] << 48);
    out[17] = (uint64_t)q->tx_number_in_block + ((uint64_t)b[51] << 32) + ((uint64_t)ZKC_LQ_AUX(q->flags) << 40) +
              ((uint64_t)ZKC_LQ_SHARD(q->flags) << 48);
    out[18] = (uint64_t)ZKC_LQ_RW(q->flags) + 2 * (uint64_t)ZKC_LQ_SERVICE(q->flags);
    out[19] = ZKC_LQ_ROLLBACK(q->flags);
}

void orc_log_query_flatten(const zkc_log_query *q, uint64_t out[36]) {
    int n = 0;
    for (int i = 0; i < 5; i++) out[n++] = q->address[i];
    for (int i = 0; i < 8; i++) out[n++] = q->key[i];
    for (int i = 0; i < 8; i++) out[n++] = q->read_value[i];
    for (int i = 0; i < 8; i++) out[n++] = q->written_value[i];
    out[n++] = ZKC_LQ_AUX(q->flags);
    out[n++] = ZKC_LQ_RW(q->flags);
    out[n++] = ZKC_LQ_ROLLBACK(q->flags);
    out[n++] = ZKC_LQ_SERVICE(q->flags);
    out[n++] = ZKC_LQ_SHARD(q->flags);
    out[n++] = q->tx_number_in_block;
    out[n++] = q->timestamp;
}

/* one absorption of a 20-element encoding chained on a 4-element state; rounds (optional): the three
 * full sponge states */
void orc_log_queue_absorb(uint64_t chain[4], const uint64_t enc[20], uint64_t *rounds /* [3][12] or NULL */) {
    uint64_t s[12];
    memset(s, 0, sizeof s);
    memcpy(s, enc, 64);
    orc_poseidon2_permutation(s);
    if (rounds) memcpy(rounds, s, 96);
    memcpy(s, enc + 8, 64);
    orc_poseidon2_permutation(s);
    if (rounds) memcpy(rounds + 12, s, 96);
    memcpy(s, enc + 16, 32);
    memcpy(s + 4, chain, 32);
    orc_poseidon2_permutation(s);
    if (rounds) memcpy(rounds + 24, s, 96);
    memcpy(chain, s, 32);
}

void orc_log_queue_simulate(const zkc_log_query *q, const uint32_t *extra_ts, size_t n, uint64_t *prev_tails,
                            zkc_queue_state4 *final_state) {
    uint64_t tail[4] = {0}, enc[20];
    for (size_t i = 0; i < n; i++) {
        if (prev_tails) memcpy(prev_tails + 4 * i, tail, 32);
        orc_log_query_encode(&q[i], enc);
        if (extra_ts) enc[19] += (uint64_t)extra_ts[i] << 8;
        orc_log_queue_absorb(tail, enc, NULL);
    }
    memset(final_state, 0, sizeof *final_state);
    memcpy(final_state->tail, tail, 32);
    final_state->length = (uint32_t)n;
}

size_t orc_put_queue_state4(uint64_t *dst, const zkc_queue_state4 *s) {
    memcpy(dst, s->head, 32);
    memcpy(dst + 4, s->tail, 32);
    dst[8] = s->length;
    return 9;
}
