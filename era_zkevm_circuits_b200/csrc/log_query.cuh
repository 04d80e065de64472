// LogQuery records, their 20-element packing and the 4-wide hash-chained queue
// (CircuitQueue<_, LogQuery, 8, 12, 4, 4, 20, R>) shared by log_sorter and
// storage_validity_by_grand_product.
//   LogQuery::encode   /root/reference/src/base_structures/log_query/mod.rs:121-517
//   flatten order      /root/reference/src/base_structures/log_query/mod.rs:60-101
//   queue absorption   restated in-repo at /root/reference/src/main_vm/opcodes/log.rs:505-600:
//                      empty sponge, enc[0..8], enc[8..16], enc[16..20] || previous tail; tail' = state[0..4]
#pragma once
#include "../../include/zkc_b200.h"
#include "poseidon2.cuh"

namespace zkc {

static_assert(sizeof(zkc_log_query) == 128, "zkc_log_query is a 128-byte record");

__device__ __forceinline__ zkc_log_query lq_load(const zkc_log_query *p) {
    zkc_log_query q;
    const uint4 *s = reinterpret_cast<const uint4 *>(p);
    uint4 *d = reinterpret_cast<uint4 *>(&q);
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = __ldg(s + i);
    return q;
}
__device__ __forceinline__ zkc_log_query lq_zero() {
    zkc_log_query q;
    uint4 *d = reinterpret_cast<uint4 *>(&q);
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = make_uint4(0, 0, 0, 0);
    return q;
}

// byte k of the 52-byte stream key[0..32] || address[0..20] (little-endian limbs)
__device__ __forceinline__ uint32_t lq_stream_byte(const zkc_log_query &q, int k) {
    const uint32_t w = k < 32 ? q.key[k >> 2] : q.address[(k - 32) >> 2];
    return (w >> (8 * (k & 3))) & 0xFFu;
}

__device__ __forceinline__ void lq_encode(const zkc_log_query &q, uint64_t (&e)[20]) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t w = i < 8 ? q.read_value[i] : q.written_value[i - 8];
        const uint32_t top = lq_stream_byte(q, 3 * i) | (lq_stream_byte(q, 3 * i + 1) << 8) | (lq_stream_byte(q, 3 * i + 2) << 16);
        e[i] = pack64(w, top);
    }
    e[16] = pack64(q.timestamp, lq_stream_byte(q, 48) | (lq_stream_byte(q, 49) << 8) | (lq_stream_byte(q, 50) << 16));
    e[17] = pack64(q.tx_number_in_block, lq_stream_byte(q, 51) | (ZKC_LQ_AUX(q.flags) << 8) | (ZKC_LQ_SHARD(q.flags) << 16));
    e[18] = ZKC_LQ_RW(q.flags) + 2 * ZKC_LQ_SERVICE(q.flags);
    e[19] = ZKC_LQ_ROLLBACK(q.flags);
}

// element i of the 36-variable flattening
__device__ __forceinline__ uint64_t lq_flat(const zkc_log_query &q, int i) {
    if (i < 5) return q.address[i];
    if (i < 13) return q.key[i - 5];
    if (i < 21) return q.read_value[i - 13];
    if (i < 29) return q.written_value[i - 21];
    switch (i) {
        case 29: return ZKC_LQ_AUX(q.flags);
        case 30: return ZKC_LQ_RW(q.flags);
        case 31: return ZKC_LQ_ROLLBACK(q.flags);
        case 32: return ZKC_LQ_SERVICE(q.flags);
        case 33: return ZKC_LQ_SHARD(q.flags);
        case 34: return q.tx_number_in_block;
        default: return q.timestamp;
    }
}

// rounds 0 and 1 of the absorption: depend on the encoding only
__device__ __forceinline__ void lq_absorb_head(const uint64_t (&e)[20], uint64_t (&s)[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = i < 8 ? e[i] : 0;
    poseidon2_permute(s);
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = e[8 + i];
    poseidon2_permute(s);
}
// round 2: enc[16..20] || chained 4-element state || carried capacity
__device__ __forceinline__ void lq_absorb_tail(const uint64_t (&e)[20], const uint64_t (&chain)[4], uint64_t (&s)[12]) {
#pragma unroll
    for (int i = 0; i < 4; i++) { s[i] = e[16 + i]; s[4 + i] = chain[i]; }
    poseidon2_permute(s);
}

__device__ inline int put_queue_state4(uint64_t *dst, const zkc_queue_state4 &s) {
    for (int i = 0; i < 4; i++) dst[i] = s.head[i];
    for (int i = 0; i < 4; i++) dst[4 + i] = s.tail[i];
    dst[8] = s.length;
    return 9;
}

// produce_fs_challenges for 4-wide queues, /root/reference/src/utils.rs:12-78: sponge over
// tail || len || tail || len (10 elements), 2 x 20 challenges squeezed 8 at a time, ch[.][0] = 1
__device__ inline void fs_challenges_4(const zkc_queue_state4 &a, const zkc_queue_state4 &b, uint64_t (*ch)[21]) {
    uint64_t in[10];
    for (int i = 0; i < 4; i++) { in[i] = a.tail[i]; in[5 + i] = b.tail[i]; }
    in[4] = a.length; in[9] = b.length;
    uint64_t s[12];
    sponge_init(s, 10);
    for (int off = 0; off < 10; off += 8) {
        for (int j = 0; j < 8; j++) s[j] = off + j < 10 ? in[off + j] : 0;
        poseidon2_permute(s);
    }
    int can_take = 8;
    for (int rep = 0; rep < 2; rep++) {
        ch[rep][0] = 1;
        for (int k = 1; k < 21; k++) {
            if (can_take == 0) { poseidon2_permute(s); can_take = 8; }
            uint64_t v = 0;
            for (int j = 0; j < 8; j++) if (j == 8 - can_take) v = s[j];
            ch[rep][k] = v;
            can_take--;
        }
    }
}

}  // namespace zkc
