// Poseidon2 over Goldilocks, width 12 / rate 8 / capacity 4: the round function `R` of every
// reference entry point (CircuitRoundFunction<F, 8, 12, 4>, /root/reference/src/utils.rs:15;
// instantiated as boojum's Poseidon2Goldilocks, /root/reference/src/ram_permutation/mod.rs:411).
// One permutation per thread, the whole state in registers; round constants sit in constant
// memory (warp-uniform index -> broadcast).  The external layer is evaluated with the
// add-chain form of M4 and the inner layer as "sum + 2^s_i * x_i", so the only 64x64
// multiplications are the x^7 S-boxes.
#pragma once
#include "gl.cuh"

namespace zkc {

static __constant__ uint64_t P2_RC[360] = {
#include "poseidon2_rc.inc"
};

__device__ __forceinline__ uint64_t p2_sbox(uint64_t x) {
    const uint64_t x2 = gl_sqr(x), x3 = gl_mul(x2, x), x4 = gl_sqr(x2);
    return gl_mul(x3, x4);
}

// 128-bit lazy accumulator for sums of < 2^32 canonical terms
struct Acc96 {
    uint64_t lo;
    uint32_t hi;
    __device__ __forceinline__ void add(uint64_t v) {
        lo += v;
        hi += (lo < v);
    }
};

// [5 7 1 3; 4 6 1 1; 1 3 5 7; 1 1 4 6] * x, on plain integers (coefficients sum to <= 16)
__device__ __forceinline__ void p2_m4(const uint64_t *x, uint64_t *o) {
    const uint64_t t0 = gl_add(x[0], x[1]), t1 = gl_add(x[2], x[3]);
    const uint64_t t2 = gl_add(gl_add(x[1], x[1]), t1), t3 = gl_add(gl_add(x[3], x[3]), t0);
    uint64_t q = gl_add(t1, t1);
    const uint64_t t4 = gl_add(gl_add(q, q), t3);
    q = gl_add(t0, t0);
    const uint64_t t5 = gl_add(gl_add(q, q), t2);
    o[0] = gl_add(t3, t5);
    o[1] = t5;
    o[2] = gl_add(t2, t4);
    o[3] = t4;
}

// circ(2*M4, M4, M4)
__device__ __forceinline__ void p2_external(uint64_t (&s)[12]) {
    uint64_t t[12];
    p2_m4(s, t);
    p2_m4(s + 4, t + 4);
    p2_m4(s + 8, t + 8);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint64_t sum = gl_add(gl_add(t[i], t[4 + i]), t[8 + i]);
        s[i] = gl_add(t[i], sum);
        s[4 + i] = gl_add(t[4 + i], sum);
        s[8 + i] = gl_add(t[8 + i], sum);
    }
}

// J + diag(2^shift_i)
__device__ __forceinline__ void p2_inner(uint64_t (&s)[12]) {
    constexpr int SH[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};
    Acc96 a{0, 0};
#pragma unroll
    for (int i = 0; i < 12; i++) a.add(s[i]);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        Acc96 b = a;
        if (SH[i] == 0) {
            b.add(s[i]);
        } else {
            const uint64_t lo = s[i] << SH[i];
            b.lo += lo;
            b.hi += (uint32_t)(s[i] >> (64 - SH[i])) + (b.lo < lo);
        }
        s[i] = gl_reduce96(b.lo, b.hi);
    }
}

__device__ __forceinline__ void p2_full_round(uint64_t (&s)[12], int round) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = p2_sbox(gl_add(s[i], P2_RC[12 * round + i]));
    p2_external(s);
}

__device__ __forceinline__ void poseidon2_permute(uint64_t (&s)[12]) {
    p2_external(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) p2_full_round(s, r);
#pragma unroll 1
    for (int r = 4; r < 26; r++) {
        s[0] = p2_sbox(gl_add(s[0], P2_RC[12 * r]));
        p2_inner(s);
    }
#pragma unroll 1
    for (int r = 26; r < 30; r++) p2_full_round(s, r);
}

// R::create_empty_state + R::apply_length_specialization (/root/reference/src/utils.rs:31-33):
// zero state with the encoding length in the last capacity element.
__device__ __forceinline__ void sponge_init(uint64_t (&s)[12], uint64_t length) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    s[11] = length;
}

// commit_encoding, /root/reference/src/fsm_input_output/mod.rs:281-326: absorb-with-replacement of
// zero-padded 8-chunks, first 4 state elements out.  `in` may be global or local memory.
__device__ inline void commit_encoding_dev(const uint64_t *in, int n, uint64_t out[4]) {
    uint64_t s[12];
    sponge_init(s, (uint64_t)n);
    for (int off = 0; off < n; off += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) s[j] = off + j < n ? in[off + j] : 0;
        poseidon2_permute(s);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) out[j] = s[j];
}

}  // namespace zkc
