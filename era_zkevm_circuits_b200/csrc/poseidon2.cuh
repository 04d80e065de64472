// Poseidon2 over Goldilocks, width 12 / rate 8 / capacity 4: the round function `R` of every
// reference entry point (CircuitRoundFunction<F, 8, 12, 4>, /root/reference/src/utils.rs:15;
// instantiated as boojum's Poseidon2Goldilocks, /root/reference/src/ram_permutation/mod.rs:411).
// One permutation per thread, the whole state in registers, values in the non-canonical domain of
// gl.cuh until the very end.  Round constants sit in constant memory (warp-uniform index ->
// broadcast).  The linear layers are evaluated on 96-bit lazy accumulators (M4 as its add chain,
// the inner layer as "sum + 2^s_i * x_i") and reduced once per lane, so the only 64 x 64
// multiplications are the x^7 S-boxes.
#pragma once
#include "gl.cuh"

namespace zkc {

static __constant__ uint64_t P2_RC[360] = {
#include "poseidon2_rc.inc"
};

__device__ __forceinline__ uint64_t p2_sbox_nc(uint64_t x) {
    const uint64_t x2 = gl_sqr_nc(x), x3 = gl_mul_nc(x2, x), x4 = gl_sqr_nc(x2);
    return gl_mul_nc(x3, x4);
}

// [5 7 1 3; 4 6 1 1; 1 3 5 7; 1 1 4 6] * x on integers (row sums 16 -> < 2^68)
__device__ __forceinline__ void p2_m4(const uint64_t *x, Acc96 *o) {
    Acc96 t0 = acc96(x[0]); t0.add(x[1]);
    Acc96 t1 = acc96(x[2]); t1.add(x[3]);
    Acc96 t2 = acc96(x[1]).shl(1); t2.add(t1);
    Acc96 t3 = acc96(x[3]).shl(1); t3.add(t0);
    Acc96 t4 = t1.shl(2); t4.add(t3);
    Acc96 t5 = t0.shl(2); t5.add(t2);
    o[0] = t3; o[0].add(t5);
    o[1] = t5;
    o[2] = t2; o[2].add(t4);
    o[3] = t4;
}

// circ(2*M4, M4, M4); row sums 64 -> < 2^70
__device__ __forceinline__ void p2_external(uint64_t (&s)[12]) {
    Acc96 t[12];
    p2_m4(s, t);
    p2_m4(s + 4, t + 4);
    p2_m4(s + 8, t + 8);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        Acc96 sum = t[i]; sum.add(t[4 + i]); sum.add(t[8 + i]);
#pragma unroll
        for (int b = 0; b < 3; b++) {
            Acc96 v = t[4 * b + i]; v.add(sum);
            s[4 * b + i] = v.reduce_nc();
        }
    }
}

// J + diag(2^shift_i)
__device__ __forceinline__ void p2_inner(uint64_t (&s)[12]) {
    constexpr int SH[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};
    Acc96 sum = acc96(s[0]);
#pragma unroll
    for (int i = 1; i < 12; i++) sum.add(s[i]);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        Acc96 v = SH[i] ? acc96(s[i]).shl(SH[i]) : acc96(s[i]);
        v.add(sum);
        s[i] = v.reduce_nc();
    }
}

__device__ __forceinline__ void p2_full_round(uint64_t (&s)[12], int round) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = p2_sbox_nc(gl_add_nc_canon(s[i], P2_RC[12 * round + i]));
    p2_external(s);
}

// in: canonical or nc; out: canonical
__device__ __forceinline__ void poseidon2_permute(uint64_t (&s)[12]) {
    p2_external(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) p2_full_round(s, r);
#pragma unroll 1
    for (int r = 4; r < 26; r++) {
        s[0] = p2_sbox_nc(gl_add_nc_canon(s[0], P2_RC[12 * r]));
        p2_inner(s);
    }
#pragma unroll 1
    for (int r = 26; r < 30; r++) p2_full_round(s, r);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
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

// the same as one out-of-line copy, for kernels that commit at several places (finalize kernels): one instance of the
// permutation in the kernel keeps the register allocation of the rest of it out of the spill range
static __device__ __noinline__ void commit_encoding_call(const uint64_t *in, int n, uint64_t *out) {
    uint64_t o[4];
    commit_encoding_dev(in, n, o);
#pragma unroll
    for (int j = 0; j < 4; j++) out[j] = o[j];
}

}  // namespace zkc
