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

// ---- the same permutation spread over 12 lanes ------------------------------------------------------------------------
// The sequential sponges (FSM commitments: 31 dependent permutations over a VM state; queue tail chains) are pure
// latency with one thread per permutation: ~19 k instructions on one lane, ~25 us.  Here lane i of a 16-lane group
// holds s[i] (i < 12; lanes 12..15 carry zeros and only take part in the shuffles): the S-boxes of a full round run
// side by side and the linear layers are a handful of shuffles.  `gm` is the group's lane mask (0xFFFF or 0xFFFF0000);
// the two groups of a warp are independent (different trip counts are fine), the 16 lanes of a group must stay together.
static __device__ const uint64_t P2_RC_LANES[360] = {
#include "poseidon2_rc.inc"
};

__device__ __forceinline__ Acc96 acc96_mul_small(uint64_t x, uint32_t c) {  // c < 2^32 really; used with c <= 7
    const uint64_t lo = (uint64_t)(uint32_t)x * c;
    const uint64_t hi = (uint64_t)(uint32_t)(x >> 32) * c + (lo >> 32);
    return Acc96{(uint32_t)lo, (uint32_t)hi, (uint32_t)(hi >> 32)};
}
__device__ __forceinline__ uint64_t p2c_shfl(unsigned gm, uint64_t v, int src) { return __shfl_sync(gm, v, src, 16); }
__device__ __forceinline__ Acc96 p2c_shfl(unsigned gm, const Acc96 &a, int src) {
    return Acc96{__shfl_sync(gm, a.l0, src, 16), __shfl_sync(gm, a.l1, src, 16), __shfl_sync(gm, a.h, src, 16)};
}
__device__ __forceinline__ Acc96 p2c_shfl_xor(unsigned gm, const Acc96 &a, int m) {
    return Acc96{__shfl_xor_sync(gm, a.l0, m, 16), __shfl_xor_sync(gm, a.l1, m, 16), __shfl_xor_sync(gm, a.h, m, 16)};
}

// circ(2*M4, M4, M4): row (i & 3) of M4 over the 4 lanes of the block, then t_i + (sum of the three blocks' t)
__device__ __forceinline__ uint64_t p2c_external(unsigned gm, uint64_t x, int i) {
    const uint32_t row = (uint32_t)(0x6411753111643175ull >> (16 * (i & 3))) & 0xFFFFu;  // nibble k = M4[i & 3][k]
    Acc96 t = acc96_mul_small(p2c_shfl(gm, x, (i & 12) | 0), row & 15u);
    t.add(acc96_mul_small(p2c_shfl(gm, x, (i & 12) | 1), (row >> 4) & 15u));
    t.add(acc96_mul_small(p2c_shfl(gm, x, (i & 12) | 2), (row >> 8) & 15u));
    t.add(acc96_mul_small(p2c_shfl(gm, x, (i & 12) | 3), row >> 12));
    const int a = i < 8 ? i + 4 : (i < 12 ? i - 8 : i), b = i < 4 ? i + 8 : (i < 12 ? i - 4 : i);
    const Acc96 ta = p2c_shfl(gm, t, a), tb = p2c_shfl(gm, t, b);
    Acc96 o = t.shl(1);
    o.add(ta);
    o.add(tb);
    return i < 12 ? o.reduce_nc() : 0ull;
}

// J + diag(2^shift_i): butterfly sum over the group (lanes 12..15 add zero)
__device__ __forceinline__ uint64_t p2c_inner(unsigned gm, uint64_t x, int i) {
    Acc96 sum = acc96(x);
#pragma unroll
    for (int m = 1; m < 16; m <<= 1) sum.add(p2c_shfl_xor(gm, sum, m));
    const int sh = (int)((0xC36D92508BE4ull >> (4 * i)) & 15);  // nibble i = {4,14,11,8,0,5,2,9,13,6,3,12}[i]
    Acc96 v = acc96(x).shl(sh);
    v.add(sum);
    return i < 12 ? v.reduce_nc() : 0ull;
}

// in: canonical or nc in lanes 0..11 of the group, zero in 12..15; out: canonical
__device__ __forceinline__ uint64_t poseidon2_permute_coop(unsigned gm, uint64_t x, int i) {
    const int ci = i < 12 ? i : 0;
    x = p2c_external(gm, x, i);
#pragma unroll 1
    for (int r = 0; r < 4; r++) x = p2c_external(gm, p2_sbox_nc(gl_add_nc_canon(x, P2_RC_LANES[12 * r + ci])), i);
#pragma unroll 1
    for (int r = 4; r < 26; r++) {
        if (i == 0) x = p2_sbox_nc(gl_add_nc_canon(x, P2_RC[12 * r]));
        x = p2c_inner(gm, x, i);
    }
#pragma unroll 1
    for (int r = 26; r < 30; r++) x = p2c_external(gm, p2_sbox_nc(gl_add_nc_canon(x, P2_RC_LANES[12 * r + ci])), i);
    return i < 12 ? gl_canon(x) : 0ull;
}

// commit_encoding by one 16-lane group: `in` (global or shared, visible to every lane of the group), n the same in all
// 16 lanes; lane j < 4 returns out[j]
__device__ __forceinline__ uint64_t commit_encoding_coop(unsigned gm, const uint64_t *in, int n, int i) {
    uint64_t x = i == 11 ? (uint64_t)n : 0ull;
    for (int off = 0; off < n; off += 8) {
        if (i < 8) x = off + i < n ? in[off + i] : 0ull;
        x = poseidon2_permute_coop(gm, x, i);
    }
    return x;
}

// produce_fs_challenges (/root/reference/src/utils.rs:12-78) by one 16-lane group: `in` holds the n absorbed elements
// (visible to the whole group).  Two repetitions of `per_rep` challenges each, ch[rep * per_rep] = 1 and the rest taken
// from ONE stream of squeezed rate elements, 8 per permutation, that runs on across the repetitions.
__device__ __forceinline__ void fs_challenges_coop(unsigned gm, const uint64_t *in, int n, int per_rep, uint64_t *ch, int i) {
    uint64_t x = commit_encoding_coop(gm, in, n, i);  // the same sponge: length in the capacity, zero-padded 8-chunks
    const int per = per_rep - 1, total = 2 * per;
    for (int t = 0; 8 * t < total; t++) {
        if (t) x = poseidon2_permute_coop(gm, x, i);
        const int idx = 8 * t + i;
        if (i < 8 && idx < total) ch[(idx / per) * per_rep + 1 + idx % per] = x;
    }
    if (i < 2) ch[i * per_rep] = 1;
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
