// Goldilocks field arithmetic in registers (p = 2^64 - 2^32 + 1), sm_100a.
// Values are canonical (< p) at every function boundary.  No Montgomery form: the special shape of
// p turns the 128 -> 64 bit reduction into two 32-bit-limb corrections (2^64 = 2^32 - 1,
// 2^96 = -1 mod p).  Field type of the reference: boojum::field::goldilocks::GoldilocksField
// (/root/reference/src/ram_permutation/mod.rs:405).
#pragma once
#include <cstdint>

namespace zkc {

constexpr uint64_t GL_P = 0xFFFFFFFF00000001ull;
constexpr uint64_t GL_EPS = 0xFFFFFFFFull;

__device__ __forceinline__ uint64_t gl_canon(uint64_t x) { return x >= GL_P ? x - GL_P : x; }

__device__ __forceinline__ uint64_t gl_add(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    // a, b < p: a + b < 2p.  Wrapped past 2^64 or landed in [p, 2^64): subtract p once.
    return (s < a || s >= GL_P) ? s - GL_P : s;
}
__device__ __forceinline__ uint64_t gl_sub(uint64_t a, uint64_t b) {
    uint64_t d = a - b;
    return a < b ? d + GL_P : d;
}
__device__ __forceinline__ uint64_t gl_neg(uint64_t a) { return a ? GL_P - a : 0; }

// x = lo + 2^64 * hi  ->  x mod p, canonical
__device__ __forceinline__ uint64_t gl_reduce128(uint64_t lo, uint64_t hi) {
    const uint64_t hh = hi >> 32, hl = hi & GL_EPS;
    uint64_t t0 = lo - hh;
    if (lo < hh) t0 -= GL_EPS;             // borrowed 2^64 = eps (mod p)
    const uint64_t t1 = (hl << 32) - hl;   // hl * eps
    uint64_t r = t0 + t1;
    if (r < t1) r += GL_EPS;               // carried 2^64 = eps (mod p)
    return gl_canon(r);
}
__device__ __forceinline__ uint64_t gl_mul(uint64_t a, uint64_t b) {
    return gl_reduce128(a * b, __umul64hi(a, b));
}
// a * b + c   (Num::fma with unit coefficients, /root/reference/src/utils.rs:112-128)
__device__ __forceinline__ uint64_t gl_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t lo = a * b, hi = __umul64hi(a, b);
    lo += c;
    hi += (lo < c);
    return gl_reduce128(lo, hi);
}
__device__ __forceinline__ uint64_t gl_sqr(uint64_t a) { return gl_mul(a, a); }

// value = lo + 2^64 * hi with hi < 2^32 (sums of a few small multiples): cheaper reduction
__device__ __forceinline__ uint64_t gl_reduce96(uint64_t lo, uint32_t hi) {
    const uint64_t t1 = ((uint64_t)hi << 32) - hi;
    uint64_t r = lo + t1;
    if (r < t1) r += GL_EPS;
    return gl_canon(r);
}

__device__ __forceinline__ uint64_t shfl_up64(uint64_t v, int d) {
    return __shfl_up_sync(0xffffffffu, v, d);
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}

}  // namespace zkc
