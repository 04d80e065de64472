/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
 * (era_zkevm_circuits_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.
 *
 * Goldilocks field F_p, p = 2^64 - 2^32 + 1, plain C restatement.
 * The reference takes the field from the un-vendored crate `boojum`
 * (boojum::field::goldilocks::GoldilocksField, e.g. /root/reference/src/ram_permutation/mod.rs:405);
 * the arithmetic is pinned by definition (integers mod p) and cross-checked against Python big
 * ints in tests/test_oracle_field.py.  All values are kept canonical (< p).
 */
#ifndef ORC_GL_H
#define ORC_GL_H
#include <stdint.h>

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL

typedef unsigned __int128 u128;

static inline uint64_t gl_canon(uint64_t x) { return x - (GL_P & (0 - (uint64_t)(x >= GL_P))); }

static inline uint64_t gl_add(uint64_t a, uint64_t b) {
    uint64_t s;
    uint64_t c = __builtin_add_overflow(a, b, &s);
    /* a, b < p: wrapped past 2^64 or landed in [p, 2^64) -> subtract p once */
    return s - (GL_P & (0 - (c | (uint64_t)(s >= GL_P))));
}
static inline uint64_t gl_sub(uint64_t a, uint64_t b) { return a - b + (GL_P & (0 - (uint64_t)(a < b))); }
static inline uint64_t gl_neg(uint64_t a) { return a ? GL_P - a : 0; }
/* 2^64 = eps, 2^96 = -1 (mod p): x = lo + eps*hi_lo - hi_hi; branch-free (the carries are data
 * dependent coin flips) */
static inline uint64_t gl_reduce128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & GL_EPS, t0, r;
    uint64_t b = __builtin_sub_overflow(lo, hh, &t0);
    t0 -= GL_EPS & (0 - b);
    uint64_t c = __builtin_add_overflow(t0, hl * GL_EPS, &r);
    r += GL_EPS & (0 - c);
    return gl_canon(r);
}
static inline uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128((u128)a * b); }
/* k0*a*b + k1*c with k0 = k1 = 1: Num::fma as used at /root/reference/src/utils.rs:112-128 */
static inline uint64_t gl_fma(uint64_t a, uint64_t b, uint64_t c) { return gl_reduce128((u128)a * b + c); }
static inline uint64_t gl_pow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; }
    return r;
}
static inline uint64_t gl_inv(uint64_t a) { return gl_pow(a, GL_P - 2); }

#endif
