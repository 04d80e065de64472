// 256-bit integers on little-endian u32 limbs (the UInt256 of the opcode gadgets, /root/reference/src/main_vm/opcodes/*.rs):
// add / sub with carry, the 512-bit product, Knuth division, shifts.  Shared by the cycle kernels and the gadget-cell kernel.
#pragma once
#include <cstdint>

namespace zkc {

// ---- 256-bit helpers on little-endian u32 limbs -------------------------------------------------------------
struct U256 {
    uint32_t v[8];
};
__device__ __forceinline__ bool u256_is_zero(const U256 &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
__device__ __forceinline__ uint32_t u256_add(const U256 &a, const U256 &b, U256 &c) {
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { const uint64_t t = (uint64_t)a.v[i] + b.v[i] + carry; c.v[i] = (uint32_t)t; carry = t >> 32; }
    return (uint32_t)carry;
}
__device__ __forceinline__ uint32_t u256_sub(const U256 &a, const U256 &b, U256 &c) {
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { const uint64_t t = (uint64_t)a.v[i] - b.v[i] - borrow; c.v[i] = (uint32_t)t; borrow = (t >> 32) & 1; }
    return (uint32_t)borrow;
}
static __device__ void u256_mul(const U256 &a, const U256 &b, U256 &lo, U256 &hi) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint64_t t = (uint64_t)a.v[i] * b.v[j] + r[i + j] + carry;
            r[i + j] = (uint32_t)t; carry = t >> 32;
        }
        r[i + 8] = (uint32_t)carry;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { lo.v[i] = r[i]; hi.v[i] = r[i + 8]; }
}
__device__ __forceinline__ bool u256_ge(const U256 &a, const U256 &b) {
    U256 t;
    return u256_sub(a, b, t) == 0;
}
// q = a / b, r = a % b for b != 0: Knuth's algorithm D on 32-bit limbs (at most 8 quotient digits, each one 64/32
// division + a multiply-subtract), instead of 256 shift-subtract steps that every lane of a warp would wait for
static __device__ void u256_divrem(const U256 &a, const U256 &b, U256 &q, U256 &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) { q.v[i] = 0; r.v[i] = 0; }
    int n = 8;
    while (n > 1 && b.v[n - 1] == 0) n--;
    if (n == 1) {
        uint64_t rem = 0;
        const uint32_t d = b.v[0];
        for (int i = 7; i >= 0; i--) {
            const uint64_t cur = (rem << 32) | a.v[i];
            q.v[i] = (uint32_t)(cur / d);
            rem = cur % d;
        }
        r.v[0] = (uint32_t)rem;
        return;
    }
    const int sh = __clz(b.v[n - 1]);
    uint32_t v[8], u[9];
    for (int i = n - 1; i > 0; i--) v[i] = sh ? (b.v[i] << sh) | (b.v[i - 1] >> (32 - sh)) : b.v[i];
    v[0] = b.v[0] << sh;
    u[8] = sh ? a.v[7] >> (32 - sh) : 0;
    for (int i = 7; i > 0; i--) u[i] = sh ? (a.v[i] << sh) | (a.v[i - 1] >> (32 - sh)) : a.v[i];
    u[0] = a.v[0] << sh;
    for (int j = 8 - n; j >= 0; j--) {
        const uint64_t num = ((uint64_t)u[j + n] << 32) | u[j + n - 1];
        uint64_t qhat = num / v[n - 1], rhat = num % v[n - 1];
        while (qhat >= (1ull << 32) || qhat * v[n - 2] > ((rhat << 32) | u[j + n - 2])) {
            qhat--;
            rhat += v[n - 1];
            if (rhat >= (1ull << 32)) break;
        }
        // u[j .. j+n] -= qhat * v
        int64_t borrow = 0;
        uint64_t carry = 0;
        for (int i = 0; i < n; i++) {
            const uint64_t p = qhat * v[i] + carry;
            carry = p >> 32;
            const int64_t t = (int64_t)u[i + j] - (int64_t)(uint32_t)p + borrow;
            u[i + j] = (uint32_t)t;
            borrow = t >> 32;  // 0 or -1
        }
        const int64_t t = (int64_t)u[j + n] - (int64_t)carry + borrow;
        u[j + n] = (uint32_t)t;
        if (t < 0) {  // qhat was one too large: add the divisor back
            qhat--;
            uint64_t c = 0;
            for (int i = 0; i < n; i++) {
                const uint64_t x = (uint64_t)u[i + j] + v[i] + c;
                u[i + j] = (uint32_t)x;
                c = x >> 32;
            }
            u[j + n] += (uint32_t)c;
        }
        q.v[j] = (uint32_t)qhat;
    }
    for (int i = 0; i < n; i++) r.v[i] = sh ? (u[i] >> sh) | ((uint64_t)u[i + 1] << (32 - sh)) : u[i];
}
// (a << s) mod 2^256 and a >> (256 - s) for s in [0, 255]: the two halves of a * 2^s (shifts.rs:95-96)
static __device__ void u256_shl_wide(const U256 &a, uint32_t s, U256 &lo, U256 &hi) {
    const uint32_t limbs = s >> 5, bits = s & 31;
    uint32_t w[17];
#pragma unroll
    for (int i = 0; i < 17; i++) w[i] = 0;
    for (int i = 0; i < 8; i++) {  // dynamic limb offset: small loop in local memory
        const uint64_t t = (uint64_t)a.v[i] << bits;
        w[i + limbs] |= (uint32_t)t;
        w[i + limbs + 1] |= (uint32_t)(t >> 32);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { lo.v[i] = w[i]; hi.v[i] = w[i + 8]; }
}
static __device__ void u256_shr(const U256 &a, uint32_t s, U256 &q) {
    const uint32_t limbs = s >> 5, bits = s & 31;
    for (int i = 0; i < 8; i++) {
        const uint32_t lo = i + limbs < 8 ? a.v[i + limbs] : 0, hi = i + limbs + 1 < 8 ? a.v[i + limbs + 1] : 0;
        q.v[i] = bits ? (lo >> bits) | (hi << (32 - bits)) : lo;
    }
}

}  // namespace zkc
