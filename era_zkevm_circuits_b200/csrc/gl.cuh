// Goldilocks field arithmetic in registers (p = 2^64 - 2^32 + 1), sm_100a.
// No Montgomery form: the special shape of p turns the 128 -> 64 bit reduction into two
// 32-bit-limb corrections (2^64 = 2^32 - 1, 2^96 = -1 mod p), written as PTX carry chains so
// that nothing is spent on 64-bit compares.  Field type of the reference:
// boojum::field::goldilocks::GoldilocksField (/root/reference/src/ram_permutation/mod.rs:405).
//
// Two value domains:
//   canonical  : < p.  Everything that leaves a kernel (witness cells) is canonical.
//   "nc"       : any uint64_t representing its residue mod p.  The *_nc functions accept and
//                return nc values; gl_canon() brings one back.  Poseidon2 runs entirely in nc.
#pragma once
#include <cstdint>

namespace zkc {

constexpr uint64_t GL_P = 0xFFFFFFFF00000001ull;
constexpr uint64_t GL_EPS = 0xFFFFFFFFull;

__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// z >= p  <=>  z + eps carries out of 64 bits, and then z + eps - 2^64 = z - p
__device__ __forceinline__ uint64_t gl_canon(uint64_t z) {
    uint32_t t0, t1, c;
    asm("add.cc.u32 %0, %3, 0xffffffff;\n\t"
        "addc.cc.u32 %1, %4, 0;\n\t"
        "addc.u32 %2, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(c)
        : "r"((uint32_t)z), "r"((uint32_t)(z >> 32)));
    return c ? pack64(t0, t1) : z;
}

// r0 + r1*2^32 + r2*2^64 + r3*2^96  (mod p), result nc
__device__ __forceinline__ uint64_t gl_reduce_limbs_nc(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    uint32_t z0, z1;
    asm("{\n\t"
        ".reg .u32 m, x0, x1, y0, y1;\n\t"
        "sub.cc.u32 x0, %2, %5;\n\t"   // (r0 + r1*2^32) - r3        [2^96 = -1]
        "subc.cc.u32 x1, %3, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"        // borrow -> m = 0xffffffff = eps
        "sub.cc.u32 x0, x0, m;\n\t"    // borrowed 2^64 = eps: subtract it back
        "subc.u32 x1, x1, 0;\n\t"
        "sub.cc.u32 y0, 0, %4;\n\t"    // r2 * eps = (r2 << 32) - r2  [2^64 = eps]
        "subc.u32 y1, %4, 0;\n\t"
        "add.cc.u32 %0, x0, y0;\n\t"
        "addc.cc.u32 %1, x1, y1;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "sub.u32 m, 0, m;\n\t"         // carry -> eps
        "add.cc.u32 %0, %0, m;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=r"(z0), "=r"(z1)
        : "r"(r0), "r"(r1), "r"(r2), "r"(r3));
    return pack64(z0, z1);
}

// full 64 x 64 -> 128 product as four 32 x 32 + 64 multiply-adds (IMAD.WIDE), then the limb reduction
__device__ __forceinline__ uint64_t gl_mul_nc(uint64_t a, uint64_t b) {
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    const uint64_t p00 = (uint64_t)a0 * b0;
    const uint64_t m1 = (uint64_t)a0 * b1 + (p00 >> 32);
    const uint64_t m2 = (uint64_t)a1 * b0 + (uint32_t)m1;
    const uint64_t hi = (uint64_t)a1 * b1 + (m1 >> 32) + (m2 >> 32);
    return gl_reduce_limbs_nc((uint32_t)p00, (uint32_t)m2, (uint32_t)hi, (uint32_t)(hi >> 32));
}
__device__ __forceinline__ uint64_t gl_sqr_nc(uint64_t a) { return gl_mul_nc(a, a); }

// a (nc) + b (canonical) -> nc: a single conditional "+ eps" absorbs the carry
__device__ __forceinline__ uint64_t gl_add_nc_canon(uint64_t a, uint64_t b) {
    uint32_t z0, z1;
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        "add.cc.u32 %0, %2, %4;\n\t"
        "addc.cc.u32 %1, %3, %5;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "sub.u32 m, 0, m;\n\t"
        "add.cc.u32 %0, %0, m;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=r"(z0), "=r"(z1)
        : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)));
    return pack64(z0, z1);
}

// ---- canonical-domain helpers (witness cells, scans) ----------------------------------------------
__device__ __forceinline__ uint64_t gl_add(uint64_t a, uint64_t b) { return gl_canon(gl_add_nc_canon(a, b)); }
__device__ __forceinline__ uint64_t gl_sub(uint64_t a, uint64_t b) {
    uint64_t d = a - b;
    return a < b ? d + GL_P : d;
}
__device__ __forceinline__ uint64_t gl_neg(uint64_t a) { return a ? GL_P - a : 0; }
__device__ __forceinline__ uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_canon(gl_mul_nc(a, b)); }
__device__ __forceinline__ uint64_t gl_sqr(uint64_t a) { return gl_mul(a, a); }
// a * b + c   (Num::fma with unit coefficients, /root/reference/src/utils.rs:112-128)
__device__ __forceinline__ uint64_t gl_fma(uint64_t a, uint64_t b, uint64_t c) {
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    const uint64_t p00 = (uint64_t)a0 * b0 + (uint32_t)c;          // < 2^64: (2^32-1)^2 + 2^32 - 1
    const uint64_t m1 = (uint64_t)a0 * b1 + (p00 >> 32);
    const uint64_t m2 = (uint64_t)a1 * b0 + (uint32_t)m1;
    const uint64_t m3 = (m2 & GL_EPS) + (c >> 32);                  // limb 1 plus the high half of c
    const uint64_t hi = (uint64_t)a1 * b1 + (m1 >> 32) + (m2 >> 32) + (m3 >> 32);
    return gl_canon(gl_reduce_limbs_nc((uint32_t)p00, (uint32_t)m3, (uint32_t)hi, (uint32_t)(hi >> 32)));
}

// x^(p - 2), p - 2 = 2^64 - 2^32 - 1 = (2^32 - 2) * 2^32 + (2^32 - 1): 63 squarings + 10 multiplications; 0 -> 0
// (the inverse witness of boojum's ZeroCheckGate)
__device__ __forceinline__ uint64_t gl_pow2k_nc(uint64_t x, int k) {
#pragma unroll 1
    for (int i = 0; i < k; i++) x = gl_sqr_nc(x);
    return x;
}
__device__ inline uint64_t gl_inv(uint64_t x) {
    const uint64_t t2 = gl_mul_nc(gl_sqr_nc(x), x);                  // x^(2^2 - 1)
    const uint64_t t4 = gl_mul_nc(gl_pow2k_nc(t2, 2), t2);           // 2^4 - 1
    const uint64_t t8 = gl_mul_nc(gl_pow2k_nc(t4, 4), t4);           // 2^8 - 1
    const uint64_t t16 = gl_mul_nc(gl_pow2k_nc(t8, 8), t8);          // 2^16 - 1
    const uint64_t t24 = gl_mul_nc(gl_pow2k_nc(t16, 8), t8);         // 2^24 - 1
    const uint64_t t28 = gl_mul_nc(gl_pow2k_nc(t24, 4), t4);         // 2^28 - 1
    const uint64_t t30 = gl_mul_nc(gl_pow2k_nc(t28, 2), t2);         // 2^30 - 1
    const uint64_t t31 = gl_mul_nc(gl_sqr_nc(t30), x);               // 2^31 - 1
    const uint64_t t32 = gl_mul_nc(gl_sqr_nc(t31), x);               // 2^32 - 1
    const uint64_t hi = gl_pow2k_nc(gl_sqr_nc(t31), 32);             // (2^32 - 2) * 2^32
    return gl_canon(gl_mul_nc(hi, t32));
}

// 96-bit lazy accumulator for sums of a few dozen u64 terms; reduced once
struct Acc96 {
    uint32_t l0, l1, h;
    __device__ __forceinline__ void add(uint64_t v) {
        asm("add.cc.u32 %0, %0, %3;\n\t"
            "addc.cc.u32 %1, %1, %4;\n\t"
            "addc.u32 %2, %2, 0;"
            : "+r"(l0), "+r"(l1), "+r"(h)
            : "r"((uint32_t)v), "r"((uint32_t)(v >> 32)));
    }
    __device__ __forceinline__ void add(const Acc96 &o) {
        asm("add.cc.u32 %0, %0, %3;\n\t"
            "addc.cc.u32 %1, %1, %4;\n\t"
            "addc.u32 %2, %2, %5;"
            : "+r"(l0), "+r"(l1), "+r"(h)
            : "r"(o.l0), "r"(o.l1), "r"(o.h));
    }
    // this * 2^k, k in [1, 31]
    __device__ __forceinline__ Acc96 shl(int k) const {
        Acc96 r;
        r.h = __funnelshift_l(l1, h, k);
        r.l1 = __funnelshift_l(l0, l1, k);
        r.l0 = l0 << k;
        return r;
    }
    // lo + 2^64 * h with h < 2^32:  2^64 = eps  ->  lo + (h << 32) - h, one carry fix-up; nc result
    __device__ __forceinline__ uint64_t reduce_nc() const {
        uint32_t z0, z1;
        asm("{\n\t"
            ".reg .u32 m, y0, y1;\n\t"
            "sub.cc.u32 y0, 0, %4;\n\t"
            "subc.u32 y1, %4, 0;\n\t"
            "add.cc.u32 %0, %2, y0;\n\t"
            "addc.cc.u32 %1, %3, y1;\n\t"
            "addc.u32 m, 0, 0;\n\t"
            "sub.u32 m, 0, m;\n\t"
            "add.cc.u32 %0, %0, m;\n\t"
            "addc.u32 %1, %1, 0;\n\t"
            "}"
            : "=r"(z0), "=r"(z1)
            : "r"(l0), "r"(l1), "r"(h));
        return pack64(z0, z1);
    }
};
__device__ __forceinline__ Acc96 acc96(uint64_t v) { return Acc96{(uint32_t)v, (uint32_t)(v >> 32), 0u}; }

__device__ __forceinline__ uint64_t shfl_up64(uint64_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }

}  // namespace zkc
