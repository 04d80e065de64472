// Single-pass tile scan with decoupled look-back, generic over an associative (not necessarily
// commutative) operator.  Used for
//   - the running grand products (accumulate_grand_products, /root/reference/src/utils.rs:81-137, is a
//     sequential chain acc <- select(flag, acc * contribution, acc); field multiplication is exact and
//     associative so the chain is an inclusive multiplicative scan, bit-identical in any association),
//   - counters the reference threads through its loops (num_nondeterministic_writes,
//     ram_permutation/mod.rs:281-290; result-queue lengths),
//   - the per-cell state machine of the storage sorter (segmented sums / "last setter" indices,
//     storage_validity_by_grand_product/mod.rs:654-825).
// Warp level: shuffle scans; tile level: shared memory; grid level: per-tile {aggregate, inclusive}
// records published with release/acquire flags; tiles are handed out by an atomic ticket so every
// predecessor of a running tile is already running (forward progress).
#pragma once
#include "gl.cuh"

namespace zkc {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

struct ScanGlobal {
    unsigned int ticket;  // next tile to hand out
    unsigned int pad[3];
};

__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// V: trivially copyable, sizeof(V) % 8 == 0, alignof(V) == 8
template <class V>
struct alignas(16) TileStateT {
    V agg;
    V inc;
    uint32_t flag;  // 0 = nothing, 1 = aggregate published, 2 = inclusive published
    uint32_t pad;
};
template <class V>
struct ScanSharedT {
    V warp_total[SCAN_WARPS];
    V warp_excl[SCAN_WARPS];
    V tile_prefix;
    unsigned int tile;
};

template <class V>
__device__ __forceinline__ V v_shfl_up(const V &v, int d) {
    static_assert(sizeof(V) % 8 == 0, "scan value must be a whole number of 64-bit words");
    V r;
    const uint64_t *s = reinterpret_cast<const uint64_t *>(&v);
    uint64_t *o = reinterpret_cast<uint64_t *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(V) / 8); i++) o[i] = __shfl_up_sync(0xffffffffu, s[i], d);
    return r;
}
template <class V>
__device__ __forceinline__ V v_shfl_xor(const V &v, int d) {
    V r;
    const uint64_t *s = reinterpret_cast<const uint64_t *>(&v);
    uint64_t *o = reinterpret_cast<uint64_t *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(V) / 8); i++) o[i] = __shfl_xor_sync(0xffffffffu, s[i], d);
    return r;
}
template <class V>
__device__ __forceinline__ V v_shfl(const V &v, int src) {
    V r;
    const uint64_t *s = reinterpret_cast<const uint64_t *>(&v);
    uint64_t *o = reinterpret_cast<uint64_t *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(V) / 8); i++) o[i] = __shfl_sync(0xffffffffu, s[i], src);
    return r;
}
template <class V>
__device__ __forceinline__ void v_store_cg(V *dst, const V &v) {
    const uint64_t *s = reinterpret_cast<const uint64_t *>(&v);
    uint64_t *o = reinterpret_cast<uint64_t *>(dst);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(V) / 8); i++) __stcg(o + i, s[i]);
}
template <class V>
__device__ __forceinline__ V v_load_cg(const V *src) {
    V r;
    const uint64_t *s = reinterpret_cast<const uint64_t *>(src);
    uint64_t *o = reinterpret_cast<uint64_t *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(V) / 8); i++) o[i] = __ldcg(s + i);
    return r;
}

template <class V>
__device__ __forceinline__ unsigned int scan_take_ticket(ScanGlobal *g, ScanSharedT<V> &sh) {
    if (threadIdx.x == 0) sh.tile = atomicAdd(&g->ticket, 1u);
    __syncthreads();
    return sh.tile;
}

// Tile-wide scan of one value per thread.  Op: struct with static V identity() and
// static V combine(const V &earlier, const V &later).  Returns the EXCLUSIVE prefix of this thread's
// element over the whole grid order (seeded with `init`, consumed by tile 0); inclusive = exclusive o v.
// All SCAN_THREADS threads of the CTA must call.
template <class V, class Op>
__device__ __forceinline__ V scan_tile_generic(const V &v, unsigned int tile, const V &init, TileStateT<V> *states,
                                               ScanSharedT<V> &sh, V &inclusive) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    V x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        V y = v_shfl_up(x, d);
        if (lane >= d) x = Op::combine(y, x);
    }
    if (lane == 31) sh.warp_total[warp] = x;
    __syncthreads();
    if (warp == 0) {
        V t = lane < SCAN_WARPS ? sh.warp_total[lane] : Op::identity();
        V ti = t;
#pragma unroll
        for (int d = 1; d < SCAN_WARPS; d <<= 1) {
            V y = v_shfl_up(ti, d);
            if (lane >= d) ti = Op::combine(y, ti);
        }
        V te = v_shfl_up(ti, 1);
        if (lane == 0) te = Op::identity();
        if (lane < SCAN_WARPS) sh.warp_excl[lane] = te;
        const V aggregate = v_shfl(ti, SCAN_WARPS - 1);
        TileStateT<V> *me = states + tile;
        V prefix;
        if (tile == 0) {
            prefix = init;
        } else {
            if (lane == 0) {
                v_store_cg(&me->agg, aggregate);
                st_release_u32(&me->flag, 1u);
            }
            prefix = Op::identity();
            int look = (int)tile - 1 - lane;  // lane 0 = nearest predecessor
            while (true) {
                uint32_t f = 2;
                if (look >= 0) {
                    do { f = ld_acquire_u32(&states[look].flag); } while (f == 0);
                }
                const unsigned incl_mask = __ballot_sync(0xffffffffu, f == 2);
                const int first = incl_mask ? __ffs(incl_mask) - 1 : 32;
                V c = Op::identity();
                if (look >= 0 && lane <= first) c = lane == first ? v_load_cg(&states[look].inc) : v_load_cg(&states[look].agg);
                // ordered butterfly: higher lanes hold EARLIER tiles.  Neighbours first (d = 1, 2, 4, ...), so that every
                // step joins two ADJACENT runs of tiles: required for operators that are not commutative
#pragma unroll
                for (int d = 1; d <= 16; d <<= 1) {
                    const V o = v_shfl_xor(c, d);
                    c = (lane & d) ? Op::combine(c, o) : Op::combine(o, c);
                }
                prefix = Op::combine(c, prefix);
                if (incl_mask) break;
                look -= 32;
            }
            // a placeholder lane (look < 0) can only be "first" together with tile 0's real inclusive
            // record at a lower lane, so `init` is always folded in through tile 0
        }
        if (lane == 0) {
            v_store_cg(&me->inc, Op::combine(prefix, aggregate));
            st_release_u32(&me->flag, 2u);
            sh.tile_prefix = prefix;
        }
    }
    __syncthreads();
    V e = v_shfl_up(x, 1);
    if (lane == 0) e = Op::identity();
    const V excl = Op::combine(Op::combine(sh.tile_prefix, sh.warp_excl[warp]), e);
    inclusive = Op::combine(excl, v);
    return excl;
}

// ---- the grand-product element: 4 running products + one u32 counter --------------------------------
struct ScanVal {
    uint64_t p[4];
    uint32_t c;
    uint32_t pad;
};
struct ScanValOp {
    static __device__ __forceinline__ ScanVal identity() { return ScanVal{{1, 1, 1, 1}, 0, 0}; }
    static __device__ __forceinline__ ScanVal combine(const ScanVal &a, const ScanVal &b) {
        ScanVal r;
#pragma unroll
        for (int i = 0; i < 4; i++) r.p[i] = gl_mul(a.p[i], b.p[i]);
        r.c = a.c + b.c;
        r.pad = 0;
        return r;
    }
};
using TileState = TileStateT<ScanVal>;
using ScanShared = ScanSharedT<ScanVal>;
__device__ __forceinline__ ScanVal scan_identity() { return ScanValOp::identity(); }
__device__ __forceinline__ ScanVal scan_tile(const ScanVal &v, unsigned int tile, const ScanVal &init, TileState *states,
                                             ScanShared &sh, ScanVal &inclusive) {
    return scan_tile_generic<ScanVal, ScanValOp>(v, tile, init, states, sh, inclusive);
}

}  // namespace zkc
