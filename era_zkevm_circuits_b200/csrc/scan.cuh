// Single-pass tile scan with decoupled look-back for the running grand products
// (accumulate_grand_products, /root/reference/src/utils.rs:81-137, is a sequential chain
// acc <- select(flag, acc * contribution, acc); field multiplication is exact and associative so
// the chain is an inclusive multiplicative scan, bit-identical in any association order).
// One scan element carries the 4 products (2 repetitions x lhs/rhs) and one u32 counter
// (e.g. num_nondeterministic_writes, ram_permutation/mod.rs:281-290).
// Warp level: shuffle scans; tile level: shared memory; grid level: per-tile {aggregate,
// inclusive} records published with release/acquire flags; tiles are handed out by an atomic
// ticket so every predecessor of a running tile is already running (forward progress).
#pragma once
#include "gl.cuh"

namespace zkc {

struct ScanVal {
    uint64_t p[4];
    uint32_t c;
};

__device__ __forceinline__ ScanVal scan_identity() { return ScanVal{{1, 1, 1, 1}, 0}; }
__device__ __forceinline__ ScanVal scan_combine(const ScanVal &a, const ScanVal &b) {
    ScanVal r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.p[i] = gl_mul(a.p[i], b.p[i]);
    r.c = a.c + b.c;
    return r;
}
__device__ __forceinline__ ScanVal scan_shfl_up(const ScanVal &v, int d) {
    ScanVal r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.p[i] = shfl_up64(v.p[i], d);
    r.c = __shfl_up_sync(0xffffffffu, v.c, d);
    return r;
}
__device__ __forceinline__ ScanVal scan_shfl_xor(const ScanVal &v, int d) {
    ScanVal r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.p[i] = __shfl_xor_sync(0xffffffffu, v.p[i], d);
    r.c = __shfl_xor_sync(0xffffffffu, v.c, d);
    return r;
}
__device__ __forceinline__ ScanVal scan_shfl(const ScanVal &v, int src) {
    ScanVal r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.p[i] = shfl64(v.p[i], src);
    r.c = __shfl_sync(0xffffffffu, v.c, src);
    return r;
}

struct alignas(16) TileState {
    uint64_t agg[4];
    uint64_t inc[4];
    uint32_t agg_c, inc_c;
    uint32_t flag;  // 0 = nothing, 1 = aggregate published, 2 = inclusive published
    uint32_t pad;
};

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

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

struct ScanShared {
    ScanVal warp_total[SCAN_WARPS];
    ScanVal warp_excl[SCAN_WARPS];
    ScanVal tile_prefix;
    unsigned int tile;
};

__device__ __forceinline__ unsigned int scan_take_ticket(ScanGlobal *g, ScanShared &sh) {
    if (threadIdx.x == 0) sh.tile = atomicAdd(&g->ticket, 1u);
    __syncthreads();
    return sh.tile;
}

// Tile-wide scan of one value per thread.  Returns the EXCLUSIVE prefix of this thread's element
// over the whole grid order (seeded with `init`, used by tile 0); `inclusive` = exclusive o v.
// All SCAN_THREADS threads of the CTA must call.
__device__ __forceinline__ ScanVal scan_tile(const ScanVal &v, unsigned int tile, const ScanVal &init,
                                             TileState *states, ScanShared &sh, ScanVal &inclusive) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // warp inclusive scan
    ScanVal x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        ScanVal y = scan_shfl_up(x, d);
        if (lane >= d) x = scan_combine(y, x);
    }
    if (lane == 31) sh.warp_total[warp] = x;
    __syncthreads();
    if (warp == 0) {
        // scan the warp totals, publish the tile aggregate, look back
        ScanVal t = lane < SCAN_WARPS ? sh.warp_total[lane] : scan_identity();
        ScanVal ti = t;
#pragma unroll
        for (int d = 1; d < SCAN_WARPS; d <<= 1) {
            ScanVal y = scan_shfl_up(ti, d);
            if (lane >= d) ti = scan_combine(y, ti);
        }
        ScanVal te = scan_shfl_up(ti, 1);
        if (lane == 0) te = scan_identity();
        if (lane < SCAN_WARPS) sh.warp_excl[lane] = te;
        const ScanVal aggregate = scan_shfl(ti, SCAN_WARPS - 1);
        TileState *me = states + tile;
        ScanVal prefix;
        if (tile == 0) {
            prefix = init;
        } else {
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 4; i++) __stcg(&me->agg[i], aggregate.p[i]);
                __stcg(&me->agg_c, aggregate.c);
                st_release_u32(&me->flag, 1u);
            }
            prefix = scan_identity();
            int look = (int)tile - 1 - lane;
            while (true) {
                uint32_t f = 2;
                if (look >= 0) {
                    do { f = ld_acquire_u32(&states[look].flag); } while (f == 0);
                }
                const unsigned incl_mask = __ballot_sync(0xffffffffu, f == 2);
                const int first = incl_mask ? __ffs(incl_mask) - 1 : 32;
                ScanVal c = scan_identity();
                if (look >= 0 && lane <= first) {
                    const TileState *s = states + look;
                    if (lane == first) {
#pragma unroll
                        for (int i = 0; i < 4; i++) c.p[i] = __ldcg(&s->inc[i]);
                        c.c = __ldcg(&s->inc_c);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i++) c.p[i] = __ldcg(&s->agg[i]);
                        c.c = __ldcg(&s->agg_c);
                    }
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) c = scan_combine(c, scan_shfl_xor(c, d));
                prefix = scan_combine(c, prefix);
                if (incl_mask) break;
                look -= 32;
            }
            // an identity placeholder lane (look < 0) can only be "first" together with tile 0's
            // real inclusive record at a lower lane, so `init` is always folded in via tile 0
        }
        if (lane == 0) {
            const ScanVal inc = scan_combine(prefix, aggregate);
#pragma unroll
            for (int i = 0; i < 4; i++) __stcg(&me->inc[i], inc.p[i]);
            __stcg(&me->inc_c, inc.c);
            st_release_u32(&me->flag, 2u);
            sh.tile_prefix = prefix;
        }
    }
    __syncthreads();
    ScanVal e = scan_shfl_up(x, 1);
    if (lane == 0) e = scan_identity();
    ScanVal excl = scan_combine(scan_combine(sh.tile_prefix, sh.warp_excl[warp]), e);
    inclusive = scan_combine(excl, v);
    return excl;
}

}  // namespace zkc
