// The RESULT queue of the sorter circuits (log_sorter/mod.rs:395, storage_validity_by_grand_product/
// mod.rs:705): a conditional CircuitQueue::push per loop iteration.  Rounds 0-1 of a push depend on
// the pushed item only and are computed by the row kernels, which leave per row
//   r2in[row][8] = enc[16..20] || capacity after round 1,   meta[row] = (pushes before this row) << 1 | pushed
// Round 2 consumes the previous tail: either verified against host-supplied tails or rebuilt by the
// sequential chain kernel (1 permutation per executed push).
// `Dev` is the per-call device block of the circuit (members rq0, limit, hint_bad, failed_checks, first_bad).
#pragma once
#include "poseidon2.cuh"

namespace zkc {

// The chain when the caller supplies no tails: one permutation per push, each consuming the previous tail.  The chain itself
// cannot be cut, but ONE permutation can: it runs on 12 cooperating lanes (poseidon2_permute_coop: S-boxes of a full round side
// by side, linear layers as shuffles), ~5x shorter than on one lane.  Launch with one warp; lanes 0..15 work.
template <class Dev>
__global__ void rq_chain_kernel(const Dev *d, const uint64_t *__restrict__ r2in, const uint32_t *__restrict__ meta,
                                uint64_t *__restrict__ tails) {
    const int i = threadIdx.x;
    if (blockIdx.x != 0 || i >= 16) return;
    const unsigned gm = 0xFFFFu;
    uint64_t tail = i < 4 ? d->rq0.tail[i] : 0ull;  // lanes 0..3 hold the running tail
    const size_t limit = d->limit;
    size_t k = 0;
    for (size_t row = 0; row < limit; row++) {
        if (!(meta[row] & 1u)) continue;  // uniform over the group
        const uint64_t from_tail = __shfl_sync(gm, tail, (i - 4) & 3, 16);
        uint64_t x = 0;
        if (i < 4) x = r2in[8 * row + i];
        else if (i < 8) x = from_tail;
        else if (i < 12) x = r2in[8 * row + 4 + (i - 8)];
        x = poseidon2_permute_coop(gm, x, i);
        if (i < 4) { tail = x; tails[4 * k + i] = x; }
        k++;
    }
}

template <class Dev, int COL_ROUND2, int COL_TAIL, uint32_t HINT_BIT>
__global__ void __launch_bounds__(256)
rq_push_kernel(Dev *d, const uint64_t *__restrict__ r2in, const uint32_t *__restrict__ meta,
               const uint64_t *__restrict__ tails, size_t n_tails, uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const uint32_t m = meta[row];
    const size_t k = m >> 1;
    const bool add = m & 1u;
    uint64_t before[4], s[12];
    bool ok = true;
    if (k == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) before[i] = d->rq0.tail[i];
    } else if (k - 1 < n_tails) {
#pragma unroll
        for (int i = 0; i < 4; i++) before[i] = __ldg(tails + 4 * (k - 1) + i);
    } else {
        ok = false;
#pragma unroll
        for (int i = 0; i < 4; i++) before[i] = 0;
    }
    const ulonglong2 *in = reinterpret_cast<const ulonglong2 *>(r2in + 8 * row);
    const ulonglong2 a = in[0], b = in[1], c = in[2], e = in[3];
    s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y;
#pragma unroll
    for (int i = 0; i < 4; i++) s[4 + i] = before[i];
    s[8] = c.x; s[9] = c.y; s[10] = e.x; s[11] = e.y;
    poseidon2_permute(s);
    if (add) {
        if (k < n_tails) {
#pragma unroll
            for (int i = 0; i < 4; i++) ok &= __ldg(tails + 4 * k + i) == s[i];
        } else ok = false;
    }
    if (trace) {
#pragma unroll
        for (int i = 0; i < 12; i++) trace[(size_t)(COL_ROUND2 + i) * limit + row] = s[i];
#pragma unroll
        for (int i = 0; i < 4; i++) trace[(size_t)(COL_TAIL + i) * limit + row] = add ? s[i] : before[i];
    }
    if (!ok) {
        d->hint_bad = 1;
        atomicOr(&d->failed_checks, HINT_BIT);
        atomicMin(&d->first_bad, ((unsigned long long)row << 16) | HINT_BIT);
    }
}

}  // namespace zkc
