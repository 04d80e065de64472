// Shared pieces of the precompile circuits (keccak256_round_function, sha256_round_function): the memory queue
// they push into (MemoryQueue = FullStateCircuitQueue<_, MemoryQuery, 8, 12, 4, 8, R>; push rule restated in-repo at
// /root/reference/src/main_vm/utils.rs:194-212).  The call kernels leave, per executed push, its 8-element encoding
// (push order) and, per (row, slot), the number of pushes executed up to and including that slot (bit 31 = this slot
// pushed).  The tail after every slot is then either verified against host-supplied states or rebuilt by the
// sequential chain kernel (one permutation per push).
// `Dev`: per-call device block with members mq0, limit, hint_bad, failed_checks, first_bad.
#pragma once
#include "poseidon2.cuh"

namespace zkc {

__device__ __forceinline__ void mq_encode(uint32_t ts, uint32_t page, uint32_t index, uint32_t rw, const uint32_t *v, uint64_t *e) {
    // MemoryQuery::encode, base_structures/memory_query/mod.rs:103-221 (is_ptr = false)
    e[0] = ts; e[1] = page;
    e[2] = (uint64_t)index | ((uint64_t)rw << 32);
    e[3] = (uint64_t)v[0] | ((uint64_t)(v[5] & 0xFFFFFFu) << 32);
    e[4] = (uint64_t)v[1] | ((uint64_t)(v[5] >> 24) << 32) | ((uint64_t)(v[6] & 0xFFFFu) << 40);
    e[5] = (uint64_t)v[2] | ((uint64_t)(v[6] >> 16) << 32) | ((uint64_t)(v[7] & 0xFFu) << 48);
    e[6] = (uint64_t)v[3] | ((uint64_t)(v[7] >> 8) << 32);
    e[7] = v[4];
}


// the chain when the caller supplies no states: one permutation per push on 12 cooperating lanes (poseidon2_permute_coop; the
// full-state queue's next tail is the whole permutation output, so lane i simply keeps s[i]).  Launch with one warp.
template <class Dev, int SLOTS>
__global__ void pc_mem_chain_kernel(const Dev *d, const uint64_t *__restrict__ push_enc, const uint32_t *__restrict__ slot_meta,
                                    uint64_t *__restrict__ states) {
    const int i = threadIdx.x;
    if (blockIdx.x != 0 || i >= 16) return;
    const unsigned gm = 0xFFFFu;
    uint64_t x = i < 12 ? d->mq0.tail[i] : 0ull;
    const size_t limit = d->limit;
    const uint32_t n = limit ? (slot_meta[SLOTS * (limit - 1) + SLOTS - 1] & 0x7FFFFFFFu) : 0;
    for (uint32_t k = 0; k < n; k++) {
        if (i < 8) x = push_enc[8 * (size_t)k + i];
        x = poseidon2_permute_coop(gm, x, i);
        if (i < 12) states[12 * (size_t)k + i] = x;
    }
}

// one thread per (row, slot); the last slot of a row is the digest write (tail columns at WRITE_COL), the others are
// reads (tail columns at READ_COL0 + slot * READ_STRIDE); the queue length follows the tail
template <class Dev, int SLOTS, int READ_COL0, int READ_STRIDE, int WRITE_COL, uint32_t HINT_BIT>
__global__ void __launch_bounds__(256)
pc_memq_kernel(Dev *d, const uint64_t *__restrict__ push_enc, const uint32_t *__restrict__ slot_meta,
               const uint64_t *__restrict__ states, size_t n_states, bool verify, uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= SLOTS * limit) return;
    const size_t row = t / SLOTS;
    const int slot = (int)(t % SLOTS);
    const uint32_t m = slot_meta[t];
    const uint32_t ord = m & 0x7FFFFFFFu;
    const bool pushed = m >> 31;
    uint64_t cur[12];
    bool ok = true;
    if (ord == 0) {
#pragma unroll
        for (int i = 0; i < 12; i++) cur[i] = d->mq0.tail[i];
    } else if (ord - 1 < n_states) {
#pragma unroll
        for (int i = 0; i < 12; i++) cur[i] = __ldg(states + 12 * (size_t)(ord - 1) + i);
    } else {
        ok = false;
#pragma unroll
        for (int i = 0; i < 12; i++) cur[i] = 0;
    }
    if (pushed && verify && ok) {
        uint64_t s[12];
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = push_enc[8 * (size_t)(ord - 1) + i];
        if (ord == 1) {
#pragma unroll
            for (int i = 8; i < 12; i++) s[i] = d->mq0.tail[i];
        } else {
#pragma unroll
            for (int i = 8; i < 12; i++) s[i] = __ldg(states + 12 * (size_t)(ord - 2) + i);
        }
        poseidon2_permute(s);
#pragma unroll
        for (int i = 0; i < 12; i++) ok &= s[i] == cur[i];
    }
    if (trace) {
        const int base = slot < SLOTS - 1 ? READ_COL0 + slot * READ_STRIDE : WRITE_COL;
#pragma unroll
        for (int i = 0; i < 12; i++) trace[(size_t)(base + i) * limit + row] = cur[i];
        trace[(size_t)(base + 12) * limit + row] = d->mq0.length + ord;
    }
    if (!ok) {
        d->hint_bad = 1;
        atomicOr(&d->failed_checks, HINT_BIT);
        atomicMin(&d->first_bad, ((unsigned long long)row << 16) | HINT_BIT);
    }
}

}  // namespace zkc
