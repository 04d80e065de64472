// The memory-queue relations every main_vm cycle evaluates whatever its opcode (include/zkc_b200.h, ZKC_VM_MEMORY_SPONGE_COLUMNS):
// opcode fetch, src0 read, dst0 write -- query encoding absorbed with replacement into the running memory-queue tail, the
// Poseidon2 permutation of that initial state, the tail / length selected by the access flag.
//   may_be_read_memory_for_code             /root/reference/src/main_vm/utils.rs:129-233
//   may_be_read_memory_for_source_operand   /root/reference/src/main_vm/utils.rs:388-522
//   may_be_write_memory                     /root/reference/src/main_vm/cycle.rs:799-935
//   enforce_sponges                         /root/reference/src/main_vm/cycle.rs:937-957
// One thread per cycle: three DEPENDENT permutations (the tail of one step is the capacity of the next), so the kernel is bound by
// the integer pipe like every other Poseidon2 kernel of the engine (3 x ~19 k instructions per cycle); 41 trace columns + 27
// snapshot words in, 112 columns out (896 B per cycle), every store coalesced across the warp's 32 consecutive cycles.
#include "ctx.cuh"
#include "poseidon2.cuh"

namespace zkc {

// MemoryQuery::encode, /root/reference/src/base_structures/memory_query/mod.rs:103-221 (v: the 8 limbs of the value)
__device__ __forceinline__ void vmq_encode(uint32_t ts, uint32_t page, uint32_t index, uint32_t rw, uint32_t is_ptr, const uint32_t (&v)[8], uint64_t (&e)[8]) {
    e[0] = ts; e[1] = page;
    e[2] = (uint64_t)index | ((uint64_t)rw << 32) | ((uint64_t)(is_ptr & 1) << 33);
    e[3] = (uint64_t)v[0] | ((uint64_t)(v[5] & 0xFFFFFFu) << 32);
    e[4] = (uint64_t)v[1] | ((uint64_t)(v[5] >> 24) << 32) | ((uint64_t)(v[6] & 0xFFFFu) << 40);
    e[5] = (uint64_t)v[2] | ((uint64_t)(v[6] >> 16) << 32) | ((uint64_t)(v[7] & 0xFFu) << 48);
    e[6] = (uint64_t)v[3] | ((uint64_t)(v[7] >> 8) << 32);
    e[7] = v[4];
}

// one step: INIT (12), FINAL (12), STATE_AFTER (12), LENGTH_AFTER (1) at column `col`
__device__ __forceinline__ void vmq_step(const uint64_t (&enc)[8], bool execute, uint64_t (&state)[12], uint32_t &len, uint64_t *out, size_t limit, int col) {
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        s[i] = i < 8 ? enc[i] : state[i];  // absorb with replacement
        out[(size_t)(col + i) * limit] = s[i];
    }
    poseidon2_permute(s);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        out[(size_t)(col + 12 + i) * limit] = s[i];
        if (execute) state[i] = s[i];       // Num::parallel_select
        out[(size_t)(col + 24 + i) * limit] = state[i];
    }
    len += execute ? 1u : 0u;               // UInt32::conditionally_select(new_len_candidate, current)
    out[(size_t)(col + 36) * limit] = len;
}

__global__ void __launch_bounds__(128)
vm_memory_sponges_kernel(const uint64_t *__restrict__ trace, const zkc_vm_state *__restrict__ snapshots, size_t limit, size_t n_instances,
                         uint64_t *__restrict__ out_all) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= limit * n_instances) return;
    const size_t inst = g / limit, row = g - inst * limit;
    const uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit + row;
    const zkc_vm_state *st = snapshots + inst * (limit + 1) + row;
    uint64_t *out = out_all + inst * (size_t)ZKC_VMQ_NUM_COLS * limit + row;
#define T(c) __ldg(t + (size_t)(c) * limit)
    const uint64_t props = T(ZKC_VM_PROPS);
    const uint64_t own_sponges = (1ull << ZKC_VM_BIT_TYPE(ZKC_OP_UMA)) | (1ull << ZKC_VM_BIT_TYPE(ZKC_OP_LOG)) | (1ull << ZKC_VM_BIT_TYPE(ZKC_OP_NEAR_CALL)) |
                                 (1ull << ZKC_VM_BIT_TYPE(ZKC_OP_FAR_CALL)) | (1ull << ZKC_VM_BIT_TYPE(ZKC_OP_RET));
    out[(size_t)ZKC_VMQ_SELECTED * limit] = (props & own_sponges) == 0;
    uint64_t state[12], enc[8];
#pragma unroll
    for (int i = 0; i < 12; i++) state[i] = __ldg(&st->memory_queue_state[i]);
    uint32_t len = __ldg(&st->memory_queue_length);
    const uint32_t ts = __ldg(&st->timestamp);
    uint32_t v[8];
    {   // opcode fetch
        const bool read_opcode = T(ZKC_VM_SHOULD_READ_OPCODE) != 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { const uint32_t w = (uint32_t)T(ZKC_VM_CODE_WORD + i); v[i] = read_opcode ? w : 0u; }
        vmq_encode(ts, __ldg(&st->current_context.code_page), (uint32_t)T(ZKC_VM_SUPER_PC), 0, 0, v, enc);
        vmq_step(enc, read_opcode, state, len, out, limit, ZKC_VMQ_FETCH_INIT);
    }
    {   // src0 read
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (uint32_t)T(ZKC_VM_SRC0_FROM_MEMORY + 1 + i);
        vmq_encode(ts, (uint32_t)T(ZKC_VM_SRC0_PAGE), (uint32_t)T(ZKC_VM_SRC0_INDEX), 0, (uint32_t)T(ZKC_VM_SRC0_FROM_MEMORY), v, enc);
        vmq_step(enc, T(ZKC_VM_SHOULD_READ_SRC0) != 0, state, len, out, limit, ZKC_VMQ_SRC0_INIT);
    }
    {   // dst0 write
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (uint32_t)T(ZKC_VM_DST0 + 1 + i);
        vmq_encode(ts + 3u, (uint32_t)T(ZKC_VM_DST0_PAGE), (uint32_t)T(ZKC_VM_DST0_INDEX), 1, (uint32_t)T(ZKC_VM_DST0), v, enc);
        vmq_step(enc, T(ZKC_VM_PERFORM_DST0_MEMORY_WRITE) != 0, state, len, out, limit, ZKC_VMQ_DST0_INIT);
    }
#undef T
}

}  // namespace zkc

using namespace zkc;

extern "C" int zkc_main_vm_memory_sponge_cells(zkc_ctx *ctx, const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances,
                                               int on_device, uint64_t *sponge_trace) {
    if (!ctx || ((limit * n_instances) && (!trace || !snapshots || !sponge_trace))) return ZKC_ERR_INVALID_ARGUMENT;
    const size_t rows = limit * n_instances, n_snaps = (limit + 1) * n_instances;
    if (!rows) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t *dt = trace;
    const zkc_vm_state *ds = snapshots;
    uint64_t *dg = sponge_trace;
    if (!on_device) {
        char *blk = (char *)ctx->scratch(zkc_carver::bytes(rows * ZKC_VM_NUM_COLS, 8) + zkc_carver::bytes(n_snaps, sizeof(zkc_vm_state)) +
                                         zkc_carver::bytes(rows * ZKC_VMQ_NUM_COLS, 8));
        if (!blk) return ZKC_ERR_CUDA;
        zkc_carver cv(blk);
        uint64_t *bt = cv.take<uint64_t>(rows * ZKC_VM_NUM_COLS);
        zkc_vm_state *bs = cv.take<zkc_vm_state>(n_snaps);
        dg = cv.take<uint64_t>(rows * ZKC_VMQ_NUM_COLS);
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(bt, trace, rows * ZKC_VM_NUM_COLS * 8, cudaMemcpyHostToDevice, s));
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(bs, snapshots, n_snaps * sizeof(zkc_vm_state), cudaMemcpyHostToDevice, s));
        dt = bt; ds = bs;
    }
    ZKC_LAUNCH(ctx, "vm_memory_sponges", vm_memory_sponges_kernel, (unsigned)((rows + 127) / 128), 128, 0, dt, ds, limit, n_instances, dg);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    if (!on_device) ZKC_CUDA(ctx, st, cudaMemcpyAsync(sponge_trace, dg, rows * ZKC_VMQ_NUM_COLS * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    return ZKC_OK;
}
