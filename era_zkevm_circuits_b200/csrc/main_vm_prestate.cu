// zkc_main_vm_prestate_cells (include/zkc_b200.h, ZKC_VM_PRESTATE_COLUMNS): the cells create_prestate allocates on the way to the
// values the DENSE trace names -- cycle control, the opcode select inside the code word, the four register selector masks, the
// 15-step register select chains, operand locations, the src0 selects, the operand swap and the pointer-erasure flags
// (/root/reference/src/main_vm/pre_state.rs:71-519; the row statement and its citations are in main_vm_prestate_row.cuh).
// One thread per cycle: 29 coalesced trace columns + 148 words of its snapshot record in (the 15 registers are 540 contiguous
// bytes of the 1 176-byte record: each of a thread's five 128-byte lines is read once from L2 and then served by L1 across the
// unrolled chain), 428 columns out -- an HBM-write-bound stream, 3 424 B per cycle against ~0.8 KB read.
#include "ctx.cuh"
#include "main_vm_prestate_row.cuh"
#include "main_vm_writeback_row.cuh"

namespace zkc {

__global__ void __launch_bounds__(128)
vm_prestate_kernel(const uint64_t *__restrict__ trace, const zkc_vm_state *__restrict__ snapshots, size_t limit, size_t n_instances,
                   uint64_t *__restrict__ out_all) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= limit * n_instances) return;
    const size_t inst = g / limit, row = g - inst * limit;
    vm_prestate_row(trace + inst * (size_t)ZKC_VM_NUM_COLS * limit + row, snapshots + inst * (limit + 1) + row,
                    out_all + inst * (size_t)ZKC_VMP_NUM_COLS * limit + row, limit);
}

// zkc_main_vm_writeback_cells (ZKC_VM_WRITEBACK_COLUMNS): the register write-back of the state diffs, cycle.rs:158-433 (row statement and
// citations in main_vm_writeback_row.cuh).  One thread per cycle: 24 coalesced trace columns + the 15 registers of its snapshot and r1
// of the next one in, 513 columns out -- HBM-write-bound, 4 104 B per cycle.
__global__ void __launch_bounds__(128)
vm_writeback_kernel(const uint64_t *__restrict__ trace, const zkc_vm_state *__restrict__ snapshots, size_t limit, size_t n_instances,
                    vm_writeback_masks lists, uint64_t *__restrict__ out_all) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= limit * n_instances) return;
    const size_t inst = g / limit, row = g - inst * limit;
    vm_writeback_row(trace + inst * (size_t)ZKC_VM_NUM_COLS * limit + row, snapshots + inst * (limit + 1) + row, lists,
                     out_all + inst * (size_t)ZKC_VMW_NUM_COLS * limit + row, limit);
}

}  // namespace zkc

using namespace zkc;

extern "C" int zkc_main_vm_prestate_cells(zkc_ctx *ctx, const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances,
                                          int on_device, uint64_t *prestate_trace) {
    if (!ctx || ((limit * n_instances) && (!trace || !snapshots || !prestate_trace))) return ZKC_ERR_INVALID_ARGUMENT;
    const size_t rows = limit * n_instances, n_snaps = (limit + 1) * n_instances;
    if (!rows) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t *dt = trace;
    const zkc_vm_state *ds = snapshots;
    uint64_t *dp = prestate_trace;
    if (!on_device) {
        char *blk = (char *)ctx->scratch(zkc_carver::bytes(rows * ZKC_VM_NUM_COLS, 8) + zkc_carver::bytes(n_snaps, sizeof(zkc_vm_state)) +
                                         zkc_carver::bytes(rows * ZKC_VMP_NUM_COLS, 8));
        if (!blk) return ZKC_ERR_CUDA;
        zkc_carver cv(blk);
        uint64_t *bt = cv.take<uint64_t>(rows * ZKC_VM_NUM_COLS);
        zkc_vm_state *bs = cv.take<zkc_vm_state>(n_snaps);
        dp = cv.take<uint64_t>(rows * ZKC_VMP_NUM_COLS);
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(bt, trace, rows * ZKC_VM_NUM_COLS * 8, cudaMemcpyHostToDevice, s));
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(bs, snapshots, n_snaps * sizeof(zkc_vm_state), cudaMemcpyHostToDevice, s));
        dt = bt; ds = bs;
    }
    ZKC_LAUNCH(ctx, "vm_prestate", vm_prestate_kernel, (unsigned)((rows + 127) / 128), 128, 0, dt, ds, limit, n_instances, dp);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    if (!on_device) ZKC_CUDA(ctx, st, cudaMemcpyAsync(prestate_trace, dp, rows * ZKC_VMP_NUM_COLS * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    return ZKC_OK;
}

extern "C" int zkc_main_vm_writeback_cells(zkc_ctx *ctx, const zkc_vm_isa *isa, const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit,
                                           size_t n_instances, int on_device, uint64_t *writeback_trace) {
    if (!ctx || !isa || ((limit * n_instances) && (!trace || !snapshots || !writeback_trace))) return ZKC_ERR_INVALID_ARGUMENT;
    const size_t rows = limit * n_instances, n_snaps = (limit + 1) * n_instances;
    if (!rows) return ZKC_OK;
    vm_writeback_masks lists = {0u, 0u};   // far_call.rs:1048-1070: the calling-convention register lists as 15-bit masks
    for (uint32_t r = 0; r < ZKC_VM_REGISTERS; r++) {
        if (r >= isa->call_system_abi_registers[0] && r < isa->call_system_abi_registers[1]) lists.system_abi |= 1u << r;
        if ((r >= isa->call_reserved_range[0] && r < isa->call_reserved_range[1]) || r == isa->call_implicit_parameter_reg_idx) lists.reserved |= 1u << r;
    }
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t *dt = trace;
    const zkc_vm_state *ds = snapshots;
    uint64_t *dw = writeback_trace;
    if (!on_device) {
        char *blk = (char *)ctx->scratch(zkc_carver::bytes(rows * ZKC_VM_NUM_COLS, 8) + zkc_carver::bytes(n_snaps, sizeof(zkc_vm_state)) +
                                         zkc_carver::bytes(rows * ZKC_VMW_NUM_COLS, 8));
        if (!blk) return ZKC_ERR_CUDA;
        zkc_carver cv(blk);
        uint64_t *bt = cv.take<uint64_t>(rows * ZKC_VM_NUM_COLS);
        zkc_vm_state *bs = cv.take<zkc_vm_state>(n_snaps);
        dw = cv.take<uint64_t>(rows * ZKC_VMW_NUM_COLS);
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(bt, trace, rows * ZKC_VM_NUM_COLS * 8, cudaMemcpyHostToDevice, s));
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(bs, snapshots, n_snaps * sizeof(zkc_vm_state), cudaMemcpyHostToDevice, s));
        dt = bt; ds = bs;
    }
    ZKC_LAUNCH(ctx, "vm_writeback", vm_writeback_kernel, (unsigned)((rows + 127) / 128), 128, 0, dt, ds, limit, n_instances, lists, dw);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    if (!on_device) ZKC_CUDA(ctx, st, cudaMemcpyAsync(writeback_trace, dw, rows * ZKC_VMW_NUM_COLS * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    return ZKC_OK;
}
