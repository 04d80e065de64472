// TEST INFRASTRUCTURE.  Compiles the row statement the CUDA kernel runs (csrc/main_vm_prestate_row.cuh, __host__ __device__) with
// g++ so that a box without a GPU can compare it with the oracle (tests/test_prestate_row_host.py).  Not part of the product:
// libzkc_b200.so has no CPU path.
#include "../../era_zkevm_circuits_b200/csrc/main_vm_prestate_row.cuh"
#include "../../era_zkevm_circuits_b200/csrc/main_vm_writeback_row.cuh"

extern "C" void prestate_rows_host(const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances, uint64_t *out_all) {
    for (size_t g = 0; g < limit * n_instances; g++) {   // the kernel's index arithmetic, one "thread" at a time
        const size_t inst = g / limit, row = g - inst * limit;
        zkc::vm_prestate_row(trace + inst * (size_t)ZKC_VM_NUM_COLS * limit + row, snapshots + inst * (limit + 1) + row,
                             out_all + inst * (size_t)ZKC_VMP_NUM_COLS * limit + row, limit);
    }
}

// the same for vm_writeback_kernel; the list masks are built as zkc_main_vm_writeback_cells builds them
extern "C" void writeback_rows_host(const zkc_vm_isa *isa, const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances,
                                    uint64_t *out_all) {
    zkc::vm_writeback_masks lists = {0u, 0u};
    for (uint32_t r = 0; r < ZKC_VM_REGISTERS; r++) {
        if (r >= isa->call_system_abi_registers[0] && r < isa->call_system_abi_registers[1]) lists.system_abi |= 1u << r;
        if ((r >= isa->call_reserved_range[0] && r < isa->call_reserved_range[1]) || r == isa->call_implicit_parameter_reg_idx) lists.reserved |= 1u << r;
    }
    for (size_t g = 0; g < limit * n_instances; g++) {
        const size_t inst = g / limit, row = g - inst * limit;
        zkc::vm_writeback_row(trace + inst * (size_t)ZKC_VM_NUM_COLS * limit + row, snapshots + inst * (limit + 1) + row, lists,
                              out_all + inst * (size_t)ZKC_VMW_NUM_COLS * limit + row, limit);
    }
}
