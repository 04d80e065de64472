// Constraint evaluation of a finished main_vm witness trace (the role of the reference's `check_if_satisfied` over the
// cells vm_cycle allocates): one thread per row streams the 276 columns of its cycle once (column-major: a warp reads 256
// contiguous bytes per column) and re-evaluates every relation that is local to the row:
//   - booleanity / range of the flag, register-index, immediate and u32-limb columns (Boolean / UIntX allocation checks);
//   - opcode decoding (decoded_opcode.rs:395-527): the opcode word out of the code word, variant / condition / register /
//     immediate fields, the opcode-table lookup (price, property bit spread), exception masks (:120-157);
//   - the relations of the arithmetic opcodes between src0, src1, dst0, dst1 and the flags: AddSubRelation
//     (opcodes/mod.rs:101-125), MulDivRelation (:129-180: a * b + rem == lo + 2^256 * hi on u32 limbs), bitwise (binop.rs);
//   - dst0 / dst1 as dot products: zero when their update flags are clear (cycle.rs:199-246);
//   - sponge columns: a relation that is not enforced carries zeros.
// Relations that span rows (state carried to the next cycle, queue chains) are what zkc_main_vm_entry_point itself
// verifies against the snapshots.  The kernel is a pure stream: 2 208 algorithmic bytes per row, HBM-bound.
#include "ctx.cuh"

namespace zkc {

struct VmCheckDev {
    unsigned long long violations, first_bad;
    uint32_t failed_checks, pad;
};

__device__ __forceinline__ bool prop_bit(uint64_t props, int bit) { return (props >> bit) & 1; }

#ifndef VM_CHECK_THREADS
#define VM_CHECK_THREADS 128
#endif
#ifndef VM_CHECK_MIN_BLOCKS
#define VM_CHECK_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(VM_CHECK_THREADS, VM_CHECK_MIN_BLOCKS)
vm_check_kernel(VmCheckDev *out, const zkc_vm_isa *__restrict__ isa, const uint64_t *__restrict__ trace, size_t limit, size_t n_instances) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= limit * n_instances) return;
    const size_t inst = g / limit, row = g - inst * limit;
    const uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit + row;
#define TR(col) __ldg(t + (size_t)(col) * limit)
    uint32_t bad = 0;
    // ---- booleans and ranges --------------------------------------------------------------------------------------------
    const uint64_t skip = TR(ZKC_VM_SHOULD_SKIP_CYCLE), pending = TR(ZKC_VM_PENDING_EXCEPTION_IN), read_op = TR(ZKC_VM_SHOULD_READ_OPCODE);
    const uint64_t cond = TR(ZKC_VM_CONDITION), out_of_ergs = TR(ZKC_VM_OUT_OF_ERGS), kexc = TR(ZKC_VM_KERNEL_MODE_EXCEPTION),
                   sexc = TR(ZKC_VM_STATIC_EXCEPTION), full = TR(ZKC_VM_CALLSTACK_IS_FULL), expl = TR(ZKC_VM_EXPLICIT_PANIC),
                   mpanic = TR(ZKC_VM_MASK_INTO_PANIC), mnop = TR(ZKC_VM_MASK_INTO_NOP);
    const uint64_t read_src0 = TR(ZKC_VM_SHOULD_READ_SRC0), dst0_mem = TR(ZKC_VM_DST0_PERFORMS_MEMORY_ACCESS), swap = TR(ZKC_VM_SWAP_OPERANDS),
                   mem_write = TR(ZKC_VM_PERFORM_DST0_MEMORY_WRITE), upd0 = TR(ZKC_VM_DST0_UPDATE_REGISTER), upd1 = TR(ZKC_VM_DST1_UPDATE_REGISTER),
                   pend_out = TR(ZKC_VM_PENDING_EXCEPTION_OUT);
    uint64_t bools = skip | pending | read_op | cond | out_of_ergs | kexc | sexc | full | expl | mpanic | mnop | read_src0 | dst0_mem | swap | mem_write |
                     upd0 | upd1 | pend_out | TR(ZKC_VM_FLAGS_OUT) | TR(ZKC_VM_FLAGS_OUT + 1) | TR(ZKC_VM_FLAGS_OUT + 2) | TR(ZKC_VM_SRC0_FROM_MEMORY) |
                     TR(ZKC_VM_SRC0) | TR(ZKC_VM_SRC1) | TR(ZKC_VM_DST0) | TR(ZKC_VM_DST1);
    if (bools > 1) bad |= ZKC_VMV_BOOLEAN;
    const uint64_t super_pc = TR(ZKC_VM_SUPER_PC), sub_pc = TR(ZKC_VM_SUB_PC);
    const uint64_t src0_r = TR(ZKC_VM_SRC0_REG), src1_r = TR(ZKC_VM_SRC1_REG), dst0_r = TR(ZKC_VM_DST0_REG), dst1_r = TR(ZKC_VM_DST1_REG);
    const uint64_t imm0 = TR(ZKC_VM_IMM0), imm1 = TR(ZKC_VM_IMM1);
    if (sub_pc > 3 || super_pc >> 14 || (src0_r | src1_r | dst0_r | dst1_r) > 15 || (imm0 | imm1) >> 16 ||
        (TR(ZKC_VM_SRC0_INDEX) | TR(ZKC_VM_DST0_INDEX) | TR(ZKC_VM_SP_AFTER_SRC0) | TR(ZKC_VM_NEW_SP) | TR(ZKC_VM_PC_OUT)) >> 16)
        bad |= ZKC_VMV_RANGE;
    uint32_t cw[8], a[8], b[8], d0[8], d1[8];
    uint64_t limbs = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint64_t c = TR(ZKC_VM_CODE_WORD + i), x = TR(ZKC_VM_SRC0 + 1 + i), y = TR(ZKC_VM_SRC1 + 1 + i), z = TR(ZKC_VM_DST0 + 1 + i),
                       w = TR(ZKC_VM_DST1 + 1 + i);
        limbs |= c | x | y | z | w | TR(ZKC_VM_SRC0_FROM_MEMORY + 1 + i);
        cw[i] = (uint32_t)c; a[i] = (uint32_t)x; b[i] = (uint32_t)y; d0[i] = (uint32_t)z; d1[i] = (uint32_t)w;
    }
    limbs |= TR(ZKC_VM_ERGS_COST) | TR(ZKC_VM_DIRTY_ERGS_LEFT) | TR(ZKC_VM_ERGS_OUT) | TR(ZKC_VM_SRC0_PAGE) | TR(ZKC_VM_DST0_PAGE) |
             TR(ZKC_VM_HEAP_BOUND_OUT) | TR(ZKC_VM_AUX_HEAP_BOUND_OUT) | TR(ZKC_VM_MEMQ_LENGTH_OUT) | TR(ZKC_VM_DEPTH_OUT);
    if (limbs >> 32) bad |= ZKC_VMV_RANGE;
    // ---- decoding ---------------------------------------------------------------------------------------------------------
    uint32_t op_lo = 0, op_hi = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) if ((int)sub_pc == i) { op_lo = cw[6 - 2 * i]; op_hi = cw[7 - 2 * i]; }
    if (skip) { op_lo = (uint32_t)isa->nop_opcode_encoding; op_hi = (uint32_t)(isa->nop_opcode_encoding >> 32); }
    if (pending) { op_lo = (uint32_t)isa->panic_opcode_encoding; op_hi = (uint32_t)(isa->panic_opcode_encoding >> 32); }
    if (TR(ZKC_VM_OPCODE) != op_lo || TR(ZKC_VM_OPCODE + 1) != op_hi) bad |= ZKC_VMV_DECODE;
    const uint32_t variant = op_lo & 0x7FF;
    if (TR(ZKC_VM_VARIANT) != variant || TR(ZKC_VM_CONDITION_IDX) != ((op_lo >> 13) & 7) || imm0 != (op_hi & 0xFFFF) || imm1 != (op_hi >> 16)) bad |= ZKC_VMV_DECODE;
    const uint64_t props_full = isa->opcode_props[variant];
    constexpr uint64_t MASK48 = (1ull << ZKC_VM_DESCRIPTION_BITS_FLATTENED) - 1;
    const uint32_t aux = (uint32_t)(props_full >> ZKC_VM_DESCRIPTION_BITS_FLATTENED);
    if (TR(ZKC_VM_ERGS_COST) != (skip ? 0u : isa->opcode_price[variant]) || expl != ((aux >> ZKC_VM_AUX_EXPLICIT_PANIC) & 1)) bad |= ZKC_VMV_DECODE;
    if (mpanic != (expl | out_of_ergs | kexc | sexc | full) || mnop != (uint64_t)(!mpanic && !cond)) bad |= ZKC_VMV_EXCEPTION_MASKS;
    if (kexc && !((aux >> ZKC_VM_AUX_KERNEL_MODE) & 1)) bad |= ZKC_VMV_EXCEPTION_MASKS;
    if (sexc && ((aux >> ZKC_VM_AUX_CAN_BE_USED_IN_STATIC) & 1)) bad |= ZKC_VMV_EXCEPTION_MASKS;
    if (out_of_ergs && TR(ZKC_VM_DIRTY_ERGS_LEFT) != 0) bad |= ZKC_VMV_EXCEPTION_MASKS;
    uint64_t props = props_full & MASK48;
    if (mpanic) props = isa->panic_bitspread & MASK48;
    if (mnop) props = isa->nop_bitspread & MASK48;
    if (TR(ZKC_VM_PROPS) != props) bad |= ZKC_VMV_DECODE;
    const bool masked = mpanic || mnop;
    uint32_t sregs = (op_lo >> 16) & 0xFF, dregs = op_lo >> 24;
    if (masked) { sregs = 0; dregs = 0; }
    if (src0_r != (sregs & 15) || src1_r != (sregs >> 4) || dst0_r != (dregs & 15) || dst1_r != (dregs >> 4)) bad |= ZKC_VMV_DECODE;
    // ---- the arithmetic relations -------------------------------------------------------------------------------------------
#define TYPE(tt) prop_bit(props, ZKC_VM_BIT_TYPE(tt))
#define VAR(v) prop_bit(props, ZKC_VM_BIT_VARIANT(v))
    const bool set_flags = prop_bit(props, ZKC_VM_BIT_FLAG(ZKC_VM_SET_FLAGS_FLAG_IDX));
    const uint64_t f0 = TR(ZKC_VM_FLAGS_OUT), f1 = TR(ZKC_VM_FLAGS_OUT + 1), f2 = TR(ZKC_VM_FLAGS_OUT + 2);
    bool d0_zero = true, d1_zero = true, b_zero = true;
#pragma unroll
    for (int i = 0; i < 8; i++) { d0_zero &= d0[i] == 0; d1_zero &= d1[i] == 0; b_zero &= b[i] == 0; }
    if (TYPE(ZKC_OP_ADD) || TYPE(ZKC_OP_SUB)) {  // a + b = c + 2^256 * of (add), a = c + b - 2^256 * of (sub): enforce_addition_relation
        const bool sub = TYPE(ZKC_OP_SUB);
        uint64_t carry = 0;
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint64_t s = (uint64_t)(sub ? d0[i] : a[i]) + b[i] + carry;
            ok &= (uint32_t)s == (sub ? a[i] : d0[i]);
            carry = s >> 32;
        }
        if (!ok) bad |= ZKC_VMV_ADD_SUB;
        if (set_flags && (f0 != carry || f1 != (uint64_t)d0_zero || f2 != (uint64_t)!(carry || d0_zero))) bad |= ZKC_VMV_FLAGS;
    }
    if (TYPE(ZKC_OP_MUL) || TYPE(ZKC_OP_DIV)) {  // a * b + rem = lo + 2^256 * hi: enforce_mul_relation (8 x 8 u32 schoolbook)
        const bool div = TYPE(ZKC_OP_DIV);
        // mul: a * b = d0 + 2^256 d1.  div: d0 (quotient) * b + d1 (remainder) = a, remainder < b (b != 0); b == 0: both zero
        uint32_t x[8], r[16];  // (a pointer select would push a / d0 into local memory)
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = div ? d0[i] : a[i];
        const uint32_t (&y)[8] = b;
#pragma unroll
        for (int i = 0; i < 16; i++) r[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t carry = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint64_t tt = (uint64_t)x[i] * y[j] + r[i + j] + carry;
                r[i + j] = (uint32_t)tt; carry = tt >> 32;
            }
            r[i + 8] = (uint32_t)carry;
        }
        bool ok = true;
        if (!div) {
#pragma unroll
            for (int i = 0; i < 8; i++) ok &= r[i] == d0[i] && r[8 + i] == d1[i];
            const bool of = !d1_zero;
            if (set_flags && (f0 != (uint64_t)of || f1 != (uint64_t)d0_zero || f2 != (uint64_t)(!of && !d0_zero))) bad |= ZKC_VMV_FLAGS;
        } else if (b_zero) {
            ok = d0_zero && d1_zero;
            if (set_flags && (f0 != 1 || f1 != 0 || f2 != 0)) bad |= ZKC_VMV_FLAGS;
        } else {
            uint64_t carry = 0;
            bool lt = false;  // remainder < divisor, MSW first
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint64_t s = (uint64_t)r[i] + d1[i] + carry;
                ok &= (uint32_t)s == a[i];
                carry = s >> 32;
                ok &= r[8 + i] == 0;
            }
            ok &= carry == 0;
            bool decided = false;
#pragma unroll
            for (int i = 7; i >= 0; i--) if (!decided && d1[i] != b[i]) { lt = d1[i] < b[i]; decided = true; }
            ok &= lt;
            if (set_flags && (f0 != 0 || f1 != (uint64_t)d0_zero || f2 != (uint64_t)d1_zero)) bad |= ZKC_VMV_FLAGS;
        }
        if (!ok) bad |= ZKC_VMV_MUL_DIV;
    }
    if (TYPE(ZKC_OP_BINOP)) {
        const bool is_or = VAR(ZKC_VAR_BINOP_OR), is_and = VAR(ZKC_VAR_BINOP_AND);
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 8; i++) ok &= d0[i] == (is_or ? (a[i] | b[i]) : (is_and ? (a[i] & b[i]) : (a[i] ^ b[i])));
        if (!ok) bad |= ZKC_VMV_BINOP;
        if (set_flags && (f0 != 0 || f1 != (uint64_t)d0_zero || f2 != 0)) bad |= ZKC_VMV_FLAGS;
    }
#undef TYPE
#undef VAR
    // ---- selection: dst0 / dst1 are dot products of (flag, candidate) pairs; memory write needs a memory destination -------
    if (!upd0 && !mem_write && !(d0_zero && TR(ZKC_VM_DST0) == 0)) bad |= ZKC_VMV_SELECTION;
    if (!upd1 && !(d1_zero && TR(ZKC_VM_DST1) == 0)) bad |= ZKC_VMV_SELECTION;
    if (mem_write && !dst0_mem) bad |= ZKC_VMV_SELECTION;
    if (upd0 && mem_write) bad |= ZKC_VMV_SELECTION;
    // ---- sponge columns: zeros unless enforced; the opcode-specific block is zero for the plain opcodes -------------------
    uint64_t stray = 0;
#pragma unroll
    for (int k = 0; k < ZKC_VM_NUM_SPONGES; k++) {  // fully unrolled: 117 independent loads, issued as far ahead as registers allow
        const uint64_t enf = TR(ZKC_VM_SPONGE_ENFORCE + k);
        if (enf > 1) bad |= ZKC_VMV_BOOLEAN;
        uint64_t any = 0;
#pragma unroll
        for (int j = 0; j < 12; j++) {
            const uint64_t v = TR(ZKC_VM_SPONGE_FINAL + 12 * k + j);
            any |= v;
            if (v >= ZKC_GL_P) bad |= ZKC_VMV_RANGE;
        }
        if (!enf) stray |= any;
    }
    if (stray) bad |= ZKC_VMV_SPONGE;
    if (read_op != TR(ZKC_VM_SPONGE_ENFORCE) || TR(ZKC_VM_SPONGE_ENFORCE + 1) < read_src0 || TR(ZKC_VM_SPONGE_ENFORCE + 2) < mem_write) bad |= ZKC_VMV_SPONGE;
    {
        const bool family = prop_bit(props, ZKC_VM_BIT_TYPE(ZKC_OP_UMA)) || prop_bit(props, ZKC_VM_BIT_TYPE(ZKC_OP_LOG)) ||
                            prop_bit(props, ZKC_VM_BIT_TYPE(ZKC_OP_NEAR_CALL)) || prop_bit(props, ZKC_VM_BIT_TYPE(ZKC_OP_FAR_CALL)) ||
                            prop_bit(props, ZKC_VM_BIT_TYPE(ZKC_OP_RET));
        uint64_t any = 0;
#pragma unroll
        for (int i = 0; i < ZKC_VM_OP_AUX_COLS; i++) any |= TR(ZKC_VM_OP_AUX + i);
        if (!family && any) bad |= ZKC_VMV_SELECTION;
    }
    // the remaining columns are streamed too (every cell is read once): forward / rollback queue ends are field elements
    {
        uint64_t big = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) big |= (uint64_t)(TR(ZKC_VM_FORWARD_TAIL_OUT + i) >= ZKC_GL_P) | (uint64_t)(TR(ZKC_VM_ROLLBACK_HEAD_OUT + i) >= ZKC_GL_P);
        if (big || (TR(ZKC_VM_FORWARD_TAIL_OUT + 4) | TR(ZKC_VM_ROLLBACK_HEAD_OUT + 4)) >> 32) bad |= ZKC_VMV_RANGE;
    }
#undef TR
    if (bad) {
        atomicAdd(&out->violations, 1ull);
        atomicOr(&out->failed_checks, bad);
        atomicMin(&out->first_bad, ((unsigned long long)g << 16) | (bad & 0xFFFFu));
    }
}

}  // namespace zkc

using namespace zkc;

extern "C" int zkc_main_vm_check_trace(zkc_ctx *ctx, const zkc_vm_isa *isa, const uint64_t *trace, size_t limit, size_t n_instances,
                                       int on_device, uint64_t *violations, zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !isa || !violations || ((limit * n_instances) && !trace)) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    *violations = 0;
    const size_t rows = limit * n_instances;
    if (!rows) return ZKC_OK;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const bool dev = on_device != 0;
    size_t bytes = zkc_carver::bytes(1, sizeof(VmCheckDev)) + zkc_carver::bytes(1, sizeof(zkc_vm_isa));
    if (!dev) bytes += zkc_carver::bytes(rows * ZKC_VM_NUM_COLS, 8);
    void *blk = ctx->scratch(bytes);
    VmCheckDev *h = (VmCheckDev *)ctx->pinned(sizeof(VmCheckDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    VmCheckDev *d = cv.take<VmCheckDev>(1);
    zkc_vm_isa *disa = cv.take<zkc_vm_isa>(1);
    cudaStream_t s = ctx->stream;
    h->violations = 0; h->first_bad = ~0ull; h->failed_checks = 0; h->pad = 0;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof *h, cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(disa, isa, sizeof *isa, cudaMemcpyHostToDevice, s));
    const uint64_t *dt = trace;
    if (!dev) {
        uint64_t *buf = cv.take<uint64_t>(rows * ZKC_VM_NUM_COLS);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(buf, trace, rows * ZKC_VM_NUM_COLS * 8, cudaMemcpyHostToDevice, s));
        dt = buf;
    }
    ZKC_LAUNCH(ctx, "vm_check", vm_check_kernel, (unsigned)((rows + VM_CHECK_THREADS - 1) / VM_CHECK_THREADS), VM_CHECK_THREADS, 0, d, disa, dt, limit,
               n_instances);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof *h, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    *violations = h->violations;
    status->failed_checks = h->failed_checks;
    if (h->violations) {
        status->code = ZKC_ERR_UNSATISFIED;
        status->first_bad_row = (int64_t)(h->first_bad >> 16);
    }
    return status->code;
}
