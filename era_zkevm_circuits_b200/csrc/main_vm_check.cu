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
#define VM_CHECK_MIN_BLOCKS 3
#endif
// R rows per thread.  R = 2: the thread owns the row pair (2k, 2k + 1) and reads every column with ONE 128-bit load
// (LDG.E.128: twice the bytes in flight per load instruction, half the instructions per byte); the pair is 16-byte aligned
// when `limit` is even.  R = 1 (odd limit, any alignment): 64-bit loads.  Values that are only range-checked or OR-reduced
// are consumed as they arrive; the few columns the relations need stay in registers as 32-bit limbs -- no array is indexed
// with a run-time value, so nothing lives in local memory.
template <int R> struct VmCells { uint64_t v[R]; };
template <int R> __device__ __forceinline__ VmCells<R> vm_ld(const uint64_t *p);
template <> __device__ __forceinline__ VmCells<1> vm_ld<1>(const uint64_t *p) { return VmCells<1>{{__ldg(p)}}; }
template <> __device__ __forceinline__ VmCells<2> vm_ld<2>(const uint64_t *p) {
    const ulonglong2 q = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    return VmCells<2>{{q.x, q.y}};
}

// the 26 boolean columns, in the order of their bit in the per-row flag mask
enum : int { VB_SKIP = 0, VB_PENDING, VB_READ_OP, VB_COND, VB_OUT_OF_ERGS, VB_KEXC, VB_SEXC, VB_FULL, VB_EXPL, VB_MPANIC, VB_MNOP, VB_READ_SRC0,
              VB_DST0_MEM, VB_SWAP, VB_MEM_WRITE, VB_UPD0, VB_UPD1, VB_PEND_OUT, VB_F0, VB_F1, VB_F2, VB_SRC0_MEM_PTR, VB_A_PTR, VB_B_PTR,
              VB_D0_PTR, VB_D1_PTR, VB_COUNT };
__device__ constexpr int VB_COL[VB_COUNT] = {
    ZKC_VM_SHOULD_SKIP_CYCLE, ZKC_VM_PENDING_EXCEPTION_IN, ZKC_VM_SHOULD_READ_OPCODE, ZKC_VM_CONDITION, ZKC_VM_OUT_OF_ERGS, ZKC_VM_KERNEL_MODE_EXCEPTION,
    ZKC_VM_STATIC_EXCEPTION, ZKC_VM_CALLSTACK_IS_FULL, ZKC_VM_EXPLICIT_PANIC, ZKC_VM_MASK_INTO_PANIC, ZKC_VM_MASK_INTO_NOP, ZKC_VM_SHOULD_READ_SRC0,
    ZKC_VM_DST0_PERFORMS_MEMORY_ACCESS, ZKC_VM_SWAP_OPERANDS, ZKC_VM_PERFORM_DST0_MEMORY_WRITE, ZKC_VM_DST0_UPDATE_REGISTER, ZKC_VM_DST1_UPDATE_REGISTER,
    ZKC_VM_PENDING_EXCEPTION_OUT, ZKC_VM_FLAGS_OUT, ZKC_VM_FLAGS_OUT + 1, ZKC_VM_FLAGS_OUT + 2, ZKC_VM_SRC0_FROM_MEMORY, ZKC_VM_SRC0, ZKC_VM_SRC1,
    ZKC_VM_DST0, ZKC_VM_DST1};

template <int R>
__global__ void __launch_bounds__(VM_CHECK_THREADS, VM_CHECK_MIN_BLOCKS)
vm_check_kernel(VmCheckDev *out, const zkc_vm_isa *__restrict__ isa, const uint64_t *__restrict__ trace, size_t limit, size_t n_instances) {
    const size_t per_inst = limit / R;  // R == 2: limit is even
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= per_inst * n_instances) return;
    const size_t inst = u / per_inst, row = (u - inst * per_inst) * R;
    const uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit + row;
#define LD(col) vm_ld<R>(t + (size_t)(col) * limit)
#define FOR_R _Pragma("unroll") for (int r = 0; r < R; r++)
#define STAGE asm volatile("" ::: "memory")  // keeps the loads of a later stage from being hoisted over this one (registers)
    uint32_t bad[R], fb[R];
    FOR_R { bad[r] = 0; fb[r] = 0; }
    // ---- booleans: range-checked as they arrive, kept as one bit each --------------------------------------------------------
    {
        uint64_t acc[R];
        FOR_R acc[r] = 0;
#pragma unroll
        for (int k = 0; k < VB_COUNT; k++) {
            const VmCells<R> v = LD(VB_COL[k]);
            FOR_R { acc[r] |= v.v[r]; fb[r] |= ((uint32_t)v.v[r] & 1u) << k; }
        }
        FOR_R if (acc[r] > 1) bad[r] |= ZKC_VMV_BOOLEAN;
    }
#define B(k) ((fb[r] >> (k)) & 1u)
    STAGE;
    // ---- small integers --------------------------------------------------------------------------------------------------
    uint32_t sub_pc[R], regs4[R], imm[R];
    {
        const VmCells<R> super_pc = LD(ZKC_VM_SUPER_PC), sp = LD(ZKC_VM_SUB_PC);
        const VmCells<R> s0 = LD(ZKC_VM_SRC0_REG), s1 = LD(ZKC_VM_SRC1_REG), t0 = LD(ZKC_VM_DST0_REG), t1 = LD(ZKC_VM_DST1_REG);
        const VmCells<R> i0 = LD(ZKC_VM_IMM0), i1 = LD(ZKC_VM_IMM1);
        const VmCells<R> x0 = LD(ZKC_VM_SRC0_INDEX), x1 = LD(ZKC_VM_DST0_INDEX), x2 = LD(ZKC_VM_SP_AFTER_SRC0), x3 = LD(ZKC_VM_NEW_SP), x4 = LD(ZKC_VM_PC_OUT);
        FOR_R {
            if (sp.v[r] > 3 || super_pc.v[r] >> 14 || (s0.v[r] | s1.v[r] | t0.v[r] | t1.v[r]) > 15 || (i0.v[r] | i1.v[r]) >> 16 ||
                (x0.v[r] | x1.v[r] | x2.v[r] | x3.v[r] | x4.v[r]) >> 16)
                bad[r] |= ZKC_VMV_RANGE;
            sub_pc[r] = (uint32_t)sp.v[r];
            regs4[r] = (uint32_t)s0.v[r] | ((uint32_t)s1.v[r] << 8) | ((uint32_t)t0.v[r] << 16) | ((uint32_t)t1.v[r] << 24);
            imm[r] = ((uint32_t)i0.v[r] & 0xFFFFu) | ((uint32_t)i1.v[r] << 16);
        }
    }
    STAGE;
    // ---- decoding: the opcode word out of the code word, the table lookups, the exception masks --------------------------------
    uint64_t props[R];
    uint64_t limbs[R];
    FOR_R limbs[r] = 0;
    constexpr uint64_t MASK48 = (1ull << ZKC_VM_DESCRIPTION_BITS_FLATTENED) - 1;
    {
        uint32_t op_lo[R], op_hi[R];
        FOR_R { op_lo[r] = 0; op_hi[r] = 0; }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const VmCells<R> lo = LD(ZKC_VM_CODE_WORD + 6 - 2 * i), hi = LD(ZKC_VM_CODE_WORD + 7 - 2 * i);
            FOR_R {
                limbs[r] |= lo.v[r] | hi.v[r];
                if ((int)sub_pc[r] == i) { op_lo[r] = (uint32_t)lo.v[r]; op_hi[r] = (uint32_t)hi.v[r]; }
            }
        }
        const VmCells<R> opc0 = LD(ZKC_VM_OPCODE), opc1 = LD(ZKC_VM_OPCODE + 1), variant_c = LD(ZKC_VM_VARIANT), cond_idx = LD(ZKC_VM_CONDITION_IDX),
                         props_c = LD(ZKC_VM_PROPS), ergs_cost = LD(ZKC_VM_ERGS_COST), dirty_ergs = LD(ZKC_VM_DIRTY_ERGS_LEFT);
        FOR_R {
            limbs[r] |= ergs_cost.v[r] | dirty_ergs.v[r];
            if (B(VB_SKIP)) { op_lo[r] = (uint32_t)isa->nop_opcode_encoding; op_hi[r] = (uint32_t)(isa->nop_opcode_encoding >> 32); }
            if (B(VB_PENDING)) { op_lo[r] = (uint32_t)isa->panic_opcode_encoding; op_hi[r] = (uint32_t)(isa->panic_opcode_encoding >> 32); }
            if (opc0.v[r] != op_lo[r] || opc1.v[r] != op_hi[r]) bad[r] |= ZKC_VMV_DECODE;
            const uint32_t variant = op_lo[r] & 0x7FF;
            if (variant_c.v[r] != variant || cond_idx.v[r] != ((op_lo[r] >> 13) & 7) || imm[r] != op_hi[r]) bad[r] |= ZKC_VMV_DECODE;
            const uint64_t props_full = isa->opcode_props[variant];
            const uint32_t aux = (uint32_t)(props_full >> ZKC_VM_DESCRIPTION_BITS_FLATTENED);
            if (ergs_cost.v[r] != (B(VB_SKIP) ? 0u : isa->opcode_price[variant]) || B(VB_EXPL) != ((aux >> ZKC_VM_AUX_EXPLICIT_PANIC) & 1)) bad[r] |= ZKC_VMV_DECODE;
            if (B(VB_MPANIC) != (B(VB_EXPL) | B(VB_OUT_OF_ERGS) | B(VB_KEXC) | B(VB_SEXC) | B(VB_FULL)) || B(VB_MNOP) != (uint32_t)(!B(VB_MPANIC) && !B(VB_COND)))
                bad[r] |= ZKC_VMV_EXCEPTION_MASKS;
            if (B(VB_KEXC) && !((aux >> ZKC_VM_AUX_KERNEL_MODE) & 1)) bad[r] |= ZKC_VMV_EXCEPTION_MASKS;
            if (B(VB_SEXC) && ((aux >> ZKC_VM_AUX_CAN_BE_USED_IN_STATIC) & 1)) bad[r] |= ZKC_VMV_EXCEPTION_MASKS;
            if (B(VB_OUT_OF_ERGS) && dirty_ergs.v[r] != 0) bad[r] |= ZKC_VMV_EXCEPTION_MASKS;
            uint64_t pr = props_full & MASK48;
            if (B(VB_MPANIC)) pr = isa->panic_bitspread & MASK48;
            if (B(VB_MNOP)) pr = isa->nop_bitspread & MASK48;
            if (props_c.v[r] != pr) bad[r] |= ZKC_VMV_DECODE;
            props[r] = pr;
            uint32_t sd = op_lo[r] >> 16;  // src nibbles | dst nibbles << 8
            if (B(VB_MPANIC) || B(VB_MNOP)) sd = 0;
            if (regs4[r] != ((sd & 15) | (((sd >> 4) & 15) << 8) | (((sd >> 8) & 15) << 16) | ((sd >> 12) << 24))) bad[r] |= ZKC_VMV_DECODE;
        }
    }
    STAGE;
    // ---- u32 columns that are only range-checked -----------------------------------------------------------------------------------
    {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const VmCells<R> m = LD(ZKC_VM_SRC0_FROM_MEMORY + 1 + i);
            FOR_R limbs[r] |= m.v[r];
        }
        const VmCells<R> l0 = LD(ZKC_VM_ERGS_OUT), l1 = LD(ZKC_VM_SRC0_PAGE), l2 = LD(ZKC_VM_DST0_PAGE), l3 = LD(ZKC_VM_HEAP_BOUND_OUT),
                         l4 = LD(ZKC_VM_AUX_HEAP_BOUND_OUT), l5 = LD(ZKC_VM_MEMQ_LENGTH_OUT), l6 = LD(ZKC_VM_DEPTH_OUT);
        FOR_R limbs[r] |= l0.v[r] | l1.v[r] | l2.v[r] | l3.v[r] | l4.v[r] | l5.v[r] | l6.v[r];
    }
    STAGE;
    // ---- the arithmetic relations -------------------------------------------------------------------------------------------
#define TYPE(tt) prop_bit(props[r], ZKC_VM_BIT_TYPE(tt))
#define VAR(v) prop_bit(props[r], ZKC_VM_BIT_VARIANT(v))
    {
        uint32_t a[R][8], b[R][8], d0[R][8], d1[R][8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const VmCells<R> x = LD(ZKC_VM_SRC0 + 1 + i), y = LD(ZKC_VM_SRC1 + 1 + i), z = LD(ZKC_VM_DST0 + 1 + i), w = LD(ZKC_VM_DST1 + 1 + i);
            FOR_R {
                limbs[r] |= x.v[r] | y.v[r] | z.v[r] | w.v[r];
                a[r][i] = (uint32_t)x.v[r]; b[r][i] = (uint32_t)y.v[r]; d0[r][i] = (uint32_t)z.v[r]; d1[r][i] = (uint32_t)w.v[r];
            }
            if (i % 4 == 3) STAGE;  // 16 loads in flight at a time: the cells shrink to 32-bit limbs before the next batch
        }
        FOR_R {
            if (limbs[r] >> 32) bad[r] |= ZKC_VMV_RANGE;
            const bool set_flags = prop_bit(props[r], ZKC_VM_BIT_FLAG(ZKC_VM_SET_FLAGS_FLAG_IDX));
            const uint32_t f0 = B(VB_F0), f1 = B(VB_F1), f2 = B(VB_F2);
            bool b_zero = true, d0_zero = true, d1_zero = true;
#pragma unroll
            for (int i = 0; i < 8; i++) { d0_zero &= d0[r][i] == 0; d1_zero &= d1[r][i] == 0; b_zero &= b[r][i] == 0; }
            const bool is_sub = TYPE(ZKC_OP_SUB), is_div = TYPE(ZKC_OP_DIV);
            if (TYPE(ZKC_OP_ADD) || is_sub) {  // a + b = c + 2^256 * of (add), a = c + b - 2^256 * of (sub): enforce_addition_relation
                uint32_t carry = 0;
                bool ok = true;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint64_t sum = (uint64_t)(is_sub ? d0[r][i] : a[r][i]) + b[r][i] + carry;
                    ok &= (uint32_t)sum == (is_sub ? a[r][i] : d0[r][i]);
                    carry = (uint32_t)(sum >> 32);
                }
                if (!ok) bad[r] |= ZKC_VMV_ADD_SUB;
                if (set_flags && (f0 != carry || f1 != (uint32_t)d0_zero || f2 != (uint32_t)!(carry || d0_zero))) bad[r] |= ZKC_VMV_FLAGS;
            }
            if (TYPE(ZKC_OP_MUL) || is_div) {  // a * b + rem = lo + 2^256 * hi: enforce_mul_relation (8 x 8 u32 schoolbook)
                // mul: a * b = d0 + 2^256 d1.  div: d0 (quotient) * b + d1 (remainder) = a, remainder < b (b != 0); b == 0: both zero
                uint32_t x[8], p[16];
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = is_div ? d0[r][i] : a[r][i];
#pragma unroll
                for (int i = 0; i < 16; i++) p[i] = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    uint64_t carry = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint64_t tt = (uint64_t)x[i] * b[r][j] + p[i + j] + carry;
                        p[i + j] = (uint32_t)tt; carry = tt >> 32;
                    }
                    p[i + 8] = (uint32_t)carry;
                }
                bool ok = true;
                if (!is_div) {
#pragma unroll
                    for (int i = 0; i < 8; i++) ok &= p[i] == d0[r][i] && p[8 + i] == d1[r][i];
                    const bool of = !d1_zero;
                    if (set_flags && (f0 != (uint32_t)of || f1 != (uint32_t)d0_zero || f2 != (uint32_t)(!of && !d0_zero))) bad[r] |= ZKC_VMV_FLAGS;
                } else if (b_zero) {
                    ok = d0_zero && d1_zero;
                    if (set_flags && (f0 != 1 || f1 != 0 || f2 != 0)) bad[r] |= ZKC_VMV_FLAGS;
                } else {
                    uint64_t carry = 0;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint64_t sum = (uint64_t)p[i] + d1[r][i] + carry;
                        ok &= (uint32_t)sum == a[r][i];
                        carry = sum >> 32;
                        ok &= p[8 + i] == 0;
                    }
                    ok &= carry == 0;
                    bool lt = false, decided = false;  // remainder < divisor, MSW first
#pragma unroll
                    for (int i = 7; i >= 0; i--) if (!decided && d1[r][i] != b[r][i]) { lt = d1[r][i] < b[r][i]; decided = true; }
                    ok &= lt;
                    if (set_flags && (f0 != 0 || f1 != (uint32_t)d0_zero || f2 != (uint32_t)d1_zero)) bad[r] |= ZKC_VMV_FLAGS;
                }
                if (!ok) bad[r] |= ZKC_VMV_MUL_DIV;
            }
            if (TYPE(ZKC_OP_BINOP)) {
                const bool is_or = VAR(ZKC_VAR_BINOP_OR), is_and = VAR(ZKC_VAR_BINOP_AND);
                bool ok = true;
#pragma unroll
                for (int i = 0; i < 8; i++) ok &= d0[r][i] == (is_or ? (a[r][i] | b[r][i]) : (is_and ? (a[r][i] & b[r][i]) : (a[r][i] ^ b[r][i])));
                if (!ok) bad[r] |= ZKC_VMV_BINOP;
                if (set_flags && (f0 != 0 || f1 != (uint32_t)d0_zero || f2 != 0)) bad[r] |= ZKC_VMV_FLAGS;
            }
            // ---- selection: dst0 / dst1 are dot products of (flag, candidate) pairs; memory write needs a memory destination ----
            if (!B(VB_UPD0) && !B(VB_MEM_WRITE) && !(d0_zero && B(VB_D0_PTR) == 0)) bad[r] |= ZKC_VMV_SELECTION;
            if (!B(VB_UPD1) && !(d1_zero && B(VB_D1_PTR) == 0)) bad[r] |= ZKC_VMV_SELECTION;
            if (B(VB_MEM_WRITE) && !B(VB_DST0_MEM)) bad[r] |= ZKC_VMV_SELECTION;
            if (B(VB_UPD0) && B(VB_MEM_WRITE)) bad[r] |= ZKC_VMV_SELECTION;
        }
    }
#undef VAR
    STAGE;
    // ---- sponge columns: zeros unless enforced; the opcode-specific block is zero for the plain opcodes -------------------
    {
        uint64_t stray[R];
        FOR_R stray[r] = 0;
#pragma unroll
        for (int k = 0; k < ZKC_VM_NUM_SPONGES; k++) {  // one sponge slot (enforce flag + 12 outputs) per batch of loads
            const VmCells<R> enf = LD(ZKC_VM_SPONGE_ENFORCE + k);
            uint64_t any[R];
            FOR_R { any[r] = 0; if (enf.v[r] > 1) bad[r] |= ZKC_VMV_BOOLEAN; }
#pragma unroll
            for (int j = 0; j < 12; j++) {
                const VmCells<R> v = LD(ZKC_VM_SPONGE_FINAL + 12 * k + j);
                FOR_R { any[r] |= v.v[r]; if (v.v[r] >= ZKC_GL_P) bad[r] |= ZKC_VMV_RANGE; }
            }
            FOR_R {
                if (!enf.v[r]) stray[r] |= any[r];
                if (k == 0 && B(VB_READ_OP) != enf.v[r]) bad[r] |= ZKC_VMV_SPONGE;
                if (k == 1 && enf.v[r] < B(VB_READ_SRC0)) bad[r] |= ZKC_VMV_SPONGE;
                if (k == 2 && enf.v[r] < B(VB_MEM_WRITE)) bad[r] |= ZKC_VMV_SPONGE;
            }
            STAGE;  // 13 independent (128-bit) loads in flight per thread at a time
        }
        FOR_R if (stray[r]) bad[r] |= ZKC_VMV_SPONGE;
    }
    {
        uint64_t any[R];
        FOR_R any[r] = 0;
#pragma unroll
        for (int i = 0; i < ZKC_VM_OP_AUX_COLS; i++) {
            const VmCells<R> v = LD(ZKC_VM_OP_AUX + i);
            FOR_R any[r] |= v.v[r];
            if (i % 12 == 11) STAGE;
        }
        FOR_R {
            const bool family = TYPE(ZKC_OP_UMA) || TYPE(ZKC_OP_LOG) || TYPE(ZKC_OP_NEAR_CALL) || TYPE(ZKC_OP_FAR_CALL) || TYPE(ZKC_OP_RET);
            if (!family && any[r]) bad[r] |= ZKC_VMV_SELECTION;
        }
    }
#undef TYPE
    // the remaining columns are streamed too (every cell is read once): forward / rollback queue ends are field elements
    {
        uint64_t big[R];
        FOR_R big[r] = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const VmCells<R> x = LD(ZKC_VM_FORWARD_TAIL_OUT + i), y = LD(ZKC_VM_ROLLBACK_HEAD_OUT + i);
            FOR_R big[r] |= (uint64_t)(x.v[r] >= ZKC_GL_P) | (uint64_t)(y.v[r] >= ZKC_GL_P);
        }
        const VmCells<R> x = LD(ZKC_VM_FORWARD_TAIL_OUT + 4), y = LD(ZKC_VM_ROLLBACK_HEAD_OUT + 4);
        FOR_R if (big[r] || (x.v[r] | y.v[r]) >> 32) bad[r] |= ZKC_VMV_RANGE;
    }
#undef LD
#undef B
#undef STAGE
    FOR_R if (bad[r]) {
        atomicAdd(&out->violations, 1ull);
        atomicOr(&out->failed_checks, bad[r]);
        atomicMin(&out->first_bad, ((unsigned long long)(inst * limit + row + r) << 16) | (bad[r] & 0xFFFFu));
    }
#undef FOR_R
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
    // row pairs with 128-bit loads when every column of every instance starts 16-byte aligned
    if (limit % 2 == 0 && (reinterpret_cast<uintptr_t>(dt) & 15) == 0) {
        ZKC_LAUNCH(ctx, "vm_check", vm_check_kernel<2>, (unsigned)((rows / 2 + VM_CHECK_THREADS - 1) / VM_CHECK_THREADS), VM_CHECK_THREADS, 0, d, disa, dt,
                   limit, n_instances);
    } else {
        ZKC_LAUNCH(ctx, "vm_check", vm_check_kernel<1>, (unsigned)((rows + VM_CHECK_THREADS - 1) / VM_CHECK_THREADS), VM_CHECK_THREADS, 0, d, disa, dt, limit,
                   n_instances);
    }
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
