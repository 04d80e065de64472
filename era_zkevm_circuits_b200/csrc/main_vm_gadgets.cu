// Cells of the arithmetic opcode gadgets, evaluated obliviously (include/zkc_b200.h, ZKC_VM_GADGET_COLUMNS): one thread per cycle
// reads its src0 / src1 operands and property bits from the finished DENSE trace and writes the 724 cells the add/sub, binop,
// mul/div and shift gadgets of the reference allocate whatever the opcode, plus the relations vm_cycle enforces once per cycle:
//   RegisterInputView::from_input_value   /root/reference/src/main_vm/register_input_view.rs:27-53
//   apply_add_sub                         /root/reference/src/main_vm/opcodes/add_sub.rs:8-166
//   apply_binop, get_binop_subresults     /root/reference/src/main_vm/opcodes/binop.rs:14-244
//   apply_mul_div                         /root/reference/src/main_vm/opcodes/mul_div.rs:199-417
//   apply_shifts, get_shift_constant      /root/reference/src/main_vm/opcodes/shifts.rs:8-221
//   relation selection / enforcement      /root/reference/src/main_vm/cycle.rs:619-670, opcodes/mod.rs:101-180
// No lane diverges on the opcode (every gadget runs on every row, selections are predicated), the reads are 17 coalesced
// columns, the writes 724: an HBM-write-bound stream (5 792 B per cycle out).
#include "ctx.cuh"
#include "u256.cuh"

namespace zkc {

__device__ __forceinline__ U256 g_sel(bool flag, const U256 &a, const U256 &b) {  // UInt32::parallel_select
    U256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = flag ? a.v[i] : b.v[i];
    return r;
}

__global__ void __launch_bounds__(128)
vm_gadgets_kernel(const uint64_t *__restrict__ trace, size_t limit, size_t n_instances, uint64_t *__restrict__ out_all) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= limit * n_instances) return;
    const size_t inst = g / limit, row = g - inst * limit;
    const uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit + row;
    uint64_t *out = out_all + inst * (size_t)ZKC_VMG_NUM_COLS * limit + row;
#define G(col, i) out[(size_t)((col) + (i)) * limit]
#define PUT8(col, x) _Pragma("unroll") for (int i_ = 0; i_ < 8; i_++) G(col, i_) = (x).v[i_]
    const uint64_t props = __ldg(t + (size_t)ZKC_VM_PROPS * limit);
    U256 a, b, zero;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a.v[i] = (uint32_t)__ldg(t + (size_t)(ZKC_VM_SRC0 + 1 + i) * limit); b.v[i] = (uint32_t)__ldg(t + (size_t)(ZKC_VM_SRC1 + 1 + i) * limit);
        zero.v[i] = 0;
    }
#define BIT(n) (((props >> (n)) & 1) != 0)
    const bool set_flags = BIT(ZKC_VM_BIT_FLAG(ZKC_VM_SET_FLAGS_FLAG_IDX));
#pragma unroll
    for (int i = 0; i < 32; i++) { G(ZKC_VMG_SRC0_BYTES, i) = (a.v[i / 4] >> (8 * (i % 4))) & 0xFF; G(ZKC_VMG_SRC1_BYTES, i) = (b.v[i / 4] >> (8 * (i % 4))) & 0xFF; }

    // ---- add_sub.rs ----------------------------------------------------------------------------------------------------
    U256 add_r, sub_r;
    const uint32_t add_of = u256_add(a, b, add_r), sub_uf = u256_sub(a, b, sub_r);
    const bool apply_add = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_ADD)), apply_sub = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_SUB)), as_any = apply_add || apply_sub;
    const U256 as_result = g_sel(apply_add, add_r, sub_r), new_b = g_sel(apply_add, a, sub_r), new_c = g_sel(apply_add, add_r, a);
    const uint32_t new_of = apply_add ? add_of : sub_uf;
    bool as_zero = true;
#pragma unroll
    for (int i = 0; i < 8; i++) { G(ZKC_VMG_ADDSUB_LIMB_IS_ZERO, i) = as_result.v[i] == 0; as_zero &= as_result.v[i] == 0; }
    PUT8(ZKC_VMG_ADD_RESULT, add_r); G(ZKC_VMG_ADD_OF, 0) = add_of; PUT8(ZKC_VMG_SUB_RESULT, sub_r); G(ZKC_VMG_SUB_UF, 0) = sub_uf;
    PUT8(ZKC_VMG_ADDSUB_RESULT, as_result); PUT8(ZKC_VMG_ADDSUB_NEW_B, new_b); PUT8(ZKC_VMG_ADDSUB_NEW_C, new_c);
    G(ZKC_VMG_ADDSUB_NEW_OF, 0) = new_of; G(ZKC_VMG_ADDSUB_RESULT_IS_ZERO, 0) = as_zero; G(ZKC_VMG_ADDSUB_GT, 0) = !(new_of || as_zero);
    G(ZKC_VMG_ADDSUB_APPLY_ANY, 0) = as_any; G(ZKC_VMG_ADDSUB_UPDATE_FLAGS, 0) = as_any && set_flags;

    // ---- binop.rs --------------------------------------------------------------------------------------------------------
    {
        U256 and_c, or_c, xor_c;
#pragma unroll
        for (int i = 0; i < 8; i++) { and_c.v[i] = a.v[i] & b.v[i]; or_c.v[i] = a.v[i] | b.v[i]; xor_c.v[i] = a.v[i] ^ b.v[i]; }
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const uint64_t an = (and_c.v[i / 4] >> (8 * (i % 4))) & 0xFF, orr = (or_c.v[i / 4] >> (8 * (i % 4))) & 0xFF, xo = (xor_c.v[i / 4] >> (8 * (i % 4))) & 0xFF;
            G(ZKC_VMG_BINOP_COMPOSITE, i) = an | (orr << 16) | (xo << 32);
            G(ZKC_VMG_BINOP_ALL_RESULTS, 3 * i) = an; G(ZKC_VMG_BINOP_ALL_RESULTS, 3 * i + 1) = orr; G(ZKC_VMG_BINOP_ALL_RESULTS, 3 * i + 2) = xo;
        }
        const bool is_and = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_BINOP_AND)), is_or = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_BINOP_OR));
        const U256 res = g_sel(is_or, or_c, g_sel(is_and, and_c, xor_c));
        bool z = true;
#pragma unroll
        for (int i = 0; i < 8; i++) { G(ZKC_VMG_BINOP_LIMB_IS_ZERO, i) = res.v[i] == 0; z &= res.v[i] == 0; }
        PUT8(ZKC_VMG_BINOP_AND, and_c); PUT8(ZKC_VMG_BINOP_OR, or_c); PUT8(ZKC_VMG_BINOP_XOR, xor_c); PUT8(ZKC_VMG_BINOP_RESULT, res);
        G(ZKC_VMG_BINOP_RESULT_IS_ZERO, 0) = z; G(ZKC_VMG_BINOP_UPDATE_FLAGS, 0) = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_BINOP)) && set_flags;
    }

    // ---- mul_div.rs ------------------------------------------------------------------------------------------------------
    const bool apply_mul = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_MUL)), apply_div = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_DIV)), md_any = apply_mul || apply_div;
    U256 mul_lo, mul_hi, quot, rem, div_sub;
    u256_mul(a, b, mul_lo, mul_hi);
    const bool divisor_z = u256_is_zero(b);
    if (divisor_z) { quot = zero; rem = a; }  // mul_div.rs:119-123
    else u256_divrem(a, b, quot, rem);
    const U256 md_rem = g_sel(apply_mul, zero, rem), md_a = g_sel(apply_mul, a, quot), md_low = g_sel(apply_mul, mul_lo, a),
               md_high = g_sel(apply_mul, mul_hi, zero);
    const uint32_t div_less = u256_sub(rem, b, div_sub);
    {
        const U256 r0 = g_sel(apply_mul, mul_lo, quot), r1 = g_sel(apply_mul, mul_hi, rem);
        const bool high_z = u256_is_zero(mul_hi), low_z = u256_is_zero(mul_lo), of_mul = !high_z, eq_mul = low_z, gt_mul = !of_mul && !eq_mul;
        const bool quot_z = u256_is_zero(quot), rem_z = u256_is_zero(rem), mask = apply_div && divisor_z;
        const bool of_div = divisor_z, eq_div = !divisor_z && quot_z, gt_div = !divisor_z && rem_z;
        PUT8(ZKC_VMG_MUL_LOW, mul_lo); PUT8(ZKC_VMG_MUL_HIGH, mul_hi); PUT8(ZKC_VMG_DIV_QUOTIENT, quot); PUT8(ZKC_VMG_DIV_REMAINDER, rem);
        PUT8(ZKC_VMG_MULDIV_RESULT_0, r0); PUT8(ZKC_VMG_MULDIV_RESULT_1_UNMASKED, r1);
        PUT8(ZKC_VMG_MULDIV_REM_TO_ENFORCE, md_rem); PUT8(ZKC_VMG_MULDIV_A_TO_ENFORCE, md_a);
        PUT8(ZKC_VMG_MULDIV_MUL_LOW_TO_ENFORCE, md_low); PUT8(ZKC_VMG_MULDIV_MUL_HIGH_TO_ENFORCE, md_high);
        G(ZKC_VMG_MUL_HIGH_IS_ZERO, 0) = high_z; G(ZKC_VMG_MUL_LOW_IS_ZERO, 0) = low_z; G(ZKC_VMG_MUL_OF, 0) = of_mul; G(ZKC_VMG_MUL_GT, 0) = gt_mul;
        G(ZKC_VMG_DIV_DIVISOR_IS_ZERO, 0) = divisor_z; G(ZKC_VMG_DIV_QUOTIENT_IS_ZERO, 0) = quot_z; G(ZKC_VMG_DIV_REMAINDER_IS_ZERO, 0) = rem_z;
        PUT8(ZKC_VMG_DIV_SUB_RESULT, div_sub); G(ZKC_VMG_DIV_REMAINDER_IS_LESS, 0) = div_less; G(ZKC_VMG_DIV_MASK_REMAINDER, 0) = mask;
#pragma unroll
        for (int i = 0; i < 8; i++) G(ZKC_VMG_MULDIV_RESULT_1, i) = mask ? 0u : r1.v[i];
        G(ZKC_VMG_DIV_EQ, 0) = eq_div; G(ZKC_VMG_DIV_GT, 0) = gt_div;
        G(ZKC_VMG_MULDIV_OF, 0) = apply_mul ? of_mul : of_div; G(ZKC_VMG_MULDIV_EQ, 0) = apply_mul ? eq_mul : eq_div;
        G(ZKC_VMG_MULDIV_GT, 0) = apply_mul ? gt_mul : gt_div;
        G(ZKC_VMG_MULDIV_APPLY_ANY, 0) = md_any; G(ZKC_VMG_MULDIV_SET_FLAGS, 0) = md_any && set_flags;
    }

    // ---- shifts.rs: the divisor / multiplier is 2^full_shift, so both "unchecked results" are shifts --------------------------
    const bool apply_shift = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_SHIFT));
    U256 shc, rq, rr, ll, lh, sh_sub;
    uint32_t sh_less;
    bool apply_left;
    {
        const bool is_rol = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_SHIFT_ROL)), is_ror = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_SHIFT_ROR)),
                   is_shr = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_SHIFT_SHR));
        const bool is_cyclic = is_rol || is_ror, is_right = is_ror || is_shr;
        const uint32_t shift = b.v[0] & 0xFF;
        const bool shift_z = shift == 0;
        const uint32_t inverted = 256 - shift;
        const bool change = is_ror && !shift_z;
        const uint32_t full = change ? inverted : shift;
#pragma unroll
        for (int i = 0; i < 8; i++) shc.v[i] = (full >> 5) == (uint32_t)i ? 1u << (full & 31) : 0u;  // tables/bitshift.rs:22-33
        const bool is_right_shift = is_right && !is_cyclic;
        apply_left = apply_shift && !is_right_shift;
        u256_shr(a, full, rq);                                   // a / 2^full
        u256_shl_wide(a, full, ll, lh);                          // a * 2^full
        {   // a mod 2^full = a - (rq << full)
            U256 back, hi_unused;
            u256_shl_wide(rq, full, back, hi_unused);
            u256_sub(a, back, rr);
        }
        sh_less = u256_sub(rr, shc, sh_sub);
        const U256 sh_rem = g_sel(apply_left, zero, rr), sh_a = g_sel(apply_left, a, rq), sh_low = g_sel(apply_left, ll, a),
                   sh_high = g_sel(apply_left, lh, zero), temp = g_sel(is_right_shift, rq, ll);
        U256 fin;
#pragma unroll
        for (int i = 0; i < 8; i++) fin.v[i] = (is_cyclic ? lh.v[i] : 0u) + temp.v[i];
        G(ZKC_VMG_SHIFT_AMOUNT, 0) = shift; G(ZKC_VMG_SHIFT_IS_ZERO, 0) = shift_z; G(ZKC_VMG_SHIFT_INVERTED, 0) = inverted;
        G(ZKC_VMG_SHIFT_CHANGE_FLAG, 0) = change; G(ZKC_VMG_SHIFT_FULL, 0) = full; PUT8(ZKC_VMG_SHIFT_CONSTANT, shc);
        G(ZKC_VMG_SHIFT_IS_RIGHT, 0) = is_right_shift; PUT8(ZKC_VMG_SHIFT_RSHIFT_Q, rq); PUT8(ZKC_VMG_SHIFT_RSHIFT_R, rr);
        G(ZKC_VMG_SHIFT_APPLY_LEFT, 0) = apply_left; PUT8(ZKC_VMG_SHIFT_LSHIFT_LOW, ll); PUT8(ZKC_VMG_SHIFT_LSHIFT_HIGH, lh);
        PUT8(ZKC_VMG_SHIFT_REM_TO_ENFORCE, sh_rem); PUT8(ZKC_VMG_SHIFT_A_TO_ENFORCE, sh_a);
        PUT8(ZKC_VMG_SHIFT_MUL_LOW_TO_ENFORCE, sh_low); PUT8(ZKC_VMG_SHIFT_MUL_HIGH_TO_ENFORCE, sh_high);
        PUT8(ZKC_VMG_SHIFT_SUB_RESULT, sh_sub); G(ZKC_VMG_SHIFT_REMAINDER_IS_LESS, 0) = sh_less;
        PUT8(ZKC_VMG_SHIFT_TEMP_RESULT, temp); PUT8(ZKC_VMG_SHIFT_RESULT, fin);
        G(ZKC_VMG_SHIFT_RESULT_IS_ZERO, 0) = u256_is_zero(fin); G(ZKC_VMG_SHIFT_SET_FLAGS, 0) = apply_shift && set_flags;
    }

    // ---- cycle.rs:619-670: candidates in push order add_sub, mul_div, shifts; the last pushed is the default -----------------
    {
        const U256 rc = g_sel(md_any, div_sub, g_sel(as_any, as_result, sh_sub));
        PUT8(ZKC_VMG_RANGE_CHECK, rc);
        const U256 ra = g_sel(md_any || as_any, b, shc), rb = g_sel(md_any, div_sub, g_sel(as_any, new_b, sh_sub)),
                   rcc = g_sel(md_any, rem, g_sel(as_any, new_c, rr));
        const uint32_t of = md_any ? div_less : (as_any ? new_of : sh_less);
        PUT8(ZKC_VMG_ADDREL_A, ra); PUT8(ZKC_VMG_ADDREL_B, rb); PUT8(ZKC_VMG_ADDREL_C, rcc); G(ZKC_VMG_ADDREL_OF, 0) = of;
        uint64_t carry = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { carry = ((uint64_t)ra.v[i] + rb.v[i] + carry) >> 32; G(ZKC_VMG_ADDREL_CARRY, i) = carry; }
        const U256 ma = g_sel(md_any, md_a, g_sel(apply_left, a, rq)), mb = g_sel(md_any, b, shc),
                   mrem = g_sel(md_any, md_rem, g_sel(apply_left, zero, rr)), mlow = g_sel(md_any, md_low, g_sel(apply_left, ll, a)),
                   mhigh = g_sel(md_any, md_high, g_sel(apply_left, lh, zero));
        PUT8(ZKC_VMG_MULREL_A, ma); PUT8(ZKC_VMG_MULREL_B, mb); PUT8(ZKC_VMG_MULREL_REM, mrem); PUT8(ZKC_VMG_MULREL_LOW, mlow); PUT8(ZKC_VMG_MULREL_HIGH, mhigh);
        uint32_t partial[16];
#pragma unroll
        for (int i = 0; i < 16; i++) partial[i] = i < 8 ? mrem.v[i] : 0u;
#pragma unroll
        for (int ai = 0; ai < 8; ai++) {
            uint32_t overflow = 0;
#pragma unroll
            for (int bi = 0; bi < 8; bi++) {
                const uint64_t p = (uint64_t)ma.v[ai] * mb.v[bi] + partial[ai + bi] + overflow;
                partial[ai + bi] = (uint32_t)p; overflow = (uint32_t)(p >> 32);
                G(ZKC_VMG_MULREL_PARTIAL_LOW, 8 * ai + bi) = (uint32_t)p; G(ZKC_VMG_MULREL_PARTIAL_HIGH, 8 * ai + bi) = overflow;
            }
            partial[ai + 8] += overflow;
            G(ZKC_VMG_MULREL_ROW_END, ai) = partial[ai + 8];
        }
    }
#undef BIT
#undef G
#undef PUT8
}


// ---- the second block (ZKC_VM_STATE_GADGET_COLUMNS): ptr, jump and context gadgets, which read the state the cycle starts from ------
//   apply_ptr      /root/reference/src/main_vm/opcodes/ptr.rs:6-183
//   apply_jump     /root/reference/src/main_vm/opcodes/jump.rs:3-38
//   apply_context  /root/reference/src/main_vm/opcodes/context.rs:7-307
// One thread per cycle: 21 coalesced trace columns + 27 words of its snapshot record in, 87 columns out.
__global__ void __launch_bounds__(128)
vm_state_gadgets_kernel(const uint64_t *__restrict__ trace, const zkc_vm_state *__restrict__ snapshots, size_t limit, size_t n_instances,
                        uint64_t *__restrict__ out_all) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= limit * n_instances) return;
    const size_t inst = g / limit, row = g - inst * limit;
    const uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit + row;
    const zkc_vm_state *st = snapshots + inst * (limit + 1) + row;
    uint64_t *out = out_all + inst * (size_t)ZKC_VMS_NUM_COLS * limit + row;
#define S(col, i) out[(size_t)((col) + (i)) * limit]
    const uint64_t props = __ldg(t + (size_t)ZKC_VM_PROPS * limit);
#define BIT(n) (((props >> (n)) & 1) != 0)
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = (uint32_t)__ldg(t + (size_t)(ZKC_VM_SRC0 + 1 + i) * limit); b[i] = (uint32_t)__ldg(t + (size_t)(ZKC_VM_SRC1 + 1 + i) * limit);
    }
    const bool a_ptr = __ldg(t + (size_t)ZKC_VM_SRC0 * limit) != 0, b_ptr = __ldg(t + (size_t)ZKC_VM_SRC1 * limit) != 0;

    // ---- ptr.rs ------------------------------------------------------------------------------------------------------------
    {
        const bool should_apply = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_PTR));
        const bool v_add = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_PTR_ADD)), v_sub = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_PTR_SUB)),
                   v_pack = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_PTR_PACK)), v_shrink = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_PTR_SHRINK));
        const bool src1_is_integer = !b_ptr, args_valid = a_ptr && src1_is_integer, args_invalid = !args_valid;
        bool hi_zero = true, lo_zero = true;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const bool z = b[i] == 0;
            S(ZKC_VMS_PTR_SRC1_LIMB_IS_ZERO, i) = z;
            if (i >= 1) hi_zero &= z;
            if (i < 4) lo_zero &= z;
        }
        const bool arith = v_add || v_sub, too_large = !hi_zero && arith, dirty_pack = !lo_zero && v_pack;
        const uint32_t add_r = a[0] + b[0], sub_r = a[0] - b[0], shr_r = a[3] - b[0];
        const bool add_of = add_r < a[0], sub_uf = a[0] < b[0], shr_uf = a[3] < b[0];
        const bool add_panic = v_add && add_of, sub_panic = v_sub && sub_uf, shr_panic = v_shrink && shr_uf;
        const bool any_panic = args_invalid || too_large || dirty_pack || add_panic || sub_panic || shr_panic;
        const uint32_t low_if_add = v_add ? add_r : a[0], low_if_add_or_sub = v_sub ? sub_r : low_if_add, b96_if_shrink = v_shrink ? shr_r : a[3];
        const uint32_t lowest32 = v_pack ? a[0] : low_if_add_or_sub, b96 = v_pack ? a[3] : b96_if_shrink;
        S(ZKC_VMS_PTR_SRC1_IS_INTEGER, 0) = src1_is_integer; S(ZKC_VMS_PTR_ARGS_VALID, 0) = args_valid; S(ZKC_VMS_PTR_ARGS_INVALID, 0) = args_invalid;
        S(ZKC_VMS_PTR_SRC1_32_256_IS_ZERO, 0) = hi_zero; S(ZKC_VMS_PTR_SRC1_0_128_IS_ZERO, 0) = lo_zero; S(ZKC_VMS_PTR_SRC1_32_256_IS_NONZERO, 0) = !hi_zero;
        S(ZKC_VMS_PTR_ARITH_VARIANT, 0) = arith; S(ZKC_VMS_PTR_TOO_LARGE_OFFSET, 0) = too_large; S(ZKC_VMS_PTR_SRC1_0_128_IS_NONZERO, 0) = !lo_zero;
        S(ZKC_VMS_PTR_DIRTY_PACK, 0) = dirty_pack; S(ZKC_VMS_PTR_ADD_RESULT, 0) = add_r; S(ZKC_VMS_PTR_ADD_OF, 0) = add_of; S(ZKC_VMS_PTR_ADD_PANIC, 0) = add_panic;
        S(ZKC_VMS_PTR_SUB_RESULT, 0) = sub_r; S(ZKC_VMS_PTR_SUB_UF, 0) = sub_uf; S(ZKC_VMS_PTR_SUB_PANIC, 0) = sub_panic;
        S(ZKC_VMS_PTR_SHRINK_RESULT, 0) = shr_r; S(ZKC_VMS_PTR_SHRINK_UF, 0) = shr_uf; S(ZKC_VMS_PTR_SHRINK_PANIC, 0) = shr_panic;
        S(ZKC_VMS_PTR_ANY_PANIC, 0) = any_panic; S(ZKC_VMS_PTR_SHOULD_PANIC, 0) = should_apply && any_panic; S(ZKC_VMS_PTR_OK, 0) = !any_panic;
        S(ZKC_VMS_PTR_UPDATE_REGISTER, 0) = should_apply && !any_panic;
        S(ZKC_VMS_PTR_LOW_IF_ADD, 0) = low_if_add; S(ZKC_VMS_PTR_LOW_IF_ADD_OR_SUB, 0) = low_if_add_or_sub; S(ZKC_VMS_PTR_96_128_IF_SHRINK, 0) = b96_if_shrink;
        S(ZKC_VMS_PTR_LOWEST32, 0) = lowest32; S(ZKC_VMS_PTR_96_128, 0) = b96;
        S(ZKC_VMS_PTR_DST0, 0) = a_ptr; S(ZKC_VMS_PTR_DST0, 1) = lowest32; S(ZKC_VMS_PTR_DST0, 2) = a[1]; S(ZKC_VMS_PTR_DST0, 3) = a[2]; S(ZKC_VMS_PTR_DST0, 4) = b96;
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint32_t h = v_pack ? b[4 + i] : a[4 + i]; S(ZKC_VMS_PTR_HIGHEST_128, i) = h; S(ZKC_VMS_PTR_DST0, 5 + i) = h; }
    }

    // ---- jump.rs ---------------------------------------------------------------------------------------------------------
    S(ZKC_VMS_JUMP_DST, 0) = a[0] & 0xFFFFu;

    // ---- context.rs ------------------------------------------------------------------------------------------------------
    {
        const zkc_vm_context *c = &st->current_context;
        const bool should_apply = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_CONTEXT));
        const bool is_this = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_THIS)), is_caller = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_CALLER)),
                   is_code = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_CODE_ADDRESS)), is_meta = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_META)),
                   is_ergs = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_ERGS_LEFT)), is_get_u128 = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_GET_U128)),
                   is_set_u128 = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_SET_U128)),
                   is_set_ergs = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_SET_ERGS_PER_PUBDATA)),
                   is_inc_tx = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_INC_TX_NUMBER));
        const bool read_only = is_set_u128 || is_set_ergs || is_inc_tx;
        S(ZKC_VMS_CTX_WRITE_TO_CONTEXT, 0) = should_apply && is_set_u128; S(ZKC_VMS_CTX_SET_PUBDATA_ERGS, 0) = should_apply && is_set_ergs;
        S(ZKC_VMS_CTX_INCREMENT_TX, 0) = should_apply && is_inc_tx; S(ZKC_VMS_CTX_READ_ONLY, 0) = read_only; S(ZKC_VMS_CTX_WRITE_LIKE, 0) = !read_only;
        S(ZKC_VMS_CTX_WRITE_TO_DST0, 0) = should_apply && !read_only;
        const uint32_t tx = __ldg(&st->tx_number_in_block);
        S(ZKC_VMS_CTX_INCREMENTED_TX_NUMBER, 0) = tx + 1u; S(ZKC_VMS_CTX_TX_OF, 0) = tx == 0xFFFFFFFFu;
        const uint32_t meta_hi = (__ldg(&c->this_shard_id) & 0xFFu) | (__ldg(&c->caller_shard_id) & 0xFFu) << 8 | (__ldg(&c->code_shard_id) & 0xFFu) << 16;
        S(ZKC_VMS_CTX_META_HIGHEST, 0) = meta_hi;
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = 0;
        r[0] = is_ergs ? (uint32_t)__ldg(t + (size_t)ZKC_VM_DIRTY_ERGS_LEFT * limit) : (uint32_t)__ldg(t + (size_t)ZKC_VM_NEW_SP * limit);
        S(ZKC_VMS_CTX_LOW_U32, 0) = r[0];
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint32_t v = __ldg(&c->context_u128_value_composite[i]); if (is_get_u128) r[i] = v; S(ZKC_VMS_CTX_RESULT_128, i) = r[i]; }
#pragma unroll
        for (int i = 0; i < 5; i++) { const uint32_t v = __ldg(&c->this_address[i]); if (is_this) r[i] = v; S(ZKC_VMS_CTX_RESULT_160_THIS, i) = r[i]; }
#pragma unroll
        for (int i = 0; i < 5; i++) { const uint32_t v = __ldg(&c->caller[i]); if (is_caller) r[i] = v; S(ZKC_VMS_CTX_RESULT_160_CALLER, i) = r[i]; }
#pragma unroll
        for (int i = 0; i < 5; i++) { const uint32_t v = __ldg(&c->code_address[i]); if (is_code) r[i] = v; S(ZKC_VMS_CTX_RESULT_160_CODE, i) = r[i]; }
        const uint32_t meta[8] = {__ldg(&st->ergs_per_pubdata_byte), 0u, __ldg(&c->heap_upper_bound), __ldg(&c->aux_heap_upper_bound), 0u, 0u, 0u, meta_hi};
#pragma unroll
        for (int i = 0; i < 8; i++) S(ZKC_VMS_CTX_RESULT_256, i) = is_meta ? meta[i] : r[i];
    }
#undef BIT
#undef S
}

}  // namespace zkc

using namespace zkc;

extern "C" int zkc_main_vm_gadget_cells(zkc_ctx *ctx, const uint64_t *trace, size_t limit, size_t n_instances, int on_device, uint64_t *gadget_trace) {
    if (!ctx || ((limit * n_instances) && (!trace || !gadget_trace))) return ZKC_ERR_INVALID_ARGUMENT;
    const size_t rows = limit * n_instances;
    if (!rows) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t *dt = trace;
    uint64_t *dg = gadget_trace;
    if (!on_device) {
        char *blk = (char *)ctx->scratch(zkc_carver::bytes(rows * ZKC_VM_NUM_COLS, 8) + zkc_carver::bytes(rows * ZKC_VMG_NUM_COLS, 8));
        if (!blk) return ZKC_ERR_CUDA;
        zkc_carver cv(blk);
        uint64_t *bt = cv.take<uint64_t>(rows * ZKC_VM_NUM_COLS);
        dg = cv.take<uint64_t>(rows * ZKC_VMG_NUM_COLS);
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(bt, trace, rows * ZKC_VM_NUM_COLS * 8, cudaMemcpyHostToDevice, s));
        dt = bt;
    }
    ZKC_LAUNCH(ctx, "vm_gadgets", vm_gadgets_kernel, (unsigned)((rows + 127) / 128), 128, 0, dt, limit, n_instances, dg);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    if (!on_device) ZKC_CUDA(ctx, st, cudaMemcpyAsync(gadget_trace, dg, rows * ZKC_VMG_NUM_COLS * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    return ZKC_OK;
}

extern "C" int zkc_main_vm_state_gadget_cells(zkc_ctx *ctx, const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances,
                                              int on_device, uint64_t *gadget_trace) {
    if (!ctx || ((limit * n_instances) && (!trace || !snapshots || !gadget_trace))) return ZKC_ERR_INVALID_ARGUMENT;
    const size_t rows = limit * n_instances, n_snaps = (limit + 1) * n_instances;
    if (!rows) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t *dt = trace;
    const zkc_vm_state *ds = snapshots;
    uint64_t *dg = gadget_trace;
    if (!on_device) {
        char *blk = (char *)ctx->scratch(zkc_carver::bytes(rows * ZKC_VM_NUM_COLS, 8) + zkc_carver::bytes(n_snaps, sizeof(zkc_vm_state)) +
                                         zkc_carver::bytes(rows * ZKC_VMS_NUM_COLS, 8));
        if (!blk) return ZKC_ERR_CUDA;
        zkc_carver cv(blk);
        uint64_t *bt = cv.take<uint64_t>(rows * ZKC_VM_NUM_COLS);
        zkc_vm_state *bs = cv.take<zkc_vm_state>(n_snaps);
        dg = cv.take<uint64_t>(rows * ZKC_VMS_NUM_COLS);
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(bt, trace, rows * ZKC_VM_NUM_COLS * 8, cudaMemcpyHostToDevice, s));
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(bs, snapshots, n_snaps * sizeof(zkc_vm_state), cudaMemcpyHostToDevice, s));
        dt = bt; ds = bs;
    }
    ZKC_LAUNCH(ctx, "vm_state_gadgets", vm_state_gadgets_kernel, (unsigned)((rows + 127) / 128), 128, 0, dt, ds, limit, n_instances, dg);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    if (!on_device) ZKC_CUDA(ctx, st, cudaMemcpyAsync(gadget_trace, dg, rows * ZKC_VMS_NUM_COLS * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    return ZKC_OK;
}
