/* ORACLE -- TEST INFRASTRUCTURE ONLY (see gl.h header).
 *
 * The cells the arithmetic opcode gadgets of the reference allocate on EVERY cycle, whatever the opcode, and the relations
 * vm_cycle enforces once per cycle (one cycle at a time, in the order the reference evaluates them):
 *   RegisterInputView::from_input_value   /root/reference/src/main_vm/register_input_view.rs:27-53
 *   apply_add_sub                         /root/reference/src/main_vm/opcodes/add_sub.rs:8-166 (+ :168-282 the unchecked results)
 *   apply_binop, get_binop_subresults     /root/reference/src/main_vm/opcodes/binop.rs:14-244
 *   apply_mul_div                         /root/reference/src/main_vm/opcodes/mul_div.rs:199-417 (+ :20-163 the unchecked results)
 *   apply_shifts, get_shift_constant      /root/reference/src/main_vm/opcodes/shifts.rs:8-221, /root/reference/src/tables/bitshift.rs
 *   relation selection                    /root/reference/src/main_vm/cycle.rs:619-670
 *   enforce_addition_relation / enforce_mul_relation   /root/reference/src/main_vm/opcodes/mod.rs:101-180
 * Column meaning: include/zkc_b200.h, ZKC_VM_GADGET_COLUMNS.  Pinning: PARITY UNPINNED against the reference (no main_vm vector
 * in it); the arithmetic itself is checked against Python integers (tests/test_oracle_main_vm_gadgets.py).
 */
#include "oracle.h"
#include <string.h>

static int add256(const uint32_t *a, const uint32_t *b, uint32_t *c) {  /* add_sub.rs:181-194 */
    uint64_t carry = 0;
    for (int i = 0; i < 8; i++) { const uint64_t s = (uint64_t)a[i] + b[i] + carry; c[i] = (uint32_t)s; carry = s >> 32; }
    return (int)carry;
}
static int sub256(const uint32_t *a, const uint32_t *b, uint32_t *c) {  /* add_sub.rs:239-252 */
    uint64_t borrow = 0;
    for (int i = 0; i < 8; i++) { const uint64_t d = (uint64_t)a[i] - b[i] - borrow; c[i] = (uint32_t)d; borrow = (d >> 32) & 1; }
    return (int)borrow;
}
static void mul256(const uint32_t *a, const uint32_t *b, uint32_t *lo, uint32_t *hi) {  /* U256::full_mul */
    uint32_t r[16];
    memset(r, 0, sizeof r);
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 8; j++) { const uint64_t t = (uint64_t)a[i] * b[j] + r[i + j] + carry; r[i + j] = (uint32_t)t; carry = t >> 32; }
        r[i + 8] = (uint32_t)carry;
    }
    memcpy(lo, r, 32); memcpy(hi, r + 8, 32);
}
static int is_zero256(const uint32_t *a) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a[i]; return o == 0; }
static int ge256(const uint32_t *a, const uint32_t *b) {
    for (int i = 7; i >= 0; i--) if (a[i] != b[i]) return a[i] > b[i];
    return 1;
}
/* mul_div.rs:110-123: (0, a) for a zero divisor; schoolbook binary long division otherwise */
static void divrem256(const uint32_t *a, const uint32_t *b, uint32_t *q, uint32_t *r) {
    memset(q, 0, 32); memset(r, 0, 32);
    if (is_zero256(b)) { memcpy(r, a, 32); return; }
    for (int bit = 255; bit >= 0; bit--) {
        uint32_t top = r[7] >> 31;
        for (int i = 7; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 31);
        r[0] = (r[0] << 1) | ((a[bit >> 5] >> (bit & 31)) & 1);
        if (top || ge256(r, b)) { uint32_t t[8]; sub256(r, b, t); memcpy(r, t, 32); q[bit >> 5] |= 1u << (bit & 31); }
    }
}
static void sel8(int flag, const uint32_t *a, const uint32_t *b, uint32_t *out) { memcpy(out, flag ? a : b, 32); }  /* UInt32::parallel_select */

#define G(col, i) out[(size_t)((col) + (i)) * limit + row]
static void put8(uint64_t *out, size_t limit, size_t row, int col, const uint32_t *v) { for (int i = 0; i < 8; i++) G(col, i) = v[i]; }

static void gadget_row(uint64_t props, const uint32_t *a, const uint32_t *b, uint64_t *out, size_t limit, size_t row) {
#define BIT(n) (int)((props >> (n)) & 1)
    static const uint32_t zero8[8] = {0};
    const int set_flags = BIT(ZKC_VM_BIT_FLAG(ZKC_VM_SET_FLAGS_FLAG_IDX));
    /* register_input_view.rs:36-46 */
    for (int i = 0; i < 32; i++) { G(ZKC_VMG_SRC0_BYTES, i) = (a[i / 4] >> (8 * (i % 4))) & 0xFF; G(ZKC_VMG_SRC1_BYTES, i) = (b[i / 4] >> (8 * (i % 4))) & 0xFF; }

    /* ---- add_sub.rs ------------------------------------------------------------------------------------------------ */
    uint32_t add_r[8], sub_r[8], as_result[8], new_b[8], new_c[8];
    const int add_of = add256(a, b, add_r), sub_uf = sub256(a, b, sub_r);
    const int apply_add = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_ADD)), apply_sub = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_SUB));
    sel8(apply_add, add_r, sub_r, as_result);          /* :51-56 */
    sel8(apply_add, a, sub_r, new_b);                  /* :91-96: relation a = src1, b, c, of */
    sel8(apply_add, add_r, a, new_c);                  /* :98-103 */
    const int new_of = apply_add ? add_of : sub_uf;
    int as_zero = 1;
    for (int i = 0; i < 8; i++) { G(ZKC_VMG_ADDSUB_LIMB_IS_ZERO, i) = as_result[i] == 0; as_zero &= as_result[i] == 0; }
    const int as_gt = !(new_of || as_zero), as_any = apply_add || apply_sub;
    put8(out, limit, row, ZKC_VMG_ADD_RESULT, add_r); G(ZKC_VMG_ADD_OF, 0) = add_of;
    put8(out, limit, row, ZKC_VMG_SUB_RESULT, sub_r); G(ZKC_VMG_SUB_UF, 0) = sub_uf;
    put8(out, limit, row, ZKC_VMG_ADDSUB_RESULT, as_result); put8(out, limit, row, ZKC_VMG_ADDSUB_NEW_B, new_b);
    put8(out, limit, row, ZKC_VMG_ADDSUB_NEW_C, new_c);
    G(ZKC_VMG_ADDSUB_NEW_OF, 0) = new_of; G(ZKC_VMG_ADDSUB_RESULT_IS_ZERO, 0) = as_zero; G(ZKC_VMG_ADDSUB_GT, 0) = as_gt;
    G(ZKC_VMG_ADDSUB_APPLY_ANY, 0) = as_any; G(ZKC_VMG_ADDSUB_UPDATE_FLAGS, 0) = as_any && set_flags;

    /* ---- binop.rs ---------------------------------------------------------------------------------------------------- */
    {
        uint32_t and_c[8] = {0}, or_c[8] = {0}, xor_c[8] = {0}, res[8];
        for (int i = 0; i < 32; i++) {
            const uint32_t x = (a[i / 4] >> (8 * (i % 4))) & 0xFF, y = (b[i / 4] >> (8 * (i % 4))) & 0xFF;
            const uint64_t an = x & y, orr = x | y, xo = x ^ y;
            G(ZKC_VMG_BINOP_COMPOSITE, i) = an | (orr << 16) | (xo << 32);  /* BinopTable row, split at :171-178 */
            G(ZKC_VMG_BINOP_ALL_RESULTS, 3 * i) = an; G(ZKC_VMG_BINOP_ALL_RESULTS, 3 * i + 1) = orr; G(ZKC_VMG_BINOP_ALL_RESULTS, 3 * i + 2) = xo;
            and_c[i / 4] |= (uint32_t)an << (8 * (i % 4)); or_c[i / 4] |= (uint32_t)orr << (8 * (i % 4)); xor_c[i / 4] |= (uint32_t)xo << (8 * (i % 4));
        }
        const int is_and = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_BINOP_AND)), is_or = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_BINOP_OR));
        sel8(is_and, and_c, xor_c, res);               /* :85 */
        if (is_or) memcpy(res, or_c, 32);              /* :86 */
        int z = 1;
        for (int i = 0; i < 8; i++) { G(ZKC_VMG_BINOP_LIMB_IS_ZERO, i) = res[i] == 0; z &= res[i] == 0; }
        put8(out, limit, row, ZKC_VMG_BINOP_AND, and_c); put8(out, limit, row, ZKC_VMG_BINOP_OR, or_c); put8(out, limit, row, ZKC_VMG_BINOP_XOR, xor_c);
        put8(out, limit, row, ZKC_VMG_BINOP_RESULT, res);
        G(ZKC_VMG_BINOP_RESULT_IS_ZERO, 0) = z;
        G(ZKC_VMG_BINOP_UPDATE_FLAGS, 0) = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_BINOP)) && set_flags;
    }

    /* ---- mul_div.rs -------------------------------------------------------------------------------------------------- */
    uint32_t mul_lo[8], mul_hi[8], quot[8], rem[8], md_rem[8], md_a[8], md_low[8], md_high[8], div_sub[8];
    const int apply_mul = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_MUL)), apply_div = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_DIV)), md_any = apply_mul || apply_div;
    int div_less;
    {
        uint32_t r0[8], r1[8], r1m[8];
        mul256(a, b, mul_lo, mul_hi);
        divrem256(a, b, quot, rem);
        sel8(apply_mul, mul_lo, quot, r0); sel8(apply_mul, mul_hi, rem, r1);              /* :258-269 */
        sel8(apply_mul, zero8, rem, md_rem); sel8(apply_mul, a, quot, md_a);              /* :281-288 */
        sel8(apply_mul, mul_lo, a, md_low); sel8(apply_mul, mul_hi, zero8, md_high);      /* :290-297 */
        const int high_z = is_zero256(mul_hi), low_z = is_zero256(mul_lo), of_mul = !high_z, eq_mul = low_z, gt_mul = !of_mul && !eq_mul;
        const int divisor_z = is_zero256(b), quot_z = is_zero256(quot), rem_z = is_zero256(rem);
        div_less = sub256(rem, b, div_sub);                                               /* :330-331 */
        const int mask = apply_div && divisor_z;                                          /* :354 */
        for (int i = 0; i < 8; i++) r1m[i] = mask ? 0 : r1[i];
        const int of_div = divisor_z, eq_div = !divisor_z && quot_z, gt_div = !divisor_z && rem_z;
        put8(out, limit, row, ZKC_VMG_MUL_LOW, mul_lo); put8(out, limit, row, ZKC_VMG_MUL_HIGH, mul_hi);
        put8(out, limit, row, ZKC_VMG_DIV_QUOTIENT, quot); put8(out, limit, row, ZKC_VMG_DIV_REMAINDER, rem);
        put8(out, limit, row, ZKC_VMG_MULDIV_RESULT_0, r0); put8(out, limit, row, ZKC_VMG_MULDIV_RESULT_1_UNMASKED, r1);
        put8(out, limit, row, ZKC_VMG_MULDIV_REM_TO_ENFORCE, md_rem); put8(out, limit, row, ZKC_VMG_MULDIV_A_TO_ENFORCE, md_a);
        put8(out, limit, row, ZKC_VMG_MULDIV_MUL_LOW_TO_ENFORCE, md_low); put8(out, limit, row, ZKC_VMG_MULDIV_MUL_HIGH_TO_ENFORCE, md_high);
        G(ZKC_VMG_MUL_HIGH_IS_ZERO, 0) = high_z; G(ZKC_VMG_MUL_LOW_IS_ZERO, 0) = low_z; G(ZKC_VMG_MUL_OF, 0) = of_mul; G(ZKC_VMG_MUL_GT, 0) = gt_mul;
        G(ZKC_VMG_DIV_DIVISOR_IS_ZERO, 0) = divisor_z; G(ZKC_VMG_DIV_QUOTIENT_IS_ZERO, 0) = quot_z; G(ZKC_VMG_DIV_REMAINDER_IS_ZERO, 0) = rem_z;
        put8(out, limit, row, ZKC_VMG_DIV_SUB_RESULT, div_sub); G(ZKC_VMG_DIV_REMAINDER_IS_LESS, 0) = div_less;
        G(ZKC_VMG_DIV_MASK_REMAINDER, 0) = mask; put8(out, limit, row, ZKC_VMG_MULDIV_RESULT_1, r1m);
        G(ZKC_VMG_DIV_EQ, 0) = eq_div; G(ZKC_VMG_DIV_GT, 0) = gt_div;
        G(ZKC_VMG_MULDIV_OF, 0) = apply_mul ? of_mul : of_div; G(ZKC_VMG_MULDIV_EQ, 0) = apply_mul ? eq_mul : eq_div;
        G(ZKC_VMG_MULDIV_GT, 0) = apply_mul ? gt_mul : gt_div;
        G(ZKC_VMG_MULDIV_APPLY_ANY, 0) = md_any; G(ZKC_VMG_MULDIV_SET_FLAGS, 0) = md_any && set_flags;
    }

    /* ---- shifts.rs ----------------------------------------------------------------------------------------------------- */
    uint32_t shc[8], sh_rem[8], sh_a[8], sh_low[8], sh_high[8], sh_sub[8], sh_rr[8];
    const int apply_shift = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_SHIFT));
    int sh_less;
    {
        const int is_rol = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_SHIFT_ROL)), is_ror = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_SHIFT_ROR)),
                  is_shr = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_SHIFT_SHR));
        const int is_cyclic = is_rol || is_ror, is_right = is_ror || is_shr;
        const uint32_t shift = b[0] & 0xFF;                                                /* :57 */
        const int shift_z = shift == 0;
        const uint32_t inverted = 256 - shift;                                             /* :65, a field element (256 for shift 0) */
        const int change = is_ror && !shift_z;
        const uint32_t full = change ? inverted : shift;                                   /* :71, 8 bits again */
        memset(shc, 0, sizeof shc); shc[full >> 5] = 1u << (full & 31);                    /* tables/bitshift.rs:22-33: 2^full */
        const int is_right_shift = is_right && !is_cyclic;                                 /* :78-81 */
        const int apply_left = apply_shift && !is_right_shift;                             /* :84-87 */
        uint32_t rq[8], rr[8], ll[8], lh[8], temp[8], fin[8];
        divrem256(a, shc, rq, rr); mul256(a, shc, ll, lh);
        memcpy(sh_rr, rr, 32);
        sel8(apply_left, zero8, rr, sh_rem); sel8(apply_left, a, rq, sh_a);                /* :99-101 */
        sel8(apply_left, ll, a, sh_low); sel8(apply_left, lh, zero8, sh_high);             /* :103-105 */
        sh_less = sub256(rr, shc, sh_sub);                                                 /* :117-118 */
        sel8(is_right_shift, rq, ll, temp);                                                /* :136 */
        for (int i = 0; i < 8; i++) fin[i] = (is_cyclic ? lh[i] : 0u) + temp[i];           /* :141-152 */
        G(ZKC_VMG_SHIFT_AMOUNT, 0) = shift; G(ZKC_VMG_SHIFT_IS_ZERO, 0) = shift_z; G(ZKC_VMG_SHIFT_INVERTED, 0) = inverted;
        G(ZKC_VMG_SHIFT_CHANGE_FLAG, 0) = change; G(ZKC_VMG_SHIFT_FULL, 0) = full; put8(out, limit, row, ZKC_VMG_SHIFT_CONSTANT, shc);
        G(ZKC_VMG_SHIFT_IS_RIGHT, 0) = is_right_shift; put8(out, limit, row, ZKC_VMG_SHIFT_RSHIFT_Q, rq); put8(out, limit, row, ZKC_VMG_SHIFT_RSHIFT_R, rr);
        G(ZKC_VMG_SHIFT_APPLY_LEFT, 0) = apply_left; put8(out, limit, row, ZKC_VMG_SHIFT_LSHIFT_LOW, ll); put8(out, limit, row, ZKC_VMG_SHIFT_LSHIFT_HIGH, lh);
        put8(out, limit, row, ZKC_VMG_SHIFT_REM_TO_ENFORCE, sh_rem); put8(out, limit, row, ZKC_VMG_SHIFT_A_TO_ENFORCE, sh_a);
        put8(out, limit, row, ZKC_VMG_SHIFT_MUL_LOW_TO_ENFORCE, sh_low); put8(out, limit, row, ZKC_VMG_SHIFT_MUL_HIGH_TO_ENFORCE, sh_high);
        put8(out, limit, row, ZKC_VMG_SHIFT_SUB_RESULT, sh_sub); G(ZKC_VMG_SHIFT_REMAINDER_IS_LESS, 0) = sh_less;
        put8(out, limit, row, ZKC_VMG_SHIFT_TEMP_RESULT, temp); put8(out, limit, row, ZKC_VMG_SHIFT_RESULT, fin);
        G(ZKC_VMG_SHIFT_RESULT_IS_ZERO, 0) = is_zero256(fin); G(ZKC_VMG_SHIFT_SET_FLAGS, 0) = apply_shift && set_flags;
    }

    /* ---- cycle.rs:619-670: the candidates in push order add_sub, mul_div, shifts; the LAST pushed is the default ----------- */
    {
        uint32_t rc[8], ra_[8], rb_[8], rcc[8];
        memcpy(rc, sh_sub, 32);                                    /* :620-627 */
        if (as_any) memcpy(rc, as_result, 32);
        if (md_any) memcpy(rc, div_sub, 32);
        put8(out, limit, row, ZKC_VMG_RANGE_CHECK, rc);
        /* AddSubRelation { a, b, c, of }: shifts (a = shift constant, b = sub result, c = rshift_r), add_sub (a = src1, new_b, new_c),
         * mul_div (a = src1, b = sub result, c = remainder) */
        int of = sh_less;
        memcpy(ra_, shc, 32); memcpy(rb_, sh_sub, 32); memcpy(rcc, sh_rr, 32);
        if (as_any) { memcpy(ra_, b, 32); memcpy(rb_, new_b, 32); memcpy(rcc, new_c, 32); of = new_of; }
        if (md_any) { memcpy(ra_, b, 32); memcpy(rb_, div_sub, 32); memcpy(rcc, rem, 32); of = div_less; }
        put8(out, limit, row, ZKC_VMG_ADDREL_A, ra_); put8(out, limit, row, ZKC_VMG_ADDREL_B, rb_); put8(out, limit, row, ZKC_VMG_ADDREL_C, rcc);
        G(ZKC_VMG_ADDREL_OF, 0) = of;
        uint64_t carry = 0;                                        /* opcodes/mod.rs:107-117 */
        for (int i = 0; i < 8; i++) { carry = ((uint64_t)ra_[i] + rb_[i] + carry) >> 32; G(ZKC_VMG_ADDREL_CARRY, i) = carry; }
        /* MulDivRelation { a, b, rem, mul_low, mul_high }: shifts by default, mul_div when it applies */
        uint32_t ma[8], mb[8], mrem[8], mlow[8], mhigh[8];
        memcpy(ma, sh_a, 32); memcpy(mb, shc, 32); memcpy(mrem, sh_rem, 32); memcpy(mlow, sh_low, 32); memcpy(mhigh, sh_high, 32);
        if (md_any) { memcpy(ma, md_a, 32); memcpy(mb, b, 32); memcpy(mrem, md_rem, 32); memcpy(mlow, md_low, 32); memcpy(mhigh, md_high, 32); }
        put8(out, limit, row, ZKC_VMG_MULREL_A, ma); put8(out, limit, row, ZKC_VMG_MULREL_B, mb); put8(out, limit, row, ZKC_VMG_MULREL_REM, mrem);
        put8(out, limit, row, ZKC_VMG_MULREL_LOW, mlow); put8(out, limit, row, ZKC_VMG_MULREL_HIGH, mhigh);
        uint32_t partial[16];                                      /* opcodes/mod.rs:146-169 */
        memset(partial, 0, sizeof partial); memcpy(partial, mrem, 32);
        for (int ai = 0; ai < 8; ai++) {
            uint32_t overflow = 0;
            for (int bi = 0; bi < 8; bi++) {
                const uint64_t t = (uint64_t)ma[ai] * mb[bi] + partial[ai + bi] + overflow;  /* UInt32::fma_with_carry */
                partial[ai + bi] = (uint32_t)t; overflow = (uint32_t)(t >> 32);
                G(ZKC_VMG_MULREL_PARTIAL_LOW, 8 * ai + bi) = (uint32_t)t; G(ZKC_VMG_MULREL_PARTIAL_HIGH, 8 * ai + bi) = overflow;
            }
            partial[ai + 8] += overflow;                           /* add_no_overflow */
            G(ZKC_VMG_MULREL_ROW_END, ai) = partial[ai + 8];
        }
    }
#undef BIT
}

void orc_main_vm_gadget_cells(const uint64_t *trace, size_t limit, size_t n_instances, uint64_t *out_all) {
    for (size_t inst = 0; inst < n_instances; inst++) {
        const uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit;
        uint64_t *out = out_all + inst * (size_t)ZKC_VMG_NUM_COLS * limit;
        for (size_t row = 0; row < limit; row++) {
            uint32_t a[8], b[8];
            for (int i = 0; i < 8; i++) { a[i] = (uint32_t)t[(size_t)(ZKC_VM_SRC0 + 1 + i) * limit + row]; b[i] = (uint32_t)t[(size_t)(ZKC_VM_SRC1 + 1 + i) * limit + row]; }
            gadget_row(t[(size_t)ZKC_VM_PROPS * limit + row], a, b, out, limit, row);
        }
    }
}

/* ---- the second block: apply_ptr, apply_jump, apply_context (ZKC_VM_STATE_GADGET_COLUMNS) --------------------------------------
 *   apply_ptr      /root/reference/src/main_vm/opcodes/ptr.rs:6-183
 *   apply_jump     /root/reference/src/main_vm/opcodes/jump.rs:3-38
 *   apply_context  /root/reference/src/main_vm/opcodes/context.rs:7-307
 *   apply_nop      /root/reference/src/main_vm/opcodes/nop.rs:4-24 (allocates nothing)
 * Pinning: PARITY UNPINNED against the reference; the values are checked against an independent Python statement
 * (tests/test_oracle_main_vm_gadgets.py). */
#define S(col, i) out[(size_t)((col) + (i)) * limit + row]
static void state_gadget_row(uint64_t props, int a_ptr, const uint32_t *a, int b_ptr, const uint32_t *b, uint32_t new_sp, uint32_t ergs_left,
                             const zkc_vm_state *st, uint64_t *out, size_t limit, size_t row) {
#define BIT(n) (int)((props >> (n)) & 1)
    /* ---- ptr.rs ---- */
    {
        const int should_apply = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_PTR));
        const int v_add = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_PTR_ADD)), v_sub = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_PTR_SUB)),
                  v_pack = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_PTR_PACK)), v_shrink = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_PTR_SHRINK));
        const int src1_is_integer = !b_ptr;                                   /* :53 */
        const int args_valid = a_ptr && src1_is_integer, args_invalid = !args_valid;   /* :56-57 */
        int lz[8], hi_zero = 1, lo_zero = 1;
        for (int i = 0; i < 8; i++) { lz[i] = b[i] == 0; S(ZKC_VMS_PTR_SRC1_LIMB_IS_ZERO, i) = lz[i]; }
        for (int i = 1; i < 8; i++) hi_zero &= lz[i];                         /* :64 */
        for (int i = 0; i < 4; i++) lo_zero &= lz[i];                         /* :65 */
        const int hi_nonzero = !hi_zero, arith = v_add || v_sub, too_large = hi_nonzero && arith;   /* :67-71 */
        const int lo_nonzero = !lo_zero, dirty_pack = lo_nonzero && v_pack;   /* :74-76 */
        const uint64_t sum = (uint64_t)a[0] + b[0];                           /* :79 */
        const uint32_t add_r = (uint32_t)sum; const int add_of = (int)(sum >> 32), add_panic = v_add && add_of;
        const uint32_t sub_r = a[0] - b[0]; const int sub_uf = a[0] < b[0], sub_panic = v_sub && sub_uf;          /* :82-83 */
        const uint32_t shr_r = a[3] - b[0]; const int shr_uf = a[3] < b[0], shr_panic = v_shrink && shr_uf;       /* :85-86 */
        const int any_panic = args_invalid || too_large || dirty_pack || add_panic || sub_panic || shr_panic;    /* :88-98 */
        const int should_panic = should_apply && any_panic, ok = !any_panic, update = should_apply && ok;        /* :100-102 */
        const uint32_t low_if_add = v_add ? add_r : a[0];                     /* :107-112 */
        const uint32_t low_if_add_or_sub = v_sub ? sub_r : low_if_add;        /* :115-120 */
        const uint32_t b96_if_shrink = v_shrink ? shr_r : a[3];               /* :123-128 */
        uint32_t highest[4];
        for (int i = 0; i < 4; i++) highest[i] = v_pack ? b[4 + i] : a[4 + i];   /* :130-145 */
        const uint32_t lowest32 = v_pack ? a[0] : low_if_add_or_sub;          /* :147-152 */
        const uint32_t b96 = v_pack ? a[3] : b96_if_shrink;                   /* :154-159 */
        S(ZKC_VMS_PTR_SRC1_IS_INTEGER, 0) = src1_is_integer; S(ZKC_VMS_PTR_ARGS_VALID, 0) = args_valid; S(ZKC_VMS_PTR_ARGS_INVALID, 0) = args_invalid;
        S(ZKC_VMS_PTR_SRC1_32_256_IS_ZERO, 0) = hi_zero; S(ZKC_VMS_PTR_SRC1_0_128_IS_ZERO, 0) = lo_zero; S(ZKC_VMS_PTR_SRC1_32_256_IS_NONZERO, 0) = hi_nonzero;
        S(ZKC_VMS_PTR_ARITH_VARIANT, 0) = arith; S(ZKC_VMS_PTR_TOO_LARGE_OFFSET, 0) = too_large; S(ZKC_VMS_PTR_SRC1_0_128_IS_NONZERO, 0) = lo_nonzero;
        S(ZKC_VMS_PTR_DIRTY_PACK, 0) = dirty_pack; S(ZKC_VMS_PTR_ADD_RESULT, 0) = add_r; S(ZKC_VMS_PTR_ADD_OF, 0) = add_of; S(ZKC_VMS_PTR_ADD_PANIC, 0) = add_panic;
        S(ZKC_VMS_PTR_SUB_RESULT, 0) = sub_r; S(ZKC_VMS_PTR_SUB_UF, 0) = sub_uf; S(ZKC_VMS_PTR_SUB_PANIC, 0) = sub_panic;
        S(ZKC_VMS_PTR_SHRINK_RESULT, 0) = shr_r; S(ZKC_VMS_PTR_SHRINK_UF, 0) = shr_uf; S(ZKC_VMS_PTR_SHRINK_PANIC, 0) = shr_panic;
        S(ZKC_VMS_PTR_ANY_PANIC, 0) = any_panic; S(ZKC_VMS_PTR_SHOULD_PANIC, 0) = should_panic; S(ZKC_VMS_PTR_OK, 0) = ok; S(ZKC_VMS_PTR_UPDATE_REGISTER, 0) = update;
        S(ZKC_VMS_PTR_LOW_IF_ADD, 0) = low_if_add; S(ZKC_VMS_PTR_LOW_IF_ADD_OR_SUB, 0) = low_if_add_or_sub; S(ZKC_VMS_PTR_96_128_IF_SHRINK, 0) = b96_if_shrink;
        for (int i = 0; i < 4; i++) S(ZKC_VMS_PTR_HIGHEST_128, i) = highest[i];
        S(ZKC_VMS_PTR_LOWEST32, 0) = lowest32; S(ZKC_VMS_PTR_96_128, 0) = b96;
        S(ZKC_VMS_PTR_DST0, 0) = a_ptr;                                       /* :161-176 */
        S(ZKC_VMS_PTR_DST0, 1) = lowest32; S(ZKC_VMS_PTR_DST0, 2) = a[1]; S(ZKC_VMS_PTR_DST0, 3) = a[2]; S(ZKC_VMS_PTR_DST0, 4) = b96;
        for (int i = 0; i < 4; i++) S(ZKC_VMS_PTR_DST0, 5 + i) = highest[i];
    }
    /* ---- jump.rs:27-33: UInt16::from_le_bytes of the two low bytes of src0 ---- */
    S(ZKC_VMS_JUMP_DST, 0) = a[0] & 0xFFFF;
    /* ---- context.rs ---- */
    {
        const zkc_vm_context *c = &st->current_context;
        const int should_apply = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_CONTEXT));
        const int is_this = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_THIS)), is_caller = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_CALLER)),
                  is_code = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_CODE_ADDRESS)), is_meta = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_META)),
                  is_ergs = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_ERGS_LEFT)), is_get_u128 = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_GET_U128)),
                  is_set_u128 = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_SET_U128)),
                  is_set_ergs = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_SET_ERGS_PER_PUBDATA)), is_inc_tx = BIT(ZKC_VM_BIT_VARIANT(ZKC_VAR_CONTEXT_INC_TX_NUMBER));
        const int read_only = is_set_u128 || is_set_ergs || is_inc_tx, write_like = !read_only;   /* :120-124 */
        S(ZKC_VMS_CTX_WRITE_TO_CONTEXT, 0) = should_apply && is_set_u128; S(ZKC_VMS_CTX_SET_PUBDATA_ERGS, 0) = should_apply && is_set_ergs;   /* :115-116 */
        S(ZKC_VMS_CTX_INCREMENT_TX, 0) = should_apply && is_inc_tx;          /* :117 */
        S(ZKC_VMS_CTX_READ_ONLY, 0) = read_only; S(ZKC_VMS_CTX_WRITE_LIKE, 0) = write_like; S(ZKC_VMS_CTX_WRITE_TO_DST0, 0) = should_apply && write_like;   /* :126 */
        const uint64_t tx = (uint64_t)st->tx_number_in_block + 1;            /* :130-133 */
        S(ZKC_VMS_CTX_INCREMENTED_TX_NUMBER, 0) = (uint32_t)tx; S(ZKC_VMS_CTX_TX_OF, 0) = tx >> 32;
        const uint32_t meta_hi = (c->this_shard_id & 0xFF) | (c->caller_shard_id & 0xFF) << 8 | (c->code_shard_id & 0xFF) << 16;   /* :145-165 */
        S(ZKC_VMS_CTX_META_HIGHEST, 0) = meta_hi;
        uint32_t r[8];
        memset(r, 0, sizeof r);
        r[0] = is_ergs ? ergs_left : new_sp;                                  /* :190-208 */
        S(ZKC_VMS_CTX_LOW_U32, 0) = r[0];
        if (is_get_u128) memcpy(r, c->context_u128_value_composite, 16);      /* :212-223 */
        for (int i = 0; i < 4; i++) S(ZKC_VMS_CTX_RESULT_128, i) = r[i];
        if (is_this) memcpy(r, c->this_address, 20);                          /* :235-245 */
        for (int i = 0; i < 5; i++) S(ZKC_VMS_CTX_RESULT_160_THIS, i) = r[i];
        if (is_caller) memcpy(r, c->caller, 20);                              /* :247-257 */
        for (int i = 0; i < 5; i++) S(ZKC_VMS_CTX_RESULT_160_CALLER, i) = r[i];
        if (is_code) memcpy(r, c->code_address, 20);                          /* :259-269 */
        for (int i = 0; i < 5; i++) S(ZKC_VMS_CTX_RESULT_160_CODE, i) = r[i];
        if (is_meta) {                                                        /* :167-186, :284-285 */
            r[0] = st->ergs_per_pubdata_byte; r[1] = 0; r[2] = c->heap_upper_bound; r[3] = c->aux_heap_upper_bound;
            r[4] = r[5] = r[6] = 0; r[7] = meta_hi;
        }
        for (int i = 0; i < 8; i++) S(ZKC_VMS_CTX_RESULT_256, i) = r[i];
    }
#undef BIT
}

void orc_main_vm_state_gadget_cells(const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances, uint64_t *out_all) {
    for (size_t inst = 0; inst < n_instances; inst++) {
        const uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit;
        const zkc_vm_state *snaps = snapshots + inst * (limit + 1);
        uint64_t *out = out_all + inst * (size_t)ZKC_VMS_NUM_COLS * limit;
        for (size_t row = 0; row < limit; row++) {
            uint32_t a[8], b[8];
            for (int i = 0; i < 8; i++) { a[i] = (uint32_t)t[(size_t)(ZKC_VM_SRC0 + 1 + i) * limit + row]; b[i] = (uint32_t)t[(size_t)(ZKC_VM_SRC1 + 1 + i) * limit + row]; }
            state_gadget_row(t[(size_t)ZKC_VM_PROPS * limit + row], (int)t[(size_t)ZKC_VM_SRC0 * limit + row], a, (int)t[(size_t)ZKC_VM_SRC1 * limit + row], b,
                             (uint32_t)t[(size_t)ZKC_VM_NEW_SP * limit + row], (uint32_t)t[(size_t)ZKC_VM_DIRTY_ERGS_LEFT * limit + row], snaps + row, out, limit, row);
        }
    }
}

/* ---- the memory-queue relations of every cycle (ZKC_VM_MEMORY_SPONGE_COLUMNS): opcode fetch, src0 read, dst0 write --------------
 *   may_be_read_memory_for_code             /root/reference/src/main_vm/utils.rs:129-233
 *   may_be_read_memory_for_source_operand   /root/reference/src/main_vm/utils.rs:388-522
 *   may_be_write_memory                     /root/reference/src/main_vm/cycle.rs:799-935
 *   enforce_sponges                         /root/reference/src/main_vm/cycle.rs:937-957
 * Pinning: PARITY UNPINNED (Poseidon2 has no known-answer vector in the reference); checked against the second Python
 * restatement of the permutation and against the DENSE trace's own enforced slots (tests/test_oracle_main_vm_gadgets.py). */
static void memq_step(const uint64_t enc[8], int execute, uint64_t state[12], uint32_t *len, uint64_t *out, size_t limit, size_t row, int col_init) {
    uint64_t s[12];
    memcpy(s, enc, 64); memcpy(s + 8, state + 8, 32);                      /* absorb with replacement */
    for (int i = 0; i < 12; i++) out[(size_t)(col_init + i) * limit + row] = s[i];
    orc_poseidon2_permutation(s);
    for (int i = 0; i < 12; i++) out[(size_t)(col_init + 12 + i) * limit + row] = s[i];
    if (execute) { memcpy(state, s, 96); (*len)++; }                       /* Num::parallel_select / UInt32::conditionally_select */
    for (int i = 0; i < 12; i++) out[(size_t)(col_init + 24 + i) * limit + row] = state[i];
    out[(size_t)(col_init + 36) * limit + row] = *len;
}

void orc_main_vm_memory_sponge_cells(const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances, uint64_t *out_all) {
    for (size_t inst = 0; inst < n_instances; inst++) {
        const uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit;
        const zkc_vm_state *snaps = snapshots + inst * (limit + 1);
        uint64_t *out = out_all + inst * (size_t)ZKC_VMQ_NUM_COLS * limit;
        for (size_t row = 0; row < limit; row++) {
#define T(c) t[(size_t)(c) * limit + row]
            const zkc_vm_state *st = snaps + row;
            const uint64_t props = T(ZKC_VM_PROPS);
            uint64_t state[12], enc[8];
            uint32_t len = st->memory_queue_length;
            memcpy(state, st->memory_queue_state, 96);
            out[(size_t)ZKC_VMQ_SELECTED * limit + row] =
                !(((props >> ZKC_VM_BIT_TYPE(ZKC_OP_UMA)) | (props >> ZKC_VM_BIT_TYPE(ZKC_OP_LOG)) | (props >> ZKC_VM_BIT_TYPE(ZKC_OP_NEAR_CALL)) |
                   (props >> ZKC_VM_BIT_TYPE(ZKC_OP_FAR_CALL)) | (props >> ZKC_VM_BIT_TYPE(ZKC_OP_RET))) & 1);
            zkc_memory_query q;
            /* opcode fetch: pre_state.rs:133-167; the word read is CODE_WORD when the read happens, zero otherwise */
            memset(&q, 0, sizeof q);
            const int read_opcode = (int)T(ZKC_VM_SHOULD_READ_OPCODE);
            q.timestamp = st->timestamp; q.memory_page = st->current_context.code_page; q.index = (uint32_t)T(ZKC_VM_SUPER_PC);
            for (int i = 0; i < 8; i++) q.value[i] = read_opcode ? (uint32_t)T(ZKC_VM_CODE_WORD + i) : 0u;
            orc_memory_query_encode(&q, enc);
            memq_step(enc, read_opcode, state, &len, out, limit, row, ZKC_VMQ_FETCH_INIT);
            /* src0 read: pre_state.rs:374-392, the same timestamp */
            memset(&q, 0, sizeof q);
            q.timestamp = st->timestamp; q.memory_page = (uint32_t)T(ZKC_VM_SRC0_PAGE); q.index = (uint32_t)T(ZKC_VM_SRC0_INDEX);
            q.is_ptr = (uint32_t)T(ZKC_VM_SRC0_FROM_MEMORY) & 1;
            for (int i = 0; i < 8; i++) q.value[i] = (uint32_t)T(ZKC_VM_SRC0_FROM_MEMORY + 1 + i);
            orc_memory_query_encode(&q, enc);
            memq_step(enc, (int)T(ZKC_VM_SHOULD_READ_SRC0), state, &len, out, limit, row, ZKC_VMQ_SRC0_INIT);
            /* dst0 write: cycle.rs:248-284, timestamp_for_dst_write = timestamp + 3 (pre_state.rs:143-149) */
            memset(&q, 0, sizeof q);
            q.timestamp = st->timestamp + 3; q.memory_page = (uint32_t)T(ZKC_VM_DST0_PAGE); q.index = (uint32_t)T(ZKC_VM_DST0_INDEX);
            q.rw_flag = 1; q.is_ptr = (uint32_t)T(ZKC_VM_DST0) & 1;
            for (int i = 0; i < 8; i++) q.value[i] = (uint32_t)T(ZKC_VM_DST0 + 1 + i);
            orc_memory_query_encode(&q, enc);
            memq_step(enc, (int)T(ZKC_VM_PERFORM_DST0_MEMORY_WRITE), state, &len, out, limit, row, ZKC_VMQ_DST0_INIT);
#undef T
        }
    }
}

/* ---- cells of create_prestate that are not columns of the DENSE trace (ZKC_VM_PRESTATE_COLUMNS) ------------------------------
 *   create_prestate                              /root/reference/src/main_vm/pre_state.rs:71-519
 *   should_read_memory                           /root/reference/src/main_vm/utils.rs:106-120
 *   resolve_memory_region_and_index_for_source   /root/reference/src/main_vm/utils.rs:237-305
 *   resolve_memory_region_and_index_for_dest     /root/reference/src/main_vm/utils.rs:307-386
 *   register selector masks                      /root/reference/src/main_vm/decoded_opcode.rs:192-202
 * Pinning: PARITY UNPINNED against the reference; checked against an independent Python statement and against the DENSE trace's
 * own results (tests/test_oracle_main_vm_gadgets.py). */
#define PC(col, i) out[(size_t)((col) + (i)) * limit + row]
static void put_reg(uint64_t *out, size_t limit, size_t row, int col, const zkc_vm_register *r) {
    PC(col, 0) = r->is_pointer & 1;
    for (int i = 0; i < 8; i++) PC(col, 1 + i) = r->value[i];
}
static uint32_t reg_mask(uint32_t idx) { return idx ? 1u << (idx - 1) : 0u; }   /* tables/integer_to_boolean_mask.rs:33-41 */

void orc_main_vm_prestate_cells(const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances, uint64_t *out_all) {
    for (size_t inst = 0; inst < n_instances; inst++) {
        const uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit;
        const zkc_vm_state *snaps = snapshots + inst * (limit + 1);
        uint64_t *out = out_all + inst * (size_t)ZKC_VMP_NUM_COLS * limit;
        for (size_t row = 0; row < limit; row++) {
#define T(c) t[(size_t)(c) * limit + row]
            const zkc_vm_state *st = snaps + row;
            const zkc_vm_context *c = &st->current_context;
            const uint64_t props = T(ZKC_VM_PROPS);
#define BIT(n) (int)((props >> (n)) & 1)
            /* ---- cycle control, pre_state.rs:88-156 ---- */
            const int should_skip = (int)T(ZKC_VM_SHOULD_SKIP_CYCLE), pending = (int)T(ZKC_VM_PENDING_EXCEPTION_IN);
            const int execute_cycle = !should_skip;                                   /* :92 */
            PC(ZKC_VMP_EXECUTE_CYCLE, 0) = execute_cycle;
            PC(ZKC_VMP_SHOULD_TRY_TO_READ_OPCODE, 0) = execute_cycle && !pending;      /* :98: execute_cycle.mask_negated(pending_exception) */
            PC(ZKC_VMP_PENDING_EXCEPTION_TAKEN_DOWN, 0) = pending && !pending;         /* :103-105: masked by itself */
            const uint32_t pc1 = c->pc + 1;                                           /* :109-111, UInt16 */
            PC(ZKC_VMP_PC_PLUS_ONE, 0) = pc1 & 0xFFFF; PC(ZKC_VMP_PC_PLUS_ONE_OF, 0) = pc1 >> 16;
            const uint32_t super_pc = (uint32_t)T(ZKC_VM_SUPER_PC), sub_pc = (uint32_t)T(ZKC_VM_SUB_PC);
            const int pages_eq = st->previous_code_page == c->code_page, spc_eq = super_pc == st->previous_super_pc;   /* utils.rs:114-115 */
            PC(ZKC_VMP_CODE_PAGES_ARE_EQUAL, 0) = pages_eq; PC(ZKC_VMP_SUPER_PC_ARE_EQUAL, 0) = spc_eq;
            PC(ZKC_VMP_CAN_SKIP_READ, 0) = pages_eq && spc_eq; PC(ZKC_VMP_SHOULD_READ_FOR_NEW_PC, 0) = !(pages_eq && spc_eq);
            for (int i = 0; i < 4; i++) PC(ZKC_VMP_TIMESTAMPS, i) = (uint32_t)(st->timestamp + 1 + i);   /* :144-150, increment_unchecked */
            PC(ZKC_VMP_NEXT_CYCLE_TIMESTAMP, 0) = should_skip ? st->timestamp : (uint32_t)(st->timestamp + 4);   /* :151-156 */
            /* ---- the opcode inside the code word, :183-214 ---- */
            {
                const uint32_t m = reg_mask(sub_pc);
                uint32_t lo = (uint32_t)T(ZKC_VM_CODE_WORD + 6), hi = (uint32_t)T(ZKC_VM_CODE_WORD + 7);
                for (int k = 0; k < 3; k++) {
                    const int bit = (int)((m >> k) & 1);
                    PC(ZKC_VMP_SUBPC_BITMASK, k) = bit;
                    if (bit) { lo = (uint32_t)T(ZKC_VM_CODE_WORD + 4 - 2 * k); hi = (uint32_t)T(ZKC_VM_CODE_WORD + 5 - 2 * k); }
                    PC(ZKC_VMP_OPCODE_SELECT_CHAIN, 2 * k) = lo; PC(ZKC_VMP_OPCODE_SELECT_CHAIN, 2 * k + 1) = hi;
                }
            }
            /* ---- register selectors and the select chains, decoded_opcode.rs:192-202, pre_state.rs:303-329 ---- */
            const uint32_t idx[4] = {(uint32_t)T(ZKC_VM_SRC0_REG), (uint32_t)T(ZKC_VM_SRC1_REG), (uint32_t)T(ZKC_VM_DST0_REG), (uint32_t)T(ZKC_VM_DST1_REG)};
            const int sel_col[4] = {ZKC_VMP_SRC0_SELECTORS, ZKC_VMP_SRC1_SELECTORS, ZKC_VMP_DST0_SELECTORS, ZKC_VMP_DST1_SELECTORS};
            for (int k = 0; k < 4; k++)
                for (int r = 0; r < 15; r++) PC(sel_col[k], r) = (reg_mask(idx[k]) >> r) & 1;
            zkc_vm_register draft_src0, src1_register;
            memset(&draft_src0, 0, sizeof draft_src0); memset(&src1_register, 0, sizeof src1_register);
            uint32_t dst0_low = 0;
            for (int r = 0; r < 15; r++) {
                if ((reg_mask(idx[0]) >> r) & 1) draft_src0 = st->registers[r];
                if ((reg_mask(idx[1]) >> r) & 1) src1_register = st->registers[r];
                if ((reg_mask(idx[2]) >> r) & 1) dst0_low = st->registers[r].value[0];
                put_reg(out, limit, row, ZKC_VMP_DRAFT_SRC0_CHAIN + 9 * r, &draft_src0);
                put_reg(out, limit, row, ZKC_VMP_SRC1_REGISTER_CHAIN + 9 * r, &src1_register);
                PC(ZKC_VMP_DST0_REG_LOW_CHAIN, r) = dst0_low;
            }
            const uint32_t src0_lowest = draft_src0.value[0] & 0xFFFF, dst0_lowest = dst0_low & 0xFFFF;   /* :310, :329 */
            PC(ZKC_VMP_SRC0_REG_LOWEST, 0) = src0_lowest; PC(ZKC_VMP_DST0_REG_LOWEST, 0) = dst0_lowest;
            const uint32_t stack_page = c->base_page + 1;                            /* :341-343 */
            PC(ZKC_VMP_STACK_PAGE, 0) = stack_page; PC(ZKC_VMP_HEAP_PAGE, 0) = (uint32_t)(stack_page + 1); PC(ZKC_VMP_AUX_HEAP_PAGE, 0) = (uint32_t)(stack_page + 2);
            const int not_nop = !BIT(ZKC_VM_BIT_TYPE(ZKC_OP_NOP));
            PC(ZKC_VMP_NOT_NOP, 0) = not_nop;
            const uint32_t imm0 = (uint32_t)T(ZKC_VM_IMM0), imm1 = (uint32_t)T(ZKC_VM_IMM1), sp = c->sp & 0xFFFF;
            uint32_t sp_after_src0;
            {   /* utils.rs:237-305 */
                const int use_code = BIT(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_CODE_PAGE)), abs_ = BIT(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_STACK_ABSOLUTE)),
                          rel = BIT(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_STACK_OFFSET)), pp = BIT(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_STACK_PUSH_POP));
                const uint32_t idx_abs = (src0_lowest + imm0) & 0xFFFF, idx_rel = (sp - idx_abs) & 0xFFFF;
                const int use_stack = abs_ || rel || pp;
                PC(ZKC_VMP_SRC_ABSOLUTE_MODE, 0) = use_code || abs_; PC(ZKC_VMP_SRC_INDEX_FOR_ABSOLUTE, 0) = idx_abs; PC(ZKC_VMP_SRC_INDEX_FOR_RELATIVE, 0) = idx_rel;
                PC(ZKC_VMP_SRC_USE_STACK, 0) = use_stack; PC(ZKC_VMP_SRC_DID_READ_UNMASKED, 0) = use_stack || use_code;
                sp_after_src0 = pp ? idx_rel : sp;
            }
            {   /* utils.rs:307-386 */
                const int abs_ = BIT(ZKC_VM_BIT_DST_MODE(ZKC_MODE_STACK_ABSOLUTE)), rel = BIT(ZKC_VM_BIT_DST_MODE(ZKC_MODE_STACK_OFFSET)),
                          pp = BIT(ZKC_VM_BIT_DST_MODE(ZKC_MODE_STACK_PUSH_POP));
                const uint32_t idx_abs = (dst0_lowest + imm1) & 0xFFFF, idx_push = (sp_after_src0 + idx_abs) & 0xFFFF, idx_rel = (sp_after_src0 - idx_abs) & 0xFFFF;
                PC(ZKC_VMP_DST_INDEX_FOR_ABSOLUTE, 0) = idx_abs; PC(ZKC_VMP_DST_INDEX_FOR_RELATIVE_WITH_PUSH, 0) = idx_push; PC(ZKC_VMP_DST_INDEX_FOR_RELATIVE, 0) = idx_rel;
                PC(ZKC_VMP_DST_DID_WRITE_UNMASKED, 0) = abs_ || rel || pp; PC(ZKC_VMP_DST_INDEX_SOMEWHAT_RELATIVE, 0) = pp ? sp_after_src0 : idx_rel;
            }
            /* ---- src0 selects, swap, pointer erasure: :403-479 ---- */
            zkc_vm_register from_mem, src0, imm_reg;
            memset(&from_mem, 0, sizeof from_mem); memset(&imm_reg, 0, sizeof imm_reg);
            from_mem.is_pointer = (uint32_t)T(ZKC_VM_SRC0_FROM_MEMORY) & 1;
            for (int i = 0; i < 8; i++) from_mem.value[i] = (uint32_t)T(ZKC_VM_SRC0_FROM_MEMORY + 1 + i);
            imm_reg.value[0] = imm0;
            src0 = BIT(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_REG_ONLY)) ? draft_src0 : from_mem;    /* :403-406 */
            put_reg(out, limit, row, ZKC_VMP_SRC0_AFTER_USE_REG, &src0);
            if (BIT(ZKC_VM_BIT_SRC_MODE(ZKC_MODE_IMM16))) src0 = imm_reg;                  /* :408-413 */
            put_reg(out, limit, row, ZKC_VMP_SRC0_AFTER_USE_IMM, &src0);
            const int is_ptr_op = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_PTR));
            const int asym = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_SUB)) || BIT(ZKC_VM_BIT_TYPE(ZKC_OP_DIV)) || BIT(ZKC_VM_BIT_TYPE(ZKC_OP_SHIFT));   /* :421-432 */
            const int t0 = asym && BIT(ZKC_VM_BIT_FLAG(ZKC_VM_SWAP_OPERANDS_FLAG_IDX)), t1 = is_ptr_op && BIT(ZKC_VM_BIT_FLAG(ZKC_VM_SWAP_OPERANDS_PTR_FLAG_IDX));
            PC(ZKC_VMP_SWAP_IS_ASSYMMETRIC, 0) = asym; PC(ZKC_VMP_SWAP_T0, 0) = t0; PC(ZKC_VMP_SWAP_T1, 0) = t1;
            const int swap = (int)T(ZKC_VM_SWAP_OPERANDS);
            const zkc_vm_register a = swap ? src1_register : src0, b = swap ? src0 : src1_register;   /* :451-454 */
            put_reg(out, limit, row, ZKC_VMP_SRC0_SWAPPED, &a); put_reg(out, limit, row, ZKC_VMP_SRC1_SWAPPED, &b);
            const int not_kernel = !(c->is_kernel_mode & 1);                          /* :458 */
            const int keeps = BIT(ZKC_VM_BIT_TYPE(ZKC_OP_RET)) || is_ptr_op || BIT(ZKC_VM_BIT_TYPE(ZKC_OP_UMA)) || BIT(ZKC_VM_BIT_TYPE(ZKC_OP_FAR_CALL));
            PC(ZKC_VMP_NOT_KERNEL_MODE, 0) = not_kernel; PC(ZKC_VMP_KEEPS_POINTERS, 0) = keeps; PC(ZKC_VMP_SHOULD_ERASE, 0) = !keeps;   /* :459-474 */
            PC(ZKC_VMP_SHOULD_ERASE_SRC0, 0) = (a.is_pointer & 1) && !keeps && not_kernel;   /* :475 */
            PC(ZKC_VMP_SHOULD_ERASE_SRC1, 0) = (b.is_pointer & 1) && not_kernel;             /* :478 */
#undef BIT
#undef T
        }
    }
}

/* ---- the register write-back of the state diffs (ZKC_VM_WRITEBACK_COLUMNS) ---------------------------------------------------------
 *   dst0 / dst1 update flags, write_as_dst0      /root/reference/src/main_vm/cycle.rs:160-189, :303-330
 *   specific updates, markers, zero-out          /root/reference/src/main_vm/cycle.rs:349-375
 *   far call register conventions                /root/reference/src/main_vm/opcodes/call_ret_impl/far_call.rs:1041-1070
 *   far return register conventions              /root/reference/src/main_vm/opcodes/call_ret_impl/ret.rs:442-464
 *   is_pointer dot products and selects          /root/reference/src/main_vm/cycle.rs:377-412
 *   value select chain                           /root/reference/src/main_vm/cycle.rs:415-433
 * One register at a time, each step of the reference's chain as an `if`.  Pinning: PARITY UNPINNED against the reference; checked
 * against an independent Python statement and against snapshot i + 1 (tests/test_oracle_main_vm_gadgets.py). */
void orc_main_vm_writeback_cells(const zkc_vm_isa *isa, const uint64_t *trace, const zkc_vm_state *snapshots, size_t limit, size_t n_instances,
                                 uint64_t *out_all) {
    for (size_t inst = 0; inst < n_instances; inst++) {
        const uint64_t *t = trace + inst * (size_t)ZKC_VM_NUM_COLS * limit;
        const zkc_vm_state *snaps = snapshots + inst * (limit + 1);
        uint64_t *out = out_all + inst * (size_t)ZKC_VMW_NUM_COLS * limit;
        for (size_t row = 0; row < limit; row++) {
#define T(c) t[(size_t)(c) * limit + row]
            const zkc_vm_state *st = snaps + row, *nx = snaps + row + 1;
            const uint64_t props = T(ZKC_VM_PROPS);
#define TYPE(n) (int)((props >> ZKC_VM_BIT_TYPE(n)) & 1)
            const int capable = TYPE(ZKC_OP_ADD) || TYPE(ZKC_OP_SUB) || TYPE(ZKC_OP_MUL) || TYPE(ZKC_OP_DIV) || TYPE(ZKC_OP_BINOP) ||
                                TYPE(ZKC_OP_SHIFT) || TYPE(ZKC_OP_PTR);   /* the gadgets whose dst0 candidate may go to memory */
            const int update_register = (int)T(ZKC_VM_DST0_UPDATE_REGISTER), memory_write = (int)T(ZKC_VM_PERFORM_DST0_MEMORY_WRITE);
            const int memory_access = (int)T(ZKC_VM_DST0_PERFORMS_MEMORY_ACCESS);
            const int potentially = capable && (update_register || memory_write);   /* :165: one opcode per cycle, so one flag at most */
            const int register_only = !capable && update_register;                 /* :179 */
            PC(ZKC_VMW_DST0_UPDATE_POTENTIALLY_TO_MEMORY, 0) = potentially; PC(ZKC_VMW_CAN_UPDATE_DST0_AS_REGISTER_ONLY, 0) = register_only;
            PC(ZKC_VMW_DST0_PERFORMS_REG_UPDATE, 0) = !memory_access; PC(ZKC_VMW_DST0_REG_UPDATE_T, 0) = !memory_access && potentially;   /* :303-309 */
            /* far call: execute, system_call after the kernel-target mask (far_call.rs:396-431), cleanup_register (:1045-1046), r2 (:1021-1039) */
            const int far_call = TYPE(ZKC_OP_FAR_CALL);
            uint32_t abi[8], target[8];
            for (int i = 0; i < 8; i++) { abi[i] = (uint32_t)T(ZKC_VM_SRC0 + 1 + i); target[i] = (uint32_t)T(ZKC_VM_SRC1 + 1 + i); }
            const int target_is_kernel = (target[0] >> 16) == 0 && target[1] == 0 && target[2] == 0 && target[3] == 0 && target[4] == 0;
            const int constructor_call = ((abi[7] >> 16) & 0xFF) != 0 && (st->current_context.is_kernel_mode & 1);   /* byte 30 */
            const int system_call = ((abi[7] >> 24) & 0xFF) != 0 && target_is_kernel;                                 /* byte 31 */
            const int cleanup = far_call && !system_call;
            const int far_return = TYPE(ZKC_OP_RET) && !(st->current_context.is_local_call & 1);   /* ret.rs:442, call_ret.rs:133 */
            PC(ZKC_VMW_FAR_CALL_UPDATE, 0) = far_call; PC(ZKC_VMW_FAR_CALL_NON_SYSTEM, 0) = !system_call; PC(ZKC_VMW_FAR_CALL_CLEANUP_REGISTER, 0) = cleanup;
            PC(ZKC_VMW_FAR_RETURN_UPDATE, 0) = far_return; PC(ZKC_VMW_FAR_CALL_NEW_R2_LOW, 0) = (uint64_t)constructor_call + 2 * (uint64_t)system_call;
            zkc_vm_register dst0, dst1;
            dst0.is_pointer = (uint32_t)T(ZKC_VM_DST0) & 1; dst1.is_pointer = (uint32_t)T(ZKC_VM_DST1) & 1;
            for (int i = 0; i < 8; i++) { dst0.value[i] = (uint32_t)T(ZKC_VM_DST0 + 1 + i); dst1.value[i] = (uint32_t)T(ZKC_VM_DST1 + 1 + i); }
            const uint32_t dst0_index = (uint32_t)T(ZKC_VM_DST0_REG), dst1_index = (uint32_t)T(ZKC_VM_DST1_REG);
            for (uint32_t r = 0; r < ZKC_VM_REGISTERS; r++) {
                zkc_vm_register reg = st->registers[r];
                reg.is_pointer &= 1;
                const int write0 = update_register && dst0_index == r + 1, write1 = dst1_index == r + 1;   /* :329-330 */
                const int in_abi = r >= isa->call_system_abi_registers[0] && r < isa->call_system_abi_registers[1];
                const int in_reserved = (r >= isa->call_reserved_range[0] && r < isa->call_reserved_range[1]) || r == isa->call_implicit_parameter_reg_idx;
                const int marker = ((in_abi || in_reserved) && far_call) || (r >= 1 && far_return);        /* :357-360 */
                const int zero_out = (in_abi && cleanup) || (in_reserved && far_call) || (r >= 1 && far_return);   /* :367-370 */
                const int far_call_sets = far_call && r < 2, far_return_sets = far_return && r == 0;
                zkc_vm_register new_far = {0, {0}};
                if (far_call_sets && r == 0) new_far = nx->registers[0];          /* final_fat_ptr.into_register: hinted, see the header */
                if (far_call_sets && r == 1) new_far.value[0] = (uint32_t)constructor_call + 2 * (uint32_t)system_call;
                const zkc_vm_register new_ret = far_return_sets ? nx->registers[0] : new_far;   /* only read when far_return_sets */
                const int any0 = write0 || far_call_sets || far_return_sets || marker;   /* :377 */
                const uint32_t is_ptr_as0 = (write0 ? dst0.is_pointer : 0) + (far_call_sets ? (new_far.is_pointer & 1) : 0) +
                                            (far_return_sets ? (new_ret.is_pointer & 1) : 0);   /* + marker * false, :380-388 */
                PC(ZKC_VMW_WRITE_AS_DST0, r) = write0; PC(ZKC_VMW_REMOVE_PTR_MARKER, r) = marker; PC(ZKC_VMW_ZERO_OUT, r) = zero_out;
                PC(ZKC_VMW_ANY_PTR_UPDATE_AS_DST0, r) = any0; PC(ZKC_VMW_IS_PTR_AS_DST0, r) = is_ptr_as0;
                if (any0) reg.is_pointer = is_ptr_as0;                                          /* :389-394 */
                PC(ZKC_VMW_IS_PTR_AFTER_DST0, r) = reg.is_pointer;
                const uint32_t is_ptr_as1 = write1 ? dst1.is_pointer : 0;                       /* :396-404 */
                PC(ZKC_VMW_IS_PTR_AS_DST1, r) = is_ptr_as1;
                if (write1) reg.is_pointer = is_ptr_as1;                                        /* :405-410 */
                PC(ZKC_VMW_IS_PTR_AFTER_DST1, r) = reg.is_pointer;
                /* the value chain, :415-433: dst0, specific updates in push order (far call, then far return: call_ret.rs:436-449), zero-out, dst1 */
                if (write0) memcpy(reg.value, dst0.value, 32);
                for (int i = 0; i < 8; i++) PC(ZKC_VMW_VALUE_AFTER_DST0, 8 * r + i) = reg.value[i];
                if (far_call_sets) memcpy(reg.value, new_far.value, 32);
                if (r < 2) for (int i = 0; i < 8; i++) PC(ZKC_VMW_VALUE_AFTER_FAR_CALL, 8 * r + i) = reg.value[i];
                if (far_return_sets) memcpy(reg.value, new_ret.value, 32);
                if (r == 0) for (int i = 0; i < 8; i++) PC(ZKC_VMW_VALUE_AFTER_FAR_RETURN, i) = reg.value[i];
                if (zero_out) memset(reg.value, 0, 32);
                for (int i = 0; i < 8; i++) PC(ZKC_VMW_VALUE_AFTER_ZERO_OUT, 8 * r + i) = reg.value[i];
                if (write1) memcpy(reg.value, dst1.value, 32);
                for (int i = 0; i < 8; i++) PC(ZKC_VMW_VALUE_AFTER_DST1, 8 * r + i) = reg.value[i];
            }
#undef TYPE
#undef T
        }
    }
}
