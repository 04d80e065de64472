// RAM permutation circuit on sm_100a: witness generation (ram_permutation_entry_point,
// /root/reference/src/ram_permutation/mod.rs:31-210) and constraint re-evaluation of the
// finished trace.  One thread per loop iteration of partial_accumulate_inner (:246); the
// sequential state the reference threads through the loop is recovered row-parallel:
//   - queue heads: from the previous-state column of the raw queue witness
//     (ram_permutation/input.rs:105-116), verified link by link;
//   - previous key / value: the neighbouring row's sorted item;
//   - running grand products and the non-deterministic write counter: one decoupled
//     look-back scan (scan.cuh).
#include "ctx.cuh"
#include "poseidon2.cuh"
#include "scan.cuh"

namespace zkc {

struct RamDev {
    zkc_ram_closed_form io;       // in: H2D copy of the caller's struct; out: fsm output filled in
    zkc_ram_options opt;
    uint64_t n_unsorted, n_sorted, limit;
    // prologue
    uint64_t ch[2][9];
    uint64_t acc0[4];             // rep*2 + side
    uint32_t nnw0, start;
    zkc_queue_state12 uq0, sq0;
    uint64_t commit_obs_in[4], commit_fsm_in[4];
    // row kernel
    uint64_t acc_final[4];
    uint32_t nnw_final, pad0;
    uint64_t head_final[2][12];
    zkc_memory_query last_sorted;
    // status
    unsigned long long first_bad;  // min over failing rows of (row << 16 | checks); ~0 = none
    uint32_t failed_checks;
    uint32_t hint_bad;
    uint32_t prologue_checks, pad1;
    // finalize
    uint64_t commitment[4];
    zkc_status status;
    unsigned long long violations;
};

__device__ __forceinline__ void ram_encode(const zkc_memory_query &q, uint64_t (&e)[8]) {
    // memory_query/mod.rs:103-221
    const uint32_t *v = q.value;
    e[0] = q.timestamp;
    e[1] = q.memory_page;
    e[2] = (uint64_t)q.index | ((uint64_t)(q.rw_flag & 1) << 32) | ((uint64_t)(q.is_ptr & 1) << 33);
    e[3] = (uint64_t)v[0] | ((uint64_t)(v[5] & 0xFFFFFFu) << 32);
    e[4] = (uint64_t)v[1] | ((uint64_t)(v[5] >> 24) << 32) | ((uint64_t)(v[6] & 0xFFFFu) << 40);
    e[5] = (uint64_t)v[2] | ((uint64_t)(v[6] >> 16) << 32) | ((uint64_t)(v[7] & 0xFFu) << 48);
    e[6] = (uint64_t)v[3] | ((uint64_t)(v[7] >> 8) << 32);
    e[7] = v[4];
}

__device__ __forceinline__ zkc_memory_query ram_load_query(const zkc_memory_query *p) {
    zkc_memory_query q;
    const uint4 *s = reinterpret_cast<const uint4 *>(p);
    uint4 *d = reinterpret_cast<uint4 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] = __ldg(s + i);
    return q;
}
__device__ __forceinline__ zkc_memory_query ram_zero_query() {
    zkc_memory_query q;
    uint4 *d = reinterpret_cast<uint4 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] = make_uint4(0, 0, 0, 0);
    return q;
}

__device__ int put_queue_state12(uint64_t *dst, const zkc_queue_state12 &s) {
    for (int i = 0; i < 12; i++) dst[i] = s.head[i];
    for (int i = 0; i < 12; i++) dst[12 + i] = s.tail[i];
    dst[24] = s.length;
    return 25;
}
// CSVarLengthEncodable order of RamPermutationFSMInputOutput, ram_permutation/input.rs:52-62
__device__ int ram_encode_fsm(const zkc_ram_fsm &f, uint64_t *dst) {
    int n = 0;
    dst[n++] = f.lhs_accumulator[0]; dst[n++] = f.lhs_accumulator[1];
    dst[n++] = f.rhs_accumulator[0]; dst[n++] = f.rhs_accumulator[1];
    n += put_queue_state12(dst + n, f.current_unsorted_queue_state);
    n += put_queue_state12(dst + n, f.current_sorted_queue_state);
    for (int i = 0; i < 3; i++) dst[n++] = f.previous_sorting_key[i];
    for (int i = 0; i < 2; i++) dst[n++] = f.previous_full_key[i];
    for (int i = 0; i < 8; i++) dst[n++] = f.previous_value[i];
    dst[n++] = f.previous_is_ptr;
    dst[n++] = f.num_nondeterministic_writes;
    return n;
}

// ---- prologue: FSM start selection, Fiat-Shamir challenges, input commitments ------------------
// three warps, one 16-lane group each, every permutation spread over 12 lanes (poseidon2_permute_coop)
__global__ void ram_prologue_kernel(RamDev *d) {
    __shared__ uint64_t buf[3][72];
    const int warp = threadIdx.x >> 5, i = threadIdx.x & 31;
    if (i >= 16) return;
    const unsigned gm = 0xFFFFu;
    const zkc_ram_closed_form &io = d->io;
    const zkc_ram_input_data &obs = io.observable_input;
    if (warp == 0) {
        if (i == 0) {
            const bool start = io.start_flag != 0;
            d->start = start;
            d->uq0 = start ? obs.unsorted_queue_initial_state : io.hidden_fsm_input.current_unsorted_queue_state;
            d->sq0 = start ? obs.sorted_queue_initial_state : io.hidden_fsm_input.current_sorted_queue_state;
            for (int k = 0; k < 2; k++) {
                d->acc0[k * 2 + 0] = start ? 1 : io.hidden_fsm_input.lhs_accumulator[k];
                d->acc0[k * 2 + 1] = start ? 1 : io.hidden_fsm_input.rhs_accumulator[k];
            }
            d->nnw0 = start ? 0 : io.hidden_fsm_input.num_nondeterministic_writes;
            uint32_t checks = 0;
            for (int k = 0; k < 12; k++)
                if (obs.unsorted_queue_initial_state.head[k] | obs.sorted_queue_initial_state.head[k])
                    checks |= ZKC_RAM_CHK_TRIVIAL_HEAD;
            if (d->uq0.length != d->sq0.length) checks |= ZKC_RAM_CHK_LENGTHS_EQUAL;
            d->prologue_checks = checks;
            // produce_fs_challenges, utils.rs:12-78, over tail || len || tail || len (26 elements)
            uint64_t *in = buf[0];
            for (int k = 0; k < 12; k++) in[k] = obs.unsorted_queue_initial_state.tail[k];
            in[12] = obs.unsorted_queue_initial_state.length;
            for (int k = 0; k < 12; k++) in[13 + k] = obs.sorted_queue_initial_state.tail[k];
            in[25] = obs.sorted_queue_initial_state.length;
        }
        __syncwarp(gm);
        fs_challenges_coop(gm, buf[0], 26, 9, &d->ch[0][0], i);
    } else {
        uint64_t *b = buf[warp];
        int n = 0;
        if (i == 0) {
            if (warp == 1) {
                n = put_queue_state12(b, obs.unsorted_queue_initial_state);
                n += put_queue_state12(b + n, obs.sorted_queue_initial_state);
                b[n++] = obs.non_deterministic_bootloader_memory_snapshot_length;
            } else {
                n = ram_encode_fsm(io.hidden_fsm_input, b);
            }
        }
        __syncwarp(gm);
        n = __shfl_sync(gm, n, 0, 16);
        const uint64_t c = commit_encoding_coop(gm, b, n, i);
        if (i < 4) (warp == 1 ? d->commit_obs_in : d->commit_fsm_in)[i] = c;
    }
}

// sequential reconstruction of the previous-state column when the caller does not supply it (only for callers without a raw
// queue witness): two warps, one hash chain each, every permutation on 12 cooperating lanes (poseidon2_permute_coop)
__global__ void ram_chain_kernel(RamDev *d, const zkc_memory_query *unsorted, const zkc_memory_query *sorted,
                                 uint64_t *uprev, uint64_t *sprev, size_t rows) {
    const int k = threadIdx.x >> 5, i = threadIdx.x & 31;
    if (i >= 16 || k > 1) return;
    const unsigned gm = 0xFFFFu;
    const zkc_memory_query *q = k ? sorted : unsorted;
    uint64_t *out = k ? sprev : uprev;
    const zkc_queue_state12 &q0 = k ? d->sq0 : d->uq0;
    uint64_t x = i < 12 ? q0.head[i] : 0ull;
    for (size_t r = 0; r < rows; r++) {
        if (i < 12) out[12 * r + i] = x;
        zkc_memory_query it = q[r];
        uint64_t e[8];
        ram_encode(it, e);
#pragma unroll
        for (int j = 0; j < 8; j++) if (i == j) x = e[j];
        x = poseidon2_permute_coop(gm, x, i);
    }
}

__device__ __forceinline__ void ram_report(RamDev *d, size_t row, uint32_t checks) {
    if (!checks) return;
    atomicOr(&d->failed_checks, checks);
    atomicMin(&d->first_bad, ((unsigned long long)row << 16) | checks);
}

// ---- the row kernel ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS)
ram_rows_kernel(RamDev *d, const zkc_memory_query *__restrict__ unsorted, const uint64_t *__restrict__ uprev,
                const zkc_memory_query *__restrict__ sorted, const uint64_t *__restrict__ sprev,
                uint64_t *__restrict__ trace, ScanGlobal *sg, TileState *tiles) {
    __shared__ ScanShared sh;
    const unsigned int tile = scan_take_ticket(sg, sh);
    const size_t limit = d->limit;
    const size_t row = (size_t)tile * SCAN_THREADS + threadIdx.x;
    const bool in_range = row < limit;
    const uint32_t ulen0 = d->uq0.length, slen0 = d->sq0.length;
    const bool u_empty = row >= ulen0, s_empty = row >= slen0;
    const bool can_pop = in_range && !u_empty;
    const size_t active_rows = limit < ulen0 ? limit : ulen0;  // rows that pop
    const uint32_t heap_page = d->opt.bootloader_heap_page ? d->opt.bootloader_heap_page : ZKC_BOOTLOADER_HEAP_PAGE_DEFAULT;
    uint32_t checks = 0;
    if (in_range && u_empty != s_empty) checks |= ZKC_RAM_CHK_EMPTY_SYNC;

#define TR(col) trace[(size_t)(col) * limit + row]
    const bool wr = in_range && trace != nullptr;
    zkc_memory_query si = ram_zero_query();
    uint64_t contrib[4];  // rep*2 + side
    // ---- two pops: item, encoding, head <- P(enc || head[8..12]) --------------------------------
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        const zkc_memory_query *recs = k ? sorted : unsorted;
        const uint64_t *prev = k ? sprev : uprev;
        const size_t n_rec = k ? d->n_sorted : d->n_unsorted;
        const zkc_queue_state12 &q0 = k ? d->sq0 : d->uq0;
        zkc_memory_query it = ram_zero_query();
        if (can_pop && row < n_rec) it = ram_load_query(recs + row);
        uint64_t e[8], s[12];
        ram_encode(it, e);
        if (can_pop) {
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = e[i];
            bool hint_ok = true;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const uint64_t h = __ldg(prev + 12 * row + i);
                if (i >= 8) s[i] = h;
                if (row == 0 && h != q0.head[i]) hint_ok = false;
            }
            poseidon2_permute(s);
            if (row + 1 < active_rows) {
#pragma unroll
                for (int i = 0; i < 12; i++) hint_ok &= __ldg(prev + 12 * (row + 1) + i) == s[i];
            } else {
#pragma unroll
                for (int i = 0; i < 12; i++) d->head_final[k][i] = s[i];
            }
            if (!hint_ok) { checks |= ZKC_RAM_CHK_QUEUE_HINT; d->hint_bad = 1; }
        } else {
            // nothing popped: the head keeps its value.  Rows past the end of the queue see the
            // fully drained queue, whose head equals its tail (enforce_consistency, :161-162)
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = ulen0 == 0 ? q0.head[i] : q0.tail[i];
        }
        if (wr) {
            const int base = k ? ZKC_RAM_SORTED_ITEM : ZKC_RAM_UNSORTED_ITEM;
            TR(base + 0) = it.timestamp; TR(base + 1) = it.memory_page; TR(base + 2) = it.index;
            TR(base + 3) = it.rw_flag & 1; TR(base + 4) = it.is_ptr & 1;
#pragma unroll
            for (int i = 0; i < 8; i++) TR(base + 5 + i) = it.value[i];
#pragma unroll
            for (int i = 0; i < 8; i++) TR(base + 13 + i) = e[i];
#pragma unroll
            for (int i = 0; i < 12; i++) TR(base + 21 + i) = s[i];
            const uint32_t len0 = k ? slen0 : ulen0;
            const size_t popped_now = row + 1 < active_rows ? row + 1 : active_rows;
            TR(base + 33) = len0 >= popped_now ? len0 - (uint32_t)popped_now : 0;
            const int bytes = k ? ZKC_RAM_SORTED_ENC_BYTES : ZKC_RAM_UNSORTED_ENC_BYTES;  // memory_query/mod.rs:133-135
#pragma unroll
            for (int l = 0; l < 3; l++)
#pragma unroll
                for (int b = 0; b < 4; b++) TR(bytes + 4 * l + b) = (it.value[5 + l] >> (8 * b)) & 0xFFu;
        }
        // utils.rs:104-129 contribution chains
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            uint64_t c = d->ch[rep][8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                c = gl_fma(e[i], d->ch[rep][i], c);
                if (wr) TR(ZKC_RAM_GP_CHAIN + (rep * 2 + k) * 8 + i) = c;
            }
            contrib[rep * 2 + k] = c;
        }
        if (k == 1) si = it;
    }

    // ---- :260-290 non-deterministic writes ------------------------------------------------------
    const bool ts_is_zero = si.timestamp == 0;
    const bool page_is_heap = si.memory_page == heap_page;
    const bool is_write = si.rw_flag & 1, is_ptr = si.is_ptr & 1;
    const bool is_nondet = can_pop && ts_is_zero && page_is_heap && is_write && !is_ptr;

    // ---- :296-362 ordering and read/write consistency against the previous row ------------------
    uint32_t prev_sk[3], prev_fk[2], prev_val[8], prev_is_ptr;
    if (row == 0) {
        const zkc_ram_fsm &f = d->io.hidden_fsm_input;
#pragma unroll
        for (int i = 0; i < 3; i++) prev_sk[i] = f.previous_sorting_key[i];
        prev_fk[0] = f.previous_full_key[0]; prev_fk[1] = f.previous_full_key[1];
#pragma unroll
        for (int i = 0; i < 8; i++) prev_val[i] = f.previous_value[i];
        prev_is_ptr = f.previous_is_ptr & 1;
    } else {
        zkc_memory_query pq = ram_zero_query();
        if (in_range && row - 1 < active_rows && row - 1 < d->n_sorted) pq = ram_load_query(sorted + row - 1);
        prev_sk[0] = pq.timestamp; prev_sk[1] = pq.index; prev_sk[2] = pq.memory_page;
        prev_fk[0] = pq.index; prev_fk[1] = pq.memory_page;
#pragma unroll
        for (int i = 0; i < 8; i++) prev_val[i] = pq.value[i];
        prev_is_ptr = pq.is_ptr & 1;
    }
    const uint32_t sk[3] = {si.timestamp, si.index, si.memory_page};
    uint32_t diff[3], bor[3];
    uint32_t borrow = 0;
    bool keys_equal = true;
#pragma unroll
    for (int i = 0; i < 3; i++) {  // previous - current, least significant limb first
        const uint64_t dd = (uint64_t)prev_sk[i] - sk[i] - borrow;
        diff[i] = (uint32_t)dd;
        borrow = (uint32_t)(dd >> 32) & 1u;
        bor[i] = borrow;
        keys_equal &= diff[i] == 0;
    }
    const bool prev_smaller = borrow;
    const bool not_start = !d->start;
    const bool first = row == 0;
    if (can_pop && (!first || not_start) && !prev_smaller) checks |= ZKC_RAM_CHK_ASCENDING;
    const bool same_cell = si.index == prev_fk[0] && si.memory_page == prev_fk[1];
    bool value_equal = true, value_is_zero = true;
#pragma unroll
    for (int i = 0; i < 8; i++) { value_equal &= si.value[i] == prev_val[i]; value_is_zero &= si.value[i] == 0; }
    const bool not_rw = !is_write;
    const bool is_zero = value_is_zero && !is_ptr;
    const bool ptr_equality = prev_is_ptr == (uint32_t)is_ptr;
    const bool value_and_ptr_equal = value_equal && ptr_equality;
    bool read_uninit, check_equality;
    if (!first) {
        read_uninit = !same_cell && not_rw;
        check_equality = same_cell && not_rw;
    } else {
        read_uninit = (not_start && !same_cell && not_rw) || (!not_start && not_rw);
        check_equality = same_cell && not_rw && not_start;
    }
    if (in_range && read_uninit && !is_zero) checks |= ZKC_RAM_CHK_UNINIT_READ_ZERO;
    if (in_range && check_equality && !value_and_ptr_equal) checks |= ZKC_RAM_CHK_READ_CONSISTENT;

    // ---- running products + write counter ---------------------------------------------------------
    ScanVal v = scan_identity();
    if (can_pop) {
#pragma unroll
        for (int i = 0; i < 4; i++) v.p[i] = contrib[i];
    }
    v.c = is_nondet;
    ScanVal init;
#pragma unroll
    for (int i = 0; i < 4; i++) init.p[i] = d->acc0[i];
    init.c = d->nnw0;
    ScanVal incl;
    const ScanVal excl = scan_tile(v, tile, init, tiles, sh, incl);

    if (wr) {
        TR(ZKC_RAM_UNSORTED_IS_EMPTY) = u_empty; TR(ZKC_RAM_SORTED_IS_EMPTY) = s_empty; TR(ZKC_RAM_CAN_POP) = can_pop;
        TR(ZKC_RAM_TS_IS_ZERO) = ts_is_zero; TR(ZKC_RAM_PAGE_IS_BOOTLOADER_HEAP) = page_is_heap;
        TR(ZKC_RAM_IS_NONDET_WRITE) = is_nondet; TR(ZKC_RAM_NUM_NONDET_WRITES) = incl.c;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            TR(ZKC_RAM_CMP_DIFF + i) = diff[i]; TR(ZKC_RAM_CMP_BORROW + i) = bor[i];
            TR(ZKC_RAM_CMP_LIMB_EQ + i) = diff[i] == 0;
        }
        TR(ZKC_RAM_KEYS_EQUAL) = keys_equal; TR(ZKC_RAM_PREV_KEY_SMALLER) = prev_smaller;
        TR(ZKC_RAM_SAME_CELL) = same_cell; TR(ZKC_RAM_VALUE_EQUAL) = value_equal;
        TR(ZKC_RAM_VALUE_IS_ZERO) = value_is_zero; TR(ZKC_RAM_IS_ZERO) = is_zero;
        TR(ZKC_RAM_PTR_EQUALITY) = ptr_equality; TR(ZKC_RAM_VALUE_AND_PTR_EQUAL) = value_and_ptr_equal;
        TR(ZKC_RAM_READ_UNINIT) = read_uninit; TR(ZKC_RAM_CHECK_EQUALITY) = check_equality;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            TR(ZKC_RAM_GP_NEW + i) = can_pop ? incl.p[i] : gl_mul(excl.p[i], contrib[i]);
            TR(ZKC_RAM_GP_ACC + i) = incl.p[i];
        }
        // differences / per-limb flags of the equality gadgets; their inverse witnesses: ram_inverse_kernel
        TR(ZKC_RAM_PAGE_DIFF) = gl_sub(si.memory_page, heap_page);
        TR(ZKC_RAM_CELL_DIFF + 0) = gl_sub(si.index, prev_fk[0]); TR(ZKC_RAM_CELL_DIFF + 1) = gl_sub(si.memory_page, prev_fk[1]);
        TR(ZKC_RAM_CELL_LIMB_EQ + 0) = si.index == prev_fk[0]; TR(ZKC_RAM_CELL_LIMB_EQ + 1) = si.memory_page == prev_fk[1];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            TR(ZKC_RAM_VALUE_DIFF + i) = gl_sub(si.value[i], prev_val[i]); TR(ZKC_RAM_VALUE_LIMB_EQ + i) = si.value[i] == prev_val[i];
            TR(ZKC_RAM_VALUE_ZERO_DIFF + i) = si.value[i]; TR(ZKC_RAM_VALUE_ZERO_LIMB_EQ + i) = si.value[i] == 0;
        }
        TR(ZKC_RAM_PTR_DIFF) = gl_sub(prev_is_ptr, (uint32_t)is_ptr);
    }
    if (in_range && row == limit - 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) d->acc_final[i] = incl.p[i];
        d->nnw_final = incl.c;
        d->last_sorted = si;
    }
    if (in_range) ram_report(d, row, checks);
#undef TR
}

// ---- inverse witnesses of the zero checks (ZeroCheckGate: x * inv = 1 - is_zero) -------------------------------------------
// 28 field inversions per row as ONE: Montgomery's trick over the row's values (zeros replaced by one and mapped back to a
// zero witness), x^(p - 2) by an addition chain.  Reads the values from the trace the row kernel wrote.
constexpr int RAM_INV_N = 28;
__device__ constexpr int RAM_INV_SRC[RAM_INV_N] = {
    -1, -2, ZKC_RAM_SORTED_ITEM + 0, ZKC_RAM_PAGE_DIFF, ZKC_RAM_CMP_DIFF, ZKC_RAM_CMP_DIFF + 1, ZKC_RAM_CMP_DIFF + 2, ZKC_RAM_CELL_DIFF,
    ZKC_RAM_CELL_DIFF + 1, ZKC_RAM_VALUE_DIFF, ZKC_RAM_VALUE_DIFF + 1, ZKC_RAM_VALUE_DIFF + 2, ZKC_RAM_VALUE_DIFF + 3, ZKC_RAM_VALUE_DIFF + 4,
    ZKC_RAM_VALUE_DIFF + 5, ZKC_RAM_VALUE_DIFF + 6, ZKC_RAM_VALUE_DIFF + 7, ZKC_RAM_VALUE_ZERO_DIFF, ZKC_RAM_VALUE_ZERO_DIFF + 1,
    ZKC_RAM_VALUE_ZERO_DIFF + 2, ZKC_RAM_VALUE_ZERO_DIFF + 3, ZKC_RAM_VALUE_ZERO_DIFF + 4, ZKC_RAM_VALUE_ZERO_DIFF + 5,
    ZKC_RAM_VALUE_ZERO_DIFF + 6, ZKC_RAM_VALUE_ZERO_DIFF + 7, ZKC_RAM_PTR_DIFF, -3, -3};
__device__ constexpr int RAM_INV_DST[RAM_INV_N] = {
    ZKC_RAM_UNSORTED_LEN_INV, ZKC_RAM_SORTED_LEN_INV, ZKC_RAM_TS_INV, ZKC_RAM_PAGE_DIFF_INV, ZKC_RAM_CMP_DIFF_INV, ZKC_RAM_CMP_DIFF_INV + 1,
    ZKC_RAM_CMP_DIFF_INV + 2, ZKC_RAM_CELL_DIFF_INV, ZKC_RAM_CELL_DIFF_INV + 1, ZKC_RAM_VALUE_DIFF_INV, ZKC_RAM_VALUE_DIFF_INV + 1,
    ZKC_RAM_VALUE_DIFF_INV + 2, ZKC_RAM_VALUE_DIFF_INV + 3, ZKC_RAM_VALUE_DIFF_INV + 4, ZKC_RAM_VALUE_DIFF_INV + 5, ZKC_RAM_VALUE_DIFF_INV + 6,
    ZKC_RAM_VALUE_DIFF_INV + 7, ZKC_RAM_VALUE_ZERO_DIFF_INV, ZKC_RAM_VALUE_ZERO_DIFF_INV + 1, ZKC_RAM_VALUE_ZERO_DIFF_INV + 2,
    ZKC_RAM_VALUE_ZERO_DIFF_INV + 3, ZKC_RAM_VALUE_ZERO_DIFF_INV + 4, ZKC_RAM_VALUE_ZERO_DIFF_INV + 5, ZKC_RAM_VALUE_ZERO_DIFF_INV + 6,
    ZKC_RAM_VALUE_ZERO_DIFF_INV + 7, ZKC_RAM_PTR_DIFF_INV, -1, -1};
__global__ void __launch_bounds__(128)
ram_inverse_kernel(const RamDev *d, uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const uint32_t ulen0 = d->uq0.length, slen0 = d->sq0.length;
    const size_t active_rows = limit < ulen0 ? limit : ulen0, popped_before = row < active_rows ? row : active_rows;
    uint64_t x[RAM_INV_N], pre[RAM_INV_N];
#pragma unroll
    for (int i = 0; i < RAM_INV_N; i++) {
        if (RAM_INV_SRC[i] == -1) x[i] = ulen0 - (uint32_t)popped_before;                           // queue lengths before the pop
        else if (RAM_INV_SRC[i] == -2) x[i] = slen0 >= popped_before ? slen0 - (uint32_t)popped_before : 0;
        else if (RAM_INV_SRC[i] == -3) x[i] = 1;                                                     // padding of the batch
        else x[i] = trace[(size_t)RAM_INV_SRC[i] * limit + row];
    }
    uint64_t acc = 1;
#pragma unroll
    for (int i = 0; i < RAM_INV_N; i++) { pre[i] = acc; acc = gl_mul(acc, x[i] ? x[i] : 1ull); }
    acc = gl_inv(acc);
#pragma unroll
    for (int i = RAM_INV_N - 1; i >= 0; i--) {
        const uint64_t inv = gl_mul(acc, pre[i]);
        acc = gl_mul(acc, x[i] ? x[i] : 1ull);
        if (RAM_INV_DST[i] >= 0) trace[(size_t)RAM_INV_DST[i] * limit + row] = x[i] ? inv : 0ull;
    }
}

// FullStateCircuitQueue::push of whole queues (the reference test builds its inputs this way,
// ram_permutation/mod.rs:506-515): one thread per independent queue, each a sequential hash chain.
__global__ void memory_queue_simulate_kernel(const zkc_memory_query *__restrict__ recs, size_t n_per_queue,
                                             size_t n_queues, uint64_t *__restrict__ prev_states,
                                             zkc_queue_state12 *__restrict__ final_states) {
    const size_t qi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= n_queues) return;
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    for (size_t r = 0; r < n_per_queue; r++) {
        const size_t g = qi * n_per_queue + r;
        if (prev_states) {
#pragma unroll
            for (int i = 0; i < 12; i++) prev_states[12 * g + i] = s[i];
        }
        zkc_memory_query it = ram_load_query(recs + g);
        uint64_t e[8];
        ram_encode(it, e);
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = e[i];
        poseidon2_permute(s);
    }
    zkc_queue_state12 &o = final_states[qi];
#pragma unroll
    for (int i = 0; i < 12; i++) { o.head[i] = 0; o.tail[i] = s[i]; }
    o.length = (uint32_t)n_per_queue;
    o._pad = 0;
}

// ---- finalize: entry-point enforcements, FSM output, commitment ---------------------------------
__global__ void ram_finalize_kernel(RamDev *d) {
    // lane 0 does the scalar bookkeeping, the permutations of the two commitments run on 12 lanes
    __shared__ uint64_t e_out[72], compact[24];
    __shared__ uint32_t sh_completed;
    const int lane = threadIdx.x & 31;
    if (lane >= 16) return;
    const unsigned gm = 0xFFFFu;
    if (lane == 0) {
    zkc_ram_closed_form &io = d->io;
    const size_t limit = d->limit;
    const uint32_t len0 = d->uq0.length;
    const size_t popped = limit < len0 ? limit : len0;
    zkc_ram_fsm out;
    memset(&out, 0, sizeof out);
    out.current_unsorted_queue_state = d->uq0;
    out.current_sorted_queue_state = d->sq0;
    if (popped > 0) {
        for (int i = 0; i < 12; i++) {
            out.current_unsorted_queue_state.head[i] = d->head_final[0][i];
            out.current_sorted_queue_state.head[i] = d->head_final[1][i];
        }
    }
    out.current_unsorted_queue_state.length = len0 - (uint32_t)popped;
    const size_t spopped = d->sq0.length < popped ? d->sq0.length : popped;
    out.current_sorted_queue_state.length = d->sq0.length - (uint32_t)spopped;
    if (limit > 0) {
        for (int i = 0; i < 2; i++) {
            out.lhs_accumulator[i] = d->acc_final[i * 2 + 0];
            out.rhs_accumulator[i] = d->acc_final[i * 2 + 1];
        }
        out.num_nondeterministic_writes = d->nnw_final;
        const zkc_memory_query &q = d->last_sorted;
        out.previous_sorting_key[0] = q.timestamp; out.previous_sorting_key[1] = q.index; out.previous_sorting_key[2] = q.memory_page;
        out.previous_full_key[0] = q.index; out.previous_full_key[1] = q.memory_page;
        for (int i = 0; i < 8; i++) out.previous_value[i] = q.value[i];
        out.previous_is_ptr = q.is_ptr & 1;
    } else {
        for (int i = 0; i < 2; i++) {
            out.lhs_accumulator[i] = d->acc0[i * 2 + 0];
            out.rhs_accumulator[i] = d->acc0[i * 2 + 1];
        }
        out.num_nondeterministic_writes = d->nnw0;
        for (int i = 0; i < 3; i++) out.previous_sorting_key[i] = io.hidden_fsm_input.previous_sorting_key[i];
        for (int i = 0; i < 2; i++) out.previous_full_key[i] = io.hidden_fsm_input.previous_full_key[i];
        for (int i = 0; i < 8; i++) out.previous_value[i] = io.hidden_fsm_input.previous_value[i];
        out.previous_is_ptr = io.hidden_fsm_input.previous_is_ptr;
    }
    uint32_t checks = 0;
    // :161-162
    const zkc_queue_state12 *qs[2] = {&out.current_unsorted_queue_state, &out.current_sorted_queue_state};
    for (int k = 0; k < 2; k++)
        if (qs[k]->length == 0)
            for (int i = 0; i < 12; i++)
                if (qs[k]->head[i] != qs[k]->tail[i]) checks |= ZKC_RAM_CHK_QUEUE_CONSISTENCY;
    const bool completed = out.current_unsorted_queue_state.length == 0;  // :164
    if (completed) {
        for (int i = 0; i < 2; i++)
            if (out.lhs_accumulator[i] != out.rhs_accumulator[i]) checks |= ZKC_RAM_CHK_GRAND_PRODUCT;  // :166-168
        if (out.num_nondeterministic_writes != io.observable_input.non_deterministic_bootloader_memory_snapshot_length)
            checks |= ZKC_RAM_CHK_NONDET_COUNT;  // :170-175
    }
    checks |= d->failed_checks | d->prologue_checks;
    uint64_t e_exp[69];
    const int n_out = ram_encode_fsm(out, e_out);
    zkc_status st;
    st.code = ZKC_OK; st.cuda_error = 0; st.first_bad_row = -1; st.failed_checks = checks; st.reserved = 0;
    if (d->first_bad != ~0ull) st.first_bad_row = (int64_t)(d->first_bad >> 16);
    if (checks) st.code = ZKC_ERR_UNSATISFIED;
    if (d->hint_bad) st.code = ZKC_ERR_QUEUE_WITNESS_INCONSISTENT;
    if (d->opt.compare_expected) {  // hook_compare_witness, fsm_input_output/mod.rs:102-133
        ram_encode_fsm(io.hidden_fsm_output, e_exp);
        bool same = (io.completion_flag != 0) == completed;
        for (int i = 0; i < n_out; i++) same &= e_out[i] == e_exp[i];
        if (!same && st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
    }
    io.hidden_fsm_output = out;
    io.completion_flag = completed;
    // ClosedFormInputCompactForm::from_full_form + commitment, fsm_input_output/mod.rs:178-255
    compact[0] = d->start; compact[1] = completed;
    for (int i = 0; i < 4; i++) {
        compact[2 + i] = d->commit_obs_in[i];
        compact[6 + i] = 0;  // observable output is `()`: commitment of the empty encoding is 0, and masked unless completed
        compact[10 + i] = d->start ? 0 : d->commit_fsm_in[i];
    }
    d->status = st;
    sh_completed = completed;
    }
    __syncwarp(gm);
    const uint64_t c_out = commit_encoding_coop(gm, e_out, 69, lane);
    if (lane < 4) compact[14 + lane] = sh_completed ? 0 : c_out;
    __syncwarp(gm);
    const uint64_t f = commit_encoding_coop(gm, compact, 18, lane);
    if (lane < 4) d->commitment[lane] = f;
}


// ---- constraint evaluation of a finished trace --------------------------------------------------
// One thread per row re-evaluates every relation the loop body of partial_accumulate_inner places
// (boolean / range, queue length bookkeeping, MemoryQuery::encode packing, UIntXAddGate borrow
// chain, zero-check / equality flags, conditional enforcements, the FMA chain and the accumulator
// update; with ZKC_GATES_ROUND_FUNCTION also the Poseidon2 link of both queue heads).  Streaming
// read of all ZKC_RAM_NUM_COLS columns (+ the previous row of the carried ones, an L1/L2 hit).
enum : uint32_t {
    RAMV_BOOLEAN = ZKC_RAMV_BOOLEAN, RAMV_QUEUE_LEN = ZKC_RAMV_QUEUE_LEN, RAMV_ENCODING = ZKC_RAMV_ENCODING,
    RAMV_ROUND_FUNCTION = ZKC_RAMV_ROUND_FUNCTION, RAMV_NONDET = ZKC_RAMV_NONDET, RAMV_COMPARISON = ZKC_RAMV_COMPARISON,
    RAMV_FLAGS = ZKC_RAMV_FLAGS, RAMV_ENFORCE = ZKC_RAMV_ENFORCE, RAMV_GP_CHAIN = ZKC_RAMV_GP_CHAIN, RAMV_GP_ACC = ZKC_RAMV_GP_ACC,
};

#ifndef RAM_CHECK_THREADS
#define RAM_CHECK_THREADS 128
#endif
#ifndef RAM_CHECK_MIN_BLOCKS
#define RAM_CHECK_MIN_BLOCKS 2  // measured per 2^20 rows (profiles/README.md): 2 CTAs / SM 0.51 ms, 3: 0.53 - 0.54, 4 (spills): 0.61; one row per thread at 4 / 5 / 6 CTAs: 0.60 / 0.55 / 0.55
#endif
#ifndef RAM_CHECK_PREFETCH
#define RAM_CHECK_PREFETCH 0  // measured slower on B200 (0.78 vs 0.53 ms per 2^20 rows): kept for the record
#endif
#ifndef RAM_CHECK_GP_MERGE
#define RAM_CHECK_GP_MERGE 0
#endif
#ifndef RAM_CHECK_PAIRS
#define RAM_CHECK_PAIRS 1
#endif
// R rows per thread.  R = 2: the thread owns the row pair (2k, 2k + 1) and reads every column with ONE 128-bit load (LDG.E.128);
// the value a relation takes from the previous row is the pair's other element, or one 8-byte load of row 2k - 1 (the
// neighbouring thread's line: an L1 hit).  R = 1 (odd limit, or with the Poseidon2 link, whose 12-element states would not fit
// twice): 64-bit loads.  The relations run in stages -- flags, the two queue pops (item, packing, byte decomposition, the FMA
// chains, the head), ordering, value flags, zero-check witnesses -- and the loads of a stage are not hoisted above the previous
// one (STAGE): the compiler keeps a stage's loads together and in flight while the previous stage's field arithmetic runs.
template <int R> struct RamCells { uint64_t v[R]; };
template <int R> __device__ __forceinline__ RamCells<R> ram_ld(const uint64_t *p);
template <> __device__ __forceinline__ RamCells<1> ram_ld<1>(const uint64_t *p) { return RamCells<1>{{__ldg(p)}}; }
template <> __device__ __forceinline__ RamCells<2> ram_ld<2>(const uint64_t *p) {
    const ulonglong2 q = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    return RamCells<2>{{q.x, q.y}};
}

template <bool ROUND_FUNCTION, int R>
__global__ void __launch_bounds__(RAM_CHECK_THREADS, RAM_CHECK_MIN_BLOCKS)
ram_check_kernel(RamDev *d, const uint64_t *__restrict__ trace) {
    __shared__ uint64_t ch[2][9];
    if (threadIdx.x < 18) ch[threadIdx.x / 9][threadIdx.x % 9] = d->ch[threadIdx.x / 9][threadIdx.x % 9];
    __syncthreads();
    const size_t limit = d->limit;
    const size_t row0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * R;  // R == 2: limit is even
    if (row0 >= limit) return;
    const bool first = row0 == 0;
    const bool start = d->start;
    const uint64_t *t = trace + row0;
    typedef RamCells<R> Cells;
#define LD(col) ram_ld<R>(t + (size_t)(col) * limit)
#define PRV(col, at_first) (first ? (uint64_t)(at_first) : __ldg(t + (size_t)(col) * limit - 1))  // the cell of row0 - 1
#define FOR_R _Pragma("unroll") for (int r = 0; r < R; r++)
#define STAGE asm volatile("" ::: "memory")
#if RAM_CHECK_PREFETCH
    // L2 prefetch of the columns two stages ahead: the loads of a stage then cost an L2 round trip instead of an HBM one, with no
    // registers held while the line is on its way
#define PF(col0, n) _Pragma("unroll 1") for (int c_ = 0; c_ < (n); c_++) asm volatile("prefetch.global.L2 [%0];" ::"l"(t + (size_t)((col0) + c_) * limit))
#else
#define PF(col0, n)
#endif
    PF(ZKC_RAM_UNSORTED_IS_EMPTY, 3 + 13 + 8); PF(ZKC_RAM_UNSORTED_LEN, 1); PF(ZKC_RAM_UNSORTED_LEN_INV, 1); PF(ZKC_RAM_UNSORTED_ENC_BYTES, 12);
    PF(ZKC_RAM_GP_CHAIN, 8); PF(ZKC_RAM_GP_CHAIN + 16, 8); PF(ZKC_RAM_GP_NEW, 8);
    PF(ZKC_RAM_UNSORTED_HEAD, 12);
    uint32_t bad[R];
    uint64_t u_empty[R], s_empty[R], can_pop[R];
    {
        const Cells ue = LD(ZKC_RAM_UNSORTED_IS_EMPTY), se = LD(ZKC_RAM_SORTED_IS_EMPTY), cp = LD(ZKC_RAM_CAN_POP);
        FOR_R {
            u_empty[r] = ue.v[r]; s_empty[r] = se.v[r]; can_pop[r] = cp.v[r];
            bad[r] = ((u_empty[r] | s_empty[r] | can_pop[r]) > 1 || u_empty[r] != s_empty[r] || can_pop[r] != 1 - u_empty[r]) ? RAMV_BOOLEAN : 0u;
        }
    }
    const uint32_t heap_page = d->opt.bootloader_heap_page ? d->opt.bootloader_heap_page : ZKC_BOOTLOADER_HEAP_PAGE_DEFAULT;
    uint32_t it[R][13];     // u32 cells (range-checked below); the sorted item survives the loop
    bool cells_ok[R];
    FOR_R cells_ok[r] = true;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int base = k ? ZKC_RAM_SORTED_ITEM : ZKC_RAM_UNSORTED_ITEM;
        const zkc_queue_state12 &q0 = k ? d->sq0 : d->uq0;
        if (k == 0) { PF(ZKC_RAM_SORTED_ITEM, 13 + 8); PF(ZKC_RAM_SORTED_LEN, 1); PF(ZKC_RAM_SORTED_LEN_INV, 1); PF(ZKC_RAM_SORTED_ENC_BYTES, 12); }
        else { PF(ZKC_RAM_TS_IS_ZERO, 4); PF(ZKC_RAM_TS_INV, 3); PF(ZKC_RAM_CMP_DIFF, 11); PF(ZKC_RAM_CMP_DIFF_INV, 3); }
        {
            uint64_t range[R];
            FOR_R range[r] = 0;
#pragma unroll
            for (int i = 0; i < 13; i++) {
                const Cells c = LD(base + i);
                FOR_R { it[r][i] = (uint32_t)c.v[r]; range[r] |= c.v[r]; }
            }
            FOR_R if ((range[r] >> 32) || (it[r][3] | it[r][4]) > 1) bad[r] |= RAMV_BOOLEAN;
            // queue length: is_empty <=> previous length == 0, length decrements on a pop
            const Cells len = LD(base + 33), inv = LD(k ? ZKC_RAM_SORTED_LEN_INV : ZKC_RAM_UNSORTED_LEN_INV);
            const uint64_t p0 = PRV(base + 33, q0.length);
            FOR_R {
                const uint64_t lp = r ? len.v[r ? r - 1 : 0] : p0;
                if ((k ? s_empty[r] : u_empty[r]) != (uint64_t)(lp == 0) || len.v[r] + can_pop[r] != lp) bad[r] |= RAMV_QUEUE_LEN;
                const uint64_t flag = k ? s_empty[r] : u_empty[r];  // :247 / :248 is_empty: the inverse witness of the length before the pop
                cells_ok[r] &= flag <= 1 && gl_mul(lp, inv.v[r]) == 1 - flag && (flag == 0 || lp == 0);
            }
        }
        // MemoryQuery::encode, memory_query/mod.rs:103-221, and the byte decomposition of value limbs 5, 6, 7 (:133-135)
        uint64_t enc[R][8];
        {
            uint64_t e[R][8];
            FOR_R {
                zkc_memory_query q;
                q.timestamp = (uint32_t)it[r][0]; q.memory_page = (uint32_t)it[r][1]; q.index = (uint32_t)it[r][2];
                q.rw_flag = (uint32_t)it[r][3]; q.is_ptr = (uint32_t)it[r][4];
#pragma unroll
                for (int i = 0; i < 8; i++) q.value[i] = (uint32_t)it[r][5 + i];
                ram_encode(q, e[r]);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const Cells c = LD(base + 13 + i);
                FOR_R { enc[r][i] = c.v[r]; if (c.v[r] != e[r][i]) bad[r] |= RAMV_ENCODING; }
            }
            const int bytes = k ? ZKC_RAM_SORTED_ENC_BYTES : ZKC_RAM_UNSORTED_ENC_BYTES;
#pragma unroll
            for (int l = 0; l < 3; l++) {
                uint64_t limb[R], range[R];
                FOR_R { limb[r] = 0; range[r] = 0; }
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const Cells c = LD(bytes + 4 * l + b);
                    FOR_R { range[r] |= c.v[r]; limb[r] |= c.v[r] << (8 * b); }
                }
                FOR_R cells_ok[r] &= (range[r] >> 8) == 0 && limb[r] == it[r][5 + 5 + l];
            }
        }
        STAGE;
        if (k == 0) { PF(ZKC_RAM_GP_CHAIN + 8, 8); PF(ZKC_RAM_GP_CHAIN + 24, 8); PF(ZKC_RAM_SORTED_HEAD, 12); }
        else { PF(ZKC_RAM_SAME_CELL, 8); PF(ZKC_RAM_CELL_DIFF, 6); PF(ZKC_RAM_VALUE_DIFF, 24); }
        // utils.rs:104-135: the two FMA chains over this queue's encoding and the accumulator update
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            const int g = rep * 2 + k;
            uint64_t c[R];
            FOR_R c[r] = ch[rep][8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const Cells cell = LD(ZKC_RAM_GP_CHAIN + g * 8 + i);
                FOR_R { if (cell.v[r] != gl_fma(enc[r][i], ch[rep][i], c[r])) bad[r] |= RAMV_GP_CHAIN; c[r] = cell.v[r]; }
            }
            const Cells nw = LD(ZKC_RAM_GP_NEW + g), acc = LD(ZKC_RAM_GP_ACC + g);
            const uint64_t p0 = PRV(ZKC_RAM_GP_ACC + g, d->acc0[g]);
            FOR_R {
                const uint64_t acc_prev = r ? acc.v[r ? r - 1 : 0] : p0;
                if (nw.v[r] != gl_mul(acc_prev, c[r]) || acc.v[r] != (can_pop[r] ? nw.v[r] : acc_prev)) bad[r] |= RAMV_GP_ACC;
            }
#if !RAM_CHECK_GP_MERGE
            STAGE;
#endif
        }
        STAGE;
        // head' = can_pop ? P(enc || head[8..12]) : head
        {
            bool same[R];
            FOR_R same[r] = true;
            uint64_t st[ROUND_FUNCTION ? R : 1][12], hcur[ROUND_FUNCTION ? R : 1][12];
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const Cells c = LD(base + 21 + i);
                const uint64_t p0 = PRV(base + 21 + i, q0.head[i]);
                FOR_R {
                    const uint64_t hp = r ? c.v[r ? r - 1 : 0] : p0;
                    same[r] &= hp == c.v[r];
                    if (ROUND_FUNCTION) { st[r][i] = i < 8 ? enc[r][i] : hp; hcur[r][i] = c.v[r]; }
                }
            }
            FOR_R {
                if (!can_pop[r] && !same[r]) bad[r] |= RAMV_ROUND_FUNCTION;
                if (ROUND_FUNCTION && can_pop[r]) {
                    poseidon2_permute(st[r]);
#pragma unroll
                    for (int i = 0; i < 12; i++) if (st[r][i] != hcur[r][i]) bad[r] |= RAMV_ROUND_FUNCTION;
                }
            }
        }
        STAGE;
    }
    PF(ZKC_RAM_VALUE_ZERO_DIFF, 24); PF(ZKC_RAM_PTR_DIFF, 2);
    // ---- the row before: the sorted item of row - 1 (the pair's other element, one load per column for the pair's first row) ----
    uint32_t prev_sk[R][3], prev_fk[R][2], prev_val[R][8], prev_is_ptr[R];  // u32 cells of the row before (range-checked on their own row)
    {
        const zkc_ram_fsm &f = d->io.hidden_fsm_input;
        prev_sk[0][0] = (uint32_t)PRV(ZKC_RAM_SORTED_ITEM + 0, f.previous_sorting_key[0]);
        prev_sk[0][1] = (uint32_t)PRV(ZKC_RAM_SORTED_ITEM + 2, f.previous_sorting_key[1]);
        prev_sk[0][2] = (uint32_t)PRV(ZKC_RAM_SORTED_ITEM + 1, f.previous_sorting_key[2]);
        prev_fk[0][0] = first ? f.previous_full_key[0] : prev_sk[0][1];
        prev_fk[0][1] = first ? f.previous_full_key[1] : prev_sk[0][2];
#pragma unroll
        for (int i = 0; i < 8; i++) prev_val[0][i] = (uint32_t)PRV(ZKC_RAM_SORTED_ITEM + 5 + i, f.previous_value[i]);
        prev_is_ptr[0] = (uint32_t)PRV(ZKC_RAM_SORTED_ITEM + 4, f.previous_is_ptr & 1);
#pragma unroll
        for (int r = 1; r < R; r++) {
            prev_sk[r][0] = it[r - 1][0]; prev_sk[r][1] = it[r - 1][2]; prev_sk[r][2] = it[r - 1][1];
            prev_fk[r][0] = it[r - 1][2]; prev_fk[r][1] = it[r - 1][1];
#pragma unroll
            for (int i = 0; i < 8; i++) prev_val[r][i] = it[r - 1][5 + i];
            prev_is_ptr[r] = it[r - 1][4];
        }
    }
    // ---- :260-290 non-deterministic writes; :296-304 borrow chain: prev - cur - borrow_in = diff - 2^32 * borrow_out ------------
    uint64_t prev_smaller[R];
    {
        const Cells tz = LD(ZKC_RAM_TS_IS_ZERO), ph = LD(ZKC_RAM_PAGE_IS_BOOTLOADER_HEAP), nd = LD(ZKC_RAM_IS_NONDET_WRITE), nnw = LD(ZKC_RAM_NUM_NONDET_WRITES);
        const uint64_t nnw0 = PRV(ZKC_RAM_NUM_NONDET_WRITES, d->nnw0);
        const Cells tsi = LD(ZKC_RAM_TS_INV), pd = LD(ZKC_RAM_PAGE_DIFF), pdi = LD(ZKC_RAM_PAGE_DIFF_INV);
        FOR_R {
            auto zero_check = [&](uint64_t x, uint64_t inv, uint64_t flag) { return flag <= 1 && gl_mul(x, inv) == 1 - flag && (flag == 0 || x == 0); };
            const uint64_t rw = it[r][3], is_ptr = it[r][4];
            const uint64_t nnw_prev = r ? nnw.v[r ? r - 1 : 0] : nnw0;
            if (tz.v[r] != (uint64_t)(it[r][0] == 0) || ph.v[r] != (uint64_t)(it[r][1] == heap_page) ||
                nd.v[r] != (can_pop[r] & tz.v[r] & ph.v[r] & rw & (1 - is_ptr)) || nnw.v[r] != nnw_prev + nd.v[r])
                bad[r] |= RAMV_NONDET;
            cells_ok[r] &= zero_check(it[r][0], tsi.v[r], tz.v[r]);
            cells_ok[r] &= pd.v[r] == gl_sub(it[r][1], heap_page) && zero_check(pd.v[r], pdi.v[r], ph.v[r]);
        }
    }
    STAGE;
    {
        uint64_t borrow[R], all_eq[R];
        FOR_R { borrow[r] = 0; all_eq[r] = 1; }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const Cells diff = LD(ZKC_RAM_CMP_DIFF + i), bo = LD(ZKC_RAM_CMP_BORROW + i), leq = LD(ZKC_RAM_CMP_LIMB_EQ + i), dinv = LD(ZKC_RAM_CMP_DIFF_INV + i);
            FOR_R {
                const uint64_t df = diff.v[r], b = bo.v[r], le = leq.v[r];
                const uint64_t sk = i == 0 ? it[r][0] : (i == 1 ? it[r][2] : it[r][1]);
                if ((df >> 32) || b > 1 || le != (uint64_t)(df == 0) || (uint64_t)prev_sk[r][i] + (b << 32) != df + sk + borrow[r]) bad[r] |= RAMV_COMPARISON;
                borrow[r] = b;
                all_eq[r] &= le;
                cells_ok[r] &= le <= 1 && gl_mul(df, dinv.v[r]) == 1 - le && (le == 0 || df == 0);
            }
        }
        const Cells ke = LD(ZKC_RAM_KEYS_EQUAL), ps = LD(ZKC_RAM_PREV_KEY_SMALLER);
        FOR_R {
            prev_smaller[r] = ps.v[r];
            if (ke.v[r] != all_eq[r] || ps.v[r] != borrow[r]) bad[r] |= RAMV_COMPARISON;
        }
    }
    STAGE;
    // ---- :318-357 flags and the conditional enforcements :312-316, :336, :340, :351, :356 ---------------------------------------
    uint64_t same_cell[R], value_equal[R], value_is_zero[R], ptr_eq[R];
    {
        const Cells sc = LD(ZKC_RAM_SAME_CELL), ve = LD(ZKC_RAM_VALUE_EQUAL), vz_ = LD(ZKC_RAM_VALUE_IS_ZERO), iz = LD(ZKC_RAM_IS_ZERO);
        const Cells pe = LD(ZKC_RAM_PTR_EQUALITY), vp = LD(ZKC_RAM_VALUE_AND_PTR_EQUAL), ru_ = LD(ZKC_RAM_READ_UNINIT), ce_ = LD(ZKC_RAM_CHECK_EQUALITY);
        FOR_R {
            same_cell[r] = sc.v[r]; value_equal[r] = ve.v[r]; value_is_zero[r] = vz_.v[r]; ptr_eq[r] = pe.v[r];
            const uint64_t rw = it[r][3], is_ptr = it[r][4];
            bool veq = true, vz = true;
#pragma unroll
            for (int i = 0; i < 8; i++) { veq &= it[r][5 + i] == prev_val[r][i]; vz &= it[r][5 + i] == 0; }
            const uint64_t not_rw = 1 - rw, not_start = start ? 0 : 1;
            const bool row_is_first = first && r == 0;
            uint64_t ru, ce;
            if (!row_is_first) { ru = (1 - sc.v[r]) & not_rw; ce = sc.v[r] & not_rw; }
            else { ru = (not_start & (1 - sc.v[r]) & not_rw) | ((1 - not_start) & not_rw); ce = sc.v[r] & not_rw & not_start; }
            if (sc.v[r] != (uint64_t)(it[r][2] == prev_fk[r][0] && it[r][1] == prev_fk[r][1]) || ve.v[r] != (uint64_t)veq ||
                vz_.v[r] != (uint64_t)vz || iz.v[r] != (vz_.v[r] & (1 - is_ptr)) || pe.v[r] != (uint64_t)(prev_is_ptr[r] == is_ptr) ||
                vp.v[r] != (ve.v[r] & pe.v[r]) || ru_.v[r] != ru || ce_.v[r] != ce)
                bad[r] |= RAMV_FLAGS;
            const uint64_t enforce_order = row_is_first ? (can_pop[r] & not_start) : can_pop[r];
            if ((enforce_order & (1 - prev_smaller[r])) | (ru_.v[r] & (1 - iz.v[r])) | (ce_.v[r] & (1 - vp.v[r]))) bad[r] |= RAMV_ENFORCE;
        }
    }
    STAGE;
    // ---- gadget cells: differences and zero-check witnesses (x * inv = 1 - flag, flag * x = 0) of the equality gadgets ---------
    {
        auto zero_check = [&](uint64_t x, uint64_t inv, uint64_t flag) { return flag <= 1 && gl_mul(x, inv) == 1 - flag && (flag == 0 || x == 0); };
        uint64_t cell_and[R], val_and[R], zero_and[R];
        FOR_R { cell_and[r] = 1; val_and[r] = 1; zero_and[r] = 1; }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const Cells df = LD(ZKC_RAM_CELL_DIFF + i), eq = LD(ZKC_RAM_CELL_LIMB_EQ + i), inv = LD(ZKC_RAM_CELL_DIFF_INV + i);
            FOR_R {
                cells_ok[r] &= df.v[r] == gl_sub(i ? it[r][1] : it[r][2], prev_fk[r][i]) && zero_check(df.v[r], inv.v[r], eq.v[r]);
                cell_and[r] &= eq.v[r];
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const Cells df = LD(ZKC_RAM_VALUE_DIFF + i), eq = LD(ZKC_RAM_VALUE_LIMB_EQ + i), inv = LD(ZKC_RAM_VALUE_DIFF_INV + i);
            const Cells zf = LD(ZKC_RAM_VALUE_ZERO_DIFF + i), zq = LD(ZKC_RAM_VALUE_ZERO_LIMB_EQ + i), zinv = LD(ZKC_RAM_VALUE_ZERO_DIFF_INV + i);
            FOR_R {
                cells_ok[r] &= df.v[r] == gl_sub(it[r][5 + i], prev_val[r][i]) && zero_check(df.v[r], inv.v[r], eq.v[r]);
                val_and[r] &= eq.v[r];
                cells_ok[r] &= zf.v[r] == it[r][5 + i] && zero_check(zf.v[r], zinv.v[r], zq.v[r]);
                zero_and[r] &= zq.v[r];
            }
            if (i == 3) STAGE;
        }
        const Cells pdf = LD(ZKC_RAM_PTR_DIFF), pdi = LD(ZKC_RAM_PTR_DIFF_INV);
        FOR_R {
            cells_ok[r] &= pdf.v[r] == gl_sub(prev_is_ptr[r], it[r][4]) && zero_check(pdf.v[r], pdi.v[r], ptr_eq[r]);
            cells_ok[r] &= cell_and[r] == same_cell[r] && val_and[r] == value_equal[r] && zero_and[r] == value_is_zero[r];  // Boolean::multi_and of the limb flags
            if (!cells_ok[r]) bad[r] |= ZKC_RAMV_GADGET_CELLS;
        }
    }
#undef LD
#undef PF
#undef PRV
#undef FOR_R
#undef STAGE
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (bad[r]) {
            atomicAdd(&d->violations, 1ull);
            atomicOr(&d->failed_checks, bad[r]);
            atomicMin(&d->first_bad, ((unsigned long long)(row0 + r) << 16) | bad[r]);
        }
    }
}

}  // namespace zkc

using namespace zkc;

extern "C" int zkc_ram_permutation_entry_point(zkc_ctx *ctx, zkc_ram_closed_form *io, const zkc_memory_query *unsorted,
                                               const uint64_t *unsorted_prev_states, size_t n_unsorted,
                                               const zkc_memory_query *sorted, const uint64_t *sorted_prev_states,
                                               size_t n_sorted, size_t limit, const zkc_ram_options *options,
                                               int on_device, uint64_t *trace, uint64_t commitment[ZKC_COMMITMENT_LEN],
                                               zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !commitment || (n_unsorted && !unsorted) || (n_sorted && !sorted) || limit > 0xFFFFFFFFull) {
        status->code = ZKC_ERR_INVALID_ARGUMENT;
        return ZKC_ERR_INVALID_ARGUMENT;
    }
    const zkc_queue_state12 &uq = io->start_flag ? io->observable_input.unsorted_queue_initial_state
                                                 : io->hidden_fsm_input.current_unsorted_queue_state;
    const size_t need = limit < uq.length ? limit : uq.length;
    if (n_unsorted < need || n_sorted < need) {  // the reference would panic popping an exhausted witness deque
        status->code = ZKC_ERR_INVALID_ARGUMENT;
        return ZKC_ERR_INVALID_ARGUMENT;
    }
    const bool in_dev = on_device & ZKC_INPUTS_ON_DEVICE, trace_dev = on_device & ZKC_TRACE_ON_DEVICE;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const size_t tiles = (limit + SCAN_THREADS - 1) / SCAN_THREADS;
    const bool need_chain = need && (!unsorted_prev_states || !sorted_prev_states);
    size_t bytes = zkc_carver::bytes(1, sizeof(RamDev)) + zkc_carver::bytes(1, sizeof(ScanGlobal)) +
                   zkc_carver::bytes(tiles + 1, sizeof(TileState));
    if (!in_dev) bytes += 2 * zkc_carver::bytes(need, sizeof(zkc_memory_query)) + 2 * zkc_carver::bytes(need * 12, 8);
    else if (need_chain) bytes += 2 * zkc_carver::bytes(need * 12, 8);
    if (trace && !trace_dev) bytes += zkc_carver::bytes((size_t)ZKC_RAM_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    RamDev *h = (RamDev *)ctx->pinned(sizeof(RamDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    RamDev *d = cv.take<RamDev>(1);
    ScanGlobal *sg = cv.take<ScanGlobal>(1);
    TileState *ts = cv.take<TileState>(tiles + 1);
    cudaStream_t s = ctx->stream;

    memset(h, 0, sizeof(RamDev));
    h->io = *io;
    if (options) h->opt = *options;
    h->n_unsorted = n_unsorted; h->n_sorted = n_sorted; h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(RamDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(sg, 0, (char *)(ts + tiles + 1) - (char *)sg, s));

    const zkc_memory_query *du = unsorted, *dsq = sorted;
    const uint64_t *dup = unsorted_prev_states, *dsp = sorted_prev_states;
    uint64_t *dtrace = trace;
    if (!in_dev) {
        zkc_memory_query *bu = cv.take<zkc_memory_query>(need), *bs = cv.take<zkc_memory_query>(need);
        uint64_t *bup = cv.take<uint64_t>(need * 12), *bsp = cv.take<uint64_t>(need * 12);
        if (need) {
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bu, unsorted, need * sizeof(zkc_memory_query), cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bs, sorted, need * sizeof(zkc_memory_query), cudaMemcpyHostToDevice, s));
            if (!need_chain) {
                ZKC_CUDA(ctx, status, cudaMemcpyAsync(bup, unsorted_prev_states, need * 96, cudaMemcpyHostToDevice, s));
                ZKC_CUDA(ctx, status, cudaMemcpyAsync(bsp, sorted_prev_states, need * 96, cudaMemcpyHostToDevice, s));
            }
        }
        du = bu; dsq = bs; dup = bup; dsp = bsp;
    } else if (need_chain) {
        dup = cv.take<uint64_t>(need * 12); dsp = cv.take<uint64_t>(need * 12);
    }
    if (trace && !trace_dev) dtrace = cv.take<uint64_t>((size_t)ZKC_RAM_NUM_COLS * limit);

    ZKC_LAUNCH(ctx, "ram_prologue", ram_prologue_kernel, 1, 96, 0, d);
    if (need_chain) ZKC_LAUNCH(ctx, "ram_chain", ram_chain_kernel, 1, 64, 0, d, du, dsq, (uint64_t *)dup, (uint64_t *)dsp, need);
    if (tiles) ZKC_LAUNCH(ctx, "ram_rows", ram_rows_kernel, (unsigned)tiles, SCAN_THREADS, 0, d, du, dup, dsq, dsp, dtrace, sg, ts);
    if (tiles && dtrace) ZKC_LAUNCH(ctx, "ram_inverse", ram_inverse_kernel, (unsigned)((limit + 127) / 128), 128, 0, d, dtrace);
    ZKC_LAUNCH(ctx, "ram_finalize", ram_finalize_kernel, 1, 32, 0, d);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(RamDev), cudaMemcpyDeviceToHost, s));
    if (!trace_dev && trace && limit)
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(trace, dtrace, (size_t)ZKC_RAM_NUM_COLS * limit * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    io->hidden_fsm_output = h->io.hidden_fsm_output;
    io->completion_flag = h->io.completion_flag;
    memcpy(commitment, h->commitment, 32);
    *status = h->status;
    return status->code;
}

extern "C" int zkc_memory_queue_simulate(zkc_ctx *ctx, const zkc_memory_query *records, size_t n_per_queue,
                                         size_t n_queues, uint64_t *prev_states, zkc_queue_state12 *final_states,
                                         int on_device) {
    if (!ctx || !final_states || (n_per_queue && n_queues && !records)) return ZKC_ERR_INVALID_ARGUMENT;
    if (!n_queues) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    const size_t n = n_per_queue * n_queues;
    const zkc_memory_query *dr = records;
    uint64_t *dp = prev_states;
    zkc_queue_state12 *df = final_states;
    cudaStream_t s = ctx->stream;
    if (!on_device) {
        size_t bytes = zkc_carver::bytes(n, sizeof(zkc_memory_query)) + zkc_carver::bytes(n * 12, 8) +
                       zkc_carver::bytes(n_queues, sizeof(zkc_queue_state12));
        void *blk = ctx->scratch(bytes);
        if (!blk) return ZKC_ERR_CUDA;
        zkc_carver cv(blk);
        zkc_memory_query *br = cv.take<zkc_memory_query>(n);
        dp = prev_states ? cv.take<uint64_t>(n * 12) : nullptr;
        df = cv.take<zkc_queue_state12>(n_queues);
        if (n) ZKC_CUDA(ctx, st, cudaMemcpyAsync(br, records, n * sizeof(zkc_memory_query), cudaMemcpyHostToDevice, s));
        dr = br;
    }
    ZKC_LAUNCH(ctx, "memory_queue_simulate", memory_queue_simulate_kernel, (unsigned)((n_queues + 31) / 32), 32, 0, dr,
               n_per_queue, n_queues, dp, df);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    if (!on_device) {
        if (prev_states && n) ZKC_CUDA(ctx, st, cudaMemcpyAsync(prev_states, dp, n * 96, cudaMemcpyDeviceToHost, s));
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(final_states, df, n_queues * sizeof(zkc_queue_state12), cudaMemcpyDeviceToHost, s));
        ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    }
    return ZKC_OK;
}

extern "C" int zkc_ram_permutation_check_trace(zkc_ctx *ctx, const zkc_ram_closed_form *io, const uint64_t *trace,
                                               size_t limit, const zkc_ram_options *options, uint32_t gates,
                                               int on_device, uint64_t *violations, zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !violations || (limit && !trace)) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    size_t bytes = zkc_carver::bytes(1, sizeof(RamDev));
    if (!on_device) bytes += zkc_carver::bytes((size_t)ZKC_RAM_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    RamDev *h = (RamDev *)ctx->pinned(sizeof(RamDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    RamDev *d = cv.take<RamDev>(1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(RamDev));
    h->io = *io;
    if (options) h->opt = *options;
    h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(RamDev), cudaMemcpyHostToDevice, s));
    const uint64_t *dt = trace;
    if (!on_device && limit) {
        uint64_t *b = cv.take<uint64_t>((size_t)ZKC_RAM_NUM_COLS * limit);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(b, trace, (size_t)ZKC_RAM_NUM_COLS * limit * 8, cudaMemcpyHostToDevice, s));
        dt = b;
    }
    ZKC_LAUNCH(ctx, "ram_prologue", ram_prologue_kernel, 1, 96, 0, d);
    if (limit) {
        const bool rf = gates == 0 || (gates & ZKC_GATES_ROUND_FUNCTION);
        const bool pairs = RAM_CHECK_PAIRS && !rf && limit % 2 == 0 && ((uintptr_t)dt & 15) == 0;  // row pairs: 128-bit loads
        const size_t threads = pairs ? limit / 2 : limit;
        const unsigned grid = (unsigned)((threads + RAM_CHECK_THREADS - 1) / RAM_CHECK_THREADS);
        if (rf) ZKC_LAUNCH(ctx, "ram_check_rf", (ram_check_kernel<true, 1>), grid, RAM_CHECK_THREADS, 0, d, dt);
        else if (pairs) ZKC_LAUNCH(ctx, "ram_check", (ram_check_kernel<false, 2>), grid, RAM_CHECK_THREADS, 0, d, dt);
        else ZKC_LAUNCH(ctx, "ram_check", (ram_check_kernel<false, 1>), grid, RAM_CHECK_THREADS, 0, d, dt);
    }
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(RamDev), cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    *violations = h->violations;
    status->failed_checks = h->failed_checks;
    if (h->violations) {
        status->code = ZKC_ERR_UNSATISFIED;
        status->first_bad_row = (int64_t)(h->first_bad >> 16);
    }
    return status->code;
}
