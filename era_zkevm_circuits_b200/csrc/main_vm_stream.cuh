// Transport forms of the main_vm call over PCIe (include/zkc_b200.h, "transport forms"): the segmented input stream
// (host encoder + the device kernels that expand a segment blob into the column layout the cycle kernel reads) and the
// PACKED trace (device kernel that narrows the COMPACT-layout trace into typed columns + sparse records).
// Included by main_vm.cu (one translation unit with the cycle kernels).
#pragma once
#include <thread>

namespace zkc {

// ---- PACKED layout: kind / slot of every DENSE-layout column -----------------------------------------------------------------
struct VmPkTable {
    uint8_t kind[ZKC_VM_NUM_COLS];
    uint16_t slot[ZKC_VM_NUM_COLS];
    uint32_t counts[7];
};
__host__ __device__ constexpr int vm_pk_kind_of(int c) {
    if (c >= ZKC_VM_OP_AUX) return ZKC_VM_PK_AUX_RECORD;
    if (c >= ZKC_VM_SPONGE_ENFORCE) return ZKC_VM_PK_SPONGE_RECORD;
    if (c >= ZKC_VM_FORWARD_TAIL_OUT) return ZKC_VM_PK_AUX_RECORD;  // forward tail (4 + length), rollback head (4 + length)
    if ((c >= ZKC_VM_CODE_WORD && c < ZKC_VM_CODE_WORD + 8) || (c >= ZKC_VM_SRC0_FROM_MEMORY && c < ZKC_VM_SRC0_FROM_MEMORY + 9) ||
        (c >= ZKC_VM_DST1 && c < ZKC_VM_DST1 + 9))
        return ZKC_VM_PK_LIMB_RECORD;
    if (c == ZKC_VM_PROPS) return ZKC_VM_PK_U64;
    if (c == ZKC_VM_SUPER_PC || c == ZKC_VM_VARIANT || c == ZKC_VM_IMM0 || c == ZKC_VM_IMM1 || c == ZKC_VM_SRC0_INDEX ||
        c == ZKC_VM_SP_AFTER_SRC0 || c == ZKC_VM_DST0_INDEX || c == ZKC_VM_NEW_SP || c == ZKC_VM_PC_OUT)
        return ZKC_VM_PK_U16;
    if (c <= ZKC_VM_SHOULD_READ_OPCODE || c == ZKC_VM_SUB_PC || c == ZKC_VM_CONDITION_IDX || c == ZKC_VM_CONDITION ||
        (c >= ZKC_VM_OUT_OF_ERGS && c <= ZKC_VM_MASK_INTO_NOP) || (c >= ZKC_VM_SRC0_REG && c <= ZKC_VM_DST1_REG) ||
        c == ZKC_VM_SHOULD_READ_SRC0 || c == ZKC_VM_DST0_PERFORMS_MEMORY_ACCESS || c == ZKC_VM_SRC0_FROM_MEMORY || c == ZKC_VM_SWAP_OPERANDS ||
        c == ZKC_VM_SRC0 || c == ZKC_VM_SRC1 || c == ZKC_VM_DST0 || c == ZKC_VM_DST1 ||
        (c >= ZKC_VM_PERFORM_DST0_MEMORY_WRITE && c <= ZKC_VM_PENDING_EXCEPTION_OUT))
        return ZKC_VM_PK_U8;
    return ZKC_VM_PK_U32;
}
__host__ __device__ constexpr VmPkTable vm_pk_make() {
    VmPkTable t{};
    for (int c = 0; c < ZKC_VM_NUM_COLS; c++) {
        const int k = vm_pk_kind_of(c);
        t.kind[c] = (uint8_t)k;
        if (k == ZKC_VM_PK_LIMB_RECORD) {  // slot = 9 * record kind + index in v[]
            t.slot[c] = (uint16_t)(c >= ZKC_VM_DST1 ? 9 * ZKC_VM_LIMB_DST1 + (c - ZKC_VM_DST1)
                                   : (c >= ZKC_VM_SRC0_FROM_MEMORY ? 9 * ZKC_VM_LIMB_SRC0_FROM_MEMORY + (c - ZKC_VM_SRC0_FROM_MEMORY) : c - ZKC_VM_CODE_WORD));
            t.counts[k]++;
        } else t.slot[c] = (uint16_t)t.counts[k]++;
    }
    return t;
}
constexpr VmPkTable VM_PK = vm_pk_make();
constexpr int VM_PK_N8 = (int)VM_PK.counts[0], VM_PK_N16 = (int)VM_PK.counts[1], VM_PK_N32 = (int)VM_PK.counts[2], VM_PK_N64 = (int)VM_PK.counts[3];
static_assert(VM_PK.counts[ZKC_VM_PK_AUX_RECORD] == 58 && VM_PK.counts[ZKC_VM_PK_SPONGE_RECORD] == 117 && VM_PK.counts[ZKC_VM_PK_LIMB_RECORD] == 26, "record columns");
static_assert(sizeof(zkc_vm_limb_record) == 48, "limb record layout");
static_assert(sizeof(zkc_vm_aux_record) == 8 + 58 * 8, "aux record layout");

struct VmPackOut {
    uint8_t *c8; uint16_t *c16; uint32_t *c32; uint64_t *c64;
    size_t rows;  // column pitch of the four blocks (n_instances * limit)
    zkc_vm_aux_record *aux; unsigned long long *n_aux; unsigned long long aux_cap;
    zkc_vm_limb_record *limb; unsigned long long *n_limb;  // capacity: 3 per row of the chunk (nothing is dropped on the device)
};

template <int C, int END>
struct VmPackCols {
    static __device__ __forceinline__ void run(const uint64_t *__restrict__ t, size_t limit, const VmPackOut &o, size_t g) {
        constexpr int kind = VM_PK.kind[C];
        constexpr size_t slot = VM_PK.slot[C];
        const uint64_t v = __ldg(t + (size_t)C * limit);
        if constexpr (kind == ZKC_VM_PK_U8) o.c8[slot * o.rows + g] = (uint8_t)v;
        else if constexpr (kind == ZKC_VM_PK_U16) o.c16[slot * o.rows + g] = (uint16_t)v;
        else if constexpr (kind == ZKC_VM_PK_U32) o.c32[slot * o.rows + g] = (uint32_t)v;
        else if constexpr (kind == ZKC_VM_PK_U64) o.c64[slot * o.rows + g] = v;
        VmPackCols<C + 1, END>::run(t, limit, o, g);
    }
};
template <int END>
struct VmPackCols<END, END> {
    static __device__ __forceinline__ void run(const uint64_t *__restrict__, size_t, const VmPackOut &, size_t) {}
};

// COMPACT-layout device trace [n_inst][ZKC_VM_COMPACT_COLS][limit] -> typed columns + aux records, rows [r0, r0 + cnt) of every instance
__global__ void __launch_bounds__(256)
vm_pack_kernel(const uint64_t *__restrict__ dense, size_t limit, size_t n_inst, size_t r0, size_t cnt, VmPackOut o) {
    const size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = l < cnt * n_inst;
    const unsigned lane = threadIdx.x & 31;
    bool emit = false;
    size_t g = 0;
    const uint64_t *t = dense;
    if (valid) {
        const size_t inst = l / cnt, row = r0 + (l - inst * cnt);
        g = inst * limit + row;
        t = dense + inst * (size_t)ZKC_VM_COMPACT_COLS * limit + row;
        VmPackCols<0, ZKC_VM_FORWARD_TAIL_OUT>::run(t, limit, o, g);
        uint64_t any = 0, moved = row == 0;
#pragma unroll
        for (int i = 0; i < 10; i++) {
            const uint64_t v = __ldg(t + (size_t)(ZKC_VM_FORWARD_TAIL_OUT + i) * limit);
            if (row) moved |= v ^ __ldg(t + (size_t)(ZKC_VM_FORWARD_TAIL_OUT + i) * limit - 1);
        }
#pragma unroll
        for (int i = 0; i < ZKC_VM_OP_AUX_COLS; i++) any |= __ldg(t + (size_t)(ZKC_VM_COMPACT_OP_AUX + i) * limit);
        emit = (any | moved) != 0;
    }
    const unsigned b = __ballot_sync(0xffffffffu, emit);
    if (b) {
        const int leader = __ffs(b) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(o.n_aux, (unsigned long long)__popc(b));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (emit) {
            const unsigned long long pos = base + __popc(b & ((1u << lane) - 1));
            if (pos < o.aux_cap) {
                zkc_vm_aux_record &r = o.aux[pos];
                r.row = (uint32_t)g; r.reserved = 0;
                for (int i = 0; i < ZKC_VM_OP_AUX_COLS; i++) r.op_aux[i] = __ldg(t + (size_t)(ZKC_VM_COMPACT_OP_AUX + i) * limit);
                for (int i = 0; i < 10; i++) r.queue_ends[i] = __ldg(t + (size_t)(ZKC_VM_FORWARD_TAIL_OUT + i) * limit);
            }
        }
    }
    // ---- limb records: code word (on an opcode fetch / row 0), src0 memory operand, dst1 (when not zero) -------------------
    uint32_t cw[9], sm[9], d1[9];
    bool e_cw = false, e_sm = false, e_d1 = false;
    if (valid) {
        const size_t row = g % limit;
        uint32_t any_sm = 0, any_d1 = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            cw[i] = i < 8 ? (uint32_t)__ldg(t + (size_t)(ZKC_VM_CODE_WORD + i) * limit) : 0u;
            sm[i] = (uint32_t)__ldg(t + (size_t)(ZKC_VM_SRC0_FROM_MEMORY + i) * limit); any_sm |= sm[i];
            d1[i] = (uint32_t)__ldg(t + (size_t)(ZKC_VM_DST1 + i) * limit); any_d1 |= d1[i];
        }
        e_cw = row == 0 || __ldg(t + (size_t)ZKC_VM_SHOULD_READ_OPCODE * limit) != 0;
        e_sm = any_sm != 0; e_d1 = any_d1 != 0;
    }
    const unsigned b0 = __ballot_sync(0xffffffffu, e_cw), b1 = __ballot_sync(0xffffffffu, e_sm), b2 = __ballot_sync(0xffffffffu, e_d1);
    const unsigned total = __popc(b0) + __popc(b1) + __popc(b2);
    if (!total) return;
    unsigned long long lbase = 0;
    if (lane == 0) lbase = atomicAdd(o.n_limb, (unsigned long long)total);
    lbase = __shfl_sync(0xffffffffu, lbase, 0);
    const unsigned below = (1u << lane) - 1;
    auto put = [&](unsigned long long pos, uint32_t kind, const uint32_t (&v)[9]) {
        zkc_vm_limb_record &r = o.limb[pos];
        r.row = (uint32_t)g; r.kind = kind; r.reserved = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) r.v[i] = v[i];
    };
    if (e_cw) put(lbase + __popc(b0 & below), ZKC_VM_LIMB_CODE_WORD, cw);
    if (e_sm) put(lbase + __popc(b0) + __popc(b1 & below), ZKC_VM_LIMB_SRC0_FROM_MEMORY, sm);
    if (e_d1) put(lbase + __popc(b0) + __popc(b1) + __popc(b2 & below), ZKC_VM_LIMB_DST1, d1);
}

// ---- expansion of a segment blob into the columns ---------------------------------------------------------------------------------
// dense words: cols[word[d] * stride + base + i] = vals[d * n + i]
__global__ void __launch_bounds__(256)
vm_expand_dense_kernel(const uint16_t *__restrict__ words, const uint32_t *__restrict__ vals, size_t n, uint32_t *__restrict__ cols, size_t stride, size_t base) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = blockIdx.y;
    cols[(size_t)words[d] * stride + base + i] = __ldg(vals + (size_t)d * n + i);
}
// sparse state words: a warp takes 32 consecutive change records of one word; together they cover one contiguous span of the
// word's column, which the warp fills with coalesced stores (each element finds its record by a 5-step search over the
// lanes' run starts)
__global__ void __launch_bounds__(128)
vm_expand_runs_kernel(const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ index, const uint32_t *__restrict__ value, uint32_t n_total,
                      uint32_t *__restrict__ cols, size_t stride, size_t base) {
    const int w = blockIdx.y, lane = threadIdx.x & 31;
    const uint32_t lo = __ldg(offsets + w), hi = __ldg(offsets + w + 1);
    const uint32_t j0 = lo + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (j0 >= hi) return;
    const uint32_t j = j0 + lane;
    const uint32_t start = j < hi ? __ldg(index + j) : n_total, val = j < hi ? __ldg(value + j) : 0u;
    const uint32_t S = __shfl_sync(0xffffffffu, start, 0);
    const uint32_t E = j0 + 32 < hi ? __ldg(index + j0 + 32) : n_total;
    uint32_t *dst = cols + (size_t)w * stride + base;
    for (uint32_t i0 = S; i0 < E; i0 += 32) {
        const uint32_t i = i0 + lane;
        int r = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const uint32_t s = __shfl_sync(0xffffffffu, start, (r + step) & 31);
            if (r + step < 32 && s <= i) r += step;
        }
        const uint32_t v = __shfl_sync(0xffffffffu, val, r);
        if (i < E) dst[i] = v;
    }
}
// sparse witness words: the segment's rows are zeroed first; one thread per (cycle, value) record
__global__ void __launch_bounds__(256)
vm_expand_scatter_kernel(const uint32_t *__restrict__ offsets, int n_words, const uint32_t *__restrict__ index, const uint32_t *__restrict__ value,
                         uint32_t n_records, uint32_t *__restrict__ cols, size_t stride, size_t base) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_records) return;
    int a = 0, b = n_words;  // last word whose offset is <= j
    while (b - a > 1) {
        const int m = (a + b) >> 1;
        if (__ldg(offsets + m) <= j) a = m; else b = m;
    }
    cols[(size_t)a * stride + base + __ldg(index + j)] = __ldg(value + j);
}

// ---- host encoder: records of one instance -> segmented stream ---------------------------------------------------------------------
struct VmStreamOwner {  // what zkc_vm_encode_input_stream allocates: the public struct first
    zkc_vm_input_stream pub;
    std::vector<zkc_vm_input_segment> segs;
    void *arena = nullptr;
    bool pinned = false;
};

static inline size_t vm_align16(size_t x) { return (x + 15) & ~(size_t)15; }

struct VmSegPlan {
    size_t first, n;
    std::vector<uint32_t> st_count, wt_count;  // change records / non-zero values per word
    std::vector<uint8_t> st_dense, wt_dense;
    zkc_vm_segment_header h;
    size_t arena_off;
};

static void vm_plan_segment(const zkc_vm_state *snaps, const zkc_vm_cycle_witness *wit, VmSegPlan &p) {
    p.st_count.assign(ZKC_VM_STATE_WORDS, 1);
    p.wt_count.assign(ZKC_VM_WITNESS_WORDS, 0);
    const uint32_t *s = reinterpret_cast<const uint32_t *>(snaps + p.first);
    for (size_t i = 1; i <= p.n; i++) {
        const uint32_t *a = s + (i - 1) * ZKC_VM_STATE_WORDS, *b = s + i * ZKC_VM_STATE_WORDS;
        for (int w = 0; w < ZKC_VM_STATE_WORDS; w++) p.st_count[w] += a[w] != b[w];
    }
    const uint32_t *q = reinterpret_cast<const uint32_t *>(wit + p.first);
    for (size_t i = 0; i < p.n; i++)
        for (int w = 0; w < ZKC_VM_WITNESS_WORDS; w++) p.wt_count[w] += q[i * ZKC_VM_WITNESS_WORDS + w] != 0;
    p.st_dense.assign(ZKC_VM_STATE_WORDS, 0);
    p.wt_dense.assign(ZKC_VM_WITNESS_WORDS, 0);
    zkc_vm_segment_header &h = p.h;
    memset(&h, 0, sizeof h);
    h.magic = ZKC_VM_SEGMENT_MAGIC; h.first_cycle = (uint32_t)p.first; h.n_cycles = (uint32_t)p.n;
    for (int w = 0; w < ZKC_VM_STATE_WORDS; w++) {
        p.st_dense[w] = 4 * (p.n + 1) <= 8 * (size_t)p.st_count[w];
        if (p.st_dense[w]) h.n_dense_state++; else h.n_sparse_state += p.st_count[w];
    }
    for (int w = 0; w < ZKC_VM_WITNESS_WORDS; w++) {
        p.wt_dense[w] = p.n && 4 * p.n <= 8 * (size_t)p.wt_count[w];
        if (p.wt_dense[w]) h.n_dense_witness++; else h.n_sparse_witness += p.wt_count[w];
    }
    size_t off = vm_align16(sizeof h);
    h.off_dense_state_word = (uint32_t)off; off = vm_align16(off + 2 * (size_t)h.n_dense_state);
    h.off_dense_witness_word = (uint32_t)off; off = vm_align16(off + 2 * (size_t)h.n_dense_witness);
    h.off_dense_state = (uint32_t)off; off = vm_align16(off + 4 * (size_t)h.n_dense_state * (p.n + 1));
    h.off_dense_witness = (uint32_t)off; off = vm_align16(off + 4 * (size_t)h.n_dense_witness * p.n);
    h.off_sparse_state_offsets = (uint32_t)off; off = vm_align16(off + 4 * (ZKC_VM_STATE_WORDS + 1));
    h.off_sparse_state_index = (uint32_t)off; off = vm_align16(off + 4 * (size_t)h.n_sparse_state);
    h.off_sparse_state_value = (uint32_t)off; off = vm_align16(off + 4 * (size_t)h.n_sparse_state);
    h.off_sparse_witness_offsets = (uint32_t)off; off = vm_align16(off + 4 * (ZKC_VM_WITNESS_WORDS + 1));
    h.off_sparse_witness_index = (uint32_t)off; off = vm_align16(off + 4 * (size_t)h.n_sparse_witness);
    h.off_sparse_witness_value = (uint32_t)off; off = vm_align16(off + 4 * (size_t)h.n_sparse_witness);
    h.blob_bytes = (uint32_t)off;
}

static void vm_fill_segment(const zkc_vm_state *snaps, const zkc_vm_cycle_witness *wit, const VmSegPlan &p, char *blob) {
    const zkc_vm_segment_header &h = p.h;
    memset(blob, 0, h.blob_bytes);
    memcpy(blob, &h, sizeof h);
    uint16_t *dsw = (uint16_t *)(blob + h.off_dense_state_word), *dww = (uint16_t *)(blob + h.off_dense_witness_word);
    uint32_t *ds = (uint32_t *)(blob + h.off_dense_state), *dw = (uint32_t *)(blob + h.off_dense_witness);
    uint32_t *so = (uint32_t *)(blob + h.off_sparse_state_offsets), *si = (uint32_t *)(blob + h.off_sparse_state_index), *sv = (uint32_t *)(blob + h.off_sparse_state_value);
    uint32_t *wo = (uint32_t *)(blob + h.off_sparse_witness_offsets), *wi = (uint32_t *)(blob + h.off_sparse_witness_index), *wv = (uint32_t *)(blob + h.off_sparse_witness_value);
    // where each word's data starts
    std::vector<uint32_t> st_row(ZKC_VM_STATE_WORDS, 0), st_cur(ZKC_VM_STATE_WORDS, 0), wt_row(ZKC_VM_WITNESS_WORDS, 0), wt_cur(ZKC_VM_WITNESS_WORDS, 0);
    uint32_t nd = 0, ns = 0;
    for (int w = 0; w < ZKC_VM_STATE_WORDS; w++) {
        so[w] = ns;
        if (p.st_dense[w]) { dsw[nd] = (uint16_t)w; st_row[w] = nd++; }
        else { st_cur[w] = ns; ns += p.st_count[w]; }
    }
    so[ZKC_VM_STATE_WORDS] = ns;
    nd = 0; ns = 0;
    for (int w = 0; w < ZKC_VM_WITNESS_WORDS; w++) {
        wo[w] = ns;
        if (p.wt_dense[w]) { dww[nd] = (uint16_t)w; wt_row[w] = nd++; }
        else { wt_cur[w] = ns; ns += p.wt_count[w]; }
    }
    wo[ZKC_VM_WITNESS_WORDS] = ns;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(snaps + p.first);
    const size_t n1 = p.n + 1;
    for (size_t i = 0; i <= p.n; i++) {
        const uint32_t *b = s + i * ZKC_VM_STATE_WORDS, *a = i ? b - ZKC_VM_STATE_WORDS : b;
        for (int w = 0; w < ZKC_VM_STATE_WORDS; w++) {
            if (p.st_dense[w]) ds[(size_t)st_row[w] * n1 + i] = b[w];
            else if (i == 0 || a[w] != b[w]) { si[st_cur[w]] = (uint32_t)i; sv[st_cur[w]] = b[w]; st_cur[w]++; }
        }
    }
    const uint32_t *q = reinterpret_cast<const uint32_t *>(wit + p.first);
    for (size_t i = 0; i < p.n; i++)
        for (int w = 0; w < ZKC_VM_WITNESS_WORDS; w++) {
            const uint32_t v = q[i * ZKC_VM_WITNESS_WORDS + w];
            if (p.wt_dense[w]) dw[(size_t)wt_row[w] * p.n + i] = v;
            else if (v) { wi[wt_cur[w]] = (uint32_t)i; wv[wt_cur[w]] = v; wt_cur[w]++; }
        }
}

}  // namespace zkc

extern "C" int zkc_vm_encode_input_stream(const zkc_vm_state *snapshots, const zkc_vm_cycle_witness *witness, size_t limit, size_t segment_cycles,
                                          zkc_vm_input_stream **out, uint64_t *bytes_out) {
    using namespace zkc;
    if (!snapshots || (limit && !witness) || !out || limit > 0x0FFFFFFFull) return ZKC_ERR_INVALID_ARGUMENT;
    const size_t seg = segment_cycles ? segment_cycles : (size_t)1 << 16;
    const size_t n_seg = std::max<size_t>(1, (limit + seg - 1) / seg);
    std::vector<VmSegPlan> plans(n_seg);
    for (size_t k = 0; k < n_seg; k++) { plans[k].first = k * seg; plans[k].n = std::min(seg, limit - k * seg); }
    const unsigned n_thr = std::max(1u, std::min<unsigned>((unsigned)n_seg, std::thread::hardware_concurrency()));
    auto parallel = [&](auto fn) {
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < n_thr; t++)
            ts.emplace_back([&, t]() { for (size_t k = t; k < n_seg; k += n_thr) fn(k); });
        for (auto &t : ts) t.join();
    };
    parallel([&](size_t k) { vm_plan_segment(snapshots, witness, plans[k]); });
    size_t total = 0;
    for (auto &p : plans) { p.arena_off = total; total += (p.h.blob_bytes + 255) & ~(size_t)255; }
    VmStreamOwner *o = new VmStreamOwner();
    if (cudaHostAlloc(&o->arena, std::max<size_t>(total, 256), cudaHostAllocDefault) == cudaSuccess) o->pinned = true;
    else {  // no device (or no pinned memory left): pageable memory, the copies are then synchronous
        (void)cudaGetLastError();
        o->arena = aligned_alloc(256, std::max<size_t>(total, 256));
        if (!o->arena) { delete o; return ZKC_ERR_CUDA; }
    }
    parallel([&](size_t k) { vm_fill_segment(snapshots, witness, plans[k], (char *)o->arena + plans[k].arena_off); });
    o->segs.resize(n_seg);
    uint64_t bytes = 0;
    for (size_t k = 0; k < n_seg; k++) {
        o->segs[k].blob = (char *)o->arena + plans[k].arena_off;
        o->segs[k].blob_bytes = plans[k].h.blob_bytes;
        bytes += plans[k].h.blob_bytes;
    }
    o->pub.limit = limit; o->pub.segment_cycles = (uint32_t)seg; o->pub.n_segments = (uint32_t)n_seg; o->pub.segments = o->segs.data();
    *out = &o->pub;
    if (bytes_out) *bytes_out = bytes;
    return ZKC_OK;
}

extern "C" void zkc_vm_input_stream_free(zkc_vm_input_stream *stream) {
    if (!stream) return;
    zkc::VmStreamOwner *o = reinterpret_cast<zkc::VmStreamOwner *>(stream);
    if (o->arena) { if (o->pinned) cudaFreeHost(o->arena); else free(o->arena); }
    delete o;
}

extern "C" void zkc_vm_packed_layout(uint8_t kind[ZKC_VM_NUM_COLS], uint16_t slot[ZKC_VM_NUM_COLS], uint32_t counts[7]) {
    for (int c = 0; c < ZKC_VM_NUM_COLS; c++) { kind[c] = zkc::VM_PK.kind[c]; slot[c] = zkc::VM_PK.slot[c]; }
    for (int k = 0; k < 7; k++) counts[k] = zkc::VM_PK.counts[k];
}
