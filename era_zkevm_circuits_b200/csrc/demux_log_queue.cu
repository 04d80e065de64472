// Log-queue demultiplexer on sm_100a: demultiplex_storage_logs_enty_point
// (/root/reference/src/demux_log_queue/mod.rs:38-217) and its loop demultiplex_storage_logs_inner (:234-399) with
// push_with_optimize (:401-444), one thread per loop iteration.  It is the step between main_vm's log queue and the
// storage / events / L1-message sorters and the keccak256 / sha256 / ecrecover precompile circuits.
// Row-parallel recovery of the loop's sequential state:
//   - the popped queue's head: previous-tail column of the raw queue witness, verified link by link;
//   - the six output queues' tails: hash chains over the rows routed to each queue.  A scan over six counters gives every
//     row its position in its queue; round 2 of the push (the only one that consumes the previous tail) is verified
//     against host-supplied tails (`output_tails`, what the out-of-circuit demultiplexer produced and the downstream
//     circuits consume) or rebuilt by six sequential chains running side by side;
//   - rounds 0 and 1 of a push absorb the same encoding from the same empty sponge as the pop's first two rounds: they are
//     computed once per row (4 permutations per row instead of the circuit's 6; identical values).
#include "ctx.cuh"
#include "log_query.cuh"
#include "scan.cuh"

namespace zkc {

constexpr int DMX_Q = ZKC_DEMUX_NUM_QUEUES;

struct DmxDev {
    zkc_demux_closed_form io;
    zkc_demux_options opt;  // constants resolved by the host wrapper
    uint64_t n_records, limit;
    uint64_t tails_base[DMX_Q], n_tails[DMX_Q];  // per output queue: first tail (in units of 4 elements) and count
    // prologue
    uint32_t start, prologue_checks;
    zkc_queue_state4 iq0, oq0[DMX_Q];
    uint64_t commit_obs_in[4], commit_fsm_in[4];
    // rows
    uint32_t counts_final[DMX_Q];
    uint64_t head_final[4];
    // status
    unsigned long long first_bad;
    uint32_t failed_checks, hint_bad;
    // finalize
    uint64_t commitment[4];
    zkc_status status;
};

struct DmxVal {
    uint32_t c[DMX_Q];
};
struct DmxValOp {
    static __device__ __forceinline__ DmxVal identity() { return DmxVal{{0, 0, 0, 0, 0, 0}}; }
    static __device__ __forceinline__ DmxVal combine(const DmxVal &a, const DmxVal &b) {
        DmxVal r;
#pragma unroll
        for (int i = 0; i < DMX_Q; i++) r.c[i] = a.c[i] + b.c[i];
        return r;
    }
};
using DmxTile = TileStateT<DmxVal>;
using DmxShared = ScanSharedT<DmxVal>;

// CSVarLengthEncodable order of LogDemuxerFSMInputOutput, input.rs:24-32
static __device__ int dmx_encode_fsm(const zkc_demux_fsm &f, uint64_t *dst) {
    int n = put_queue_state4(dst, f.initial_log_queue_state);
    for (int q = 0; q < DMX_Q; q++) n += put_queue_state4(dst + n, f.output_queue_states[q]);
    return n;  // 63
}

// warp 0: start selection; warps 1 / 2: commitments to the observable input / FSM input, 12 lanes per permutation
__global__ void dmx_prologue_kernel(DmxDev *d) {
    __shared__ uint64_t buf[2][64];
    const int warp = threadIdx.x >> 5, i = threadIdx.x & 31;
    if (i >= 16) return;
    const unsigned gm = 0xFFFFu;
    const zkc_demux_closed_form &io = d->io;
    if (warp == 0) {
        if (i != 0) return;
        const bool start = io.start_flag != 0;
        d->start = start;
        d->iq0 = start ? io.initial_log_queue_state : io.hidden_fsm_input.initial_log_queue_state;
        zkc_queue_state4 empty;
        memset(&empty, 0, sizeof empty);
        for (int q = 0; q < DMX_Q; q++) d->oq0[q] = start ? empty : io.hidden_fsm_input.output_queue_states[q];  // :83-106
        uint32_t checks = 0;
        for (int k = 0; k < 4; k++)
            if (io.initial_log_queue_state.head[k]) checks |= ZKC_DMX_CHK_TRIVIAL_HEAD;
        d->prologue_checks = checks;
    } else {
        uint64_t *b = buf[warp - 1];
        int n = 0;
        if (i == 0) n = warp == 1 ? put_queue_state4(b, io.initial_log_queue_state) : dmx_encode_fsm(io.hidden_fsm_input, b);
        __syncwarp(gm);
        n = __shfl_sync(gm, n, 0, 16);
        const uint64_t c = commit_encoding_coop(gm, b, n, i);
        if (i < 4) (warp == 1 ? d->commit_obs_in : d->commit_fsm_in)[i] = c;
    }
}

__device__ __forceinline__ void dmx_report(DmxDev *d, size_t row, uint32_t checks) {
    if (!checks) return;
    atomicOr(&d->failed_checks, checks);
    atomicMin(&d->first_bad, ((unsigned long long)row << 16) | checks);
}

// ---- pass A: pop, classification, position of the row in its output queue, rounds 0-1 of the push ---------------------
__global__ void __launch_bounds__(SCAN_THREADS)
dmx_rows_kernel(DmxDev *d, const zkc_log_query *__restrict__ recs, const uint64_t *__restrict__ prev,
                uint64_t *__restrict__ trace, uint64_t *__restrict__ r2in, uint32_t *__restrict__ meta,
                uint32_t *__restrict__ counts, uint32_t *__restrict__ lists, ScanGlobal *sg, DmxTile *tiles) {
    __shared__ DmxShared sh;
    const unsigned int tile = scan_take_ticket(sg, sh);
    const size_t limit = d->limit;
    const size_t row = (size_t)tile * SCAN_THREADS + threadIdx.x;
    const bool in_range = row < limit;
    const uint32_t len0 = d->iq0.length;
    const bool queue_is_empty = row >= len0;
    const bool execute = in_range && !queue_is_empty;
    const size_t active_rows = limit < len0 ? limit : len0;
    uint32_t checks = 0;
#define TR(col) trace[(size_t)(col) * limit + row]
    const bool wr = in_range && trace != nullptr;
    zkc_log_query it = lq_zero();
    if (execute && row < d->n_records) it = lq_load(recs + row);
    uint64_t e[20], s[12];
    lq_encode(it, e);
    // rounds 0 and 1: shared by pop_front and by the push of the same encoding
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = i < 8 ? e[i] : 0;
    poseidon2_permute(s);
    if (wr) {
#pragma unroll
        for (int i = 0; i < 12; i++) TR(ZKC_DMX_PUSH_ROUND0 + i) = s[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = e[8 + i];
    poseidon2_permute(s);
    if (wr) {
#pragma unroll
        for (int i = 0; i < 12; i++) TR(ZKC_DMX_PUSH_ROUND1 + i) = s[i];
    }
    if (in_range) {
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(r2in + 8 * row);
        o[0] = make_ulonglong2(e[16], e[17]); o[1] = make_ulonglong2(e[18], e[19]);
        o[2] = make_ulonglong2(s[8], s[9]); o[3] = make_ulonglong2(s[10], s[11]);
    }
    uint64_t head[4];
    if (execute) {
        uint64_t chain[4];
        bool hint_ok = true;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            chain[i] = __ldg(prev + 4 * row + i);
            if (row == 0 && chain[i] != d->iq0.head[i]) hint_ok = false;
        }
        lq_absorb_tail(e, chain, s);
#pragma unroll
        for (int i = 0; i < 4; i++) head[i] = s[i];
        if (row + 1 < active_rows) {
#pragma unroll
            for (int i = 0; i < 4; i++) hint_ok &= __ldg(prev + 4 * (row + 1) + i) == head[i];
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) d->head_final[i] = head[i];
        }
        if (!hint_ok) { checks |= ZKC_DMX_CHK_QUEUE_HINT; d->hint_bad = 1; }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) head[i] = len0 == 0 ? d->iq0.head[i] : d->iq0.tail[i];
    }

    // ---- :285-360 classification ---------------------------------------------------------------------------
    const uint32_t aux = ZKC_LQ_AUX(it.flags);
    bool is_aux[4], is_addr[3];
#pragma unroll
    for (int i = 0; i < 4; i++) is_aux[i] = aux == d->opt.aux_bytes[i];
    const bool small = (it.address[1] | it.address[2] | it.address[3] | it.address[4]) == 0;
#pragma unroll
    for (int i = 0; i < 3; i++) is_addr[i] = small && it.address[0] == d->opt.precompile_addresses[i];
    const bool is_rollup_shard = ZKC_LQ_SHARD(it.flags) == 0;
    const bool execute_porter_storage = is_aux[0] && !is_rollup_shard && execute;
    if (execute_porter_storage) checks |= ZKC_DMX_CHK_PORTER_STORAGE;
    const bool bit[DMX_Q] = {is_aux[0] && is_rollup_shard && execute, is_aux[1] && execute, is_aux[2] && execute,
                             is_aux[3] && is_addr[0] && execute, is_aux[3] && is_addr[1] && execute,
                             is_aux[3] && is_addr[2] && execute};
    int sel = 0;
    bool any = bit[0];
#pragma unroll
    for (int q = 1; q < DMX_Q; q++) if (bit[q]) { sel = q; any = true; }
    const bool is_bitmask = (int)is_aux[0] + (int)is_aux[1] + (int)is_aux[2] + (int)is_aux[3] == 1;
    if (execute && !is_bitmask) checks |= ZKC_DMX_CHK_BITMASK;

    DmxVal v;
#pragma unroll
    for (int q = 0; q < DMX_Q; q++) v.c[q] = any && sel == q;  // the host wrapper guarantees at most one bit per row
    DmxVal incl;
    const DmxVal excl = scan_tile_generic<DmxVal, DmxValOp>(v, tile, DmxValOp::identity(), tiles, sh, incl);
    if (in_range) {
        meta[row] = (uint32_t)sel | ((uint32_t)any << 3);
#pragma unroll
        for (int q = 0; q < DMX_Q; q++) counts[DMX_Q * row + q] = incl.c[q];
        if (any) lists[(size_t)sel * limit + excl.c[sel]] = (uint32_t)row;
    }
    if (wr) {
        TR(ZKC_DMX_QUEUE_IS_EMPTY) = queue_is_empty; TR(ZKC_DMX_EXECUTE) = execute;
#pragma unroll
        for (int i = 0; i < 36; i++) TR(ZKC_DMX_ITEM + i) = lq_flat(it, i);
#pragma unroll
        for (int i = 0; i < 20; i++) TR(ZKC_DMX_ENC + i) = e[i];
#pragma unroll
        for (int i = 0; i < 4; i++) TR(ZKC_DMX_HEAD + i) = head[i];
        const size_t popped_now = row + 1 < active_rows ? row + 1 : active_rows;
        TR(ZKC_DMX_LEN) = len0 - (uint32_t)popped_now;
#pragma unroll
        for (int i = 0; i < 4; i++) TR(ZKC_DMX_IS_AUX + i) = is_aux[i];
#pragma unroll
        for (int i = 0; i < 3; i++) TR(ZKC_DMX_IS_ADDRESS + i) = is_addr[i];
        TR(ZKC_DMX_IS_ROLLUP_SHARD) = is_rollup_shard; TR(ZKC_DMX_EXECUTE_PORTER_STORAGE) = execute_porter_storage;
#pragma unroll
        for (int q = 0; q < DMX_Q; q++) {
            TR(ZKC_DMX_BITMASK + q) = bit[q];
            TR(ZKC_DMX_QUEUE_LENS + q) = d->oq0[q].length + incl.c[q];
        }
        TR(ZKC_DMX_IS_BITMASK) = is_bitmask;
        TR(ZKC_DMX_EXEC_LEN) = d->oq0[sel].length + excl.c[sel];
    }
    if (in_range && row == limit - 1) {
#pragma unroll
        for (int q = 0; q < DMX_Q; q++) d->counts_final[q] = incl.c[q];
    }
    if (in_range) dmx_report(d, row, checks);
#undef TR
}

// the six chains side by side when the caller supplies no tails: warp q, 1 permutation per push of queue q, each on 12
// cooperating lanes (poseidon2_permute_coop)
__global__ void dmx_chain_kernel(const DmxDev *d, const uint64_t *__restrict__ r2in, const uint32_t *__restrict__ lists,
                                 uint64_t *__restrict__ tails) {
    const int q = threadIdx.x >> 5, i = threadIdx.x & 31;
    if (i >= 16 || q >= DMX_Q) return;
    const unsigned gm = 0xFFFFu;
    uint64_t tail = i < 4 ? d->oq0[q].tail[i] : 0ull;  // lanes 0..3 hold the running tail
    const size_t n = d->counts_final[q], limit = d->limit;
    uint64_t *out = tails + 4 * d->tails_base[q];
    for (size_t k = 0; k < n; k++) {
        const size_t row = lists[(size_t)q * limit + k];
        const uint64_t from_tail = __shfl_sync(gm, tail, (i - 4) & 3, 16);
        uint64_t x = 0;
        if (i < 4) x = r2in[8 * row + i];
        else if (i < 8) x = from_tail;
        else if (i < 12) x = r2in[8 * row + 4 + (i - 8)];
        x = poseidon2_permute_coop(gm, x, i);
        if (i < 4) { tail = x; out[4 * k + i] = x; }
    }
}

// ---- pass B: round 2 of the push on the selected queue's state, the six queues after the row ---------------------------
__global__ void __launch_bounds__(256)
dmx_push_kernel(DmxDev *d, const uint64_t *__restrict__ r2in, const uint32_t *__restrict__ meta,
                const uint32_t *__restrict__ counts, const uint64_t *__restrict__ tails, uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const uint32_t m = meta[row];
    const int sel = m & 7;
    const bool any = m >> 3;
    if (!any && !trace) return;
    uint32_t c[DMX_Q];
#pragma unroll
    for (int q = 0; q < DMX_Q; q++) c[q] = counts[DMX_Q * row + q];
    uint32_t csel = 0;
#pragma unroll
    for (int q = 0; q < DMX_Q; q++) if (q == sel) csel = c[q];
    const size_t kb = csel - (uint32_t)any;  // pushes of the selected queue before this row
    const size_t base = d->tails_base[sel], n_t = d->n_tails[sel];
    uint64_t before[4], s[12];
    bool ok = true;
    if (kb == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) before[i] = d->oq0[sel].tail[i];
    } else if (kb - 1 < n_t) {
#pragma unroll
        for (int i = 0; i < 4; i++) before[i] = __ldg(tails + 4 * (base + kb - 1) + i);
    } else {
        ok = false;
#pragma unroll
        for (int i = 0; i < 4; i++) before[i] = 0;
    }
    const ulonglong2 *in = reinterpret_cast<const ulonglong2 *>(r2in + 8 * row);
    const ulonglong2 a = in[0], b = in[1], cc = in[2], e = in[3];
    s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y;
#pragma unroll
    for (int i = 0; i < 4; i++) s[4 + i] = before[i];
    s[8] = cc.x; s[9] = cc.y; s[10] = e.x; s[11] = e.y;
    poseidon2_permute(s);
    if (any) {
        if (kb < n_t) {
#pragma unroll
            for (int i = 0; i < 4; i++) ok &= __ldg(tails + 4 * (base + kb) + i) == s[i];
        } else ok = false;
    }
    if (trace) {
#pragma unroll
        for (int i = 0; i < 4; i++) trace[(size_t)(ZKC_DMX_EXEC_TAIL + i) * limit + row] = before[i];
#pragma unroll
        for (int i = 0; i < 12; i++) trace[(size_t)(ZKC_DMX_PUSH_ROUND2 + i) * limit + row] = s[i];
#pragma unroll
        for (int q = 0; q < DMX_Q; q++) {
            const size_t bq = d->tails_base[q], nq = d->n_tails[q];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint64_t t;
                if (any && q == sel) t = s[i];
                else if (c[q] == 0) t = d->oq0[q].tail[i];
                else t = c[q] - 1 < nq ? __ldg(tails + 4 * (bq + c[q] - 1) + i) : 0ull;
                trace[(size_t)(ZKC_DMX_QUEUE_TAILS + 4 * q + i) * limit + row] = t;
            }
        }
    }
    if (!ok) {
        d->hint_bad = 1;
        atomicOr(&d->failed_checks, ZKC_DMX_CHK_QUEUE_HINT);
        atomicMin(&d->first_bad, ((unsigned long long)row << 16) | ZKC_DMX_CHK_QUEUE_HINT);
    }
}

// ---- finalize: FSM output, consistency, commitment -------------------------------------------------------------------------
__global__ void dmx_finalize_kernel(DmxDev *d, const uint64_t *__restrict__ tails) {
    __shared__ zkc_demux_fsm out;
    __shared__ uint64_t e_out[64], o_out[56], compact[24];
    __shared__ uint32_t sh_completed;
    const int lane = threadIdx.x & 31, i = lane & 15;
    const unsigned gm = lane < 16 ? 0xFFFFu : 0xFFFF0000u;
    zkc_demux_closed_form &io = d->io;
    if (lane == 0) {
        const size_t limit = d->limit;
        const uint32_t len0 = d->iq0.length;
        const size_t popped = limit < len0 ? limit : len0;
        memset(&out, 0, sizeof out);
        out.initial_log_queue_state = d->iq0;
        if (popped > 0)
            for (int k = 0; k < 4; k++) out.initial_log_queue_state.head[k] = d->head_final[k];
        out.initial_log_queue_state.length = len0 - (uint32_t)popped;
        bool hint_bad = d->hint_bad;
        for (int q = 0; q < DMX_Q; q++) {
            zkc_queue_state4 st = d->oq0[q];
            const uint32_t n = limit ? d->counts_final[q] : 0;
            if (n) {
                if (n - 1 < d->n_tails[q]) for (int k = 0; k < 4; k++) st.tail[k] = tails[4 * (d->tails_base[q] + n - 1) + k];
                else hint_bad = true;
            }
            st.length += n;
            out.output_queue_states[q] = st;
        }
        uint32_t checks = d->failed_checks | d->prologue_checks;
        const zkc_queue_state4 &iq = out.initial_log_queue_state;
        const bool completed = iq.length == 0;
        if (completed)
            for (int k = 0; k < 4; k++)
                if (iq.head[k] != iq.tail[k]) checks |= ZKC_DMX_CHK_QUEUE_CONSISTENCY;  // :395
        zkc_queue_state4 empty;
        memset(&empty, 0, sizeof empty);
        const int n_out = dmx_encode_fsm(out, e_out);
        int n_obs = 0;
        for (int q = 0; q < DMX_Q; q++) n_obs += put_queue_state4(o_out + n_obs, completed ? out.output_queue_states[q] : empty);
        zkc_status st;
        st.code = ZKC_OK; st.cuda_error = 0; st.first_bad_row = -1; st.failed_checks = checks; st.reserved = 0;
        if (d->first_bad != ~0ull) st.first_bad_row = (int64_t)(d->first_bad >> 16);
        if (checks) st.code = ZKC_ERR_UNSATISFIED;
        if (hint_bad) { st.code = ZKC_ERR_QUEUE_WITNESS_INCONSISTENT; st.failed_checks |= ZKC_DMX_CHK_QUEUE_HINT; }
        if (d->opt.compare_expected) {
            bool same = (io.completion_flag != 0) == completed;
            uint64_t e_exp[63], o_exp[54];
            dmx_encode_fsm(io.hidden_fsm_output, e_exp);
            int n = 0;
            for (int q = 0; q < DMX_Q; q++) n += put_queue_state4(o_exp + n, io.output_queue_states[q]);
            for (int k = 0; k < n_out; k++) same &= e_out[k] == e_exp[k];
            for (int k = 0; k < n_obs; k++) same &= o_out[k] == o_exp[k];
            if (!same && st.code == ZKC_OK) st.code = ZKC_ERR_FSM_OUTPUT_MISMATCH;
        }
        io.hidden_fsm_output = out;
        for (int q = 0; q < DMX_Q; q++) io.output_queue_states[q] = completed ? out.output_queue_states[q] : empty;
        io.completion_flag = completed;
        d->status = st;
        sh_completed = completed;
    }
    __syncwarp();
    const uint64_t c = commit_encoding_coop(gm, lane < 16 ? e_out : o_out, lane < 16 ? 63 : 54, i);
    const bool completed = sh_completed;
    if (lane < 4) compact[14 + lane] = completed ? 0 : c;
    if (lane >= 16 && lane < 20) compact[6 + lane - 16] = completed ? c : 0;
    if (lane == 0) {
        compact[0] = d->start; compact[1] = completed;
        for (int k = 0; k < 4; k++) {
            compact[2 + k] = d->commit_obs_in[k];
            compact[10 + k] = d->start ? 0 : d->commit_fsm_in[k];
        }
    }
    __syncwarp();
    if (lane < 16) {
        const uint64_t f = commit_encoding_coop(gm, compact, 18, i);
        if (i < 4) d->commitment[i] = f;
    }
}


// ---- constraint evaluation of a finished trace ------------------------------------------------------------------------------
// One thread per row re-evaluates every relation the loop body of demultiplex_storage_logs_inner (mod.rs:268-393) and
// push_with_optimize (:401-447) place that is local to a row or to a row and its predecessor: booleans / ranges of the popped item,
// LogQuery::encode, queue-length / head bookkeeping, the aux-byte / address / shard classification, the six execute bits and the
// bitmask check, the selected output queue's state before the push, every output queue's tail / length after it.  With
// ZKC_GATES_ROUND_FUNCTION also the permutations: rounds 0-1 (shared by the pop and the push), the pop's round 2, the push's round 2.
template <bool ROUND_FUNCTION>
__global__ void __launch_bounds__(128)
dmx_check_kernel(DmxDev *d, unsigned long long *violations, const uint64_t *__restrict__ trace) {
    const size_t limit = d->limit;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= limit) return;
    const bool first = row == 0;
#define TR(col) __ldg(trace + (size_t)(col) * limit + row)
#define TP(col) __ldg(trace + (size_t)(col) * limit + row - 1)
    uint32_t bad = 0;
    const uint64_t is_empty = TR(ZKC_DMX_QUEUE_IS_EMPTY), execute = TR(ZKC_DMX_EXECUTE);
    if ((is_empty | execute) > 1 || execute != 1 - is_empty) bad |= ZKC_DMXV_BOOLEAN;
    uint64_t f[36], limbs = 0;
#pragma unroll
    for (int i = 0; i < 36; i++) f[i] = TR(ZKC_DMX_ITEM + i);
#pragma unroll
    for (int i = 0; i < 29; i++) limbs |= f[i];
    if ((limbs | f[34] | f[35]) >> 32 || (f[29] | f[33]) >> 8 || (f[30] | f[31] | f[32]) > 1) bad |= ZKC_DMXV_BOOLEAN;
    zkc_log_query q = lq_zero();
#pragma unroll
    for (int i = 0; i < 5; i++) q.address[i] = (uint32_t)f[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { q.key[i] = (uint32_t)f[5 + i]; q.read_value[i] = (uint32_t)f[13 + i]; q.written_value[i] = (uint32_t)f[21 + i]; }
    q.flags = ZKC_LQ_FLAGS((uint32_t)f[29], (uint32_t)f[33], (uint32_t)f[30], (uint32_t)f[31], (uint32_t)f[32]);
    q.tx_number_in_block = (uint32_t)f[34]; q.timestamp = (uint32_t)f[35];
    uint64_t e[20], enc[20];
    lq_encode(q, e);
#pragma unroll
    for (int i = 0; i < 20; i++) { enc[i] = TR(ZKC_DMX_ENC + i); if (enc[i] != e[i]) bad |= ZKC_DMXV_ENCODING; }
    // the popped queue: is_empty <=> previous length == 0, length decrements on a pop, the head only moves on a pop
    const uint64_t len_prev = first ? d->iq0.length : TP(ZKC_DMX_LEN), len = TR(ZKC_DMX_LEN);
    if (is_empty != (uint64_t)(len_prev == 0) || len + execute != len_prev) bad |= ZKC_DMXV_QUEUE_LEN;
    uint64_t head[4], head_prev[4];
    bool same = true;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        head[i] = TR(ZKC_DMX_HEAD + i);
        head_prev[i] = first ? d->iq0.head[i] : TP(ZKC_DMX_HEAD + i);
        same &= head[i] == head_prev[i];
        if (head[i] >= GL_P) bad |= ZKC_DMXV_BOOLEAN;
    }
    if (!execute && !same) bad |= ZKC_DMXV_QUEUE_LEN;
    // :285-360 classification
    const uint64_t aux = f[29], shard = f[33];
    uint64_t is_aux[4], is_addr[3], sum_aux = 0;
    const bool small = (f[1] | f[2] | f[3] | f[4]) == 0;
    uint64_t flags_or = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        is_aux[i] = TR(ZKC_DMX_IS_AUX + i); flags_or |= is_aux[i]; sum_aux += is_aux[i];
        if (is_aux[i] != (uint64_t)(aux == d->opt.aux_bytes[i])) bad |= ZKC_DMXV_FLAGS;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        is_addr[i] = TR(ZKC_DMX_IS_ADDRESS + i); flags_or |= is_addr[i];
        if (is_addr[i] != (uint64_t)(small && f[0] == d->opt.precompile_addresses[i])) bad |= ZKC_DMXV_FLAGS;
    }
    const uint64_t rollup = TR(ZKC_DMX_IS_ROLLUP_SHARD), porter = TR(ZKC_DMX_EXECUTE_PORTER_STORAGE), is_bitmask = TR(ZKC_DMX_IS_BITMASK);
    if ((flags_or | rollup | porter | is_bitmask) > 1 || rollup != (uint64_t)(shard == 0) || porter != (is_aux[0] & (1 - rollup) & execute) ||
        is_bitmask != (uint64_t)(sum_aux == 1))
        bad |= ZKC_DMXV_FLAGS;
    const uint64_t want_bit[DMX_Q] = {is_aux[0] & rollup & execute, is_aux[1] & execute, is_aux[2] & execute,
                                      is_aux[3] & is_addr[0] & execute, is_aux[3] & is_addr[1] & execute, is_aux[3] & is_addr[2] & execute};
    uint64_t bit[DMX_Q];
    int sel = 0;
    bool any = false;
#pragma unroll
    for (int k = 0; k < DMX_Q; k++) {
        bit[k] = TR(ZKC_DMX_BITMASK + k);
        if (bit[k] != want_bit[k]) bad |= ZKC_DMXV_FLAGS;
        if (bit[k] & 1) { sel = k; any = true; }
    }
    // enforcements: no porter storage (:304-305), exactly one aux class on an executed row (:383-384)
    if (porter | (execute & (1 - (is_bitmask & 1)))) bad |= ZKC_DMXV_ENFORCE;
    // push_with_optimize (:401-447): the selected queue's state before the push, every queue's tail / length after it
    uint64_t r0[12], r1[12], r2[12], exec_tail[4];
#pragma unroll
    for (int i = 0; i < 12; i++) { r0[i] = TR(ZKC_DMX_PUSH_ROUND0 + i); r1[i] = TR(ZKC_DMX_PUSH_ROUND1 + i); r2[i] = TR(ZKC_DMX_PUSH_ROUND2 + i); }
#pragma unroll
    for (int i = 0; i < 4; i++) exec_tail[i] = TR(ZKC_DMX_EXEC_TAIL + i);
    const uint64_t exec_len = TR(ZKC_DMX_EXEC_LEN);
#pragma unroll
    for (int k = 0; k < DMX_Q; k++) {
        const uint64_t lp = first ? d->oq0[k].length : TP(ZKC_DMX_QUEUE_LENS + k);
        if (TR(ZKC_DMX_QUEUE_LENS + k) != lp + bit[k]) bad |= ZKC_DMXV_OUTPUT_QUEUES;
        if (k == sel && exec_len != lp) bad |= ZKC_DMXV_OUTPUT_QUEUES;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint64_t tp = first ? d->oq0[k].tail[i] : TP(ZKC_DMX_QUEUE_TAILS + 4 * k + i);
            if (k == sel && exec_tail[i] != tp) bad |= ZKC_DMXV_OUTPUT_QUEUES;
            if (TR(ZKC_DMX_QUEUE_TAILS + 4 * k + i) != ((any && k == sel) ? r2[i] : tp)) bad |= ZKC_DMXV_OUTPUT_QUEUES;
        }
    }
    if (ROUND_FUNCTION) {
        uint64_t st[12];
#pragma unroll
        for (int i = 0; i < 12; i++) st[i] = i < 8 ? enc[i] : 0;
        poseidon2_permute(st);
#pragma unroll
        for (int i = 0; i < 12; i++) if (st[i] != r0[i]) bad |= ZKC_DMXV_ROUND_FUNCTION;
#pragma unroll
        for (int i = 0; i < 8; i++) st[i] = enc[8 + i];
        poseidon2_permute(st);
#pragma unroll
        for (int i = 0; i < 12; i++) if (st[i] != r1[i]) bad |= ZKC_DMXV_ROUND_FUNCTION;
        uint64_t sp[12];
#pragma unroll
        for (int i = 0; i < 12; i++) sp[i] = st[i];
#pragma unroll
        for (int i = 0; i < 4; i++) { st[i] = enc[16 + i]; st[4 + i] = exec_tail[i]; }  // the push: absorbs the selected queue's tail
        poseidon2_permute(st);
#pragma unroll
        for (int i = 0; i < 12; i++) if (st[i] != r2[i]) bad |= ZKC_DMXV_ROUND_FUNCTION;
        if (execute) {  // the pop: the same two rounds, then the head before the pop
#pragma unroll
            for (int i = 0; i < 4; i++) { sp[i] = enc[16 + i]; sp[4 + i] = head_prev[i]; }
            poseidon2_permute(sp);
#pragma unroll
            for (int i = 0; i < 4; i++) if (sp[i] != head[i]) bad |= ZKC_DMXV_ROUND_FUNCTION;
        }
    } else {
        uint64_t big = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) big |= (uint64_t)(r0[i] >= GL_P) | (uint64_t)(r1[i] >= GL_P) | (uint64_t)(r2[i] >= GL_P);
        if (big) bad |= ZKC_DMXV_BOOLEAN;
    }
#undef TR
#undef TP
    if (bad) {
        atomicAdd(violations, 1ull);
        atomicOr(&d->failed_checks, bad);
        atomicMin(&d->first_bad, ((unsigned long long)row << 16) | bad);
    }
}

}  // namespace zkc

using namespace zkc;

extern "C" int zkc_demux_log_queue_entry_point(zkc_ctx *ctx, zkc_demux_closed_form *io, const zkc_log_query *records,
                                               const uint64_t *prev_tails, size_t n_records, const uint64_t *output_tails,
                                               const size_t n_output_tails[ZKC_DEMUX_NUM_QUEUES], size_t limit,
                                               const zkc_demux_options *options, int on_device, uint64_t *trace,
                                               uint64_t commitment[ZKC_COMMITMENT_LEN], zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    auto invalid = [&]() { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; };
    if (!ctx || !io || !commitment || (n_records && !records) || limit > 0x7FFFFFFFull || (output_tails && !n_output_tails))
        return invalid();
    zkc_demux_options opt;
    memset(&opt, 0, sizeof opt);
    if (options) opt = *options;
    if (!opt.custom_constants) {
        const uint32_t aux[4] = {0, 1, 2, 3}, addr[3] = {0x8010u, 0x02u, 0x01u};
        memcpy(opt.aux_bytes, aux, sizeof aux);
        memcpy(opt.precompile_addresses, addr, sizeof addr);
    }
    // at most one output queue per record: the constants must be pairwise distinct
    for (int a = 0; a < 4; a++)
        for (int b = a + 1; b < 4; b++)
            if (opt.aux_bytes[a] == opt.aux_bytes[b] || opt.aux_bytes[a] > 0xFF || opt.aux_bytes[b] > 0xFF) return invalid();
    for (int a = 0; a < 3; a++)
        for (int b = a + 1; b < 3; b++)
            if (opt.precompile_addresses[a] == opt.precompile_addresses[b]) return invalid();
    const zkc_queue_state4 &iq = io->start_flag ? io->initial_log_queue_state : io->hidden_fsm_input.initial_log_queue_state;
    const size_t need = limit < iq.length ? limit : iq.length;
    if (n_records < need || (need && !prev_tails)) return invalid();
    const bool in_dev = on_device & ZKC_INPUTS_ON_DEVICE, trace_dev = on_device & ZKC_TRACE_ON_DEVICE;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    const size_t tiles = (limit + SCAN_THREADS - 1) / SCAN_THREADS;
    const bool have_tails = output_tails != nullptr;
    size_t total_tails = 0;
    if (have_tails) for (int q = 0; q < DMX_Q; q++) total_tails += n_output_tails[q];
    else total_tails = (size_t)DMX_Q * limit;
    size_t bytes = zkc_carver::bytes(1, sizeof(DmxDev)) + zkc_carver::bytes(1, sizeof(ScanGlobal)) +
                   zkc_carver::bytes(tiles + 1, sizeof(DmxTile)) + zkc_carver::bytes(limit * 8 + 8, 8) +
                   zkc_carver::bytes(limit + 1, 4) + 2 * zkc_carver::bytes((size_t)DMX_Q * limit + DMX_Q, 4);
    if (!in_dev) bytes += zkc_carver::bytes(need + 1, sizeof(zkc_log_query)) + zkc_carver::bytes(need * 4 + 4, 8);
    if (!in_dev || !have_tails) bytes += zkc_carver::bytes(total_tails * 4 + 4, 8);
    if (trace && !trace_dev) bytes += zkc_carver::bytes((size_t)ZKC_DMX_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    DmxDev *h = (DmxDev *)ctx->pinned(sizeof(DmxDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; status->cuda_error = (int)cudaErrorMemoryAllocation; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    DmxDev *d = cv.take<DmxDev>(1);
    ScanGlobal *sg = cv.take<ScanGlobal>(1);
    DmxTile *ts = cv.take<DmxTile>(tiles + 1);
    uint64_t *r2in = cv.take<uint64_t>(limit * 8 + 8);
    uint32_t *meta = cv.take<uint32_t>(limit + 1);
    uint32_t *counts = cv.take<uint32_t>((size_t)DMX_Q * limit + DMX_Q);
    uint32_t *lists = cv.take<uint32_t>((size_t)DMX_Q * limit + DMX_Q);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(DmxDev));
    h->io = *io;
    h->opt = opt;
    h->n_records = n_records; h->limit = limit;
    h->first_bad = ~0ull;
    size_t off = 0;
    for (int q = 0; q < DMX_Q; q++) {
        h->tails_base[q] = have_tails ? off : (size_t)q * limit;
        h->n_tails[q] = have_tails ? n_output_tails[q] : limit;
        off += have_tails ? n_output_tails[q] : 0;
    }
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(DmxDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(sg, 0, (char *)(ts + tiles + 1) - (char *)sg, s));
    const zkc_log_query *dr = records;
    const uint64_t *dp = prev_tails, *dtails = output_tails;
    uint64_t *dtrace = trace;
    if (!in_dev) {
        zkc_log_query *br = cv.take<zkc_log_query>(need + 1);
        uint64_t *bp = cv.take<uint64_t>(need * 4 + 4);
        if (need) {
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(br, records, need * sizeof(zkc_log_query), cudaMemcpyHostToDevice, s));
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bp, prev_tails, need * 32, cudaMemcpyHostToDevice, s));
        }
        dr = br; dp = bp;
    }
    if (!in_dev || !have_tails) {
        uint64_t *bt = cv.take<uint64_t>(total_tails * 4 + 4);
        if (have_tails && total_tails)
            ZKC_CUDA(ctx, status, cudaMemcpyAsync(bt, output_tails, total_tails * 32, cudaMemcpyHostToDevice, s));
        dtails = bt;
    }
    if (trace && !trace_dev) dtrace = cv.take<uint64_t>((size_t)ZKC_DMX_NUM_COLS * limit);

    ZKC_LAUNCH(ctx, "dmx_prologue", dmx_prologue_kernel, 1, 96, 0, d);
    if (tiles) {
        ZKC_LAUNCH(ctx, "dmx_rows", dmx_rows_kernel, (unsigned)tiles, SCAN_THREADS, 0, d, dr, dp, dtrace, r2in, meta, counts, lists, sg, ts);
        if (!have_tails) ZKC_LAUNCH(ctx, "dmx_chain", dmx_chain_kernel, 1, 32 * DMX_Q, 0, d, r2in, lists, (uint64_t *)dtails);
        ZKC_LAUNCH(ctx, "dmx_push", dmx_push_kernel, (unsigned)((limit + 255) / 256), 256, 0, d, r2in, meta, counts, dtails, dtrace);
    }
    ZKC_LAUNCH(ctx, "dmx_finalize", dmx_finalize_kernel, 1, 32, 0, d, dtails);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(DmxDev), cudaMemcpyDeviceToHost, s));
    if (!trace_dev && trace && limit)
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(trace, dtrace, (size_t)ZKC_DMX_NUM_COLS * limit * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    io->hidden_fsm_output = h->io.hidden_fsm_output;
    memcpy(io->output_queue_states, h->io.output_queue_states, sizeof io->output_queue_states);
    io->completion_flag = h->io.completion_flag;
    memcpy(commitment, h->commitment, 32);
    *status = h->status;
    return status->code;
}

extern "C" int zkc_demux_log_queue_check_trace(zkc_ctx *ctx, const zkc_demux_closed_form *io, const zkc_demux_options *options, const uint64_t *trace,
                                               size_t limit, uint32_t gates, int on_device, uint64_t *violations, zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !io || !violations || (limit && !trace)) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    zkc_demux_options opt;
    memset(&opt, 0, sizeof opt);
    if (options) opt = *options;
    if (!opt.custom_constants) {
        const uint32_t aux[4] = {0, 1, 2, 3}, addr[3] = {0x8010u, 0x02u, 0x01u};
        memcpy(opt.aux_bytes, aux, sizeof aux);
        memcpy(opt.precompile_addresses, addr, sizeof addr);
    }
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    size_t bytes = zkc_carver::bytes(1, sizeof(DmxDev)) + zkc_carver::bytes(1, 8);
    if (!on_device) bytes += zkc_carver::bytes((size_t)ZKC_DMX_NUM_COLS * limit, 8);
    void *blk = ctx->scratch(bytes);
    DmxDev *h = (DmxDev *)ctx->pinned(sizeof(DmxDev) + 8);
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    DmxDev *d = cv.take<DmxDev>(1);
    unsigned long long *dviol = cv.take<unsigned long long>(1);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof(DmxDev));
    h->io = *io;
    h->opt = opt;
    h->limit = limit;
    h->first_bad = ~0ull;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof(DmxDev), cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemsetAsync(dviol, 0, 8, s));
    const uint64_t *dt = trace;
    if (!on_device && limit) {
        uint64_t *b = cv.take<uint64_t>((size_t)ZKC_DMX_NUM_COLS * limit);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(b, trace, (size_t)ZKC_DMX_NUM_COLS * limit * 8, cudaMemcpyHostToDevice, s));
        dt = b;
    }
    ZKC_LAUNCH(ctx, "dmx_prologue", dmx_prologue_kernel, 1, 96, 0, d);
    if (limit) {
        const unsigned grid = (unsigned)((limit + 127) / 128);
        if (gates == 0 || (gates & ZKC_GATES_ROUND_FUNCTION)) ZKC_LAUNCH(ctx, "dmx_check_rf", dmx_check_kernel<true>, grid, 128, 0, d, dviol, dt);
        else ZKC_LAUNCH(ctx, "dmx_check", dmx_check_kernel<false>, grid, 128, 0, d, dviol, dt);
    }
    ZKC_CUDA(ctx, status, cudaGetLastError());
    unsigned long long *hviol = (unsigned long long *)(h + 1);
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof(DmxDev), cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(hviol, dviol, 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    *violations = *hviol;
    status->failed_checks = h->failed_checks;
    if (*hviol) {
        status->code = ZKC_ERR_UNSATISFIED;
        status->first_bad_row = (int64_t)(h->first_bad >> 16);
    }
    return status->code;
}
