// Context management and the stand-alone primitives of the C ABI (include/zkc_b200.h).
#include "ctx.cuh"
#include "poseidon2.cuh"
#include "scan.cuh"

using namespace zkc;

extern "C" const char *zkc_version(void) { return "zkc_b200 0.1 (sm_100a)"; }

extern "C" int zkc_create(int device, zkc_ctx **out) {
    if (!out) return ZKC_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return ZKC_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return ZKC_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ZKC_ERR_CUDA;
    if (prop.major < 10) return ZKC_ERR_NO_DEVICE;  // sm_100a cubins only
    zkc_ctx *c = new zkc_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return ZKC_OK;
}

extern "C" void zkc_destroy(zkc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->prof_resolve();
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    delete ctx;
}

extern "C" int zkc_set_stream(zkc_ctx *ctx, void *cuda_stream) {
    if (!ctx) return ZKC_ERR_INVALID_ARGUMENT;
    ctx->stream = (cudaStream_t)cuda_stream;
    return ZKC_OK;
}
extern "C" uint64_t zkc_launch_count(const zkc_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int zkc_sm_count(const zkc_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" int zkc_profile_enable(zkc_ctx *ctx, int enable) {
    if (!ctx) return ZKC_ERR_INVALID_ARGUMENT;
    ctx->profiling = enable != 0;
    return ZKC_OK;
}
extern "C" int zkc_profile_query(zkc_ctx *ctx, const char *name, double *ms_total, uint64_t *launches) {
    if (!ctx || !name) return ZKC_ERR_INVALID_ARGUMENT;
    ctx->prof_resolve();
    auto it = ctx->prof.find(name);
    if (ms_total) *ms_total = it == ctx->prof.end() ? 0.0 : it->second.first;
    if (launches) *launches = it == ctx->prof.end() ? 0 : it->second.second;
    return ZKC_OK;
}
extern "C" int zkc_profile_reset(zkc_ctx *ctx) {
    if (!ctx) return ZKC_ERR_INVALID_ARGUMENT;
    ctx->prof_resolve();
    ctx->prof.clear();
    return ZKC_OK;
}
extern "C" void *zkc_host_alloc(size_t bytes) {
    void *p = nullptr;
    return cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
extern "C" void zkc_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- Poseidon2 batch ------------------------------------------------------------------------
namespace zkc {
__global__ void __launch_bounds__(128) poseidon2_batch_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s[12];
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(in + 12 * i);
#pragma unroll
    for (int j = 0; j < 6; j++) { ulonglong2 v = __ldg(src + j); s[2 * j] = v.x; s[2 * j + 1] = v.y; }
    poseidon2_permute(s);
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out + 12 * i);
#pragma unroll
    for (int j = 0; j < 6; j++) dst[j] = make_ulonglong2(s[2 * j], s[2 * j + 1]);
}

__global__ void __launch_bounds__(128) commit_encoding_kernel(const uint64_t *__restrict__ in, size_t len, size_t n, uint64_t *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t o[4];
    commit_encoding_dev(in + i * len, (int)len, o);
    for (int j = 0; j < 4; j++) out[4 * i + j] = o[j];
}
}  // namespace zkc

extern "C" int zkc_poseidon2_permute(zkc_ctx *ctx, const uint64_t *states_in, uint64_t *states_out, size_t n, int on_device) {
    if (!ctx || (n && (!states_in || !states_out))) return ZKC_ERR_INVALID_ARGUMENT;
    if (!n) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    const uint64_t *din = states_in;
    uint64_t *dout = states_out;
    if (!on_device) {
        uint64_t *buf = (uint64_t *)ctx->scratch(2 * n * 96);  // separate output region: the kernel's pointers are __restrict__
        if (!buf) return ZKC_ERR_CUDA;
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(buf, states_in, n * 96, cudaMemcpyHostToDevice, ctx->stream));
        din = buf; dout = buf + n * 12;
    }
    ZKC_LAUNCH(ctx, "poseidon2_batch", poseidon2_batch_kernel, (unsigned)((n + 127) / 128), 128, 0, din, dout, n);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    if (!on_device) {
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(states_out, dout, n * 96, cudaMemcpyDeviceToHost, ctx->stream));
        ZKC_CUDA(ctx, st, cudaStreamSynchronize(ctx->stream));
    }
    return ZKC_OK;
}

extern "C" int zkc_commit_encoding(zkc_ctx *ctx, const uint64_t *inputs, size_t len, size_t n_items, uint64_t *out, int on_device) {
    if (!ctx || !out || (len && n_items && !inputs) || len > 0x7fffffff) return ZKC_ERR_INVALID_ARGUMENT;
    if (!n_items) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    const uint64_t *din = inputs;
    uint64_t *dout = out;
    if (!on_device) {
        const size_t inb = zkc_carver::bytes(n_items * len + 1, 8);
        char *buf = (char *)ctx->scratch(inb + n_items * 32);
        if (!buf) return ZKC_ERR_CUDA;
        if (len) ZKC_CUDA(ctx, st, cudaMemcpyAsync(buf, inputs, n_items * len * 8, cudaMemcpyHostToDevice, ctx->stream));
        din = (uint64_t *)buf; dout = (uint64_t *)(buf + inb);
    }
    ZKC_LAUNCH(ctx, "commit_encoding", commit_encoding_kernel, (unsigned)((n_items + 127) / 128), 128, 0, din, len, n_items, dout);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    if (!on_device) {
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(out, dout, n_items * 32, cudaMemcpyDeviceToHost, ctx->stream));
        ZKC_CUDA(ctx, st, cudaStreamSynchronize(ctx->stream));
    }
    return ZKC_OK;
}

// ---- element-wise field operations (diagnostic entry: pins the PTX carry chains of gl.cuh) ---------
namespace zkc {
__global__ void field_ops_kernel(const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t n, uint64_t *mul,
                                 uint64_t *add, uint64_t *sub, uint64_t *fma) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mul[i] = gl_mul(a[i], b[i]);
    add[i] = gl_add(a[i], b[i]);
    sub[i] = gl_sub(a[i], b[i]);
    fma[i] = gl_fma(a[i], b[i], c[i]);
}
}  // namespace zkc

extern "C" int zkc_field_ops(zkc_ctx *ctx, const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t n,
                             uint64_t *out_mul, uint64_t *out_add, uint64_t *out_sub, uint64_t *out_fma) {
    if (!ctx || !a || !b || !c || !out_mul || !out_add || !out_sub || !out_fma) return ZKC_ERR_INVALID_ARGUMENT;
    if (!n) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    uint64_t *buf = (uint64_t *)ctx->scratch(7 * n * 8);
    if (!buf) return ZKC_ERR_CUDA;
    cudaStream_t s = ctx->stream;
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(buf, a, n * 8, cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(buf + n, b, n * 8, cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(buf + 2 * n, c, n * 8, cudaMemcpyHostToDevice, s));
    ZKC_LAUNCH(ctx, "field_ops", field_ops_kernel, (unsigned)((n + 255) / 256), 256, 0, buf, buf + n, buf + 2 * n, n,
               buf + 3 * n, buf + 4 * n, buf + 5 * n, buf + 6 * n);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(out_mul, buf + 3 * n, n * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(out_add, buf + 4 * n, n * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(out_sub, buf + 5 * n, n * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(out_fma, buf + 6 * n, n * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    return ZKC_OK;
}

// ---- stand-alone grand product ---------------------------------------------------------------
namespace zkc {
struct GpParams {
    uint64_t ch[2][21];
    uint64_t acc0[4];  // lhs0, lhs1, rhs0, rhs1 (ABI order)
    uint64_t acc_final[4];
};

template <int ENC>
__global__ void __launch_bounds__(SCAN_THREADS)
grand_product_kernel(GpParams *gp, const uint64_t *__restrict__ lhs, const uint64_t *__restrict__ rhs,
                     const uint8_t *__restrict__ flags, size_t rows, uint64_t *__restrict__ acc_out,
                     uint64_t *__restrict__ chain_out, ScanGlobal *sg, TileState *tiles) {
    __shared__ ScanShared sh;
    __shared__ uint64_t ch[2][ENC + 1];
    if (threadIdx.x < 2 * (ENC + 1)) ch[threadIdx.x / (ENC + 1)][threadIdx.x % (ENC + 1)] = gp->ch[threadIdx.x / (ENC + 1)][threadIdx.x % (ENC + 1)];
    const unsigned int tile = scan_take_ticket(sg, sh);
    const size_t row = (size_t)tile * SCAN_THREADS + threadIdx.x;
    const bool in_range = row < rows;
    ScanVal v = scan_identity();
    if (in_range) {
        uint64_t c[4] = {ch[0][ENC], ch[0][ENC], ch[1][ENC], ch[1][ENC]};  // rep*2 + side
#pragma unroll
        for (int i = 0; i < ENC; i++) {
            const uint64_t l = __ldg(lhs + (size_t)i * rows + row), r = __ldg(rhs + (size_t)i * rows + row);
#pragma unroll
            for (int rep = 0; rep < 2; rep++) {
                c[rep * 2 + 0] = gl_fma(l, ch[rep][i], c[rep * 2 + 0]);
                c[rep * 2 + 1] = gl_fma(r, ch[rep][i], c[rep * 2 + 1]);
                if (chain_out) {
                    chain_out[((size_t)(rep * 2 + 0) * ENC + i) * rows + row] = c[rep * 2 + 0];
                    chain_out[((size_t)(rep * 2 + 1) * ENC + i) * rows + row] = c[rep * 2 + 1];
                }
            }
        }
        const bool f = flags ? flags[row] != 0 : true;
        if (f) { v.p[0] = c[0]; v.p[1] = c[2]; v.p[2] = c[1]; v.p[3] = c[3]; }  // -> lhs0, lhs1, rhs0, rhs1
    }
    ScanVal init;
#pragma unroll
    for (int i = 0; i < 4; i++) init.p[i] = gp->acc0[i];
    init.c = 0;
    ScanVal incl;
    scan_tile(v, tile, init, tiles, sh, incl);
    if (in_range) {
        if (acc_out) {
#pragma unroll
            for (int i = 0; i < 4; i++) acc_out[(size_t)i * rows + row] = incl.p[i];
        }
        if (row == rows - 1) {
#pragma unroll
            for (int i = 0; i < 4; i++) gp->acc_final[i] = incl.p[i];
        }
    }
}
}  // namespace zkc

namespace zkc {
// two rows per thread and column: 128-bit loads / stores over the column-major accumulators
__global__ void __launch_bounds__(256)
scale_accumulators_kernel(uint64_t *__restrict__ acc, size_t rows, uint64_t f0, uint64_t f1, uint64_t f2, uint64_t f3, int n_cols) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= rows) return;
    const int c = blockIdx.y;
    if (c >= n_cols) return;
    const uint64_t f = c == 0 ? f0 : (c == 1 ? f1 : (c == 2 ? f2 : f3));
    uint64_t *p = acc + (size_t)c * rows + i;
    if (i + 1 < rows && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        ulonglong2 v = *reinterpret_cast<ulonglong2 *>(p);
        v.x = gl_mul(v.x, f); v.y = gl_mul(v.y, f);
        *reinterpret_cast<ulonglong2 *>(p) = v;
    } else {
        p[0] = gl_mul(p[0], f);
        if (i + 1 < rows) p[1] = gl_mul(p[1], f);
    }
}
}  // namespace zkc

extern "C" int zkc_scale_accumulators(zkc_ctx *ctx, uint64_t *acc, size_t n_cols, size_t rows, const uint64_t *factors, int on_device) {
    if (!ctx || !factors || n_cols > 4 || (n_cols && rows && !acc)) return ZKC_ERR_INVALID_ARGUMENT;
    if (!rows || !n_cols) return ZKC_OK;
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    uint64_t f[4] = {1, 1, 1, 1};
    for (size_t c = 0; c < n_cols; c++) f[c] = factors[c] % 0xFFFFFFFF00000001ull;
    uint64_t *d = acc;
    cudaStream_t s = ctx->stream;
    if (!on_device) {
        d = (uint64_t *)ctx->scratch(n_cols * rows * 8);
        if (!d) return ZKC_ERR_CUDA;
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(d, acc, n_cols * rows * 8, cudaMemcpyHostToDevice, s));
    }
    ZKC_LAUNCH(ctx, "scale_accumulators", scale_accumulators_kernel, dim3((unsigned)((rows / 2 + 256) / 256), (unsigned)n_cols), 256, 0, d, rows, f[0], f[1],
               f[2], f[3], (int)n_cols);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    if (!on_device) ZKC_CUDA(ctx, st, cudaMemcpyAsync(acc, d, n_cols * rows * 8, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    return ZKC_OK;
}

namespace zkc {
struct ColCheckDev { unsigned long long violations, first_bad; uint32_t failed_checks, pad; };
// the allocation checks of a whole trace: thread = R consecutive rows (R = 2: one 128-bit load per column), loop over the
// columns with the class table in shared memory, 8 independent loads in flight per iteration; a pure HBM stream
template <int R>
__global__ void __launch_bounds__(256)
check_columns_kernel(ColCheckDev *out, const uint64_t *__restrict__ trace, size_t n_cols, size_t rows, const uint8_t *__restrict__ col_class) {
    extern __shared__ uint8_t cls[];
    for (size_t c = threadIdx.x; c < n_cols; c += blockDim.x) cls[c] = col_class[c];
    __syncthreads();
    const size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * R;
    if (row >= rows) return;
    uint32_t bad[R];
    unsigned long long where[R];
#pragma unroll
    for (int r = 0; r < R; r++) { bad[r] = 0; where[r] = ~0ull; }
    constexpr uint64_t BOUND[ZKC_COL_NUM_CLASSES] = {GL_P, 2ull, 1ull << 8, 1ull << 16, 1ull << 32};
    const uint64_t *t = trace + row;
#pragma unroll 8
    for (size_t c = 0; c < n_cols; c++) {
        uint64_t v[R];
        if constexpr (R == 2) { const ulonglong2 q = __ldg(reinterpret_cast<const ulonglong2 *>(t + c * rows)); v[0] = q.x; v[1] = q.y; }
        else v[0] = __ldg(t + c * rows);
        const uint32_t k = cls[c] < ZKC_COL_NUM_CLASSES ? cls[c] : 0u;
        const uint64_t bound = k == 0 ? BOUND[0] : (k == 1 ? BOUND[1] : (k == 2 ? BOUND[2] : (k == 3 ? BOUND[3] : BOUND[4])));
#pragma unroll
        for (int r = 0; r < R; r++)
            if (v[r] >= bound) { bad[r] |= 1u << k; if (where[r] == ~0ull) where[r] = c; }
    }
#pragma unroll
    for (int r = 0; r < R; r++)
        if (bad[r] && row + r < rows) {
            atomicAdd(&out->violations, 1ull);
            atomicOr(&out->failed_checks, bad[r]);
            atomicMin(&out->first_bad, ((unsigned long long)(row + r) << 20) | (where[r] & 0xFFFFFull));
        }
}
}  // namespace zkc

extern "C" int zkc_check_trace_columns(zkc_ctx *ctx, const uint64_t *trace, size_t n_cols, size_t rows, const uint8_t *col_class, int on_device,
                                       uint64_t *violations, uint32_t *first_bad_column, zkc_status *status) {
    zkc_status local;
    if (!status) status = &local;
    *status = zkc_status{ZKC_OK, 0, -1, 0, 0};
    if (!ctx || !violations || !col_class || n_cols > 40000 || ((n_cols * rows) && !trace)) { status->code = ZKC_ERR_INVALID_ARGUMENT; return ZKC_ERR_INVALID_ARGUMENT; }
    *violations = 0;
    if (first_bad_column) *first_bad_column = 0;
    if (!rows || !n_cols) return ZKC_OK;
    ZKC_CUDA(ctx, status, cudaSetDevice(ctx->device));
    size_t bytes = zkc_carver::bytes(1, sizeof(ColCheckDev)) + zkc_carver::bytes(n_cols, 1);
    if (!on_device) bytes += zkc_carver::bytes(n_cols * rows, 8);
    void *blk = ctx->scratch(bytes);
    ColCheckDev *h = (ColCheckDev *)ctx->pinned(sizeof(ColCheckDev));
    if (!blk || !h) { status->code = ZKC_ERR_CUDA; return ZKC_ERR_CUDA; }
    zkc_carver cv(blk);
    ColCheckDev *d = cv.take<ColCheckDev>(1);
    uint8_t *dcls = cv.take<uint8_t>(n_cols);
    cudaStream_t s = ctx->stream;
    h->violations = 0; h->first_bad = ~0ull; h->failed_checks = 0; h->pad = 0;
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(d, h, sizeof *h, cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(dcls, col_class, n_cols, cudaMemcpyHostToDevice, s));  // the class table is always host data
    const uint64_t *dt = trace;
    if (!on_device) {
        uint64_t *b = cv.take<uint64_t>(n_cols * rows);
        ZKC_CUDA(ctx, status, cudaMemcpyAsync(b, trace, n_cols * rows * 8, cudaMemcpyHostToDevice, s));
        dt = b;
    }
    if (rows % 2 == 0 && (reinterpret_cast<uintptr_t>(dt) & 15) == 0)
        ZKC_LAUNCH(ctx, "check_columns", check_columns_kernel<2>, (unsigned)((rows / 2 + 255) / 256), 256, n_cols, d, dt, n_cols, rows, dcls);
    else
        ZKC_LAUNCH(ctx, "check_columns", check_columns_kernel<1>, (unsigned)((rows + 255) / 256), 256, n_cols, d, dt, n_cols, rows, dcls);
    ZKC_CUDA(ctx, status, cudaGetLastError());
    ZKC_CUDA(ctx, status, cudaMemcpyAsync(h, d, sizeof *h, cudaMemcpyDeviceToHost, s));
    ZKC_CUDA(ctx, status, cudaStreamSynchronize(s));
    *violations = h->violations;
    status->failed_checks = h->failed_checks;
    if (h->violations) {
        status->code = ZKC_ERR_UNSATISFIED;
        status->first_bad_row = (int64_t)(h->first_bad >> 20);
        if (first_bad_column) *first_bad_column = (uint32_t)(h->first_bad & 0xFFFFFull);
    }
    return status->code;
}

extern "C" int zkc_accumulate_grand_products(zkc_ctx *ctx, const uint64_t *lhs_enc, const uint64_t *rhs_enc,
                                             const uint8_t *should_acc, size_t enc_len, size_t rows,
                                             const uint64_t *challenges, const uint64_t acc_in[4], uint64_t *acc_out,
                                             uint64_t *chain_out, uint64_t acc_final[4], int on_device) {
    if (!ctx || !challenges || !acc_in || !acc_final || (enc_len != 8 && enc_len != 20) || (rows && (!lhs_enc || !rhs_enc)))
        return ZKC_ERR_INVALID_ARGUMENT;
    if (!rows) { memcpy(acc_final, acc_in, 32); return ZKC_OK; }
    zkc_status *st = nullptr;
    ZKC_CUDA(ctx, st, cudaSetDevice(ctx->device));
    const size_t tiles = (rows + SCAN_THREADS - 1) / SCAN_THREADS;
    size_t bytes = zkc_carver::bytes(1, sizeof(GpParams)) + zkc_carver::bytes(1, sizeof(ScanGlobal)) + zkc_carver::bytes(tiles, sizeof(TileState));
    if (!on_device)
        bytes += 2 * zkc_carver::bytes(enc_len * rows, 8) + zkc_carver::bytes(rows, 1) + zkc_carver::bytes(4 * rows, 8) +
                 (chain_out ? zkc_carver::bytes(4 * enc_len * rows, 8) : 0);
    void *blk = ctx->scratch(bytes);
    GpParams *h = (GpParams *)ctx->pinned(sizeof(GpParams));
    if (!blk || !h) return ZKC_ERR_CUDA;
    zkc_carver cv(blk);
    GpParams *d = cv.take<GpParams>(1);
    ScanGlobal *sg = cv.take<ScanGlobal>(1);
    TileState *ts = cv.take<TileState>(tiles);
    cudaStream_t s = ctx->stream;
    memset(h, 0, sizeof *h);
    for (int rep = 0; rep < 2; rep++) memcpy(h->ch[rep], challenges + rep * (enc_len + 1), (enc_len + 1) * 8);
    memcpy(h->acc0, acc_in, 32);
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(d, h, sizeof *h, cudaMemcpyHostToDevice, s));
    ZKC_CUDA(ctx, st, cudaMemsetAsync(sg, 0, (char *)(ts + tiles) - (char *)sg, s));
    const uint64_t *dl = lhs_enc, *dr = rhs_enc;
    const uint8_t *df = should_acc;
    uint64_t *dacc = acc_out, *dchain = chain_out;
    if (!on_device) {
        uint64_t *bl = cv.take<uint64_t>(enc_len * rows), *br = cv.take<uint64_t>(enc_len * rows);
        uint8_t *bf = cv.take<uint8_t>(rows);
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(bl, lhs_enc, enc_len * rows * 8, cudaMemcpyHostToDevice, s));
        ZKC_CUDA(ctx, st, cudaMemcpyAsync(br, rhs_enc, enc_len * rows * 8, cudaMemcpyHostToDevice, s));
        if (should_acc) ZKC_CUDA(ctx, st, cudaMemcpyAsync(bf, should_acc, rows, cudaMemcpyHostToDevice, s));
        dl = bl; dr = br; df = should_acc ? bf : nullptr;
        dacc = acc_out ? cv.take<uint64_t>(4 * rows) : nullptr;
        dchain = chain_out ? cv.take<uint64_t>(4 * enc_len * rows) : nullptr;
    }
    if (enc_len == 8)
        ZKC_LAUNCH(ctx, "grand_product", grand_product_kernel<8>, (unsigned)tiles, SCAN_THREADS, 0, d, dl, dr, df, rows, dacc, dchain, sg, ts);
    else
        ZKC_LAUNCH(ctx, "grand_product", grand_product_kernel<20>, (unsigned)tiles, SCAN_THREADS, 0, d, dl, dr, df, rows, dacc, dchain, sg, ts);
    ZKC_CUDA(ctx, st, cudaGetLastError());
    ZKC_CUDA(ctx, st, cudaMemcpyAsync(h, d, sizeof *h, cudaMemcpyDeviceToHost, s));
    if (!on_device) {
        if (acc_out) ZKC_CUDA(ctx, st, cudaMemcpyAsync(acc_out, dacc, 4 * rows * 8, cudaMemcpyDeviceToHost, s));
        if (chain_out) ZKC_CUDA(ctx, st, cudaMemcpyAsync(chain_out, dchain, 4 * enc_len * rows * 8, cudaMemcpyDeviceToHost, s));
    }
    ZKC_CUDA(ctx, st, cudaStreamSynchronize(s));
    memcpy(acc_final, h->acc_final, 32);
    return ZKC_OK;
}
