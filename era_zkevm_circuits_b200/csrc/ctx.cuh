// Engine context shared by every entry point of the C ABI (include/zkc_b200.h): one context per
// GPU / process, one CUDA stream, grow-only device + pinned scratch, launch counter and optional
// per-kernel CUDA-event timing.  Nothing here falls back to the CPU: without a device
// zkc_create fails with ZKC_ERR_NO_DEVICE.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../../include/zkc_b200.h"

struct zkc_prof_rec {
    cudaEvent_t a, b;
    std::string name;
};

struct zkc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    bool profiling = false;
    std::vector<zkc_prof_rec> pending;
    std::vector<cudaEvent_t> event_pool;
    std::map<std::string, std::pair<double, uint64_t>> prof;
    // grow-only scratch
    void *d_scratch = nullptr;
    size_t d_scratch_bytes = 0;
    void *h_pinned = nullptr;
    size_t h_pinned_bytes = 0;
    int sm_count = 148;
    int last_cuda_error = 0;
    // copy streams of the pipelined host paths (H2D of chunk i+1 | kernels of chunk i | D2H of chunk i-1)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    // high-priority side stream for small latency-bound kernels that can run beside the main launches
    cudaStream_t aux = nullptr;
    cudaStream_t aux_stream() {
        if (!aux) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            if (cudaStreamCreateWithPriority(&aux, cudaStreamNonBlocking, hi) != cudaSuccess) aux = nullptr;
        }
        return aux;
    }
    bool copy_streams() {
        if (!copy_in && cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (!copy_out && cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking) != cudaSuccess) return false;
        return true;
    }

    void *scratch(size_t bytes) {
        if (bytes > d_scratch_bytes) {
            if (d_scratch) { cudaStreamSynchronize(stream); cudaFree(d_scratch); }
            size_t want = bytes + bytes / 4;
            if (cudaMalloc(&d_scratch, want) != cudaSuccess) { d_scratch = nullptr; d_scratch_bytes = 0; return nullptr; }
            d_scratch_bytes = want;
        }
        return d_scratch;
    }
    void *pinned(size_t bytes) {
        if (bytes > h_pinned_bytes) {
            if (h_pinned) { cudaStreamSynchronize(stream); cudaFreeHost(h_pinned); }
            if (cudaHostAlloc(&h_pinned, bytes, cudaHostAllocDefault) != cudaSuccess) { h_pinned = nullptr; h_pinned_bytes = 0; return nullptr; }
            h_pinned_bytes = bytes;
        }
        return h_pinned;
    }
    cudaEvent_t get_event() {
        if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    void prof_begin(const char *name) {
        launches++;
        if (!profiling) return;
        zkc_prof_rec r{get_event(), get_event(), name};
        cudaEventRecord(r.a, stream);
        pending.push_back(r);
    }
    void prof_end() {
        if (!profiling) return;
        cudaEventRecord(pending.back().b, stream);
    }
    void prof_resolve() {
        for (auto &r : pending) {
            cudaEventSynchronize(r.b);
            float ms = 0;
            cudaEventElapsedTime(&ms, r.a, r.b);
            auto &e = prof[r.name];
            e.first += ms; e.second += 1;
            event_pool.push_back(r.a); event_pool.push_back(r.b);
        }
        pending.clear();
    }
};

// carve a sub-allocation out of a scratch block (256-byte aligned)
struct zkc_carver {
    char *base; size_t off = 0;
    explicit zkc_carver(void *p) : base((char *)p) {}
    template <class T> T *take(size_t n) {
        T *r = (T *)(base + off);
        off += (n * sizeof(T) + 255) & ~(size_t)255;
        return r;
    }
    static size_t bytes(size_t n, size_t sz) { return (n * sz + 255) & ~(size_t)255; }
};

#define ZKC_LAUNCH(ctx, name, kernel, grid, block, smem, ...)                 \
    do {                                                                       \
        (ctx)->prof_begin(name);                                               \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);       \
        (ctx)->prof_end();                                                     \
    } while (0)

#define ZKC_CUDA(ctx, st, expr)                                                \
    do {                                                                       \
        cudaError_t e_ = (expr);                                               \
        if (e_ != cudaSuccess) {                                               \
            (ctx)->last_cuda_error = (int)e_;                                  \
            if (st) { (st)->code = ZKC_ERR_CUDA; (st)->cuda_error = (int)e_; } \
            return ZKC_ERR_CUDA;                                               \
        }                                                                      \
    } while (0)
