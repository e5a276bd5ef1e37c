// common.cuh -- shared plumbing for libbbocr.so (handle, error propagation, stream-ordered buffers, launch counting)
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/bbocr.h"

namespace bbocr {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error(code, buf);
}

#define CUDA_CHECK(expr)                                                                                   \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            ::bbocr::fail(BBOCR_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define ARG_CHECK(cond, ...)                                   \
    do {                                                       \
        if (!(cond)) ::bbocr::fail(BBOCR_E_ARG, __VA_ARGS__);  \
    } while (0)

// Host wait for a stream WITHOUT spinning: record an event created with cudaEventBlockingSync and sleep on it.
// cudaStreamSynchronize spins by default (cudaDeviceScheduleSpin on an otherwise idle context), and a handle keeps ~10 lane
// threads waiting on their streams; with one process per GPU on a shared host (8 ranks x 10 threads on 32 cores) the spinning
// waiters starve the threads that have launches to issue -- the 1 -> 8 GPU efficiency loss measured in round 1.
inline cudaError_t stream_sync(cudaStream_t st) {
    static const bool spin = getenv("BBOCR_SPIN_SYNC") != nullptr;          // A/B: the old spinning wait
    if (spin) return cudaStreamSynchronize(st);
    // one blocking event per stream (the lanes' streams live as long as their handle), created on first use
    static std::mutex mu;
    static std::map<cudaStream_t, std::pair<int, cudaEvent_t>> events;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaEvent_t ev = nullptr;
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = events.find(st);
        if (it != events.end() && it->second.first != dev) {                // a recycled stream handle on another device
            cudaEventDestroy(it->second.second);
            events.erase(it);
            it = events.end();
        }
        if (it == events.end()) {
            e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
            events[st] = std::make_pair(dev, ev);
        } else {
            ev = it->second.second;
        }
    }
    e = cudaEventRecord(ev, st);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ev);
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Stream-ordered device buffer (cudaMallocAsync pool: no device-wide sync on alloc/free).
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t s = nullptr;
    DevBuf() = default;
    DevBuf(size_t n, cudaStream_t st) { alloc(n, st); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes), s(o.s) { o.p = nullptr; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; bytes = o.bytes; s = o.s; o.p = nullptr; }
        return *this;
    }
    void alloc(size_t n, cudaStream_t st) {
        release();
        s = st;
        bytes = n ? n : 16;
        CUDA_CHECK(cudaMallocAsync(&p, bytes, s));
    }
    void release() {
        if (p) { cudaFreeAsync(p, s); p = nullptr; }
    }
    ~DevBuf() { release(); }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Pinned host staging buffer that grows on demand (one per lane; reused across pages).
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    void* get(size_t n) {
        if (n > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            cap = n + n / 4 + 4096;
            CUDA_CHECK(cudaMallocHost(&p, cap));
        }
        return p;
    }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

// ---- activation tensors: NHWC, element type float (FP32 mode) or __nv_bfloat16 (BF16 mode) --------------------------
struct Act {
    void* p = nullptr;
    void* lo = nullptr;               // non-null: split-precision tensor x = bf16(p) + bf16(lo)  (3 x bf16 products ~ FP32)
    int N = 0, H = 0, W = 0, C = 0;
    int64_t elems() const { return (int64_t)N * H * W * C; }
};

// Convolution parameters after load-time folding (BatchNorm -> per-channel scale/bias applied in the epilogue).
struct ConvW {
    int cin = 0, cout = 0, kh = 0, kw = 0, pad = 0, dil = 1;
    int cout_pad = 0;                 // cout rounded up to 16 (BF16 weight rows)
    float* w_f32 = nullptr;           // [kh*kw][cin][cout]          (FP32 path; cout contiguous)
    __nv_bfloat16* w_bf16 = nullptr;  // [kh*kw][cout_pad][cin]      (tcgen05 path; K-major rows)
    __nv_bfloat16* w_split = nullptr; // [kh*kw][cout_pad][2*cin] = [w_hi | w_lo]  (split precision: w = bf16 hi + bf16 lo)
    float* scale = nullptr;           // [cout_pad]
    float* bias = nullptr;            // [cout_pad]
};

struct LstmW {                        // one BidirectionalLSTM block (easyocr/model/modules.py)
    ConvW in_proj;                    // 256 -> 2048 : [fwd i,f,g,o | bwd i,f,g,o], bias = b_ih + b_hh
    float* w_hh = nullptr;            // [2][256 (k)][1024 (gate row)]  FP32, k-major for coalesced reads
    ConvW linear;                     // 512 -> 256
};

struct Handle;
void count_launch(Handle* h, int n = 1);

}  // namespace bbocr
