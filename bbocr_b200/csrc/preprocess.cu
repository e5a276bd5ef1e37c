// preprocess.cu -- stage 1 of the OCR path: the OpenCV/Pillow chain of
//   /root/reference/pipeline_demo/ocr_testing/preprocessing/image_preprocessor.py:17-160
// as bit-exact integer / unfused-FP32 CUDA kernels (HBM-bound byte work: coalesced, vectorised where alignment allows,
// shared-memory tiles for the stencils).  Float expressions that OpenCV/Pillow evaluate without FMA contraction use
// __fmul_rn/__fadd_rn so nvcc cannot fuse them.
#include "engine.h"

namespace bbocr {

// ------------------------------------------------------------------------------------------------------------------
// A2  cv2.cvtColor(BGR2GRAY): (B*3735 + G*19235 + R*9798 + 16384) >> 15          image_preprocessor.py:25-30
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) {
    return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

// contiguous rows: thread handles 16 pixels = 3 x 16-byte loads -> 1 x 16-byte store
__global__ void k_gray_packed(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t npix) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t p0 = t * 16;
    if (p0 + 16 <= npix) {
        const uint4* s = reinterpret_cast<const uint4*>(src + p0 * 3);
        uint4 a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2);
        uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {          // 4 pixels = 12 bytes = 3 words
            uint32_t w0 = w[q * 3], w1 = w[q * 3 + 1], w2 = w[q * 3 + 2];
            uint32_t g0 = gray_of(w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255);
            uint32_t g1 = gray_of(w0 >> 24, w1 & 255, (w1 >> 8) & 255);
            uint32_t g2 = gray_of((w1 >> 16) & 255, w1 >> 24, w2 & 255);
            uint32_t g3 = gray_of((w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24);
            o[q] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
        }
        *reinterpret_cast<uint4*>(dst + p0) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
        for (int64_t p = p0; p < npix; ++p)
            dst[p] = (uint8_t)gray_of(src[p * 3], src[p * 3 + 1], src[p * 3 + 2]);
    }
}

__global__ void k_gray_strided(const uint8_t* __restrict__ src, int stride, uint8_t* __restrict__ dst, int H, int W) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < W) {
        const uint8_t* s = src + (int64_t)y * stride + x * 3;
        dst[(int64_t)y * W + x] = (uint8_t)gray_of(s[0], s[1], s[2]);
    }
}

void pp_gray(Handle* h, cudaStream_t st, const uint8_t* bgr, int H, int W, int stride, uint8_t* out) {
    int64_t npix = (int64_t)H * W;
    if (stride == W * 3 && ((uintptr_t)bgr % 16 == 0) && ((uintptr_t)out % 16 == 0)) {
        int64_t threads = cdiv64(npix, 16);
        k_gray_packed<<<(unsigned)cdiv64(threads, 256), 256, 0, st>>>(bgr, out, npix);
    } else {
        k_gray_strided<<<dim3(cdiv(W, 256), H), 256, 0, st>>>(bgr, stride, out, H, W);
    }
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// A3  cv2.resize(INTER_CUBIC), u8 x1                                                   image_preprocessor.py:125-132
// Coefficient tables are built on the host exactly as OpenCV builds them (double coordinate, float32 Keys cubic
// A=-0.75, rint(c*2048) shorts); the kernel does the int32 horizontal pass and the float32 vertical pass.
// ------------------------------------------------------------------------------------------------------------------
static void cubic_coeffs(float x, float* c) {
    const float A = -0.75f;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}

struct CubicAxis {
    std::vector<int32_t> idx;     // [n][4] clamped source indices
    std::vector<int32_t> ci;      // [n][4] rint(c*2048)
    std::vector<float> cf;        // [n][4] raw float coefficients (T2) or ci * 2^-22 (T1 vertical)
};

static CubicAxis cubic_axis(int src, int dst, bool vertical_t1) {
    CubicAxis a;
    a.idx.resize((size_t)dst * 4);
    a.ci.resize((size_t)dst * 4);
    a.cf.resize((size_t)dst * 4);
    double inv_scale = (double)dst / src;
    double scale = 1.0 / inv_scale;
    const float s22 = 1.f / (2048.f * 2048.f);
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        float c[4];
        cubic_coeffs(f, c);
        for (int k = 0; k < 4; ++k) {
            int j = s - 1 + k;
            j = j < 0 ? 0 : (j > src - 1 ? src - 1 : j);
            a.idx[d * 4 + k] = j;
            int q = (int)lrintf(c[k] * 2048.f);
            a.ci[d * 4 + k] = q;
            a.cf[d * 4 + k] = vertical_t1 ? (float)q * s22 : c[k];
        }
    }
    return a;
}

// mode 0 (T1): H_k = sum_j p*ax (int32); out = rint(S0*b0 + (S1*b1 + (S2*b2 + S3*b3))) unfused float32;
//              the last dst_w % 8 columns use OpenCV's scalar tail (sum S_k*beta_k + 2^21) >> 22.
// mode 1 (T2): float32 real-arithmetic cubic, sequential unfused accumulation (oracle/preprocess_np.py resize_cubic T2)
__global__ void k_resize_cubic(const uint8_t* __restrict__ src, int sH, int sW, uint8_t* __restrict__ dst, int dH,
                               int dW, const int32_t* __restrict__ xidx, const int32_t* __restrict__ xci,
                               const float* __restrict__ xcf, const int32_t* __restrict__ yidx,
                               const int32_t* __restrict__ yci, const float* __restrict__ ycf, int mode) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dW || y >= dH) return;
    int4 xi = __ldg(reinterpret_cast<const int4*>(xidx) + x);
    int4 yi = __ldg(reinterpret_cast<const int4*>(yidx) + y);
    const uint8_t* r0 = src + (int64_t)yi.x * sW;
    const uint8_t* r1 = src + (int64_t)yi.y * sW;
    const uint8_t* r2 = src + (int64_t)yi.z * sW;
    const uint8_t* r3 = src + (int64_t)yi.w * sW;
    uint8_t o;
    if (mode == 0) {
        int4 ax = __ldg(reinterpret_cast<const int4*>(xci) + x);
        int S0 = r0[xi.x] * ax.x + r0[xi.y] * ax.y + r0[xi.z] * ax.z + r0[xi.w] * ax.w;
        int S1 = r1[xi.x] * ax.x + r1[xi.y] * ax.y + r1[xi.z] * ax.z + r1[xi.w] * ax.w;
        int S2 = r2[xi.x] * ax.x + r2[xi.y] * ax.y + r2[xi.z] * ax.z + r2[xi.w] * ax.w;
        int S3 = r3[xi.x] * ax.x + r3[xi.y] * ax.y + r3[xi.z] * ax.z + r3[xi.w] * ax.w;
        if (x < (dW / 8) * 8) {
            float4 b = __ldg(reinterpret_cast<const float4*>(ycf) + y);
            float acc = __fmul_rn((float)S3, b.w);
            acc = __fadd_rn(__fmul_rn((float)S2, b.z), acc);
            acc = __fadd_rn(__fmul_rn((float)S1, b.y), acc);
            acc = __fadd_rn(__fmul_rn((float)S0, b.x), acc);
            int v = __float2int_rn(acc);
            o = (uint8_t)min(max(v, 0), 255);
        } else {
            int4 by = __ldg(reinterpret_cast<const int4*>(yci) + y);
            int v = (S0 * by.x + S1 * by.y + S2 * by.z + S3 * by.w + (1 << 21)) >> 22;
            o = (uint8_t)min(max(v, 0), 255);
        }
    } else {
        float4 cx = __ldg(reinterpret_cast<const float4*>(xcf) + x);
        float4 cy = __ldg(reinterpret_cast<const float4*>(ycf) + y);
        auto hrow = [&](const uint8_t* r) {
            float a = __fmul_rn((float)r[xi.x], cx.x);
            a = __fadd_rn(a, __fmul_rn((float)r[xi.y], cx.y));
            a = __fadd_rn(a, __fmul_rn((float)r[xi.z], cx.z));
            a = __fadd_rn(a, __fmul_rn((float)r[xi.w], cx.w));
            return a;
        };
        float a = __fmul_rn(hrow(r0), cy.x);
        a = __fadd_rn(a, __fmul_rn(hrow(r1), cy.y));
        a = __fadd_rn(a, __fmul_rn(hrow(r2), cy.z));
        a = __fadd_rn(a, __fmul_rn(hrow(r3), cy.w));
        int v = __float2int_rn(a);
        o = (uint8_t)min(max(v, 0), 255);
    }
    dst[(int64_t)y * dW + x] = o;
}

// Tiled variant (the default whenever a 128 x 32 output tile needs at most RS_W x RS_H source pixels, i.e. any scale
// factor >= ~1.4): the source patch is staged in shared memory, the horizontal pass runs ONCE per (source row, output
// column) into an int32 / float plane in shared memory, and the vertical pass reads it as 16-byte vectors, four output
// pixels per thread.  Arithmetic (and therefore every output bit) is the same as k_resize_cubic.
constexpr int RT_W = 128, RT_H = 32, RS_W = 96, RS_H = 28;
__global__ void __launch_bounds__(256) k_resize_cubic_tile(const uint8_t* __restrict__ src, int sH, int sW, uint8_t* __restrict__ dst,
                                                           int dH, int dW, const int32_t* __restrict__ xidx,
                                                           const int32_t* __restrict__ xci, const float* __restrict__ xcf,
                                                           const int32_t* __restrict__ yidx, const int32_t* __restrict__ yci,
                                                           const float* __restrict__ ycf, int mode) {
    __shared__ uint8_t s_src[RS_H][RS_W];
    __shared__ __align__(16) int s_h[RS_H][RT_W];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * RT_W, y0 = blockIdx.y * RT_H;
    const int x1 = min(x0 + RT_W, dW) - 1, y1 = min(y0 + RT_H, dH) - 1;
    const int sx_lo = __ldg(xidx + 4 * x0), sx_hi = __ldg(xidx + 4 * x1 + 3);
    const int sy_lo = __ldg(yidx + 4 * y0), sy_hi = __ldg(yidx + 4 * y1 + 3);
    const int sw = sx_hi - sx_lo + 1, sh = sy_hi - sy_lo + 1;
    for (int i = tid; i < sh * sw; i += 256) {
        int r = i / sw, c = i - r * sw;
        s_src[r][c] = src[(int64_t)(sy_lo + r) * sW + sx_lo + c];
    }
    __syncthreads();
    {
        const int tx = tid & (RT_W - 1), x = x0 + tx;
        if (x <= x1) {
            int4 xi = __ldg(reinterpret_cast<const int4*>(xidx) + x);
            xi.x -= sx_lo; xi.y -= sx_lo; xi.z -= sx_lo; xi.w -= sx_lo;
            if (mode == 0) {
                const int4 ax = __ldg(reinterpret_cast<const int4*>(xci) + x);
                for (int r = tid >> 7; r < sh; r += 2) {
                    const uint8_t* p = s_src[r];
                    s_h[r][tx] = p[xi.x] * ax.x + p[xi.y] * ax.y + p[xi.z] * ax.z + p[xi.w] * ax.w;
                }
            } else {
                const float4 cx = __ldg(reinterpret_cast<const float4*>(xcf) + x);
                for (int r = tid >> 7; r < sh; r += 2) {
                    const uint8_t* p = s_src[r];
                    float a = __fmul_rn((float)p[xi.x], cx.x);
                    a = __fadd_rn(a, __fmul_rn((float)p[xi.y], cx.y));
                    a = __fadd_rn(a, __fmul_rn((float)p[xi.z], cx.z));
                    a = __fadd_rn(a, __fmul_rn((float)p[xi.w], cx.w));
                    s_h[r][tx] = __float_as_int(a);
                }
            }
        }
    }
    __syncthreads();
    const int cg = (tid & 31) * 4;
    const int xvec_end = (dW / 8) * 8;
    for (int ry = tid >> 5; ry < RT_H; ry += 8) {
        const int y = y0 + ry;
        if (y > y1) break;
        int4 yi = __ldg(reinterpret_cast<const int4*>(yidx) + y);
        const int4 S0 = *reinterpret_cast<const int4*>(&s_h[yi.x - sy_lo][cg]);
        const int4 S1 = *reinterpret_cast<const int4*>(&s_h[yi.y - sy_lo][cg]);
        const int4 S2 = *reinterpret_cast<const int4*>(&s_h[yi.z - sy_lo][cg]);
        const int4 S3 = *reinterpret_cast<const int4*>(&s_h[yi.w - sy_lo][cg]);
        const int a0[4] = {S0.x, S0.y, S0.z, S0.w}, a1[4] = {S1.x, S1.y, S1.z, S1.w};
        const int a2[4] = {S2.x, S2.y, S2.z, S2.w}, a3[4] = {S3.x, S3.y, S3.z, S3.w};
        const float4 b = __ldg(reinterpret_cast<const float4*>(ycf) + y);
        const int4 by = __ldg(reinterpret_cast<const int4*>(yci) + y);
        uint32_t packed = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int x = x0 + cg + c;
            int v;
            if (mode == 0) {
                if (x < xvec_end) {
                    float acc = __fmul_rn((float)a3[c], b.w);
                    acc = __fadd_rn(__fmul_rn((float)a2[c], b.z), acc);
                    acc = __fadd_rn(__fmul_rn((float)a1[c], b.y), acc);
                    acc = __fadd_rn(__fmul_rn((float)a0[c], b.x), acc);
                    v = __float2int_rn(acc);
                } else {
                    v = (a0[c] * by.x + a1[c] * by.y + a2[c] * by.z + a3[c] * by.w + (1 << 21)) >> 22;
                }
            } else {
                float a = __fmul_rn(__int_as_float(a0[c]), b.x);
                a = __fadd_rn(a, __fmul_rn(__int_as_float(a1[c]), b.y));
                a = __fadd_rn(a, __fmul_rn(__int_as_float(a2[c]), b.z));
                a = __fadd_rn(a, __fmul_rn(__int_as_float(a3[c]), b.w));
                v = __float2int_rn(a);
            }
            packed |= (uint32_t)min(max(v, 0), 255) << (8 * c);
        }
        uint8_t* out = dst + (int64_t)y * dW + x0 + cg;
        if (x0 + cg + 3 <= x1 && (dW & 3) == 0) {
            *reinterpret_cast<uint32_t*>(out) = packed;
        } else {
            for (int c = 0; c < 4 && x0 + cg + c <= x1; ++c) out[c] = (uint8_t)(packed >> (8 * c));
        }
    }
}

struct CubicTables {
    int32_t *xidx, *xci, *yidx, *yci;
    float *xcf, *ycf;
    bool tiled = false;            // every 128 x 32 output tile fits the shared-memory source patch
    bool fused = false;            // ... and so does every 130 x 34 tile of k_gray_resize_gauss
};

// Coefficient tables live on the device for the lifetime of the handle, one set per (source size, target size, mode):
// a batch of same-size photos builds and uploads them once.
static CubicTables cubic_tables(Handle* h, int sH, int sW, int dH, int dW, int mode) {
    std::lock_guard<std::mutex> lk(h->stat_mu);
    const std::array<int, 5> key = {sH, sW, dH, dW, mode};
    auto it = h->cubic_cache.find(key);
    if (it == h->cubic_cache.end()) {
        CubicAxis ax = cubic_axis(sW, dW, false), ay = cubic_axis(sH, dH, mode == 0);
        size_t nx = (size_t)dW * 4, ny = (size_t)dH * 4;
        bool tiled = true;
        for (int x0 = 0; x0 < dW && tiled; x0 += RT_W)
            tiled = ax.idx[4 * (size_t)(std::min(x0 + RT_W, dW) - 1) + 3] - ax.idx[4 * (size_t)x0] + 1 <= RS_W;
        for (int y0 = 0; y0 < dH && tiled; y0 += RT_H)
            tiled = ay.idx[4 * (size_t)(std::min(y0 + RT_H, dH) - 1) + 3] - ay.idx[4 * (size_t)y0] + 1 <= RS_H;
        bool fused = true;
        for (int x0 = 0; x0 < dW && fused; x0 += 128)
            fused = ax.idx[4 * (size_t)std::min(x0 + 128, dW - 1) + 3] - ax.idx[4 * (size_t)std::max(x0 - 1, 0)] + 1 <= 100;
        for (int y0 = 0; y0 < dH && fused; y0 += 32)
            fused = ay.idx[4 * (size_t)std::min(y0 + 32, dH - 1) + 3] - ay.idx[4 * (size_t)std::max(y0 - 1, 0)] + 1 <= 30;
        std::vector<int32_t> host((nx + ny) * 3);
        memcpy(&host[0], ax.idx.data(), nx * 4);
        memcpy(&host[nx], ax.ci.data(), nx * 4);
        memcpy(&host[2 * nx], ax.cf.data(), nx * 4);
        memcpy(&host[3 * nx], ay.idx.data(), ny * 4);
        memcpy(&host[3 * nx + ny], ay.ci.data(), ny * 4);
        memcpy(&host[3 * nx + 2 * ny], ay.cf.data(), ny * 4);
        if (h->cubic_cache.size() >= 32) {                   // bound the cache: sizes seen long ago are rebuilt on demand
            CUDA_CHECK(cudaDeviceSynchronize());
            for (auto& kv : h->cubic_cache) cudaFree(kv.second.first);
            h->cubic_cache.clear();
        }
        void* dev = nullptr;
        CUDA_CHECK(cudaMalloc(&dev, host.size() * 4));
        CUDA_CHECK(cudaMemcpy(dev, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
        // a pageable-memory cudaMemcpy may return while the DMA is still in flight, and the lanes' streams are non-blocking
        // (no implicit ordering against the legacy stream): make sure the tables have landed before any kernel reads them
        CUDA_CHECK(cudaDeviceSynchronize());
        it = h->cubic_cache.emplace(key, std::make_pair(dev, (tiled ? 1 : 0) | (fused ? 2 : 0))).first;
    }
    const size_t nx = (size_t)dW * 4, ny = (size_t)dH * 4;
    int32_t* b = static_cast<int32_t*>(it->second.first);
    CubicTables t;
    t.xidx = b; t.xci = b + nx; t.xcf = reinterpret_cast<float*>(b + 2 * nx);
    t.yidx = b + 3 * nx; t.yci = b + 3 * nx + ny; t.ycf = reinterpret_cast<float*>(b + 3 * nx + 2 * ny);
    t.tiled = it->second.second & 1;
    t.fused = (it->second.second & 2) != 0;
    return t;
}

void pp_resize_cubic(Handle* h, cudaStream_t st, const uint8_t* src, int sH, int sW, uint8_t* dst, int dH, int dW,
                     int mode) {
    const CubicTables t = cubic_tables(h, sH, sW, dH, dW, mode);
    if (t.tiled) {
        k_resize_cubic_tile<<<dim3(cdiv(dW, RT_W), cdiv(dH, RT_H)), 256, 0, st>>>(src, sH, sW, dst, dH, dW, t.xidx, t.xci, t.xcf,
                                                                                  t.yidx, t.yci, t.ycf, mode);
    } else {
        dim3 blk(64, 4), grd(cdiv(dW, 64), cdiv(dH, 4));
        k_resize_cubic<<<grd, blk, 0, st>>>(src, sH, sW, dst, dH, dW, t.xidx, t.xci, t.xcf, t.yidx, t.yci, t.ycf, mode);
    }
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// A4  cv2.GaussianBlur((3,3), sigma): fixed-point [k0,k1,k0]/256 both ways, BORDER_REFLECT_101, one rounding
//     (v + 2^15) >> 16.  Optionally accumulates the global pixel sum of the OUTPUT (Pillow Contrast's mean).
// ------------------------------------------------------------------------------------------------------------------
// cv2 borderInterpolate(BORDER_REFLECT_101): reflect until inside.  One reflection is the common case; tile loaders at
// the image edge ask for columns far outside small images (their values are never used, but the index must stay legal),
// and CLAHE pads an 8-pixel-wide image by 8.
__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

__global__ void k_gaussian3(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, int k0, int k1,
                            unsigned long long* __restrict__ sum_out) {
    __shared__ uint8_t tile[18][128 + 2 + 2];           // 16 rows + halo; 128 cols + halo (+pad)
    const int bx = blockIdx.x * 128, by = blockIdx.y * 16;
    for (int i = threadIdx.x; i < 18 * 130; i += blockDim.x) {
        int ty = i / 130, tx = i % 130;
        int gy = reflect101(by + ty - 1, H), gx = reflect101(bx + tx - 1, W);
        tile[ty][tx] = src[(int64_t)gy * W + gx];
    }
    __syncthreads();
    unsigned int local = 0;
    for (int i = threadIdx.x; i < 16 * 128; i += blockDim.x) {
        int ty = i / 128, tx = i % 128;
        int gy = by + ty, gx = bx + tx;
        if (gy < H && gx < W) {
            int h0 = tile[ty][tx] * k0 + tile[ty][tx + 1] * k1 + tile[ty][tx + 2] * k0;
            int h1 = tile[ty + 1][tx] * k0 + tile[ty + 1][tx + 1] * k1 + tile[ty + 1][tx + 2] * k0;
            int h2 = tile[ty + 2][tx] * k0 + tile[ty + 2][tx + 1] * k1 + tile[ty + 2][tx + 2] * k0;
            int v = (h0 * k0 + h1 * k1 + h2 * k0 + (1 << 15)) >> 16;
            dst[(int64_t)gy * W + gx] = (uint8_t)v;
            local += v;
        }
    }
    if (sum_out) {
        for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
        __shared__ unsigned int wsum[8];
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long s = 0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += wsum[i];
            atomicAdd(sum_out, s);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// A2 + A3 + A4 fused for the chain: BGR -> gray -> INTER_CUBIC -> 3x3 Gaussian (+ global sum of the output) in one
// kernel.  A block owns 128 x 32 OUTPUT pixels; it resizes the 130 x 34 pixels the Gaussian needs (the halo columns /
// rows are the BORDER_REFLECT_101 images of real columns / rows, so they are resized like any other) from a source
// patch staged (and gray-converted) in shared memory.  Per-pixel arithmetic is that of the three single kernels, so the
// bytes are identical; the gray plane and the resized plane never go to HBM.
// ------------------------------------------------------------------------------------------------------------------
constexpr int FT_W = 128, FT_H = 32, FR_W = FT_W + 2, FR_P = 136, FR_H = FT_H + 2, FS_W = 100, FS_H = 30;

// one resized pixel from its four horizontally-filtered rows (mode 0: OpenCV's float32 vector path / integer tail; mode 1: T2)
__device__ __forceinline__ uint32_t cubic_vertical(int a0, int a1, int a2, int a3, const float4& b, const int4& by, int mode, bool vec) {
    int v;
    if (mode == 0) {
        if (vec) {
            float acc = __fmul_rn((float)a3, b.w);
            acc = __fadd_rn(__fmul_rn((float)a2, b.z), acc);
            acc = __fadd_rn(__fmul_rn((float)a1, b.y), acc);
            acc = __fadd_rn(__fmul_rn((float)a0, b.x), acc);
            v = __float2int_rn(acc);
        } else {
            v = (a0 * by.x + a1 * by.y + a2 * by.z + a3 * by.w + (1 << 21)) >> 22;
        }
    } else {
        float a = __fmul_rn(__int_as_float(a0), b.x);
        a = __fadd_rn(a, __fmul_rn(__int_as_float(a1), b.y));
        a = __fadd_rn(a, __fmul_rn(__int_as_float(a2), b.z));
        a = __fadd_rn(a, __fmul_rn(__int_as_float(a3), b.w));
        v = __float2int_rn(a);
    }
    return (uint32_t)min(max(v, 0), 255);
}

__global__ void __launch_bounds__(256) k_gray_resize_gauss(const uint8_t* __restrict__ src, int channels, int stride, int sH, int sW,
                                                           uint8_t* __restrict__ dst, int dH, int dW,
                                                           const int32_t* __restrict__ xidx, const int32_t* __restrict__ xci,
                                                           const float* __restrict__ xcf, const int32_t* __restrict__ yidx,
                                                           const int32_t* __restrict__ yci, const float* __restrict__ ycf, int mode,
                                                           int k0, int k1, unsigned long long* __restrict__ sum_out) {
    __shared__ uint8_t s_src[FS_H][FS_W];
    __shared__ int4 s_yi[FR_H], s_yq[FR_H];
    __shared__ float4 s_yf[FR_H];
    __shared__ __align__(16) int s_h[FS_H][FR_P];
    __shared__ __align__(16) uint8_t s_r[FR_H][FR_P];
    __shared__ unsigned int wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * FT_W, y0 = blockIdx.y * FT_H;
    const int xa = max(x0 - 1, 0), xb = min(x0 + FT_W, dW - 1), ya = max(y0 - 1, 0), yb = min(y0 + FT_H, dH - 1);
    const int sx_lo = __ldg(xidx + 4 * xa), sx_hi = __ldg(xidx + 4 * xb + 3);
    const int sy_lo = __ldg(yidx + 4 * ya), sy_hi = __ldg(yidx + 4 * yb + 3);
    const int sw = sx_hi - sx_lo + 1, sh = sy_hi - sy_lo + 1;
    // phase 0: source patch (gray-converted on the way in), one warp per row; vertical taps of the 34 tile rows
    for (int r = warp; r < sh; r += 8) {
        const uint8_t* rowp = src + (int64_t)(sy_lo + r) * stride + (int64_t)sx_lo * channels;
        if (channels == 3) {
            for (int c = lane; c < sw; c += 32) {
                const uint8_t* p = rowp + 3 * c;
                s_src[r][c] = (uint8_t)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15);
            }
        } else {
            for (int c = lane; c < sw; c += 32) s_src[r][c] = rowp[c];
        }
    }
    if (tid < FR_H) {
        const int Y = reflect101(min(y0 - 1 + tid, dH), dH);              // rows past the halo are never used
        int4 yi = __ldg(reinterpret_cast<const int4*>(yidx) + Y);
        yi.x -= sy_lo; yi.y -= sy_lo; yi.z -= sy_lo; yi.w -= sy_lo;
        s_yi[tid] = yi;
        s_yq[tid] = __ldg(reinterpret_cast<const int4*>(yci) + Y);
        s_yf[tid] = __ldg(reinterpret_cast<const float4*>(ycf) + Y);
    }
    __syncthreads();
    // phase 1: horizontal pass, thread = one tile column (taps in registers) walking down the patch rows
    for (int c = tid & 127; c < FR_W; c += 128) {
        if (c >= 128 && (tid & 127) >= FR_W - 128) break;
        const int X = reflect101(min(x0 - 1 + c, dW), dW);                // columns past the halo are never used
        int4 xi = __ldg(reinterpret_cast<const int4*>(xidx) + X);
        xi.x -= sx_lo; xi.y -= sx_lo; xi.z -= sx_lo; xi.w -= sx_lo;
        if (mode == 0) {
            const int4 xc = __ldg(reinterpret_cast<const int4*>(xci) + X);
            for (int r = tid >> 7; r < sh; r += 2) {
                const uint8_t* p = s_src[r];
                s_h[r][c] = p[xi.x] * xc.x + p[xi.y] * xc.y + p[xi.z] * xc.z + p[xi.w] * xc.w;
            }
        } else {
            const float4 xc = __ldg(reinterpret_cast<const float4*>(xcf) + X);
            for (int r = tid >> 7; r < sh; r += 2) {
                const uint8_t* p = s_src[r];
                float a = __fmul_rn((float)p[xi.x], xc.x);
                a = __fadd_rn(a, __fmul_rn((float)p[xi.y], xc.y));
                a = __fadd_rn(a, __fmul_rn((float)p[xi.z], xc.z));
                a = __fadd_rn(a, __fmul_rn((float)p[xi.w], xc.w));
                s_h[r][c] = __float_as_int(a);
            }
        }
    }
    __syncthreads();
    // phase 2: vertical pass -> the 34 x 130 resized pixels the Gaussian needs, four per thread
    const int xvec_end = (dW / 8) * 8;
    const bool all_vec = x0 + FT_W + 1 <= xvec_end;                        // no column of this tile is in OpenCV's scalar tail
    for (int it = tid; it < FR_H * (FR_P / 4 - 1); it += 256) {          // 34 rows x 33 groups of four columns (0..131)
        const int ry = it / (FR_P / 4 - 1), gq = it - ry * (FR_P / 4 - 1);
        const int4 yi = s_yi[ry], by = s_yq[ry];
        const float4 b = s_yf[ry];
        const int4 S0 = *reinterpret_cast<const int4*>(&s_h[yi.x][4 * gq]);
        const int4 S1 = *reinterpret_cast<const int4*>(&s_h[yi.y][4 * gq]);
        const int4 S2 = *reinterpret_cast<const int4*>(&s_h[yi.z][4 * gq]);
        const int4 S3 = *reinterpret_cast<const int4*>(&s_h[yi.w][4 * gq]);
        uint32_t packed;
        if (all_vec || mode != 0) {
            packed = cubic_vertical(S0.x, S1.x, S2.x, S3.x, b, by, mode, true) | (cubic_vertical(S0.y, S1.y, S2.y, S3.y, b, by, mode, true) << 8) |
                     (cubic_vertical(S0.z, S1.z, S2.z, S3.z, b, by, mode, true) << 16) | (cubic_vertical(S0.w, S1.w, S2.w, S3.w, b, by, mode, true) << 24);
        } else {
            const int xc0 = x0 - 1 + 4 * gq;
            packed = cubic_vertical(S0.x, S1.x, S2.x, S3.x, b, by, 0, reflect101(min(xc0, dW), dW) < xvec_end) |
                     (cubic_vertical(S0.y, S1.y, S2.y, S3.y, b, by, 0, reflect101(min(xc0 + 1, dW), dW) < xvec_end) << 8) |
                     (cubic_vertical(S0.z, S1.z, S2.z, S3.z, b, by, 0, reflect101(min(xc0 + 2, dW), dW) < xvec_end) << 16) |
                     (cubic_vertical(S0.w, S1.w, S2.w, S3.w, b, by, 0, reflect101(min(xc0 + 3, dW), dW) < xvec_end) << 24);
        }
        reinterpret_cast<uint32_t*>(s_r[ry])[gq] = packed;
    }
    __syncthreads();
    // phase 3: 3x3 Gaussian of the resized tile, four pixels per thread, + the global sum of the output
    unsigned int local = 0;
    const int j = lane;
    const bool st_vec = (dW & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0;
    for (int oy = warp; oy < FT_H; oy += 8) {
        const int y = y0 + oy, x = x0 + 4 * j;
        if (y >= dH || x >= dW) continue;
        int hsum[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const uint32_t w0 = reinterpret_cast<const uint32_t*>(s_r[oy + r])[j], w1 = reinterpret_cast<const uint32_t*>(s_r[oy + r])[j + 1];
            const int p[6] = {(int)(w0 & 255u), (int)((w0 >> 8) & 255u), (int)((w0 >> 16) & 255u), (int)(w0 >> 24),
                              (int)(w1 & 255u), (int)((w1 >> 8) & 255u)};
#pragma unroll
            for (int c = 0; c < 4; ++c) hsum[r][c] = (p[c] + p[c + 2]) * k0 + p[c + 1] * k1;
        }
        uint32_t packed = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int v = ((hsum[0][c] + hsum[2][c]) * k0 + hsum[1][c] * k1 + (1 << 15)) >> 16;
            packed |= (uint32_t)v << (8 * c);
            if (x + c < dW) local += v;
        }
        uint8_t* o = dst + (int64_t)y * dW + x;
        if (st_vec && x + 3 < dW) {
            *reinterpret_cast<uint32_t*>(o) = packed;
        } else {
            for (int c = 0; c < 4 && x + c < dW; ++c) o[c] = (uint8_t)(packed >> (8 * c));
        }
    }
    if (sum_out) {
        for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
        if (lane == 0) wsum[warp] = local;
        __syncthreads();
        if (tid == 0) {
            unsigned long long s = 0;
            for (int i = 0; i < 8; ++i) s += wsum[i];
            atomicAdd(sum_out, s);
        }
    }
}

void gaussian3_kernel_q8(float sigma, int* k0, int* k1) {
    // OpenCV getGaussianKernelBitExact -> 8-bit fixed point; centre tap absorbs the rounding residue.
    // sigma <= 0 with ksize 3 selects cv2's fixed table [0.25, 0.5, 0.25] (enhanced_extractor.py:254 uses it).
    if (!(sigma > 0.f)) { *k0 = 64; *k1 = 128; return; }
    double e = exp(-1.0 / (2.0 * (double)sigma * sigma));
    double norm = 1.0 + 2.0 * e;
    int q0 = (int)lrint(e / norm * 256.0), q1 = (int)lrint(1.0 / norm * 256.0);
    q1 += 256 - (2 * q0 + q1);
    *k0 = q0;
    *k1 = q1;
}

void pp_gaussian3(Handle* h, cudaStream_t st, const uint8_t* src, uint8_t* dst, int H, int W, float sigma,
                  unsigned long long* sum_out) {
    int k0, k1;
    gaussian3_kernel_q8(sigma, &k0, &k1);
    k_gaussian3<<<dim3(cdiv(W, 128), cdiv(H, 16)), 256, 0, st>>>(src, dst, H, W, k0, k1, sum_out);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// gray (channels == 3) -> INTER_CUBIC -> GaussianBlur 3x3, with the global sum of the result; one launch when the
// geometry fits the fused kernel's shared-memory patch (any up-scale), else the three single kernels.
static int pp_gray_resize_gauss(Handle* h, cudaStream_t st, const uint8_t* bgr, int H, int W, int stride, uint8_t* dst, int dH,
                                int dW, int mode, float sigma, unsigned long long* sum_out) {
    const CubicTables t = cubic_tables(h, H, W, dH, dW, mode);
    if (t.fused) {
        int k0, k1;
        gaussian3_kernel_q8(sigma, &k0, &k1);
        k_gray_resize_gauss<<<dim3(cdiv(dW, FT_W), cdiv(dH, FT_H)), 256, 0, st>>>(bgr, 3, stride, H, W, dst, dH, dW, t.xidx, t.xci,
                                                                                  t.xcf, t.yidx, t.yci, t.ycf, mode, k0, k1, sum_out);
        count_launch(h);
        CUDA_CHECK(cudaGetLastError());
        return 1;
    }
    DevBuf gray((size_t)H * W, st), resized((size_t)dH * dW, st);
    pp_gray(h, st, bgr, H, W, stride, gray.as<uint8_t>());
    pp_resize_cubic(h, st, gray.as<uint8_t>(), H, W, resized.as<uint8_t>(), dH, dW, mode);
    pp_gaussian3(h, st, resized.as<uint8_t>(), dst, dH, dW, sigma, sum_out);
    return 3;
}

// ------------------------------------------------------------------------------------------------------------------
// A5/A6  Pillow ImageEnhance.Contrast / Brightness = ImagingBlend extrapolation: a 256-entry byte LUT.
//   contrast: t = m + f*(p - m), m = int(mean + 0.5);  brightness: t = 0 + f*(p - 0);  out = t<=0 ? 0 : t>=255 ? 255 : (u8)t
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t pil_blend_px(int base, int p, float f) {
    float t = __fadd_rn((float)base, __fmul_rn(f, (float)(p - base)));
    if (f >= 0.f && f <= 1.f) return (uint8_t)t;
    return t <= 0.f ? 0 : (t >= 255.f ? 255 : (uint8_t)t);
}

__global__ void k_sum_u8(const uint8_t* __restrict__ src, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long local = 0;
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    for (; i + 16 <= n; i += stride) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) local += __dp4a(w[q], 0x01010101u, 0u);
    }
    if (i < n && i + 16 > n)
        for (int64_t j = i; j < n; ++j) local += src[j];
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, local);
}

// lut = brightness(contrast(p)); either factor may be disabled with f <= 0 (identity)
__global__ void k_build_tone_lut(const unsigned long long* __restrict__ sum, double npix, float contrast,
                                 float brightness, uint8_t* __restrict__ lut) {
    int p = threadIdx.x;
    int v = p;
    if (contrast > 0.f) {
        int m = (int)((double)(*sum) / npix + 0.5);
        v = pil_blend_px(m, v, contrast);
    }
    if (brightness > 0.f) v = pil_blend_px(0, v, brightness);
    lut[p] = (uint8_t)v;
}

__global__ void k_apply_lut(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t n,
                            const uint8_t* __restrict__ lut) {
    __shared__ uint8_t s[256];
    s[threadIdx.x & 255] = lut[threadIdx.x & 255];
    __syncthreads();
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i + 16 <= n) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            w[q] = s[w[q] & 255] | (s[(w[q] >> 8) & 255] << 8) | (s[(w[q] >> 16) & 255] << 16) | (s[w[q] >> 24] << 24);
        *reinterpret_cast<uint4*>(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (int64_t j = i; j < n; ++j) dst[j] = s[src[j]];
    }
}

// ---- cv2.equalizeHist (image_preprocessor.py:39-46): global 256-bin histogram -> LUT -> apply ------------------------------
__global__ void k_hist256(const uint8_t* __restrict__ src, int64_t n, unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[8][256];                      // one private histogram per warp
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    unsigned int* mine = sh[threadIdx.x >> 5];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < n; i += stride) {
        if (i + 16 <= n) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                atomicAdd(&mine[w[q] & 255], 1u);
                atomicAdd(&mine[(w[q] >> 8) & 255], 1u);
                atomicAdd(&mine[(w[q] >> 16) & 255], 1u);
                atomicAdd(&mine[w[q] >> 24], 1u);
            }
        } else {
            for (int64_t j = i; j < n; ++j) atomicAdd(&mine[src[j]], 1u);
        }
    }
    __syncthreads();
    unsigned int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    if (t) atomicAdd(&hist[threadIdx.x], t);
}

// cv::equalizeHist's LUT: i0 = first non-empty bin; scale = 255.f / (total - hist[i0]); lut[i0] = 0;
// lut[i] = saturate_cast<uchar>(sum_{i0 < j <= i} hist[j] * scale) with the int sum converted to float; a single-valued image
// maps to itself (dst.setTo(i0)).
__global__ void k_equalize_lut(const unsigned int* __restrict__ hist, int total, uint8_t* __restrict__ lut) {
    if (threadIdx.x != 0) return;
    int i = 0;
    while (i < 255 && !hist[i]) ++i;
    if ((int)hist[i] == total) {
        for (int v = 0; v < 256; ++v) lut[v] = (uint8_t)i;
        return;
    }
    const float scale = __fdiv_rn(255.f, (float)(total - (int)hist[i]));
    for (int v = 0; v <= i; ++v) lut[v] = 0;
    int sum = 0;
    for (++i; i < 256; ++i) {
        sum += (int)hist[i];
        const int r = __float2int_rn(__fmul_rn((float)sum, scale));
        lut[i] = (uint8_t)min(max(r, 0), 255);
    }
}

void pp_equalize_hist(Handle* h, cudaStream_t st, const uint8_t* src, uint8_t* dst, int H, int W) {
    const int64_t n = (int64_t)H * W;
    DevBuf tmp(256 * 4 + 256, st);
    unsigned int* hist = tmp.as<unsigned int>();
    uint8_t* lut = tmp.as<uint8_t>() + 1024;
    CUDA_CHECK(cudaMemsetAsync(hist, 0, 1024, st));
    const int blocks = (int)std::min<int64_t>(cdiv64(n, 256 * 16), 148 * 8);
    k_hist256<<<blocks, 256, 0, st>>>(src, n, hist);
    k_equalize_lut<<<1, 32, 0, st>>>(hist, (int)n, lut);
    count_launch(h, 2);
    pp_apply_lut(h, st, src, dst, n, lut);
    CUDA_CHECK(cudaGetLastError());
}

void pp_sum(Handle* h, cudaStream_t st, const uint8_t* src, int64_t n, unsigned long long* sum) {
    CUDA_CHECK(cudaMemsetAsync(sum, 0, 8, st));
    int blocks = (int)std::min<int64_t>(cdiv64(n, 256 * 16), 148 * 8);
    k_sum_u8<<<blocks, 256, 0, st>>>(src, n, sum);
    count_launch(h);
}

void pp_tone_lut(Handle* h, cudaStream_t st, const unsigned long long* sum, int64_t npix, float contrast,
                 float brightness, uint8_t* lut) {
    k_build_tone_lut<<<1, 256, 0, st>>>(sum, (double)npix, contrast, brightness, lut);
    count_launch(h);
}

void pp_apply_lut(Handle* h, cudaStream_t st, const uint8_t* src, uint8_t* dst, int64_t n, const uint8_t* lut) {
    k_apply_lut<<<(unsigned)cdiv64(cdiv64(n, 16), 256), 256, 0, st>>>(src, dst, n, lut);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// A7  cv2.createCLAHE(clip, (8,8)).apply                                              image_preprocessor.py:48-56
// ------------------------------------------------------------------------------------------------------------------
// Histogram of every tile of the (reflect-101 padded) image, pixels optionally mapped through a tone LUT first.
// grid = (8 tiles x, 8 tiles y, row chunks); block = 256 threads over a chunk of the tile's rows.
__global__ void __launch_bounds__(256) k_clahe_hist(const uint8_t* __restrict__ src, int H, int W, int tw, int th,
                                                    int rows_per_chunk, const uint8_t* __restrict__ tone,
                                                    unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[8][256];                 // one private histogram per warp: fewer same-address collisions
    __shared__ uint8_t stone[256];
    const int tid = threadIdx.x;
    for (int i = tid; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0;
    stone[tid] = tone ? tone[tid] : (uint8_t)tid;
    __syncthreads();
    unsigned int* my = sh[tid >> 5];
    const int tx = blockIdx.x, ty = blockIdx.y;
    const int y0 = blockIdx.z * rows_per_chunk, y1 = min(th, y0 + rows_per_chunk);
    const int xb = tx * tw, xe = xb + tw;              // tile columns in padded coordinates
    const int xa = xb & ~3;                             // groups of four columns aligned to 4 in the image
    const int ng = (xe - xa + 3) >> 2;
    const bool vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0;
    for (int i = tid; i < (y1 - y0) * ng; i += 256) {
        const int r = i / ng, gq = i - r * ng;
        const int gy = reflect101(ty * th + y0 + r, H);
        const int gx0 = xa + 4 * gq;
        const uint8_t* row = src + (int64_t)gy * W;
        if (vec && gx0 >= xb && gx0 + 3 < xe && gx0 + 3 < W) {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(row + gx0));
            atomicAdd(&my[stone[w & 255u]], 1u);
            atomicAdd(&my[stone[(w >> 8) & 255u]], 1u);
            atomicAdd(&my[stone[(w >> 16) & 255u]], 1u);
            atomicAdd(&my[stone[w >> 24]], 1u);
        } else {
            for (int c = 0; c < 4; ++c) {
                const int x = gx0 + c;
                if (x >= xb && x < xe) atomicAdd(&my[stone[row[reflect101(x, W)]]], 1u);
            }
        }
    }
    __syncthreads();
    unsigned int tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += sh[w][tid];
    if (tot) atomicAdd(&hist[(ty * 8 + tx) * 256 + tid], tot);
}

// One block per tile: clip, redistribute, cumulative sum, scale -> LUT.
__global__ void k_clahe_lut(const unsigned int* __restrict__ hist, int clip, float lut_scale, uint8_t* __restrict__ luts) {
    __shared__ int sh[256];
    __shared__ int red[256];
    const int t = threadIdx.x;
    int v = (int)hist[blockIdx.x * 256 + t];
    if (clip > 0) {
        red[t] = max(v - clip, 0);
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (t < o) red[t] += red[t + o];
            __syncthreads();
        }
        int excess = red[0];
        v = min(v, clip) + excess / 256;
        int r = excess % 256;
        if (r > 0) {
            int step = max(256 / r, 1);
            if (t % step == 0 && t / step < r) v += 1;
        }
    }
    sh[t] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {          // Hillis-Steele inclusive scan
        int add = t >= o ? sh[t - o] : 0;
        __syncthreads();
        sh[t] += add;
        __syncthreads();
    }
    int q = __float2int_rn(__fmul_rn((float)sh[t], lut_scale));
    luts[blockIdx.x * 256 + t] = (uint8_t)min(max(q, 0), 255);
}

struct ClaheGeom {
    int tw, th;
    float inv_tw, inv_th;
};

__device__ __forceinline__ uint8_t clahe_px(const uint8_t* __restrict__ luts, int v, int x, int y, ClaheGeom g) {
    float txf = __fsub_rn(__fmul_rn((float)x, g.inv_tw), 0.5f);
    int tx1 = (int)floorf(txf);
    float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
    int tx2 = min(tx1 + 1, 7);
    tx1 = max(tx1, 0);
    float tyf = __fsub_rn(__fmul_rn((float)y, g.inv_th), 0.5f);
    int ty1 = (int)floorf(tyf);
    float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
    int ty2 = min(ty1 + 1, 7);
    ty1 = max(ty1, 0);
    float a = (float)luts[((ty1 * 8 + tx1) << 8) + v], b = (float)luts[((ty1 * 8 + tx2) << 8) + v];
    float c = (float)luts[((ty2 * 8 + tx1) << 8) + v], d = (float)luts[((ty2 * 8 + tx2) << 8) + v];
    float top = __fadd_rn(__fmul_rn(a, xa1), __fmul_rn(b, xa));
    float bot = __fadd_rn(__fmul_rn(c, xa1), __fmul_rn(d, xa));
    float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
    int q = __float2int_rn(res);
    return (uint8_t)min(max(q, 0), 255);
}

__global__ void k_clahe_apply(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, ClaheGeom g,
                              const uint8_t* __restrict__ tone, const uint8_t* __restrict__ luts_g) {
    extern __shared__ uint8_t smem[];
    uint8_t* luts = smem;                 // 64*256
    uint8_t* stone = smem + 64 * 256;     // 256
    for (int i = threadIdx.x; i < 64 * 256 / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(luts)[i] = __ldg(reinterpret_cast<const uint4*>(luts_g) + i);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) stone[i] = tone ? tone[i] : (uint8_t)i;
    __syncthreads();
    int y = blockIdx.y;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
        int v = stone[src[(int64_t)y * W + x]];
        dst[(int64_t)y * W + x] = clahe_px(luts, v, x, y, g);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// A8  Pillow UnsharpMask(radius=1.0, percent, threshold): GaussianBlur(1.0) = 3 horizontal + 3 vertical passes of the
//     radius-0.25 extended box blur (c*ww + (l+r)*fw + 2^23) >> 24 with per-pass edge replication, then
//     out = |in-blur| > thr ? clip8(in + (in-blur)*percent/100) : in.
// Fused with the CLAHE application (tone LUT -> CLAHE LUT blend -> unsharp) when `luts_g` != nullptr so the chain
// reads its input once and writes its output once.
// ------------------------------------------------------------------------------------------------------------------
#define US_H 3      // halo of the three box passes per direction

// 128 x 32 output pixels per block, four pixels per thread in every phase:
//   0. per-column / per-row CLAHE interpolation terms of the tile (they only depend on x or on y)
//   1. tone LUT + CLAHE blend of the tile and its halo into shared memory: one aligned 32-bit load per four pixels and ONE
//      32-bit table load per pixel -- k_clahe_quadlut packs, for each of the 9 x 9 possible (tile, neighbour tile) pairs
//      per axis, the four tile-LUT entries a pixel value blends (already seen through the tone LUT) into one word
//   2. the three horizontal box passes in registers: 10 input bytes (three 32-bit words) -> 8 -> 6 -> 4 values
//   3. the three vertical passes the same way down 10 rows, then the unsharp combine and one 32-bit store per row
// Per-pass edge replication at the image border = copying the border value outwards before each pass.
// Shared tile column c <-> image x = 128 * blockIdx.x - 4 + c (4-byte aligned); columns 1..3 and 132..134 are the halo.
constexpr int UT_W = 128, UT_H = 32, UA_P = UT_W + 8, UA_H = UT_H + 2 * US_H;

__global__ void k_clahe_quadlut(const uint8_t* __restrict__ luts, const uint8_t* __restrict__ tone, uint32_t* __restrict__ quad) {
    const int qy = blockIdx.x / 9, qx = blockIdx.x % 9, v = threadIdx.x;
    const int ty1 = max(qy - 1, 0), ty2 = min(qy, 7), tx1 = max(qx - 1, 0), tx2 = min(qx, 7);
    const int p = tone ? tone[v] : v;
    quad[blockIdx.x * 256 + v] = (uint32_t)luts[((ty1 * 8 + tx1) << 8) + p] | ((uint32_t)luts[((ty1 * 8 + tx2) << 8) + p] << 8) |
                                 ((uint32_t)luts[((ty2 * 8 + tx1) << 8) + p] << 16) | ((uint32_t)luts[((ty2 * 8 + tx2) << 8) + p] << 24);
}

__device__ __forceinline__ unsigned box_px(unsigned c, unsigned l, unsigned r, unsigned ww, unsigned fw) {
    return (c * ww + (l + r) * fw + (1u << 23)) >> 24;
}

// v[0..9] (tile coordinate of v[i] = base + i) -> three box passes -> v[3..6]
__device__ __forceinline__ void box3_passes(unsigned (&v)[10], int base, int c0, int c1, bool edge, unsigned ww, unsigned fw) {
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const int lo = pass, hi = 9 - pass;               // valid input range of this pass
        if (edge) {
#pragma unroll
            for (int i = 8; i >= 0; --i)
                if (i >= lo && i < hi && base + i < c0) v[i] = v[i + 1];
#pragma unroll
            for (int i = 1; i <= 9; ++i)
                if (i > lo && i <= hi && base + i > c1) v[i] = v[i - 1];
        }
        unsigned prev = v[lo];
#pragma unroll
        for (int i = 1; i <= 8; ++i) {
            if (i > lo && i < hi) {
                unsigned cur = v[i];
                v[i] = box_px(cur, prev, v[i + 1], ww, fw);
                prev = cur;
            }
        }
    }
}

__device__ __forceinline__ uint32_t clahe_blend(uint32_t q4, float xa, float xa1, float ya, float ya1) {
    const float fa = (float)(q4 & 255u), fb = (float)((q4 >> 8) & 255u), fc = (float)((q4 >> 16) & 255u), fd = (float)(q4 >> 24);
    const float top = __fadd_rn(__fmul_rn(fa, xa1), __fmul_rn(fb, xa));
    const float bot = __fadd_rn(__fmul_rn(fc, xa1), __fmul_rn(fd, xa));
    const int q = __float2int_rn(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya)));
    return (uint32_t)min(max(q, 0), 255);
}

__global__ void __launch_bounds__(256) k_unsharp_tile(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                                                      unsigned ww, unsigned fw, int percent, int threshold, ClaheGeom g,
                                                      const uint32_t* __restrict__ quad) {
    __shared__ __align__(16) uint8_t a[UA_H][UA_P];
    __shared__ __align__(16) uint8_t hb[UA_H][UT_W];
    __shared__ __align__(16) float s_xa[UA_P];
    __shared__ __align__(16) int s_xq[UA_P];             // quad-table offset of the column: 256 * qx
    __shared__ float s_ya[UA_H];
    __shared__ int s_yq[UA_H];                            // 256 * 9 * qy
    const int tid = threadIdx.x;
    const int bx = blockIdx.x * UT_W - 4, by = blockIdx.y * UT_H - US_H;
    if (quad) {
        if (tid < UA_P) {
            float txf = __fsub_rn(__fmul_rn((float)(bx + tid), g.inv_tw), 0.5f);
            int t1 = (int)floorf(txf);
            s_xa[tid] = __fsub_rn(txf, (float)t1);
            s_xq[tid] = min(max(t1 + 1, 0), 8) << 8;
        } else if (tid < UA_P + UA_H) {
            const int r = tid - UA_P;
            float tyf = __fsub_rn(__fmul_rn((float)(by + r), g.inv_th), 0.5f);
            int t1 = (int)floorf(tyf);
            s_ya[r] = __fsub_rn(tyf, (float)t1);
            s_yq[r] = min(max(t1 + 1, 0), 8) * (9 * 256);
        }
        __syncthreads();
    }
    const bool vec_ok = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0;
    for (int it = tid; it < UA_H * (UA_P / 4); it += 256) {
        const int ty = it / (UA_P / 4), gq = it - ty * (UA_P / 4);
        const int gy = by + ty, gx = bx + 4 * gq;
        uint32_t w = 0;
        if (gy >= 0 && gy < H && gx + 3 >= 0 && gx < W) {
            const uint8_t* row = src + (int64_t)gy * W;
            if (vec_ok && gx >= 0 && gx + 3 < W) {
                w = __ldg(reinterpret_cast<const uint32_t*>(row + gx));
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (gx + c >= 0 && gx + c < W) w |= (uint32_t)row[gx + c] << (8 * c);
            }
            if (quad) {
                const float4 xa = *reinterpret_cast<const float4*>(&s_xa[4 * gq]);
                const int4 xq = *reinterpret_cast<const int4*>(&s_xq[4 * gq]);
                const float ya = s_ya[ty], ya1 = __fsub_rn(1.0f, ya);
                const uint32_t* qrow = quad + s_yq[ty];
                const uint32_t o0 = clahe_blend(__ldg(qrow + xq.x + (w & 255u)), xa.x, __fsub_rn(1.0f, xa.x), ya, ya1);
                const uint32_t o1 = clahe_blend(__ldg(qrow + xq.y + ((w >> 8) & 255u)), xa.y, __fsub_rn(1.0f, xa.y), ya, ya1);
                const uint32_t o2 = clahe_blend(__ldg(qrow + xq.z + ((w >> 16) & 255u)), xa.z, __fsub_rn(1.0f, xa.z), ya, ya1);
                const uint32_t o3 = clahe_blend(__ldg(qrow + xq.w + (w >> 24)), xa.w, __fsub_rn(1.0f, xa.w), ya, ya1);
                w = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24);
            }
        }
        reinterpret_cast<uint32_t*>(a[ty])[gq] = w;
    }
    __syncthreads();
    const int cx0 = max(0, -bx), cx1 = min(UA_P - 1, W - 1 - bx);
    const int cy0 = max(0, -by), cy1 = min(UA_H - 1, H - 1 - by);
    const bool edge_x = cx0 > 0 || cx1 < UA_P - 1, edge_y = cy0 > 0 || cy1 < UA_H - 1;
    for (int it = tid; it < UA_H * 32; it += 256) {
        const int ty = it >> 5, j = it & 31;
        const uint32_t* row = reinterpret_cast<const uint32_t*>(a[ty]) + j;
        const uint32_t w0 = row[0], w1 = row[1], w2 = row[2];       // tile columns 4j .. 4j+11; the window is 4j+1 .. 4j+10
        unsigned v[10] = {(w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24, w1 & 255u, (w1 >> 8) & 255u, (w1 >> 16) & 255u,
                          w1 >> 24, w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u};
        box3_passes(v, 4 * j + 1, cx0, cx1, edge_x, ww, fw);
        reinterpret_cast<uint32_t*>(hb[ty])[j] = v[3] | (v[4] << 8) | (v[5] << 16) | (v[6] << 24);
    }
    __syncthreads();
    {
        const int j = tid & 31, rg = tid >> 5;
        uint32_t w[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) w[i] = reinterpret_cast<const uint32_t*>(hb[4 * rg + i])[j];
        uint32_t orig[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) orig[k] = reinterpret_cast<const uint32_t*>(a[4 * rg + US_H + k])[1 + j];
        uint32_t out[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            unsigned v[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) v[i] = (w[i] >> (8 * c)) & 255u;
            box3_passes(v, 4 * rg, cy0, cy1, edge_y, ww, fw);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int s = (orig[k] >> (8 * c)) & 255u, d = s - (int)v[3 + k];
                int o = s;
                if (abs(d) > threshold) o = min(max(s + d * percent / 100, 0), 255);
                out[k] |= (uint32_t)o << (8 * c);
            }
        }
        const int gx = blockIdx.x * UT_W + 4 * j;
        const bool st_vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int gy = blockIdx.y * UT_H + 4 * rg + k;
            if (gy >= H || gx >= W) continue;
            uint8_t* o = dst + (int64_t)gy * W + gx;
            if (st_vec && gx + 3 < W) {
                *reinterpret_cast<uint32_t*>(o) = out[k];
            } else {
                for (int c = 0; c < 4 && gx + c < W; ++c) o[c] = (uint8_t)(out[k] >> (8 * c));
            }
        }
    }
}

static void pil_box_params(float radius_sigma, unsigned* ww, unsigned* fw) {
    // Pillow BoxBlur.c: _gaussian_blur_radius (all float) then ImagingLineBoxBlur8 weights
    float sigma2 = radius_sigma * radius_sigma / 3;
    float L = (float)sqrt(12.0 * sigma2 + 1.0);
    float l = (float)floor((L - 1.0) / 2.0);
    float a = (2 * l + 1) * (l * (l + 1) - 3 * sigma2);
    a /= 6 * (sigma2 - (l + 1) * (l + 1));
    float fr = l + a;
    int radius = (int)fr;
    if (radius != 0) fail(BBOCR_E_UNSUPPORTED, "unsharp: only box radius < 1 (UnsharpMask radius=1.0) is implemented");
    *ww = (unsigned)((float)(1 << 24) / (fr * 2 + 1));
    *fw = ((1u << 24) - (radius * 2 + 1) * (*ww)) / 2;
}

static ClaheGeom clahe_geom(int H, int W) {
    // OpenCV pads BOTH dimensions by 8 - (dim % 8) as soon as either is not a multiple of 8 (so a divisible dimension
    // grows by a full 8 when the other one is not divisible)
    int pw = W, ph = H;
    if (W % 8 != 0 || H % 8 != 0) { pw = W + 8 - W % 8; ph = H + 8 - H % 8; }
    ClaheGeom g;
    g.tw = pw / 8;
    g.th = ph / 8;
    g.inv_tw = 1.0f / g.tw;
    g.inv_th = 1.0f / g.th;
    return g;
}

// builds the 64 tile LUTs for `src` seen through the optional tone LUT
void pp_clahe_luts(Handle* h, cudaStream_t st, const uint8_t* src, int H, int W, float clip_limit, const uint8_t* tone,
                   unsigned int* hist /*64*256*/, uint8_t* luts /*64*256*/) {
    ClaheGeom g = clahe_geom(H, W);
    CUDA_CHECK(cudaMemsetAsync(hist, 0, 64 * 256 * 4, st));
    int rows_per_chunk = std::max(1, 8192 / g.tw);
    k_clahe_hist<<<dim3(8, 8, cdiv(g.th, rows_per_chunk)), 256, 0, st>>>(src, H, W, g.tw, g.th, rows_per_chunk, tone, hist);
    int area = g.tw * g.th, clip = 0;
    if (clip_limit > 0.0f) {
        clip = (int)((double)clip_limit * area / 256);     // static_cast<int>(clipLimit_ * tileSizeTotal / histSize), clipLimit_ double
        clip = std::max(clip, 1);
    }
    float lut_scale = 255.0f / (float)area;
    k_clahe_lut<<<64, 256, 0, st>>>(hist, clip, lut_scale, luts);
    count_launch(h, 2);
    CUDA_CHECK(cudaGetLastError());
}

void pp_clahe_apply(Handle* h, cudaStream_t st, const uint8_t* src, uint8_t* dst, int H, int W, const uint8_t* tone,
                    const uint8_t* luts) {
    ClaheGeom g = clahe_geom(H, W);
    k_clahe_apply<<<dim3(cdiv(W, 1024), H), 256, 64 * 256 + 256, st>>>(src, dst, H, W, g, tone, luts);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

void pp_unsharp(Handle* h, cudaStream_t st, const uint8_t* src, uint8_t* dst, int H, int W, int percent, int threshold,
                const uint8_t* tone, const uint8_t* clahe_luts) {
    unsigned ww, fw;
    pil_box_params(1.0f, &ww, &fw);
    ClaheGeom g = clahe_geom(H, W);
    DevBuf quad;
    if (clahe_luts) {
        quad.alloc(81 * 256 * 4, st);
        k_clahe_quadlut<<<81, 256, 0, st>>>(clahe_luts, tone, quad.as<uint32_t>());
        count_launch(h);
    }
    k_unsharp_tile<<<dim3(cdiv(W, UT_W), cdiv(H, UT_H)), 256, 0, st>>>(src, dst, H, W, ww, fw, percent, threshold, g,
                                                                        clahe_luts ? quad.as<uint32_t>() : nullptr);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// A10  cv2.adaptiveThreshold (BORDER_REPLICATE).  GAUSSIAN_C: float32 separable blur, rint -> u8 mean;
//      MEAN_C: exact integer box sum, rint(sum / block^2) in double.
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_at_rows(const uint8_t* __restrict__ src, int H, int W, int block, const float* __restrict__ k,
                          float* __restrict__ outf, int* __restrict__ outi) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* row = src + (int64_t)y * W;
    int r = block / 2;
    if (outf) {
        // generic float row filter with fused multiply-add: s = p0*k0; s = fma(p_i, k_i, s)
        float acc = __fmul_rn((float)row[min(max(x - r, 0), W - 1)], k[0]);
        for (int i = 1; i < block; ++i) {
            int xx = min(max(x - r + i, 0), W - 1);
            acc = __fmaf_rn((float)row[xx], k[i], acc);
        }
        outf[(int64_t)y * W + x] = acc;
    } else {
        int acc = 0;
        for (int i = 0; i < block; ++i) acc += row[min(max(x - r + i, 0), W - 1)];
        outi[(int64_t)y * W + x] = acc;
    }
}

__global__ void k_at_cols(const uint8_t* __restrict__ src, const float* __restrict__ inf, const int* __restrict__ ini,
                          int H, int W, int block, const float* __restrict__ k, int inv, int idelta,
                          uint8_t* __restrict__ dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int r = block / 2, mean;
    if (inf) {
        // symmetric float column filter: s = c*k_c; s = fma(r_{+j} + r_{-j}, k_j, s)
        float acc = __fmul_rn(inf[(int64_t)y * W + x], k[r]);
        for (int j = 1; j <= r; ++j) {
            int ya = min(y + j, H - 1), yb = max(y - j, 0);
            acc = __fmaf_rn(__fadd_rn(inf[(int64_t)ya * W + x], inf[(int64_t)yb * W + x]), k[r + j], acc);
        }
        mean = min(max(__float2int_rn(acc), 0), 255);
    } else {
        int acc = 0;
        for (int i = 0; i < block; ++i) acc += ini[(int64_t)min(max(y - r + i, 0), H - 1) * W + x];
        mean = min(max(__double2int_rn((double)acc * (1.0 / ((double)block * block))), 0), 255);
    }
    int d = (int)src[(int64_t)y * W + x] - mean;
    dst[(int64_t)y * W + x] = inv ? (d <= -idelta ? 255 : 0) : (d > -idelta ? 255 : 0);
}

// Tiled variant for the block sizes this repository uses (11, 31, 35; template parameter so that every tap is a
// compile-time index and the weights are immediate constant-bank operands): a 64 x 64 output tile; the source patch with
// its replicate border is staged in shared memory, the row pass produces the (64 + 2r) x 64 intermediate plane (float
// for GAUSSIAN_C, exact int sums for MEAN_C) four pixels per thread from 32-bit words, and the column pass reads it
// back as 16-byte vectors.  Same per-pixel operation order as k_at_rows / k_at_cols (a leading fma(p, k0, +0) is the
// product p * k0), hence the same bytes.
struct AtWeights { float k[36]; };
constexpr int AT_T = 64;
template <int BLOCK, bool GAUSSIAN>
__global__ void __launch_bounds__(256) k_at_tile(const uint8_t* __restrict__ src, int H, int W, const AtWeights kw, int inv,
                                                 int idelta, uint8_t* __restrict__ dst) {
    constexpr int R = BLOCK / 2, ROWS = AT_T + 2 * R, NW = (BLOCK + 3 + 3) / 4, SP = (AT_T / 4 - 1 + NW) * 4;
    __shared__ __align__(16) uint8_t s_src[ROWS][SP];
    __shared__ __align__(16) int s_mid[ROWS][AT_T];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * AT_T, y0 = blockIdx.y * AT_T;
    for (int i = tid; i < ROWS * SP; i += 256) {
        const int ty = i / SP, tx = i - ty * SP;
        const int gy = min(max(y0 - R + ty, 0), H - 1), gx = min(max(x0 - R + tx, 0), W - 1);
        s_src[ty][tx] = src[(int64_t)gy * W + gx];
    }
    __syncthreads();
    // row pass: thread -> (tile row, four output columns 4g .. 4g+3); output c uses inputs c .. c + BLOCK - 1
    for (int it = tid; it < ROWS * (AT_T / 4); it += 256) {
        const int ty = it / (AT_T / 4), g = it - ty * (AT_T / 4);
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_src[ty]) + g;
        float facc[4] = {0.f, 0.f, 0.f, 0.f};
        int iacc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int wi = 0; wi < NW; ++wi) {
            const uint32_t w = wp[wi];
#pragma unroll
            for (int bsel = 0; bsel < 4; ++bsel) {
                const int i = 4 * wi + bsel;
                const unsigned pv = (w >> (8 * bsel)) & 255u;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int t = i - c;
                    if (t >= 0 && t < BLOCK) {
                        if (GAUSSIAN) facc[c] = __fmaf_rn((float)pv, kw.k[t], facc[c]);
                        else iacc[c] += (int)pv;
                    }
                }
            }
        }
        *reinterpret_cast<int4*>(&s_mid[ty][4 * g]) =
            GAUSSIAN ? make_int4(__float_as_int(facc[0]), __float_as_int(facc[1]), __float_as_int(facc[2]), __float_as_int(facc[3]))
                     : make_int4(iacc[0], iacc[1], iacc[2], iacc[3]);
    }
    __syncthreads();
    const bool st_vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0;
    for (int it = tid; it < AT_T * (AT_T / 4); it += 256) {
        const int oy = it / (AT_T / 4), g = it - oy * (AT_T / 4);
        const int y = y0 + oy, x = x0 + 4 * g;
        if (y >= H || x >= W) continue;
        int mean[4];
        if (GAUSSIAN) {
            const int4 c4 = *reinterpret_cast<const int4*>(&s_mid[oy + R][4 * g]);
            float acc[4] = {__fmul_rn(__int_as_float(c4.x), kw.k[R]), __fmul_rn(__int_as_float(c4.y), kw.k[R]),
                            __fmul_rn(__int_as_float(c4.z), kw.k[R]), __fmul_rn(__int_as_float(c4.w), kw.k[R])};
#pragma unroll
            for (int j = 1; j <= R; ++j) {
                const int4 a4 = *reinterpret_cast<const int4*>(&s_mid[oy + R + j][4 * g]);
                const int4 b4 = *reinterpret_cast<const int4*>(&s_mid[oy + R - j][4 * g]);
                acc[0] = __fmaf_rn(__fadd_rn(__int_as_float(a4.x), __int_as_float(b4.x)), kw.k[R + j], acc[0]);
                acc[1] = __fmaf_rn(__fadd_rn(__int_as_float(a4.y), __int_as_float(b4.y)), kw.k[R + j], acc[1]);
                acc[2] = __fmaf_rn(__fadd_rn(__int_as_float(a4.z), __int_as_float(b4.z)), kw.k[R + j], acc[2]);
                acc[3] = __fmaf_rn(__fadd_rn(__int_as_float(a4.w), __int_as_float(b4.w)), kw.k[R + j], acc[3]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) mean[c] = min(max(__float2int_rn(acc[c]), 0), 255);
        } else {
            int acc[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < BLOCK; ++i) {
                const int4 a4 = *reinterpret_cast<const int4*>(&s_mid[oy + i][4 * g]);
                acc[0] += a4.x; acc[1] += a4.y; acc[2] += a4.z; acc[3] += a4.w;
            }
            const double inv_area = 1.0 / ((double)BLOCK * BLOCK);
#pragma unroll
            for (int c = 0; c < 4; ++c) mean[c] = min(max(__double2int_rn((double)acc[c] * inv_area), 0), 255);
        }
        uint32_t packed = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int d = (int)s_src[oy + R][4 * g + c + R] - mean[c];
            const uint32_t o = inv ? (d <= -idelta ? 255u : 0u) : (d > -idelta ? 255u : 0u);
            packed |= o << (8 * c);
        }
        uint8_t* o = dst + (int64_t)y * W + x;
        if (st_vec && x + 3 < W) {
            *reinterpret_cast<uint32_t*>(o) = packed;
        } else {
            for (int c = 0; c < 4 && x + c < W; ++c) o[c] = (uint8_t)(packed >> (8 * c));
        }
    }
}

template <int BLOCK>
static void at_tile_launch(cudaStream_t st, const uint8_t* src, int H, int W, const AtWeights& kw, bool gaussian, int inv, int idelta,
                           uint8_t* dst) {
    dim3 grd(cdiv(W, AT_T), cdiv(H, AT_T));
    if (gaussian) k_at_tile<BLOCK, true><<<grd, 256, 0, st>>>(src, H, W, kw, inv, idelta, dst);
    else k_at_tile<BLOCK, false><<<grd, 256, 0, st>>>(src, H, W, kw, inv, idelta, dst);
}

void pp_adaptive_threshold(Handle* h, cudaStream_t st, const uint8_t* src, uint8_t* dst, int H, int W, int method,
                           int inv, int block, float delta) {
    ARG_CHECK(block % 2 == 1 && block > 1 && block <= 255, "adaptiveThreshold: block must be odd, 3..255");
    std::vector<float> kf(block, 0.f);
    if (method == 1) {
        double sigma = ((block - 1) * 0.5 - 1) * 0.3 + 0.8;
        std::vector<double> kd(block);
        double s = 0;
        for (int i = 0; i < block; ++i) {
            double x = i - (block - 1) * 0.5;
            kd[i] = exp(-(x * x) / (2.0 * sigma * sigma));
            s += kd[i];
        }
        for (int i = 0; i < block; ++i) kf[i] = (float)(kd[i] / s);
    }
    int idelta = inv ? (int)floor((double)delta) : (int)ceil((double)delta);
    if (block == 11 || block == 31 || block == 35) {
        AtWeights kw;
        memset(&kw, 0, sizeof kw);
        memcpy(kw.k, kf.data(), block * sizeof(float));
        if (block == 11) at_tile_launch<11>(st, src, H, W, kw, method == 1, inv, idelta, dst);
        else if (block == 31) at_tile_launch<31>(st, src, H, W, kw, method == 1, inv, idelta, dst);
        else at_tile_launch<35>(st, src, H, W, kw, method == 1, inv, idelta, dst);
        count_launch(h);
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    DevBuf tmp((size_t)H * W * 4, st), kbuf(256 * 4, st);
    if (method == 1) {
        CUDA_CHECK(cudaMemcpyAsync(kbuf.p, kf.data(), block * 4, cudaMemcpyHostToDevice, st));
        CUDA_CHECK(stream_sync(st));
    }
    dim3 grd(cdiv(W, 256), H);
    k_at_rows<<<grd, 256, 0, st>>>(src, H, W, block, kbuf.as<float>(), method == 1 ? tmp.as<float>() : nullptr,
                                   method == 1 ? nullptr : tmp.as<int>());
    k_at_cols<<<grd, 256, 0, st>>>(src, method == 1 ? tmp.as<float>() : nullptr, method == 1 ? nullptr : tmp.as<int>(), H,
                                   W, block, kbuf.as<float>(), inv, idelta, dst);
    count_launch(h, 2);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// A11  preprocess_for_book_cover minus file I/O, all on one stream; device in, device out.
//   launches: fused gray+resize+gaussian(+sum), tone LUT, CLAHE hist, CLAHE lut, quad table, fused CLAHE-apply+unsharp  (6;
//   8 when the resize is a strong down-scale and the fused kernel's shared-memory patch does not fit)
// ------------------------------------------------------------------------------------------------------------------
int preprocess_launches_per_image() { return 6; }

void preprocess_chain_dev(Handle* h, cudaStream_t st, const uint8_t* bgr, int H, int W, int stride,
                          const bbocr_pp_params& p, uint8_t* out, int* outH, int* outW) {
    ARG_CHECK(H > 0 && W > 0 && p.scale > 0, "preprocess: bad geometry");
    int dH = (int)(H * (double)p.scale), dW = (int)(W * (double)p.scale);     // int(h * scale_factor) in Python doubles
    ARG_CHECK(dH > 0 && dW > 0, "preprocess: empty output");
    int64_t n = (int64_t)dH * dW;
    DevBuf blurred((size_t)n, st);
    DevBuf small(8 + 256 + 64 * 256 * 4 + 64 * 256, st);
    unsigned long long* sum = small.as<unsigned long long>();
    uint8_t* tone = small.as<uint8_t>() + 8;
    unsigned int* hist = reinterpret_cast<unsigned int*>(small.as<uint8_t>() + 8 + 256);
    uint8_t* luts = small.as<uint8_t>() + 8 + 256 + 64 * 256 * 4;
    CUDA_CHECK(cudaMemsetAsync(sum, 0, 8, st));
    pp_gray_resize_gauss(h, st, bgr, H, W, stride, blurred.as<uint8_t>(), dH, dW, p.resize_mode, p.sigma, sum);
    pp_tone_lut(h, st, sum, n, p.contrast, p.brightness, tone);
    pp_clahe_luts(h, st, blurred.as<uint8_t>(), dH, dW, p.clahe_clip, tone, hist, luts);
    pp_unsharp(h, st, blurred.as<uint8_t>(), out, dH, dW, p.sharpen_percent, 3, tone, luts);
    *outH = dH;
    *outW = dW;
}

// ------------------------------------------------------------------------------------------------------------------
// A15  deskew -- NOT in the reference (only promised in prose); defined by this repository as
//   1. ink mask = adaptiveThreshold(GAUSSIAN_C, BINARY_INV, 31, 5)
//   2. for theta_i = -max + 0.1*i degrees: projection profile of the ink pixels of every row / even columns,
//      bin = floor((y-cy)*cos - (x-cx)*sin) + R0 ; score_i = sum(bin_count^2) ; theta = first arg-max
//   3. out(x,y) = bilinear sample (float64, replicate border) of src at (cx + c*dx - s*dy, cy + s*dx + c*dy)
// float64 everywhere with the operation order of oracle/preprocess_np.py::deskew (no FMA contraction).
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_deskew_hist(const uint8_t* __restrict__ fg, int H, int W, const double* __restrict__ cs, int n_ang, int NR,
                              int R0, double cx, double cy, unsigned int* __restrict__ bins) {
    int x = (blockIdx.x * blockDim.x + threadIdx.x) * 2, y = blockIdx.y;      // every row, even columns
    if (x >= W || y >= H || !fg[(int64_t)y * W + x]) return;
    const double dx = (double)x - cx, dy = (double)y - cy;
    for (int i = 0; i < n_ang; ++i) {
        double v = __dsub_rn(__dmul_rn(dy, cs[2 * i]), __dmul_rn(dx, cs[2 * i + 1]));
        int r = (int)floor(v) + R0;
        r = min(max(r, 0), NR - 1);
        atomicAdd(&bins[(int64_t)i * NR + r], 1u);
    }
}

// Banded variant (the default): a block owns DK_ROWS image rows and DK_ANG angles.  The ink pixels of a band land, for one
// angle, in a window of at most DK_ROWS + W |sin| + 4 consecutive bins, so the block counts in shared memory and adds its
// non-empty bins to the global profile once; per ink pixel that is DK_ANG shared atomics instead of n_ang global ones.
// The bin of a pixel is computed exactly as above, so the profiles (integers) are identical.
constexpr int DK_ROWS = 32, DK_ANG = 16;
__global__ void __launch_bounds__(256) k_deskew_hist_band(const uint8_t* __restrict__ fg, int H, int W, const double* __restrict__ cs,
                                                          int n_ang, int NR, int R0, double cx, double cy, int win,
                                                          unsigned int* __restrict__ bins) {
    extern __shared__ unsigned int s_bins[];             // [DK_ANG][win]
    __shared__ int s_lo[DK_ANG];
    __shared__ double s_c[DK_ANG], s_s[DK_ANG];
    const int tid = threadIdx.x;
    const int y0 = blockIdx.x * DK_ROWS, y1 = min(y0 + DK_ROWS, H);
    const int a0 = blockIdx.y * DK_ANG, na = min(DK_ANG, n_ang - a0);
    for (int i = tid; i < DK_ANG * win; i += 256) s_bins[i] = 0;
    if (tid < na) {
        const double c = cs[2 * (a0 + tid)], s = cs[2 * (a0 + tid) + 1];
        s_c[tid] = c;
        s_s[tid] = s;
        // smallest projection over the band: rows y0..y1-1, columns 0..W-1 (c > 0 for |angle| <= 45 degrees)
        const double dy_lo = (double)y0 - cy, dx_far = s >= 0.0 ? (double)(W - 1) - cx : -cx;
        s_lo[tid] = (int)floor(dy_lo * c - dx_far * s) - 2 + R0;
    }
    __syncthreads();
    // a thread takes four consecutive even columns (8 pixels of a row); the projection is monotone along the row, so equal
    // bins are neighbours and are merged before the shared atomic
    const int wq = (W + 7) / 8;
    for (int i = tid; i < (y1 - y0) * wq; i += 256) {
        const int ry = i / wq, xq = i - ry * wq;
        const int y = y0 + ry, x = 8 * xq;
        const uint8_t* row = fg + (int64_t)y * W;
        bool ink[4];
        bool any = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ink[j] = x + 2 * j < W && row[x + 2 * j] != 0;
            any |= ink[j];
        }
        if (!any) continue;
        const double dy = (double)y - cy;
        for (int k = 0; k < na; ++k) {
            const double c = s_c[k], sn = s_s[k], dyc = __dmul_rn(dy, c);
            const int lo = s_lo[k];
            int prev = -1, cnt = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!ink[j]) continue;
                const double dx = (double)(x + 2 * j) - cx;
                int r = (int)floor(__dsub_rn(dyc, __dmul_rn(dx, sn))) + R0;
                r = min(max(r, 0), NR - 1);
                if (r == prev) { ++cnt; continue; }
                if (cnt) {
                    const int idx = prev - lo;
                    if (idx >= 0 && idx < win) atomicAdd(&s_bins[k * win + idx], (unsigned)cnt);
                    else atomicAdd(&bins[(int64_t)(a0 + k) * NR + prev], (unsigned)cnt);
                }
                prev = r;
                cnt = 1;
            }
            if (cnt) {
                const int idx = prev - lo;
                if (idx >= 0 && idx < win) atomicAdd(&s_bins[k * win + idx], (unsigned)cnt);
                else atomicAdd(&bins[(int64_t)(a0 + k) * NR + prev], (unsigned)cnt);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < na * win; i += 256) {
        const unsigned int v = s_bins[i];
        if (v) {
            const int k = i / win, r = s_lo[k] + (i - k * win);
            atomicAdd(&bins[(int64_t)(a0 + k) * NR + r], v);
        }
    }
}

// Run-based variant (the default).  Along a row the bin floor(dy c - dx s) is monotone in x, so for one (row, angle) the ink
// pixels of a bin are those of one interval of columns: with a per-row prefix count of the ink pixels (even columns) a thread
// walks the intervals of its (row, angle) -- next boundary from the real-valued estimate, then corrected with the EXACT
// per-pixel expression so that every pixel is counted in the bin the definition gives it -- and adds prefix differences.
// ~W |sin| intervals per (row, angle) instead of W/2 pixels, and no per-pixel atomics.
__global__ void __launch_bounds__(256) k_ink_prefix(const uint8_t* __restrict__ fg, int H, int W, int We, uint16_t* __restrict__ P) {
    __shared__ int s_part[256];
    const int y = blockIdx.x, tid = threadIdx.x;
    const uint8_t* row = fg + (int64_t)y * W;
    const int per = (We + 255) / 256, e0 = tid * per, e1 = min(e0 + per, We);
    int cnt = 0;
    for (int e = e0; e < e1; ++e) cnt += row[2 * e] != 0;
    s_part[tid] = cnt;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        int add = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += add;
        __syncthreads();
    }
    int run = s_part[tid] - cnt;                          // exclusive prefix of this thread's chunk
    uint16_t* out = P + (int64_t)y * (We + 1);
    for (int e = e0; e < e1; ++e) {
        out[e] = (uint16_t)run;
        run += row[2 * e] != 0;
    }
    if (e1 == We && e0 < We) out[We] = (uint16_t)run;
    if (We == 0 && tid == 0) out[0] = 0;
}

__device__ __forceinline__ int deskew_bin(double dyc, int e, double cx, double sn) {
    const double dx = (double)(2 * e) - cx;
    return (int)floor(__dsub_rn(dyc, __dmul_rn(dx, sn)));
}

__global__ void __launch_bounds__(256) k_deskew_prof(const uint16_t* __restrict__ P, int H, int W, int We,
                                                     const double* __restrict__ cs, int n_ang, int NR, int R0, double cx, double cy,
                                                     int win, unsigned int* __restrict__ bins) {
    extern __shared__ unsigned int s_bins[];             // [8][win]
    __shared__ int s_lo[8];
    const int tid = threadIdx.x, lane = tid & 31, k = tid >> 5;       // warp = one angle, lanes = the 32 rows of the band
    const int y0 = blockIdx.x * 32, a0 = blockIdx.y * 8;
    const int ai = a0 + k, y = y0 + lane;
    const bool live = ai < n_ang;
    for (int i = tid; i < 8 * win; i += 256) s_bins[i] = 0;
    double c = 1.0, sn = 0.0;
    if (live) { c = cs[2 * ai]; sn = cs[2 * ai + 1]; }
    if (lane == 0) {
        const double dy_lo = (double)y0 - cy, dx_far = sn >= 0.0 ? (double)(W - 1) - cx : -cx;
        s_lo[k] = (int)floor(dy_lo * c - dx_far * sn) - 2 + R0;
    }
    __syncthreads();
    if (live && y < H && We > 0) {
        const uint16_t* pr = P + (int64_t)y * (We + 1);
        const double dyc = __dmul_rn((double)y - cy, c);
        const double inv = sn != 0.0 ? 1.0 / sn : 0.0;
        const int lo = s_lo[k];
        int e = 0;
        int r = deskew_bin(dyc, 0, cx, sn);
        unsigned before = pr[0];
        while (e < We) {
            int en;
            if (sn == 0.0) {
                en = We;
            } else {
                // real-valued crossing of the bin edge (r for a falling profile, r + 1 for a rising one), then exact correction
                const double edge = sn > 0.0 ? (double)r : (double)(r + 1);
                const double xe = ((dyc - edge) * inv + cx) * 0.5;
                en = (int)floor(xe) + 1;
                en = min(max(en, e + 1), We);
                while (en > e + 1 && deskew_bin(dyc, en - 1, cx, sn) != r) --en;
                while (en < We && deskew_bin(dyc, en, cx, sn) == r) ++en;
            }
            const unsigned after = pr[en];
            const unsigned cnt = after - before;
            if (cnt) {
                const int rb = min(max(r + R0, 0), NR - 1);
                const int idx = rb - lo;
                if (idx >= 0 && idx < win) atomicAdd(&s_bins[k * win + idx], cnt);
                else atomicAdd(&bins[(int64_t)ai * NR + rb], cnt);
            }
            before = after;
            e = en;
            if (e < We) r = deskew_bin(dyc, e, cx, sn);
        }
    }
    __syncthreads();
    for (int i = tid; i < 8 * win; i += 256) {
        const unsigned int v = s_bins[i];
        if (v) {
            const int kk = i / win;
            atomicAdd(&bins[(int64_t)(a0 + kk) * NR + s_lo[kk] + (i - kk * win)], v);
        }
    }
}

__global__ void k_deskew_score(const unsigned int* __restrict__ bins, int NR, unsigned long long* __restrict__ score) {
    unsigned long long acc = 0;
    for (int r = threadIdx.x; r < NR; r += blockDim.x) {
        unsigned long long c = bins[(int64_t)blockIdx.x * NR + r];
        acc += c * c;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    __shared__ unsigned long long ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += ws[i];
        score[blockIdx.x] = t;
    }
}

__global__ void k_rotate_bilinear(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, double c,
                                  double s, double cx, double cy) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const double dx = (double)x - cx, dy = (double)y - cy;
    double sx = __dadd_rn(__dsub_rn(__dmul_rn(c, dx), __dmul_rn(s, dy)), cx);
    double sy = __dadd_rn(__dadd_rn(__dmul_rn(s, dx), __dmul_rn(c, dy)), cy);
    double fx0 = floor(sx), fy0 = floor(sy);
    double fx = __dsub_rn(sx, fx0), fy = __dsub_rn(sy, fy0);
    int x0 = (int)fx0, y0 = (int)fy0;
    int xa = min(max(x0, 0), W - 1), xb = min(max(x0 + 1, 0), W - 1);
    int ya = min(max(y0, 0), H - 1), yb = min(max(y0 + 1, 0), H - 1);
    double p00 = src[(int64_t)ya * W + xa], p01 = src[(int64_t)ya * W + xb];
    double p10 = src[(int64_t)yb * W + xa], p11 = src[(int64_t)yb * W + xb];
    double gx = __dsub_rn(1.0, fx), gy = __dsub_rn(1.0, fy);
    double top = __dadd_rn(__dmul_rn(gx, p00), __dmul_rn(fx, p01));
    double bot = __dadd_rn(__dmul_rn(gx, p10), __dmul_rn(fx, p11));
    double v = __dadd_rn(__dmul_rn(gy, top), __dmul_rn(fy, bot));
    int q = __double2int_rn(v);
    dst[(int64_t)y * W + x] = (uint8_t)min(max(q, 0), 255);
}

float pp_deskew(Handle* h, cudaStream_t st, const uint8_t* src, uint8_t* dst, int H, int W, float max_deg, int variant,
                std::vector<unsigned long long>* scores_out) {
    ARG_CHECK(max_deg >= 0.f && max_deg <= 45.f, "deskew: max_deg out of range");
    const int half = (int)lrint((double)max_deg / 0.1);
    const int n_ang = 2 * half + 1;
    const int NR = (int)ceil(sqrt((double)H * H + (double)W * W)) + 3, R0 = NR / 2;
    const double cx = (W - 1) * 0.5, cy = (H - 1) * 0.5;
    std::vector<double> cs((size_t)n_ang * 2), deg(n_ang);
    for (int i = 0; i < n_ang; ++i) {
        deg[i] = (double)(i - half) * 0.1;
        double rad = deg[i] * 3.141592653589793 / 180.0;
        cs[2 * i] = cos(rad);
        cs[2 * i + 1] = sin(rad);
    }
    DevBuf fg((size_t)H * W, st), dcs(cs.size() * 8, st), bins((size_t)n_ang * NR * 4, st), score((size_t)n_ang * 8, st);
    pp_adaptive_threshold(h, st, src, fg.as<uint8_t>(), H, W, 1, 1, 31, 5.0f);
    CUDA_CHECK(cudaMemcpyAsync(dcs.p, cs.data(), cs.size() * 8, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemsetAsync(bins.p, 0, (size_t)n_ang * NR * 4, st));
    const int win = DK_ROWS + (int)ceil((double)W * sin((double)max_deg * 3.141592653589793 / 180.0)) + 8;
    const int We = (W + 1) / 2;
    // variant: -1 = fastest that fits (default); 0 = per-pixel global atomics, 1 = banded shared atomics, 2 = run-based (test hook)
    if ((variant < 0 || variant == 2) && (size_t)8 * win * 4 <= 46 * 1024 && We < 65536) {
        DevBuf prefix((size_t)H * (We + 1) * 2, st);
        k_ink_prefix<<<H, 256, 0, st>>>(fg.as<uint8_t>(), H, W, We, prefix.as<uint16_t>());
        k_deskew_prof<<<dim3(cdiv(H, 32), cdiv(n_ang, 8)), 256, (size_t)8 * win * 4, st>>>(
            prefix.as<uint16_t>(), H, W, We, dcs.as<double>(), n_ang, NR, R0, cx, cy, win, bins.as<unsigned int>());
        count_launch(h);
    } else if ((variant < 0 || variant == 1) && (size_t)DK_ANG * win * 4 <= 46 * 1024) {
        k_deskew_hist_band<<<dim3(cdiv(H, DK_ROWS), cdiv(n_ang, DK_ANG)), 256, (size_t)DK_ANG * win * 4, st>>>(
            fg.as<uint8_t>(), H, W, dcs.as<double>(), n_ang, NR, R0, cx, cy, win, bins.as<unsigned int>());
    } else {
        k_deskew_hist<<<dim3(cdiv(cdiv(W, 2), 128), H), 128, 0, st>>>(fg.as<uint8_t>(), H, W, dcs.as<double>(), n_ang, NR, R0,
                                                                           cx, cy, bins.as<unsigned int>());
    }
    k_deskew_score<<<n_ang, 256, 0, st>>>(bins.as<unsigned int>(), NR, score.as<unsigned long long>());
    std::vector<unsigned long long> hs(n_ang);
    CUDA_CHECK(cudaMemcpyAsync(hs.data(), score.p, (size_t)n_ang * 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_sync(st));
    if (scores_out) *scores_out = hs;
    int best = 0;
    for (int i = 1; i < n_ang; ++i)
        if (hs[i] > hs[best]) best = i;
    k_rotate_bilinear<<<dim3(cdiv(W, 256), H), 256, 0, st>>>(src, dst, H, W, cs[2 * best], cs[2 * best + 1], cx, cy);
    count_launch(h, 3);
    CUDA_CHECK(cudaGetLastError());
    return (float)deg[best];
}

}  // namespace bbocr
