// nn.cu -- layer kernels shared by the CRAFT detector and the CRNN recogniser (NHWC activations).
//   * k_conv_generic : CUDA-core FP32-accumulate implicit GEMM (the <=1e-3 parity mode, and the fall-back for the few
//                      layer shapes the tcgen05 kernel in conv_tc.cu does not take: tiny Cin/Cout heads)
//   * k_conv_first   : direct convolution for Cin in {1,3} (CRNN conv0 / CRAFT conv1_1 reading the FP32 canvas)
//   * pooling, bilinear x2 up-sampling, row mean, fused classification tail, BiLSTM recurrence
// Upstream semantics: easyocr/craft.py, easyocr/model/modules.py, easyocr/model/vgg_model.py.
#include "engine.h"

namespace bbocr {

size_t act_elem_size(const Handle* h) { return h->precision == BBOCR_PREC_BF16 ? 2 : 4; }

Act act_alloc(Handle* h, cudaStream_t st, DevBuf& buf, int N, int H, int W, int C, bool force_f32) {
    Act a;
    a.N = N; a.H = H; a.W = W; a.C = C;
    size_t es = force_f32 ? 4 : act_elem_size(h);
    buf.alloc((size_t)a.elems() * es + 256, st);
    a.p = buf.p;
    return a;
}

__device__ __forceinline__ void load4(const float* p, float v[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float v[4]) {
    uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float v[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float v[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
}
// split-precision store: hi = bf16(v), lo = bf16(v - hi)
__device__ __forceinline__ void split_store4(__nv_bfloat16* hi, __nv_bfloat16* lo, const float v[4]) {
    float h4[4], l4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h4[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
        l4[j] = v[j] - h4[j];
    }
    store4(hi, h4);
    store4(lo, l4);
}
__device__ __forceinline__ void split_store4(float*, float*, const float*) {}      // FP32 tensors are never split
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ------------------------------------------------------------------------------------------------------------------
// generic implicit-GEMM convolution, 64x64x16 tiles, 256 threads x (4x4) micro-tiles, FP32 accumulation
// ------------------------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) k_conv_generic(const TI* __restrict__ in1, int C1, const TI* __restrict__ in2,
                                                     int C2, TO* __restrict__ out, int N, int H, int W, int OH, int OW,
                                                     const float* __restrict__ w, const float* __restrict__ scale,
                                                     const float* __restrict__ bias, int cout, int cout_pad, int kh,
                                                     int kw, int pad, int dil, int relu) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t M = (int64_t)N * OH * OW;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int a_row = tid >> 2, a_k = (tid & 3) * 4;
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;
    const int64_t m = m0 + a_row;
    const bool mvalid = m < M;
    int n_img = 0, oy = 0, ox = 0;
    if (mvalid) {
        n_img = (int)(m / ((int64_t)OH * OW));
        int r = (int)(m - (int64_t)n_img * OH * OW);
        oy = r / OW;
        ox = r - oy * OW;
    }
    const int cin = C1 + C2;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < kh * kw; ++tap) {
        const int ky = tap / kw, kx = tap - ky * kw;
        const int iy = oy - pad + ky * dil, ix = ox - pad + kx * dil;
        const bool ok = mvalid && iy >= 0 && iy < H && ix >= 0 && ix < W;
        const int64_t pix = ((int64_t)n_img * H + iy) * W + ix;
        for (int c0 = 0; c0 < cin; c0 += BK) {
            float av[4] = {0.f, 0.f, 0.f, 0.f};
            if (ok) {
                int c = c0 + a_k;
                if (c < C1) load4(in1 + pix * C1 + c, av);
                else load4(in2 + pix * C2 + (c - C1), av);
            }
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + b_n < cout_pad)
                bv = __ldg(reinterpret_cast<const float4*>(w + ((int64_t)tap * cin + c0 + b_k) * cout_pad + n0 + b_n));
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) As[a_k + i][a_row] = av[i];
            *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = bv;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
                float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                b[0] = bb.x; b[1] = bb.y; b[2] = bb.z; b[3] = bb.w;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t mm = m0 + ty * 4 + i;
        if (mm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n < cout) {
                float v = fmaf(acc[i][j], __ldg(scale + n), __ldg(bias + n));
                if (relu) v = fmaxf(v, 0.f);
                st1(out + mm * cout + n, v);
            }
        }
    }
}

static void conv_generic(Handle* h, cudaStream_t st, const ConvW& cw, const Act& in1, const Act& in2, Act& out, int flags) {
    const bool bf = h->precision == BBOCR_PREC_BF16;
    const bool out_f32 = !bf || (flags & CONV_OUT_F32);
    int64_t M = (int64_t)out.N * out.H * out.W;
    dim3 grd((unsigned)cdiv64(M, 64), cdiv(cw.cout, 64));
    int relu = (flags & CONV_RELU) ? 1 : 0;
#define LAUNCH(TI, TO)                                                                                            \
    k_conv_generic<TI, TO><<<grd, 256, 0, st>>>((const TI*)in1.p, in1.C, (const TI*)in2.p, in2.C, (TO*)out.p, out.N, \
                                                in1.H, in1.W, out.H, out.W, cw.w_f32, cw.scale, cw.bias, cw.cout,   \
                                                cw.cout_pad, cw.kh, cw.kw, cw.pad, cw.dil, relu)
    if (!bf) LAUNCH(float, float);
    else if (out_f32) LAUNCH(__nv_bfloat16, float);
    else LAUNCH(__nv_bfloat16, __nv_bfloat16);
#undef LAUNCH
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

thread_local int g_conv_scope = 0;      // 1 while the calling thread is inside the detector (CRAFT) forward pass

void conv_forward(Handle* h, cudaStream_t st, const ConvW& cw, const Act& in1, const Act& in2, Act& out, int flags,
                  Act* pooled, const uint8_t* colmask) {
    ARG_CHECK(in1.C + in2.C == cw.cin, "conv: channel mismatch (%d+%d vs %d)", in1.C, in2.C, cw.cin);
    ARG_CHECK(in1.C % 16 == 0 && in2.C % 16 == 0, "conv: channel segments must be multiples of 16");
    ARG_CHECK(out.C == cw.cout, "conv: output channel mismatch");
    ARG_CHECK(out.H == in1.H + 2 * cw.pad - cw.dil * (cw.kh - 1) && out.W == in1.W + 2 * cw.pad - cw.dil * (cw.kw - 1),
              "conv: output geometry mismatch");
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool timed = h->conv_timing && g_conv_scope == 1;     // dominant-kernel instrumentation: detector convolutions
    if (timed) {
        CUDA_CHECK(cudaEventCreate(&e0));
        CUDA_CHECK(cudaEventCreate(&e1));
        CUDA_CHECK(cudaEventRecord(e0, st));
    }
    const bool tc = h->precision == BBOCR_PREC_BF16 && !h->force_generic_conv && conv_tc_supported(cw, in1, in2);
    ARG_CHECK(!colmask || tc, "conv: a column mask needs the tcgen05 path");
    const bool pool_ok = !pooled || (out.H >= 8 && out.H % 2 == 0 && ((flags & CONV_POOL21) || out.W % 2 == 0));
    if (tc && pool_ok && !colmask && !in1.lo && !out.lo && !(flags & (CONV_OUT_F32 | CONV_POOL21)) && conv_res_supported(cw, in1, in2, out)) {
        conv_res_forward(h, st, cw, in1, in2, out, flags, pooled);      // resident weights + halo patch (low-channel 3x3 layers)
    } else if (tc && pool_ok) {
        conv_tc_forward(h, st, cw, in1, in2, out, flags, pooled, colmask);       // max-pool fused into the epilogue
    } else {
        DevBuf tmp;
        Act full = out;
        const bool split = pooled && pooled->lo;                        // split-precision tensors (bf16x3 detector)
        if (pooled && !full.p) {                                        // caller only wants the pooled tensor
            bool f32 = (flags & CONV_OUT_F32) != 0;
            full = split ? act_alloc_split(h, st, tmp, out.N, out.H, out.W, out.C) : act_alloc(h, st, tmp, out.N, out.H, out.W, out.C, f32);
        }
        ARG_CHECK(!colmask, "conv: column mask with an unfused pool is not supported");
        ARG_CHECK(tc || !in1.lo, "conv: split-precision tensors need the tcgen05 path");
        if (tc) conv_tc_forward(h, st, cw, in1, in2, full, flags, nullptr);
        else conv_generic(h, st, cw, in1, in2, full, flags);
        if (pooled) {
            const int kh = 2, kw = (flags & CONV_POOL22) ? 2 : 1;
            if (split) maxpool_split(h, st, full, *pooled, kh, kw, kh, kw, 0, 0);
            else maxpool(h, st, full, *pooled, kh, kw, kh, kw, 0, 0);
        }
    }
    if (timed) {
        CUDA_CHECK(cudaEventRecord(e1, st));
        std::lock_guard<std::mutex> g(h->stat_mu);
        h->conv_events.emplace_back(e0, e1);
        h->conv_flops += 2.0 * (double)out.N * out.H * out.W * cw.cout * cw.cin * cw.kh * cw.kw;
        h->conv_launches += 1;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// direct convolution for Cin in {1,3}: one thread per output pixel, all COUT channels in registers
// ------------------------------------------------------------------------------------------------------------------
template <int COUT, typename TO>
__global__ void __launch_bounds__(128) k_conv_first(const float* __restrict__ in, int N, int H, int W, int cs, int cin,
                                                   const float* __restrict__ w /*[9][cin][cout_pad]*/,
                                                   int cout_pad, const float* __restrict__ scale,
                                                   const float* __restrict__ bias, TO* __restrict__ out, TO* __restrict__ out_lo,
                                                   int relu) {
    __shared__ float sw[9 * 3 * COUT];
    __shared__ float ss[COUT], sb[COUT];
    for (int i = threadIdx.x; i < 9 * cin * COUT; i += blockDim.x) {
        int t = i / COUT, c = i - t * COUT;
        sw[i] = w[(int64_t)t * cout_pad + c];
    }
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) { ss[i] = scale[i]; sb[i] = bias[i]; }
    __syncthreads();
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t M = (int64_t)N * H * W;
    if (m >= M) return;
    int n_img = (int)(m / ((int64_t)H * W));
    int r = (int)(m - (int64_t)n_img * H * W);
    int oy = r / W, ox = r - oy * W;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
        int iy = oy - 1 + ky;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
            int ix = ox - 1 + kx;
            if (ix < 0 || ix >= W) continue;
            const float* p = in + (((int64_t)n_img * H + iy) * W + ix) * cs;
            for (int ci = 0; ci < cin; ++ci) {
                float v = __ldg(p + ci);
                const float* wr = sw + ((ky * 3 + kx) * cin + ci) * COUT;
#pragma unroll
                for (int c = 0; c < COUT; ++c) acc[c] = fmaf(v, wr[c], acc[c]);
            }
        }
    }
    TO* o = out + m * COUT;
#pragma unroll
    for (int c = 0; c < COUT; c += 4) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = fmaf(acc[c + j], ss[c + j], sb[c + j]);
            if (relu) v[j] = fmaxf(v[j], 0.f);
        }
        if (out_lo) split_store4(o + c, out_lo + m * COUT + c, v);      // split-precision tensor (bf16x3 detector stem)
        else store4(o + c, v);
    }
}

void conv_first(Handle* h, cudaStream_t st, const ConvW& cw, const float* in, int N, int H, int W, int cstride, Act& out,
                int flags) {
    const bool force_f32 = (flags & CONV_OUT_F32) != 0;
    ARG_CHECK(cw.kh == 3 && cw.kw == 3 && cw.pad == 1 && cw.dil == 1 && cw.cin <= 3, "conv_first: unsupported shape");
    ARG_CHECK(cw.cout == 32 || cw.cout == 64, "conv_first: cout must be 32 or 64");
    int64_t M = (int64_t)N * H * W;
    unsigned grd = (unsigned)cdiv64(M, 128);
    int relu = (flags & CONV_RELU) ? 1 : 0;
    const bool bf = h->precision == BBOCR_PREC_BF16 && !force_f32;
    ARG_CHECK(!out.lo || bf, "conv_first: split output needs the bf16 element type");
#define LAUNCH(CO, TO) \
    k_conv_first<CO, TO><<<grd, 128, 0, st>>>(in, N, H, W, cstride, cw.cin, cw.w_f32, cw.cout_pad, cw.scale, cw.bias, (TO*)out.p, (TO*)out.lo, relu)
    if (cw.cout == 64) { if (bf) LAUNCH(64, __nv_bfloat16); else LAUNCH(64, float); }
    else { if (bf) LAUNCH(32, __nv_bfloat16); else LAUNCH(32, float); }
#undef LAUNCH
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// max pooling (-inf padding), 4 channels per thread
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_maxpool(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C, int OH, int OW,
                          int kh, int kw, int sh, int sw, int ph, int pw) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int C4 = C >> 2;
    int64_t total = (int64_t)N * OH * OW * C4;
    if (idx >= total) return;
    int c = (int)(idx % C4) * 4;
    int64_t p = idx / C4;
    int ox = (int)(p % OW);
    p /= OW;
    int oy = (int)(p % OH);
    int n = (int)(p / OH);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int ky = 0; ky < kh; ++ky) {
        int iy = oy * sh - ph + ky;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < kw; ++kx) {
            int ix = ox * sw - pw + kx;
            if (ix < 0 || ix >= W) continue;
            float v[4];
            load4(in + (((int64_t)n * H + iy) * W + ix) * C + c, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) best[j] = fmaxf(best[j], v[j]);
        }
    }
    store4(out + (((int64_t)n * OH + oy) * OW + ox) * C + c, best);
}

void maxpool(Handle* h, cudaStream_t st, const Act& in, Act& out, int kh, int kw, int sh, int sw, int ph, int pw) {
    ARG_CHECK(in.C % 4 == 0 && out.C == in.C, "maxpool: channels");
    int64_t total = (int64_t)out.N * out.H * out.W * (in.C / 4);
    unsigned grd = (unsigned)cdiv64(total, 256);
    if (h->precision == BBOCR_PREC_BF16)
        k_maxpool<<<grd, 256, 0, st>>>((const __nv_bfloat16*)in.p, (__nv_bfloat16*)out.p, in.N, in.H, in.W, in.C, out.H,
                                       out.W, kh, kw, sh, sw, ph, pw);
    else
        k_maxpool<<<grd, 256, 0, st>>>((const float*)in.p, (float*)out.p, in.N, in.H, in.W, in.C, out.H, out.W, kh, kw,
                                       sh, sw, ph, pw);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// F.interpolate(mode='bilinear', align_corners=False) to exactly twice the size (general in/out ratio kept)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_upsample(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C, int OH, int OW,
                           float sy, float sx) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int C4 = C >> 2;
    int64_t total = (int64_t)N * OH * OW * C4;
    if (idx >= total) return;
    int c = (int)(idx % C4) * 4;
    int64_t p = idx / C4;
    int ox = (int)(p % OW);
    p /= OW;
    int oy = (int)(p % OH);
    int n = (int)(p / OH);
    float fy = fmaxf(sy * ((float)oy + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)ox + 0.5f) - 0.5f, 0.f);
    int y0 = (int)fy, x0 = (int)fx;
    int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    float a[4], b[4], cc[4], d[4], o[4];
    const T* base = in + (int64_t)n * H * W * C + c;
    load4(base + ((int64_t)y0 * W + x0) * C, a);
    load4(base + ((int64_t)y0 * W + x1) * C, b);
    load4(base + ((int64_t)y1 * W + x0) * C, cc);
    load4(base + ((int64_t)y1 * W + x1) * C, d);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = hy * (hx * a[j] + lx * b[j]) + ly * (hx * cc[j] + lx * d[j]);
    store4(out + (((int64_t)n * OH + oy) * OW + ox) * C + c, o);
}

// bf16, exact x2: one thread per INPUT pixel and 8 channels produces the 2x2 output quad from the clamped 3x3 input
// neighbourhood (9 x 16-byte loads for 4 x 16-byte stores instead of 16 x 8-byte loads).  Every output is evaluated with the
// same expression and the same operands as k_upsample, so the two kernels are bit-identical.  SPLIT: the tensors are hi/lo
// pairs (bf16x3 detector): values are hi + lo, the result is re-split.
template <bool SPLIT>
__global__ void __launch_bounds__(256) k_upsample2x_bf16(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ in_lo,
                                                         __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_lo, int N,
                                                         int H, int W, int C) {
    const int C8 = C >> 3;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total_x = W * C8;
    if (idx >= total_x) return;
    const int j = idx / C8, c = (idx - j * C8) * 8;
    const int i = blockIdx.y, n = blockIdx.z;
    const int OH = 2 * H, OW = 2 * W;
    const int rows[3] = {max(i - 1, 0), i, min(i + 1, H - 1)}, cols[3] = {max(j - 1, 0), j, min(j + 1, W - 1)};
    float v[3][3][8];
    const size_t base = (size_t)n * H * W * C + c;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const size_t off = base + ((size_t)rows[a] * W + cols[b]) * C;
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + off));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
                v[a][b][2 * k] = __low2float(t);
                v[a][b][2 * k + 1] = __high2float(t);
            }
            if (SPLIT) {
                const uint4 ql = __ldg(reinterpret_cast<const uint4*>(in_lo + off));
                const uint32_t wl[4] = {ql.x, ql.y, ql.z, ql.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&wl[k]);
                    v[a][b][2 * k] += __low2float(t);
                    v[a][b][2 * k + 1] += __high2float(t);
                }
            }
        }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const int oy = 2 * i + dy;
        const float fy = fmaxf(0.5f * ((float)oy + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)fy;
        const float ly = fy - (float)y0, hy = 1.f - ly;
        // (y0, y1) = (i-1, i) for the upper output row of an interior pixel, else (i, min(i+1, H-1)); row slot = y - i + 1
        const bool up = dy == 0 && i >= 1;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int ox = 2 * j + dx;
            const float fx = fmaxf(0.5f * ((float)ox + 0.5f) - 0.5f, 0.f);
            const int x0 = (int)fx;
            const float lx = fx - (float)x0, hx = 1.f - lx;
            const bool left = dx == 0 && j >= 1;
            uint32_t w[4], wl[4];
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                float o[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float a = up ? (left ? v[0][0][k + e] : v[0][1][k + e]) : (left ? v[1][0][k + e] : v[1][1][k + e]);
                    const float b = up ? (left ? v[0][1][k + e] : v[0][2][k + e]) : (left ? v[1][1][k + e] : v[1][2][k + e]);
                    const float cc = up ? (left ? v[1][0][k + e] : v[1][1][k + e]) : (left ? v[2][0][k + e] : v[2][1][k + e]);
                    const float d = up ? (left ? v[1][1][k + e] : v[1][2][k + e]) : (left ? v[2][1][k + e] : v[2][2][k + e]);
                    o[e] = hy * (hx * a + lx * b) + ly * (hx * cc + lx * d);
                }
                const __nv_bfloat162 t = __floats2bfloat162_rn(o[0], o[1]);
                w[k >> 1] = *reinterpret_cast<const uint32_t*>(&t);
                if (SPLIT) {
                    const __nv_bfloat162 tl = __floats2bfloat162_rn(o[0] - __low2float(t), o[1] - __high2float(t));
                    wl[k >> 1] = *reinterpret_cast<const uint32_t*>(&tl);
                }
            }
            const size_t oo = (((size_t)n * OH + oy) * OW + ox) * C + c;
            *reinterpret_cast<uint4*>(out + oo) = make_uint4(w[0], w[1], w[2], w[3]);
            if (SPLIT) *reinterpret_cast<uint4*>(out_lo + oo) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
    }
}

void upsample2x(Handle* h, cudaStream_t st, const Act& in, Act& out) {
    if (h->precision == BBOCR_PREC_BF16 && in.C % 8 == 0 && out.H == 2 * in.H && out.W == 2 * in.W && out.C == in.C) {
        k_upsample2x_bf16<false><<<dim3(cdiv(in.W * (in.C / 8), 256), in.H, in.N), 256, 0, st>>>(
            (const __nv_bfloat16*)in.p, nullptr, (__nv_bfloat16*)out.p, nullptr, in.N, in.H, in.W, in.C);
        count_launch(h);
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    ARG_CHECK(in.C % 4 == 0 && out.C == in.C, "upsample: channels");
    int64_t total = (int64_t)out.N * out.H * out.W * (in.C / 4);
    unsigned grd = (unsigned)cdiv64(total, 256);
    float sy = (float)in.H / (float)out.H, sx = (float)in.W / (float)out.W;
    if (h->precision == BBOCR_PREC_BF16)
        k_upsample<<<grd, 256, 0, st>>>((const __nv_bfloat16*)in.p, (__nv_bfloat16*)out.p, in.N, in.H, in.W, in.C, out.H,
                                        out.W, sy, sx);
    else
        k_upsample<<<grd, 256, 0, st>>>((const float*)in.p, (float*)out.p, in.N, in.H, in.W, in.C, out.H, out.W, sy, sx);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// AdaptiveAvgPool2d((None,1)) on the permuted feature map: mean over the H rows -> [N][1][W][C]
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_mean_rows(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)N * W * C;
    if (idx >= total) return;
    int c = (int)(idx % C);
    int64_t p = idx / C;
    int x = (int)(p % W);
    int n = (int)(p / W);
    float s = 0.f;
    for (int y = 0; y < H; ++y) s += to_f(in[(((int64_t)n * H + y) * W + x) * C + c]);
    st1(out + idx, s / (float)H);
}

void mean_rows(Handle* h, cudaStream_t st, const Act& in, Act& out) {
    int64_t total = (int64_t)in.N * in.W * in.C;
    unsigned grd = (unsigned)cdiv64(total, 256);
    if (h->precision == BBOCR_PREC_BF16)
        k_mean_rows<<<grd, 256, 0, st>>>((const __nv_bfloat16*)in.p, (__nv_bfloat16*)out.p, in.N, in.H, in.W, in.C);
    else
        k_mean_rows<<<grd, 256, 0, st>>>((const float*)in.p, (float*)out.p, in.N, in.H, in.W, in.C);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// conv_cls tail: 1x1 16->16 + ReLU, 1x1 16->2, split into the text and link maps (y.permute(0,2,3,1)[...,0/1])
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_cls_tail(const T* __restrict__ in, int64_t M, const float* __restrict__ w3, int cp3,
                           const float* __restrict__ b3, const float* __restrict__ w4, int cp4,
                           const float* __restrict__ b4, float* __restrict__ text, float* __restrict__ link) {
    __shared__ float s3[16 * 16], sb3[16], s4[16 * 2], sb4[2];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s3[i] = w3[(i / 16) * cp3 + (i % 16)];     // [cin][cout]
    for (int i = threadIdx.x; i < 32; i += blockDim.x) s4[i] = w4[(i / 2) * cp4 + (i % 2)];
    if (threadIdx.x < 16) sb3[threadIdx.x] = b3[threadIdx.x];
    if (threadIdx.x < 2) sb4[threadIdx.x] = b4[threadIdx.x];
    __syncthreads();
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float x[16], y[16];
#pragma unroll
    for (int c = 0; c < 16; c += 4) load4(in + m * 16 + c, x + c);
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) a = fmaf(x[c], s3[c * 16 + o], a);
        y[o] = fmaxf(a + sb3[o], 0.f);
    }
    float t = 0.f, l = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        t = fmaf(y[c], s4[c * 2], t);
        l = fmaf(y[c], s4[c * 2 + 1], l);
    }
    text[m] = t + sb4[0];
    link[m] = l + sb4[1];
}

void cls_tail(Handle* h, cudaStream_t st, const ConvW& c3, const ConvW& c4, const Act& in, float* text, float* link) {
    ARG_CHECK(in.C == 16 && c3.cin == 16 && c3.cout == 16 && c4.cin == 16 && c4.cout == 2, "cls_tail: shapes");
    int64_t M = (int64_t)in.N * in.H * in.W;
    unsigned grd = (unsigned)cdiv64(M, 256);
    // scale is 1 for these bias-only convolutions; bias holds the conv bias
    if (h->precision == BBOCR_PREC_BF16)
        k_cls_tail<<<grd, 256, 0, st>>>((const __nv_bfloat16*)in.p, M, c3.w_f32, c3.cout_pad, c3.bias, c4.w_f32,
                                        c4.cout_pad, c4.bias, text, link);
    else
        k_cls_tail<<<grd, 256, 0, st>>>((const float*)in.p, M, c3.w_f32, c3.cout_pad, c3.bias, c4.w_f32, c4.cout_pad,
                                        c4.bias, text, link);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}


// ---- split-precision (x = bf16 hi + bf16 lo) helpers for the recogniser in throughput mode ------------------------
Act act_alloc_split(Handle* h, cudaStream_t st, DevBuf& buf, int N, int H, int W, int C) {
    Act a;
    a.N = N; a.H = H; a.W = W; a.C = C;
    size_t half = ((size_t)a.elems() * 2 + 255) & ~(size_t)255;
    buf.alloc(half * 2 + 256, st);
    a.p = buf.p;
    a.lo = (uint8_t*)buf.p + half;
    return a;
}

// MaxPool2d on an FP32 tensor, output split
__global__ void k_maxpool_f32_split(const float* __restrict__ in, __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo,
                                    int N, int H, int W, int C, int OH, int OW, int kh, int kw, const uint8_t* __restrict__ colmask) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int C4 = C >> 2;
    int64_t total = (int64_t)N * OH * OW * C4;
    if (idx >= total) return;
    int c = (int)(idx % C4) * 4;
    int64_t p = idx / C4;
    int ox = (int)(p % OW);
    p /= OW;
    int oy = (int)(p % OH);
    int n = (int)(p / OH);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int ky = 0; ky < kh; ++ky)
        for (int kx = 0; kx < kw; ++kx) {
            float v[4];
            load4(in + (((int64_t)n * H + oy * kh + ky) * W + ox * kw + kx) * C + c, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) best[j] = fmaxf(best[j], v[j]);
        }
    if (colmask && !colmask[ox]) best[0] = best[1] = best[2] = best[3] = 0.f;      // gap between concatenated crops
    int64_t o = (((int64_t)n * OH + oy) * OW + ox) * C + c;
    split_store4(ohi + o, olo + o, best);
}

void maxpool_f32_to_split(Handle* h, cudaStream_t st, const Act& in, Act& out, int kh, int kw, const uint8_t* colmask) {
    int64_t total = (int64_t)out.N * out.H * out.W * (in.C / 4);
    k_maxpool_f32_split<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>((const float*)in.p, (__nv_bfloat16*)out.p, (__nv_bfloat16*)out.lo,
                                                                      in.N, in.H, in.W, in.C, out.H, out.W, kh, kw, colmask);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// AdaptiveAvgPool over the H rows of a split tensor -> split [N][1][W][C]
__global__ void k_mean_rows_split(const __nv_bfloat16* __restrict__ ihi, const __nv_bfloat16* __restrict__ ilo,
                                  __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo, int N, int H, int W, int C) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)N * W * C;
    if (idx >= total) return;
    int c = (int)(idx % C);
    int64_t p = idx / C;
    int x = (int)(p % W);
    int n = (int)(p / W);
    float s = 0.f;
    for (int y = 0; y < H; ++y) {
        int64_t i = (((int64_t)n * H + y) * W + x) * C + c;
        s += __bfloat162float(ihi[i]) + __bfloat162float(ilo[i]);
    }
    float v = s / (float)H;
    __nv_bfloat16 hb = __float2bfloat16_rn(v);
    ohi[idx] = hb;
    olo[idx] = __float2bfloat16_rn(v - __bfloat162float(hb));
}

void mean_rows_split(Handle* h, cudaStream_t st, const Act& in, Act& out) {
    int64_t total = (int64_t)in.N * in.W * in.C;
    k_mean_rows_split<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>((const __nv_bfloat16*)in.p, (const __nv_bfloat16*)in.lo,
                                                                    (__nv_bfloat16*)out.p, (__nv_bfloat16*)out.lo, in.N, in.H,
                                                                    in.W, in.C);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// CRNN stem in one pass: Conv(1 -> 32, 3x3, pad 1) + ReLU + MaxPool2d(2,2) -> split (hi/lo bf16) tensor, optional column
// mask (strips).  One thread per pooled pixel: the four convolution outputs of its window are evaluated with exactly the
// FMA order of k_conv_first (ky, kx ascending; a zero-padded tap adds +0), so the result is bit-identical to
// conv_first + maxpool_f32_to_split while the 32-channel FP32 full-resolution tensor never goes to HBM.
__global__ void __launch_bounds__(128) k_conv0_pool_split(const float* __restrict__ in, int N, int H, int W,
                                                          const float* __restrict__ w /*[9][1][cout_pad]*/, int cout_pad,
                                                          const float* __restrict__ scale, const float* __restrict__ bias,
                                                          __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo,
                                                          const uint8_t* __restrict__ colmask) {
    __shared__ float sw[9 * 32];
    __shared__ float ss[32], sb[32];
    for (int i = threadIdx.x; i < 9 * 32; i += blockDim.x) sw[i] = w[(int64_t)(i >> 5) * cout_pad + (i & 31)];
    if (threadIdx.x < 32) { ss[threadIdx.x] = scale[threadIdx.x]; sb[threadIdx.x] = bias[threadIdx.x]; }
    __syncthreads();
    const int OH = H >> 1, OW = W >> 1;
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (int64_t)N * OH * OW) return;
    const int n_img = (int)(m / ((int64_t)OH * OW));
    const int r = (int)(m - (int64_t)n_img * OH * OW);
    const int py = r / OW, px = r - py * OW;
    float v[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int iy = 2 * py - 1 + a, ix = 2 * px - 1 + b;
            v[a][b] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(in + ((int64_t)n_img * H + iy) * W + ix) : 0.f;
        }
    float best[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) best[c] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            float acc[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int iy = 2 * py + dy - 1 + ky, ix = 2 * px + dx - 1 + kx;
                    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;      // k_conv_first skips the padded taps
                    const float x = v[dy + ky][dx + kx];
                    const float* wr = sw + (ky * 3 + kx) * 32;
#pragma unroll
                    for (int c = 0; c < 32; ++c) acc[c] = fmaf(x, wr[c], acc[c]);
                }
#pragma unroll
            for (int c = 0; c < 32; ++c) best[c] = fmaxf(best[c], fmaxf(fmaf(acc[c], ss[c], sb[c]), 0.f));
        }
    if (colmask && !colmask[px]) {
#pragma unroll
        for (int c = 0; c < 32; ++c) best[c] = 0.f;
    }
    const int64_t o = m * 32;
#pragma unroll
    for (int c = 0; c < 32; c += 4) split_store4(ohi + o + c, olo + o + c, best + c);
}

void conv0_pool_split(Handle* h, cudaStream_t st, const ConvW& cw, const float* in, int N, int H, int W, Act& out,
                      const uint8_t* colmask) {
    ARG_CHECK(cw.kh == 3 && cw.kw == 3 && cw.pad == 1 && cw.dil == 1 && cw.cin == 1 && cw.cout == 32 && out.lo, "conv0_pool_split: shape");
    ARG_CHECK(H % 2 == 0 && W % 2 == 0 && out.H == H / 2 && out.W == W / 2 && out.C == 32 && out.N == N, "conv0_pool_split: geometry");
    const int64_t M = (int64_t)N * (H / 2) * (W / 2);
    k_conv0_pool_split<<<(unsigned)cdiv64(M, 128), 128, 0, st>>>(in, N, H, W, cw.w_f32, cw.cout_pad, cw.scale, cw.bias,
                                                                  (__nv_bfloat16*)out.p, (__nv_bfloat16*)out.lo, colmask);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ragged AdaptiveAvgPool over the H rows: crop i owns columns [col0, col0 + T) of the strip, its sequence starts at row0
__global__ void k_mean_rows_split_ragged(const __nv_bfloat16* __restrict__ ihi, const __nv_bfloat16* __restrict__ ilo,
                                         __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo, int H, int W, int C,
                                         const int* __restrict__ meta) {
    const int crop = blockIdx.y;
    const int col0 = meta[3 * crop], T = meta[3 * crop + 1], row0 = meta[3 * crop + 2];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = idx / C, c = idx - t * C;
    if (t >= T) return;
    float s = 0.f;
    for (int y = 0; y < H; ++y) {
        const int64_t i = ((int64_t)y * W + col0 + t) * C + c;
        s += __bfloat162float(ihi[i]) + __bfloat162float(ilo[i]);
    }
    const float v = s / (float)H;
    const __nv_bfloat16 hb = __float2bfloat16_rn(v);
    const int64_t o = (int64_t)(row0 + t) * C + c;
    ohi[o] = hb;
    olo[o] = __float2bfloat16_rn(v - __bfloat162float(hb));
}

void mean_rows_split_ragged(Handle* h, cudaStream_t st, const Act& in, const Act& out, const int* meta_dev, int n_crops, int t_max) {
    if (n_crops == 0) return;
    k_mean_rows_split_ragged<<<dim3(cdiv(t_max * in.C, 256), n_crops), 256, 0, st>>>(
        (const __nv_bfloat16*)in.p, (const __nv_bfloat16*)in.lo, (__nv_bfloat16*)out.p, (__nv_bfloat16*)out.lo, in.H, in.W, in.C, meta_dev);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ---- split-precision glue of the detector (precision mode bf16x3): values are hi + lo, evaluated in FP32 and re-split --
__device__ __forceinline__ void load4_split(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float v[4]) {
    float a[4], b[4];
    load4(hi, a);
    load4(lo, b);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = a[j] + b[j];
}

__global__ void k_maxpool_split(const __nv_bfloat16* __restrict__ ihi, const __nv_bfloat16* __restrict__ ilo,
                                __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo, int N, int H, int W, int C, int OH,
                                int OW, int kh, int kw, int sh, int sw, int ph, int pw) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int C4 = C >> 2;
    int64_t total = (int64_t)N * OH * OW * C4;
    if (idx >= total) return;
    int c = (int)(idx % C4) * 4;
    int64_t p = idx / C4;
    int ox = (int)(p % OW);
    p /= OW;
    int oy = (int)(p % OH);
    int n = (int)(p / OH);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int ky = 0; ky < kh; ++ky) {
        int iy = oy * sh - ph + ky;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < kw; ++kx) {
            int ix = ox * sw - pw + kx;
            if (ix < 0 || ix >= W) continue;
            float v[4];
            const int64_t i = (((int64_t)n * H + iy) * W + ix) * C + c;
            load4_split(ihi + i, ilo + i, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) best[j] = fmaxf(best[j], v[j]);
        }
    }
    const int64_t o = (((int64_t)n * OH + oy) * OW + ox) * C + c;
    split_store4(ohi + o, olo + o, best);
}

void maxpool_split(Handle* h, cudaStream_t st, const Act& in, Act& out, int kh, int kw, int sh, int sw, int ph, int pw) {
    ARG_CHECK(in.C % 4 == 0 && out.C == in.C && in.lo && out.lo, "maxpool_split: channels / split tensors");
    int64_t total = (int64_t)out.N * out.H * out.W * (in.C / 4);
    k_maxpool_split<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>((const __nv_bfloat16*)in.p, (const __nv_bfloat16*)in.lo, (__nv_bfloat16*)out.p,
                                                                  (__nv_bfloat16*)out.lo, in.N, in.H, in.W, in.C, out.H, out.W, kh, kw, sh, sw, ph, pw);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// F.interpolate(bilinear, align_corners=False) x2 on a split tensor: same expression as k_upsample on hi + lo
__global__ void k_upsample_split(const __nv_bfloat16* __restrict__ ihi, const __nv_bfloat16* __restrict__ ilo,
                                 __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo, int N, int H, int W, int C, int OH,
                                 int OW, float sy, float sx) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int C4 = C >> 2;
    int64_t total = (int64_t)N * OH * OW * C4;
    if (idx >= total) return;
    int c = (int)(idx % C4) * 4;
    int64_t p = idx / C4;
    int ox = (int)(p % OW);
    p /= OW;
    int oy = (int)(p % OH);
    int n = (int)(p / OH);
    float fy = fmaxf(sy * ((float)oy + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)ox + 0.5f) - 0.5f, 0.f);
    int y0 = (int)fy, x0 = (int)fx;
    int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    float a[4], b[4], cc[4], d[4], o[4];
    const int64_t base = (int64_t)n * H * W * C + c;
    const int64_t i00 = base + ((int64_t)y0 * W + x0) * C, i01 = base + ((int64_t)y0 * W + x1) * C;
    const int64_t i10 = base + ((int64_t)y1 * W + x0) * C, i11 = base + ((int64_t)y1 * W + x1) * C;
    load4_split(ihi + i00, ilo + i00, a);
    load4_split(ihi + i01, ilo + i01, b);
    load4_split(ihi + i10, ilo + i10, cc);
    load4_split(ihi + i11, ilo + i11, d);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = hy * (hx * a[j] + lx * b[j]) + ly * (hx * cc[j] + lx * d[j]);
    const int64_t oo = (((int64_t)n * OH + oy) * OW + ox) * C + c;
    split_store4(ohi + oo, olo + oo, o);
}

void upsample2x_split(Handle* h, cudaStream_t st, const Act& in, Act& out) {
    ARG_CHECK(in.C % 4 == 0 && out.C == in.C && in.lo && out.lo, "upsample_split: channels / split tensors");
    if (in.C % 8 == 0 && out.H == 2 * in.H && out.W == 2 * in.W) {
        k_upsample2x_bf16<true><<<dim3(cdiv(in.W * (in.C / 8), 256), in.H, in.N), 256, 0, st>>>(
            (const __nv_bfloat16*)in.p, (const __nv_bfloat16*)in.lo, (__nv_bfloat16*)out.p, (__nv_bfloat16*)out.lo, in.N, in.H, in.W, in.C);
        count_launch(h);
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    int64_t total = (int64_t)out.N * out.H * out.W * (in.C / 4);
    float sy = (float)in.H / (float)out.H, sx = (float)in.W / (float)out.W;
    k_upsample_split<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>((const __nv_bfloat16*)in.p, (const __nv_bfloat16*)in.lo, (__nv_bfloat16*)out.p,
                                                                   (__nv_bfloat16*)out.lo, in.N, in.H, in.W, in.C, out.H, out.W, sy, sx);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// conv_cls tail on an FP32 16-channel tensor regardless of the precision mode (bf16x3 detector)
void cls_tail_f32(Handle* h, cudaStream_t st, const ConvW& c3, const ConvW& c4, const Act& in, float* text, float* link) {
    ARG_CHECK(in.C == 16 && c3.cin == 16 && c3.cout == 16 && c4.cin == 16 && c4.cout == 2, "cls_tail: shapes");
    int64_t M = (int64_t)in.N * in.H * in.W;
    k_cls_tail<<<(unsigned)cdiv64(M, 256), 256, 0, st>>>((const float*)in.p, M, c3.w_f32, c3.cout_pad, c3.bias, c4.w_f32, c4.cout_pad, c4.bias,
                                                         text, link);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ---- dtype conversion between host-facing FP32 buffers and the activation dtype of the current precision mode -------
template <typename T>
__global__ void k_from_f32(const float* __restrict__ in, T* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st1(out + i, in[i]);
}
template <typename T>
__global__ void k_to_f32(const T* __restrict__ in, float* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = to_f(in[i]);
}
void act_from_f32(Handle* h, cudaStream_t st, const float* in, void* out, int64_t n) {
    unsigned g = (unsigned)cdiv64(n, 256);
    if (h->precision == BBOCR_PREC_BF16) k_from_f32<<<g, 256, 0, st>>>(in, (__nv_bfloat16*)out, n);
    else k_from_f32<<<g, 256, 0, st>>>(in, (float*)out, n);
    count_launch(h);
}
void act_to_f32(Handle* h, cudaStream_t st, const void* in, float* out, int64_t n) {
    unsigned g = (unsigned)cdiv64(n, 256);
    if (h->precision == BBOCR_PREC_BF16) k_to_f32<<<g, 256, 0, st>>>((const __nv_bfloat16*)in, out, n);
    else k_to_f32<<<g, 256, 0, st>>>((const float*)in, out, n);
    count_launch(h);
}

}  // namespace bbocr
