// recog.cu -- crop / recognise / decode:
//   utils.get_image_list + compute_ratio_and_resize (cv2 INTER_LINEAR u8), four_point_transform (warpPerspective),
//   recognition.AlignCollate + NormalizePAD (Pillow BICUBIC for tall crops, (x/255-0.5)/0.5, replicate right pad),
//   adjust_contrast_grey, vgg_model.Model.forward, softmax / greedy CTC / custom_mean.
#include "engine.h"
#include "resize.cuh"

namespace bbocr {

// ------------------------------------------------------------------------------------------------------------------
// crops: one launch for all boxes of a page.  grid = (x chunks, rows, boxes)
// ------------------------------------------------------------------------------------------------------------------
// cv2.warpPerspective(INTER_LINEAR, BORDER_CONSTANT 0) sample: fixed-point 5-bit sub-pixel, 15-bit weights
__device__ __forceinline__ uint8_t warp_px(const uint8_t* __restrict__ src, int H, int W, const double* __restrict__ M,
                                           int x, int y, int dw, int dh) {
    // OpenCV evaluates X0 at the start of each 32-wide block and adds M[0]*x1 inside it (WarpPerspectiveInvoker), in plain
    // double multiplies and adds: every operation is written with the _rn intrinsics so that nvcc cannot contract a*b + c
    // into an FMA (the contraction moved X by one 1/32-pixel step on ~2 % of the pixels in round 1)
    // block width: BLOCK_SZ = 32; bh0 = min(16, rows); bw0 = min(32 * 32 / bh0, cols)  -> 64-pixel blocks for patches >= 16 rows
    const int bw0 = min(1024 / min(16, dh), dw);
    const int xb = (x / bw0) * bw0, x1 = x - xb;
    const double dxb = (double)xb, dy = (double)y, dx1 = (double)x1;
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], dxb), __dmul_rn(M[1], dy)), M[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], dxb), __dmul_rn(M[4], dy)), M[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], dxb), __dmul_rn(M[7], dy)), M[8]);
    double Wd = __dadd_rn(W0, __dmul_rn(M[6], dx1));
    Wd = Wd ? __ddiv_rn(32.0, Wd) : 0;
    double fX = fmax((double)INT_MIN, fmin((double)INT_MAX, __dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], dx1)), Wd)));
    double fY = fmax((double)INT_MIN, fmin((double)INT_MAX, __dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], dx1)), Wd)));
    int X = __double2int_rn(fX), Y = __double2int_rn(fY);
    int sx = X >> 5, sy = Y >> 5;
    sx = max(-32768, min(32767, sx));
    sy = max(-32768, min(32767, sy));
    int ax = X & 31, ay = Y & 31;
    int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32, w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
    auto px = [&](int yy, int xx) -> int {
        return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (int)src[(int64_t)yy * W + xx] : 0;
    };
    int v = px(sy, sx) * w00 + px(sy, sx + 1) * w01 + px(sy + 1, sx) * w10 + px(sy + 1, sx + 1) * w11;
    return (uint8_t)((v + (1 << 14)) >> 15);
}

// stage A (free boxes only): warpPerspective into a scratch rectangle (desc.w x desc.h at desc.x0 = scratch offset)
__global__ void k_warp(const uint8_t* __restrict__ gray, int H, int W, const CropDesc* __restrict__ descs,
                       const double* __restrict__ mats, uint8_t* __restrict__ scratch) {
    const CropDesc d = descs[blockIdx.z];
    if (d.free_idx < 0) return;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= d.w || y >= d.h) return;
    scratch[(int64_t)d.x0 + (int64_t)y * d.w + x] = warp_px(gray, H, W, mats + d.free_idx * 9, x, y, d.w, d.h);
}

// stage B: compute_ratio_and_resize (cv2.resize INTER_LINEAR) of the source rectangle into the packed crop buffer
__global__ void k_crop_resize(const uint8_t* __restrict__ gray, int W, const uint8_t* __restrict__ scratch,
                              const CropDesc* __restrict__ descs, uint8_t* __restrict__ crops) {
    const CropDesc d = descs[blockIdx.z];
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= d.ow || y >= d.oh) return;
    const uint8_t* src;
    int stride;
    if (d.free_idx >= 0) { src = scratch + d.x0; stride = d.w; }
    else { src = gray + (int64_t)d.y0 * W + d.x0; stride = W; }
    double scale_x = 1.0 / ((double)d.ow / d.w), scale_y = 1.0 / ((double)d.oh / d.h);
    crops[(int64_t)d.off + (int64_t)y * d.ow + x] = bilinear_u8_px(src, d.h, d.w, stride, 1, 0, x, y, scale_x, scale_y);
}

void crops_dev(Handle* h, cudaStream_t st, const uint8_t* gray, int H, int W, const CropDesc* descs_dev, int n,
               const CropDesc* descs_host, const double* warp_dev, uint8_t* scratch, uint8_t* crops) {
    if (n == 0) return;
    int max_ow = 1, max_oh = 1, max_w = 0, max_h = 0;
    for (int i = 0; i < n; ++i) {
        max_ow = std::max(max_ow, descs_host[i].ow);
        max_oh = std::max(max_oh, descs_host[i].oh);
        if (descs_host[i].free_idx >= 0) { max_w = std::max(max_w, descs_host[i].w); max_h = std::max(max_h, descs_host[i].h); }
    }
    if (max_w > 0) {
        k_warp<<<dim3(cdiv(max_w, 128), max_h, n), 128, 0, st>>>(gray, H, W, descs_dev, warp_dev, scratch);
        count_launch(h);
    }
    k_crop_resize<<<dim3(cdiv(max_ow, 128), max_oh, n), 128, 0, st>>>(gray, W, scratch, descs_dev, crops);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// adjust_contrast_grey: per-crop 256-bin histogram on the device; the two percentiles (np.percentile, linear) and
// the ratio are O(256) host arithmetic in float64 exactly as NumPy does them; the remap is a device kernel.
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_crop_hist(const uint8_t* __restrict__ crops, const CropDesc* __restrict__ descs,
                            unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[256];
    const CropDesc d = descs[blockIdx.x];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int n = d.ow * d.oh;
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&sh[crops[(int64_t)d.off + i]], 1u);
    __syncthreads();
    hist[blockIdx.x * 256 + threadIdx.x] = sh[threadIdx.x];
}

// img = clip((int(img) - low + 25) * ratio, 0, 255).astype(u8) in float64 ; apply[i] == 0 leaves the crop unchanged
__global__ void k_crop_contrast(const uint8_t* __restrict__ crops, const CropDesc* __restrict__ descs,
                                const double* __restrict__ low, const double* __restrict__ ratio,
                                const int* __restrict__ apply, uint8_t* __restrict__ out) {
    const CropDesc d = descs[blockIdx.y];
    const int n = d.ow * d.oh;
    const double lo = low[blockIdx.y], r = ratio[blockIdx.y];
    const int ap = apply[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint8_t v = crops[(int64_t)d.off + i];
        if (ap) {
            double t = ((double)(int)v - lo + 25.0) * r;
            t = fmax(0.0, fmin(255.0, t));
            v = (uint8_t)t;
        }
        out[(int64_t)d.off + i] = v;
    }
}

void crop_hist_dev(Handle* h, cudaStream_t st, const uint8_t* crops, const CropDesc* descs_dev, int n, unsigned int* hist) {
    if (n == 0) return;
    k_crop_hist<<<n, 256, 0, st>>>(crops, descs_dev, hist);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

void crop_contrast_dev(Handle* h, cudaStream_t st, const uint8_t* crops, const CropDesc* descs_dev, int n,
                       const double* low, const double* ratio, const int* apply, uint8_t* out) {
    if (n == 0) return;
    k_crop_contrast<<<dim3(8, n), 256, 0, st>>>(crops, descs_dev, low, ratio, apply, out);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// Pillow Image.resize(BICUBIC) for mode L (ImagingResample 8bpc): integer coefficient tables come from the host
// (computed in double exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc), horizontal then vertical pass.
// Only tall crops (h > w) take this path; everything else is an identity resize.
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_pil_pass(const uint8_t* __restrict__ src, int sW, int sH, uint8_t* __restrict__ dst, int dW, int dH,
                           const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, int vertical) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dW || y >= dH) return;
    int o = vertical ? y : x;
    int lo = bounds[o * 2], cnt = bounds[o * 2 + 1];
    const int* k = kk + o * ksize;
    int ss = 1 << 21;
    for (int i = 0; i < cnt; ++i) {
        int v = vertical ? src[(int64_t)(lo + i) * sW + x] : src[(int64_t)y * sW + lo + i];
        ss += v * k[i];
    }
    ss >>= 22;
    dst[(int64_t)y * dW + x] = (uint8_t)min(max(ss, 0), 255);
}

static double pil_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// in1: right/bottom edge of the source box (Resample.c receives the box as single-precision floats; in0 = 0 here)
static int pil_coeffs(int in_size, float in1, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
    double scale = (double)(in1 - 0.0f) / out_size, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    double support = 2.0 * filterscale;
    int ksize = (int)ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> k(ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        double center = 0.0 + (xx + 0.5) * scale, ww = 0.0, ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            double w = pil_bicubic((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
            kk[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << 22)) : (int)(0.5 + k[x] * (1 << 22));
        }
        bounds[xx * 2] = xmin;
        bounds[xx * 2 + 1] = xmax;
    }
    return ksize;
}

// Image.reduce((fx, fy)) for mode L (libImaging/Reduce.c): out = ((count / 2 + sum of the fx x fy block clipped to the image) *
// mult[count]) >> 24; mult = division_UINT32(count, 8) comes from the host (float division); the four possible counts are
// full block, right edge, bottom edge, corner.
__global__ void k_pil_reduce(const uint8_t* __restrict__ src, int sW, int sH, uint8_t* __restrict__ dst, int dW, int dH, int fx, int fy,
                             uint32_t m_full, uint32_t m_right, uint32_t m_bottom, uint32_t m_corner) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dW) return;
    const int x0 = x * fx, y0 = y * fy, x1 = min(x0 + fx, sW), y1 = min(y0 + fy, sH);
    uint32_t ss = 0;
    for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) ss += src[(int64_t)yy * sW + xx];
    const bool pr = x1 - x0 < fx, pb = y1 - y0 < fy;
    const uint32_t cnt = (uint32_t)((x1 - x0) * (y1 - y0));
    const uint32_t mult = pr ? (pb ? m_corner : m_right) : (pb ? m_bottom : m_full);
    dst[(int64_t)y * dW + x] = (uint8_t)(((ss + cnt / 2) * mult) >> 24);
}

static uint32_t pil_division_u32(int divider) {           // Reduce.c::division_UINT32(divider, 8)
    const uint32_t max_dividend = (1u << 8) * (uint32_t)divider;
    const float max_int = (float)(1 << 30) * 4.0f;
    return (uint32_t)(max_int / (float)max_dividend);
}

void pil_reduce_dev(Handle* h, cudaStream_t st, const uint8_t* src, int sH, int sW, int fx, int fy, uint8_t* dst) {
    const int dW = (sW + fx - 1) / fx, dH = (sH + fy - 1) / fy;
    const int rx = sW % fx, ry = sH % fy;
    k_pil_reduce<<<dim3(cdiv(dW, 128), dH), 128, 0, st>>>(src, sW, sH, dst, dW, dH, fx, fy, pil_division_u32(fx * fy),
                                                         pil_division_u32(std::max(rx, 1) * fy), pil_division_u32(fx * std::max(ry, 1)),
                                                         pil_division_u32(std::max(rx, 1) * std::max(ry, 1)));
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// src (sH x sW) -> dst (dH x dW); scratch must hold sH*dW bytes.  box_w / box_h: the source box (0, 0, box_w, box_h) the
// bicubic pass samples (the whole image unless a reduce() pre-pass left a fractional extent).
void pil_resize_bicubic_dev(Handle* h, cudaStream_t st, const uint8_t* src, int sH, int sW, uint8_t* dst, int dH, int dW,
                            uint8_t* scratch, float box_w, float box_h) {
    if (box_w <= 0.f) box_w = (float)sW;
    if (box_h <= 0.f) box_h = (float)sH;
    std::vector<int> bx, kx, by, ky;
    int ksx = pil_coeffs(sW, box_w, dW, bx, kx), ksy = pil_coeffs(sH, box_h, dH, by, ky);
    std::vector<int> all;
    all.insert(all.end(), bx.begin(), bx.end());
    all.insert(all.end(), kx.begin(), kx.end());
    all.insert(all.end(), by.begin(), by.end());
    all.insert(all.end(), ky.begin(), ky.end());
    DevBuf tab(all.size() * 4, st);
    CUDA_CHECK(cudaMemcpyAsync(tab.p, all.data(), all.size() * 4, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(stream_sync(st));
    int* t = tab.as<int>();
    const uint8_t* hsrc = src;
    if (dW != sW || box_w != (float)dW) {                  // Resample.c: need_horizontal = xsize != in.xsize || box[2] != xsize
        k_pil_pass<<<dim3(cdiv(dW, 64), sH), 64, 0, st>>>(src, sW, sH, scratch, dW, sH, t, t + bx.size(), ksx, 0);
        count_launch(h);
        hsrc = scratch;
    }
    if (dH != sH || box_h != (float)dH) {
        k_pil_pass<<<dim3(cdiv(dW, 64), dH), 64, 0, st>>>(hsrc, dW, sH, dst, dW, dH, t + bx.size() + kx.size(),
                                                         t + bx.size() + kx.size() + by.size(), ksy, 1);
        count_launch(h);
    } else {
        CUDA_CHECK(cudaMemcpyAsync(dst, hsrc, (size_t)dW * dH, cudaMemcpyDeviceToDevice, st));
    }
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// NormalizePAD: ToTensor (/255), sub 0.5, div 0.5; right-pad to model_w by replicating the last column.
// aligned crops are 64 rows x resized_w columns at desc.aoff
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_crops_to_input(const uint8_t* __restrict__ aligned, const CropDesc* __restrict__ descs,
                                 float* __restrict__ inputs) {
    const CropDesc d = descs[blockIdx.z];
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= d.model_w) return;
    int sx = min(x, d.resized_w - 1);
    float v = (float)aligned[(int64_t)d.aoff + (int64_t)y * d.resized_w + sx];
    v = __fdiv_rn(v, 255.f);
    v = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
    inputs[(int64_t)d.bucket_off + ((int64_t)d.slot * 64 + y) * d.model_w + x] = v;
}

void crops_to_input_dev(Handle* h, cudaStream_t st, const uint8_t* aligned, const CropDesc* descs_dev, int n, int max_model_w,
                        float* inputs) {
    if (n == 0) return;
    k_crops_to_input<<<dim3(cdiv(max_model_w, 128), 64, n), 128, 0, st>>>(aligned, descs_dev, inputs);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// Ragged batch: every crop is written side by side into ONE strip image [64][Wtot] at column xoff[crop]; the columns
// between crops stay zero (the strip is cleared first), which is exactly the zero padding each crop's convolutions see
// when it is processed alone.
__global__ void k_crops_to_strip(const uint8_t* __restrict__ aligned, const CropDesc* __restrict__ descs, const int* __restrict__ xoff,
                                 int Wtot, float* __restrict__ strip) {
    const CropDesc d = descs[blockIdx.z];
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= d.model_w) return;
    int sx = min(x, d.resized_w - 1);
    float v = (float)aligned[(int64_t)d.aoff + (int64_t)y * d.resized_w + sx];
    v = __fdiv_rn(v, 255.f);
    v = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
    strip[(int64_t)y * Wtot + xoff[blockIdx.z] + x] = v;
}

void crops_to_strip_dev(Handle* h, cudaStream_t st, const uint8_t* aligned, const CropDesc* descs_dev, const int* xoff_dev, int n,
                        int max_model_w, int Wtot, float* strip) {
    if (n == 0) return;
    CUDA_CHECK(cudaMemsetAsync(strip, 0, (size_t)64 * Wtot * 4, st));
    k_crops_to_strip<<<dim3(cdiv(max_model_w, 128), 64, n), 128, 0, st>>>(aligned, descs_dev, xoff_dev, Wtot, strip);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------------
// vgg_model.Model.forward, split at the sequence boundary so that the recurrent half runs ONCE per page over the crops
// of every width bucket:
//   crnn_features_dev : VGG feature extractor + AdaptiveAvgPool of one bucket (N crops of identical model width Wm)
//   crnn_sequence_dev : 2 x (input-projection GEMM -> clustered BiLSTM recurrence -> Linear) + Prediction, all crops
// ------------------------------------------------------------------------------------------------------------------
// Throughput mode keeps the recogniser at FP32-class accuracy on the tensor cores: activations are carried as a pair of
// bf16 tensors (hi + lo) and every convolution / Linear is three bf16 GEMM segments x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
// accumulated in FP32 (conv_tc.cu).  The recogniser is <5 % of the page's FLOPs, so the 3x costs little, and greedy CTC
// strings then match the FP32 oracle (tests/test_gpu_recognizer.py).
bool crnn_split(const Handle* h) {
#ifdef BBOCR_DIAG              // diagnostics build only (make DIAG=1): plain bf16 recogniser for A/B timing
    static const bool plain = getenv("BBOCR_CRNN_BF16") != nullptr;
    if (plain) return false;
#endif
    return h->precision == BBOCR_PREC_BF16 && !h->force_generic_conv;
}

Act crnn_alloc_seq(Handle* h, cudaStream_t st, DevBuf& buf, int rows) {
    if (crnn_split(h)) return act_alloc_split(h, st, buf, 1, 1, rows, 256);
    return act_alloc(h, st, buf, 1, 1, rows, 256);
}

void crnn_features_dev(Handle* h, cudaStream_t st, const float* x, int N, int Wm, const Act& seq, int row0) {
    if (!h->crnn_loaded) fail(BBOCR_E_STATE, "CRNN weights not loaded (bbocr_load_crnn)");
    ARG_CHECK(N > 0 && Wm >= 64 && Wm % 4 == 0, "crnn: bad batch geometry (N=%d, W=%d)", N, Wm);
    const CrnnW& w = h->crnn;
    const Act none;
    const int R = CONV_RELU;
    const bool split = crnn_split(h);
    DevBuf b0, b1;
    auto alloc = [&](DevBuf& buf, int n, int hh, int ww, int c) {
        return split ? act_alloc_split(h, st, buf, n, hh, ww, c) : act_alloc(h, st, buf, n, hh, ww, c);
    };
    auto conv = [&](const ConvW& cw, const Act& a, DevBuf& buf, int flags) {
        Act o = alloc(buf, a.N, a.H + 2 * cw.pad - cw.dil * (cw.kh - 1), a.W + 2 * cw.pad - cw.dil * (cw.kw - 1), cw.cout);
        conv_forward(h, st, cw, a, none, o, flags);
        return o;
    };
    // conv + ReLU + max-pool in one launch (the pool runs in the tcgen05 kernel's epilogue)
    auto conv_pool = [&](const ConvW& cw, const Act& a, DevBuf& buf, int kh, int kw) {
        Act full;
        full.N = a.N; full.H = a.H; full.W = a.W; full.C = cw.cout; full.p = nullptr;
        Act pooled = alloc(buf, a.N, a.H / kh, a.W / kw, cw.cout);
        conv_forward(h, st, cw, a, none, full, R | (kw == 2 ? CONV_POOL22 : CONV_POOL21), &pooled);
        return pooled;
    };
    Act a;
    if (split) {
        a = act_alloc_split(h, st, b1, N, 32, Wm / 2, 32);
        conv0_pool_split(h, st, w.c0, x, N, 64, Wm, a);                      // conv0 + ReLU + 2x2 pool + hi/lo split in one pass
    } else {
        Act c0 = act_alloc(h, st, b0, N, 64, Wm, 32);
        conv_first(h, st, w.c0, x, N, 64, Wm, 1, c0, R);
        a = act_alloc(h, st, b1, N, 32, Wm / 2, 32);
        maxpool(h, st, c0, a, 2, 2, 2, 2, 0, 0);
    }
    a = conv_pool(w.c1, a, b0, 2, 2);    // 16 x Wm/4
    a = conv(w.c2, a, b1, R);
    a = conv_pool(w.c3, a, b0, 2, 1);    // 8 x Wm/4
    a = conv(w.c4, a, b1, R);
    a = conv_pool(w.c5, a, b0, 2, 1);    // 4 x Wm/4
    a = conv(w.c6, a, b1, R);            // 3 x (Wm/4 - 1)
    Act s;
    s.N = N; s.H = 1; s.W = a.W; s.C = 256;
    if (split) {
        s.p = (uint8_t*)seq.p + (size_t)row0 * 256 * 2;
        s.lo = (uint8_t*)seq.lo + (size_t)row0 * 256 * 2;
        mean_rows_split(h, st, a, s);
    } else {
        s.p = (uint8_t*)seq.p + (size_t)row0 * 256 * act_elem_size(h);
        mean_rows(h, st, a, s);
    }
}

// Ragged feature extractor (throughput mode): ONE launch per layer over the strip of all crops instead of one launch
// chain per width bucket.  Output columns that lie between crops are re-zeroed by every layer (column masks in the
// max-pool / tcgen05 epilogues), so each crop keeps seeing zero padding at its own borders and its features are the same
// as when it is run alone.  Widths and offsets are multiples of 16 at the input, hence even at every pooled resolution.
bool crnn_ragged(const Handle* h) {
    static const bool off = getenv("BBOCR_CRNN_BUCKETS") != nullptr;
    return crnn_split(h) && !off;
}

void crnn_features_strip_dev(Handle* h, cudaStream_t st, const float* strip, int Wtot, const uint8_t* mask1, const uint8_t* mask2,
                             const int* meta_dev, int n_crops, int t_max, const Act& seq) {
    if (!h->crnn_loaded) fail(BBOCR_E_STATE, "CRNN weights not loaded (bbocr_load_crnn)");
    ARG_CHECK(crnn_split(h) && Wtot % 16 == 0, "crnn: the ragged path needs split precision and 16-aligned strips");
    const CrnnW& w = h->crnn;
    const Act none;
    const int R = CONV_RELU;
    DevBuf b0, b1;
    auto conv = [&](const ConvW& cw, const Act& a, DevBuf& buf, int flags, const uint8_t* mask) {
        Act o = act_alloc_split(h, st, buf, a.N, a.H + 2 * cw.pad - cw.dil * (cw.kh - 1), a.W + 2 * cw.pad - cw.dil * (cw.kw - 1), cw.cout);
        conv_forward(h, st, cw, a, none, o, flags, nullptr, mask);
        return o;
    };
    auto conv_pool = [&](const ConvW& cw, const Act& a, DevBuf& buf, int kh, int kw, const uint8_t* mask) {
        Act full;
        full.N = a.N; full.H = a.H; full.W = a.W; full.C = cw.cout; full.p = nullptr;
        Act pooled = act_alloc_split(h, st, buf, a.N, a.H / kh, a.W / kw, cw.cout);
        conv_forward(h, st, cw, a, none, full, R | (kw == 2 ? CONV_POOL22 : CONV_POOL21), &pooled, mask);
        return pooled;
    };
    Act a = act_alloc_split(h, st, b1, 1, 32, Wtot / 2, 32);
    conv0_pool_split(h, st, w.c0, strip, 1, 64, Wtot, a, mask1);            // conv0 + ReLU + 2x2 pool + hi/lo split + gap mask
    a = conv_pool(w.c1, a, b0, 2, 2, mask1);        // 16 x Wtot/4   (mask at the conv's own resolution, before the pool)
    a = conv(w.c2, a, b1, R, mask2);
    a = conv_pool(w.c3, a, b0, 2, 1, mask2);        // 8 x Wtot/4
    a = conv(w.c4, a, b1, R, mask2);
    a = conv_pool(w.c5, a, b0, 2, 1, mask2);        // 4 x Wtot/4
    a = conv(w.c6, a, b1, R, nullptr);              // 3 x (Wtot/4 - 1); the columns that straddle two crops are never read
    mean_rows_split_ragged(h, st, a, seq, meta_dev, n_crops, t_max);
}

void crnn_sequence_dev(Handle* h, Lane& lane, const Act& seq, const std::vector<SeqDesc>& seqs, float* logits) {
    cudaStream_t st = lane.stream;
    const CrnnW& w = h->crnn;
    const Act none;
    const int rows = seq.W;
    const int n_seq = (int)seqs.size();
    if (rows == 0 || n_seq == 0) return;
    const bool split = crnn_split(h);
    // groups of NB sequences of similar length (longest first) -> one cluster each per direction
    const int NBg = lstm_group_size(h);
    std::vector<int> order(n_seq);
    for (int i = 0; i < n_seq; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return seqs[a].T > seqs[b].T; });
    const int n_groups = cdiv(n_seq, NBg);
    std::vector<int> meta((size_t)n_seq * 2 + (size_t)n_groups * NBg, -1);
    for (int i = 0; i < n_seq; ++i) { meta[2 * i] = seqs[i].row0; meta[2 * i + 1] = seqs[i].T; }
    for (int i = 0; i < n_seq; ++i) meta[(size_t)n_seq * 2 + i] = order[i];
    DevBuf dmeta(meta.size() * 4, st);
    {
        if (lane.in_busy) { CUDA_CHECK(stream_sync(st)); lane.in_busy = false; }
        void* pin = lane.pin_in.get(meta.size() * 4);
        memcpy(pin, meta.data(), meta.size() * 4);
        CUDA_CHECK(cudaMemcpyAsync(dmeta.p, pin, meta.size() * 4, cudaMemcpyHostToDevice, st));
        lane.in_busy = true;
    }
    const SeqDesc* seqs_dev = dmeta.as<SeqDesc>();
    const int* groups_dev = dmeta.as<int>() + (size_t)n_seq * 2;

    DevBuf bg, bh, bs[2];
    Act cur = seq;
    for (int layer = 0; layer < 2; ++layer) {
        const LstmW& l = layer == 0 ? w.l0 : w.l1;
        Act gates = act_alloc(h, st, bg, 1, 1, rows, 2048, true);               // FP32 input projections, all time steps
        conv_forward(h, st, l.in_proj, cur, none, gates, CONV_OUT_F32);
        Act hcat = split ? act_alloc_split(h, st, bh, 1, 1, rows, 512) : act_alloc(h, st, bh, 1, 1, rows, 512);
        lstm_sequences(h, lane, (const float*)gates.p, l.w_hh, seqs.data(), n_seq, seqs_dev, groups_dev, n_groups, hcat.p, hcat.lo);
        Act nxt = split ? act_alloc_split(h, st, bs[layer], 1, 1, rows, 256) : act_alloc(h, st, bs[layer], 1, 1, rows, 256);
        conv_forward(h, st, l.linear, hcat, none, nxt, 0);
        cur = nxt;
    }
    Act out;
    out.N = 1; out.H = 1; out.W = rows; out.C = w.num_class; out.p = logits;
    conv_forward(h, st, w.pred, cur, none, out, CONV_OUT_F32);
}

void crnn_forward_dev(Handle* h, Lane& lane, const float* x, int N, int Wm, float* logits) {
    const int T = Wm / 4 - 1, rows = N * T;
    DevBuf sb;
    Act seq = crnn_alloc_seq(h, lane.stream, sb, rows);
    crnn_features_dev(h, lane.stream, x, N, Wm, seq, 0);
    std::vector<SeqDesc> seqs(N);
    for (int i = 0; i < N; ++i) { seqs[i].row0 = i * T; seqs[i].T = T; }
    crnn_sequence_dev(h, lane, seq, seqs, logits);
}

// ------------------------------------------------------------------------------------------------------------------
// recognizer_predict (greedy) + CTCLabelConverter.decode_greedy + the inputs of custom_mean.
//   k_row_argmax  : one warp per time step: p = softmax(logits); p[ignore] = 0; p /= sum(p); (argmax, max p)
//   k_ctc_collapse: one warp per crop: keep t where idx[t] != idx[t-1] and idx[t] != 0 (ballot compaction)
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_row_argmax(const float* __restrict__ logits, int rows, int C, const uint8_t* __restrict__ ignore,
                             float* __restrict__ step_prob, int32_t* __restrict__ step_idx) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* lg = logits + (int64_t)row * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum_all = 0.f, sum_kept = 0.f, best = -1.f;
    int besti = INT_MAX;
    for (int c = lane; c < C; c += 32) {
        float e = expf(lg[c] - mx);
        sum_all += e;
        bool ig = ignore && ignore[c];
        if (!ig) sum_kept += e;
        float v = ig ? 0.f : e;
        if (v > best) { best = v; besti = c; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum_all += __shfl_xor_sync(0xffffffffu, sum_all, o);
        sum_kept += __shfl_xor_sync(0xffffffffu, sum_kept, o);
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane == 0) {
        step_prob[row] = (best / sum_all) / (sum_kept / sum_all);
        step_idx[row] = besti;
    }
}

__global__ void k_ctc_collapse(const int32_t* __restrict__ step_idx, const SeqDesc* __restrict__ seqs, int n_seq,
                               int32_t* __restrict__ text_idx, int32_t* __restrict__ text_len) {
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= n_seq) return;
    const int row0 = seqs[s].row0, T = seqs[s].T;
    int count = 0;
    for (int base = 0; base < T; base += 32) {
        int t = base + lane;
        int cur = t < T ? step_idx[row0 + t] : 0;
        int prev = (t > 0 && t < T) ? step_idx[row0 + t - 1] : -1;
        bool keep = t < T && cur != prev && cur != 0;
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) text_idx[row0 + count + __popc(m & ((1u << lane) - 1u))] = cur;
        count += __popc(m);
    }
    if (lane == 0) text_len[s] = count;
}

void ctc_decode_dev(Handle* h, cudaStream_t st, const float* logits, int rows, int C, const uint8_t* ignore_dev,
                    const SeqDesc* seqs_dev, int n_seq, int32_t* text_idx, int32_t* text_len, float* step_prob,
                    int32_t* step_idx) {
    if (rows == 0 || n_seq == 0) return;
    k_row_argmax<<<cdiv(rows * 32, 256), 256, 0, st>>>(logits, rows, C, ignore_dev, step_prob, step_idx);
    k_ctc_collapse<<<cdiv(n_seq * 32, 128), 128, 0, st>>>(step_idx, seqs_dev, n_seq, text_idx, text_len);
    count_launch(h, 2);
    CUDA_CHECK(cudaGetLastError());
}

// the full probability matrix of recognizer_predict for the beam-search decoders: same arithmetic as k_row_argmax
__global__ void k_row_probs(const float* __restrict__ logits, int rows, int C, const uint8_t* __restrict__ ignore,
                            float* __restrict__ probs) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* lg = logits + (int64_t)row * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum_all = 0.f, sum_kept = 0.f;
    for (int c = lane; c < C; c += 32) {
        float e = expf(lg[c] - mx);
        sum_all += e;
        if (!(ignore && ignore[c])) sum_kept += e;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum_all += __shfl_xor_sync(0xffffffffu, sum_all, o);
        sum_kept += __shfl_xor_sync(0xffffffffu, sum_kept, o);
    }
    const float norm = sum_kept / sum_all;
    for (int c = lane; c < C; c += 32) {
        const bool ig = ignore && ignore[c];
        probs[(int64_t)row * C + c] = ig ? 0.f : (expf(lg[c] - mx) / sum_all) / norm;
    }
}

void row_probs_dev(Handle* h, cudaStream_t st, const float* logits, int rows, int C, const uint8_t* ignore_dev, float* probs) {
    if (rows == 0) return;
    k_row_probs<<<cdiv(rows * 32, 256), 256, 0, st>>>(logits, rows, C, ignore_dev, probs);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// utils.make_rotated_img_list for the eligible angles: scipy.ndimage.rotate(img, 90 k, reshape=True) == np.rot90(img, k)
//   k = 1: out[i][j] = src[j][W-1-i] (W x H);  k = 2: out[i][j] = src[H-1-i][W-1-j];  k = 3: out[i][j] = src[H-1-j][i] (W x H)
__global__ void k_rotate_crops(uint8_t* __restrict__ crops, const RotDesc* __restrict__ descs) {
    const RotDesc d = descs[blockIdx.y];
    const int n = d.sh * d.sw;
    const int oh = d.k == 2 ? d.sh : d.sw, ow = d.k == 2 ? d.sw : d.sh;
    (void)oh;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int i = p / ow, j = p - i * ow;
        int sy, sx;
        if (d.k == 1) { sy = j; sx = d.sw - 1 - i; }
        else if (d.k == 2) { sy = d.sh - 1 - i; sx = d.sw - 1 - j; }
        else { sy = d.sh - 1 - j; sx = i; }
        crops[d.dst_off + p] = crops[d.src_off + sy * d.sw + sx];
    }
}

void rotate_crops_dev(Handle* h, cudaStream_t st, uint8_t* crops, const RotDesc* descs_dev, int n, int max_pixels) {
    if (n == 0) return;
    k_rotate_crops<<<dim3(std::min(64, cdiv(max_pixels, 256)), n), 256, 0, st>>>(crops, descs_dev);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace bbocr
