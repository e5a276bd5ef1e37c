// craft.cu -- detector front half: easyocr/detection.py::test_net up to the score maps.
//   resize_aspect_ratio (easyocr/imgproc.py) -> normalizeMeanVariance -> CRAFT.forward (easyocr/craft.py,
//   easyocr/model/modules.py::vgg16_bn).   SURVEY.md §8a B2-B4.
#include "engine.h"
#include "resize.cuh"

namespace bbocr {

CanvasGeom canvas_geom(int H, int W, int canvas_size, double mag_ratio) {
    CanvasGeom g;
    g.H = H; g.W = W;
    double target = mag_ratio * std::max(H, W);          // Python float arithmetic
    if (target > canvas_size) target = canvas_size;
    g.ratio = target / std::max(H, W);
    g.th = (int)(H * g.ratio);
    g.tw = (int)(W * g.ratio);
    g.H32 = g.th % 32 ? g.th + (32 - g.th % 32) : g.th;
    g.W32 = g.tw % 32 ? g.tw + (32 - g.tw % 32) : g.tw;
    return g;
}

__global__ void k_resize_bilinear_u8(const uint8_t* __restrict__ src, int sH, int sW, int sstride, int C,
                                     uint8_t* __restrict__ dst, int dH, int dW, double scale_x, double scale_y) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dW * C) return;
    int px = x / C, c = x - px * C;
    dst[(int64_t)y * dW * C + x] = bilinear_u8_px(src, sH, sW, sstride, C, c, px, y, scale_x, scale_y);
}

void resize_bilinear_u8(Handle* h, cudaStream_t st, const uint8_t* src, int sH, int sW, int sstride, int C, uint8_t* dst,
                        int dH, int dW) {
    double scale_x = 1.0 / ((double)dW / sW), scale_y = 1.0 / ((double)dH / sH);
    k_resize_bilinear_u8<<<dim3(cdiv(dW * C, 256), dH), 256, 0, st>>>(src, sH, sW, sstride, C, dst, dH, dW, scale_x, scale_y);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// canvas: zero-padded (H32 x W32) float image, then (x - mean*255) / (std*255) per channel; 4th lane = 0
__global__ void k_canvas(const uint8_t* __restrict__ img, int th, int tw, float* __restrict__ out, int H32, int W32,
                         float m0, float m1, float m2, float s0, float s1, float s2) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W32) return;
    float r = 0.f, g = 0.f, b = 0.f;
    if (y < th && x < tw) {
        const uint8_t* p = img + ((int64_t)y * tw + x) * 3;
        r = p[0]; g = p[1]; b = p[2];
    }
    float4 o;
    o.x = __fdiv_rn(__fsub_rn(r, m0), s0);
    o.y = __fdiv_rn(__fsub_rn(g, m1), s1);
    o.z = __fdiv_rn(__fsub_rn(b, m2), s2);
    o.w = 0.f;
    reinterpret_cast<float4*>(out)[(int64_t)y * W32 + x] = o;
}

// BF16 path: normalise + gather the 3x3x3 neighbourhood of every canvas pixel into 32 bf16 channels (tap*3 + c; 27..31 = 0),
// so that conv1_1 runs as a K = 32 GEMM on the tensor cores.  Outside the canvas = conv zero padding; inside the canvas
// but outside the image = the normalised zero pixel (resize_aspect_ratio pads before normalizeMeanVariance).
__global__ void __launch_bounds__(128) k_im2col_rgb(const uint8_t* __restrict__ img, int th, int tw, __nv_bfloat16* __restrict__ out,
                                                    int H32, int W32, float m0, float m1, float m2, float s0, float s1, float s2) {
    // (v - mean) / std has only 3 x 256 possible results: one table per block (6 IEEE divisions per thread instead of 27),
    // bit-identical to evaluating the expression per tap
    __shared__ uint16_t lut[3][256];
    {
        const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
        for (int i = threadIdx.x; i < 768; i += blockDim.x) {
            const int c = i >> 8, v = i & 255;
            const __nv_bfloat16 b = __float2bfloat16_rn(__fdiv_rn(__fsub_rn((float)v, mean[c]), sd[c]));
            lut[c][v] = *reinterpret_cast<const uint16_t*>(&b);
        }
    }
    __syncthreads();
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W32) return;
    uint16_t v[32];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        const bool in_canvas = yy >= 0 && yy < H32 && xx >= 0 && xx < W32;
        const bool in_img = in_canvas && yy < th && xx < tw;
        const uint8_t* p = img + ((int64_t)yy * tw + xx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int px = in_img ? (int)__ldg(p + c) : 0;
            v[t * 3 + c] = in_canvas ? lut[c][px] : (uint16_t)0;
        }
    }
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = 0;
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = (uint32_t)v[2 * i] | ((uint32_t)v[2 * i + 1] << 16);
    uint4* o = reinterpret_cast<uint4*>(out + ((int64_t)y * W32 + x) * 32);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    o[2] = make_uint4(w[8], w[9], w[10], w[11]);
    o[3] = make_uint4(w[12], w[13], w[14], w[15]);
}

// bf16x3 mode: the same gather, every normalised value as bf16 hi + bf16 lo (two [px][32] tensors)
__global__ void __launch_bounds__(128) k_im2col_rgb_split(const uint8_t* __restrict__ img, int th, int tw, __nv_bfloat16* __restrict__ ohi,
                                                          __nv_bfloat16* __restrict__ olo, int H32, int W32, float m0, float m1, float m2,
                                                          float s0, float s1, float s2) {
    __shared__ uint32_t lut[3][256];          // hi | lo << 16
    {
        const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
        for (int i = threadIdx.x; i < 768; i += blockDim.x) {
            const int c = i >> 8, v = i & 255;
            const float f = __fdiv_rn(__fsub_rn((float)v, mean[c]), sd[c]);
            const __nv_bfloat16 hb = __float2bfloat16_rn(f);
            const __nv_bfloat16 lb = __float2bfloat16_rn(f - __bfloat162float(hb));
            lut[c][v] = (uint32_t)*reinterpret_cast<const uint16_t*>(&hb) | ((uint32_t)*reinterpret_cast<const uint16_t*>(&lb) << 16);
        }
    }
    __syncthreads();
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W32) return;
    uint32_t v[32];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        const bool in_canvas = yy >= 0 && yy < H32 && xx >= 0 && xx < W32;
        const bool in_img = in_canvas && yy < th && xx < tw;
        const uint8_t* p = img + ((int64_t)yy * tw + xx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int px = in_img ? (int)__ldg(p + c) : 0;
            v[t * 3 + c] = in_canvas ? lut[c][px] : 0u;
        }
    }
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = 0;
    uint32_t wh[16], wl[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        wh[i] = (v[2 * i] & 0xffffu) | (v[2 * i + 1] << 16);
        wl[i] = (v[2 * i] >> 16) | (v[2 * i + 1] & 0xffff0000u);
    }
    // 64 bytes per pixel and tensor: two 256-bit stores each (one 32-byte sector per request)
    __nv_bfloat16* oh = ohi + ((int64_t)y * W32 + x) * 32;
    __nv_bfloat16* ol = olo + ((int64_t)y * W32 + x) * 32;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(oh + 16 * q), "r"(wh[8 * q]), "r"(wh[8 * q + 1]),
                     "r"(wh[8 * q + 2]), "r"(wh[8 * q + 3]), "r"(wh[8 * q + 4]), "r"(wh[8 * q + 5]), "r"(wh[8 * q + 6]), "r"(wh[8 * q + 7]) : "memory");
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ol + 16 * q), "r"(wl[8 * q]), "r"(wl[8 * q + 1]),
                     "r"(wl[8 * q + 2]), "r"(wl[8 * q + 3]), "r"(wl[8 * q + 4]), "r"(wl[8 * q + 5]), "r"(wl[8 * q + 6]), "r"(wl[8 * q + 7]) : "memory");
    }
}

extern thread_local int g_conv_scope;

void craft_forward_dev(Handle* h, cudaStream_t st, const uint8_t* img_dev, const CanvasGeom& g, float* text, float* link) {
    craft_forward_batch_dev(h, st, &img_dev, 1, g, text, link);
}

// nimg images of identical geometry through the network as ONE batch (NHWC with N = nimg): every layer is one launch, so
// the per-launch fixed costs (prologue, resident-weight load, tail wave) are shared and the small 1/16-resolution layers
// get nimg x the tiles.  text / link: [nimg][H32/2][W32/2] planes.
void craft_forward_batch_dev(Handle* h, cudaStream_t st, const uint8_t* const* imgs_dev, int nimg, const CanvasGeom& g, float* text,
                             float* link) {
    struct Scope { Scope() { g_conv_scope = 1; } ~Scope() { g_conv_scope = 0; } } scope_guard;
    if (!h->craft_loaded) fail(BBOCR_E_STATE, "CRAFT weights not loaded (bbocr_load_craft)");
    const CraftW& w = h->craft;
    const int H = g.H32, W = g.W32;
    ARG_CHECK(nimg >= 1 && nimg <= 16, "craft: batch of %d images", nimg);
    const float m0 = (float)(0.485 * 255.0), m1 = (float)(0.456 * 255.0), m2 = (float)(0.406 * 255.0);
    const float s0 = (float)(0.229 * 255.0), s1 = (float)(0.224 * 255.0), s2 = (float)(0.225 * 255.0);
    const bool tc_first = h->precision == BBOCR_PREC_BF16 && !h->force_generic_conv;
    const bool split = tc_first && h->det_split;        // bf16x3: every activation tensor is a hi/lo pair
    // bf16x3 stem: the gathered 32-channel stem (hi + lo) on the tensor cores like bf16 mode: 553 us per 1920x1440 page.  The
    // alternative -- the 27-tap conv1_1 on the CUDA cores straight from the FP32 canvas -- measured 827 us (its per-pixel 128-byte
    // hi / lo stores are written 8 bytes per lane), kept behind BBOCR_X3_STEM_DIRECT for A/B.
    static const bool stem_direct_x3 = getenv("BBOCR_X3_STEM_DIRECT") != nullptr;
    const bool gather = tc_first && (!split || !stem_direct_x3);
    const size_t canvas_px_bytes = gather ? 64 : 16;
    const size_t canvas_plane = (((size_t)nimg * H * W * canvas_px_bytes) + 255) & ~(size_t)255;
    // bf16x3: gather + conv1_1 fused (k_conv_stem builds the A tiles in shared memory from a 16-byte-per-pixel normalised canvas;
    // no 32-channel stem tensor in HBM).  BBOCR_STEM_FUSED=0: the two-kernel path (k_im2col_rgb_split + k_conv_tc), A/B switch
    static const bool stem_fused_on = !(getenv("BBOCR_STEM_FUSED") && atoi(getenv("BBOCR_STEM_FUSED")) == 0);
    const bool need_resize = g.th != g.H || g.tw != g.W;
    const ConvW& c11 = w.c1_1_tc;                        // the fused stem is written for CRAFT's conv1_1 (27 -> 32 gathered channels, 64 outputs)
    const bool stem_fused = stem_fused_on && split && gather && c11.w_split && c11.cin == 32 && c11.cout == 64 && c11.cout_pad == 64 &&
                            c11.kh == 1 && c11.kw == 1;
    const float mean3[3] = {m0, m1, m2}, sd3[3] = {s0, s1, s2};
    DevBuf canvas(stem_fused ? (size_t)nimg * H * W * 16 : canvas_plane * (split && gather ? 2 : 1), st), resized;
    if (need_resize) resized.alloc((size_t)g.th * g.tw * 3, st);
    for (int i = 0; i < nimg; ++i) {
        const uint8_t* src = imgs_dev[i];
        if (need_resize) {
            resize_bilinear_u8(h, st, imgs_dev[i], g.H, g.W, g.W * 3, 3, resized.as<uint8_t>(), g.th, g.tw);
            src = resized.as<uint8_t>();
        }
        if (stem_fused) {
            stem_norm_forward(h, st, src, g.th, g.tw, canvas.as<uint8_t>() + (size_t)i * H * W * 16, H, W, mean3, sd3);
            continue;
        }
        uint8_t* dst = canvas.as<uint8_t>() + (size_t)i * H * W * canvas_px_bytes;
        if (split && gather)
            k_im2col_rgb_split<<<dim3(cdiv(W, 128), H), 128, 0, st>>>(src, g.th, g.tw, reinterpret_cast<__nv_bfloat16*>(dst),
                                                                      reinterpret_cast<__nv_bfloat16*>(dst + canvas_plane), H, W, m0, m1, m2, s0, s1, s2);
        else if (gather)
            k_im2col_rgb<<<dim3(cdiv(W, 128), H), 128, 0, st>>>(src, g.th, g.tw, reinterpret_cast<__nv_bfloat16*>(dst), H, W, m0, m1, m2, s0, s1, s2);
        else
            k_canvas<<<dim3(cdiv(W, 256), H), 256, 0, st>>>(src, g.th, g.tw, reinterpret_cast<float*>(dst), H, W, m0, m1, m2, s0, s1, s2);
        count_launch(h);
    }
    CUDA_CHECK(cudaGetLastError());

    const Act none;
    auto alloc = [&](DevBuf& buf, int n, int hh, int ww, int c) {
        return split ? act_alloc_split(h, st, buf, n, hh, ww, c) : act_alloc(h, st, buf, n, hh, ww, c);
    };
    auto conv = [&](const ConvW& cw, const Act& a, const Act& b, DevBuf& buf, int flags) {
        Act o = alloc(buf, a.N, a.H + 2 * cw.pad - cw.dil * (cw.kh - 1), a.W + 2 * cw.pad - cw.dil * (cw.kw - 1), cw.cout);
        conv_forward(h, st, cw, a, b, o, flags);
        return o;
    };
    auto up2 = [&](const Act& a, DevBuf& buf) {
        Act o = alloc(buf, a.N, a.H * 2, a.W * 2, a.C);
        if (split) upsample2x_split(h, st, a, o);
        else upsample2x(h, st, a, o);
        return o;
    };
    // conv + ReLU + MaxPool2d(2,2) in one launch; keep_full also materialises the un-pooled tensor (skip connection)
    auto conv_pool = [&](const ConvW& cw, const Act& a, DevBuf* full_buf, DevBuf& pool_buf, Act* full_out) {
        Act full;
        full.N = a.N; full.H = a.H; full.W = a.W; full.C = cw.cout; full.p = nullptr;
        if (full_buf) full = alloc(*full_buf, a.N, a.H, a.W, cw.cout);
        Act pooled = alloc(pool_buf, a.N, a.H / 2, a.W / 2, cw.cout);
        conv_forward(h, st, cw, a, none, full, CONV_RELU | CONV_POOL22, &pooled);
        if (full_out) *full_out = full;
        return pooled;
    };
    const int R = CONV_RELU;
    DevBuf b0, b1, b_r22, b_r32, b_r43, b_r53;
    // slice1
    Act a = alloc(b0, nimg, H, W, 64);
    if (stem_fused && conv_stem_supported(w.c1_1_tc, a)) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (h->conv_timing) {                           // same instrumentation as conv_forward (bench.py roofline)
            CUDA_CHECK(cudaEventCreate(&e0));
            CUDA_CHECK(cudaEventCreate(&e1));
            CUDA_CHECK(cudaEventRecord(e0, st));
        }
        conv_stem_forward(h, st, w.c1_1_tc, canvas.p, a, R);
        if (h->conv_timing) {
            CUDA_CHECK(cudaEventRecord(e1, st));
            std::lock_guard<std::mutex> lk(h->stat_mu);
            h->conv_events.emplace_back(e0, e1);
            h->conv_flops += 2.0 * (double)a.N * a.H * a.W * w.c1_1_tc.cout * w.c1_1_tc.cin;
            h->conv_launches += 1;
        }
    } else if (gather) {
        ARG_CHECK(!stem_fused, "conv_stem: layer shape not supported");
        Act x32;
        x32.N = nimg; x32.H = H; x32.W = W; x32.C = 32; x32.p = canvas.p;
        if (split) x32.lo = canvas.as<uint8_t>() + canvas_plane;
        conv_forward(h, st, w.c1_1_tc, x32, none, a, R);
    } else {
        conv_first(h, st, w.c1_1, canvas.as<float>(), nimg, H, W, 4, a, R);
    }
    canvas.release();
    a = conv_pool(w.c1_2, a, nullptr, b1, nullptr);
    a = conv(w.c2_1, a, none, b0, R);
    Act r22;                                          // in-place ReLU of slice2[12] rectifies the aliased tap
    a = conv_pool(w.c2_2, a, &b_r22, b1, &r22);       // slice2 starts with the pool
    // slice2
    a = conv(w.c3_1, a, none, b0, R);
    std::swap(b0, b1);
    Act r32 = conv(w.c3_2, a, none, b_r32, R);
    // slice3
    a = conv_pool(w.c3_3, r32, nullptr, b1, nullptr);
    a = conv(w.c4_1, a, none, b0, R);
    Act r43 = conv(w.c4_2, a, none, b_r43, R);
    // slice4
    a = conv_pool(w.c4_3, r43, nullptr, b1, nullptr);
    a = conv(w.c5_1, a, none, b0, R);
    Act r53 = conv(w.c5_2, a, none, b_r53, 0);        // followed by MaxPool, not ReLU: stays the raw BN output
    // slice5
    a = alloc(b0, nimg, r53.H, r53.W, 512);
    if (split) maxpool_split(h, st, r53, a, 3, 3, 1, 1, 1, 1);
    else maxpool(h, st, r53, a, 3, 3, 1, 1, 1, 1);
    a = conv(w.fc6, a, none, b1, 0);
    Act fc7 = conv(w.fc7, a, none, b0, 0);
    // decoder
    a = conv(w.up1a, fc7, r53, b1, R);
    a = conv(w.up1b, a, none, b0, R);
    b_r53.release();
    a = up2(a, b1);
    a = conv(w.up2a, a, r43, b0, R);
    a = conv(w.up2b, a, none, b1, R);
    b_r43.release();
    a = up2(a, b0);
    a = conv(w.up3a, a, r32, b1, R);
    a = conv(w.up3b, a, none, b0, R);
    b_r32.release();
    a = up2(a, b1);
    a = conv(w.up4a, a, r22, b0, R);
    a = conv(w.up4b, a, none, b1, R);
    b_r22.release();
    a = conv(w.cls0, a, none, b0, R);
    a = conv(w.cls1, a, none, b1, R);
    if (h->precision == BBOCR_PREC_BF16 && !split && !h->force_generic_conv && conv_res_cls_tail_supported(w.cls2, w.cls3, w.cls4, a)) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;            // counted with the detector convolutions (bench.py roofline)
        if (h->conv_timing) { CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1)); CUDA_CHECK(cudaEventRecord(e0, st)); }
        conv_res_cls_tail(h, st, w.cls2, w.cls3, w.cls4, a, text, link);      // conv_cls[4] + the 1x1 tail in one launch
        if (h->conv_timing) {
            CUDA_CHECK(cudaEventRecord(e1, st));
            std::lock_guard<std::mutex> g(h->stat_mu);
            h->conv_events.emplace_back(e0, e1);
            h->conv_flops += 2.0 * (double)a.N * a.H * a.W * (16.0 * 32 * 9 + 16 * 16 + 16 * 2);
            h->conv_launches += 1;
        }
        return;
    }
    if (split) {                                       // conv_cls[4] to an FP32 tensor, then the FP32 1x1 tail
        Act o = act_alloc(h, st, b0, a.N, a.H, a.W, w.cls2.cout, true);
        conv_forward(h, st, w.cls2, a, none, o, R | CONV_OUT_F32);
        cls_tail_f32(h, st, w.cls3, w.cls4, o, text, link);
        return;
    }
    a = conv(w.cls2, a, none, b0, R);
    cls_tail(h, st, w.cls3, w.cls4, a, text, link);
}

}  // namespace bbocr
