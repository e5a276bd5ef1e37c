// api.cu -- the C ABI declared in include/bbocr.h and the readtext orchestration
//   (easyocr/easyocr.py::Reader.readtext -> detect -> recognize, batch_size == 1 semantics).
#include <condition_variable>
#include <thread>

#include "engine.h"

struct bbocr_handle : bbocr::Handle {};

using namespace bbocr;

namespace bbocr {
extern thread_local int g_conv_scope;
}

#include <chrono>

namespace {

thread_local std::string g_create_err;

struct StageTimer {        // wall-clock per readtext stage (includes the stream waits): bbocr_dbg_stage_ms
    Handle* h; int idx; std::chrono::steady_clock::time_point t0;
    StageTimer(Handle* hh, int i) : h(hh), idx(i), t0(std::chrono::steady_clock::now()) {}
    ~StageTimer() {
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::lock_guard<std::mutex> g(h->stat_mu);
        h->stage_ms[idx] += ms;
    }
};

template <typename F>
int guarded(bbocr_handle* h, F&& f) {
    if (!h) return BBOCR_E_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    try {
        CUDA_CHECK(cudaSetDevice(h->device));
        f();
        return BBOCR_OK;
    } catch (const Error& e) {
        h->err = e.what();
        cudaGetLastError();
        return e.code;
    } catch (const std::bad_alloc&) {
        h->err = "out of host memory";
        return BBOCR_E_NOMEM;
    } catch (const std::exception& e) {
        h->err = e.what();
        return BBOCR_E_ARG;
    }
}

// host -> device through the lane's pinned staging buffer
void* staging(Lane& lane, size_t bytes) {          // pinned H2D staging; waits for the previous copy out of it
    if (lane.in_busy) {
        CUDA_CHECK(stream_sync(lane.stream));
        lane.in_busy = false;
    }
    return lane.pin_in.get(bytes);
}

void upload(Lane& lane, DevBuf& dst, const void* src, size_t bytes) {
    dst.alloc(bytes, lane.stream);
    uint8_t* pin = (uint8_t*)staging(lane, bytes);
    constexpr size_t kChunk = 16u << 20;
    if (bytes <= 2 * kChunk) {
        memcpy(pin, src, bytes);
        CUDA_CHECK(cudaMemcpyAsync(dst.p, pin, bytes, cudaMemcpyHostToDevice, lane.stream));
    } else {
        // a large page (recognition-only pages are hundreds of MB): a single-threaded pageable -> pinned memcpy runs at ~6 GB/s
        // and used to be 40 % of the call.  Four threads copy 16 MB chunks; every chunk's DMA is enqueued as soon as it is staged.
        const size_t n_chunks = (bytes + kChunk - 1) / kChunk;
        std::atomic<size_t> next{0};
        std::vector<std::atomic<int>> ready(n_chunks);
        for (auto& r : ready) r.store(0);
        auto work = [&] {
            for (size_t c; (c = next.fetch_add(1)) < n_chunks;) {
                const size_t o = c * kChunk, n = std::min(kChunk, bytes - o);
                memcpy(pin + o, (const uint8_t*)src + o, n);
                ready[c].store(1, std::memory_order_release);
            }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < 3; ++t) th.emplace_back(work);
        std::thread last(work);
        cudaError_t err = cudaSuccess;
        for (size_t c = 0; c < n_chunks; ++c) {
            while (!ready[c].load(std::memory_order_acquire)) std::this_thread::yield();
            const size_t o = c * kChunk, n = std::min(kChunk, bytes - o);
            if (err == cudaSuccess) err = cudaMemcpyAsync((uint8_t*)dst.p + o, pin + o, n, cudaMemcpyHostToDevice, lane.stream);
        }
        for (auto& t : th) t.join();
        last.join();
        CUDA_CHECK(err);
    }
    lane.in_busy = true;
}

void download(Lane& lane, void* dst, const void* src_dev, size_t bytes) {
    void* pin = lane.pin_out.get(bytes);
    CUDA_CHECK(cudaMemcpyAsync(pin, src_dev, bytes, cudaMemcpyDeviceToHost, lane.stream));
    CUDA_CHECK(stream_sync(lane.stream));
    lane.in_busy = false;
    memcpy(dst, pin, bytes);
}

// One preprocessing step on host buffers: upload, run, download.
template <typename F>
int pp_step(bbocr_handle* h, const uint8_t* src, size_t in_bytes, uint8_t* out, size_t out_bytes, F&& run) {
    return guarded(h, [&] {
        ARG_CHECK(src && out && in_bytes > 0 && out_bytes > 0, "null or empty buffer");
        Lane& lane = h->lanes[0];
        DevBuf din, dout(out_bytes, lane.stream);
        upload(lane, din, src, in_bytes);
        run(lane.stream, din.as<uint8_t>(), dout.as<uint8_t>());
        download(lane, out, dout.p, out_bytes);
    });
}

}  // namespace

extern "C" {

const char* bbocr_version(void) { return "bbocr-b200 0.1 (sm_100a)"; }

int bbocr_create(int device, bbocr_handle** out) {
    if (!out) return BBOCR_E_ARG;
    *out = nullptr;
    try {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            fail(BBOCR_E_CUDA, "no CUDA device available (%s); libbbocr has no CPU path", cudaGetErrorString(e));
        ARG_CHECK(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
        CUDA_CHECK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            fail(BBOCR_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                 prop.major, prop.minor);
        std::unique_ptr<bbocr_handle> h(new bbocr_handle());
        h->device = device;
        h->sm_count = prop.multiProcessorCount;
        {
            const char* e = getenv("BBOCR_LANES");        // detector pages in flight per handle (streams + host threads)
            int nl = e ? atoi(e) : 8;
            h->n_det_lanes = std::min(std::max(nl, 1), 32);
            const char* r = getenv("BBOCR_REC_LANES");    // recogniser groups in flight
            int nr = r ? atoi(r) : 2;
            const char* b = getenv("BBOCR_DET_BATCH");    // same-size pages per detector-network launch chain
            h->det_batch = std::min(std::max(b ? atoi(b) : 2, 1), 16);
            const char* g = getenv("BBOCR_REC_GROUP");    // pages per recogniser group
            h->rec_group = std::max(1, g ? atoi(g) : 32);
            h->lanes.resize(h->n_det_lanes + std::min(std::max(nr, 1), 8));
        }
        for (auto& l : h->lanes) CUDA_CHECK(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        // keep freed blocks in the stream-ordered pool instead of returning them to the driver after every sync
        cudaMemPool_t pool;
        CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thresh = UINT64_MAX;
        CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
        {
            // Reserve the working set up front: growing the pool later means cudaMalloc-class calls that serialise
            // every lane.  180 GB of HBM3e per GPU; the default reservation covers 8 pages of 1920x1440 in flight.
            const char* e = getenv("BBOCR_POOL_GB");
            double gb = e ? atof(e) : 32.0;
            size_t free_b = 0, total_b = 0;
            CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
            size_t want = std::min((size_t)(gb * (1ull << 30)), free_b / 2);
            if (want > 0) {
                void* p = nullptr;
                cudaStream_t s0 = h->lanes[0].stream;
                if (cudaMallocAsync(&p, want, s0) == cudaSuccess) cudaFreeAsync(p, s0);
                else cudaGetLastError();
                CUDA_CHECK(stream_sync(s0));
            }
        }
        *out = h.release();
        return BBOCR_OK;
    } catch (const Error& e) {
        g_create_err = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_create_err = e.what();
        return BBOCR_E_ARG;
    }
}

void bbocr_destroy(bbocr_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& l : h->lanes)
        if (l.stream) cudaStreamDestroy(l.stream);
    for (auto& l : h->jpeg_lanes)
        if (l.stream) cudaStreamDestroy(l.stream);
    for (auto* v : {&h->lanes, &h->jpeg_lanes})
        for (auto& l : *v)
            for (auto& sc : l.scr)
                if (sc.p) cudaFree(sc.p);
    for (void* p : h->owned) cudaFree(p);
    for (auto& kv : h->cubic_cache) cudaFree(kv.second.first);
    for (auto& ev : h->conv_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    delete h;
}

const char* bbocr_last_error(const bbocr_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int bbocr_load_craft(bbocr_handle* h, const bbocr_tensor* t, int n) {
    return guarded(h, [&] { ARG_CHECK(t && n > 0, "no tensors"); load_craft(h, t, n); });
}
int bbocr_load_crnn(bbocr_handle* h, const bbocr_tensor* t, int n) {
    return guarded(h, [&] { ARG_CHECK(t && n > 0, "no tensors"); load_crnn(h, t, n); });
}
int bbocr_set_precision(bbocr_handle* h, int prec) {
    return guarded(h, [&] {
        ARG_CHECK(prec == BBOCR_PREC_FP32 || prec == BBOCR_PREC_BF16 || prec == BBOCR_PREC_BF16X3, "unknown precision %d", prec);
        h->precision = prec == BBOCR_PREC_FP32 ? BBOCR_PREC_FP32 : BBOCR_PREC_BF16;      // kernel family: CUDA cores / tcgen05
        h->det_split = prec == BBOCR_PREC_BF16X3;
    });
}
int bbocr_set_dictionary(bbocr_handle* h, const int32_t* idx, const int32_t* off, int n) {
    return guarded(h, [&] {
        ARG_CHECK(n >= 0 && (n == 0 || (idx && off)), "bad dictionary");
        h->dict.clear();
        for (int i = 0; i < n; ++i) {
            ARG_CHECK(off[i + 1] >= off[i], "dictionary offsets must ascend");
            h->dict.emplace(idx + off[i], idx + off[i + 1]);
        }
    });
}
int bbocr_ctc_beam_decode(const float* probs, int T, int C, int decoder, int beam_width, int space_idx, const int32_t* dict_idx,
                          const int32_t* dict_off, int n_dict, int32_t* text_out, int cap, int* len) {
    if (!probs || T < 0 || C < 2 || (decoder != 1 && decoder != 2) || beam_width < 1 || !text_out || !len || n_dict < 0 ||
        (n_dict > 0 && (!dict_idx || !dict_off)))
        return BBOCR_E_ARG;
    try {
        std::set<std::vector<int32_t>> dict;
        for (int i = 0; i < n_dict; ++i) dict.emplace(dict_idx + dict_off[i], dict_idx + dict_off[i + 1]);
        std::vector<int32_t> text;
        bbocr::decode_beam(probs, T, C, decoder, beam_width, space_idx, &dict, text);
        *len = (int)text.size();
        for (int i = 0; i < (int)text.size() && i < cap; ++i) text_out[i] = text[i];
    } catch (...) {
        return BBOCR_E_ARG;
    }
    return BBOCR_OK;
}
int bbocr_get_precision(const bbocr_handle* h) {
    if (!h) return BBOCR_E_ARG;
    return h->det_split ? BBOCR_PREC_BF16X3 : h->precision;
}

// ---- preprocessing ---------------------------------------------------------------------------------------------------
int bbocr_preprocess_launches_per_image(void) { return preprocess_launches_per_image(); }

int bbocr_preprocess_u8(bbocr_handle* h, const uint8_t* bgr, int H, int W, int stride, int in_on_device,
                        const bbocr_pp_params* p, uint8_t* out, int out_on_device, int* outH, int* outW) {
    return guarded(h, [&] {
        ARG_CHECK(bgr && p && out && outH && outW && H > 0 && W > 0 && stride >= W * 3, "bad arguments");
        Lane& lane = h->lanes[0];
        DevBuf din, dout;
        const uint8_t* src = bgr;
        if (!in_on_device) {
            upload(lane, din, bgr, (size_t)H * stride);
            src = din.as<uint8_t>();
        }
        int dH = (int)(H * (double)p->scale), dW = (int)(W * (double)p->scale);
        ARG_CHECK(dH > 0 && dW > 0, "empty output");
        uint8_t* dst = out;
        if (!out_on_device) {
            dout.alloc((size_t)dH * dW, lane.stream);
            dst = dout.as<uint8_t>();
        }
        preprocess_chain_dev(h, lane.stream, src, H, W, stride, *p, dst, outH, outW);
        if (!out_on_device) download(lane, out, dst, (size_t)dH * dW);
        else CUDA_CHECK(stream_sync(lane.stream));
    });
}

// n same-size photos: photo i runs on lane i % n_lanes (its own stream and pinned staging), so the seven launches of
// different photos overlap and there is one synchronisation per lane instead of one per photo.
int bbocr_preprocess_batch_u8(bbocr_handle* h, int n, const uint8_t* const* bgr, int H, int W, int stride, int in_on_device,
                              const bbocr_pp_params* p, uint8_t* const* out, int out_on_device, int* outH, int* outW) {
    return guarded(h, [&] {
        ARG_CHECK(n >= 0 && bgr && p && out && outH && outW && H > 0 && W > 0 && stride >= W * 3, "bad arguments");
        const int dH = (int)(H * (double)p->scale), dW = (int)(W * (double)p->scale);
        ARG_CHECK(dH > 0 && dW > 0, "empty output");
        *outH = dH; *outW = dW;
        const int nl = std::min<int>(h->n_det_lanes, std::max(n, 1));
        const size_t in_bytes = (size_t)H * stride, out_bytes = (size_t)dH * dW;
        for (int base = 0; base < n; base += nl) {
            const int m = std::min(nl, n - base);
            std::vector<DevBuf> din(m), dout(m);
            for (int k = 0; k < m; ++k) {
                ARG_CHECK(bgr[base + k] && out[base + k], "null image %d", base + k);
                Lane& lane = h->lanes[k];
                const uint8_t* src = bgr[base + k];
                if (!in_on_device) { upload(lane, din[k], src, in_bytes); src = din[k].as<uint8_t>(); }
                uint8_t* dst = out[base + k];
                if (!out_on_device) { dout[k].alloc(out_bytes, lane.stream); dst = dout[k].as<uint8_t>(); }
                int oh, ow;
                preprocess_chain_dev(h, lane.stream, src, H, W, stride, *p, dst, &oh, &ow);
            }
            for (int k = 0; k < m; ++k) {
                Lane& lane = h->lanes[k];
                if (!out_on_device) download(lane, out[base + k], dout[k].p, out_bytes);
                else CUDA_CHECK(stream_sync(lane.stream));
                lane.in_busy = false;
            }
        }
    });
}

int bbocr_pp_gray(bbocr_handle* h, const uint8_t* bgr, int H, int W, uint8_t* out) {
    return pp_step(h, bgr, (size_t)H * W * 3, out, (size_t)H * W,
                   [&](cudaStream_t st, const uint8_t* s, uint8_t* d) { pp_gray(h, st, s, H, W, W * 3, d); });
}
int bbocr_pp_resize_cubic(bbocr_handle* h, const uint8_t* src, int H, int W, int dstH, int dstW, int mode, uint8_t* out) {
    return pp_step(h, src, (size_t)H * W, out, (size_t)dstH * dstW,
                   [&](cudaStream_t st, const uint8_t* s, uint8_t* d) { pp_resize_cubic(h, st, s, H, W, d, dstH, dstW, mode); });
}
int bbocr_pp_gaussian3(bbocr_handle* h, const uint8_t* src, int H, int W, float sigma, uint8_t* out) {
    return pp_step(h, src, (size_t)H * W, out, (size_t)H * W,
                   [&](cudaStream_t st, const uint8_t* s, uint8_t* d) { pp_gaussian3(h, st, s, d, H, W, sigma, nullptr); });
}
static int tone_step(bbocr_handle* h, const uint8_t* src, int H, int W, float contrast, float brightness, uint8_t* out) {
    return pp_step(h, src, (size_t)H * W, out, (size_t)H * W, [&](cudaStream_t st, const uint8_t* s, uint8_t* d) {
        DevBuf small(8 + 256, st);
        unsigned long long* sum = small.as<unsigned long long>();
        uint8_t* lut = small.as<uint8_t>() + 8;
        pp_sum(h, st, s, (int64_t)H * W, sum);
        pp_tone_lut(h, st, sum, (int64_t)H * W, contrast, brightness, lut);
        pp_apply_lut(h, st, s, d, (int64_t)H * W, lut);
    });
}
int bbocr_pp_contrast(bbocr_handle* h, const uint8_t* src, int H, int W, float factor, uint8_t* out) {
    return tone_step(h, src, H, W, factor, 0.f, out);
}
int bbocr_pp_brightness(bbocr_handle* h, const uint8_t* src, int H, int W, float factor, uint8_t* out) {
    return tone_step(h, src, H, W, 0.f, factor, out);
}
int bbocr_pp_clahe(bbocr_handle* h, const uint8_t* src, int H, int W, float clip, uint8_t* out) {
    return pp_step(h, src, (size_t)H * W, out, (size_t)H * W, [&](cudaStream_t st, const uint8_t* s, uint8_t* d) {
        DevBuf small(64 * 256 * 4 + 64 * 256, st);
        unsigned int* hist = small.as<unsigned int>();
        uint8_t* luts = small.as<uint8_t>() + 64 * 256 * 4;
        pp_clahe_luts(h, st, s, H, W, clip, nullptr, hist, luts);
        pp_clahe_apply(h, st, s, d, H, W, nullptr, luts);
    });
}
int bbocr_pp_equalize_hist(bbocr_handle* h, const uint8_t* src, int H, int W, uint8_t* out) {
    if (H <= 0 || W <= 0 || (int64_t)H * W > INT32_MAX) return BBOCR_E_ARG;
    return pp_step(h, src, (size_t)H * W, out, (size_t)H * W,
                   [&](cudaStream_t st, const uint8_t* s, uint8_t* d) { pp_equalize_hist(h, st, s, d, H, W); });
}
int bbocr_pp_unsharp(bbocr_handle* h, const uint8_t* src, int H, int W, int percent, int threshold, uint8_t* out) {
    return pp_step(h, src, (size_t)H * W, out, (size_t)H * W, [&](cudaStream_t st, const uint8_t* s, uint8_t* d) {
        pp_unsharp(h, st, s, d, H, W, percent, threshold, nullptr, nullptr);
    });
}
int bbocr_pp_adaptive_threshold(bbocr_handle* h, const uint8_t* src, int H, int W, int method, int inv, int block,
                                float delta, uint8_t* out) {
    return pp_step(h, src, (size_t)H * W, out, (size_t)H * W, [&](cudaStream_t st, const uint8_t* s, uint8_t* d) {
        pp_adaptive_threshold(h, st, s, d, H, W, method, inv, block, delta);
    });
}
int bbocr_pp_deskew(bbocr_handle* h, const uint8_t* src, int H, int W, float max_deg, uint8_t* out, float* angle_out) {
    return pp_step(h, src, (size_t)H * W, out, (size_t)H * W, [&](cudaStream_t st, const uint8_t* s, uint8_t* d) {
        float a = pp_deskew(h, st, s, d, H, W, max_deg);
        if (angle_out) *angle_out = a;
    });
}

int bbocr_preprocess_scan_u8(bbocr_handle* h, const uint8_t* bgr, int H, int W, int stride, int in_on_device, float clahe_clip,
                             int block, float delta, float max_deg, uint8_t* out, int out_on_device, float* angle_out) {
    return guarded(h, [&] {
        ARG_CHECK(bgr && out && H > 0 && W > 0 && stride >= W * 3, "bad arguments");
        Lane& lane = h->lanes[0];
        cudaStream_t st = lane.stream;
        const size_t n = (size_t)H * W;
        DevBuf din, dout;
        const uint8_t* src = bgr;
        if (!in_on_device) { upload(lane, din, bgr, (size_t)H * stride); src = din.as<uint8_t>(); }
        uint8_t* dst = out;
        if (!out_on_device) { dout.alloc(n, st); dst = dout.as<uint8_t>(); }
        DevBuf gray(n, st), eq(n, st), rot(n, st), small(64 * 256 * 4 + 64 * 256, st);
        unsigned int* hist = small.as<unsigned int>();
        uint8_t* luts = small.as<uint8_t>() + 64 * 256 * 4;
        pp_gray(h, st, src, H, W, stride, gray.as<uint8_t>());
        pp_clahe_luts(h, st, gray.as<uint8_t>(), H, W, clahe_clip, nullptr, hist, luts);
        pp_clahe_apply(h, st, gray.as<uint8_t>(), eq.as<uint8_t>(), H, W, nullptr, luts);
        const float a = pp_deskew(h, st, eq.as<uint8_t>(), rot.as<uint8_t>(), H, W, max_deg);
        if (angle_out) *angle_out = a;
        pp_adaptive_threshold(h, st, rot.as<uint8_t>(), dst, H, W, 1, 0, block, delta);
        if (!out_on_device) download(lane, out, dst, n);
        else CUDA_CHECK(stream_sync(st));
        lane.in_busy = false;
    });
}

// ---- extractor glue ------------------------------------------------------------------------------------
// PIL.Image.thumbnail((m, m)) size rule (Image.py::thumbnail, default BICUBIC, reducing_gap = 2.0):
// aspect-preserving, never enlarges; round_aspect picks floor/ceil by the smaller aspect error (floor on ties), at least 1.
static void pil_thumbnail_size(int W, int H, int m, int* ow, int* oh) {
    int x = m, y = m;
    if (x >= W && y >= H) { *ow = W; *oh = H; return; }
    const double aspect = (double)W / (double)H;
    auto round_aspect = [](double number, auto key) {
        const double f = std::floor(number), c = std::ceil(number);
        const double best = key(f) <= key(c) ? f : c;
        return std::max((int)best, 1);
    };
    if ((double)x / (double)y >= aspect) x = round_aspect(y * aspect, [&](double n) { return std::fabs(aspect - n / y); });
    else y = round_aspect(x / aspect, [&](double n) { return n == 0 ? 0.0 : std::fabs(aspect - x / n); });
    *ow = x; *oh = y;
}

extern "C" int bbocr_thumbnail_u8(bbocr_handle* h, const uint8_t* src, int H, int W, int in_on_device, int max_dim, uint8_t* out,
                                  int out_on_device, int* outH, int* outW) {
    return guarded(h, [&] {
        ARG_CHECK(H > 0 && W > 0 && max_dim > 0 && outH && outW, "bad arguments");
        int ow, oh;
        pil_thumbnail_size(W, H, max_dim, &ow, &oh);
        *outW = ow; *outH = oh;
        if (!out) return;                                   // size query
        ARG_CHECK(src, "null image");
        // Image.resize(reducing_gap = 2.0): once an axis shrinks by >= 4x, factor = int(src / dst / 2.0) > 1 and the image is
        // box-reduced first (Image.reduce; the safe box of a full-image box is the full image), then the bicubic pass samples
        // the reduced image through the fractional box (0, 0, W / fx, H / fy) -- single-precision floats on Pillow's C side
        int fx = (int)((double)W / ow / 2.0), fy = (int)((double)H / oh / 2.0);
        fx = std::max(fx, 1); fy = std::max(fy, 1);
        Lane& lane = h->lanes[0];
        cudaStream_t st = lane.stream;
        DevBuf din, dout, scratch((size_t)H * std::max(ow, W) + 16, st), reduced;
        const uint8_t* s = src;
        if (!in_on_device) { upload(lane, din, src, (size_t)H * W); s = din.as<uint8_t>(); }
        uint8_t* d = out;
        if (!out_on_device) { dout.alloc((size_t)oh * ow, st); d = dout.as<uint8_t>(); }
        int sH = H, sW = W;
        float box_w = 0.f, box_h = 0.f;
        if (fx > 1 || fy > 1) {
            sW = (W + fx - 1) / fx; sH = (H + fy - 1) / fy;
            reduced.alloc((size_t)sH * sW + 16, st);
            pil_reduce_dev(h, st, s, H, W, fx, fy, reduced.as<uint8_t>());
            s = reduced.as<uint8_t>();
            box_w = (float)((double)W / fx); box_h = (float)((double)H / fy);
        }
        if (ow == sW && oh == sH && box_w == 0.f) CUDA_CHECK(cudaMemcpyAsync(d, s, (size_t)H * W, cudaMemcpyDeviceToDevice, st));
        else pil_resize_bicubic_dev(h, st, s, sH, sW, d, oh, ow, scratch.as<uint8_t>(), box_w, box_h);
        if (!out_on_device) download(lane, out, d, (size_t)oh * ow);
        else CUDA_CHECK(stream_sync(st));
    });
}

// ---- page crops ---------------------------------------------------------------------------------------
extern "C" int bbocr_autocrop_rect(bbocr_handle* h, const uint8_t* bgr, int H, int W, int channels, int stride, int in_on_device, int margin,
                                   int32_t rect[4], int* found, uint8_t* mask_out, uint8_t* merged_out, int32_t* boxes_out,
                                   int boxes_cap, int* nboxes, int32_t* otsu_out) {
    return guarded(h, [&] {
        ARG_CHECK(bgr && rect && found && H > 0 && W > 0 && margin >= 0, "bad arguments");
        ARG_CHECK((channels == 3 && stride >= W * 3) || (channels == 1 && stride == W), "channels must be 3 (BGR) or 1 (packed gray)");
        Lane& lane = h->lanes[0];
        DevBuf din;
        const uint8_t* src = bgr;
        if (!in_on_device) { upload(lane, din, bgr, (size_t)H * stride); src = din.as<uint8_t>(); }
        const bool want_dbg = mask_out || merged_out || boxes_out || nboxes || otsu_out;
        AutoCropDebug dbg;
        dbg.mask = mask_out;
        dbg.merged = merged_out;
        rect[0] = rect[1] = rect[2] = rect[3] = -1;
        *found = autocrop_dev(h, lane.stream, src, H, W, channels, stride, margin, rect, want_dbg ? &dbg : nullptr) ? 1 : 0;
        CUDA_CHECK(stream_sync(lane.stream));
        lane.in_busy = false;
        if (nboxes) *nboxes = (int)(dbg.boxes.size() / 4);
        if (boxes_out) memcpy(boxes_out, dbg.boxes.data(), std::min((size_t)std::max(boxes_cap, 0) * 16, dbg.boxes.size() * 4));
        if (otsu_out) { otsu_out[0] = dbg.otsu[0]; otsu_out[1] = dbg.otsu[1]; }
    });
}

extern "C" int bbocr_external_boxes(bbocr_handle* h, const uint8_t* binary, int H, int W, int32_t* boxes_out, int boxes_cap,
                                    int* nboxes) {
    return guarded(h, [&] {
        ARG_CHECK(binary && nboxes && H > 0 && W > 0, "bad arguments");
        Lane& lane = h->lanes[0];
        DevBuf din;
        upload(lane, din, binary, (size_t)H * W);
        std::vector<int32_t> boxes;
        external_boxes_dev(h, lane.stream, din.as<uint8_t>(), H, W, boxes);
        CUDA_CHECK(stream_sync(lane.stream));
        lane.in_busy = false;
        *nboxes = (int)(boxes.size() / 4);
        if (boxes_out) memcpy(boxes_out, boxes.data(), std::min((size_t)std::max(boxes_cap, 0) * 16, boxes.size() * 4));
    });
}

extern "C" int bbocr_rect_morph(bbocr_handle* h, const uint8_t* binary, int H, int W, int kw, int kh, int erode, uint8_t* out) {
    return pp_step(h, binary, (size_t)H * W, out, (size_t)H * W, [&](cudaStream_t st, const uint8_t* s, uint8_t* d) {
        rect_morph_dev(h, st, s, H, W, kw, kh, erode != 0, d);
    });
}

// ---- detector ---------------------------------------------------------------------------------------------------------
int bbocr_craft_forward(bbocr_handle* h, const uint8_t* img, int H, int W, int on_device, int canvas_size,
                        double mag_ratio, float* score_text, float* score_link, int* mapH, int* mapW, double* ratio) {
    return guarded(h, [&] {
        ARG_CHECK(H > 0 && W > 0 && canvas_size >= 32, "bad geometry");
        CanvasGeom g = canvas_geom(H, W, canvas_size, mag_ratio);
        if (mapH) *mapH = g.H32 / 2;
        if (mapW) *mapW = g.W32 / 2;
        if (ratio) *ratio = g.ratio;
        if (!score_text || !score_link) return;          // size query
        ARG_CHECK(img, "null image");
        Lane& lane = h->lanes[0];
        DevBuf din;
        const uint8_t* src = img;
        if (!on_device) { upload(lane, din, img, (size_t)H * W * 3); src = din.as<uint8_t>(); }
        size_t n = (size_t)(g.H32 / 2) * (g.W32 / 2);
        DevBuf maps(n * 8, lane.stream);
        craft_forward_dev(h, lane.stream, src, g, maps.as<float>(), maps.as<float>() + n);
        std::vector<float> tmp(n * 2);
        download(lane, tmp.data(), maps.p, n * 8);
        memcpy(score_text, tmp.data(), n * 4);
        memcpy(score_link, tmp.data() + n, n * 4);
    });
}

// test hook (not in include/bbocr.h): host_path != 0 runs the round-1 tail (row extents to the host, hull + calipers in
// boxes.cpp) so that the device tail can be diffed against it; *used_device reports which one produced the boxes
int bbocr_dbg_det_boxes(bbocr_handle* h, const float* textmap, const float* linkmap, int mapH, int mapW, double text_threshold,
                        double link_threshold, double low_text, int host_path, float* boxes, int cap, int* n, int* used_device) {
    return guarded(h, [&] {
        ARG_CHECK(textmap && linkmap && boxes && n && mapH > 0 && mapW > 0, "bad arguments");
        Lane& lane = h->lanes[0];
        size_t np = (size_t)mapH * mapW;
        DevBuf dt, dl;
        upload(lane, dt, textmap, np * 4);
        CUDA_CHECK(stream_sync(lane.stream));      // pin_in is reused by the next upload
        upload(lane, dl, linkmap, np * 4);
        std::vector<float> b;
        static const bool env_host = getenv("BBOCR_HOST_BOXES") != nullptr;        // A/B: hull / calipers on the host
        bool dev = false;
        if (!host_path && !env_host)
            dev = det_boxes_dev(h, lane, dt.as<float>(), dl.as<float>(), mapH, mapW, (float)text_threshold, (float)link_threshold,
                                (float)low_text, b, nullptr);
        if (!dev) {
            DetComponents dc;
            det_components_dev(h, lane, dt.as<float>(), dl.as<float>(), mapH, mapW, (float)text_threshold,
                               (float)link_threshold, (float)low_text, dc);
            boxes_from_components(dc, mapH, mapW, b);
        }
        if (used_device) *used_device = dev ? 1 : 0;
        int nb = (int)b.size() / 8;
        ARG_CHECK(nb <= cap, "box capacity %d too small for %d boxes", cap, nb);
        memcpy(boxes, b.data(), b.size() * 4);
        *n = nb;
    });
}

int bbocr_det_boxes(bbocr_handle* h, const float* textmap, const float* linkmap, int mapH, int mapW,
                    double text_threshold, double link_threshold, double low_text, float* boxes, int cap, int* n) {
    return bbocr_dbg_det_boxes(h, textmap, linkmap, mapH, mapW, text_threshold, link_threshold, low_text, 0, boxes, cap, n, nullptr);
}

int bbocr_min_area_box(const int32_t* xy, int npoints, float* out8) {
    if (!xy || !out8 || npoints <= 0) return BBOCR_E_ARG;
    try {
        min_area_box(xy, npoints, out8);
    } catch (...) {
        return BBOCR_E_ARG;
    }
    return BBOCR_OK;
}

// test hook: the per-angle projection scores of bbocr_pp_deskew computed by one of its three histogram kernels
// (0 = one global atomic per ink pixel and angle, 1 = banded shared-memory counts, 2 = run-based prefix differences)
int bbocr_dbg_deskew_scores(bbocr_handle* h, const uint8_t* src, int H, int W, float max_deg, int variant, uint64_t* scores,
                            int cap, int* n_out) {
    return pp_step(h, src, (size_t)H * W, const_cast<uint8_t*>(src), (size_t)H * W, [&](cudaStream_t st, const uint8_t* s, uint8_t* d) {
        std::vector<unsigned long long> sc;
        pp_deskew(h, st, s, d, H, W, max_deg, variant, &sc);
        if (n_out) *n_out = (int)sc.size();
        for (int i = 0; i < (int)sc.size() && i < cap; ++i) scores[i] = sc[i];
    });
}

int bbocr_dbg_convex_hull(const int32_t* xy, int n, int clockwise, int32_t* out, int* nout) {
    std::vector<int> hull;
    debug_convex_hull(xy, n, clockwise, hull);
    for (size_t i = 0; i < hull.size(); ++i) out[i] = hull[i];
    *nout = (int)hull.size();
    return 0;
}

// test hook (not part of include/bbocr.h): one convolution on host FP32 NHWC data in the handle's precision mode.
// force_generic != 0 routes BF16 mode through the CUDA-core kernel so the tcgen05 kernel can be compared against it.
int bbocr_dbg_conv(bbocr_handle* h, const float* in1, int C1, const float* in2, int C2, int N, int H, int W, const float* w,
                   const float* bias, int cout, int kh, int kw, int pad, int dil, int relu, int force_generic, float* out) {
    return guarded(h, [&] {
        Lane& lane = h->lanes[0];
        cudaStream_t st = lane.stream;
        ConvW cw = make_conv_raw(h, w, bias, cout, C1 + C2, kh, kw, pad, dil);
        Act a1, a2, o;
        DevBuf f1, f2, b1, b2, bo, fo;
        int64_t n1 = (int64_t)N * H * W * C1, n2 = (int64_t)N * H * W * C2;
        upload(lane, f1, in1, n1 * 4);
        const bool split = force_generic == 2;          // 2: split-precision (hi/lo) input, FP32 output, tensor cores
        if (split) {
            Act src;
            src.N = N; src.H = H; src.W = W; src.C = C1; src.p = f1.p;
            a1 = act_alloc_split(h, st, b1, N, H, W, C1);
            maxpool_f32_to_split(h, st, src, a1, 1, 1);
        } else {
            a1 = act_alloc(h, st, b1, N, H, W, C1);
            act_from_f32(h, st, f1.as<float>(), a1.p, n1);
        }
        if (C2 > 0) {
            upload(lane, f2, in2, n2 * 4);
            a2 = act_alloc(h, st, b2, N, H, W, C2);
            act_from_f32(h, st, f2.as<float>(), a2.p, n2);
        }
        int OH = H + 2 * pad - dil * (kh - 1), OW = W + 2 * pad - dil * (kw - 1);
        o = act_alloc(h, st, bo, N, OH, OW, cout, split);
        h->force_generic_conv = force_generic == 1;
        g_conv_scope = 1;                                // timed like a detector convolution (tools/layer_times.py)
        try {
            conv_forward(h, st, cw, a1, a2, o, (relu ? CONV_RELU : 0) | (split ? CONV_OUT_F32 : 0));
        } catch (...) {
            h->force_generic_conv = false;
            g_conv_scope = 0;
            throw;
        }
        h->force_generic_conv = false;
        g_conv_scope = 0;
        int64_t no = (int64_t)N * OH * OW * cout;
        if (split) {
            download(lane, out, o.p, no * 4);
        } else {
            fo.alloc(no * 4, st);
            act_to_f32(h, st, o.p, fo.as<float>(), no);
            download(lane, out, fo.p, no * 4);
        }
    });
}

int bbocr_group_boxes(const float* boxes, int n, double ratio, const bbocr_group_params* p, int32_t* hlist, int* nh,
                      double* flist, int* nf, int cap) {
    if (!p || !hlist || !nh || !flist || !nf || n < 0 || (n > 0 && !boxes)) return BBOCR_E_ARG;
    try {
        std::vector<int32_t> hl;
        std::vector<double> fl;
        group_boxes(boxes, n, ratio, *p, hl, fl);
        if ((int)hl.size() / 4 > cap || (int)fl.size() / 8 > cap) return BBOCR_E_ARG;
        memcpy(hlist, hl.data(), hl.size() * 4);
        memcpy(flist, fl.data(), fl.size() * 8);
        *nh = (int)hl.size() / 4;
        *nf = (int)fl.size() / 8;
    } catch (...) {
        return BBOCR_E_ARG;
    }
    return BBOCR_OK;
}

void bbocr_default_params(bbocr_params* p) {
    if (!p) return;
    p->min_size = 20;
    p->canvas_size = 2560;
    p->contrast_ths = 0.1;
    p->adjust_contrast = 0.5;
    p->text_threshold = 0.7;
    p->low_text = 0.4;
    p->link_threshold = 0.4;
    p->mag_ratio = 1.0;
    p->slope_ths = 0.1;
    p->ycenter_ths = 0.5;
    p->height_ths = 0.5;
    p->width_ths = 0.5;
    p->add_margin = 0.1;
    p->ignore = nullptr;
    p->decoder = 0;
    p->beam_width = 5;
    p->batch_mode = 0;
    p->n_rotations = 0;
    p->rotation[0] = p->rotation[1] = p->rotation[2] = 0;
    p->space_idx = 43;
}

}  // extern "C"

// ======================================================================================================================
// recognition orchestration
// ======================================================================================================================
namespace {

struct CropJob {
    CropDesc d;
    bool is_free = false;
    double box[8];          // result box (clamped ints for horizontal boxes, the free quad otherwise)
    bool tall = false;
    int base = -1;          // rotation_info: index of the unrotated job this one is a rotated copy of (-1 = it is one itself)
};

// AlignCollate geometry of a (oh x ow) crop for model width model_w: resized_w = min(ceil(64 * w / h), imgW); `tall` = the
// PIL BICUBIC resize is not the identity
void align_geometry(CropJob& j) {
    const double r2 = (double)j.d.ow / (double)j.d.oh;
    const int rw = (int)std::ceil(64 * r2);
    j.d.resized_w = rw > j.d.model_w ? j.d.model_w : rw;
    j.tall = !(j.d.oh == 64 && j.d.resized_w == j.d.ow);
}

// utils.get_image_list geometry for one horizontal box [xmin,xmax,ymin,ymax]
bool horizontal_job(const int32_t* b, int H, int W, CropJob& j) {
    int x_min = std::max(0, b[0]), x_max = std::min(b[1], W), y_min = std::max(0, b[2]), y_max = std::min(b[3], H);
    int width = x_max - x_min, height = y_max - y_min;
    if (width <= 0 || height <= 0) fail(BBOCR_E_ARG, "degenerate text box %dx%d (upstream raises here as well)", width, height);
    j.d.x0 = x_min; j.d.y0 = y_min; j.d.w = width; j.d.h = height;
    j.d.free_idx = -1;
    j.is_free = false;
    double q[8] = {(double)x_min, (double)y_min, (double)x_max, (double)y_min, (double)x_max, (double)y_max, (double)x_min, (double)y_max};
    memcpy(j.box, q, sizeof q);
    return true;
}

// calculate_ratio + compute_ratio_and_resize + AlignCollate geometry; returns false when upstream skips the box
bool finish_geometry(CropJob& j) {
    const int width = j.d.w, height = j.d.h;
    double ratio = (double)width / (double)height;
    double cr = ratio < 1.0 ? 1.0 / ratio : ratio;                 // calculate_ratio
    int new_width = (int)(64 * cr);
    if (new_width == 0) return false;
    if (ratio < 1.0) { j.d.ow = 64; j.d.oh = (int)(64 * cr); j.tall = true; }
    else { j.d.ow = (int)(64 * ratio); j.d.oh = 64; j.tall = false; }
    double max_ratio = std::max(cr, 1.0);
    j.d.model_w = (int)std::ceil(std::ceil(max_ratio)) * 64;
    // AlignCollate: ratio = w / float(h); resized_w = min(ceil(64 * ratio), imgW)
    double r2 = (double)j.d.ow / (double)j.d.oh;
    int rw = (int)std::ceil(64 * r2);
    j.d.resized_w = rw > j.d.model_w ? j.d.model_w : rw;
    if (j.d.oh == 64 && j.d.resized_w == j.d.ow) j.tall = false;    // identity resize
    return true;
}

// np.percentile(img, q) (method='linear') from a 256-bin histogram, replicating NumPy's float64 arithmetic
double percentile_from_hist(const unsigned int* hist, int64_t n, double q100) {
    double q = q100 / 100.0;
    double vi = (double)n * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;
    double prev_f = std::floor(vi);
    int64_t prev = (int64_t)prev_f, next = prev + 1;
    double gamma = vi - prev_f;
    prev = std::min<int64_t>(std::max<int64_t>(prev, 0), n - 1);
    next = std::min<int64_t>(std::max<int64_t>(next, 0), n - 1);
    auto value_at = [&](int64_t k) {
        int64_t c = 0;
        for (int v = 0; v < 256; ++v) {
            c += hist[v];
            if (k < c) return (double)v;
        }
        return 255.0;
    };
    double a = value_at(prev), b = value_at(next);
    double diff = b - a;
    double r = a + diff * gamma;
    if (gamma >= 0.5) r = b - diff * (1 - gamma);
    return r;
}

struct Recognized {
    std::vector<int32_t> text;
    double conf = 0;
};

// custom_mean over the max probabilities of the non-blank timesteps (float32 running product, float64 pow)
double confidence_of(const float* prob, const int32_t* idx, int T) {
    float prod = 1.f;
    int cnt = 0;
    for (int t = 0; t < T; ++t)
        if (idx[t] != 0) { prod = prod * prob[t]; ++cnt; }
    if (cnt == 0) return 0.0;
    return std::pow((double)prod, 2.0 / std::sqrt((double)cnt));
}

// decoder = 'beamsearch' / 'wordbeamsearch' (recognizer_predict): the strings come from the host beam search over the
// probability matrix; the confidence stays the greedy path's custom_mean, as upstream computes it for every decoder.
void beam_decode_pass(Handle* h, Lane& lane, const float* logits_dev, int rows, int C, const uint8_t* ignore_dev,
                      const std::vector<SeqDesc>& where, const bbocr_params& p, std::vector<Recognized>& out) {
    if (p.decoder == 0 || rows == 0) return;
    ARG_CHECK(p.decoder == 1 || p.decoder == 2, "unknown decoder %d", p.decoder);
    ARG_CHECK(p.beam_width >= 1, "beamWidth must be positive");
    cudaStream_t st = lane.stream;
    DevBuf dprobs((size_t)rows * C * 4, st);
    row_probs_dev(h, st, logits_dev, rows, C, ignore_dev, dprobs.as<float>());
    std::vector<float> probs((size_t)rows * C);
    download(lane, probs.data(), dprobs.p, probs.size() * 4);
    const int space_idx = p.space_idx;                          // CTCLabelConverter.dict[' ']
    const int n = (int)where.size();
    const int nt = std::max(1, std::min<int>(8, n / 4));
    std::atomic<int> next{0};
    auto work = [&] {
        for (int k; (k = next.fetch_add(1)) < n;)
            decode_beam(probs.data() + (size_t)where[k].row0 * C, where[k].T, C, p.decoder, p.beam_width, space_idx, &h->dict, out[k].text);
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

// Throughput-mode variant of recognize_pass: the crops of the pass are laid side by side in strip images (at most
// `max_cols` columns each) and the feature extractor runs once per strip; the sequence half and the decoder run once.
void recognize_pass_ragged(Handle* h, Lane& lane, std::vector<CropJob>& jobs, const std::vector<int>& which, const uint8_t* crops,
                           size_t crops_bytes, const uint8_t* ignore_dev, const bbocr_params& p, std::vector<Recognized>& out) {
    cudaStream_t st = lane.stream;
    const int n = (int)which.size();
    out.assign(n, Recognized());
    if (n == 0) return;
    constexpr int GAP = 16;
    static const int max_cols = getenv("BBOCR_REC_COLS") ? std::max(4096, atoi(getenv("BBOCR_REC_COLS"))) : 65536;
    size_t aligned_extra = 0, scratch_max = 0;
    for (int i : which)
        if (jobs[i].tall) {
            aligned_extra += (size_t)64 * jobs[i].d.resized_w;
            scratch_max = std::max(scratch_max, (size_t)jobs[i].d.oh * std::max(jobs[i].d.resized_w, jobs[i].d.ow));
        }
    DevBuf aligned_buf(crops_bytes + aligned_extra + 16, st), pil_scratch(scratch_max + 16, st);
    CUDA_CHECK(cudaMemcpyAsync(aligned_buf.p, crops, crops_bytes, cudaMemcpyDeviceToDevice, st));
    size_t acur = crops_bytes;
    std::vector<CropDesc> descs(n);
    std::vector<SeqDesc> seqs(n);
    int rows = 0, max_w = 0;
    for (int k = 0; k < n; ++k) {
        CropJob& j = jobs[which[k]];
        if (j.tall) {
            pil_resize_bicubic_dev(h, st, aligned_buf.as<uint8_t>() + j.d.off, j.d.oh, j.d.ow,
                                   aligned_buf.as<uint8_t>() + acur, 64, j.d.resized_w, pil_scratch.as<uint8_t>());
            j.d.aoff = (int)acur;
            acur += (size_t)64 * j.d.resized_w;
        } else {
            j.d.aoff = j.d.off;
        }
        descs[k] = j.d;
        const int T = j.d.model_w / 4 - 1;
        seqs[k] = SeqDesc{rows, T};
        rows += T;
        max_w = std::max(max_w, j.d.model_w);
    }
    // strips: consecutive crops until the column budget is reached
    struct Strip { int first, count, Wtot, t_max; };
    std::vector<Strip> strips;
    std::vector<int> xoff(n);
    for (int k = 0; k < n;) {
        Strip s{k, 0, 0, 0};
        while (k < n && (s.count == 0 || s.Wtot + descs[k].model_w + GAP <= max_cols)) {
            xoff[k] = s.Wtot;
            s.Wtot += descs[k].model_w + GAP;
            s.t_max = std::max(s.t_max, descs[k].model_w / 4 - 1);
            ++s.count;
            ++k;
        }
        strips.push_back(s);
    }
    // one upload: descs | xoff | meta (col0, T, row0 per crop) | masks of every strip
    size_t mask_bytes = 0;
    for (auto& s : strips) mask_bytes += (size_t)s.Wtot / 2 + s.Wtot / 4;
    const size_t o_desc = 0, o_xoff = o_desc + (size_t)n * sizeof(CropDesc), o_meta = o_xoff + (size_t)n * 4,
                 o_seq = o_meta + (size_t)n * 12, o_mask = o_seq + (size_t)n * sizeof(SeqDesc),
                 total = ((o_mask + mask_bytes + 15) & ~(size_t)15);
    std::vector<uint8_t> blob(total, 0);
    memcpy(blob.data() + o_desc, descs.data(), (size_t)n * sizeof(CropDesc));
    memcpy(blob.data() + o_seq, seqs.data(), (size_t)n * sizeof(SeqDesc));
    memcpy(blob.data() + o_xoff, xoff.data(), (size_t)n * 4);
    int* meta = reinterpret_cast<int*>(blob.data() + o_meta);
    std::vector<size_t> strip_mask_off(strips.size());
    {
        size_t mo = o_mask;
        for (size_t q = 0; q < strips.size(); ++q) {
            const Strip& s = strips[q];
            strip_mask_off[q] = mo;
            uint8_t* m1 = blob.data() + mo;
            uint8_t* m2 = m1 + s.Wtot / 2;
            for (int k = s.first; k < s.first + s.count; ++k) {
                memset(m1 + xoff[k] / 2, 1, descs[k].model_w / 2);
                memset(m2 + xoff[k] / 4, 1, descs[k].model_w / 4);
                meta[3 * k] = xoff[k] / 4;
                meta[3 * k + 1] = seqs[k].T;
                meta[3 * k + 2] = seqs[k].row0;
            }
            mo += (size_t)s.Wtot / 2 + s.Wtot / 4;
        }
    }
    DevBuf dblob;
    upload(lane, dblob, blob.data(), blob.size());
    const uint8_t* db = dblob.as<uint8_t>();
    const CropDesc* ddesc = reinterpret_cast<const CropDesc*>(db + o_desc);
    const int* dxoff = reinterpret_cast<const int*>(db + o_xoff);
    const int* dmeta = reinterpret_cast<const int*>(db + o_meta);
    DevBuf seqbuf;
    Act seq = crnn_alloc_seq(h, st, seqbuf, rows);
    for (size_t q = 0; q < strips.size(); ++q) {
        const Strip& s = strips[q];
        DevBuf strip((size_t)64 * s.Wtot * 4, st);
        crops_to_strip_dev(h, st, aligned_buf.as<uint8_t>(), ddesc + s.first, dxoff + s.first, s.count, max_w, s.Wtot, strip.as<float>());
        crnn_features_strip_dev(h, st, strip.as<float>(), s.Wtot, db + strip_mask_off[q], db + strip_mask_off[q] + s.Wtot / 2,
                                dmeta + 3 * s.first, s.count, s.t_max, seq);
    }
    const int C = h->crnn.num_class;
    DevBuf logits((size_t)rows * C * 4, st);
    crnn_sequence_dev(h, lane, seq, seqs, logits.as<float>());
    const SeqDesc* dseq = reinterpret_cast<const SeqDesc*>(db + o_seq);
    const size_t step_elems = (size_t)rows;
    DevBuf dec((step_elems * 3 + (size_t)n) * 4, st);
    int32_t* text_idx = dec.as<int32_t>();
    int32_t* step_idx = text_idx + step_elems;
    float* step_prob = reinterpret_cast<float*>(step_idx + step_elems);
    int32_t* text_len = reinterpret_cast<int32_t*>(step_prob + step_elems);
    ctc_decode_dev(h, st, logits.as<float>(), rows, C, ignore_dev, dseq, n, text_idx, text_len, step_prob, step_idx);
    std::vector<int32_t> hdec(step_elems * 3 + n);
    download(lane, hdec.data(), dec.p, hdec.size() * 4);
    const int32_t* h_text = hdec.data();
    const int32_t* h_sidx = h_text + step_elems;
    const float* h_prob = reinterpret_cast<const float*>(h_sidx + step_elems);
    const int32_t* h_len = reinterpret_cast<const int32_t*>(h_prob + step_elems);
    for (int k = 0; k < n; ++k) {
        const size_t so = (size_t)seqs[k].row0;
        out[k].text.assign(h_text + so, h_text + so + h_len[k]);
        out[k].conf = confidence_of(h_prob + so, h_sidx + so, seqs[k].T);
    }
    beam_decode_pass(h, lane, logits.as<float>(), rows, C, ignore_dev, seqs, p, out);
}

// One pass of AlignCollate -> CRNN -> decode over `jobs` whose (possibly contrast-adjusted) crops live in `crops`.
void recognize_pass(Handle* h, Lane& lane, std::vector<CropJob>& jobs, const std::vector<int>& which, const uint8_t* crops,
                    size_t crops_bytes, const uint8_t* ignore_dev, const bbocr_params& p, std::vector<Recognized>& out) {
    if (crnn_ragged(h)) {
        recognize_pass_ragged(h, lane, jobs, which, crops, crops_bytes, ignore_dev, p, out);
        return;
    }
    cudaStream_t st = lane.stream;
    const int n = (int)which.size();
    out.assign(n, Recognized());
    if (n == 0) return;
    // aligned buffer: tall crops are Pillow-resized to 64 rows, the rest alias the crop buffer
    size_t aligned_extra = 0, scratch_max = 0;
    for (int i : which)
        if (jobs[i].tall) {
            aligned_extra += (size_t)64 * jobs[i].d.resized_w;
            scratch_max = std::max(scratch_max, (size_t)jobs[i].d.oh * std::max(jobs[i].d.resized_w, jobs[i].d.ow));
        }
    DevBuf aligned_buf(crops_bytes + aligned_extra + 16, st), pil_scratch(scratch_max + 16, st);
    CUDA_CHECK(cudaMemcpyAsync(aligned_buf.p, crops, crops_bytes, cudaMemcpyDeviceToDevice, st));
    size_t acur = crops_bytes;
    // buckets by model width, in first-appearance order
    std::vector<int> bucket_w, bucket_n;
    std::vector<CropDesc> descs(n);
    for (int k = 0; k < n; ++k) {
        CropJob& j = jobs[which[k]];
        if (j.tall) {
            pil_resize_bicubic_dev(h, st, aligned_buf.as<uint8_t>() + j.d.off, j.d.oh, j.d.ow,
                                   aligned_buf.as<uint8_t>() + acur, 64, j.d.resized_w, pil_scratch.as<uint8_t>());
            j.d.aoff = (int)acur;
            acur += (size_t)64 * j.d.resized_w;
        } else {
            j.d.aoff = j.d.off;
        }
        int b = -1;
        for (size_t q = 0; q < bucket_w.size(); ++q)
            if (bucket_w[q] == j.d.model_w) b = (int)q;
        if (b < 0) { b = (int)bucket_w.size(); bucket_w.push_back(j.d.model_w); bucket_n.push_back(0); }
        j.d.slot = bucket_n[b]++;
        j.d.bucket_off = b;                                       // bucket index for now
    }
    std::vector<size_t> bucket_off(bucket_w.size());
    size_t in_floats = 0, logit_floats = 0, step_elems = 0;
    std::vector<size_t> logit_off(bucket_w.size()), step_off(bucket_w.size());
    const int C = h->crnn.num_class;
    int max_w = 0;
    for (size_t q = 0; q < bucket_w.size(); ++q) {
        bucket_off[q] = in_floats;
        in_floats += (size_t)bucket_n[q] * 64 * bucket_w[q];
        int T = bucket_w[q] / 4 - 1;
        logit_off[q] = logit_floats;
        logit_floats += (size_t)bucket_n[q] * T * C;
        step_off[q] = step_elems;
        step_elems += (size_t)bucket_n[q] * T;
        max_w = std::max(max_w, bucket_w[q]);
    }
    std::vector<int> job_bucket(n);
    for (int k = 0; k < n; ++k) {
        CropJob& j = jobs[which[k]];
        job_bucket[k] = j.d.bucket_off;
        j.d.bucket_off = (int)bucket_off[job_bucket[k]];
        descs[k] = j.d;
    }
    DevBuf ddesc, inputs(in_floats * 4, st), logits(logit_floats * 4, st);
    upload(lane, ddesc, descs.data(), descs.size() * sizeof(CropDesc));
    crops_to_input_dev(h, st, aligned_buf.as<uint8_t>(), ddesc.as<CropDesc>(), n, max_w, inputs.as<float>());
    // feature extractor per width bucket -> one flat [rows][256] sequence tensor; sequence k = (bucket, slot)
    const int rows = (int)step_elems;
    DevBuf seqbuf;
    Act seq = crnn_alloc_seq(h, st, seqbuf, rows);
    std::vector<SeqDesc> seqs;
    std::vector<int> seq_of_bucket_slot0(bucket_w.size());
    for (size_t q = 0; q < bucket_w.size(); ++q) {
        int T = bucket_w[q] / 4 - 1;
        crnn_features_dev(h, st, inputs.as<float>() + bucket_off[q], bucket_n[q], bucket_w[q], seq, (int)step_off[q]);
        seq_of_bucket_slot0[q] = (int)seqs.size();
        for (int sl = 0; sl < bucket_n[q]; ++sl) seqs.push_back(SeqDesc{(int)step_off[q] + sl * T, T});
    }
    // recurrent half + Prediction once over all crops, then greedy CTC over all rows
    crnn_sequence_dev(h, lane, seq, seqs, logits.as<float>());
    const int n_seq = (int)seqs.size();
    DevBuf dseq;
    upload(lane, dseq, seqs.data(), seqs.size() * sizeof(SeqDesc));
    // decode outputs: text_idx | step_idx (int32, rows each) | step_prob (float, rows) | text_len (n_seq)
    DevBuf dec((step_elems * 3 + (size_t)n_seq) * 4, st);
    int32_t* text_idx = dec.as<int32_t>();
    int32_t* step_idx = text_idx + step_elems;
    float* step_prob = reinterpret_cast<float*>(step_idx + step_elems);
    int32_t* text_len = reinterpret_cast<int32_t*>(step_prob + step_elems);
    ctc_decode_dev(h, st, logits.as<float>(), rows, C, ignore_dev, dseq.as<SeqDesc>(), n_seq, text_idx, text_len, step_prob,
                   step_idx);
    std::vector<int32_t> hdec(step_elems * 3 + n_seq);
    download(lane, hdec.data(), dec.p, hdec.size() * 4);
    const int32_t* h_text = hdec.data();
    const int32_t* h_sidx = h_text + step_elems;
    const float* h_prob = reinterpret_cast<const float*>(h_sidx + step_elems);
    const int32_t* h_len = reinterpret_cast<const int32_t*>(h_prob + step_elems);
    std::vector<SeqDesc> where(n);
    for (int k = 0; k < n; ++k) {
        const CropJob& j = jobs[which[k]];
        int q = job_bucket[k], T = bucket_w[q] / 4 - 1;
        size_t so = step_off[q] + (size_t)j.d.slot * T;
        int len = h_len[seq_of_bucket_slot0[q] + j.d.slot];
        out[k].text.assign(h_text + so, h_text + so + len);
        out[k].conf = confidence_of(h_prob + so, h_sidx + so, T);
        where[k] = SeqDesc{(int)so, T};
    }
    beam_decode_pass(h, lane, logits.as<float>(), rows, C, ignore_dev, where, p, out);
}

// ---- readtext = detect (per page) -> recognize (per GROUP of pages) -> assemble (per page) ---------------------------------
// Pages are independent, and with batch_size == 1 semantics every crop has its own model width, so the recogniser's result
// for a crop does not depend on which other crops share its launch.  The detector half therefore runs page by page on the
// detector lanes, while the recogniser half runs once over the crops of a whole group of pages: a few large launches
// (and full 128-crop LSTM clusters) instead of ~120 tiny ones per page.
struct PageWork {
    int index = 0;
    int H = 0, W = 0;
    int n_labels = 0;
    std::vector<CropJob> jobs;
    DevBuf dcrops;                 // packed u8 crops of this page (device)
    size_t crops_bytes = 0;
    std::vector<Recognized> rec;   // final per-job result
    int n_crops_run = 0;
};

// utils.get_image_list for the boxes of one page (batch_size == 1 semantics: every box has its own max_width): crop jobs in
// upstream order (horizontal list, then free list) and the packed u8 crops on the device
void build_crops(Handle* h, Lane& lane, const uint8_t* gray, int H, int W, const std::vector<int32_t>& hlist,
                 const std::vector<double>& flist, const bbocr_params& p, PageWork& pw) {
    cudaStream_t st = lane.stream;
    std::vector<CropJob>& jobs = pw.jobs;
    std::vector<double> mats;
    size_t scratch_bytes = 0;
    ARG_CHECK(p.n_rotations >= 0 && p.n_rotations <= 3, "rotation_info: at most three orientations");
    for (int r = 0; r < p.n_rotations; ++r)
        ARG_CHECK(p.rotation[r] == 90 || p.rotation[r] == 180 || p.rotation[r] == 270, "rotation_info: eligible values are 90, 180 and 270");
    // upstream's batched branch of Reader.recognize (batch_size > 1 on a GPU reader, or rotation_info): get_image_list over
    // the whole page = free boxes first, one max_width, crops stably sorted by the y of their first corner
    const bool batch_mode = p.batch_mode != 0 || p.n_rotations > 0;
    auto add_horizontal = [&] {
        for (size_t i = 0; i + 4 <= hlist.size(); i += 4) {
            CropJob j;
            horizontal_job(&hlist[i], H, W, j);
            if (finish_geometry(j)) jobs.push_back(j);
        }
    };
    if (!batch_mode) add_horizontal();
    for (size_t i = 0; i + 8 <= flist.size(); i += 8) {
        CropJob j;
        int mwid, mhei;
        double M[9];
        free_box_transform(&flist[i], &mwid, &mhei, M);
        if (mwid <= 0 || mhei <= 0) fail(BBOCR_E_ARG, "degenerate free-form box");
        j.is_free = true;
        memcpy(j.box, &flist[i], 64);
        j.d.free_idx = (int)mats.size() / 9;
        mats.insert(mats.end(), M, M + 9);
        j.d.x0 = (int)scratch_bytes; j.d.y0 = 0; j.d.w = mwid; j.d.h = mhei;
        scratch_bytes += (size_t)mwid * mhei;
        if (finish_geometry(j)) jobs.push_back(j);
    }
    if (batch_mode) {
        add_horizontal();
        int shared_w = 64;
        for (auto& j : jobs) shared_w = std::max(shared_w, j.d.model_w);
        std::stable_sort(jobs.begin(), jobs.end(), [](const CropJob& a, const CropJob& b) { return a.box[1] < b.box[1]; });
        for (auto& j : jobs) { j.d.model_w = shared_w; align_geometry(j); }
    }
    const int n = (int)jobs.size();
    if (n > 0) {
        size_t crops_bytes = 0;
        for (auto& j : jobs) { j.d.off = (int)crops_bytes; crops_bytes += (size_t)j.d.ow * j.d.oh; }
        const size_t base_bytes = crops_bytes;
        // rotation_info (utils.make_rotated_img_list): for every angle a rotated copy of every crop, appended in that order
        std::vector<RotDesc> rots;
        int max_pixels = 0;
        for (int r = 0; r < p.n_rotations; ++r)
            for (int i = 0; i < n; ++i) {
                CropJob j = jobs[i];
                const int k = p.rotation[r] / 90;
                rots.push_back(RotDesc{jobs[i].d.off, (int)crops_bytes, jobs[i].d.oh, jobs[i].d.ow, k});
                max_pixels = std::max(max_pixels, jobs[i].d.oh * jobs[i].d.ow);
                if (k != 2) std::swap(j.d.ow, j.d.oh);
                j.d.off = (int)crops_bytes;
                j.base = i;
                align_geometry(j);
                crops_bytes += (size_t)j.d.ow * j.d.oh;
                jobs.push_back(j);
            }
        pw.crops_bytes = crops_bytes;
        std::vector<CropDesc> descs(n);
        for (int i = 0; i < n; ++i) descs[i] = jobs[i].d;
        DevBuf ddesc, dmats, drots, dscratch(scratch_bytes + 16, st);
        pw.dcrops.alloc(crops_bytes + 16, st);
        upload(lane, ddesc, descs.data(), descs.size() * sizeof(CropDesc));
        if (!mats.empty()) {
            CUDA_CHECK(stream_sync(st));
            upload(lane, dmats, mats.data(), mats.size() * 8);
        }
        crops_dev(h, st, gray, H, W, ddesc.as<CropDesc>(), n, descs.data(), dmats.as<double>(), dscratch.as<uint8_t>(),
                  pw.dcrops.as<uint8_t>());
        if (!rots.empty()) {
            CUDA_CHECK(stream_sync(st));
            upload(lane, drots, rots.data(), rots.size() * sizeof(RotDesc));
            rotate_crops_dev(h, st, pw.dcrops.as<uint8_t>(), drots.as<RotDesc>(), (int)rots.size(), max_pixels);
        }
        (void)base_bytes;
    }
    pw.rec.assign(jobs.size(), Recognized());
}

// Reader.detect + utils.get_image_list for one page on one lane; leaves the page's crops on the device
// k pages of identical size on one lane: the detector network runs them as ONE batch (every layer one launch), the
// per-page part (components, boxes, crops) follows page by page.
void detect_pages(Handle* h, Lane& lane, const bbocr_image* const* imgs, int k, const bbocr_params& p, PageWork* const* pws) {
    cudaStream_t st = lane.stream;
    if (!h->craft_loaded || !h->crnn_loaded) fail(BBOCR_E_STATE, "weights not loaded");
    ARG_CHECK(k >= 1 && k <= 16, "bad detector batch");
    const int H = imgs[0]->H, W = imgs[0]->W;
    std::vector<DevBuf> dcolor(k), dgray(k);
    std::vector<const uint8_t*> color(k), gray(k);
    // host pages: one pinned staging area for the whole batch (no stream sync between the pages' uploads)
    size_t pin_total = 0;
    std::vector<size_t> pin_off(k, 0);
    for (int i = 0; i < k; ++i) {
        const bbocr_image& img = *imgs[i];
        ARG_CHECK(img.color && img.H == H && img.W == W && H > 0 && W > 0, "bad image");
        if (!img.on_device) {
            pin_off[i] = pin_total;
            pin_total += (((size_t)H * W * (img.gray ? 4 : 3)) + 255) & ~(size_t)255;
        }
    }
    uint8_t* pin_base = pin_total ? (uint8_t*)staging(lane, pin_total) : nullptr;
    for (int i = 0; i < k; ++i) {
        const bbocr_image& img = *imgs[i];
        pws[i]->H = H; pws[i]->W = W;
        color[i] = img.color;
        gray[i] = img.gray;
        if (!img.on_device) {
            size_t cb = (size_t)H * W * 3, gb = img.gray ? (size_t)H * W : 0;
            uint8_t* pin = pin_base + pin_off[i];
            memcpy(pin, img.color, cb);
            if (gb) memcpy(pin + cb, img.gray, gb);
            dcolor[i].alloc(cb + gb, st);
            CUDA_CHECK(cudaMemcpyAsync(dcolor[i].p, pin, cb + gb, cudaMemcpyHostToDevice, st));
            lane.in_busy = true;
            color[i] = dcolor[i].as<uint8_t>();
            gray[i] = gb ? dcolor[i].as<uint8_t>() + cb : nullptr;
        }
        if (!gray[i]) {                                   // reformat_input: cv2.cvtColor(image, COLOR_BGR2GRAY)
            dgray[i].alloc((size_t)H * W, st);
            pp_gray(h, st, color[i], H, W, W * 3, dgray[i].as<uint8_t>());
            gray[i] = dgray[i].as<uint8_t>();
        }
    }
    // ---- detect -------------------------------------------------------------------------------------------------
    std::unique_ptr<StageTimer> tm(new StageTimer(h, 0));
    CanvasGeom g = canvas_geom(H, W, p.canvas_size, p.mag_ratio);
    const int mh = g.H32 / 2, mw = g.W32 / 2;
    const size_t plane = (size_t)mh * mw;
    DevBuf maps((size_t)k * plane * 8, st);
    float* text = maps.as<float>();
    float* link = text + (size_t)k * plane;
    craft_forward_batch_dev(h, st, color.data(), k, g, text, link);
    for (int i = 0; i < k; ++i) {
        tm.reset(new StageTimer(h, 1));
        std::vector<float> boxes;
        static const bool host_boxes = getenv("BBOCR_HOST_BOXES") != nullptr;      // A/B: hull / calipers on the host (round 1)
        int n_labels = 0;
        if (host_boxes || !det_boxes_dev(h, lane, text + i * plane, link + i * plane, mh, mw, (float)p.text_threshold,
                                         (float)p.link_threshold, (float)p.low_text, boxes, &n_labels)) {
            DetComponents dc;
            det_components_dev(h, lane, text + i * plane, link + i * plane, mh, mw, (float)p.text_threshold, (float)p.link_threshold,
                               (float)p.low_text, dc);
            n_labels = dc.n_labels;
            boxes_from_components(dc, mh, mw, boxes);
        }
        pws[i]->n_labels = n_labels;
        tm.reset(new StageTimer(h, 2));
        bbocr_group_params gp{p.slope_ths, p.ycenter_ths, p.height_ths, p.width_ths, p.add_margin, p.min_size};
        std::vector<int32_t> hlist;
        std::vector<double> flist;
        group_boxes(boxes.data(), (int)boxes.size() / 8, g.ratio, gp, hlist, flist);
        build_crops(h, lane, gray[i], H, W, hlist, flist, p, *pws[i]);
    }
    tm.reset();
    CUDA_CHECK(stream_sync(st));                 // the pages' crops are complete; pinned staging is free again
    lane.in_busy = false;
}

void detect_page(Handle* h, Lane& lane, const bbocr_image& img, const bbocr_params& p, PageWork& pw) {
    const bbocr_image* ip = &img;
    PageWork* pp = &pw;
    detect_pages(h, lane, &ip, 1, p, &pp);
}

// Reader.recognize over the crops of a group of pages (pass 1, then the contrast-retry pass for low-confidence crops)
void recognize_group(Handle* h, Lane& lane, const std::vector<PageWork*>& pages, const bbocr_params& p) {
    cudaStream_t st = lane.stream;
    // union of the pages' jobs; crop offsets rebased into one packed buffer
    std::vector<CropJob> jobs;
    std::vector<std::pair<int, int>> origin;       // (page slot, job index)
    size_t crops_bytes = 0;
    std::vector<size_t> page_base(pages.size());
    for (size_t q = 0; q < pages.size(); ++q) {
        page_base[q] = crops_bytes;
        for (size_t i = 0; i < pages[q]->jobs.size(); ++i) {
            CropJob j = pages[q]->jobs[i];
            j.d.off += (int)crops_bytes;
            jobs.push_back(j);
            origin.emplace_back((int)q, (int)i);
        }
        crops_bytes += pages[q]->crops_bytes;
    }
    const int n = (int)jobs.size();
    if (n == 0) return;
    ARG_CHECK(crops_bytes < (size_t)INT32_MAX, "crop batch too large");
    DevBuf dignore;
    const uint8_t* ignore_dev = nullptr;
    if (p.ignore) {
        upload(lane, dignore, p.ignore, h->crnn.num_class);
        CUDA_CHECK(stream_sync(st));
        lane.in_busy = false;
        ignore_dev = dignore.as<uint8_t>();
    }
    DevBuf dcrops(crops_bytes + 16, st);
    for (size_t q = 0; q < pages.size(); ++q)
        if (pages[q]->crops_bytes)
            CUDA_CHECK(cudaMemcpyAsync(dcrops.as<uint8_t>() + page_base[q], pages[q]->dcrops.p, pages[q]->crops_bytes,
                                       cudaMemcpyDeviceToDevice, st));
    std::vector<int> all(n);
    for (int i = 0; i < n; ++i) all[i] = i;
    std::unique_ptr<StageTimer> tm(new StageTimer(h, 3));
    std::vector<Recognized> rec1;
    recognize_pass(h, lane, jobs, all, dcrops.as<uint8_t>(), crops_bytes, ignore_dev, p, rec1);
    std::vector<int> low;
    for (int i = 0; i < n; ++i)
        if (rec1[i].conf < p.contrast_ths) low.push_back(i);
    std::vector<Recognized> final_rec = rec1;
    tm.reset(new StageTimer(h, 4));
    if (!low.empty()) {
        // second round: adjust_contrast_grey(target = adjust_contrast) on the low-confidence crops
        const int nl = (int)low.size();
        std::vector<CropDesc> ldesc(nl);
        for (int k = 0; k < nl; ++k) ldesc[k] = jobs[low[k]].d;
        DevBuf dl, dhist((size_t)nl * 256 * 4, st), dadj(crops_bytes + 16, st);
        CUDA_CHECK(stream_sync(st));
        upload(lane, dl, ldesc.data(), ldesc.size() * sizeof(CropDesc));
        crop_hist_dev(h, st, dcrops.as<uint8_t>(), dl.as<CropDesc>(), nl, dhist.as<unsigned int>());
        std::vector<unsigned int> hist((size_t)nl * 256);
        download(lane, hist.data(), dhist.p, hist.size() * 4);
        std::vector<double> par((size_t)nl * 2);
        std::vector<int> apply(nl);
        for (int k = 0; k < nl; ++k) {
            int64_t np = (int64_t)ldesc[k].ow * ldesc[k].oh;
            double high = percentile_from_hist(&hist[(size_t)k * 256], np, 90.0);
            double lo = percentile_from_hist(&hist[(size_t)k * 256], np, 10.0);
            double contrast = (high - lo) / std::max(10.0, high + lo);
            apply[k] = contrast < p.adjust_contrast;
            par[k] = lo;
            par[(size_t)nl + k] = 200.0 / std::max(10.0, high - lo);
        }
        DevBuf dpar, dapply;
        upload(lane, dpar, par.data(), par.size() * 8);
        CUDA_CHECK(stream_sync(st));
        upload(lane, dapply, apply.data(), (size_t)nl * 4);
        crop_contrast_dev(h, st, dcrops.as<uint8_t>(), dl.as<CropDesc>(), nl, dpar.as<double>(), dpar.as<double>() + nl,
                          dapply.as<int>(), dadj.as<uint8_t>());
        CUDA_CHECK(stream_sync(st));
        std::vector<Recognized> rec2;
        recognize_pass(h, lane, jobs, low, dadj.as<uint8_t>(), crops_bytes, ignore_dev, p, rec2);
        for (int k = 0; k < nl; ++k) {
            pages[origin[low[k]].first]->n_crops_run += 1;
            if (!(rec1[low[k]].conf > rec2[k].conf)) final_rec[low[k]] = rec2[k];
        }
    }
    tm.reset();
    for (int i = 0; i < n; ++i) {
        PageWork* pw = pages[origin[i].first];
        pw->rec[origin[i].second] = std::move(final_rec[i]);
        pw->n_crops_run += 1;
    }
}

// result assembly (horizontal boxes in group order, then free boxes)
bbocr_results* assemble_page(PageWork& pw) {
    // rotation_info (utils.set_result_with_confidence): every box keeps the orientation with the highest confidence, the
    // first one on ties (Python's max); the rotated copies follow the n base jobs angle by angle
    int n = (int)pw.jobs.size();
    for (int i = 0; i < (int)pw.jobs.size(); ++i)
        if (pw.jobs[i].base >= 0) { n = i; break; }
    for (int i = n; i < (int)pw.jobs.size(); ++i) {
        const int b = pw.jobs[i].base;
        if (pw.rec[i].conf > pw.rec[b].conf) pw.rec[b] = pw.rec[i];
    }
    bbocr_results* r = new bbocr_results();
    memset(r, 0, sizeof *r);
    r->n_components = pw.n_labels;
    r->n_crops = pw.n_crops_run;
    r->n = n;
    r->box = new double[(size_t)std::max(n, 1) * 8];
    r->is_free = new uint8_t[std::max(n, 1)];
    r->text_off = new int32_t[n + 1];
    r->conf = new double[std::max(n, 1)];
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += pw.rec[i].text.size();
    r->text_idx = new int32_t[std::max<size_t>(total, 1)];
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
        memcpy(r->box + (size_t)i * 8, pw.jobs[i].box, 64);
        r->is_free[i] = pw.jobs[i].is_free;
        r->text_off[i] = (int32_t)off;
        memcpy(r->text_idx + off, pw.rec[i].text.data(), pw.rec[i].text.size() * 4);
        off += pw.rec[i].text.size();
        r->conf[i] = pw.rec[i].conf;
    }
    r->text_off[n] = (int32_t)off;
    return r;
}

// Reader.readtext for one page on one lane
bbocr_results* readtext_page(Handle* h, Lane& lane, const bbocr_image& img, const bbocr_params& p) {
    PageWork pw;
    detect_page(h, lane, img, p, pw);
    std::vector<PageWork*> one{&pw};
    recognize_group(h, lane, one, p);
    return assemble_page(pw);
}

}  // namespace

extern "C" {

void bbocr_results_free(bbocr_results* r) {
    if (!r) return;
    delete[] r->box;
    delete[] r->is_free;
    delete[] r->text_off;
    delete[] r->text_idx;
    delete[] r->conf;
    delete r;
}

int bbocr_readtext(bbocr_handle* h, const bbocr_image* img, const bbocr_params* p, bbocr_results** out) {
    return guarded(h, [&] {
        ARG_CHECK(img && out, "null argument");
        bbocr_params dp;
        if (!p) { bbocr_default_params(&dp); p = &dp; }
        *out = readtext_page(h, h->lanes[0], *img, *p);
    });
}

int bbocr_recognize(bbocr_handle* h, const uint8_t* gray, int H, int W, int on_device, const int32_t* hlist, int nh,
                    const double* flist, int nf, const bbocr_params* p, bbocr_results** out) {
    return guarded(h, [&] {
        ARG_CHECK(gray && out && H > 0 && W > 0 && nh >= 0 && nf >= 0 && (nh == 0 || hlist) && (nf == 0 || flist), "bad arguments");
        if (!h->crnn_loaded) fail(BBOCR_E_STATE, "weights not loaded");
        bbocr_params dp;
        if (!p) { bbocr_default_params(&dp); p = &dp; }
        Lane& lane = h->lanes[0];
        DevBuf dgray;
        const uint8_t* g = gray;
        if (!on_device) { upload(lane, dgray, gray, (size_t)H * W); g = dgray.as<uint8_t>(); }
        PageWork pw;
        pw.H = H; pw.W = W;
        std::vector<int32_t> hl(hlist, hlist + (size_t)nh * 4);
        std::vector<double> fl(flist, flist + (size_t)nf * 8);
        build_crops(h, lane, g, H, W, hl, fl, *p, pw);
        CUDA_CHECK(stream_sync(lane.stream));
        lane.in_busy = false;
        std::vector<PageWork*> one{&pw};
        recognize_group(h, lane, one, *p);
        *out = assemble_page(pw);
    });
}

int bbocr_readtext_batch(bbocr_handle* h, int n, const bbocr_image* imgs, const bbocr_params* p, bbocr_results** out) {
    return guarded(h, [&] {
        ARG_CHECK(n >= 0 && (n == 0 || (imgs && out)), "null argument");
        bbocr_params dp;
        if (!p) { bbocr_default_params(&dp); p = &dp; }
        for (int i = 0; i < n; ++i) out[i] = nullptr;
        if (n == 0) return;
        // Detector lanes (stream + pinned staging + host thread each) take pages off a shared counter; recogniser lanes
        // take GROUPS of finished pages off a ready queue (detection of later pages overlaps recognition of earlier ones).
        const int nd = std::min<int>(h->n_det_lanes, cdiv(n, h->det_batch));
        const int nr = std::min<int>((int)h->lanes.size() - h->n_det_lanes, std::max(1, cdiv(n, h->rec_group)));
        std::vector<PageWork> work(n);
        // processing order: pages grouped by size, largest first (stable), so that same-size pages of a MIXED batch are
        // neighbours for the detector's pair batching and the long pages start first; results stay indexed by input position
        std::vector<int> order(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            const int64_t pa = (int64_t)imgs[a].H * imgs[a].W, pb = (int64_t)imgs[b].H * imgs[b].W;
            if (pa != pb) return pa > pb;
            return imgs[a].W > imgs[b].W;
        });
        std::atomic<int> next{0};
        std::mutex qmu;
        std::condition_variable qcv;
        std::vector<int> ready;
        int det_done = 0;              // pages that left the detector (finished or abandoned)
        bool failed = false;
        std::vector<std::string> errs(nd + nr);
        std::vector<int> codes(nd + nr, 0);
        auto record = [&](int slot, int code, const std::string& msg) {
            std::lock_guard<std::mutex> g(qmu);
            codes[slot] = code;
            errs[slot] = msg;
            failed = true;
            next.store(n);
            qcv.notify_all();
        };
        std::vector<std::thread> threads;
        for (int l = 0; l < nd; ++l)
            threads.emplace_back([&, l] {
                try {
                    CUDA_CHECK(cudaSetDevice(h->device));
                    const int db = h->det_batch;
                    for (int i0; (i0 = next.fetch_add(db)) < n;) {
                        const int kmax = std::min(db, n - i0);
                        for (int j = 0; j < kmax;) {
                            // consecutive pages of identical size go through the detector network as one batch
                            int k = 1;
                            while (j + k < kmax && imgs[order[i0 + j + k]].H == imgs[order[i0 + j]].H &&
                                   imgs[order[i0 + j + k]].W == imgs[order[i0 + j]].W) ++k;
                            const bbocr_image* ip[16];
                            PageWork* pp[16];
                            for (int q = 0; q < k; ++q) {
                                const int idx = order[i0 + j + q];
                                work[idx].index = idx;
                                ip[q] = &imgs[idx];
                                pp[q] = &work[idx];
                            }
                            detect_pages(h, h->lanes[l], ip, k, *p, pp);
                            std::lock_guard<std::mutex> g(qmu);
                            for (int q = 0; q < k; ++q) ready.push_back(order[i0 + j + q]);
                            det_done += k;
                            qcv.notify_all();
                            j += k;
                        }
                    }
                } catch (const Error& e) {
                    record(l, e.code, e.what());
                } catch (const std::exception& e) {
                    record(l, BBOCR_E_ARG, e.what());
                }
                std::lock_guard<std::mutex> g(qmu);
                qcv.notify_all();
            });
        for (int l = 0; l < nr; ++l)
            threads.emplace_back([&, l] {
                try {
                    CUDA_CHECK(cudaSetDevice(h->device));
                    Lane& lane = h->lanes[h->n_det_lanes + l];
                    for (;;) {
                        std::vector<PageWork*> group;
                        {
                            std::unique_lock<std::mutex> g(qmu);
                            qcv.wait(g, [&] { return failed || (int)ready.size() >= h->rec_group || det_done >= n; });
                            if (failed) return;
                            if (ready.empty()) return;              // det_done == n and nothing left
                            const int take = std::min<int>((int)ready.size(), h->rec_group);
                            for (int k = 0; k < take; ++k) group.push_back(&work[ready[k]]);
                            ready.erase(ready.begin(), ready.begin() + take);
                        }
#ifdef BBOCR_DIAG          // diagnostics build only (make DIAG=1): detector-only throughput; never in the shipped library
                        static const bool skip_rec = getenv("BBOCR_DIAG_SKIP_REC") != nullptr;
                        if (!skip_rec)
#endif
                        recognize_group(h, lane, group, *p);
                        for (PageWork* pw : group) {
                            out[pw->index] = assemble_page(*pw);
                            pw->dcrops.release();
                        }
                    }
                } catch (const Error& e) {
                    record(nd + l, e.code, e.what());
                } catch (const std::exception& e) {
                    record(nd + l, BBOCR_E_ARG, e.what());
                }
            });
        for (auto& t : threads) t.join();
        for (size_t l = 0; l < codes.size(); ++l)
            if (codes[l]) {
                for (int i = 0; i < n; ++i) { bbocr_results_free(out[i]); out[i] = nullptr; }
                throw Error(codes[l], errs[l]);
            }
    });
}

// ---- image decode ------------------------------------------------------------------------------------
int bbocr_jpeg_info(const uint8_t* data, size_t n, int* H, int* W, int* channels, int* orientation) {
    if (!data || !H || !W || !channels || !orientation) return BBOCR_E_ARG;
    try {
        bbocr::jpeg_info(data, n, H, W, channels, orientation);
    } catch (const bbocr::Error& e) {
        return e.code;
    } catch (...) {
        return BBOCR_E_ARG;
    }
    return BBOCR_OK;
}

int bbocr_jpeg_coefficients(const uint8_t* data, size_t n, int16_t* out, int64_t cap_blocks, int64_t* n_blocks) {
    if (!data || !n_blocks) return BBOCR_E_ARG;
    try {
        *n_blocks = bbocr::jpeg_coefficients_host(data, n, out, cap_blocks);
    } catch (const bbocr::Error& e) {
        return e.code;
    } catch (...) {
        return BBOCR_E_ARG;
    }
    return BBOCR_OK;
}

int bbocr_jpeg_decode(bbocr_handle* h, const uint8_t* data, size_t n, int ignore_orientation, uint8_t* out_bgr, uint8_t* out_gray,
                      int out_on_device, int* H, int* W) {
    return guarded(h, [&] {
        ARG_CHECK(data && n > 0 && H && W, "bad arguments");
        Lane& lane = h->lanes[0];
        if (out_on_device || (!out_bgr && !out_gray)) {
            jpeg_decode_dev(h, lane, data, n, ignore_orientation, out_bgr, out_gray, H, W);
            CUDA_CHECK(stream_sync(lane.stream));
            lane.in_busy = false;
            return;
        }
        int ch = 0, o = 0;
        jpeg_info(data, n, H, W, &ch, &o);
        if (ignore_orientation && o >= 5) std::swap(*H, *W);
        const size_t px = (size_t)*H * *W;
        DevBuf dbgr, dgray;
        if (out_bgr) dbgr.alloc(px * 3, lane.stream);
        if (out_gray) dgray.alloc(px, lane.stream);
        jpeg_decode_dev(h, lane, data, n, ignore_orientation, out_bgr ? dbgr.as<uint8_t>() : nullptr,
                        out_gray ? dgray.as<uint8_t>() : nullptr, H, W);
        if (out_bgr) download(lane, out_bgr, dbgr.p, px * 3);
        if (out_gray) download(lane, out_gray, dgray.p, px);
        CUDA_CHECK(stream_sync(lane.stream));
        lane.in_busy = false;
    });
}

int bbocr_jpeg_decode_batch(bbocr_handle* h, int n, const uint8_t* const* data, const size_t* sizes, int ignore_orientation,
                            uint8_t* const* out_bgr, uint8_t* const* out_gray) {
    return guarded(h, [&] {
        ARG_CHECK(n >= 0 && (n == 0 || (data && sizes)) && (out_bgr || out_gray), "bad arguments");
        if (n == 0) return;
        // Host threads parse / stage / launch GROUPS of files, each thread over its own two lanes (the staging of one group
        // fills while the other group's work drains); a group's entropy decoding is ONE launch over all its restart intervals.
        static const int kThreads = getenv("BBOCR_JPEG_THREADS") ? std::min(32, std::max(1, atoi(getenv("BBOCR_JPEG_THREADS")))) : 4;
        static const int kGroup = getenv("BBOCR_JPEG_GROUP") ? std::min(64, std::max(1, atoi(getenv("BBOCR_JPEG_GROUP")))) : 8;
        constexpr int kLanesPerThread = 2;
        if (h->jpeg_lanes.empty()) {
            h->jpeg_lanes.resize(kThreads * kLanesPerThread);
            for (auto& l : h->jpeg_lanes) CUDA_CHECK(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        }
        const int n_groups = cdiv(n, kGroup);
        const int nt = std::min(kThreads, n_groups);
        std::vector<int> codes(nt, 0);
        std::vector<std::string> errs(nt);
        std::atomic<int> next_group{0};
        auto work = [&](int t) {
            try {
                CUDA_CHECK(cudaSetDevice(h->device));
                std::vector<JpegJob> jobs;
                int k = 0;
                for (int gi; (gi = next_group.fetch_add(1)) < n_groups; ++k) {
                    jobs.clear();
                    for (int i = gi * kGroup; i < std::min(n, (gi + 1) * kGroup); ++i) {
                        ARG_CHECK(data[i] && sizes[i] > 0, "bad arguments");
                        jobs.push_back(JpegJob{data[i], sizes[i], out_bgr ? out_bgr[i] : nullptr, out_gray ? out_gray[i] : nullptr, 0, 0});
                    }
                    jpeg_decode_group_dev(h, h->jpeg_lanes[t * kLanesPerThread + k % kLanesPerThread], jobs.data(), (int)jobs.size(),
                                          ignore_orientation);
                }
                for (int l = 0; l < kLanesPerThread; ++l) {
                    Lane& lane = h->jpeg_lanes[t * kLanesPerThread + l];
                    CUDA_CHECK(stream_sync(lane.stream));
                    lane.in_busy = false;
                }
            } catch (const Error& e) {
                codes[t] = e.code;
                errs[t] = e.what();
            } catch (const std::exception& e) {
                codes[t] = BBOCR_E_ARG;
                errs[t] = e.what();
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& t : th) t.join();
        for (int t = 0; t < nt; ++t)
            if (codes[t]) {
                for (auto& l : h->jpeg_lanes) { cudaStreamSynchronize(l.stream); l.in_busy = false; }
                throw Error(codes[t], errs[t]);
            }
    });
}

// single-stage recogniser entry points (parity-test surface)
int bbocr_crop_horizontal(bbocr_handle* h, const uint8_t* gray, int H, int W, const int32_t box[4], uint8_t* out, int cap,
                          int* outH, int* outW, int* model_w) {
    return guarded(h, [&] {
        ARG_CHECK(gray && box && out && outH && outW && model_w && H > 0 && W > 0, "bad arguments");
        Lane& lane = h->lanes[0];
        CropJob j;
        horizontal_job(box, H, W, j);
        if (!finish_geometry(j)) { *outW = *outH = 0; *model_w = 0; return; }
        ARG_CHECK(j.d.ow * j.d.oh <= cap, "crop buffer too small");
        j.d.off = 0;
        DevBuf dg, dd, dc((size_t)j.d.ow * j.d.oh + 16, lane.stream);
        upload(lane, dg, gray, (size_t)H * W);
        CUDA_CHECK(stream_sync(lane.stream));
        upload(lane, dd, &j.d, sizeof(CropDesc));
        crops_dev(h, lane.stream, dg.as<uint8_t>(), H, W, dd.as<CropDesc>(), 1, &j.d, nullptr, nullptr, dc.as<uint8_t>());
        download(lane, out, dc.p, (size_t)j.d.ow * j.d.oh);
        *outW = j.d.ow; *outH = j.d.oh; *model_w = j.d.model_w;
    });
}

int bbocr_crop_free(bbocr_handle* h, const uint8_t* gray, int H, int W, const double quad[8], uint8_t* out, int cap,
                    int* outH, int* outW, int* model_w) {
    return guarded(h, [&] {
        ARG_CHECK(gray && quad && out && outH && outW && model_w && H > 0 && W > 0, "bad arguments");
        Lane& lane = h->lanes[0];
        CropJob j;
        int mwid, mhei;
        double M[9];
        free_box_transform(quad, &mwid, &mhei, M);
        ARG_CHECK(mwid > 0 && mhei > 0, "degenerate free-form box");
        j.d.free_idx = 0; j.d.x0 = 0; j.d.y0 = 0; j.d.w = mwid; j.d.h = mhei;
        if (!finish_geometry(j)) { *outW = *outH = 0; *model_w = 0; return; }
        ARG_CHECK(j.d.ow * j.d.oh <= cap, "crop buffer too small");
        j.d.off = 0;
        DevBuf dg, dd, dm, ds((size_t)mwid * mhei + 16, lane.stream), dc((size_t)j.d.ow * j.d.oh + 16, lane.stream);
        upload(lane, dg, gray, (size_t)H * W);
        CUDA_CHECK(stream_sync(lane.stream));
        upload(lane, dd, &j.d, sizeof(CropDesc));
        CUDA_CHECK(stream_sync(lane.stream));
        upload(lane, dm, M, 72);
        crops_dev(h, lane.stream, dg.as<uint8_t>(), H, W, dd.as<CropDesc>(), 1, &j.d, dm.as<double>(), ds.as<uint8_t>(),
                  dc.as<uint8_t>());
        download(lane, out, dc.p, (size_t)j.d.ow * j.d.oh);
        *outW = j.d.ow; *outH = j.d.oh; *model_w = j.d.model_w;
    });
}

int bbocr_crnn_forward(bbocr_handle* h, const float* x, int N, int Wm, float* logits) {
    return guarded(h, [&] {
        ARG_CHECK(x && logits && N > 0 && Wm >= 64 && Wm % 64 == 0, "bad arguments");
        Lane& lane = h->lanes[0];
        const int T = Wm / 4 - 1, C = h->crnn.num_class;
        DevBuf dx, dl((size_t)N * T * C * 4, lane.stream);
        upload(lane, dx, x, (size_t)N * 64 * Wm * 4);
        crnn_forward_dev(h, lane, dx.as<float>(), N, Wm, dl.as<float>());
        download(lane, logits, dl.p, (size_t)N * T * C * 4);
    });
}

int bbocr_ctc_decode(bbocr_handle* h, const float* logits, int N, int T, int C, const uint8_t* ignore, int32_t* text_idx,
                     int32_t* text_len, double* conf) {
    return guarded(h, [&] {
        ARG_CHECK(logits && text_idx && text_len && conf && N > 0 && T > 0 && C > 1, "bad arguments");
        Lane& lane = h->lanes[0];
        cudaStream_t st = lane.stream;
        DevBuf dlg, dig;
        upload(lane, dlg, logits, (size_t)N * T * C * 4);
        if (ignore) {
            CUDA_CHECK(stream_sync(st));
            upload(lane, dig, ignore, C);
        }
        size_t se = (size_t)N * T;
        DevBuf dec((se * 3 + N) * 4, st);
        int32_t* d_text = dec.as<int32_t>();
        int32_t* d_sidx = d_text + se;
        float* d_prob = reinterpret_cast<float*>(d_sidx + se);
        int32_t* d_len = reinterpret_cast<int32_t*>(d_prob + se);
        std::vector<SeqDesc> seqs(N);
        for (int i = 0; i < N; ++i) { seqs[i].row0 = i * T; seqs[i].T = T; }
        DevBuf dseq;
        CUDA_CHECK(stream_sync(st));
        upload(lane, dseq, seqs.data(), seqs.size() * sizeof(SeqDesc));
        ctc_decode_dev(h, st, dlg.as<float>(), N * T, C, ignore ? dig.as<uint8_t>() : nullptr, dseq.as<SeqDesc>(), N, d_text,
                       d_len, d_prob, d_sidx);
        std::vector<int32_t> hd(se * 3 + N);
        download(lane, hd.data(), dec.p, hd.size() * 4);
        memcpy(text_idx, hd.data(), se * 4);
        memcpy(text_len, hd.data() + se * 3, (size_t)N * 4);
        const float* prob = reinterpret_cast<const float*>(hd.data() + se * 2);
        for (int i = 0; i < N; ++i) conf[i] = confidence_of(prob + (size_t)i * T, hd.data() + se + (size_t)i * T, T);
    });
}

// test/diagnostic hook: cumulative wall-clock ms per readtext stage [craft enqueue, det (waits for CRAFT), boxes+crops,
// recognise pass 1, pass 2]; resets on read
int bbocr_dbg_stage_ms(bbocr_handle* h, double* out5) {
    if (!h || !out5) return BBOCR_E_ARG;
    std::lock_guard<std::mutex> g(h->stat_mu);
    for (int i = 0; i < 5; ++i) { out5[i] = h->stage_ms[i]; h->stage_ms[i] = 0; }
    return 0;
}

int64_t bbocr_launch_count(const bbocr_handle* h) { return h ? h->launches.load() : 0; }
void bbocr_reset_launch_count(bbocr_handle* h) {
    if (h) h->launches.store(0);
}
int bbocr_enable_conv_timing(bbocr_handle* h, int on) {
    return guarded(h, [&] { h->conv_timing = on != 0; });
}
int bbocr_conv_stats(bbocr_handle* h, double* ms, int64_t* launches, double* flops) {
    return guarded(h, [&] {
        CUDA_CHECK(cudaDeviceSynchronize());
        std::lock_guard<std::mutex> g(h->stat_mu);
        double total = 0;
        for (auto& ev : h->conv_events) {
            float t = 0;
            CUDA_CHECK(cudaEventElapsedTime(&t, ev.first, ev.second));
            total += t;
            cudaEventDestroy(ev.first);
            cudaEventDestroy(ev.second);
        }
        h->conv_events.clear();
        if (ms) *ms = total;
        if (launches) *launches = h->conv_launches;
        if (flops) *flops = h->conv_flops;
        h->conv_launches = 0;
        h->conv_flops = 0;
    });
}

}  // extern "C"
