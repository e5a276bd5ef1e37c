// autocrop.cu -- the text-region crop heuristic of BB-OCR's extractor on the device.
//
// Reference: pipeline_demo/extractor/enhanced_extractor.py:239-372 (_auto_crop_text_region): gray -> GaussianBlur 3x3 ->
// CLAHE(2.0) -> four text cues (adaptive mean 35/10 INV, adaptive Gaussian 31/5 INV, Otsu INV, Otsu of the Sobel
// magnitude) OR-ed into a mask -> two rectangle-morphology variants OR-ed -> boundingRect of every external contour ->
// area filter / union / margin on the host.  Everything per pixel runs here, bit-exact against cv2 (oracle/autocrop_np.py
// lists the cv2 facts this relies on, each pinned by a CPU test):
//   * the mask is produced as a BIT image (one warp ballot = one 32-pixel word), so the morphology works on 32 pixels
//     per register: a w x h rectangle dilation is an OR of funnel-shifted words; erosion is the dilation of the complement;
//     iterations and consecutive passes are folded:  merged = dil13x5( ero19x7(dil17x5 m) | ero31x11(dil29x9 m) )
//   * RETR_EXTERNAL == 8-connected foreground components that touch the background region connected to the image frame
//     (background 4-connected): ONE union-find over all pixels labels foreground (8-conn.) and background (4-conn.)
//     together; a component is external iff one of its pixels is on the image border or 4-adjacent to a frame-connected
//     background pixel.
#include <climits>

#include "engine.h"

namespace bbocr {

void gaussian3_kernel_q8(float sigma, int* k0, int* k1);

namespace {

__device__ __forceinline__ int uf_find(const int* __restrict__ L, int i) {
    int p = L[i];
    while (p != i) { i = p; p = L[i]; }
    return i;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) { int old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { int old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

// ---- Sobel magnitude + both histograms ------------------------------------------------------------------------------
// grad = min(255, min(255,|gx|) + min(255,|gy|)), 3x3 Sobel, BORDER_REFLECT_101.  hist[0..255] = eq, hist[256..511] = grad.
constexpr int GT_W = 64, GT_H = 16;
__global__ void __launch_bounds__(256) k_ac_grad_hist(const uint8_t* __restrict__ eq, int H, int W, uint8_t* __restrict__ grad,
                                                      unsigned int* __restrict__ hist) {
    __shared__ uint8_t tile[GT_H + 2][GT_W + 2];
    __shared__ unsigned int sh[512];
    int tid = threadIdx.x;
    for (int i = tid; i < 512; i += 256) sh[i] = 0;
    int x0 = blockIdx.x * GT_W, y0 = blockIdx.y * GT_H;
    for (int i = tid; i < (GT_H + 2) * (GT_W + 2); i += 256) {
        int ty = i / (GT_W + 2), tx = i % (GT_W + 2);
        int y = y0 + ty - 1, x = x0 + tx - 1;
        // BORDER_REFLECT_101 (a 1-pixel halo needs one reflection); single-row / single-column images fold onto 0
        y = y < 0 ? -y : (y >= H ? 2 * H - 2 - y : y);
        x = x < 0 ? -x : (x >= W ? 2 * W - 2 - x : x);
        y = min(max(y, 0), H - 1);
        x = min(max(x, 0), W - 1);
        tile[ty][tx] = eq[(int64_t)y * W + x];
    }
    __syncthreads();
    for (int i = tid; i < GT_H * GT_W; i += 256) {
        int ty = i / GT_W, tx = i % GT_W;
        int y = y0 + ty, x = x0 + tx;
        if (y >= H || x >= W) continue;
        int a = tile[ty][tx], b = tile[ty][tx + 1], c = tile[ty][tx + 2];
        int d = tile[ty + 1][tx], e = tile[ty + 1][tx + 1], f = tile[ty + 1][tx + 2];
        int g = tile[ty + 2][tx], hh = tile[ty + 2][tx + 1], k = tile[ty + 2][tx + 2];
        int gx = (c + 2 * f + k) - (a + 2 * d + g);
        int gy = (g + 2 * hh + k) - (a + 2 * b + c);
        int m = min(255, min(255, abs(gx)) + min(255, abs(gy)));
        grad[(int64_t)y * W + x] = (uint8_t)m;
        atomicAdd(&sh[e], 1u);
        atomicAdd(&sh[256 + m], 1u);
    }
    __syncthreads();
    for (int i = tid; i < 512; i += 256)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// cv2 getThreshVal_Otsu_8u in plain (uncontracted) double arithmetic; thread t handles histogram t.
__global__ void k_ac_otsu(const unsigned int* __restrict__ hist, double npix, int* __restrict__ thr) {
    if (threadIdx.x >= 2) return;
    const unsigned int* h = hist + 256 * threadIdx.x;
    const double eps = (double)1.1920928955078125e-07f;
    double scale = __ddiv_rn(1.0, npix);
    double mu = 0.0;
    for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)h[i]));
    mu = __dmul_rn(mu, scale);
    double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
    int max_val = 0;
    for (int i = 0; i < 256; ++i) {
        double p_i = __dmul_rn((double)h[i], scale);
        mu1 = __dmul_rn(mu1, q1);
        q1 = __dadd_rn(q1, p_i);
        double q2 = __dsub_rn(1.0, q1);
        if (fmin(q1, q2) < eps || fmax(q1, q2) > __dsub_rn(1.0, eps)) continue;
        mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
        double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
        double diff = __dsub_rn(mu1, mu2);
        double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), diff), diff);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    thr[threadIdx.x] = max_val;
}

// mask bit = thr_mean | thr_gaus | (eq <= t_eq) | (grad > t_grad); one ballot per 32 pixels.  Row pitch Ww words,
// bit b of word j = pixel 32 j + b; bits beyond W stay 0.
__global__ void k_ac_mask(const uint8_t* __restrict__ thr_mean, const uint8_t* __restrict__ thr_gaus,
                          const uint8_t* __restrict__ eq, const uint8_t* __restrict__ grad, const int* __restrict__ thr, int H,
                          int W, int Ww, uint32_t* __restrict__ bits) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    bool on = false;
    if (x < W) {
        int64_t i = (int64_t)y * W + x;
        on = thr_mean[i] || thr_gaus[i] || (int)eq[i] <= thr[0] || (int)grad[i] > thr[1];
    }
    unsigned m = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < Ww) bits[(int64_t)y * Ww + (x >> 5)] = m;
}

// four pixels per thread (32-bit loads of the four planes), nibbles OR-reduced over 8 lanes into one word; needs W % 4 == 0
__global__ void __launch_bounds__(256) k_ac_mask4(const uint32_t* __restrict__ thr_mean, const uint32_t* __restrict__ thr_gaus,
                                                  const uint32_t* __restrict__ eq, const uint32_t* __restrict__ grad,
                                                  const int* __restrict__ thr, int H, int W, int Ww, uint32_t* __restrict__ bits) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, lane = threadIdx.x & 31;
    const int W4 = W >> 2;
    uint32_t nib = 0;
    if (g < W4) {
        const int64_t i = (int64_t)y * W4 + g;
        const uint32_t m = __ldg(thr_mean + i) | __ldg(thr_gaus + i), e = __ldg(eq + i), gr = __ldg(grad + i);
        const int t0 = thr[0], t1 = thr[1];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const bool on = ((m >> (8 * c)) & 255u) || (int)((e >> (8 * c)) & 255u) <= t0 || (int)((gr >> (8 * c)) & 255u) > t1;
            nib |= (on ? 1u : 0u) << c;
        }
    }
    uint32_t v = nib << (4 * (lane & 7));
    v |= __shfl_xor_sync(0xffffffffu, v, 1);
    v |= __shfl_xor_sync(0xffffffffu, v, 2);
    v |= __shfl_xor_sync(0xffffffffu, v, 4);
    if ((lane & 7) == 0 && (g >> 3) < Ww) bits[(int64_t)y * Ww + (g >> 3)] = v;
}

__global__ void k_ac_pack(const uint8_t* __restrict__ src, int H, int W, int Ww, uint32_t* __restrict__ bits) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    bool on = x < W && src[(int64_t)y * W + x] != 0;
    unsigned m = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < Ww) bits[(int64_t)y * Ww + (x >> 5)] = m;
}

__global__ void k_ac_unpack(const uint32_t* __restrict__ bits, int H, int W, int Ww, uint8_t* __restrict__ dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < W) dst[(int64_t)y * W + x] = ((bits[(int64_t)y * Ww + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
}

// (2rx+1) x (2ry+1) rectangle dilation of a bit image (pixels outside the image do not contribute), rx <= 31.
// inv_in: dilate the complement (of `a | b`); inv_out: complement the result  => inv_in && inv_out = erosion.
__global__ void k_bit_dilate(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int H, int W, int Ww, int rx, int ry,
                             int inv_in, int inv_out, uint32_t* __restrict__ dst) {
    int j = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (j >= Ww) return;
    const uint32_t tail = (W & 31) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;
    auto load = [&](int yy, int jj) -> uint32_t {
        if (jj < 0 || jj >= Ww) return 0u;
        uint32_t v = a[(int64_t)yy * Ww + jj];
        if (b) v |= b[(int64_t)yy * Ww + jj];
        if (inv_in) v = ~v;
        if (jj == Ww - 1) v &= tail;
        return v;
    };
    // rows first (plain OR of words), then ONE horizontal spread of the 96-bit window by shift doubling: a value that already
    // holds the OR of shifts 0..k extends to 0..k+s with one more shift by s <= k+1, so rx = 15 costs 4 steps per side
    uint32_t p = 0, c = 0, n = 0;
    int ya = max(y - ry, 0), yb = min(y + ry, H - 1);
    for (int yy = ya; yy <= yb; ++yy) {
        p |= load(yy, j - 1);
        c |= load(yy, j);
        n |= load(yy, j + 1);
    }
    uint64_t L = ((uint64_t)c << 32) | p, R = ((uint64_t)n << 32) | c;
    for (int cover = 0; cover < rx;) {
        const int step = min(cover + 1, rx - cover);
        L |= L << step;
        R |= R >> step;
        cover += step;
    }
    uint32_t acc = (uint32_t)(L >> 32) | (uint32_t)R;
    if (inv_out) acc = ~acc;
    if (j == Ww - 1) acc &= tail;
    dst[(int64_t)y * Ww + j] = acc;
}

// ---- components: foreground 8-connected, background 4-connected, one label array --------------------------------------
__global__ void k_cc_init(const uint32_t* __restrict__ bits, int H, int W, int Ww, int* __restrict__ L) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int lane = threadIdx.x & 31;
    uint32_t word = bits[(int64_t)y * Ww + (x >> 5)];
    uint32_t same = ((word >> lane) & 1u) ? word : ~word;
    uint32_t z = ~same & ((1u << lane) - 1u);
    int start = z ? 32 - __clz(z) : 0;
    L[y * W + x] = y * W + x - lane + start;
}

__device__ __forceinline__ unsigned bit_at(const uint32_t* __restrict__ bits, int Ww, int x, int y) {
    return (bits[(int64_t)y * Ww + (x >> 5)] >> (x & 31)) & 1u;
}

__global__ void k_cc_merge(const uint32_t* __restrict__ bits, int H, int W, int Ww, int* __restrict__ L) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    bool inb = x < W;
    int lane = threadIdx.x & 31;
    int i = y * W + x;
    unsigned v = 0, u = 0;
    uint32_t word = 0;
    bool up_same = false;
    if (inb) {
        word = bits[(int64_t)y * Ww + (x >> 5)];
        v = (word >> lane) & 1u;
        if (y > 0) {
            u = bit_at(bits, Ww, x, y - 1);
            up_same = (u == v);
        }
    }
    unsigned um = __ballot_sync(0xffffffffu, up_same);
    if (!inb) return;
    if (up_same) {
        bool dup = lane > 0 && ((um >> (lane - 1)) & 1u) && (((word >> (lane - 1)) & 1u) == v);
        if (!dup) uf_union(L, i, i - W);
    }
    if (v && y > 0 && !u) {                                  // diagonals only matter when the pixel above is background
        if (x > 0 && bit_at(bits, Ww, x - 1, y - 1)) uf_union(L, i, i - W - 1);
        if (x < W - 1 && bit_at(bits, Ww, x + 1, y - 1)) uf_union(L, i, i - W + 1);
    }
    if (lane == 0 && x > 0 && bit_at(bits, Ww, x - 1, y) == v) uf_union(L, i, i - 1);
}

__global__ void k_cc_compress(int n, int* __restrict__ L) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) L[i] = uf_find(L, i);
}

// background pixels on the image border belong to the frame-connected background
__global__ void k_cc_border(const uint32_t* __restrict__ bits, int H, int W, int Ww, const int* __restrict__ L,
                            uint8_t* __restrict__ outer) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int x, y;
    if (t < W) { x = t; y = 0; }
    else if (t < 2 * W) { x = t - W; y = H - 1; }
    else if (t < 2 * W + H) { x = 0; y = t - 2 * W; }
    else if (t < 2 * W + 2 * H) { x = W - 1; y = t - 2 * W - H; }
    else return;
    if (!bit_at(bits, Ww, x, y)) outer[L[y * W + x]] = 1;
}

__global__ void k_cc_assign(const uint32_t* __restrict__ bits, int H, int W, int Ww, const int* __restrict__ L, int* __restrict__ cid,
                            int* __restrict__ count) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int i = y * W + x;
    if (L[i] == i && bit_at(bits, Ww, x, y)) cid[i] = atomicAdd(count, 1);
}

__global__ void k_cc_box_init(int n, int* __restrict__ box) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    box[5 * i + 0] = INT_MAX;
    box[5 * i + 1] = INT_MAX;
    box[5 * i + 2] = -1;
    box[5 * i + 3] = -1;
    box[5 * i + 4] = 0;
}

// per component: min x, min y, max x, max y, external flag
__global__ void k_cc_stats(const uint32_t* __restrict__ bits, int H, int W, int Ww, const int* __restrict__ L,
                           const int* __restrict__ cid, const uint8_t* __restrict__ outer, int* __restrict__ box) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    if (!bit_at(bits, Ww, x, y)) return;
    int i = y * W + x;
    bool l = x > 0 && bit_at(bits, Ww, x - 1, y), r = x < W - 1 && bit_at(bits, Ww, x + 1, y);
    bool u = y > 0 && bit_at(bits, Ww, x, y - 1), d = y < H - 1 && bit_at(bits, Ww, x, y + 1);
    if (l && r && u && d) return;                            // interior pixel: contributes nothing
    int* b = box + 5 * cid[L[i]];
    if (!l) { atomicMin(&b[0], x); atomicMin(&b[1], y); atomicMax(&b[3], y); }
    if (!r) atomicMax(&b[2], x);
    bool ext = x == 0 || y == 0 || x == W - 1 || y == H - 1;
    if (!ext && !l) ext = outer[L[i - 1]];
    if (!ext && !r) ext = outer[L[i + 1]];
    if (!ext && !u) ext = outer[L[i - W]];
    if (!ext && !d) ext = outer[L[i + W]];
    if (ext) b[4] = 1;
}

}  // namespace

// boundingRect of every RETR_EXTERNAL contour of a bit image -> (x, y, w, h) rows sorted by (y, x, w, h)
static void external_boxes_bits(Handle* h, cudaStream_t st, const uint32_t* bits, int H, int W, int Ww, std::vector<int32_t>& out) {
    out.clear();
    const int n = H * W;
    DevBuf L((size_t)n * 4, st), cid((size_t)n * 4, st), outer((size_t)n, st), cnt(4, st);
    dim3 grd(cdiv(W, 256), H);
    CUDA_CHECK(cudaMemsetAsync(outer.p, 0, (size_t)n, st));
    CUDA_CHECK(cudaMemsetAsync(cnt.p, 0, 4, st));
    k_cc_init<<<grd, 256, 0, st>>>(bits, H, W, Ww, L.as<int>());
    k_cc_merge<<<grd, 256, 0, st>>>(bits, H, W, Ww, L.as<int>());
    k_cc_compress<<<cdiv(n, 256), 256, 0, st>>>(n, L.as<int>());
    k_cc_border<<<cdiv(2 * W + 2 * H, 256), 256, 0, st>>>(bits, H, W, Ww, L.as<int>(), outer.as<uint8_t>());
    k_cc_assign<<<grd, 256, 0, st>>>(bits, H, W, Ww, L.as<int>(), cid.as<int>(), cnt.as<int>());
    count_launch(h, 5);
    CUDA_CHECK(cudaGetLastError());
    int ncomp = 0;
    CUDA_CHECK(cudaMemcpyAsync(&ncomp, cnt.p, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_sync(st));
    if (ncomp == 0) return;
    DevBuf box((size_t)ncomp * 20, st);
    k_cc_box_init<<<cdiv(ncomp, 256), 256, 0, st>>>(ncomp, box.as<int>());
    k_cc_stats<<<grd, 256, 0, st>>>(bits, H, W, Ww, L.as<int>(), cid.as<int>(), outer.as<uint8_t>(), box.as<int>());
    count_launch(h, 2);
    CUDA_CHECK(cudaGetLastError());
    std::vector<int> hb((size_t)ncomp * 5);
    CUDA_CHECK(cudaMemcpyAsync(hb.data(), box.p, (size_t)ncomp * 20, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_sync(st));
    std::vector<std::array<int32_t, 4>> rows;
    for (int k = 0; k < ncomp; ++k)
        if (hb[5 * k + 4]) rows.push_back({hb[5 * k + 1], hb[5 * k + 0], hb[5 * k + 2] - hb[5 * k + 0] + 1, hb[5 * k + 3] - hb[5 * k + 1] + 1});
    std::sort(rows.begin(), rows.end());                     // (y, x, w, h)
    for (auto& r : rows) { out.push_back(r[1]); out.push_back(r[0]); out.push_back(r[2]); out.push_back(r[3]); }
}

static void bit_dilate(Handle* h, cudaStream_t st, const uint32_t* a, const uint32_t* b, int H, int W, int Ww, int kw, int kh,
                       bool erode, uint32_t* dst) {
    k_bit_dilate<<<dim3(cdiv(Ww, 64), H), 64, 0, st>>>(a, b, H, W, Ww, kw / 2, kh / 2, erode, erode, dst);
    count_launch(h);
}

// enhanced_extractor.py:288-333 -- area filter, union, inflate-if-small, margin
static bool crop_rect_from_boxes(const std::vector<int32_t>& boxes, int H, int W, int margin, int32_t rect[4]) {
    const double img_area = (double)((int64_t)H * W);
    bool any = false;
    int64_t x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    for (size_t k = 0; k + 3 < boxes.size(); k += 4) {
        int64_t x = boxes[k], y = boxes[k + 1], w = boxes[k + 2], hh = boxes[k + 3];
        double area = (double)(w * hh);
        if (area < 0.0001 * img_area || area > 0.10 * img_area) continue;
        if (!any) { x0 = x; y0 = y; x1 = x + w; y1 = y + hh; any = true; }
        else { x0 = std::min(x0, x); y0 = std::min(y0, y); x1 = std::max(x1, x + w); y1 = std::max(y1, y + hh); }
    }
    if (!any) return false;
    if ((double)((x1 - x0) * (y1 - y0)) < 0.12 * img_area) {
        int64_t pad = (int64_t)(0.03 * (double)std::max(W, H));
        x0 = std::max<int64_t>(0, x0 - pad); y0 = std::max<int64_t>(0, y0 - pad);
        x1 = std::min<int64_t>(W, x1 + pad); y1 = std::min<int64_t>(H, y1 + pad);
    }
    x0 = std::max<int64_t>(0, x0 - margin); y0 = std::max<int64_t>(0, y0 - margin);
    x1 = std::min<int64_t>(W, x1 + margin); y1 = std::min<int64_t>(H, y1 + margin);
    if (x1 <= x0 || y1 <= y0) return false;
    rect[0] = (int32_t)x0; rect[1] = (int32_t)y0; rect[2] = (int32_t)x1; rect[3] = (int32_t)y1;
    return true;
}

bool autocrop_dev(Handle* h, cudaStream_t st, const uint8_t* bgr, int H, int W, int channels, int stride, int margin, int32_t rect[4],
                  AutoCropDebug* dbg) {
    ARG_CHECK(H > 0 && W > 0 && (int64_t)H * W < (int64_t)INT_MAX, "autocrop: bad geometry");
    const size_t n = (size_t)H * W;
    const int Ww = cdiv(W, 32);
    const size_t nb = (size_t)H * Ww * 4;
    DevBuf gray(channels == 3 ? n : 16, st), blur(n, st), eq(n, st), tmean(n, st), tgaus(n, st), grad(n, st);
    DevBuf small(64 * 256 * 4 + 64 * 256 + 512 * 4 + 16, st);
    unsigned int* chist = small.as<unsigned int>();
    uint8_t* luts = small.as<uint8_t>() + 64 * 256 * 4;
    unsigned int* hist = reinterpret_cast<unsigned int*>(small.as<uint8_t>() + 64 * 256 * 4 + 64 * 256);
    int* thr = reinterpret_cast<int*>(hist + 512);
    if (channels == 3) pp_gray(h, st, bgr, H, W, stride, gray.as<uint8_t>());
    pp_gaussian3(h, st, channels == 3 ? gray.as<uint8_t>() : bgr, blur.as<uint8_t>(), H, W, 0.f, nullptr);
    pp_clahe_luts(h, st, blur.as<uint8_t>(), H, W, 2.0f, nullptr, chist, luts);
    pp_clahe_apply(h, st, blur.as<uint8_t>(), eq.as<uint8_t>(), H, W, nullptr, luts);
    pp_adaptive_threshold(h, st, eq.as<uint8_t>(), tmean.as<uint8_t>(), H, W, 0, 1, 35, 10.f);
    pp_adaptive_threshold(h, st, eq.as<uint8_t>(), tgaus.as<uint8_t>(), H, W, 1, 1, 31, 5.f);
    CUDA_CHECK(cudaMemsetAsync(hist, 0, 512 * 4, st));
    k_ac_grad_hist<<<dim3(cdiv(W, GT_W), cdiv(H, GT_H)), 256, 0, st>>>(eq.as<uint8_t>(), H, W, grad.as<uint8_t>(), hist);
    k_ac_otsu<<<1, 32, 0, st>>>(hist, (double)n, thr);
    DevBuf mask(nb, st), a1(nb, st), a2(nb, st), b1(nb, st), b2(nb, st), merged(nb, st);
    dim3 grd(cdiv(W, 256), H);
    if ((W & 3) == 0)
        k_ac_mask4<<<dim3(cdiv(Ww * 8, 256), H), 256, 0, st>>>(tmean.as<uint32_t>(), tgaus.as<uint32_t>(), eq.as<uint32_t>(),
                                                               grad.as<uint32_t>(), thr, H, W, Ww, mask.as<uint32_t>());
    else
        k_ac_mask<<<grd, 256, 0, st>>>(tmean.as<uint8_t>(), tgaus.as<uint8_t>(), eq.as<uint8_t>(), grad.as<uint8_t>(), thr, H, W, Ww,
                                       mask.as<uint32_t>());
    count_launch(h, 3);
    bit_dilate(h, st, mask.as<uint32_t>(), nullptr, H, W, Ww, 17, 5, false, a1.as<uint32_t>());
    bit_dilate(h, st, a1.as<uint32_t>(), nullptr, H, W, Ww, 19, 7, true, a2.as<uint32_t>());
    bit_dilate(h, st, mask.as<uint32_t>(), nullptr, H, W, Ww, 29, 9, false, b1.as<uint32_t>());
    bit_dilate(h, st, b1.as<uint32_t>(), nullptr, H, W, Ww, 31, 11, true, b2.as<uint32_t>());
    bit_dilate(h, st, a2.as<uint32_t>(), b2.as<uint32_t>(), H, W, Ww, 13, 5, false, merged.as<uint32_t>());
    CUDA_CHECK(cudaGetLastError());
    std::vector<int32_t> boxes;
    external_boxes_bits(h, st, merged.as<uint32_t>(), H, W, Ww, boxes);
    if (dbg) {
        dbg->boxes = boxes;
        CUDA_CHECK(cudaMemcpyAsync(dbg->otsu, thr, 8, cudaMemcpyDeviceToHost, st));
        DevBuf u8(n, st);
        for (int which = 0; which < 2; ++which) {
            uint8_t* dst = which ? dbg->merged : dbg->mask;
            if (!dst) continue;
            k_ac_unpack<<<grd, 256, 0, st>>>((which ? merged : mask).as<uint32_t>(), H, W, Ww, u8.as<uint8_t>());
            CUDA_CHECK(cudaMemcpyAsync(dst, u8.p, n, cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(stream_sync(st));
        }
        CUDA_CHECK(stream_sync(st));
    }
    return crop_rect_from_boxes(boxes, H, W, margin, rect);
}

void external_boxes_dev(Handle* h, cudaStream_t st, const uint8_t* binary_dev, int H, int W, std::vector<int32_t>& boxes) {
    ARG_CHECK(H > 0 && W > 0 && (int64_t)H * W < (int64_t)INT_MAX, "external_boxes: bad geometry");
    const int Ww = cdiv(W, 32);
    DevBuf bits((size_t)H * Ww * 4, st);
    k_ac_pack<<<dim3(cdiv(W, 256), H), 256, 0, st>>>(binary_dev, H, W, Ww, bits.as<uint32_t>());
    count_launch(h);
    external_boxes_bits(h, st, bits.as<uint32_t>(), H, W, Ww, boxes);
}

void rect_morph_dev(Handle* h, cudaStream_t st, const uint8_t* binary_dev, int H, int W, int kw, int kh, bool erode, uint8_t* out_dev) {
    ARG_CHECK(H > 0 && W > 0 && kw >= 1 && kh >= 1 && (kw & 1) && (kh & 1) && kw <= 63, "rect_morph: odd kernel sizes, width <= 63");
    const int Ww = cdiv(W, 32);
    DevBuf bits((size_t)H * Ww * 4, st), res((size_t)H * Ww * 4, st);
    dim3 grd(cdiv(W, 256), H);
    k_ac_pack<<<grd, 256, 0, st>>>(binary_dev, H, W, Ww, bits.as<uint32_t>());
    bit_dilate(h, st, bits.as<uint32_t>(), nullptr, H, W, Ww, kw, kh, erode, res.as<uint32_t>());
    k_ac_unpack<<<grd, 256, 0, st>>>(res.as<uint32_t>(), H, W, Ww, out_dev);
    count_launch(h, 2);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace bbocr
