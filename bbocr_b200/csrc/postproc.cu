// postproc.cu -- craft_utils.getDetBoxes_core on the device:
//   text/link thresholding (cv2.threshold, strict >), 4-connected component labelling by union-find whose roots are
//   the minimum linear index (=> ascending roots == OpenCV's raster label order), per-label area / bbox / max text
//   score, the area>=10 and max>=text_threshold filters, and per-row x-extents of (label minus link-only pixels) --
//   all the per-pixel work -- and, per kept label, the square dilation of those extents, the convex hull and the min-area
//   rectangle (cv2.minAreaRect + boxPoints restated in geom.cuh, the same code boxes.cpp runs on the host), the 0.1 "diamond"
//   rule and the clockwise start: k_det_boxes.  One D2H copy of the finished boxes per page; nothing per label on the host.
#include "engine.h"
#include "geom.cuh"

namespace bbocr {

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

// One warp covers 32 consecutive pixels of a row.  Horizontal runs inside the warp are labelled with the run start
// straight from the ballot (no atomics); flags: bit0 = foreground, bit1 = link-only pixel (link && !text).
__global__ void k_ccl_init(const float* __restrict__ text, const float* __restrict__ link, int h, int w, float low_text,
                           float link_thr, int* __restrict__ L, uint8_t* __restrict__ flags) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    bool inb = x < w;
    int i = y * w + x;
    bool t = false, l = false;
    if (inb) { t = text[i] > low_text; l = link[i] > link_thr; }
    bool fg = t || l;
    unsigned mask = __ballot_sync(0xffffffffu, fg);
    if (!inb) return;
    int lane = threadIdx.x & 31;
    int label = -1;
    if (fg) {
        unsigned z = ~mask & ((1u << lane) - 1u);
        int start = z ? 32 - __clz(z) : 0;
        label = i - lane + start;
    }
    L[i] = label;
    flags[i] = (uint8_t)((fg ? 1 : 0) | ((l && !t) ? 2 : 0));
}

// Vertical unions (once per maximal run of vertically-adjacent foreground pairs inside the warp) and the horizontal
// union across the warp boundary.
__global__ void k_ccl_merge(int h, int w, int* __restrict__ L, const uint8_t* __restrict__ flags) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    bool inb = x < w;
    int i = y * w + x;
    bool fg = inb && (flags[i] & 1);
    bool up = fg && y > 0 && (flags[i - w] & 1);
    unsigned um = __ballot_sync(0xffffffffu, up);
    int lane = threadIdx.x & 31;
    if (up && !(lane > 0 && ((um >> (lane - 1)) & 1u))) uf_union(L, i, i - w);
    if (fg && lane == 0 && x > 0 && (flags[i - 1] & 1)) uf_union(L, i, i - 1);
}

__global__ void k_ccl_compress(int n, int* __restrict__ L) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && L[i] >= 0) L[i] = uf_find(L, i);
}

// ---- raster-order ranking of roots (OpenCV label k = rank + 1) -------------------------------------------------------
__global__ void k_count_roots(int n, const int* __restrict__ L, int* __restrict__ block_counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int c = __syncthreads_count(i < n && L[i] == i);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

__global__ void k_scan_blocks(int nb, int* __restrict__ block_counts, int* __restrict__ header) {
    __shared__ int carry;
    __shared__ int s[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nb ? block_counts[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int add = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < nb) block_counts[i] = carry + s[threadIdx.x] - v;      // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) header[0] = carry;                             // number of components
}

__global__ void k_assign_ranks(int n, const int* __restrict__ L, const int* __restrict__ block_offsets,
                               int* __restrict__ rank_of_root) {
    __shared__ int warp_counts[32];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool root = i < n && L[i] == i;
    unsigned m = __ballot_sync(0xffffffffu, root);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_counts[wid] = __popc(m);
    __syncthreads();
    if (root) {
        int off = block_offsets[blockIdx.x];
        for (int k = 0; k < wid; ++k) off += warp_counts[k];
        rank_of_root[i] = off + __popc(m & ((1u << lane) - 1u));
    }
}

__device__ __forceinline__ int float_key(float v) {
    int b = __float_as_int(v);
    return b >= 0 ? b : b ^ 0x7fffffff;
}

struct CompStats {        // per component, indexed by rank
    int area, minx, maxx, miny, maxy, maxkey;
};

__global__ void k_stats_init(int ncap, CompStats* __restrict__ st) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < ncap) {
        CompStats s;
        s.area = 0; s.minx = INT_MAX; s.maxx = -1; s.miny = INT_MAX; s.maxy = -1; s.maxkey = INT_MIN;
        st[k] = s;
    }
}

__global__ void k_stats(const float* __restrict__ text, int h, int w, const int* __restrict__ L,
                        const int* __restrict__ rank_of_root, CompStats* __restrict__ st, int* __restrict__ comp_of_px, int cap) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    bool inb = x < w;
    int i = y * w + x;
    int k = -1;
    if (inb && L[i] >= 0) k = rank_of_root[L[i]];
    if (k >= cap) k = -1;                                // more components than the statistics table holds: flagged by k_select2
    if (inb) comp_of_px[i] = k;
    // warp aggregation when every foreground lane belongs to the same component (the common case inside a blob)
    unsigned fgm = __ballot_sync(0xffffffffu, k >= 0);
    if (fgm == 0) return;
    int leader = __ffs(fgm) - 1;
    int k0 = __shfl_sync(0xffffffffu, k, leader);
    bool uniform = __all_sync(0xffffffffu, k < 0 || k == k0);
    int key = k >= 0 ? float_key(text[i]) : INT_MIN;
    if (uniform) {
        int mx = key;
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == leader) {
            int x0 = x - (int)(threadIdx.x & 31);
            CompStats* s = st + k0;
            atomicAdd(&s->area, __popc(fgm));
            atomicMin(&s->minx, x0 + leader);
            atomicMax(&s->maxx, x0 + 31 - __clz(fgm));
            atomicMin(&s->miny, y);
            atomicMax(&s->maxy, y);
            atomicMax(&s->maxkey, mx);
        }
    } else if (k >= 0) {
        CompStats* s = st + k;
        atomicAdd(&s->area, 1);
        atomicMin(&s->minx, x);
        atomicMax(&s->maxx, x);
        atomicMin(&s->miny, y);
        atomicMax(&s->maxy, y);
        atomicMax(&s->maxkey, key);
    }
}

// keep[k] = area >= 10 && max(text) >= text_threshold ; row_off = exclusive scan of kept heights.  Single block.
__global__ void k_select(const int* __restrict__ header, const CompStats* __restrict__ st, float text_thr,
                         int* __restrict__ row_off, int* __restrict__ header_out) {
    __shared__ int carry;
    __shared__ int s[1024];
    const int n = header[0];
    const int thr_key = float_key(text_thr);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        int k = base + threadIdx.x;
        int v = 0;
        if (k < n) {
            CompStats c = st[k];
            bool keep = c.area >= 10 && !(c.maxkey < thr_key);
            v = keep ? (c.maxy - c.miny + 1) : 0;
        }
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int add = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += add;
            __syncthreads();
        }
        if (k < n) row_off[k] = v ? carry + s[threadIdx.x] - v : -1;
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) header_out[1] = carry;                         // total rows of kept components
}

__global__ void k_rows_init(int n, int* __restrict__ rmin, int* __restrict__ rmax) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { rmin[i] = INT_MAX; rmax[i] = -1; }
}

__global__ void k_rows(int h, int w, const int* __restrict__ comp_of_px, const uint8_t* __restrict__ flags,
                       const CompStats* __restrict__ st, const int* __restrict__ row_off, int* __restrict__ rmin,
                       int* __restrict__ rmax) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    int i = y * w + x;
    int k = comp_of_px[i];
    if (k < 0 || (flags[i] & 2)) return;
    int off = row_off[k];
    if (off < 0) return;
    int r = off + y - st[k].miny;
    atomicMin(&rmin[r], x);
    atomicMax(&rmax[r], x);
}

// ---- the whole detector tail on the device --------------------------------------------------------------------------------
// header: [0] = components found, [1] = total rows of kept components, [2] = kept components, [3] = overflow flags
// (1: more components than `cap`; 2: a kept component taller than the per-block point budget; 4: more boxes than box_cap)
__global__ void k_select2(int* __restrict__ header, const CompStats* __restrict__ st, float text_thr, int cap,
                          int* __restrict__ row_off, int* __restrict__ kept_rank) {
    __shared__ int carry_rows, carry_kept;
    __shared__ int s[1024], s2[1024];
    int n = header[0];
    if (threadIdx.x == 0) {
        carry_rows = 0; carry_kept = 0;
        header[3] = n > cap ? 1 : 0;
    }
    if (n > cap) n = cap;
    const int thr_key = float_key(text_thr);
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        int k = base + threadIdx.x;
        int v = 0, kp = 0;
        if (k < n) {
            CompStats c = st[k];
            bool keep = c.area >= 10 && !(c.maxkey < thr_key);
            v = keep ? (c.maxy - c.miny + 1) : 0;
            kp = keep ? 1 : 0;
        }
        s[threadIdx.x] = v;
        s2[threadIdx.x] = kp;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int add = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            int add2 = threadIdx.x >= o ? s2[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += add;
            s2[threadIdx.x] += add2;
            __syncthreads();
        }
        if (k < n) {
            row_off[k] = kp ? carry_rows + s[threadIdx.x] - v : -1;
            kept_rank[k] = kp ? carry_kept + s2[threadIdx.x] - 1 : -1;
        }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_rows += s[1023]; carry_kept += s2[1023]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { header[1] = carry_rows; header[2] = carry_kept; }
}

constexpr int BOX_MAXPTS = 2048;          // <= 2 points per dilated row: components up to 1024 score-map rows
constexpr int BOX_THREADS = 128;

struct BoxSmem {
    geom::P2i pts[BOX_MAXPTS];
    geom::P2i sp[BOX_MAXPTS];             // the points in cv::convexHull's sort order
    unsigned long long keys[BOX_MAXPTS];
    int sorted[BOX_MAXPTS];
    int stack[BOX_MAXPTS + 2];
    int hullbuf[BOX_MAXPTS];
    geom::P2f hp[BOX_MAXPTS];             // first: dmin / dmax of the dilated rows (aliased, dead before the hull)
    geom::P2f vect[BOX_MAXPTS];
    float inv_len[BOX_MAXPTS];
    int np, l, r, t, b;
};

// One block per kept component (grid-stride): craft_utils.getDetBoxes_core from "segmap" to the rolled box.
//   segmap = label minus link-only pixels, given as per-row extents [rmin, rmax];  cv2.dilate with a (1 + niter)^2 rectangle
//   inside the clipped ROI turns row r into the span [rmin - lo, rmax + hi] on rows [r - lo, r + hi]; minAreaRect of all
//   non-zero pixels == minAreaRect of the per-row end points.
__global__ void __launch_bounds__(BOX_THREADS) k_det_boxes(int* __restrict__ header, const CompStats* __restrict__ st,
                                                           const int* __restrict__ row_off, const int* __restrict__ kept_rank,
                                                           const int* __restrict__ rmin_all, const int* __restrict__ rmax_all, int mh,
                                                           int mw, int cap, float* __restrict__ boxes, int box_cap) {
    extern __shared__ __align__(16) unsigned char box_smem_raw[];
    BoxSmem& S = *reinterpret_cast<BoxSmem*>(box_smem_raw);
    int* dmin = reinterpret_cast<int*>(S.hp);
    int* dmax = reinterpret_cast<int*>(S.vect);
    const int tid = threadIdx.x;
    int n = header[0];
    if (n > cap) n = cap;
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        const int off = row_off[k];
        if (off < 0) continue;                                   // block-uniform
        const int slot = kept_rank[k];
        if (slot >= box_cap) { if (tid == 0) atomicOr(&header[3], 4); continue; }
        const CompStats c = st[k];
        const int x = c.minx, y = c.miny, w = c.maxx - c.minx + 1, h = c.maxy - c.miny + 1, size = c.area;
        const int niter = (int)(sqrt((double)((int64_t)size * min(w, h)) / (double)((int64_t)w * h)) * 2);
        int sx = x - niter, ex = x + w + niter + 1, sy = y - niter, ey = y + h + niter + 1;
        if (sx < 0) sx = 0;
        if (sy < 0) sy = 0;
        if (ex >= mw) ex = mw;
        if (ey >= mh) ey = mh;
        const int ks = 1 + niter, anchor = ks / 2;
        const int lo = ks - 1 - anchor, hi = anchor;              // a source pixel q covers [q - lo, q + hi] on each axis
        const int R = ey - sy;
        float* out = boxes + (size_t)slot * 8;
        if (2 * R > BOX_MAXPTS) {                                // taller than the block's point budget: host fallback
            if (tid == 0) atomicOr(&header[3], 2);
            continue;
        }
        const int* rmin = rmin_all + off;
        const int* rmax = rmax_all + off;
        __syncthreads();                                         // previous component's hull arrays are dead
        for (int i = tid; i < R; i += BOX_THREADS) {
            const int yy = sy + i;
            int a = INT_MAX, e = -1;
            const int r0 = max(0, yy - hi - y), r1 = min(h - 1, yy + lo - y);
            for (int r = r0; r <= r1; ++r) {
                const int m0 = rmin[r], m1 = rmax[r];
                if (m0 > m1) continue;                            // row holds only link-only pixels
                a = min(a, max(m0 - lo, sx));
                e = max(e, min(m1 + hi, ex - 1));
            }
            dmin[i] = a;
            dmax[i] = e;
        }
        __syncthreads();
        if (tid == 0) {
            int np = 0, l = INT_MAX, r_ = -1, t = INT_MAX, b = -1;
            for (int i = 0; i < R; ++i) {
                const int a = dmin[i], e = dmax[i], yy = sy + i;
                if (a > e) continue;
                S.pts[np].x = a; S.pts[np].y = yy; ++np;
                if (e != a) { S.pts[np].x = e; S.pts[np].y = yy; ++np; }
                l = min(l, a); r_ = max(r_, e);
                t = min(t, yy); b = max(b, yy);
            }
            S.np = np; S.l = l; S.r = r_; S.t = t; S.b = b;
        }
        __syncthreads();
        const int np = S.np;
        if (np == 0) {                                           // cv2.minAreaRect of an empty set: the zero box
            if (tid < 8) out[tid] = 0.f;
            continue;
        }
        // cv::convexHull sorts the points by (x, y, position): bitonic sort of packed keys
        int npow = 1;
        while (npow < np) npow <<= 1;
        for (int i = tid; i < npow; i += BOX_THREADS)
            S.keys[i] = i < np ? (((unsigned long long)S.pts[i].x << 40) | ((unsigned long long)S.pts[i].y << 20) | (unsigned long long)i)
                               : ~0ull;
        __syncthreads();
        for (int kk = 2; kk <= npow; kk <<= 1)
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < npow; i += BOX_THREADS) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = S.keys[i], b = S.keys[ixj];
                        const bool up = (i & kk) == 0;
                        if ((a > b) == up) { S.keys[i] = b; S.keys[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        for (int i = tid; i < np; i += BOX_THREADS) {
            const int idx = (int)(S.keys[i] & 0xfffffull);
            S.sorted[i] = idx;
            S.sp[i] = S.pts[idx];
        }
        __syncthreads();
        if (tid == 0) {
            const int hn = geom::convex_hull(S.sp, S.sorted, np, S.stack, S.hullbuf, false);
            float box[8], rolled[8];
            geom::min_area_box_from_hull(S.pts, S.hullbuf, hn, S.hp, S.inv_len, S.vect, box);
            geom::finish_det_box(box, S.l, S.t, S.r, S.b, rolled);
#pragma unroll
            for (int i = 0; i < 8; ++i) out[i] = rolled[i];
        }
    }
}

// craft_utils.getDetBoxes_core, device only: boxes (n x 8 floats, label order) with ONE synchronisation.  Returns false when a
// capacity was exceeded (the caller then takes the host path det_components_dev + boxes_from_components).
bool det_boxes_dev(Handle* h, Lane& lane, const float* text, const float* link, int mh, int mw, float text_threshold,
                   float link_threshold, float low_text, std::vector<float>& boxes, int* n_labels) {
    cudaStream_t st = lane.stream;
    const int n = mh * mw;
    ARG_CHECK(n > 0 && (int64_t)mh * mw < (1ll << 30), "score map size");
    constexpr int CAP = 32768, BOX_CAP = 8192, FIRST = 1024;
    const int nblocks = cdiv(n, 1024);
    DevBuf bL((size_t)n * 4, st), bflags((size_t)n, st), bcounts((size_t)nblocks * 4, st);
    DevBuf brank((size_t)n * 4, st), bcomp((size_t)n * 4, st);
    DevBuf bstats((size_t)CAP * sizeof(CompStats), st), browoff((size_t)CAP * 4, st), bkept((size_t)CAP * 4, st);
    DevBuf brmin((size_t)n * 4, st), brmax((size_t)n * 4, st);
    DevBuf bout(16 + (size_t)BOX_CAP * 32, st);                  // header | boxes
    int* header = bout.as<int>();
    float* dboxes = reinterpret_cast<float*>(bout.as<uint8_t>() + 16);
    int* L = bL.as<int>();
    uint8_t* flags = bflags.as<uint8_t>();
    CompStats* stats = bstats.as<CompStats>();
    dim3 grd(cdiv(mw, 256), mh);
    k_ccl_init<<<grd, 256, 0, st>>>(text, link, mh, mw, low_text, link_threshold, L, flags);
    k_ccl_merge<<<grd, 256, 0, st>>>(mh, mw, L, flags);
    k_ccl_compress<<<cdiv(n, 256), 256, 0, st>>>(n, L);
    k_count_roots<<<nblocks, 1024, 0, st>>>(n, L, bcounts.as<int>());
    k_scan_blocks<<<1, 1024, 0, st>>>(nblocks, bcounts.as<int>(), header);
    k_assign_ranks<<<nblocks, 1024, 0, st>>>(n, L, bcounts.as<int>(), brank.as<int>());
    k_stats_init<<<cdiv(CAP, 256), 256, 0, st>>>(CAP, stats);
    k_stats<<<grd, 256, 0, st>>>(text, mh, mw, L, brank.as<int>(), stats, bcomp.as<int>(), CAP);
    k_select2<<<1, 1024, 0, st>>>(header, stats, text_threshold, CAP, browoff.as<int>(), bkept.as<int>());
    k_rows_init<<<cdiv(n, 256), 256, 0, st>>>(n, brmin.as<int>(), brmax.as<int>());
    k_rows<<<grd, 256, 0, st>>>(mh, mw, bcomp.as<int>(), flags, stats, browoff.as<int>(), brmin.as<int>(), brmax.as<int>());
    static std::once_flag once_attr[64];
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    std::call_once(once_attr[dev & 63], [] {
        CUDA_CHECK(cudaFuncSetAttribute(k_det_boxes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BoxSmem)));
    });
    k_det_boxes<<<2 * h->sm_count, BOX_THREADS, sizeof(BoxSmem), st>>>(header, stats, browoff.as<int>(), bkept.as<int>(), brmin.as<int>(),
                                                                    brmax.as<int>(), mh, mw, CAP, dboxes, BOX_CAP);
    count_launch(h, 12);
    CUDA_CHECK(cudaGetLastError());
    uint8_t* pin = (uint8_t*)lane.pin_out.get(16 + (size_t)BOX_CAP * 32);
    CUDA_CHECK(cudaMemcpyAsync(pin, bout.p, 16 + (size_t)FIRST * 32, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_sync(st));
    const int* hdr = reinterpret_cast<const int*>(pin);
    if (n_labels) *n_labels = hdr[0];
    if (hdr[3] != 0) return false;
    const int nk = hdr[2];
    if (nk > FIRST) {
        CUDA_CHECK(cudaMemcpyAsync(pin + 16 + (size_t)FIRST * 32, bout.as<uint8_t>() + 16 + (size_t)FIRST * 32, (size_t)(nk - FIRST) * 32,
                                   cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_sync(st));
    }
    boxes.assign(reinterpret_cast<const float*>(pin + 16), reinterpret_cast<const float*>(pin + 16) + (size_t)nk * 8);
    return true;
}

void det_components_dev(Handle* h, Lane& lane, const float* text, const float* link, int mh, int mw,
                        float text_threshold, float link_threshold, float low_text, DetComponents& out) {
    cudaStream_t st = lane.stream;
    const int n = mh * mw;
    ARG_CHECK(n > 0 && (int64_t)mh * mw < (1ll << 30), "score map size");
    const int nblocks = cdiv(n, 1024);
    DevBuf bL((size_t)n * 4, st), bflags((size_t)n, st), bcounts((size_t)nblocks * 4, st), bhdr(16, st);
    DevBuf brank((size_t)n * 4, st), bcomp((size_t)n * 4, st);
    int* L = bL.as<int>();
    uint8_t* flags = bflags.as<uint8_t>();
    dim3 grd(cdiv(mw, 256), mh);
    k_ccl_init<<<grd, 256, 0, st>>>(text, link, mh, mw, low_text, link_threshold, L, flags);
    k_ccl_merge<<<grd, 256, 0, st>>>(mh, mw, L, flags);
    k_ccl_compress<<<cdiv(n, 256), 256, 0, st>>>(n, L);
    k_count_roots<<<nblocks, 1024, 0, st>>>(n, L, bcounts.as<int>());
    k_scan_blocks<<<1, 1024, 0, st>>>(nblocks, bcounts.as<int>(), bhdr.as<int>());
    k_assign_ranks<<<nblocks, 1024, 0, st>>>(n, L, bcounts.as<int>(), brank.as<int>());
    count_launch(h, 6);
    int hdr[2] = {0, 0};
    CUDA_CHECK(cudaMemcpyAsync(hdr, bhdr.p, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_sync(st));
    const int ncomp = hdr[0];
    out = DetComponents();
    out.n_labels = ncomp;
    if (ncomp == 0) return;
    DevBuf bstats((size_t)ncomp * sizeof(CompStats), st), browoff((size_t)ncomp * 4, st);
    DevBuf brmin((size_t)n * 4, st), brmax((size_t)n * 4, st);
    CompStats* stats = bstats.as<CompStats>();
    k_stats_init<<<cdiv(ncomp, 256), 256, 0, st>>>(ncomp, stats);
    k_stats<<<grd, 256, 0, st>>>(text, mh, mw, L, brank.as<int>(), stats, bcomp.as<int>(), ncomp);
    k_select<<<1, 1024, 0, st>>>(bhdr.as<int>(), stats, text_threshold, browoff.as<int>(), bhdr.as<int>());
    k_rows_init<<<cdiv(n, 256), 256, 0, st>>>(n, brmin.as<int>(), brmax.as<int>());
    k_rows<<<grd, 256, 0, st>>>(mh, mw, bcomp.as<int>(), flags, stats, browoff.as<int>(), brmin.as<int>(), brmax.as<int>());
    count_launch(h, 5);
    CUDA_CHECK(cudaGetLastError());
    // D2H: header, stats, row offsets, then the used part of the row extents
    size_t fixed = 16 + (size_t)ncomp * (sizeof(CompStats) + 4);
    uint8_t* pin = (uint8_t*)lane.pin_out.get(fixed);
    CUDA_CHECK(cudaMemcpyAsync(pin, bhdr.p, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(pin + 16, stats, (size_t)ncomp * sizeof(CompStats), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(pin + 16 + (size_t)ncomp * sizeof(CompStats), browoff.p, (size_t)ncomp * 4,
                               cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_sync(st));
    const int total_rows = ((int*)pin)[1];
    std::vector<CompStats> hs(ncomp);
    std::vector<int> hoff(ncomp);
    memcpy(hs.data(), pin + 16, (size_t)ncomp * sizeof(CompStats));
    memcpy(hoff.data(), pin + 16 + (size_t)ncomp * sizeof(CompStats), (size_t)ncomp * 4);
    out.row_min.resize(total_rows);
    out.row_max.resize(total_rows);
    if (total_rows > 0) {
        int* pr = (int*)lane.pin_out.get((size_t)total_rows * 8);
        CUDA_CHECK(cudaMemcpyAsync(pr, brmin.p, (size_t)total_rows * 4, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(pr + total_rows, brmax.p, (size_t)total_rows * 4, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_sync(st));
        memcpy(out.row_min.data(), pr, (size_t)total_rows * 4);
        memcpy(out.row_max.data(), pr + total_rows, (size_t)total_rows * 4);
    }
    for (int k = 0; k < ncomp; ++k) {
        if (hoff[k] < 0) continue;
        out.comp_x.push_back(hs[k].minx);
        out.comp_y.push_back(hs[k].miny);
        out.comp_w.push_back(hs[k].maxx - hs[k].minx + 1);
        out.comp_h.push_back(hs[k].maxy - hs[k].miny + 1);
        out.comp_area.push_back(hs[k].area);
        out.row_off.push_back(hoff[k]);
    }
}

}  // namespace bbocr
