// conv_tc.cu -- implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a):
//     D[128 pixels x BN channels] (FP32, TMEM)  +=  A[128 x BK] (bf16, smem)  x  B[BN x BK]^T (bf16, smem)
// for every filter tap and every BK-channel block.  No im2col buffer exists anywhere: the A tile of tap (ky,kx) is
// fetched by ONE TMA tiled load of the (BK ch x TW x TH x 1) box at spatial offset (kx*dil - pad, ky*dil - pad) of the NHWC
// activation tensor; TMA's out-of-bounds zero fill IS the convolution's zero padding.  1x1 convolutions / Linear layers
// use a flat 2-D view (channels x pixels).  Channel concatenation (U-Net skips) is two tensor maps walked back to back.
//
// Persistent CTAs (static round-robin over output tiles) with warp roles
//   warp 0    : TMA producer (one lane)       -> full[stage]   (mbarrier expect_tx / complete_tx)
//   warp 1    : tcgen05.mma issuer (one lane) -> empty[stage]  (tcgen05.commit); tfull[acc] after a tile's last k-block
//   warp 2    : TMEM allocate / free (2 accumulator stages of BN columns: tile i+1 accumulates while tile i drains)
//   warps 4-7 : epilogue: tcgen05.ld 32x32b -> folded-BN scale/bias (+ReLU) -> optional fused 2x2 / 2x1 max-pool by
//               warp shuffles -> bf16 / fp32 NHWC stores -> tempty[acc]
//
// Operand tiles are the canonical K-major SWIZZLE_128B (BK = 64) / SWIZZLE_64B (BK = 32) layouts that the TMA box
// produces directly; the UMMA shared-memory descriptors walk them in 32-byte (UMMA_K = 16) steps.
#include <cuda.h>

#include "engine.h"

namespace bbocr {

namespace {

constexpr int BM = 128;

struct TcParams {
    int C1, C2;                 // channels of the two concatenated inputs (C2 may be 0)
    int w_lo_off;               // PAIR: K offset of the w_lo half inside a weight row [w_hi (cin) | w_lo (cin)]
    int taps_w, taps;           // kw, kh*kw
    int pad, dil;
    int TW, TH, tiles_x, tiles_y;
    int OH, OW, NIMG;
    int flat;                   // 1: 1x1 conv over a flat [pixels][channels] view
    int64_t M;                  // total output pixels
    int cout, BN, n_tiles;      // n_tiles = cout_pad / BN
    int m_tiles, total_tiles;
    int relu, out_f32;
    int pool;                   // 0 none, 1 = MaxPool2d(2,2), 2 = MaxPool2d((2,1),(2,1))  (tile must be 16 wide)
    int write_full;             // store the un-pooled tensor as well (skip connections)
    void* out;                  // un-pooled output [N][OH][OW][cout]
    void* out2;                 // pooled output    [N][OH/2][OW/(pool==1?2:1)][cout]
    void* out_lo;               // split-precision output: out = bf16 hi part, out_lo = bf16(v - hi)
    void* out2_lo;
    int split_out;
    const float* scale;
    const float* bias;
    const uint8_t* colmask;     // optional [OW]: output columns with 0 are written as zeros (gaps between concatenated crops)
    int stages;
    int p_stages, patch_al;     // halo-patch kernel: patch pipeline depth, bytes of one (1024-aligned) patch
    int ncat;                   // split precision, BN <= 128: the w_hi and w_lo tiles sit back to back in shared memory and are
                                // used as ONE B operand of 2*BN rows, so a K = 16 step is two MMAs instead of three:
                                //   D[:, 0:2BN] += x_hi * [w_hi ; w_lo]      D[:, 0:BN] += x_lo * w_hi
                                // (same products, one A-operand read less per step: the low-channel layers are bound by the
                                // operand reads from shared memory, not by the math); the epilogue adds the two column halves
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(=1)<<16 | SBO>>4<<32 |
// version(=1)<<46 | layout<<61 ; SWIZZLE_128B = 2 (8 rows x 128 B atoms, SBO = 1024), SWIZZLE_64B = 4 (SBO = 512)
template <int BK>
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    constexpr uint64_t layout = (BK == 64) ? 2 : 4;
    constexpr uint64_t sbo = (BK == 64) ? 1024 : 512;
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait: the epilogue issues the load of column group c + 1 before it works on group c
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct TileCoord {
    int n0, img, x0, y0;
    int64_t m0;
};
__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, int tile) {
    TileCoord t;
    const int mt = tile / p.n_tiles;
    t.n0 = (tile - mt * p.n_tiles) * p.BN;        // N tiles of one M tile run on neighbouring CTAs: the A tile is shared in L2
    t.img = 0; t.x0 = 0; t.y0 = 0; t.m0 = 0;
    if (p.flat) {
        t.m0 = (int64_t)mt * BM;
    } else {
        int q = mt;
        int tx = q % p.tiles_x;
        q /= p.tiles_x;
        int ty = q % p.tiles_y;
        t.img = q / p.tiles_y;
        t.x0 = tx * p.TW;
        t.y0 = ty * p.TH;
    }
    return t;
}

// 256-bit global store (sm_100: STG.E.256): one 32-byte sector per request instead of two 16-byte halves
__device__ __forceinline__ void st_global_256(void* p, const uint32_t* w) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

template <typename TO>
__device__ __forceinline__ void store16(TO* o, const float* f, int nbase, int cout);
template <>
__device__ __forceinline__ void store16<float>(float* o, const float* f, int nbase, int cout) {
    if ((cout & 7) == 0 && nbase + 16 <= cout && (reinterpret_cast<uintptr_t>(o) & 31) == 0) {
        st_global_256(o, reinterpret_cast<const uint32_t*>(f));
        st_global_256(o + 8, reinterpret_cast<const uint32_t*>(f) + 8);
    } else if ((cout & 3) == 0 && nbase + 16 <= cout) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else {
        for (int j = 0; j < 16; ++j)
            if (nbase + j < cout) o[j] = f[j];
    }
}
template <>
__device__ __forceinline__ void store16<__nv_bfloat16>(__nv_bfloat16* o, const float* f, int nbase, int cout) {
    if ((cout & 7) == 0 && nbase + 16 <= cout) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&t);
        }
        if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) {
            st_global_256(o, w);
        } else {
            *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(o + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        }
    } else {
        for (int j = 0; j < 16; ++j)
            if (nbase + j < cout) o[j] = __float2bfloat16_rn(f[j]);
    }
}

// split-precision store: hi = bf16(v), lo = bf16(v - hi).  Packed conversions (cvt.rn.bf16x2.f32: 8 + 8 per group instead of 16 scalar
// conversions and 16 packs); al32 = every 16-channel group of both tensors starts on a 32-byte boundary (a property of the layer,
// uniform over the grid: the 256-bit store needs no per-thread alignment branch)
__device__ __forceinline__ void store_split(void* hi_base, void* lo_base, int64_t off, const float* f, int nbase, int cout, bool al32) {
    __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(hi_base) + off;
    __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(lo_base) + off;
    if ((cout & 7) == 0 && nbase + 16 <= cout) {
        uint32_t wh[8], wl[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            const uint32_t hw = *reinterpret_cast<const uint32_t*>(&h2);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(f[2 * j] - __uint_as_float(hw << 16), f[2 * j + 1] - __uint_as_float(hw & 0xffff0000u));
            wh[j] = hw;
            wl[j] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        if (al32) {
            st_global_256(oh, wh);
            st_global_256(ol, wl);
        } else {
            *reinterpret_cast<uint4*>(oh) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
            *reinterpret_cast<uint4*>(oh + 8) = make_uint4(wh[4], wh[5], wh[6], wh[7]);
            *reinterpret_cast<uint4*>(ol) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
            *reinterpret_cast<uint4*>(ol + 8) = make_uint4(wl[4], wl[5], wl[6], wl[7]);
        }
    } else {
        for (int j = 0; j < 16; ++j)
            if (nbase + j < cout) {
                const __nv_bfloat16 hb = __float2bfloat16_rn(f[j]);
                oh[j] = hb;
                ol[j] = __float2bfloat16_rn(f[j] - __bfloat162float(hb));
            }
    }
}

// Epilogue warps (4..7) of both kernels: tcgen05.ld 32x32b -> folded-BN scale/bias (+ReLU) -> optional fused 2x2 / 2x1 max-pool
// by warp shuffles -> bf16 / fp32 / split NHWC stores -> tempty[acc].  TMEM lane r = pixel (r / TW, r % TW) of the tile; with
// TW in {8, 16} the 2x2 (2x1) pooling window lives in lanes {l, l^1, l^TW, l^TW^1} ({l, l^TW}).
// EW = 8 (kernels that run one CTA per SM with BN = 128): two warps per TMEM lane quarter, each takes half of the columns.
template <bool NCAT, int EW = 4>
__device__ __forceinline__ void tc_epilogue(const TcParams& p, uint32_t tmem_base, int warp, int lane, uint64_t* tfull_bar,
                                            uint64_t* tempty_bar, const float* s_sb /* shared [scale 256 | bias 256] or nullptr */,
                                            float* s_pool = nullptr /* shared, 512 floats per epilogue warp, or nullptr */) {
        const int wq = warp & 3;
        const int r = wq * 32 + lane;
        // 2x2 pooling through shared memory (s_pool): lane = (window w, column quad q).  Window w of the warp's 32 pixels =
        // lanes {b, b^1, b^TW, b^TW^1} with b = w's bits spread around bit 0 and bit log2(TW)
        float* my_pool = s_pool ? s_pool + (warp - 4) * 512 : nullptr;
        const bool al32 = (p.cout & 15) == 0 && ((reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.out_lo) |
                                                  reinterpret_cast<uintptr_t>(p.out2) | reinterpret_cast<uintptr_t>(p.out2_lo)) & 31) == 0;
        const int pw = lane >> 2, pq = lane & 3;
        const int pb = ((pw << 1) & (p.TW - 1)) | ((((pw << 1) & ~(p.TW - 1))) << 1);
        int ti = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
            const TileCoord tc = tile_coord(p, tile);
            const int as = ti & 1;
            const uint32_t aph = (ti >> 1) & 1;
            mbar_wait(&tfull_bar[as], aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            int64_t pix = -1, pix2 = -1;
            bool colvalid = true;
            if (p.flat) {
                if (tc.m0 + r < p.M) pix = tc.m0 + r;
            } else {
                const int hl = r / p.TW, wl = r - hl * p.TW;
                const int y = tc.y0 + hl, x = tc.x0 + wl;
                if (y < p.OH && x < p.OW) pix = ((int64_t)tc.img * p.OH + y) * p.OW + x;
                if (p.colmask && x < p.OW) colvalid = __ldg(p.colmask + x) != 0;
                if (p.pool) {
                    // lane = (hl % (32 / TW)) * TW + wl ; the 2x2 (2x1) window lives in lanes {l, l^1, l^TW, l^TW^1} ({l, l^TW})
                    const int POH = p.OH >> 1, POW = p.pool == 1 ? p.OW >> 1 : p.OW;
                    const int py = y >> 1, px = p.pool == 1 ? x >> 1 : x;
                    const bool writer = (lane & p.TW) == 0 && (p.pool == 2 || (lane & 1) == 0);
                    if (writer && py < POH && px < POW) pix2 = ((int64_t)tc.img * POH + py) * POW + px;
                }
            }
            // pooled pixel of the window this lane serves on the shared-memory pooling path (the window's base lane computed it)
            const int64_t pix2w = (p.pool == 1 && my_pool) ? __shfl_sync(0xffffffffu, pix2, pb) : -1;
            const uint32_t trow = tmem_base + (uint32_t)(as * (NCAT ? 2 * p.BN : p.BN)) + ((uint32_t)(wq * 32) << 16);
            uint32_t vn[16];
            const int cw = p.BN / (EW / 4), c_lo = ((warp - 4) >> 2) * cw, c_hi = c_lo + cw;      // this warp's column range
            tmem_ld16_nowait(trow + c_lo, vn);
            for (int c = c_lo; c < c_hi; c += 16) {
                uint32_t v[16];
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = vn[j];
                if (NCAT) {                         // x_hi * w_lo accumulated in the second half of the stage
                    // (loading this half one group ahead like the first one was tried: 120 registers, conv2_2 497 -> 535 us)
                    uint32_t v2[16];
                    tmem_ld16(trow + p.BN + c, v2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
                }
                if (c + 16 < c_hi) tmem_ld16_nowait(trow + c + 16, vn);     // in flight while this group is scaled, split and stored
                float f[16], sc[16], bi[16];
                if (s_sb) {                         // one N tile: the layer's scale / bias sit in shared memory (8 LDS.128 per group
                                                    // instead of 32 uniform global loads: those were ~40 % of the epilogue's stalls)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 a4 = *reinterpret_cast<const float4*>(s_sb + c + 4 * q);
                        const float4 b4 = *reinterpret_cast<const float4*>(s_sb + 256 + c + 4 * q);
                        sc[4 * q] = a4.x; sc[4 * q + 1] = a4.y; sc[4 * q + 2] = a4.z; sc[4 * q + 3] = a4.w;
                        bi[4 * q] = b4.x; bi[4 * q + 1] = b4.y; bi[4 * q + 2] = b4.z; bi[4 * q + 3] = b4.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        sc[j] = __ldg(p.scale + tc.n0 + c + j);                     // padded to cout_pad
                        bi[j] = __ldg(p.bias + tc.n0 + c + j);
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float a = fmaf(__uint_as_float(v[j]), sc[j], bi[j]);
                    f[j] = p.relu ? fmaxf(a, 0.f) : a;
                }
                if (p.colmask && !colvalid) {       // recogniser strips only: gaps between concatenated crops
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = 0.f;
                }
                const int nbase = tc.n0 + c;
                if (pix >= 0 && (!p.pool || p.write_full)) {
                    if (p.split_out) store_split(p.out, p.out_lo, pix * p.cout + nbase, f, nbase, p.cout, al32);
                    else if (p.out_f32) store16(reinterpret_cast<float*>(p.out) + pix * p.cout + nbase, f, nbase, p.cout);
                    else store16(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.cout + nbase, f, nbase, p.cout);
                }
                if (p.pool == 1 && my_pool) {
                    // the warp's 32 x 16 activations go through shared memory (row = lane, 64 bytes, 16-byte chunk k at
                    // k ^ ((lane >> 1) & 3)); lane (w, q) reads columns 4q .. 4q+3 of its window's four pixels: 4 + 4 128-bit
                    // shared accesses and 12 max instead of 32 shuffles and 32 max, and every lane splits / stores 4 pooled
                    // values instead of a quarter of the lanes 16 (the others computing theirs for nothing)
                    const int sw = (lane >> 1) & 3;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        *reinterpret_cast<float4*>(my_pool + lane * 16 + ((k ^ sw) << 2)) = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
                    __syncwarp();
                    float4 m4;
                    {
                        const int r0 = pb, r1 = pb ^ 1, r2 = pb ^ p.TW, r3 = pb ^ p.TW ^ 1;
                        const float4 a0 = *reinterpret_cast<const float4*>(my_pool + r0 * 16 + ((pq ^ ((r0 >> 1) & 3)) << 2));
                        const float4 a1 = *reinterpret_cast<const float4*>(my_pool + r1 * 16 + ((pq ^ ((r1 >> 1) & 3)) << 2));
                        const float4 a2 = *reinterpret_cast<const float4*>(my_pool + r2 * 16 + ((pq ^ ((r2 >> 1) & 3)) << 2));
                        const float4 a3 = *reinterpret_cast<const float4*>(my_pool + r3 * 16 + ((pq ^ ((r3 >> 1) & 3)) << 2));
                        m4.x = fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x));
                        m4.y = fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y));
                        m4.z = fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z));
                        m4.w = fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w));
                    }
                    __syncwarp();                                   // the next group overwrites the staging rows
                    if (pix2w >= 0) {
                        const float m[4] = {m4.x, m4.y, m4.z, m4.w};
                        const int nb = nbase + 4 * pq;
                        const int64_t off = pix2w * p.cout + nb;
                        if ((p.cout & 3) == 0 && nb + 4 <= p.cout) {
                            if (p.split_out) {
                                const __nv_bfloat162 h01 = __floats2bfloat162_rn(m[0], m[1]), h23 = __floats2bfloat162_rn(m[2], m[3]);
                                const uint32_t w01 = *reinterpret_cast<const uint32_t*>(&h01), w23 = *reinterpret_cast<const uint32_t*>(&h23);
                                const __nv_bfloat162 l01 = __floats2bfloat162_rn(m[0] - __uint_as_float(w01 << 16), m[1] - __uint_as_float(w01 & 0xffff0000u));
                                const __nv_bfloat162 l23 = __floats2bfloat162_rn(m[2] - __uint_as_float(w23 << 16), m[3] - __uint_as_float(w23 & 0xffff0000u));
                                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + off) = make_uint2(w01, w23);
                                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2_lo) + off) =
                                    make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
                            } else if (p.out_f32) {
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out2) + off) = m4;
                            } else {
                                const __nv_bfloat162 h01 = __floats2bfloat162_rn(m[0], m[1]), h23 = __floats2bfloat162_rn(m[2], m[3]);
                                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + off) =
                                    make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                            }
                        } else {
                            for (int j = 0; j < 4; ++j) {
                                if (nb + j >= p.cout) break;
                                if (p.split_out) {
                                    const __nv_bfloat16 hb = __float2bfloat16_rn(m[j]);
                                    reinterpret_cast<__nv_bfloat16*>(p.out2)[off + j] = hb;
                                    reinterpret_cast<__nv_bfloat16*>(p.out2_lo)[off + j] = __float2bfloat16_rn(m[j] - __bfloat162float(hb));
                                } else if (p.out_f32) {
                                    reinterpret_cast<float*>(p.out2)[off + j] = m[j];
                                } else {
                                    reinterpret_cast<__nv_bfloat16*>(p.out2)[off + j] = __float2bfloat16_rn(m[j]);
                                }
                            }
                        }
                    }
                } else if (p.pool) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float m = f[j];
                        if (p.pool == 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, p.TW));
                        f[j] = m;
                    }
                    if (pix2 >= 0) {
                        if (p.split_out) store_split(p.out2, p.out2_lo, pix2 * p.cout + nbase, f, nbase, p.cout, al32);
                        else if (p.out_f32) store16(reinterpret_cast<float*>(p.out2) + pix2 * p.cout + nbase, f, nbase, p.cout);
                        else store16(reinterpret_cast<__nv_bfloat16*>(p.out2) + pix2 * p.cout + nbase, f, nbase, p.cout);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
    }

// PAIR = 1: split-precision operands.  A pipeline stage holds the hi AND lo halves of the activation tile and of the
// weight tile (four TMA boxes), and every K = 16 step issues three MMAs into the same accumulator:
//     x_hi * w_hi  +  x_lo * w_hi  +  x_hi * w_lo        (the dropped x_lo * w_lo term is 2^-16-class)
// so the operand stream through L2 -> SM is 2x the single-precision stream for 3x the tensor work.
// Maps: PAIR = 0: tmA1 = in1, tmA2 = in2 (channel concat).  PAIR = 1: tmA1/tmA2 = in1 hi/lo, tmA3/tmA4 = in2 hi/lo.
template <int BK, int PAIR, bool NCAT>
__global__ void __launch_bounds__(256, 2) k_conv_tc(const __grid_constant__ CUtensorMap tmA1,
                                                 const __grid_constant__ CUtensorMap tmA2,
                                                 const __grid_constant__ CUtensorMap tmA3,
                                                 const __grid_constant__ CUtensorMap tmA4,
                                                 const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[8], empty_bar[8], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_sb[512];            // scale[256] | bias[256] of a layer with one N tile
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int A_BYTES = BM * BK * 2;
    const int B_BYTES = p.BN * BK * 2;
    const int STAGE_BYTES = (PAIR ? 2 : 1) * (A_BYTES + B_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb1 = p.C1 / BK, kb2 = p.C2 / BK;
    const int kiters = p.taps * (kb1 + kb2);
    uint32_t ncols = 32;
    while ((int)ncols < (NCAT ? 4 : 2) * p.BN) ncols <<= 1;
    const bool sb_res = p.n_tiles == 1;                  // BN = cout_pad <= 256
    if (sb_res)
        for (int i = threadIdx.x; i < p.BN; i += blockDim.x) { s_sb[i] = p.scale[i]; s_sb[256 + i] = p.bias[i]; }

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer ----------------
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const TileCoord tc = tile_coord(p, tile);
            for (int tap = 0; tap < p.taps; ++tap) {
                const int ky = tap / p.taps_w, kx = tap - ky * p.taps_w;
                const int cx = tc.x0 - p.pad + kx * p.dil, cy = tc.y0 - p.pad + ky * p.dil;
                for (int kb = 0; kb < kb1 + kb2; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + (size_t)s * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], (uint32_t)STAGE_BYTES);
                    const int seg = kb < kb1 ? 0 : 1;
                    const int c0 = (seg == 0 ? kb : kb - kb1) * BK;
                    if (PAIR) {
                        const CUtensorMap* th = seg == 0 ? &tmA1 : &tmA3;
                        const CUtensorMap* tl = seg == 0 ? &tmA2 : &tmA4;
                        uint8_t* sb = sa + 2 * A_BYTES;
                        if (p.flat) {
                            tma_load_2d(sa, th, &full_bar[s], c0, (int)tc.m0);
                            tma_load_2d(sa + A_BYTES, tl, &full_bar[s], c0, (int)tc.m0);
                        } else {
                            tma_load_4d(sa, th, &full_bar[s], c0, cx, cy, tc.img);
                            tma_load_4d(sa + A_BYTES, tl, &full_bar[s], c0, cx, cy, tc.img);
                        }
                        tma_load_3d(sb, &tmB, &full_bar[s], kb * BK, tc.n0, tap);
                        tma_load_3d(sb + B_BYTES, &tmB, &full_bar[s], p.w_lo_off + kb * BK, tc.n0, tap);
                    } else {
                        const CUtensorMap* tma = seg == 0 ? &tmA1 : &tmA2;
                        uint8_t* sb = sa + A_BYTES;
                        if (p.flat) tma_load_2d(sa, tma, &full_bar[s], c0, (int)tc.m0);
                        else tma_load_4d(sa, tma, &full_bar[s], c0, cx, cy, tc.img);
                        tma_load_3d(sb, &tmB, &full_bar[s], kb * BK, tc.n0, tap);
                    }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer ----------------
        // instruction descriptor: D = F32 (1<<4), A = B = BF16 (1<<7, 1<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        int it = 0, ti = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
            const int as = ti & 1;
            const uint32_t aph = (ti >> 1) & 1;
            mbar_wait(&tempty_bar[as], aph ^ 1);                 // the epilogue has drained this accumulator stage
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_acc = tmem_base + (uint32_t)(as * (NCAT ? 2 * p.BN : p.BN));
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * p.BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int k = 0; k < kiters; ++k, ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
                if (PAIR) {
                    const uint64_t ahi = umma_desc<BK>(sa), alo = umma_desc<BK>(sa + A_BYTES);
                    const uint64_t bhi = umma_desc<BK>(sa + 2 * A_BYTES), blo = umma_desc<BK>(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        if (NCAT) {
                            umma_bf16(tmem_acc, ahi + 2 * kk, bhi + 2 * kk, idesc2, (k > 0 || kk > 0) ? 1u : 0u);
                            umma_bf16(tmem_acc, alo + 2 * kk, bhi + 2 * kk, idesc, 1u);
                        } else {
                            umma_bf16(tmem_acc, ahi + 2 * kk, bhi + 2 * kk, idesc, (k > 0 || kk > 0) ? 1u : 0u);
                            umma_bf16(tmem_acc, alo + 2 * kk, bhi + 2 * kk, idesc, 1u);
                            umma_bf16(tmem_acc, ahi + 2 * kk, blo + 2 * kk, idesc, 1u);
                        }
                    }
                } else {
                    const uint64_t adesc = umma_desc<BK>(sa), bdesc = umma_desc<BK>(sa + A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk)      // +32 bytes (2 x 16 B units) per UMMA_K = 16 step inside the swizzle atom
                        umma_bf16(tmem_acc, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tfull_bar[as]);
        }
    } else if (warp >= 4) {
        tc_epilogue<NCAT>(p, tmem_base, warp, lane, tfull_bar, tempty_bar, sb_res ? s_sb : nullptr);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// ---- halo-patch variant for split-precision 3x3 / pad 1 convolutions ------------------------------------------------------
// k_conv_tc fetches one shifted 128-pixel A tile per filter tap: nine tiles (hi AND lo) per k-block for an output tile whose
// taps overlap almost completely.  Here the input patch of an 8 x 16 output tile -- 10 x 18 pixels with the 1-px halo, TMA
// out-of-bounds zero fill = the padding -- is loaded ONCE per k-block (hi + lo), and tap (ky, kx) is a shifted UMMA descriptor
// into it: start = patch + (ky * 10 + kx) * PIX, stride between the 8-row groups (SBO) = one patch row (the trick of
// conv_res.cu; SWIZZLE_128B / 64B are functions of absolute shared-memory address bits for the TMA write and the tensor-core
// read alike).  Only the weight tiles (hi + lo per tap) stream through the ring.  A traffic per k-block: 2 x 23 KB instead of
// 2 x 9 x 16 KB; the low-channel layers of the detector (Cout <= 128), where the A tile dominates the operand stream, leave
// the L2 -> SM limit.  Three MMAs per K = 16 step as in k_conv_tc<.., PAIR = 1>.
__device__ __forceinline__ uint64_t desc_kmajor_sbo(uint32_t saddr, uint32_t sbo_bytes, uint64_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (layout << 61);
}

template <int BK, bool NCAT, int EW = 4>
__global__ void __launch_bounds__(128 + 32 * EW, EW == 4 ? 2 : 1) k_conv_tc_patch(const __grid_constant__ CUtensorMap tmA1,
                                                       const __grid_constant__ CUtensorMap tmA2,
                                                       const __grid_constant__ CUtensorMap tmA3,
                                                       const __grid_constant__ CUtensorMap tmA4,
                                                       const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[8], empty_bar[8], pfull_bar[4], pempty_bar[4], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_sb[512];            // scale[256] | bias[256] of a layer with one N tile
    __shared__ __align__(16) float s_pool[EW * 512];     // 2x2 pooling: 32 pixels x 16 columns per epilogue warp
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int PIX = BK * 2;                              // bytes per pixel row of the operand tiles
    constexpr int PW = 10, PH = 18;                          // patch: 8 x 16 output pixels + halo
    constexpr int PPITCH = PW * PIX;
    constexpr int PATCH_BYTES = PH * PPITCH;
    constexpr uint64_t LAYOUT = BK == 64 ? 2 : 4;
    const int B_BYTES = p.BN * BK * 2;
    const int PSTAGE = 2 * p.patch_al, RSTAGE = 2 * B_BYTES;
    uint8_t* ring = smem + (size_t)p.p_stages * PSTAGE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb1 = p.C1 / BK, kb2 = p.C2 / BK;
    uint32_t ncols = 32;
    while ((int)ncols < (NCAT ? 4 : 2) * p.BN) ncols <<= 1;
    const bool sb_res = p.n_tiles == 1;                  // BN = cout_pad <= 256
    if (sb_res)
        for (int i = threadIdx.x; i < p.BN; i += blockDim.x) { s_sb[i] = p.scale[i]; s_sb[256 + i] = p.bias[i]; }

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < p.p_stages; ++s) { mbar_init(&pfull_bar[s], 1); mbar_init(&pempty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer: one (hi, lo) patch per (tile, k-block), one (hi, lo) weight tile per tap ----------------
        // The patch of item j + 1 is requested in the MIDDLE of item j's nine weight tiles: late enough that its stage has been
        // released (the MMA issuer runs `stages` ring slots behind the producer, so item j - 1 has retired by tap `stages`),
        // early enough to land before item j + 1 starts (requested after the nine tiles it stalled every tile switch).
        const int nkb = kb1 + kb2;
        const int my_tiles = blockIdx.x < p.total_tiles ? (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const int items = my_tiles * nkb;
        auto load_patch = [&](int j) {
            const int tile = blockIdx.x + (j / nkb) * gridDim.x, kb = j % nkb;
            const TileCoord tc = tile_coord(p, tile);
            const int ps = j % p.p_stages;
            mbar_wait(&pempty_bar[ps], ((j / p.p_stages) & 1) ^ 1);
            const int seg = kb < kb1 ? 0 : 1;
            const int c0 = (seg == 0 ? kb : kb - kb1) * BK;
            uint8_t* pa = smem + (size_t)ps * PSTAGE;
            mbar_expect_tx(&pfull_bar[ps], (uint32_t)(2 * PATCH_BYTES));
            tma_load_4d(pa, seg == 0 ? &tmA1 : &tmA3, &pfull_bar[ps], c0, tc.x0 - 1, tc.y0 - 1, tc.img);
            tma_load_4d(pa + p.patch_al, seg == 0 ? &tmA2 : &tmA4, &pfull_bar[ps], c0, tc.x0 - 1, tc.y0 - 1, tc.img);
        };
        int it = 0;
        if (items > 0) load_patch(0);
        const int tap_pf = p.stages < 8 ? p.stages : 8;
        for (int j = 0; j < items; ++j) {
            const int tile = blockIdx.x + (j / nkb) * gridDim.x, kb = j % nkb;
            const int n0 = (tile % p.n_tiles) * p.BN;
            for (int tap = 0; tap < 9; ++tap, ++it) {
                const int s = it % p.stages;
                mbar_wait(&empty_bar[s], ((it / p.stages) & 1) ^ 1);
                uint8_t* sb = ring + (size_t)s * RSTAGE;
                mbar_expect_tx(&full_bar[s], (uint32_t)RSTAGE);
                tma_load_3d(sb, &tmB, &full_bar[s], kb * BK, n0, tap);
                tma_load_3d(sb + B_BYTES, &tmB, &full_bar[s], p.w_lo_off + kb * BK, n0, tap);
                if (tap == tap_pf && j + 1 < items) load_patch(j + 1);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        int it = 0, pit = 0, ti = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
            const int as = ti & 1;
            mbar_wait(&tempty_bar[as], ((ti >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_acc = tmem_base + (uint32_t)(as * (NCAT ? 2 * p.BN : p.BN));
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * p.BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < kb1 + kb2; ++kb, ++pit) {
                const int ps = pit % p.p_stages;
                mbar_wait(&pfull_bar[ps], (pit / p.p_stages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t pa = smem_u32(smem + (size_t)ps * PSTAGE);
                const uint64_t ahi0 = desc_kmajor_sbo(pa, PPITCH, LAYOUT), alo0 = desc_kmajor_sbo(pa + p.patch_al, PPITCH, LAYOUT);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap, ++it) {
                    const int s = it % p.stages;
                    mbar_wait(&full_bar[s], (it / p.stages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sb = smem_u32(ring + (size_t)s * RSTAGE);
                    const uint64_t bhi = umma_desc<BK>(sb), blo = umma_desc<BK>(sb + B_BYTES);
                    const int ky = tap / 3, kx = tap % 3;
                    const uint64_t off = (uint64_t)(((ky * PW + kx) * PIX) >> 4);
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        if (NCAT) {
                            umma_bf16(tmem_acc, ahi0 + off + 2 * kk, bhi + 2 * kk, idesc2, (kb > 0 || tap > 0 || kk > 0) ? 1u : 0u);
                            umma_bf16(tmem_acc, alo0 + off + 2 * kk, bhi + 2 * kk, idesc, 1u);
                        } else {
                            umma_bf16(tmem_acc, ahi0 + off + 2 * kk, bhi + 2 * kk, idesc, (kb > 0 || tap > 0 || kk > 0) ? 1u : 0u);
                            umma_bf16(tmem_acc, alo0 + off + 2 * kk, bhi + 2 * kk, idesc, 1u);
                            umma_bf16(tmem_acc, ahi0 + off + 2 * kk, blo + 2 * kk, idesc, 1u);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&pempty_bar[ps]);
            }
            umma_commit(&tfull_bar[as]);
        }
    } else if (warp >= 4) {
        tc_epilogue<NCAT, EW>(p, tmem_base, warp, lane, tfull_bar, tempty_bar, sb_res ? s_sb : nullptr, s_pool);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// ---- fused detector stem (bf16x3): gather + conv1_1 in ONE kernel -----------------------------------------------------------------
// k_im2col_rgb_split wrote the gathered 32-channel stem (hi + lo: 354 MB per 1920x1440 page) and conv1_1 read it back through
// TMA.  Here k_stem_norm writes the normalised canvas ONCE, 16 bytes per pixel ((hi | lo << 16) per channel: the very words
// k_im2col_rgb_split stored, 44 MB per page, L2-resident), and four producer warps build the A tiles in shared memory
// themselves: thread r of a tile fetches the nine neighbours of pixel m0 + r with one 128-bit load each and writes its 64-byte
// hi and lo rows in the K-major SWIZZLE_64B layout the UMMA descriptor expects (16-byte chunk c of row r at chunk
// c ^ ((r >> 1) & 3): the swizzle TMA would have applied).  The weight tile (w_hi | w_lo, 8 KB) is resident; MMAs and epilogue
// are those of k_conv_tc<32, 1, false>, so the layer's output is bit-identical to the two-kernel path.
// (First version: 27 byte loads + 27 table look-ups per pixel in the producer -- ~530 instructions per tile on one warp,
// 10 k cycles per tile, 390 us per page, no faster than the two kernels it replaced.)
struct StemParams {
    const uint4* norm;               // [nimg][H][W] normalised canvas, {c0, c1, c2, 0} with c = bf16 hi | bf16 lo << 16
    int H, W;
};

constexpr int STEM_STAGES = 2;
constexpr int STEM_OUT_TILE = BM * 128;        // one staged output tile: 128 pixels x 64 channels bf16 (hi or lo)

// u8 image (th x tw x 3, placed top-left on the H x W canvas) -> normalised canvas; canvas pixels outside the image hold the
// normalised value of 0 like upstream's zero-filled canvas (imgproc.resize_aspect_ratio + normalizeMeanVariance)
__global__ void __launch_bounds__(256) k_stem_norm(const uint8_t* __restrict__ img, int th, int tw, uint4* __restrict__ out, int H, int W,
                                                   float m0, float m1, float m2, float s0, float s1, float s2) {
    __shared__ uint32_t lut[3][256];                                    // hi | lo << 16
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
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    uint32_t v[3];
    const bool in_img = y < th && x < tw;
    const uint8_t* px = img + ((int64_t)y * tw + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = lut[c][in_img ? (int)__ldg(px + c) : 0];
    out[(int64_t)y * W + x] = make_uint4(v[0], v[1], v[2], 0u);
}

// Epilogue (64 output channels, split hi / lo): the four epilogue warps stage the tile in shared memory in the SWIZZLE_128B image
// (pixel row = 128 bytes; 16-byte chunk q of row r at chunk q ^ (r & 7): conflict-free 128-bit shared stores) and one thread
// hands it to TMA: full 128-byte lines leave the SM instead of one 32-byte sector per lane and request.  Two staging buffers
// (the accumulator stage picks one), so the store of tile t is read out of shared memory while tile t + 1 is computed.
__global__ void __launch_bounds__(384, 2) k_conv_stem(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOh,
                                                      const __grid_constant__ CUtensorMap tmOl, const TcParams p, const StemParams sp) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t afull_bar[STEM_STAGES], aempty_bar[STEM_STAGES], bfull_bar, tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_scale[64], s_bias[64];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int BK = 32;
    constexpr int A_BYTES = BM * BK * 2;                                // 8 KB per (hi or lo) tile
    const int B_BYTES = p.BN * BK * 2;                                  // BN = 64: 4 KB
    uint8_t* sB = smem + (size_t)STEM_STAGES * 2 * A_BYTES;
    uint8_t* sOut = sB + 8192;                                          // [2 buffers][hi | lo][128 rows x 128 B]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t ncols = 32;
    while ((int)ncols < 2 * p.BN) ncols <<= 1;
    if (threadIdx.x < 64) {
        s_scale[threadIdx.x] = p.scale[threadIdx.x];
        s_bias[threadIdx.x] = p.bias[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < STEM_STAGES; ++s) { mbar_init(&afull_bar[s], 128); mbar_init(&aempty_bar[s], 1); }
        mbar_init(&bfull_bar, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOl) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0 && lane == 0) {
        // the resident weight tile: w_hi at sB, w_lo behind it
        mbar_expect_tx(&bfull_bar, (uint32_t)(2 * B_BYTES));
        tma_load_3d(sB, &tmB, &bfull_bar, 0, 0, 0);
        tma_load_3d(sB + B_BYTES, &tmB, &bfull_bar, p.w_lo_off, 0, 0);
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        mbar_wait(&bfull_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t bhi = umma_desc<BK>(smem_u32(sB)), blo = umma_desc<BK>(smem_u32(sB + B_BYTES));
        int ti = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
            const int as = ti & 1;
            mbar_wait(&tempty_bar[as], ((ti >> 1) & 1) ^ 1);
            const int s = ti % STEM_STAGES;
            mbar_wait(&afull_bar[s], (ti / STEM_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_acc = tmem_base + (uint32_t)(as * p.BN);
            const uint32_t sa = smem_u32(smem + (size_t)s * 2 * A_BYTES);
            const uint64_t ahi = umma_desc<BK>(sa), alo = umma_desc<BK>(sa + A_BYTES);
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
                umma_bf16(tmem_acc, ahi + 2 * kk, bhi + 2 * kk, idesc, kk > 0 ? 1u : 0u);
                umma_bf16(tmem_acc, alo + 2 * kk, bhi + 2 * kk, idesc, 1u);
                umma_bf16(tmem_acc, ahi + 2 * kk, blo + 2 * kk, idesc, 1u);
            }
            umma_commit(&aempty_bar[s]);
            umma_commit(&tfull_bar[as]);
        }
    } else if (warp >= 4 && warp < 8) {
        // ---------------- epilogue: TMEM -> scale / bias / ReLU -> hi | lo -> staged tile -> TMA store ----------------
        const int wq = warp & 3, r = wq * 32 + lane;
        const bool leader = threadIdx.x == 128;
        const int rsw = r & 7;
        int ti = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
            const int as = ti & 1;
            mbar_wait(&tfull_bar[as], (ti >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (ti >= 2) {                   // staging buffer `as`: the store issued two tiles ago has been read out of it
                if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            uint8_t* sh = sOut + (size_t)as * 2 * STEM_OUT_TILE + (size_t)r * 128;
            uint8_t* sl = sh + STEM_OUT_TILE;
            const uint32_t trow = tmem_base + (uint32_t)(as * p.BN) + ((uint32_t)(wq * 32) << 16);
            uint32_t vn[16];
            tmem_ld16_nowait(trow, vn);
#pragma unroll
            for (int c = 0; c < 64; c += 16) {
                uint32_t v[16];
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = vn[j];
                if (c + 16 < 64) tmem_ld16_nowait(trow + c + 16, vn);
                uint32_t wh[8], wl[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float f0 = fmaf(__uint_as_float(v[2 * j]), s_scale[c + 2 * j], s_bias[c + 2 * j]);
                    float f1 = fmaf(__uint_as_float(v[2 * j + 1]), s_scale[c + 2 * j + 1], s_bias[c + 2 * j + 1]);
                    if (p.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);       // hi = bf16(v)
                    const uint32_t hw = *reinterpret_cast<const uint32_t*>(&h2);
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(f0 - __uint_as_float(hw << 16), f1 - __uint_as_float(hw & 0xffff0000u));
                    wh[j] = hw;
                    wl[j] = *reinterpret_cast<const uint32_t*>(&l2);                // lo = bf16(v - hi)
                }
                const int q = c >> 3;                                               // 16-byte chunks q, q + 1 of the 128-byte row
                *reinterpret_cast<uint4*>(sh + ((q ^ rsw) << 4)) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                *reinterpret_cast<uint4*>(sh + (((q + 1) ^ rsw) << 4)) = make_uint4(wh[4], wh[5], wh[6], wh[7]);
                *reinterpret_cast<uint4*>(sl + ((q ^ rsw) << 4)) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
                *reinterpret_cast<uint4*>(sl + (((q + 1) ^ rsw) << 4)) = make_uint4(wl[4], wl[5], wl[6], wl[7]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // staged rows -> visible to the TMA engine
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (leader) {
                const int m0 = tile * BM;                                           // rows beyond M are clipped by the tensor map
                const uint32_t s0 = smem_u32(sOut + (size_t)as * 2 * STEM_OUT_TILE);
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmOh), "r"(s0), "r"(0),
                             "r"(m0)
                             : "memory");
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmOl),
                             "r"(s0 + (uint32_t)STEM_OUT_TILE), "r"(0), "r"(m0)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all stores have landed before the CTA exits
    } else if (warp >= 8) {
        // ---------------- gather producers: thread r builds row r of the A tiles ----------------
        const int r = threadIdx.x - 256;
        const uint32_t plane = (uint32_t)(sp.H * sp.W), W = (uint32_t)sp.W;
        const int sw = (r >> 1) & 3;
        int ti = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
            const int s = ti % STEM_STAGES;
            const uint32_t m = (uint32_t)tile * BM + (uint32_t)r;      // M < 2^31 (checked by the host side)
            uint32_t v[32];
#pragma unroll
            for (int i = 27; i < 32; ++i) v[i] = 0;
            const bool live = (int64_t)m < p.M;
            const uint32_t n_img = m / plane, rem = m - n_img * plane;
            const int y = (int)(rem / W), x = (int)(rem - (uint32_t)y * W);
            const uint4* row = sp.norm + (size_t)n_img * plane + (size_t)y * W + x;
            // the nine loads are issued before the stage is waited for: their latency overlaps the wait
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int dy = t / 3 - 1, dx = t % 3 - 1;
                const bool ok = live && y + dy >= 0 && y + dy < sp.H && x + dx >= 0 && x + dx < sp.W;
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (ok) q = __ldg(row + dy * (int)W + dx);
                v[t * 3] = q.x; v[t * 3 + 1] = q.y; v[t * 3 + 2] = q.z;
            }
            mbar_wait(&aempty_bar[s], ((ti / STEM_STAGES) & 1) ^ 1);
            uint8_t* ahi = smem + (size_t)s * 2 * A_BYTES + (size_t)r * 64;
            uint8_t* alo = ahi + A_BYTES;
#pragma unroll
            for (int c = 0; c < 4; ++c) {                              // 16-byte chunk c = channels 8c .. 8c+7
                uint32_t wh[4], wl[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t a = v[8 * c + 2 * j], b = v[8 * c + 2 * j + 1];
                    wh[j] = __byte_perm(a, b, 0x5410);                 // (a & 0xffff) | (b << 16)
                    wl[j] = __byte_perm(a, b, 0x7632);                 // (a >> 16) | (b & 0xffff0000)
                }
                *reinterpret_cast<uint4*>(ahi + ((c ^ sw) << 4)) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                *reinterpret_cast<uint4*>(alo + ((c ^ sw) << 4)) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
            mbar_arrive(&afull_bar[s]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// ---- host side: tensor maps --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    });
    if (!fn) fail(BBOCR_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    return fn;
}

CUtensorMap make_map(void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int bk) {
    CUtensorMap m;
    cuuint64_t gd[5];
    cuuint64_t gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gd, gs, bx, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(BBOCR_E_CUDA, "cuTensorMapEncodeTiled failed (%d) rank %d", (int)r, rank);
    return m;
}

int pick_bk(const ConvW& cw, const Act& in1, const Act& in2) {
    if (in1.C % 64 == 0 && in2.C % 64 == 0) return 64;
    if (in1.C % 32 == 0 && in2.C % 32 == 0) return 32;
    return 0;
}

}  // namespace

CUtensorMap tc_make_map(void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int bk) {
    return make_map(base, rank, dims, strides_bytes, box, bk);
}

bool conv_tc_supported(const ConvW& cw, const Act& in1, const Act& in2) {
    if (pick_bk(cw, in1, in2) == 0) return false;
    if (cw.cout_pad > 128 && cw.cout_pad % 128 != 0) return false;
    if (!cw.w_bf16) return false;
    const bool flat = cw.kh == 1 && cw.kw == 1 && cw.pad == 0;
    if (!flat && in1.H < 4) return false;
    if ((int64_t)in1.N * in1.H * in1.W > (1ll << 31) - 256) return false;
    return true;
}

// pooled: optional second output (fused MaxPool2d(2,2) when flags & CONV_POOL22, MaxPool2d((2,1)) when CONV_POOL21);
// out.p may be null together with a pooled output when only the pooled tensor is needed.
void conv_tc_forward(Handle* h, cudaStream_t st, const ConvW& cw, const Act& in1, const Act& in2, Act& out, int flags,
                     Act* pooled, const uint8_t* colmask) {
    int bk = pick_bk(cw, in1, in2);
    const bool split_in = in1.lo != nullptr;
    if (split_in) {
        ARG_CHECK(cw.w_split && (in2.C == 0 || in2.lo), "split-precision convolution needs split weights and split inputs");
    }
    TcParams p;
    p.C1 = in1.C; p.C2 = in2.C;
    p.w_lo_off = cw.cin;
    p.split_out = out.lo != nullptr ? 1 : 0;
    p.out_lo = out.lo;
    p.out2_lo = nullptr;
    p.taps_w = cw.kw; p.taps = cw.kh * cw.kw;
    p.pad = cw.pad; p.dil = cw.dil;
    p.OH = out.H; p.OW = out.W; p.NIMG = out.N;
    p.M = (int64_t)out.N * out.H * out.W;
    p.flat = (cw.kh == 1 && cw.kw == 1 && cw.pad == 0) ? 1 : 0;
    p.cout = cw.cout;
    static const int bn_max = getenv("BBOCR_TC_BN") ? atoi(getenv("BBOCR_TC_BN")) : 256;
    p.BN = cw.cout_pad <= 128 ? cw.cout_pad : ((bn_max >= 256 && cw.cout_pad % 256 == 0) ? 256 : 128);
    if (p.BN == 256) {
        // One CTA per SM works through ceil(tiles / SMs) tiles of cost ~BN each.  For the small layers (1/16 resolution,
        // recogniser strips) the wave quantisation of 256-wide tiles costs more than their better operand reuse saves:
        // take 128-wide tiles when they shorten the longest SM's queue by more than 10 % (1x1) / 30 % (3x3: 128-wide tiles
        // pull 1.5x the operand bytes per FLOP through the L2->SM fabric, measured on fc6).
        const int64_t mt = cdiv64((int64_t)out.N * out.H * out.W, BM);
        const int64_t w256 = cdiv64(mt * (cw.cout_pad / 256), h->sm_count) * 256, w128 = cdiv64(mt * (cw.cout_pad / 128), h->sm_count) * 128;
        if (w256 * 10 > w128 * (p.flat ? 11 : 13)) p.BN = 128;
    }
    p.n_tiles = cw.cout_pad / p.BN;
    p.relu = (flags & CONV_RELU) ? 1 : 0;
    p.out_f32 = (flags & CONV_OUT_F32) ? 1 : 0;
    p.out = out.p;
    p.out2 = nullptr;
    p.pool = 0;
    p.write_full = 1;
    p.scale = cw.scale;
    p.bias = cw.bias;
    p.colmask = colmask;
    // halo-patch kernel: split-precision 3x3 / pad 1 layers with at least 16 output rows (BBOCR_TC_PATCH=0: A/B switch)
    static const bool patch_on = !(getenv("BBOCR_TC_PATCH") && atoi(getenv("BBOCR_TC_PATCH")) == 0);
    const bool patch = patch_on && split_in && !p.flat && cw.kh == 3 && cw.kw == 3 && cw.pad == 1 && cw.dil == 1 && out.H >= 16 && out.W >= 8;
    if (p.flat) { p.TW = 128; p.TH = 1; p.tiles_x = p.tiles_y = 1; }
    else if (patch) { p.TW = 8; p.TH = 16; }
    else if (out.H >= 8) { p.TW = 16; p.TH = 8; }
    else { p.TW = 32; p.TH = 4; }
    if (pooled) {
        ARG_CHECK(!p.flat && (p.TW == 16 || p.TW == 8), "fused pooling needs the 16x8 or 8x16 spatial tile");
        p.pool = (flags & CONV_POOL22) ? 1 : 2;
        ARG_CHECK(out.H % 2 == 0 && (p.pool == 2 || out.W % 2 == 0), "fused pooling needs even output dimensions");
        ARG_CHECK(pooled->H == out.H / 2 && pooled->W == (p.pool == 1 ? out.W / 2 : out.W) && pooled->C == out.C, "pooled geometry");
        p.out2 = pooled->p;
        p.out2_lo = pooled->lo;
        if (pooled->lo) p.split_out = 1;
        p.write_full = out.p != nullptr;
    }
    if (p.flat) p.m_tiles = (int)cdiv64(p.M, BM);
    else {
        p.tiles_x = cdiv(out.W, p.TW);
        p.tiles_y = cdiv(out.H, p.TH);
        p.m_tiles = p.tiles_x * p.tiles_y * out.N;
    }
    p.total_tiles = p.m_tiles * p.n_tiles;
    static const int smem_budget_kb = getenv("BBOCR_TC_SMEM_KB") ? atoi(getenv("BBOCR_TC_SMEM_KB")) : 100;
    static const int ctas_per_sm = getenv("BBOCR_TC_CTAS") ? atoi(getenv("BBOCR_TC_CTAS")) : 2;
    int budget = (p.BN > 128 ? 200 : smem_budget_kb) * 1024;
    p.p_stages = 0;
    p.patch_al = 0;
    int stage_bytes;
    size_t smem;
    int ctas = p.BN > 128 ? 1 : ctas_per_sm;
    if (patch) {
        // patches: 2 stages of (hi, lo) 10 x 18 px; ring: (hi, lo) weight tiles, one per tap.  64-channel k-blocks when three
        // ring stages still fit next to the patches, else 32-channel ones; two CTAs per SM when the whole CTA stays under 100 KB
        auto plan = [&](int k, int& ring_stages, size_t& total) {
            const int pal = ((18 * 10 * k * 2) + 1023) & ~1023;
            const int rst = 2 * p.BN * k * 2;
            const int avail = 200 * 1024 - 2 * 2 * pal;
            ring_stages = std::min(8, avail / rst);
            total = (size_t)2 * 2 * pal + (size_t)ring_stages * rst + 1024;
            return pal;
        };
        int rs = 0;
        size_t tot = 0;
        int pal = plan(bk, rs, tot);
        if (bk == 64 && rs < 3) { bk = 32; pal = plan(bk, rs, tot); }
        ARG_CHECK(rs >= 2, "conv_tc patch: tile does not fit");
        if (p.BN <= 64 && tot > 100 * 1024) {
            // narrow layers (short MMAs, A-read bound): two CTAs per SM hide the tile switch and the barrier latencies better
            // than one deep ring; 32-channel k-blocks halve the patches so that both fit
            const int pal32 = ((18 * 10 * 32 * 2) + 1023) & ~1023, rst32 = 2 * p.BN * 32 * 2;
            const int rs2 = std::min(8, (100 * 1024 - 1024 - 4 * pal32) / rst32);
            if (rs2 >= 4) { bk = 32; pal = pal32; rs = rs2; tot = (size_t)4 * pal + (size_t)rs * rst32 + 1024; }
        }
        p.p_stages = 2;
        p.patch_al = pal;
        p.stages = rs;
        stage_bytes = 2 * p.BN * bk * 2;
        smem = tot;
        ctas = tot <= 100 * 1024 ? 2 : 1;
    } else {
        // a paired (hi + lo) stage is twice as large: fall back to 32-channel k-blocks when fewer than three 64-channel stages fit
        if (split_in && bk == 64 && budget / (2 * (BM * 64 * 2 + p.BN * 64 * 2)) < 3) bk = 32;
        stage_bytes = (split_in ? 2 : 1) * (BM * bk * 2 + p.BN * bk * 2);
        p.stages = std::min(8, std::max(2, budget / stage_bytes));
        smem = (size_t)p.stages * stage_bytes + 1024;
    }
    (void)stage_bytes;
    // two-MMA scheme for split-precision layers with BN <= 128 and enough k-iterations for the MMA stream to matter
    // (BBOCR_TC_NCAT=0: the three-MMA scheme everywhere, A/B switch)
    static const bool ncat_on = !(getenv("BBOCR_TC_NCAT") && atoi(getenv("BBOCR_TC_NCAT")) == 0);
    // cout_pad <= 128 (not BN <= 128): BN of the wider layers depends on the batch geometry, and the two schemes add the same
    // products in a different order -- a layer must use ONE scheme whatever else shares its launch (batched == single, bit for bit)
    p.ncat = (ncat_on && split_in && cw.cout_pad <= 128 && p.BN % 8 == 0 && p.taps * ((in1.C + in2.C) / bk) >= 4) ? 1 : 0;

    auto act_map = [&](const Act& a) {
        if (p.flat) {
            uint64_t dims[2] = {(uint64_t)a.C, (uint64_t)((int64_t)a.N * a.H * a.W)};
            uint64_t str[1] = {(uint64_t)a.C * 2};
            uint32_t box[2] = {(uint32_t)bk, (uint32_t)BM};
            return make_map(a.p, 2, dims, str, box, bk);
        }
        uint64_t dims[4] = {(uint64_t)a.C, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
        uint64_t str[3] = {(uint64_t)a.C * 2, (uint64_t)a.W * a.C * 2, (uint64_t)a.H * a.W * a.C * 2};
        uint32_t box[4] = {(uint32_t)bk, (uint32_t)(patch ? 10 : p.TW), (uint32_t)(patch ? 18 : p.TH), 1};
        return make_map(a.p, 4, dims, str, box, bk);
    };
    CUtensorMap mA1 = act_map(in1);
    CUtensorMap mA2 = in2.C > 0 ? act_map(in2) : mA1;
    CUtensorMap mA3 = mA1, mA4 = mA1;
    if (split_in) {
        Act lo = in1;
        lo.p = in1.lo;
        mA2 = act_map(lo);
        if (in2.C > 0) {
            mA3 = act_map(in2);
            Act lo2 = in2;
            lo2.p = in2.lo;
            mA4 = act_map(lo2);
        }
    }
    const uint64_t wcin = split_in ? (uint64_t)cw.cin * 2 : (uint64_t)cw.cin;
    uint64_t wd[3] = {wcin, (uint64_t)cw.cout_pad, (uint64_t)p.taps};
    uint64_t ws[2] = {wcin * 2, (uint64_t)cw.cout_pad * wcin * 2};
    uint32_t wb[3] = {(uint32_t)bk, (uint32_t)p.BN, 1};
    CUtensorMap mB = make_map(split_in ? cw.w_split : cw.w_bf16, 3, wd, ws, wb, bk);

    // persistent grid: a multiple of the SM count (148 on B200), never more CTAs than tiles
    const unsigned grid = (unsigned)std::min<int64_t>(p.total_tiles, (int64_t)h->sm_count * ctas);
    if (!h->tc_attr_set) {          // per device (one handle = one device)
        // dynamic limit = the device's opt-in maximum per block minus the kernel's static shared memory (the patch kernels hold
        // 10-18 KB of pooling / scale-bias staging statically)
        int dev = 0, optin = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        CUDA_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        auto set_smem = [&](const void* fn) {
            cudaFuncAttributes fa;
            CUDA_CHECK(cudaFuncGetAttributes(&fa, fn));
            CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
        };
#define TC_SMEM(...) set_smem(reinterpret_cast<const void*>(&__VA_ARGS__))
        TC_SMEM(k_conv_tc<64, 0, false>); TC_SMEM(k_conv_tc<32, 0, false>);
        TC_SMEM(k_conv_tc<64, 1, false>); TC_SMEM(k_conv_tc<32, 1, false>);
        TC_SMEM(k_conv_tc<64, 1, true>); TC_SMEM(k_conv_tc<32, 1, true>);
        TC_SMEM(k_conv_tc_patch<64, false>); TC_SMEM(k_conv_tc_patch<32, false>);
        TC_SMEM(k_conv_tc_patch<64, true>); TC_SMEM(k_conv_tc_patch<32, true>);
        TC_SMEM(k_conv_tc_patch<64, false, 8>); TC_SMEM(k_conv_tc_patch<32, false, 8>);
        TC_SMEM(k_conv_tc_patch<64, true, 8>); TC_SMEM(k_conv_tc_patch<32, true, 8>);
#undef TC_SMEM
        h->tc_attr_set = true;
    }
    // one CTA per SM and BN = 128 (conv2_x, up2b): the four epilogue warps of the single resident CTA -- one per scheduler, ~0.2
    // instructions per cycle each -- are the layer's critical path; eight warps split the columns.  (Layers with BN = 256 are
    // tensor-bound, layers with BN <= 64 run two CTAs per SM.)  BBOCR_TC_EW8=0: A/B switch
    static const bool ew8_on = !(getenv("BBOCR_TC_EW8") && atoi(getenv("BBOCR_TC_EW8")) == 0);
    if (patch && ew8_on && ctas == 1 && p.BN == 128) {
        if (p.ncat) {
            if (bk == 64) k_conv_tc_patch<64, true, 8><<<grid, 384, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
            else k_conv_tc_patch<32, true, 8><<<grid, 384, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
        } else {
            if (bk == 64) k_conv_tc_patch<64, false, 8><<<grid, 384, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
            else k_conv_tc_patch<32, false, 8><<<grid, 384, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
        }
    } else if (patch && p.ncat) {
        if (bk == 64) k_conv_tc_patch<64, true><<<grid, 256, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
        else k_conv_tc_patch<32, true><<<grid, 256, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
    } else if (patch) {
        if (bk == 64) k_conv_tc_patch<64, false><<<grid, 256, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
        else k_conv_tc_patch<32, false><<<grid, 256, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
    } else if (split_in && p.ncat) {
        if (bk == 64) k_conv_tc<64, 1, true><<<grid, 256, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
        else k_conv_tc<32, 1, true><<<grid, 256, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
    } else if (split_in) {
        if (bk == 64) k_conv_tc<64, 1, false><<<grid, 256, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
        else k_conv_tc<32, 1, false><<<grid, 256, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
    } else {
        if (bk == 64) k_conv_tc<64, 0, false><<<grid, 256, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
        else k_conv_tc<32, 0, false><<<grid, 256, smem, st>>>(mA1, mA2, mA3, mA4, mB, p);
    }
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

// conv1_1 of the bf16x3 detector straight from the u8 images (k_conv_stem): out = split NHWC [nimg][H][W][64]
bool conv_stem_supported(const ConvW& cw, const Act& out) {
    return cw.w_split && cw.cin == 32 && cw.cout == 64 && cw.cout_pad == 64 && cw.kh == 1 && cw.kw == 1 && out.lo != nullptr && out.C == 64;
}

// normalised canvas of one image for k_conv_stem: dst = [H][W] uint4
void stem_norm_forward(Handle* h, cudaStream_t st, const uint8_t* img, int th, int tw, void* dst, int H, int W, const float* mean,
                       const float* sd) {
    k_stem_norm<<<dim3(cdiv(W, 256), H), 256, 0, st>>>(img, th, tw, reinterpret_cast<uint4*>(dst), H, W, mean[0], mean[1], mean[2], sd[0],
                                                        sd[1], sd[2]);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

void conv_stem_forward(Handle* h, cudaStream_t st, const ConvW& cw, const void* norm, Act& out, int flags) {
    ARG_CHECK(conv_stem_supported(cw, out), "conv_stem: unsupported layer");
    TcParams p{};
    p.C1 = 32; p.C2 = 0;
    p.w_lo_off = cw.cin;
    p.split_out = 1;
    p.out = out.p; p.out_lo = out.lo; p.out2 = nullptr; p.out2_lo = nullptr;
    p.taps_w = 1; p.taps = 1; p.pad = 0; p.dil = 1;
    p.OH = out.H; p.OW = out.W; p.NIMG = out.N;
    p.M = (int64_t)out.N * out.H * out.W;
    ARG_CHECK(p.M < (int64_t)1 << 31, "conv_stem: batch of %lld pixels", (long long)p.M);
    p.flat = 1;
    p.cout = cw.cout; p.BN = cw.cout_pad; p.n_tiles = 1;
    p.relu = (flags & CONV_RELU) ? 1 : 0;
    p.out_f32 = 0; p.pool = 0; p.write_full = 1;
    p.scale = cw.scale; p.bias = cw.bias; p.colmask = nullptr;
    p.TW = 128; p.TH = 1; p.tiles_x = p.tiles_y = 1;
    p.m_tiles = (int)cdiv64(p.M, BM);
    p.total_tiles = p.m_tiles;
    p.stages = STEM_STAGES; p.p_stages = 0; p.patch_al = 0; p.ncat = 0;
    StemParams sp{};
    sp.norm = reinterpret_cast<const uint4*>(norm);
    sp.H = out.H; sp.W = out.W;
    const uint64_t wcin = (uint64_t)cw.cin * 2;
    uint64_t wd[3] = {wcin, (uint64_t)cw.cout_pad, 1};
    uint64_t ws[2] = {wcin * 2, (uint64_t)cw.cout_pad * wcin * 2};
    uint32_t wb[3] = {32, (uint32_t)p.BN, 1};
    CUtensorMap mB = make_map(cw.w_split, 3, wd, ws, wb, 32);
    // output maps: [M pixels][64 channels] bf16, one 128-pixel x 128-byte box per tile (SWIZZLE_128B staging image)
    uint64_t od[2] = {64, (uint64_t)p.M};
    uint64_t os[1] = {128};
    uint32_t ob[2] = {64, (uint32_t)BM};
    CUtensorMap mOh = make_map(out.p, 2, od, os, ob, 64), mOl = make_map(out.lo, 2, od, os, ob, 64);
    const size_t smem = (size_t)STEM_STAGES * 2 * BM * 32 * 2 + 8192 + (size_t)4 * STEM_OUT_TILE + 1024;
    {
        std::lock_guard<std::mutex> g(h->stat_mu);
        if (!h->stem_attr_set) {
            CUDA_CHECK(cudaFuncSetAttribute(k_conv_stem, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            h->stem_attr_set = true;
        }
    }
    const unsigned grid = (unsigned)std::min<int64_t>(p.total_tiles, (int64_t)h->sm_count * 2);
    k_conv_stem<<<grid, 384, smem, st>>>(mB, mOh, mOl, p, sp);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace bbocr
