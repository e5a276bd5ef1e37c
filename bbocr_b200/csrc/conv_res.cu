// conv_res.cu -- 3x3 / pad 1 implicit-GEMM convolution for the LOW-CHANNEL layers, with every operand reused from shared
// memory (sm_100a).
//
// conv_tc.cu fetches one shifted 128-pixel tile AND one weight tile per filter tap.  For Cin, Cout <= 128 that is 24-32 KB
// of L2->SM traffic per 128..256 tensor-core cycles: the layers at full and half resolution (conv1_2, conv2_x, the last
// decoder stages) end up bound by the ~60 B/clk an SM can ingest, not by the tensor pipe
// (measured with ncu at the start of round 1: 9x the input volume crossed the L2->SM fabric).  Here
//   * the CTA's weight slice  [9 taps][Cin][BN]  is loaded ONCE per persistent CTA and stays resident (<= 147 KB);
//   * the input patch of an output tile -- 8 px wide, 16*MT px high, plus the 1-px halo; TMA out-of-bounds zero fill is the
//     convolution padding -- is loaded ONCE per 64-channel block (10 x (16 MT + 2) px x 128 B);
//   * tap (ky,kx) of M-tile mt is nothing but a shifted UMMA descriptor into that patch:
//         start = patch + ((16 mt + ky) * 10 + kx) * 128 B,   8-row group stride (SBO) = one patch row = 1280 B.
//     SWIZZLE_128B is a function of absolute shared-memory address bits for both the TMA write and the tensor-core read,
//     so a start address that is not atom-aligned needs no base offset (verified on B200: tests/test_gpu_conv_tc.py).
// L2->SM traffic drops from 9 x (16 + BN/8) KB to 23 KB per 128 output pixels; the kernel is then bound by the tensor
// core's own operand reads from shared memory (N = 64: 48 clk per K = 16 step).
//
// Warp roles as in conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer (warp-uniform loop, elect.sync for the issue),
// warp 2 TMEM allocation, warps 4-7 epilogue (folded BN scale/bias, ReLU, optional fused 2x2 / 2x1 max-pool, bf16 NHWC
// stores) on double-buffered TMEM accumulators.  Each CTA owns one BN-wide slice of the output channels for its whole
// life (that is what keeps the weights resident) and walks the output tiles with a static stride.
#include <cuda.h>

#include "engine.h"

namespace bbocr {

CUtensorMap tc_make_map(void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int bk);

namespace {

// Geometry of one variant: KB channels per k-block (64 -> 128-byte pixel rows, SWIZZLE_128B; 32 -> 64-byte rows, SWIZZLE_64B),
// TAPS = 9 (3x3, pad 1: 1-px halo) or 1 (1x1: the patch is the tile itself), MT M-tiles of 8 x 16 px stacked vertically.
template <int KB, int TAPS, int MT>
struct Geo {
    static constexpr int PIX = KB * 2;                       // bytes per pixel row of the operand tiles
    static constexpr int HALO = TAPS == 9 ? 1 : 0;
    static constexpr int PW = 8 + 2 * HALO;                  // patch width in pixels
    static constexpr int PH = 16 * MT + 2 * HALO;            // patch height
    static constexpr int PPITCH = PW * PIX;                  // bytes per patch row = stride between 8-pixel groups
    static constexpr int PATCH_BYTES = PH * PPITCH;
    static constexpr int KSTEPS = KB / 16;                   // UMMA_K = 16 steps per k-block
    static constexpr uint64_t LAYOUT = KB == 64 ? 2 : 4;     // UMMA layout code: SWIZZLE_128B / SWIZZLE_64B
};

struct ResParams {
    int tiles_x, tiles_y, OH, OW, NIMG;
    int cout, BN, n_tiles, m_tiles;
    int relu, out_f32, pool, write_full;
    void* out;
    void* out2;
    const float* scale;
    const float* bias;
    int p_stages, patch_stride;
    int b_stages;                       // STREAM variant: depth of the weight-tile ring
    long long* trace;                   // optional: per-role cycle sums of CTA 0 (diagnostics)
    // fused classifier tail (EPI == 1): relu(conv) [16 ch] -> 1x1 16->16 + ReLU -> 1x1 16->2 -> text / link planes (FP32)
    const float* w3; const float* b3; const float* w4; const float* b4;
    int cp3, cp4;
    float* text; float* link;
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
        "RWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra RDONE;\n\t"
        "bra RWAIT_LOOP;\n\t"
        "RDONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
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
// K-major swizzled descriptor with an explicit stride between 8-row groups
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr, uint32_t sbo_bytes, uint64_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (layout << 61);
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

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void store16(float* o, const float* f, int nbase, int cout) {
    if ((cout & 3) == 0 && nbase + 16 <= cout) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else {
        for (int j = 0; j < 16; ++j)
            if (nbase + j < cout) o[j] = f[j];
    }
}
__device__ __forceinline__ void store16(__nv_bfloat16* o, const float* f, int nbase, int cout) {
    if ((cout & 7) == 0 && nbase + 16 <= cout) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(o + 8) = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
        for (int j = 0; j < 16; ++j)
            if (nbase + j < cout) o[j] = __float2bfloat16_rn(f[j]);
    }
}

struct RTile {
    int img, x0, y0;
};
template <int MT>
__device__ __forceinline__ RTile rtile(const ResParams& p, int m) {
    RTile t;
    int tx = m % p.tiles_x;
    m /= p.tiles_x;
    int ty = m % p.tiles_y;
    t.img = m / p.tiles_y;
    t.x0 = tx * 8;
    t.y0 = ty * 16 * MT;
    return t;
}

// STREAM = 0: the CTA's weight slice is resident.  STREAM = 1 (Cin = 128, 128-wide channel slices): weight tiles stream
// through a ring, one per (k-block, tap), each feeding the MT stacked M-tiles of the resident patch.
template <int KB, int TAPS, int NKB, int MT, int EPI, int STREAM>
__global__ void __launch_bounds__(384, 1) k_conv_res(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmP,
                                                     const ResParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t wfull, pfull[8], pempty[8], tfull[2], tempty[2], bfull[8], bempty[8];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using G = Geo<KB, TAPS, MT>;
    constexpr int PATCH_BYTES = G::PATCH_BYTES;
    const int W_TILE = p.BN * G::PIX;                      // one (tap, k-block) weight tile
    uint8_t* w_base = smem;
    uint8_t* patch_base = smem + (((STREAM ? p.b_stages : TAPS * NKB) * W_TILE + 1023) & ~1023);
    // epilogue staging (TMA-store sources): 2 x [128 px][BN] bf16 full-resolution tiles, 2 x [32 px][BN] pooled tiles
    const int ST_BYTES = 128 * p.BN * 2, PST_BYTES = 32 * p.BN * 2;
    uint8_t* st_base = patch_base + p.p_stages * p.patch_stride;
    uint8_t* pst_base = st_base + (NKB == 1 ? 2 : 1) * ST_BYTES;      // one tile per epilogue team
    __shared__ __align__(16) float s_scale[128];
    __shared__ __align__(16) float s_bias[128];
    __shared__ __align__(16) float s_w3[EPI == 1 ? 256 : 4], s_b3[EPI == 1 ? 16 : 4], s_w4[EPI == 1 ? 32 : 4], s_b4[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t ncols = 32;
    while ((int)ncols < 2 * MT * p.BN) ncols <<= 1;
    // this CTA's slice of the output channels and its share of the M tiles
    const int nt = blockIdx.x % p.n_tiles, n0 = nt * p.BN;
    const int m_first = blockIdx.x / p.n_tiles, m_step = gridDim.x / p.n_tiles;
    const bool tr = p.trace != nullptr && blockIdx.x == 0;

    if (threadIdx.x >= 128 && threadIdx.x < 128 + p.BN) {       // the CTA's channel slice never changes: keep scale/bias in smem (BN <= 128)
        s_scale[threadIdx.x - 128] = p.scale[n0 + threadIdx.x - 128];
        s_bias[threadIdx.x - 128] = p.bias[n0 + threadIdx.x - 128];
    }
    if (EPI == 1) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_w3[i] = p.w3[(i / 16) * p.cp3 + (i % 16)];     // [cin][cout]
        if (threadIdx.x < 32) s_w4[threadIdx.x] = p.w4[(threadIdx.x / 2) * p.cp4 + (threadIdx.x % 2)];
        if (threadIdx.x < 16) s_b3[threadIdx.x] = p.b3[threadIdx.x];
        if (threadIdx.x < 2) s_b4[threadIdx.x] = p.b4[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        mbar_init(&wfull, 1);
        for (int s = 0; s < p.p_stages; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); }
        if (STREAM) for (int s = 0; s < p.b_stages; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], (NKB == 1 && MT == 2) ? 8 : 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
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

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer: the resident weight slice once, then one patch per (tile, k-block) ----------
            if (!STREAM) {
                mbar_expect_tx(&wfull, (uint32_t)(TAPS * NKB * W_TILE));
                for (int tap = 0; tap < TAPS; ++tap)
                    for (int kb = 0; kb < NKB; ++kb) tma_load_3d(w_base + (tap * NKB + kb) * W_TILE, &tmB, &wfull, kb * KB, n0, tap);
            }
            int pit = 0, bit = 0;
            for (int m = m_first; m < p.m_tiles; m += m_step) {
                const RTile tc = rtile<MT>(p, m);
                for (int kb = 0; kb < NKB; ++kb, ++pit) {
                    const int ps = pit % p.p_stages;
                    mbar_wait(&pempty[ps], ((pit / p.p_stages) & 1) ^ 1);
                    mbar_expect_tx(&pfull[ps], (uint32_t)PATCH_BYTES);
                    tma_load_4d(patch_base + ps * p.patch_stride, &tmA, &pfull[ps], kb * KB, tc.x0 - G::HALO, tc.y0 - G::HALO, tc.img);
                    if (STREAM) {
                        for (int tap = 0; tap < TAPS; ++tap, ++bit) {
                            const int bs = bit % p.b_stages;
                            mbar_wait(&bempty[bs], ((bit / p.b_stages) & 1) ^ 1);
                            mbar_expect_tx(&bfull[bs], (uint32_t)W_TILE);
                            tma_load_3d(w_base + bs * W_TILE, &tmB, &bfull[bs], kb * KB, n0, tap);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (whole warp walks the loop; one elected lane issues) ----------------
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t w_addr = smem_u32(w_base), p_addr = smem_u32(patch_base);
        long long c_tempty = 0, c_pfull = 0, c_issue = 0, t0 = 0;
        int pit = 0, ti = 0, bit = 0;
        if (!STREAM) mbar_wait(&wfull, 0);
        for (int m = m_first; m < p.m_tiles; m += m_step, ++ti) {
            const int as = ti & 1;
            if (tr) t0 = clock64();
            mbar_wait(&tempty[as], ((ti >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tr) { long long t1 = clock64(); c_tempty += t1 - t0; t0 = t1; }
            const uint32_t tacc = tmem_base + (uint32_t)(as * MT * p.BN);
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb, ++pit) {
                const int ps = pit % p.p_stages;
                mbar_wait(&pfull[ps], (pit / p.p_stages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tr) { long long t1 = clock64(); c_pfull += t1 - t0; t0 = t1; }
                if (!STREAM) {
                    if (elect_one()) {
                        const uint64_t a0 = desc_kmajor(p_addr + ps * p.patch_stride, G::PPITCH, G::LAYOUT);
                        const uint64_t b0 = desc_kmajor(w_addr + kb * W_TILE, 8 * G::PIX, G::LAYOUT);
                        const uint32_t wstep = (uint32_t)(NKB * W_TILE) >> 4;            // descriptor units (16 B) per tap
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                            for (int tap = 0; tap < TAPS; ++tap) {
                                const int ky = tap / 3, kx = tap % 3;
                                const uint64_t ad = a0 + (uint64_t)((((mt * 16 + ky) * G::PW + kx) * G::PIX) >> 4);
                                const uint64_t bd = b0 + (uint64_t)(tap * wstep);
#pragma unroll
                                for (int kk = 0; kk < G::KSTEPS; ++kk)
                                    umma_bf16(tacc + (uint32_t)(mt * p.BN), ad + 2 * kk, bd + 2 * kk, idesc, (kb | tap | kk) ? 1u : 0u);
                            }
                        umma_commit(&pempty[ps]);
                        if (kb == NKB - 1) umma_commit(&tfull[as]);
                    }
                } else {
                    const uint64_t a0 = desc_kmajor(p_addr + ps * p.patch_stride, G::PPITCH, G::LAYOUT);
#pragma unroll
                    for (int tap = 0; tap < TAPS; ++tap, ++bit) {
                        const int bs = bit % p.b_stages;
                        mbar_wait(&bfull[bs], (bit / p.b_stages) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (elect_one()) {
                            const int ky = tap / 3, kx = tap % 3;
                            const uint64_t bd = desc_kmajor(w_addr + bs * W_TILE, 8 * G::PIX, G::LAYOUT);
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) {
                                const uint64_t ad = a0 + (uint64_t)((((mt * 16 + ky) * G::PW + kx) * G::PIX) >> 4);
#pragma unroll
                                for (int kk = 0; kk < G::KSTEPS; ++kk)
                                    umma_bf16(tacc + (uint32_t)(mt * p.BN), ad + 2 * kk, bd + 2 * kk, idesc, (kb | tap | kk) ? 1u : 0u);
                            }
                            umma_commit(&bempty[bs]);
                            if (tap == TAPS - 1) {
                                umma_commit(&pempty[ps]);
                                if (kb == NKB - 1) umma_commit(&tfull[as]);
                            }
                        }
                        __syncwarp();
                    }
                }
                __syncwarp();
                if (tr) { long long t1 = clock64(); c_issue += t1 - t0; t0 = t1; }
            }
        }
        if (tr && lane == 0) { p.trace[0] = c_tempty; p.trace[1] = c_pfull; p.trace[2] = c_issue; p.trace[3] = ti; }
    } else if (warp >= 4) {
        // ---------------- epilogue: M-tile mt = 8 px wide x 16 px high; TMEM lane r = hl * 8 + wl ----------------
        // TMEM -> registers -> folded BN (+ReLU) -> bf16 -> swizzled smem tile -> ONE TMA store per tile (the TMA unit does
        // the address generation and clips partial tiles); the 2x2 max-pool is taken from the staged bf16 tile (max commutes
        // with the rounding) and stored the same way.
        // Two independent teams of four warps take alternate M-tiles (own staging buffer, own named barrier), so an M-tile's
        // epilogue may take twice the tensor-core time of an M-tile before it stalls the MMA issuer.  Two-k-block layers
        // spend twice as long per M-tile on the tensor core and get by with one team (their weights need the smem).
        constexpr int N_TEAMS = NKB == 1 ? 2 : 1;
        const int team = (warp - 4) >> 2;
        const int bar_id = 1 + team;
        const int wq = warp & 3;
        const int r = wq * 32 + lane;                      // TMEM lane = row of the staged tile
        const int et = (threadIdx.x - 128) & 127;          // 0..127 within the team
        const bool leader = et == 0;
        // the tile is staged in sub-tiles of SUB = min(BN, 64) channels (one TMA store each: 128-byte rows are the widest a
        // swizzled box takes)
        const int SUB = p.BN < 64 ? p.BN : 64, n_sub = p.BN / SUB;
        const int chunks = SUB >> 3;                       // 16-byte chunks per pixel (8 for SUB = 64, 4 for SUB = 32)
        const int row_bytes = SUB * 2;
        // 16-byte chunk c of row q sits at (c ^ f(q)) : SWIZZLE_128B f = q & 7 (128-byte rows), SWIZZLE_64B f = (q >> 1) & 3
        auto swz = [&](int q, int c) { return SUB == 64 ? (c ^ (q & 7)) : (c ^ ((q >> 1) & 3)); };
        long long c_wait = 0, c_work = 0, t0 = 0;
        int ti = 0;
        uint8_t* st_team = st_base + team * ST_BYTES;
        uint8_t* pst_team = pst_base + team * PST_BYTES;
        if (team < N_TEAMS)
        for (int m = m_first; m < p.m_tiles; m += m_step, ++ti) {
            if (MT == 1 && (ti % N_TEAMS) != team) continue;
            const RTile tc = rtile<MT>(p, m);
            const int as = ti & 1;
            if (tr) t0 = clock64();
            mbar_wait(&tfull[as], (ti >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tr) { long long t1 = clock64(); c_wait += t1 - t0; t0 = t1; }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                if (MT == 2 && (mt % N_TEAMS) != team) continue;
                const uint32_t trow = tmem_base + (uint32_t)((as * MT + mt) * p.BN) + ((uint32_t)(wq * 32) << 16);
                if (EPI == 1) {
                    // classifier tail in registers: this thread's pixel never leaves the SM between conv_cls[4] and the maps
                    uint32_t v[16];
                    tmem_ld16(trow, v);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                    float x[16], y3[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) x[j] = fmaxf(fmaf(__uint_as_float(v[j]), s_scale[j], s_bias[j]), 0.f);
#pragma unroll
                    for (int o = 0; o < 16; ++o) {
                        float a = 0.f;
#pragma unroll
                        for (int c = 0; c < 16; ++c) a = fmaf(x[c], s_w3[c * 16 + o], a);
                        y3[o] = fmaxf(a + s_b3[o], 0.f);
                    }
                    float t = 0.f, l = 0.f;
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        t = fmaf(y3[c], s_w4[c * 2], t);
                        l = fmaf(y3[c], s_w4[c * 2 + 1], l);
                    }
                    const int y = tc.y0 + mt * 16 + (r >> 3), xx = tc.x0 + (r & 7);
                    if (y < p.OH && xx < p.OW) {
                        const int64_t pix = ((int64_t)tc.img * p.OH + y) * p.OW + xx;
                        p.text[pix] = t + s_b4[0];
                        p.link[pix] = l + s_b4[1];
                    }
                    continue;
                }
                for (int sub = 0; sub < n_sub; ++sub) {
                uint8_t* st = st_team + sub * (128 * 128);
                uint8_t* pst = pst_team + sub * (32 * 128);
                uint32_t w[32];                                // the thread's pixel: up to 64 bf16 channels
#pragma unroll
                for (int c = 0; c < 64; c += 32) {
                    if (c < SUB) {
                        uint32_t v[32];
                        tmem_ld32(trow + sub * 64 + c, v);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 sc = *reinterpret_cast<const float4*>(s_scale + sub * 64 + c + j);
                            const float4 bi = *reinterpret_cast<const float4*>(s_bias + sub * 64 + c + j);
                            float a0 = fmaf(__uint_as_float(v[j]), sc.x, bi.x), a1 = fmaf(__uint_as_float(v[j + 1]), sc.y, bi.y);
                            float a2 = fmaf(__uint_as_float(v[j + 2]), sc.z, bi.z), a3 = fmaf(__uint_as_float(v[j + 3]), sc.w, bi.w);
                            if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
                            __nv_bfloat162 lo2 = __floats2bfloat162_rn(a0, a1), hi2 = __floats2bfloat162_rn(a2, a3);
                            w[(c + j) >> 1] = *reinterpret_cast<uint32_t*>(&lo2);
                            w[((c + j) >> 1) + 1] = *reinterpret_cast<uint32_t*>(&hi2);
                        }
                    }
                }
                if ((mt == MT - 1 || N_TEAMS == 2) && sub == n_sub - 1) {            // this team is done with the tile's accumulators
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                }
                // the TMA store that read this team's staging tile last time has finished reading it (it had the whole
                // accumulator drain above to do so)
                if (leader) {
                    if (n_sub == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < chunks)
                        *reinterpret_cast<uint4*>(st + r * row_bytes + (swz(r, q) << 4)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (p.pool) {
                    // pooled tile: 4 px wide x 8 px high; thread -> pooled pixel pp, 16-byte chunk(s) of its channels
                    const int per = chunks >> 2;               // chunks per thread (2 for BN = 64, 1 for BN = 32)
                    const int pp = et >> 2, py = pp >> 2, px = pp & 3;
                    const int r00 = (2 * py) * 8 + 2 * px;
                    for (int k = 0; k < per; ++k) {
                        const int c = (et & 3) * per + k;
                        const uint4 q00 = *reinterpret_cast<const uint4*>(st + r00 * row_bytes + (swz(r00, c) << 4));
                        const uint4 q01 = *reinterpret_cast<const uint4*>(st + (r00 + 1) * row_bytes + (swz(r00 + 1, c) << 4));
                        const uint4 q10 = *reinterpret_cast<const uint4*>(st + (r00 + 8) * row_bytes + (swz(r00 + 8, c) << 4));
                        const uint4 q11 = *reinterpret_cast<const uint4*>(st + (r00 + 9) * row_bytes + (swz(r00 + 9, c) << 4));
                        auto mx = [](uint32_t a, uint32_t b) {
                            __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a), y = *reinterpret_cast<__nv_bfloat162*>(&b);
                            __nv_bfloat162 z = __hmax2(x, y);
                            return *reinterpret_cast<uint32_t*>(&z);
                        };
                        uint4 o;
                        o.x = mx(mx(q00.x, q01.x), mx(q10.x, q11.x));
                        o.y = mx(mx(q00.y, q01.y), mx(q10.y, q11.y));
                        o.z = mx(mx(q00.z, q01.z), mx(q10.z, q11.z));
                        o.w = mx(mx(q00.w, q01.w), mx(q10.w, q11.w));
                        *reinterpret_cast<uint4*>(pst + pp * row_bytes + (swz(pp, c) << 4)) = o;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                }
                if (leader) {
                    const int y0 = tc.y0 + mt * 16;
                    if (!p.pool || p.write_full)
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmO),
                                     "r"(smem_u32(st)), "r"(n0 + sub * 64), "r"(tc.x0), "r"(y0), "r"(tc.img)
                                     : "memory");
                    if (p.pool)
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmP),
                                     "r"(smem_u32(pst)), "r"(n0 + sub * 64), "r"(tc.x0 >> 1), "r"(y0 >> 1), "r"(tc.img)
                                     : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                            }
            }
            if (tr) { long long t1 = clock64(); c_work += t1 - t0; t0 = t1; }
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all stores have landed before the CTA exits
        if (tr && warp == 4 && lane == 0) { p.trace[4] = c_wait; p.trace[5] = c_work; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

int res_mode() {          // BBOCR_RES=0 off (A/B against conv_tc.cu), 2 = force wherever supported (tests), 3 = force, no streamed variant
    static const int m = getenv("BBOCR_RES") ? atoi(getenv("BBOCR_RES")) : 1;
    return m;
}

struct ResPlan {
    int kb, taps, nkb, BN, MT, p_stages, patch_stride;
    int stream, b_stages;
    size_t smem;
};
bool res_plan(const ConvW& cw, const Act& in1, const Act& out, ResPlan& pl, bool tail = false) {
    pl.taps = cw.kh * cw.kw;
    if (in1.C % 64 == 0) { pl.kb = 64; pl.nkb = in1.C / 64; }
    else if (in1.C == 32) { pl.kb = 32; pl.nkb = 1; }
    else return false;
    if (pl.nkb < 1 || pl.nkb > 2) return false;
    if (pl.taps == 1 && pl.kb != 32) return false;              // the 1x1 variant exists for the K = 32 stem (conv1_1)
    pl.stream = 0;
    pl.b_stages = 0;
    if (pl.nkb == 2 && !tail && pl.taps == 9 && cw.cout % 128 == 0 && out.H >= 32 && res_mode() != 3) {
        // Cin = 128, Cout a multiple of 128: 128-wide channel slices (tensor-bound at the full N = 128 rate), two stacked
        // M-tiles per patch, weight tiles streamed through a ring (each 16 KB tile feeds 8 MMAs = 512 tensor cycles)
        pl.stream = 1;
        pl.BN = 128;
        pl.MT = 2;
        pl.patch_stride = (((16 * 2 + 2) * 10 * 128) + 1023) & ~1023;
        pl.p_stages = 2;
        const size_t staging = (size_t)(128 + 32) * 128 * 2;                      // one epilogue team
        const size_t budget = 222 * 1024 - 1024 - 2 * (size_t)pl.patch_stride - staging;
        pl.b_stages = (int)std::min<size_t>(8, budget / (128 * 128));
        if (pl.b_stages < 3) return false;
        pl.smem = (size_t)pl.b_stages * 128 * 128 + 2 * (size_t)pl.patch_stride + staging + 1024;
        return true;
    }
    // other two-k-block layers would be tensor-bound at N = 64 (48 clk per step) and lose to conv_tc.cu: forced mode only
    if (pl.nkb == 2 && res_mode() < 2) return false;
    pl.BN = std::min(cw.cout_pad, 64);
    if (tail) { if (cw.cout != 16 || cw.cout_pad != 16) return false; }
    else if ((pl.BN != 32 && pl.BN != 64) || cw.cout % pl.BN != 0) return false;      // staged tile rows of 64 / 128 bytes
    const int pix = pl.kb * 2, halo = pl.taps == 9 ? 1 : 0;
    const size_t wbytes = ((size_t)pl.taps * pl.nkb * pl.BN * pix + 1023) & ~(size_t)1023;
    const size_t staging = (size_t)(pl.nkb == 1 ? 2 : 1) * (128 + 32) * pl.BN * 2;       // one tile + pooled tile per epilogue team
    const size_t budget = 222 * 1024 - 1024 - wbytes - staging;
    for (int mt = 2; mt >= 1; --mt) {
        if (mt == 2 && (out.H < 32 || pl.nkb > 1)) continue;
        if (2 * mt * pl.BN > 512) continue;
        const int stride = (((16 * mt + 2 * halo) * (8 + 2 * halo) * pix) + 1023) & ~1023;
        const int stages = (int)std::min<size_t>(4, budget / stride);
        if (stages >= 2) {
            pl.MT = mt;
            pl.p_stages = stages;
            pl.patch_stride = stride;
            pl.smem = wbytes + (size_t)stages * stride + staging + 1024;
            return true;
        }
    }
    return false;
}

template <int KB, int TAPS, int NKB, int MT, int EPI = 0, int STREAM = 0>
void res_launch(Handle* h, int grid, size_t smem, cudaStream_t st, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mO,
                const CUtensorMap& mP, const ResParams& p) {
    // the opt-in shared-memory size is a per-device function attribute: remember it per handle (one handle = one device)
    const int key = ((((KB * 10 + TAPS) * 10 + NKB) * 10 + MT) * 10 + EPI) * 10 + STREAM;
    {
        std::lock_guard<std::mutex> g(h->stat_mu);
        if (!h->res_attr_done.count(key)) {
            CUDA_CHECK(cudaFuncSetAttribute(k_conv_res<KB, TAPS, NKB, MT, EPI, STREAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
            h->res_attr_done.insert(key);
        }
    }
    k_conv_res<KB, TAPS, NKB, MT, EPI, STREAM><<<grid, 384, smem, st>>>(mA, mB, mO, mP, p);
}

}  // namespace

bool conv_res_supported(const ConvW& cw, const Act& in1, const Act& in2, const Act& out) {
    if (!res_mode()) return false;
    const bool k3 = cw.kh == 3 && cw.kw == 3 && cw.pad == 1 && cw.dil == 1, k1 = cw.kh == 1 && cw.kw == 1 && cw.pad == 0;
    if (!k3 && !k1) return false;
    if (in2.C != 0 || !cw.w_bf16) return false;
    if (out.H < 16 || out.W < 8 || out.H % 2 || out.W % 2) return false;
    // worth it only where many tiles amortise the resident-weight load
    if (res_mode() != 2 && (int64_t)out.N * out.H * out.W < (int64_t)128 * 148 * 4) return false;      // 2: tests force it
    ResPlan pl;
    return res_plan(cw, in1, out, pl);
}

void conv_res_forward(Handle* h, cudaStream_t st, const ConvW& cw, const Act& in1, const Act& in2, Act& out, int flags,
                      Act* pooled) {
    ResPlan pl;
    ARG_CHECK(res_plan(cw, in1, out, pl), "conv_res: unsupported geometry");
    ResParams p;
    p.OH = out.H; p.OW = out.W; p.NIMG = out.N;
    p.cout = cw.cout;
    p.BN = pl.BN;
    p.n_tiles = cw.cout_pad / pl.BN;
    p.tiles_x = cdiv(out.W, 8);
    p.tiles_y = cdiv(out.H, 16 * pl.MT);
    p.m_tiles = p.tiles_x * p.tiles_y * out.N;
    p.relu = (flags & CONV_RELU) ? 1 : 0;
    ARG_CHECK(!(flags & CONV_OUT_F32) && !(flags & CONV_POOL21), "conv_res: bf16 outputs and 2x2 pooling only");
    p.out_f32 = 0;
    p.out = out.p;
    p.out2 = nullptr;
    p.pool = 0;
    p.write_full = 1;
    p.scale = cw.scale;
    p.bias = cw.bias;
    p.p_stages = pl.p_stages;
    p.patch_stride = pl.patch_stride;
    p.b_stages = pl.b_stages;
    p.trace = nullptr;
    if (pooled) {
        p.pool = (flags & CONV_POOL22) ? 1 : 2;
        ARG_CHECK(out.H % 2 == 0 && (p.pool == 2 || out.W % 2 == 0), "fused pooling needs even output dimensions");
        p.out2 = pooled->p;
        p.write_full = out.p != nullptr;
    }
    uint64_t dims[4] = {(uint64_t)in1.C, (uint64_t)in1.W, (uint64_t)in1.H, (uint64_t)in1.N};
    uint64_t str[3] = {(uint64_t)in1.C * 2, (uint64_t)in1.W * in1.C * 2, (uint64_t)in1.H * in1.W * in1.C * 2};
    const int halo = pl.taps == 9 ? 1 : 0;
    uint32_t box[4] = {(uint32_t)pl.kb, (uint32_t)(8 + 2 * halo), (uint32_t)(16 * pl.MT + 2 * halo), 1};
    CUtensorMap mA = tc_make_map(in1.p, 4, dims, str, box, pl.kb);
    uint64_t wd[3] = {(uint64_t)cw.cin, (uint64_t)cw.cout_pad, (uint64_t)pl.taps};
    uint64_t ws[2] = {(uint64_t)cw.cin * 2, (uint64_t)cw.cout_pad * cw.cin * 2};
    uint32_t wb[3] = {(uint32_t)pl.kb, (uint32_t)pl.BN, 1};
    CUtensorMap mB = tc_make_map(cw.w_bf16, 3, wd, ws, wb, pl.kb);
    // TMA-store maps: full-resolution tile [BN ch x 8 px x 16 rows], pooled tile [BN ch x 4 px x 8 rows]
    auto out_map = [&](void* base, int OH, int OW, uint32_t bw, uint32_t bh) {
        uint64_t od[4] = {(uint64_t)cw.cout, (uint64_t)OW, (uint64_t)OH, (uint64_t)out.N};
        uint64_t os[3] = {(uint64_t)cw.cout * 2, (uint64_t)OW * cw.cout * 2, (uint64_t)OH * OW * cw.cout * 2};
        const int sub = std::min(pl.BN, 64);                 // staged sub-tiles of at most 64 channels (128-byte rows)
        uint32_t ob[4] = {(uint32_t)sub, bw, bh, 1};
        return tc_make_map(base, 4, od, os, ob, sub);
    };
    CUtensorMap mO = p.write_full && out.p ? out_map(out.p, out.H, out.W, 8, 16) : mA;
    CUtensorMap mP = p.pool ? out_map(p.out2, out.H / 2, out.W / 2, 4, 8) : mO;
    // persistent grid: one CTA per SM, a multiple of the number of channel slices
    int grid = std::min(h->sm_count, p.m_tiles * p.n_tiles);
    grid -= grid % p.n_tiles;
    ARG_CHECK(grid >= p.n_tiles, "conv_res: grid too small");
    static const bool want_trace = getenv("BBOCR_RES_TRACE") != nullptr;
    DevBuf dtrace;
    if (want_trace) {
        dtrace.alloc(8 * 8, st);
        CUDA_CHECK(cudaMemsetAsync(dtrace.p, 0, 64, st));
        p.trace = dtrace.as<long long>();
    }
    const int key = pl.kb * 1000 + pl.taps * 100 + pl.nkb * 10 + pl.MT + (pl.stream ? 100000 : 0);
    switch (key) {
        case 100000 + 64000 + 900 + 20 + 2: res_launch<64, 9, 2, 2, 0, 1>(h, grid, pl.smem, st, mA, mB, mO, mP, p); break;
        case 64000 + 900 + 10 + 2: res_launch<64, 9, 1, 2>(h, grid, pl.smem, st, mA, mB, mO, mP, p); break;
        case 64000 + 900 + 10 + 1: res_launch<64, 9, 1, 1>(h, grid, pl.smem, st, mA, mB, mO, mP, p); break;
        case 64000 + 900 + 20 + 1: res_launch<64, 9, 2, 1>(h, grid, pl.smem, st, mA, mB, mO, mP, p); break;
        case 32000 + 900 + 10 + 2: res_launch<32, 9, 1, 2>(h, grid, pl.smem, st, mA, mB, mO, mP, p); break;
        case 32000 + 900 + 10 + 1: res_launch<32, 9, 1, 1>(h, grid, pl.smem, st, mA, mB, mO, mP, p); break;
        case 32000 + 100 + 10 + 2: res_launch<32, 1, 1, 2>(h, grid, pl.smem, st, mA, mB, mO, mP, p); break;
        case 32000 + 100 + 10 + 1: res_launch<32, 1, 1, 1>(h, grid, pl.smem, st, mA, mB, mO, mP, p); break;
        default: fail(BBOCR_E_ARG, "conv_res: no kernel variant for kb %d taps %d nkb %d MT %d", pl.kb, pl.taps, pl.nkb, pl.MT);
    }
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
    if (want_trace) {
        long long t[8];
        CUDA_CHECK(cudaMemcpyAsync(t, p.trace, 64, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_sync(st));
        const double n = t[3] > 0 ? (double)t[3] : 1.0;
        fprintf(stderr, "[conv_res trace] %dx%d k%d cin %d cout %d BN %d MT %d stages %d b-stages %d grid %d | CTA0: %lld tiles; per tile: mma wait-tmem %.0f "
                        "wait-patch %.0f issue %.0f | epilogue wait %.0f work %.0f cycles\n",
                out.H, out.W, cw.kh, cw.cin, cw.cout, pl.BN, pl.MT, pl.p_stages, pl.b_stages, grid, t[3], t[0] / n, t[1] / n, t[2] / n, t[4] / n, t[5] / n);
    }
}

// conv_cls[4] (3x3, 32 -> 16, ReLU) with the rest of the classifier (1x1 16 -> 16 + ReLU, 1x1 16 -> 2) in its epilogue:
// writes the text and link score maps directly.  craft.py::CRAFT.conv_cls tail; SURVEY.md §8a B4.
bool conv_res_cls_tail_supported(const ConvW& c2, const ConvW& c3, const ConvW& c4, const Act& in) {
    if (!res_mode() || !c2.w_bf16) return false;
    if (c2.kh != 3 || c2.kw != 3 || c2.pad != 1 || c2.dil != 1 || in.C != 32) return false;
    if (c3.cin != 16 || c3.cout != 16 || c4.cin != 16 || c4.cout != 2 || c3.kh != 1 || c4.kh != 1) return false;
    if (in.H < 32 || in.W < 8) return false;
    ResPlan pl;
    return res_plan(c2, in, in, pl, true) && pl.MT == 2;
}

void conv_res_cls_tail(Handle* h, cudaStream_t st, const ConvW& c2, const ConvW& c3, const ConvW& c4, const Act& in, float* text,
                       float* link) {
    ResPlan pl;
    ARG_CHECK(res_plan(c2, in, in, pl, true) && pl.MT == 2, "conv_res_cls_tail: unsupported geometry");
    ResParams p;
    memset(&p, 0, sizeof p);
    p.OH = in.H; p.OW = in.W; p.NIMG = in.N;
    p.cout = 16; p.BN = 16; p.n_tiles = 1;
    p.tiles_x = cdiv(in.W, 8);
    p.tiles_y = cdiv(in.H, 32);
    p.m_tiles = p.tiles_x * p.tiles_y * in.N;
    p.relu = 1;
    p.scale = c2.scale; p.bias = c2.bias;
    p.p_stages = pl.p_stages; p.patch_stride = pl.patch_stride;
    p.w3 = c3.w_f32; p.b3 = c3.bias; p.cp3 = c3.cout_pad;
    p.w4 = c4.w_f32; p.b4 = c4.bias; p.cp4 = c4.cout_pad;
    p.text = text; p.link = link;
    uint64_t dims[4] = {(uint64_t)in.C, (uint64_t)in.W, (uint64_t)in.H, (uint64_t)in.N};
    uint64_t str[3] = {(uint64_t)in.C * 2, (uint64_t)in.W * in.C * 2, (uint64_t)in.H * in.W * in.C * 2};
    uint32_t box[4] = {32, 10, 34, 1};
    CUtensorMap mA = tc_make_map(in.p, 4, dims, str, box, 32);
    uint64_t wd[3] = {(uint64_t)c2.cin, (uint64_t)c2.cout_pad, 9};
    uint64_t ws[2] = {(uint64_t)c2.cin * 2, (uint64_t)c2.cout_pad * c2.cin * 2};
    uint32_t wb[3] = {32, 16, 1};
    CUtensorMap mB = tc_make_map(c2.w_bf16, 3, wd, ws, wb, 32);
    const int grid = std::min(h->sm_count, p.m_tiles);
    res_launch<32, 9, 1, 2, 1>(h, grid, pl.smem, st, mA, mB, mA, mA, p);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace bbocr
