// conv_halo.cu -- 3x3 / pad 1 / dilation 1 implicit-GEMM convolution with operand reuse in shared memory (sm_100a).
//
// conv_tc.cu fetches one shifted 128-pixel tile per filter tap, so every input pixel crosses the L2->SM fabric nine times
// and the wide layers end up bound by that traffic (profiles/r1_conv_tc_ncu_full.summary.txt).  Here one TMA load brings
// the (16*MT + 2) x 16 pixel x 64 channel input patch of a tile (with its halo; out-of-bounds zero fill = padding) into
// shared memory ONCE per 64-channel block, and all nine taps are issued from it: tap (ky,kx) of M-tile mt is the UMMA
// descriptor  start = patch + mt*16 rows + (ky*16 + kx) pixels,  stride between 8-pixel groups = one patch row
// (16 px * 128 B = 2048 B), base_offset = 0: the swizzle is a function of absolute smem address bits, so a start that is
// kx rows into a 1024-byte atom needs no correction (verified on B200 by tests/test_gpu_conv_tc.py).  The output tile is
// 8 px wide x 16*MT px high (MT = 2: two M=128 MMAs share every weight tile).  Weights stream per tap through their own
// ring.  Same warp roles / persistent tiles / double-buffered TMEM / fused max-pool epilogue as conv_tc.cu.
#include <cuda.h>

#include "engine.h"

namespace bbocr {

namespace {

constexpr int PW = 16;                 // patch width in pixels (8 output columns + halo, padded to a multiple of 8)
constexpr int PATCH_ROW = PW * 128;    // bytes per patch row (64 bf16 channels per pixel)

struct HaloParams {
    int C1, C2;
    int MT;                            // M-tiles (of 8 x 16 px) stacked vertically per CTA tile: 1 or 2
    int tiles_x, tiles_y, OH, OW, NIMG;
    int cout, BN, n_tiles, total_tiles;
    int relu, out_f32, pool, write_full;
    int base_off_mode;
    void* out;
    void* out2;
    const float* scale;
    const float* bias;
    int b_stages;
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
        "HWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra HDONE;\n\t"
        "bra HWAIT_LOOP;\n\t"
        "HDONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
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
// K-major SWIZZLE_128B descriptor with explicit stride between 8-row groups and base offset
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
           ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
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

struct HTile {
    int n0, img, x0, y0;
};
__device__ __forceinline__ HTile htile(const HaloParams& p, int tile) {
    HTile t;
    int mt = tile / p.n_tiles;
    t.n0 = (tile - mt * p.n_tiles) * p.BN;
    int tx = mt % p.tiles_x;
    mt /= p.tiles_x;
    int ty = mt % p.tiles_y;
    t.img = mt / p.tiles_y;
    t.x0 = tx * 8;
    t.y0 = ty * 16 * p.MT;
    return t;
}

__global__ void __launch_bounds__(256, 1) k_conv3x3_halo(const __grid_constant__ CUtensorMap tmA1,
                                                         const __grid_constant__ CUtensorMap tmA2,
                                                         const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t pfull[2], pempty[2], bfull[8], bempty[8], tfull[2], tempty[2];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int PATCH_BYTES = (16 * p.MT + 2) * PATCH_ROW;          // multiple of 1024
    const int B_BYTES = p.BN * 128;
    uint8_t* patch_base = smem;
    uint8_t* b_base = smem + 2 * PATCH_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb1 = p.C1 / 64, kb2 = p.C2 / 64, nkb = kb1 + kb2;
    uint32_t ncols = 32;
    while ((int)ncols < 2 * p.MT * p.BN) ncols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        for (int s = 0; s < p.b_stages; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
        // ---------------- TMA producer: one patch per 64-channel block, nine weight tiles per patch ----------------
        int pit = 0, bit = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const HTile tc = htile(p, tile);
            for (int kb = 0; kb < nkb; ++kb, ++pit) {
                const int ps = pit & 1;
                mbar_wait(&pempty[ps], ((pit >> 1) & 1) ^ 1);
                mbar_expect_tx(&pfull[ps], (uint32_t)PATCH_BYTES);
                const bool second = kb >= kb1;
                tma_load_4d(patch_base + ps * PATCH_BYTES, second ? &tmA2 : &tmA1, &pfull[ps], (second ? kb - kb1 : kb) * 64,
                            tc.x0 - 1, tc.y0 - 1, tc.img);
                for (int tap = 0; tap < 9; ++tap, ++bit) {
                    const int bs = bit % p.b_stages;
                    mbar_wait(&bempty[bs], ((bit / p.b_stages) & 1) ^ 1);
                    mbar_expect_tx(&bfull[bs], (uint32_t)B_BYTES);
                    tma_load_3d(b_base + bs * B_BYTES, &tmB, &bfull[bs], kb * 64, tc.n0, tap);
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int pit = 0, bit = 0, ti = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
            const int as = ti & 1;
            mbar_wait(&tempty[as], ((ti >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + (uint32_t)(as * p.MT * p.BN);
            for (int kb = 0; kb < nkb; ++kb, ++pit) {
                const int ps = pit & 1;
                mbar_wait(&pfull[ps], (pit >> 1) & 1);
                const uint32_t pa = smem_u32(patch_base + ps * PATCH_BYTES);
                for (int tap = 0; tap < 9; ++tap, ++bit) {
                    const int bs = bit % p.b_stages;
                    mbar_wait(&bfull[bs], (bit / p.b_stages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int ky = tap / 3, kx = tap - ky * 3;
                    const uint64_t bdesc = desc_sw128(smem_u32(b_base + bs * B_BYTES), 1024, 0);
                    for (int mt = 0; mt < p.MT; ++mt) {
                        const uint32_t start = pa + (uint32_t)((mt * 16 + ky) * PATCH_ROW + kx * 128);
                        const uint64_t adesc = desc_sw128(start, PATCH_ROW, p.base_off_mode ? (uint32_t)kx : 0u);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(tacc + (uint32_t)(mt * p.BN), adesc + 2 * kk, bdesc + 2 * kk, idesc,
                                      (kb > 0 || tap > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(&bempty[bs]);
                }
                umma_commit(&pempty[ps]);
            }
            umma_commit(&tfull[as]);
        }
    } else if (warp >= 4) {
        // ---------------- epilogue: M-tile mt = 8 px wide x 16 px high; TMEM lane r = hl * 8 + wl ----------------
        const int wq = warp & 3;
        const int r = wq * 32 + lane;
        const int hl = r >> 3, wl = r & 7;
        int ti = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
            const HTile tc = htile(p, tile);
            const int as = ti & 1;
            mbar_wait(&tfull[as], (ti >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int mt = 0; mt < p.MT; ++mt) {
                const int y = tc.y0 + mt * 16 + hl, x = tc.x0 + wl;
                int64_t pix = -1, pix2 = -1;
                if (y < p.OH && x < p.OW) pix = ((int64_t)tc.img * p.OH + y) * p.OW + x;
                if (p.pool) {      // window partners: lane ^ 1 (x), lane ^ 8 (y)
                    const int POH = p.OH >> 1, POW = p.pool == 1 ? p.OW >> 1 : p.OW;
                    const int py = y >> 1, px = p.pool == 1 ? x >> 1 : x;
                    const bool writer = (lane & 8) == 0 && (p.pool == 2 || (lane & 1) == 0);
                    if (writer && py < POH && px < POW) pix2 = ((int64_t)tc.img * POH + py) * POW + px;
                }
                const uint32_t trow = tmem_base + (uint32_t)((as * p.MT + mt) * p.BN) + ((uint32_t)(wq * 32) << 16);
                for (int c = 0; c < p.BN; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(trow + c, v);
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = tc.n0 + c + j;
                        float a = fmaf(__uint_as_float(v[j]), __ldg(p.scale + n), __ldg(p.bias + n));
                        f[j] = p.relu ? fmaxf(a, 0.f) : a;
                    }
                    const int nbase = tc.n0 + c;
                    if (pix >= 0 && (!p.pool || p.write_full)) {
                        if (p.out_f32) store16(reinterpret_cast<float*>(p.out) + pix * p.cout + nbase, f, nbase, p.cout);
                        else store16(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.cout + nbase, f, nbase, p.cout);
                    }
                    if (p.pool) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float m = f[j];
                            if (p.pool == 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
                            f[j] = m;
                        }
                        if (pix2 >= 0) {
                            if (p.out_f32) store16(reinterpret_cast<float*>(p.out2) + pix2 * p.cout + nbase, f, nbase, p.cout);
                            else store16(reinterpret_cast<__nv_bfloat16*>(p.out2) + pix2 * p.cout + nbase, f, nbase, p.cout);
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

}  // namespace

CUtensorMap tc_make_map(void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int bk);

int conv_halo_mode() {
    // 0 off, 2 on.  (1 = descriptor base_offset = kx: measured WRONG on B200 -- the tensor core applies the 128-byte
    // swizzle to absolute shared-memory address bits, so a start address kx rows into an atom needs base_offset 0.)
    // Default OFF: correct, but on B200 it is slower than conv_tc.cu (900 vs 1300 TFLOP/s on conv4_2): with one CTA per
    // SM the single MMA-issuing thread needs ~150 cycles per tcgen05.mma, which N = 128 instructions (64 tensor cycles)
    // cannot hide; kept as a measured experiment for the round-2 work on a 2-CTA (cta_group::2, N = 256) variant.
    static const int m = getenv("BBOCR_HALO") ? atoi(getenv("BBOCR_HALO")) : 0;
    return m;
}

bool conv_halo_supported(const ConvW& cw, const Act& in1, const Act& in2, const Act& out) {
    if (!conv_halo_mode()) return false;
    if (cw.kh != 3 || cw.kw != 3 || cw.pad != 1 || cw.dil != 1) return false;
    if (in1.C % 64 != 0 || in2.C % 64 != 0) return false;
    if (cw.cout_pad > 128 && cw.cout_pad % 128 != 0) return false;
    if (out.H < 16 || out.W < 8) return false;
    return true;
}

void conv_halo_forward(Handle* h, cudaStream_t st, const ConvW& cw, const Act& in1, const Act& in2, Act& out, int flags,
                       Act* pooled) {
    HaloParams p;
    p.C1 = in1.C; p.C2 = in2.C;
    p.OH = out.H; p.OW = out.W; p.NIMG = out.N;
    p.cout = cw.cout;
    p.BN = cw.cout_pad <= 128 ? cw.cout_pad : 128;
    p.n_tiles = cw.cout_pad / p.BN;
    static const int mt_max = getenv("BBOCR_HALO_MT") ? atoi(getenv("BBOCR_HALO_MT")) : 2;
    p.MT = (out.H >= 32 && mt_max >= 2) ? 2 : 1;
    p.tiles_x = cdiv(out.W, 8);
    p.tiles_y = cdiv(out.H, 16 * p.MT);
    p.total_tiles = p.tiles_x * p.tiles_y * out.N * p.n_tiles;
    p.relu = (flags & CONV_RELU) ? 1 : 0;
    p.out_f32 = (flags & CONV_OUT_F32) ? 1 : 0;
    p.out = out.p;
    p.out2 = nullptr;
    p.pool = 0;
    p.write_full = 1;
    p.base_off_mode = conv_halo_mode() == 1 ? 1 : 0;
    p.scale = cw.scale;
    p.bias = cw.bias;
    if (pooled) {
        p.pool = (flags & CONV_POOL22) ? 1 : 2;
        ARG_CHECK(out.H % 2 == 0 && (p.pool == 2 || out.W % 2 == 0), "fused pooling needs even output dimensions");
        p.out2 = pooled->p;
        p.write_full = out.p != nullptr;
    }
    const int patch_bytes = (16 * p.MT + 2) * PATCH_ROW, b_bytes = p.BN * 128;
    p.b_stages = std::min(8, std::max(2, (222 * 1024 - 2 * patch_bytes) / b_bytes));
    const size_t smem = (size_t)2 * patch_bytes + (size_t)p.b_stages * b_bytes + 1024;
    auto act_map = [&](const Act& a) {
        uint64_t dims[4] = {(uint64_t)a.C, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
        uint64_t str[3] = {(uint64_t)a.C * 2, (uint64_t)a.W * a.C * 2, (uint64_t)a.H * a.W * a.C * 2};
        uint32_t box[4] = {64, (uint32_t)PW, (uint32_t)(16 * p.MT + 2), 1};
        return tc_make_map(a.p, 4, dims, str, box, 64);
    };
    CUtensorMap mA1 = act_map(in1);
    CUtensorMap mA2 = in2.C > 0 ? act_map(in2) : mA1;
    uint64_t wd[3] = {(uint64_t)cw.cin, (uint64_t)cw.cout_pad, 9};
    uint64_t ws[2] = {(uint64_t)cw.cin * 2, (uint64_t)cw.cout_pad * cw.cin * 2};
    uint32_t wb[3] = {64, (uint32_t)p.BN, 1};
    CUtensorMap mB = tc_make_map(cw.w_bf16, 3, wd, ws, wb, 64);
    const unsigned grid = (unsigned)std::min<int64_t>(p.total_tiles, h->sm_count);
    if (!h->halo_attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(k_conv3x3_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        h->halo_attr_set = true;
    }
    k_conv3x3_halo<<<grid, 256, smem, st>>>(mA1, mA2, mB, p);
    count_launch(h);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace bbocr
