// lstm_mma.cu -- BiLSTM recurrence of the CRNN batched ACROSS CROPS on the tensor cores (throughput mode, sm_100a).
//   easyocr/model/modules.py::BidirectionalLSTM (nn.LSTM(256, 256, bidirectional=True); gate order i,f,g,o; h0 = c0 = 0).
//   SURVEY.md §8a B11 ("the recurrent step runs in a persistent kernel batched across crops").
//
// lstm.cu walks 8 crops per 8-CTA cluster on the FP32 cores: ~5 us per time step whatever the batch, i.e. the stage costs
// SM-time proportional to crops x steps.  Here the CROPS are the UMMA M dimension: one 16-CTA cluster owns (128 crops, one
// direction); CTA r keeps the 64 gate columns of hidden units [16r, 16r+16) resident in shared memory as the B operand
// and every step is
//     D[128 crops x 64 gate columns] (FP32, TMEM)  =  H[128 x 256] . W_r[64 x 256]^T
// evaluated in split precision (FP32-class, the recogniser's accuracy contract: tests/test_gpu_recognizer.py):
//     H = H_hi + H_lo, W = W_hi + W_lo (bf16 pairs);  D = H_hi W_hi + H_lo W_hi + H_hi W_lo   (48 tcgen05.mma, K = 16 each)
// TMEM lane = crop, so thread (crop, 8 units) reads its four gates with tcgen05.ld, adds the input projections, applies the
// cell update with the state in registers and writes h(t) (a) to the layer output and (b) as hi/lo bf16 into a small
// global staging matrix [128 crops][256 units] that stays in L2.  One cluster barrier later every CTA re-loads the whole
// H operand with TMA (8 boxes of 64 units x 128 crops, SWIZZLE_128B = the UMMA A layout).  Going through L2 + TMA instead of
// distributed shared memory is deliberate: every CTA needs all of H each step (128 KiB) and DSMEM moves ~20 B/clk/SM,
// TMA from L2 several times that.
//
//   gates_in : [rows][2048] FP32, row = seq.row0 + t, columns [fwd i,f,g,o | bwd i,f,g,o]   (x_t W_ih^T + b_ih + b_hh)
//   w_hh     : [2][256 k][1024] FP32 (k-major)
//   out      : [rows][512] = [h_fwd(t) | h_bwd(t)]  bf16, split bf16 (out + out_lo) or FP32
#include <cuda.h>

#include "engine.h"

namespace bbocr {

CUtensorMap tc_make_map(void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int bk);

namespace {

constexpr int MB = 128;                // crops per cluster = UMMA M
constexpr int CLM = 16;                // CTAs per cluster (non-portable size; one cluster per GPC on B200)
constexpr int UN = 16;                 // hidden units per CTA
constexpr int NG = 4 * UN;             // gate columns per CTA = UMMA N
constexpr int H_TILE = MB * 128;       // one 64-unit k-block of the A operand (128 rows x 128 B)
constexpr int W_TILE = NG * 128;       // one 64-unit k-block of the B operand
constexpr int H_BYTES = 8 * H_TILE;    // hi kb0..3 | lo kb0..3   = 128 KiB
constexpr int W_BYTES = 8 * W_TILE;    //                          =  64 KiB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LM_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LM_DONE;\n\t"
        "bra LM_WAIT;\n\t"
        "LM_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// K-major SWIZZLE_128B matrix descriptor (8-row x 128-byte atoms, 1024 bytes between 8-row groups)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

// fast, FP32-class gate non-linearities (ex2.approx + rcp.approx: ~1e-6 absolute; the split-precision products are 2^-16)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return __fdividef(2.f, 1.f + __expf(-2.f * x)) - 1.f; }

__device__ __forceinline__ uint4 pack8(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&t);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// out_mode: 0 = bf16, 1 = split bf16 (out + out_lo), 2 = FP32
__global__ void __launch_bounds__(256, 1)
    k_lstm_mma(const __grid_constant__ CUtensorMap tmS, const float* __restrict__ gates_in, const float* __restrict__ w_hh,
               void* __restrict__ out, void* __restrict__ out_lo, int out_mode, __nv_bfloat16* __restrict__ stage,
               const SeqDesc* __restrict__ seqs, const int* __restrict__ groups /*[n_groups][MB], -1 = empty*/, int fence_gpu,
               long long* __restrict__ trace /* optional: clock64 stamps of CTA 0, 8 per step */) {
    extern __shared__ uint8_t lm_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lm_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* Hs = sm;                        // A operand: [hi|lo][4 k-blocks][128 crops x 128 B]
    uint8_t* Ws = sm + H_BYTES;              // B operand: [hi|lo][4 k-blocks][64 gate columns x 128 B]
    __shared__ uint64_t full_bar[8], mma_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int s_tmax[8];
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    const int group = blockIdx.x / CLM, dir = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // resident weight slice, split hi/lo, in the UMMA K-major SWIZZLE_128B image:
    //   row n = g*16 + u  <-  W_hh[dir][k][g*256 + 16r + u];  byte (n, k) = kb*W_TILE + (n/8)*1024 + (n%8)*128 + ((k%64/8) ^ (n%8))*16 + (k%8)*2
    const float* wd = w_hh + (size_t)dir * 256 * 1024;
    for (int i = tid; i < 256 * NG; i += 256) {
        const int k = i / NG, n = i - k * NG;
        const int g = n >> 4, u = n & 15;
        const float w = __ldg(wd + (size_t)k * 1024 + g * 256 + UN * r + u);
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        const int kb = k >> 6, kk = k & 63;
        const int off = kb * W_TILE + (n >> 3) * 1024 + (n & 7) * 128 + ((((kk >> 3) ^ (n & 7)) & 7) << 4) + (kk & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(Ws + off) = hi;
        *reinterpret_cast<__nv_bfloat16*>(Ws + 4 * W_TILE + off) = lo;
    }
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full_bar[i])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmS) : "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // thread -> (crop = TMEM lane, 8 of the CTA's 16 units)
    const int crop = (warp & 3) * 32 + lane, half = warp >> 2;
    int row0 = 0, T = 0;
    {
        const int s = groups[group * MB + crop];
        if (s >= 0) { row0 = seqs[s].row0; T = seqs[s].T; }
    }
    int Tmax = T;
    for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
    if (lane == 0) s_tmax[warp] = Tmax;
    asm volatile("fence.proxy.async;" ::: "memory");            // generic-proxy smem writes (W) -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    Tmax = 0;
    for (int i = 0; i < 8; ++i) Tmax = max(Tmax, s_tmax[i]);
    const uint32_t tmem_acc = tmem_slot;
    // instruction descriptor: D = F32 (1<<4), A = B = BF16 (1<<7, 1<<10), K-major, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NG >> 3) << 17) | ((uint32_t)(MB >> 4) << 24);
    const uint32_t hs_addr = smem_u32(Hs), ws_addr = smem_u32(Ws);
    // staging rows of this (group, direction): [parity][hi|lo][128 crops]
    const int stage_row0 = (group * 2 + dir) * 4 * MB;
    float cstate[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cstate[j] = 0.f;

    const bool tr = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
#define LM_STAMP(i) do { if (tr && s < 64) trace[s * 8 + (i)] = clock64(); } while (0)
    for (int s = 0; s < Tmax; ++s) {
        LM_STAMP(0);
        const bool active = s < T;
        const int t = dir ? T - 1 - s : s;
        // this step's input projections (in flight while the cluster synchronises and the tensor core works)
        float pre[4][8];
        if (active) {
            const float* gp = gates_in + (size_t)(row0 + t) * 2048 + dir * 1024 + UN * r + 8 * half;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(gp + g * 256));
                const float4 b = __ldg(reinterpret_cast<const float4*>(gp + g * 256 + 4));
                pre[g][0] = a.x; pre[g][1] = a.y; pre[g][2] = a.z; pre[g][3] = a.w;
                pre[g][4] = b.x; pre[g][5] = b.y; pre[g][6] = b.z; pre[g][7] = b.w;
            }
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int j = 0; j < 8; ++j) pre[g][j] = 0.f;
        }
        if (s > 0) {
            // every CTA of the cluster has stored its slice of h(s-1): release/acquire at cluster scope
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");      // arrived at the end of step s-1
            LM_STAMP(1);
            const uint32_t ph = (uint32_t)((s - 1) & 1);
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                asm volatile("fence.proxy.async;" ::: "memory");
                const int rb = stage_row0 + (s & 1) * 2 * MB;
                // one mbarrier per 16 KiB box so that the MMAs of a k-block start as soon as its operand has landed
#pragma unroll
                for (int hl = 0; hl < 2; ++hl)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint32_t bar = smem_u32(&full_bar[hl * 4 + kb]);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)H_TILE) : "memory");
                        asm volatile(
                            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                                hs_addr + (uint32_t)((hl * 4 + kb) * H_TILE)),
                            "l"(&tmS), "r"(bar), "r"(kb * 64), "r"(rb + hl * MB)
                            : "memory");
                    }
                // D = H_hi W_hi + H_hi W_lo (as the hi boxes arrive) + H_lo W_hi
#pragma unroll
                for (int hl = 0; hl < 2; ++hl)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        mbar_wait(&full_bar[hl * 4 + kb], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t ad = desc_sw128(hs_addr + (hl * 4 + kb) * H_TILE);
                        const uint64_t bhi = desc_sw128(ws_addr + kb * W_TILE), blo = desc_sw128(ws_addr + (4 + kb) * W_TILE);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) mma_bf16(tmem_acc, ad + 2 * kk, bhi + 2 * kk, idesc, (hl | kb | kk) ? 1u : 0u);
                        if (hl == 0) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) mma_bf16(tmem_acc, ad + 2 * kk, blo + 2 * kk, idesc, 1u);
                        }
                    }
                LM_STAMP(2);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar))
                             : "memory");
                LM_STAMP(3);
            }
            mbar_wait(&mma_bar, ph);
            LM_STAMP(4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[4][8];
            const uint32_t trow = tmem_acc + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(8 * half);
#pragma unroll
            for (int g = 0; g < 4; ++g) tmem_ld8(trow + g * UN, v[g]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int j = 0; j < 8; ++j) pre[g][j] += __uint_as_float(v[g][j]);
            LM_STAMP(5);
        }
        // gates, cell and hidden state of (crop, 8 units)
        float hv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float ig = sigmoid_fast(pre[0][j]);
            const float fg = sigmoid_fast(pre[1][j]);
            const float gg = tanh_fast(pre[2][j]);
            const float og = sigmoid_fast(pre[3][j]);
            const float c = fg * cstate[j] + ig * gg;
            cstate[j] = active ? c : cstate[j];
            hv[j] = active ? og * tanh_fast(c) : 0.f;
        }
        float hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            hi[j] = __bfloat162float(__float2bfloat16_rn(hv[j]));
            lo[j] = hv[j] - hi[j];
        }
        LM_STAMP(6);
        if (s + 1 < Tmax) {
            // h(s) for the next step's operand: staging[parity (s+1)&1][hi|lo][crop][16r + 8*half ..]
            __nv_bfloat16* sp = stage + ((size_t)(stage_row0 + ((s + 1) & 1) * 2 * MB + crop) * 256 + UN * r + 8 * half);
            *reinterpret_cast<uint4*>(sp) = pack8(hi);
            *reinterpret_cast<uint4*>(sp + (size_t)MB * 256) = pack8(lo);
            // generic-proxy global writes -> visible to the peers' TMA reads (async proxy); the release/acquire pair of the
            // cluster barrier at the top of the next step orders them across the CTAs
            if (fence_gpu) __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
            // arrive now, wait at the top of the next step: the layer-output stores below and the next step's input
            // projections are in flight while the cluster synchronises
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        }
        if (active) {
            const size_t o = (size_t)(row0 + t) * 512 + dir * 256 + UN * r + 8 * half;
            if (out_mode == 2) {
                float* op = reinterpret_cast<float*>(out) + o;
                *reinterpret_cast<float4*>(op) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                *reinterpret_cast<float4*>(op + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
            } else {
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + o) = pack8(hi);
                if (out_mode == 1) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out_lo) + o) = pack8(lo);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        LM_STAMP(7);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_acc) : "memory");
}

}  // namespace

int lstm_mma_group_size() { return MB; }

void lstm_sequences_mma(Handle* h, Lane& lane, const float* gates_in, const float* w_hh, int n_seq, const SeqDesc* seqs_dev,
                        const int* groups_dev, int n_groups, void* out, void* out_lo, int out_mode) {
    cudaStream_t st = lane.stream;
    if (n_seq == 0 || n_groups == 0) return;
#ifdef BBOCR_DIAG              // diagnostics build only (make DIAG=1): results are garbage; never in the shipped library
    static const bool diag_skip = getenv("BBOCR_DIAG_SKIP_LSTM") != nullptr;
    if (diag_skip) return;
#endif
    const size_t smem = (size_t)H_BYTES + W_BYTES + 1024;
    // staging: [group][dir][parity][hi|lo][128 crops][256 units] bf16 (256 KiB per (group, direction); lives in L2)
    const uint64_t stage_rows = (uint64_t)n_groups * 2 * 4 * MB;
    DevBuf stage(stage_rows * 256 * 2, st);
    uint64_t dims[2] = {256, stage_rows};
    uint64_t str[1] = {512};
    uint32_t box[2] = {64, (uint32_t)MB};
    CUtensorMap tmS = tc_make_map(stage.p, 2, dims, str, box, 64);
    if (!h->lstm_mma_attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(k_lstm_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_CHECK(cudaFuncSetAttribute(k_lstm_mma, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        h->lstm_mma_attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CLM * n_groups, 2);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLM;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static const bool want_trace = getenv("BBOCR_LSTM_TRACE") != nullptr;
    static const int fence_gpu = getenv("BBOCR_LSTM_FENCE") ? atoi(getenv("BBOCR_LSTM_FENCE")) : 0;
    DevBuf dtrace;
    long long* trace = nullptr;
    if (want_trace) {
        dtrace.alloc(64 * 8 * 8, st);
        CUDA_CHECK(cudaMemsetAsync(dtrace.p, 0, 64 * 8 * 8, st));
        trace = dtrace.as<long long>();
    }
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_lstm_mma, tmS, gates_in, w_hh, out, out_lo, out_mode, stage.as<__nv_bfloat16>(), seqs_dev,
                                  groups_dev, fence_gpu, trace));
    count_launch(h);
    if (want_trace) {           // diagnostic: per-phase cycles of CTA 0, averaged over steps 8..39
        std::vector<long long> t(64 * 8);
        CUDA_CHECK(cudaMemcpyAsync(t.data(), trace, t.size() * 8, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_sync(st));
        double acc[8] = {0};
        int cnt = 0;
        for (int s = 8; s < 40; ++s) {
            if (!t[s * 8 + 7] || !t[(s + 1) * 8]) break;
            for (int i = 0; i < 7; ++i) acc[i] += (double)(t[s * 8 + i + 1] - t[s * 8 + i]);
            acc[7] += (double)(t[(s + 1) * 8] - t[s * 8]);
            ++cnt;
        }
        if (cnt)
            fprintf(stderr, "[lstm_mma trace] groups=%d steps=%d cycles: gin+barrier %.0f | tma %.0f | mma-issue %.0f | mma-wait %.0f | "
                            "tmem-ld %.0f | gates+out %.0f | stage+fence %.0f | step total %.0f\n",
                    n_groups, cnt, acc[0] / cnt, acc[1] / cnt, acc[2] / cnt, acc[3] / cnt, acc[4] / cnt, acc[5] / cnt, acc[6] / cnt, acc[7] / cnt);
    }
}

}  // namespace bbocr
