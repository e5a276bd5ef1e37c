// lstm_tc.cu -- the BiLSTM recurrence with the per-step mat-vec on the tensor cores (throughput mode).
//
// Same decomposition as lstm.cu (8-CTA cluster per (group of crops, direction); CTA r owns hidden units [32r, 32r+32),
// i.e. 128 gate columns, resident in shared memory for the whole sequence; new hidden state exchanged through distributed
// shared memory + one cluster barrier per step), but the 128x256 by 256xNB product of every step is ONE chain of
// tcgen05.mma kind::tf32 instructions (M = 128 gate columns, N = 16 crops, K = 256 in 32 steps of 8) accumulating in
// TMEM.  Operands stay FP32 in shared memory (TF32 reads the top 19 bits), written by the threads directly in the
// canonical K-major SWIZZLE_128B layout:
//     byte offset of element (row, k) = kblock * (rows * 128) + (row / 8) * 1024 + (row % 8) * 128
//                                       + (((k % 32) / 4) ^ (row % 8)) * 16 + (k % 4) * 4 ,   kblock = k / 32
// The cell state, the gate non-linearities and the input projections stay FP32.
#include <cooperative_groups.h>

#include "engine.h"

namespace cg = cooperative_groups;

namespace bbocr {

namespace {

constexpr int NBT = 16;          // crops per cluster = UMMA N
constexpr int CL = 8;
constexpr int UNITS = 32;
constexpr int COLS = 128;        // UMMA M
constexpr int W_BYTES = 8 * COLS * 128;          // 8 k-blocks x 128 rows x 128 B = 128 KiB
constexpr int H_BYTES = 8 * NBT * 128;           // 8 k-blocks x 16 rows x 128 B  = 16 KiB per buffer

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int swz(int row, int kk) {         // within one k-block (32 floats per row)
    return (row >> 3) * 1024 + (row & 7) * 128 + ((((kk >> 2) ^ (row & 7)) & 7) << 4) + ((kk & 3) << 2);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void st_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename TO>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256, 1)
    k_lstm_tc(const float* __restrict__ gates_in, const float* __restrict__ w_hh, TO* __restrict__ out,
              const SeqDesc* __restrict__ seqs, const int* __restrict__ groups /*[n_groups][NBT], -1 = empty*/) {
    extern __shared__ uint8_t lraw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lraw) + 1023) & ~(uintptr_t)1023);
    uint8_t* Ws = sm;                               // W_BYTES
    uint8_t* Hs = sm + W_BYTES;                     // 2 x H_BYTES
    float* G = reinterpret_cast<float*>(Hs + 2 * H_BYTES);     // [4 gates][NBT][32 units]
    __shared__ uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int s_tmax[8];
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();
    const int group = blockIdx.x / CL, dir = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // resident weight slice in UMMA A layout: row c = g*32 + u  <-  W_hh[dir][k][g*256 + 32r + u]
    const float* wd = w_hh + (size_t)dir * 256 * 1024;
    for (int i = tid; i < 256 * COLS; i += 256) {
        int k = i / COLS, c = i - k * COLS;
        int g = c >> 5, u = c & 31;
        float v = __ldg(wd + (size_t)k * 1024 + g * 256 + UNITS * r + u);
        *reinterpret_cast<float*>(Ws + (k >> 5) * (COLS * 128) + swz(c, k & 31)) = v;
    }
    for (int i = tid; i < 2 * H_BYTES / 4; i += 256) reinterpret_cast<float*>(Hs)[i] = 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // phase-2 role: thread -> crop pb = tid / 16, units pu, pu + 1 (pu = (tid % 16) * 2)
    const int pb = tid >> 4, pu = (tid & 15) * 2;
    int row0 = 0, T = 0;
    {
        int s = groups[group * NBT + pb];
        if (s >= 0) { row0 = seqs[s].row0; T = seqs[s].T; }
    }
    int Tmax = T;
    for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
    if (lane == 0) s_tmax[warp] = Tmax;
    asm volatile("fence.proxy.async;" ::: "memory");            // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    Tmax = 0;
    for (int i = 0; i < 8; ++i) Tmax = max(Tmax, s_tmax[i]);
    const uint32_t tmem_acc = tmem_slot;
    float cstate[2] = {0.f, 0.f};
    cluster.sync();                                             // peers' buffers are zeroed before anyone writes into them

    // instruction descriptor: D = F32 (1<<4), A = B = TF32 (2<<7, 2<<10), K-major, N = 16 (>>3 at bit 17), M = 128 (>>4 at bit 24)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NBT >> 3) << 17) | ((uint32_t)(COLS >> 4) << 24);
    const uint32_t ws_addr = smem_u32(Ws);
    int cur = 0;
    for (int s = 0; s < Tmax; ++s) {
        // ---- mat-vec on the tensor core: D[128 x 16] = Ws[128 x 256] * Hs[cur][16 x 256]^T --------------------------
        if (tid == 0) {
            const uint32_t hs_addr = smem_u32(Hs + cur * H_BYTES);
#pragma unroll 1
            for (int kb = 0; kb < 8; ++kb) {
                const uint64_t ad = desc_sw128(ws_addr + kb * (COLS * 128));
                const uint64_t bd = desc_sw128(hs_addr + kb * (NBT * 128));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t accum = (kb | ks) ? 1u : 0u;
                    asm volatile(
                        "{\n\t"
                        ".reg .pred p;\n\t"
                        "setp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
                        "}\n" ::"r"(tmem_acc),
                        "l"(ad + 2 * ks), "l"(bd + 2 * ks), "r"(idesc), "r"(accum)
                        : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar))
                         : "memory");
        }
        // ---- meanwhile: this step's input projections for the thread's (crop, 2 units) --------------------------------
        const bool active = s < T;
        const int t = dir ? T - 1 - s : s;
        float gin[4][2];
#pragma unroll
        for (int g = 0; g < 4; ++g) gin[g][0] = gin[g][1] = 0.f;
        if (active) {
            const float* gp = gates_in + (size_t)(row0 + t) * 2048 + dir * 1024 + UNITS * r + pu;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float2 v = __ldg(reinterpret_cast<const float2*>(gp + g * 256));
                gin[g][0] = v.x; gin[g][1] = v.y;
            }
        }
        // ---- accumulator -> shared staging G[gate][crop][unit] ---------------------------------------------------------
        {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "LSTM_WAIT:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra LSTM_DONE;\n\t"
                "bra LSTM_WAIT;\n\t"
                "LSTM_DONE:\n\t"
                "}\n" ::"r"(smem_u32(&mma_bar)),
                "r"((uint32_t)(s & 1))
                : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int g = warp & 3, chalf = warp >> 2;          // TMEM lane quarter = gate; warps 4-7 take crops 8..15
            uint32_t v[8];
            const uint32_t taddr = tmem_acc + ((uint32_t)(g * 32) << 16) + chalf * 8;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j) G[(g * NBT + chalf * 8 + j) * UNITS + lane] = __uint_as_float(v[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // ---- gates, cell, hidden state; push h(t) into every peer's next-step operand buffer --------------------------
        {
            float hv[2] = {0.f, 0.f};
            if (active) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float pre[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) pre[g] = gin[g][q] + G[(g * NBT + pb) * UNITS + pu + q];
                    float ig = 1.f / (1.f + expf(-pre[0]));
                    float fg = 1.f / (1.f + expf(-pre[1]));
                    float gg = tanhf(pre[2]);
                    float og = 1.f / (1.f + expf(-pre[3]));
                    cstate[q] = fg * cstate[q] + ig * gg;
                    hv[q] = og * tanhf(cstate[q]);
                }
                TO* o = out + (size_t)(row0 + t) * 512 + dir * 256 + UNITS * r + pu;
                st_out(o, hv[0]);
                st_out(o + 1, hv[1]);
            }
            // element (crop pb, k = 32r + pu): k-block r of the B operand; pu even => the float2 stays inside a 16-byte chunk
            const int off = (cur ^ 1) * H_BYTES + r * (NBT * 128) + swz(pb, pu);
#pragma unroll
            for (int q = 0; q < CL; ++q) {
                uint8_t* remote = cluster.map_shared_rank(Hs, q);
                *reinterpret_cast<float2*>(remote + off) = make_float2(hv[0], hv[1]);
            }
        }
        asm volatile("fence.proxy.async;" ::: "memory");        // the peers' tensor cores read these stores next step
        cluster.sync();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        cur ^= 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_acc) : "memory");
}

}  // namespace

void lstm_sequences_tc(Handle* h, Lane& lane, const float* gates_in, const float* w_hh, int n_seq, const SeqDesc* seqs_dev,
                       const int* groups_dev, int n_groups, void* out) {
    cudaStream_t st = lane.stream;
    if (n_seq == 0) return;
    const size_t smem = (size_t)W_BYTES + 2 * H_BYTES + 4 * NBT * UNITS * sizeof(float) + 1024;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL * n_groups, 2);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (h->precision == BBOCR_PREC_BF16) {
        CUDA_CHECK(cudaFuncSetAttribute(k_lstm_tc<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_lstm_tc<__nv_bfloat16>, gates_in, w_hh, (__nv_bfloat16*)out, seqs_dev, groups_dev));
    } else {
        CUDA_CHECK(cudaFuncSetAttribute(k_lstm_tc<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_lstm_tc<float>, gates_in, w_hh, (float*)out, seqs_dev, groups_dev));
    }
    count_launch(h);
}

int lstm_tc_group_size() { return NBT; }

}  // namespace bbocr
