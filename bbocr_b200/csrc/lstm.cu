// lstm.cu -- BiLSTM recurrence of the CRNN (easyocr/model/modules.py::BidirectionalLSTM, nn.LSTM(256, 256,
// bidirectional=True); gate order i,f,g,o; h0 = c0 = 0).  SURVEY.md §8a B11, §7.3-6.
//
// The recurrence is a chain of T dependent 256x1024 mat-vecs per crop and direction: latency-bound, and W_hh
// (1 MiB in FP32) does not fit one SM.  One thread-block CLUSTER of 8 CTAs therefore owns one (group of NB crops,
// direction): CTA r keeps the 128 gate columns of hidden units [32r, 32r+32) resident in its shared memory (128 KiB)
// for the whole sequence, every step each CTA computes its 32 units for all NB crops, and the new hidden state is
// exchanged through distributed shared memory (one 128-float block per peer) followed by one cluster barrier.
// The input projections x_t W_ih^T + b_ih + b_hh for all time steps come from one tensor-core GEMM (conv_tc.cu).
//
//   gates_in : [rows][2048] FP32, row = seq.row0 + t, columns [fwd i,f,g,o | bwd i,f,g,o]
//   w_hh     : [2][256 k][1024] FP32 (k-major)
//   out      : [rows][512] = [h_fwd(t) | h_bwd(t)]   (activation dtype of the precision mode)
#include <cooperative_groups.h>

#include "engine.h"

namespace cg = cooperative_groups;

namespace bbocr {

namespace {

constexpr int NB = 8;            // crops per cluster (phase 2 maps one thread to each (crop, unit) pair: NB * 32 = 256)
constexpr int CL = 8;            // CTAs per cluster
constexpr int UNITS = 32;        // hidden units per CTA
constexpr int COLS = 4 * UNITS;  // gate columns per CTA
constexpr int KPARTS = 8;        // k-split of the 256-long dot products: one warp per 32-wide k range

__device__ __forceinline__ void st_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename TO>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256, 1)
    k_lstm_cluster(const float* __restrict__ gates_in, const float* __restrict__ w_hh, TO* __restrict__ out,
                   __nv_bfloat16* __restrict__ out_lo /* split-precision output when non-null (TO = bf16) */,
                   const SeqDesc* __restrict__ seqs, const int* __restrict__ groups /*[n_groups][NB], -1 = empty*/) {
    extern __shared__ __align__(16) float lsm[];
    float* Ws = lsm;                               // [256][COLS]
    float* hbuf = Ws + 256 * COLS;                 // [2][NB][256]
    float* part = hbuf + 2 * NB * 256;             // [KPARTS][NB][COLS]
    float* hstage = part + KPARTS * COLS * NB;     // [NB][UNITS]
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();
    const int group = blockIdx.x / CL, dir = blockIdx.y;
    const int tid = threadIdx.x;

    // resident weight slice: Ws[k][g*32 + u] = W_hh[dir][k][g*256 + 32r + u]
    const float* wd = w_hh + (size_t)dir * 256 * 1024;
    for (int i = tid; i < 256 * COLS; i += 256) {
        int k = i / COLS, c = i - k * COLS;
        int g = c >> 5, u = c & 31;
        Ws[i] = __ldg(wd + (size_t)k * 1024 + g * 256 + UNITS * r + u);
    }
    for (int i = tid; i < 2 * NB * 256; i += 256) hbuf[i] = 0.f;

    // phase-2 role: thread (b, u) for tid < NB*32
    const int pb = tid >> 5, pu = tid & 31;
    int row0 = 0, T = 0;
    {
        int s = groups[group * NB + pb];
        if (s >= 0) { row0 = seqs[s].row0; T = seqs[s].T; }
    }
    int Tmax = T;
    for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
    __shared__ int s_tmax[8];
    if ((tid & 31) == 0) s_tmax[tid >> 5] = Tmax;
    __syncthreads();
    Tmax = 0;
    for (int i = 0; i < NB; ++i) Tmax = max(Tmax, s_tmax[i]);
    float cstate = 0.f;
    cluster.sync();                                 // every CTA's hbuf is zeroed before any peer writes into it

    // phase-1 role: warp kp owns k in [32 kp, 32 kp + 32); lane cgp owns gate columns [4 cgp, 4 cgp + 4).  All lanes of a
    // warp read the same h values (shared-memory broadcast) and consecutive 16-byte weight vectors (conflict-free).
    const int kp = tid >> 5, cgp = tid & 31;
    int cur = 0;
    for (int s = 0; s < Tmax; ++s) {
        // prefetch this step's input projections (latency hidden behind the mat-vec)
        float gin[4] = {0.f, 0.f, 0.f, 0.f};
        const bool active = s < T;
        const int t = dir ? T - 1 - s : s;
        if (active) {
            const float* g = gates_in + (size_t)(row0 + t) * 2048 + dir * 1024 + UNITS * r + pu;
            gin[0] = __ldg(g); gin[1] = __ldg(g + 256); gin[2] = __ldg(g + 512); gin[3] = __ldg(g + 768);
        }
        float acc[NB][4];
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.f;
        const float* hb = hbuf + cur * NB * 256;
        const int k0 = kp * 32;
#pragma unroll 2
        for (int k = k0; k < k0 + 32; k += 4) {
            float4 w0 = *reinterpret_cast<const float4*>(Ws + (k + 0) * COLS + 4 * cgp);
            float4 w1 = *reinterpret_cast<const float4*>(Ws + (k + 1) * COLS + 4 * cgp);
            float4 w2 = *reinterpret_cast<const float4*>(Ws + (k + 2) * COLS + 4 * cgp);
            float4 w3 = *reinterpret_cast<const float4*>(Ws + (k + 3) * COLS + 4 * cgp);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                float4 hv = *reinterpret_cast<const float4*>(hb + b * 256 + k);
                acc[b][0] = fmaf(w0.x, hv.x, acc[b][0]); acc[b][1] = fmaf(w0.y, hv.x, acc[b][1]);
                acc[b][2] = fmaf(w0.z, hv.x, acc[b][2]); acc[b][3] = fmaf(w0.w, hv.x, acc[b][3]);
                acc[b][0] = fmaf(w1.x, hv.y, acc[b][0]); acc[b][1] = fmaf(w1.y, hv.y, acc[b][1]);
                acc[b][2] = fmaf(w1.z, hv.y, acc[b][2]); acc[b][3] = fmaf(w1.w, hv.y, acc[b][3]);
                acc[b][0] = fmaf(w2.x, hv.z, acc[b][0]); acc[b][1] = fmaf(w2.y, hv.z, acc[b][1]);
                acc[b][2] = fmaf(w2.z, hv.z, acc[b][2]); acc[b][3] = fmaf(w2.w, hv.z, acc[b][3]);
                acc[b][0] = fmaf(w3.x, hv.w, acc[b][0]); acc[b][1] = fmaf(w3.y, hv.w, acc[b][1]);
                acc[b][2] = fmaf(w3.z, hv.w, acc[b][2]); acc[b][3] = fmaf(w3.w, hv.w, acc[b][3]);
            }
        }
        // partial sums: part[kp][b][col] (column fastest: conflict-free 16-byte stores and phase-2 loads)
#pragma unroll
        for (int b = 0; b < NB; ++b)
            *reinterpret_cast<float4*>(part + ((kp * NB + b) * COLS + 4 * cgp)) = make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
        __syncthreads();
        // phase 2: gates, cell and hidden state of unit pu for crop pb
        {
            float hv = 0.f;
            if (active) {
                float pre[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float a = gin[g];
#pragma unroll
                    for (int q = 0; q < KPARTS; ++q) a += part[(q * NB + pb) * COLS + g * UNITS + pu];
                    pre[g] = a;
                }
                float ig = 1.f / (1.f + expf(-pre[0]));
                float fg = 1.f / (1.f + expf(-pre[1]));
                float gg = tanhf(pre[2]);
                float og = 1.f / (1.f + expf(-pre[3]));
                cstate = fg * cstate + ig * gg;
                hv = og * tanhf(cstate);
                const size_t oidx = (size_t)(row0 + t) * 512 + dir * 256 + UNITS * r + pu;
                st_out(out + oidx, hv);
                if (out_lo) out_lo[oidx] = __float2bfloat16_rn(hv - __bfloat162float(__float2bfloat16_rn(hv)));
            }
            hstage[pb * UNITS + pu] = hv;
        }
        __syncthreads();
        // phase 3: warp q pushes this CTA's NB x 32 block into peer q's next-step buffer (distributed shared memory)
        {
            const int q = tid >> 5, l = tid & 31;
            float* remote = cluster.map_shared_rank(hbuf, q) + (cur ^ 1) * NB * 256;
#pragma unroll
            for (int b = 0; b < NB; ++b) remote[b * 256 + UNITS * r + l] = hstage[b * UNITS + l];
        }
        cluster.sync();
        cur ^= 1;
    }
}

}  // namespace

// recurrence kernel: FP32 mode = the CUDA-core cluster kernel above (parity mode).  BF16 (throughput) mode = lstm_mma.cu
// (BBOCR_LSTM=cluster selects the CUDA-core kernel in throughput mode as well, for comparison).
static int lstm_variant(const Handle* h) {
    if (h->precision != BBOCR_PREC_BF16 || h->force_generic_conv) return 0;
    static const int v = [] {
        const char* e = getenv("BBOCR_LSTM");
        if (!e) return 2;
        if (!strcmp(e, "cluster")) return 0;
        return 2;
    }();
    return v;
}

// seqs: device array of n_seq descriptors; the host copy is used to form groups of NB sequences of similar length
void lstm_sequences(Handle* h, Lane& lane, const float* gates_in, const float* w_hh, const SeqDesc* seqs_host, int n_seq,
                    const SeqDesc* seqs_dev, const int* groups_dev, int n_groups, void* out, void* out_lo) {
    cudaStream_t st = lane.stream;
    if (n_seq == 0) return;
    const int variant = lstm_variant(h);
    if (variant == 2) {            // crops on the UMMA M dimension, split precision (lstm_mma.cu)
        lstm_sequences_mma(h, lane, gates_in, w_hh, n_seq, seqs_dev, groups_dev, n_groups, out, out_lo, out_lo ? 1 : 0);
        return;
    }
    const size_t smem = (size_t)(256 * COLS + 2 * NB * 256 + KPARTS * COLS * NB + NB * UNITS) * sizeof(float);
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
        cudaFuncSetAttribute(k_lstm_cluster<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_lstm_cluster<__nv_bfloat16>, gates_in, w_hh, (__nv_bfloat16*)out, (__nv_bfloat16*)out_lo, seqs_dev, groups_dev));
    } else {
        CUDA_CHECK(cudaFuncSetAttribute(k_lstm_cluster<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_lstm_cluster<float>, gates_in, w_hh, (float*)out, (__nv_bfloat16*)nullptr, seqs_dev, groups_dev));
    }
    count_launch(h);
}

int lstm_group_size(const Handle* h) {
    const int v = lstm_variant(h);
    return v == 2 ? lstm_mma_group_size() : NB;
}

}  // namespace bbocr
