// engine.h -- internal interfaces between the translation units of libbbocr.so
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <memory>
#include <set>

#include "common.cuh"

namespace bbocr {

struct CraftW {
    ConvW c1_1_tc;      // conv1_1 as a 1x1 convolution over the 32-channel gathered neighbourhood (tensor-core path)
    ConvW c1_1, c1_2, c2_1, c2_2, c3_1, c3_2, c3_3, c4_1, c4_2, c4_3, c5_1, c5_2, fc6, fc7;
    ConvW up1a, up1b, up2a, up2b, up3a, up3b, up4a, up4b, cls0, cls1, cls2, cls3, cls4;
};
struct CrnnW {
    ConvW c0, c1, c2, c3, c4, c5, c6;
    LstmW l0, l1;
    ConvW pred;
    int num_class = 97;
};

struct LaneScratch {                // grow-only device buffer owned by ONE lane: reused in stream order, so it never creates the
    void* p = nullptr;              // cross-stream reuse dependencies a shared cudaMallocAsync pool inserts between lanes
    size_t cap = 0;
    void* get(size_t n) {
        if (n > cap) {
            if (p) cudaFree(p);     // synchronises the device; happens only while the buffers grow to the working size
            p = nullptr;
            cap = n + n / 8 + 4096;
            CUDA_CHECK(cudaMalloc(&p, cap));
        }
        return p;
    }
};

struct Lane {                       // one in-flight page: a stream + its pinned staging
    cudaStream_t stream = nullptr;
    PinnedBuf pin_in, pin_out;
    bool in_busy = false;           // an async H2D copy out of pin_in may still be in flight
    LaneScratch scr[3];             // jpeg.cu: coefficient blocks, sample planes, tables + entropy bytes
    bool scr_dirty = false;         // jpeg.cu: scr[0] may hold coefficients of an image whose IDCT was never enqueued
};

struct Handle {
    int device = 0;
    int sm_count = 148;
    int precision = BBOCR_PREC_FP32;
    bool det_split = false;         // BBOCR_PREC_BF16X3: the detector runs in split precision as well (3 x bf16 products ~ FP32)
    std::string err;
    std::mutex mu;                  // one public call at a time per handle (Reader is shared between threads:
                                    // batch_processor_enhanced.py:215-216 + enhanced_extractor.py:97-98)
    std::atomic<int64_t> launches{0};
    bool craft_loaded = false, crnn_loaded = false;
    CraftW craft;
    CrnnW crnn;
    std::vector<void*> owned;       // weight allocations
    std::vector<Lane> lanes;         // [0, n_det_lanes) detector lanes, then the recogniser lanes
    int n_det_lanes = 8;
    int det_batch = 2;               // same-size pages that share one detector-network launch chain
    int rec_group = 32;              // pages per recogniser launch group (bbocr_readtext_batch)
    // dominant-kernel instrumentation (bench.py roofline): CUDA events around every implicit-GEMM conv launch
    bool conv_timing = false;
    bool force_generic_conv = false; // test hook: route BF16-mode convolutions through the CUDA-core kernel
    bool tc_attr_set = false, lstm_attr_set = false, lstm_mma_attr_set = false, stem_attr_set = false;
    std::mutex stat_mu;
    std::map<std::array<int, 5>, std::pair<void*, int>> cubic_cache;   // preprocess.cu: INTER_CUBIC tables per geometry
    std::set<int> res_attr_done;     // conv_res.cu kernel variants whose smem attribute is set on this device
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> conv_events;
    double conv_flops = 0;
    double stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t conv_launches = 0;
    std::vector<Lane> jpeg_lanes;          // streams + pinned staging of bbocr_jpeg_decode_batch (created on first use)
    std::set<std::vector<int32_t>> dict;   // wordbeamsearch dictionary as class-index words (bbocr_set_dictionary)
};

inline void count_launch(Handle* h, int n) { h->launches.fetch_add(n, std::memory_order_relaxed); }

// ---- preprocess.cu ---------------------------------------------------------------------------------------------------
void pp_gray(Handle*, cudaStream_t, const uint8_t* bgr, int H, int W, int stride, uint8_t* out);
void pp_resize_cubic(Handle*, cudaStream_t, const uint8_t* src, int sH, int sW, uint8_t* dst, int dH, int dW, int mode);
void pp_gaussian3(Handle*, cudaStream_t, const uint8_t* src, uint8_t* dst, int H, int W, float sigma,
                  unsigned long long* sum_out);
void pp_sum(Handle*, cudaStream_t, const uint8_t* src, int64_t n, unsigned long long* sum);
void pp_tone_lut(Handle*, cudaStream_t, const unsigned long long* sum, int64_t npix, float contrast, float brightness,
                 uint8_t* lut);
void pp_apply_lut(Handle*, cudaStream_t, const uint8_t* src, uint8_t* dst, int64_t n, const uint8_t* lut);
void pp_equalize_hist(Handle*, cudaStream_t, const uint8_t* src, uint8_t* dst, int H, int W);      // cv2.equalizeHist
void pp_clahe_luts(Handle*, cudaStream_t, const uint8_t* src, int H, int W, float clip, const uint8_t* tone,
                   unsigned int* hist, uint8_t* luts);
void pp_clahe_apply(Handle*, cudaStream_t, const uint8_t* src, uint8_t* dst, int H, int W, const uint8_t* tone,
                    const uint8_t* luts);
void pp_unsharp(Handle*, cudaStream_t, const uint8_t* src, uint8_t* dst, int H, int W, int percent, int threshold,
                const uint8_t* tone, const uint8_t* clahe_luts);
void pp_adaptive_threshold(Handle*, cudaStream_t, const uint8_t* src, uint8_t* dst, int H, int W, int method, int inv,
                           int block, float delta);
void preprocess_chain_dev(Handle*, cudaStream_t, const uint8_t* bgr, int H, int W, int stride, const bbocr_pp_params&,
                          uint8_t* out, int* outH, int* outW);
int preprocess_launches_per_image();
float pp_deskew(Handle*, cudaStream_t, const uint8_t* src, uint8_t* dst, int H, int W, float max_deg, int variant = -1,
                std::vector<unsigned long long>* scores_out = nullptr);

// ---- autocrop.cu (enhanced_extractor.py:239-372) ---------------------------------------------------------
struct AutoCropDebug {             // optional parity outputs (host memory)
    uint8_t* mask = nullptr;       // HxW composite text mask (0/255)
    uint8_t* merged = nullptr;     // HxW mask after the two morphology variants
    int otsu[2] = {0, 0};          // Otsu thresholds of the equalised image and of the Sobel magnitude
    std::vector<int32_t> boxes;    // (x, y, w, h) of every external contour of `merged`, sorted by (y, x, w, h)
};
// img: device HxWx3 BGR (row stride in bytes) or a packed HxW gray plane (channels = 1).  Returns false for the reference's "no crop" (None); rect = (x0, y0, x1, y1).
bool autocrop_dev(Handle*, cudaStream_t, const uint8_t* img, int H, int W, int channels, int stride, int margin, int32_t rect[4],
                  AutoCropDebug* dbg);
void external_boxes_dev(Handle*, cudaStream_t, const uint8_t* binary_dev, int H, int W, std::vector<int32_t>& boxes);
void rect_morph_dev(Handle*, cudaStream_t, const uint8_t* binary_dev, int H, int W, int kw, int kh, bool erode, uint8_t* out_dev);

// ---- nn.cu : layer kernels (NHWC; T = float | __nv_bfloat16 chosen by Handle::precision) ------------------------------
enum ConvFlags { CONV_RELU = 1, CONV_OUT_F32 = 2, CONV_POOL22 = 4, CONV_POOL21 = 8 };
// out = epilogue(conv(concat_channels(in1, in2)))   in2 may be empty (C == 0).  'same' geometry unless pad says otherwise.
// `pooled` (optional): also produce MaxPool2d(2,2) [CONV_POOL22] / MaxPool2d((2,1)) [CONV_POOL21] of the output; out.p may
// then be null when only the pooled tensor is needed.
void conv_forward(Handle*, cudaStream_t, const ConvW&, const Act& in1, const Act& in2, Act& out, int flags,
                  Act* pooled = nullptr, const uint8_t* colmask = nullptr);     // colmask [out.W]: 0 = write zeros (tcgen05 path only)
// cin in {1,3(stored 4)} direct convolution from an FP32 NHWC tensor (canvas / crop batch)
void conv_first(Handle*, cudaStream_t, const ConvW&, const float* in, int N, int H, int W, int cstride, Act& out,
                int flags);
void maxpool(Handle*, cudaStream_t, const Act& in, Act& out, int kh, int kw, int sh, int sw, int ph, int pw);
void upsample2x(Handle*, cudaStream_t, const Act& in, Act& out);          // bilinear, align_corners=False
void mean_rows(Handle*, cudaStream_t, const Act& in, Act& out);           // AdaptiveAvgPool2d((None,1)) after permute
void cls_tail(Handle*, cudaStream_t, const ConvW& c3, const ConvW& c4, const Act& in, float* text, float* link);
// split-precision (hi + lo bf16) glue of the detector in bf16x3 mode
void maxpool_split(Handle*, cudaStream_t, const Act& in, Act& out, int kh, int kw, int sh, int sw, int ph, int pw);
void upsample2x_split(Handle*, cudaStream_t, const Act& in, Act& out);
void cls_tail_f32(Handle*, cudaStream_t, const ConvW& c3, const ConvW& c4, const Act& in, float* text, float* link);

Act act_alloc_split(Handle*, cudaStream_t, DevBuf& buf, int N, int H, int W, int C);
void maxpool_f32_to_split(Handle*, cudaStream_t, const Act& in, Act& out, int kh, int kw, const uint8_t* colmask = nullptr);
// CRNN stem fused: Conv(1->32) + ReLU + MaxPool2d(2,2) -> split tensor (bit-identical to conv_first + maxpool_f32_to_split)
void conv0_pool_split(Handle*, cudaStream_t, const ConvW&, const float* in, int N, int H, int W, Act& out, const uint8_t* colmask = nullptr);
// ragged AdaptiveAvgPool: crop i = columns [meta[3i], meta[3i] + meta[3i+1]) of `in` ([1][H][W][C] split) -> rows meta[3i+2].. of out
void mean_rows_split_ragged(Handle*, cudaStream_t, const Act& in, const Act& out, const int* meta_dev, int n_crops, int t_max);
void mean_rows_split(Handle*, cudaStream_t, const Act& in, Act& out);
void act_from_f32(Handle*, cudaStream_t, const float* in, void* out, int64_t n);
void act_to_f32(Handle*, cudaStream_t, const void* in, float* out, int64_t n);
size_t act_elem_size(const Handle*);
Act act_alloc(Handle*, cudaStream_t, DevBuf& buf, int N, int H, int W, int C, bool force_f32 = false);

// ---- conv_tc.cu : tcgen05 implicit GEMM ------------------------------------------------------------------------------
bool conv_tc_supported(const ConvW&, const Act& in1, const Act& in2);
bool conv_stem_supported(const ConvW&, const Act& out);
void stem_norm_forward(Handle*, cudaStream_t, const uint8_t* img, int th, int tw, void* dst, int H, int W, const float* mean,
                       const float* sd);
void conv_stem_forward(Handle*, cudaStream_t, const ConvW&, const void* norm, Act& out, int flags);
void conv_tc_forward(Handle*, cudaStream_t, const ConvW&, const Act& in1, const Act& in2, Act& out, int flags, Act* pooled,
                     const uint8_t* colmask = nullptr);

// conv_res.cu : 3x3 / pad 1 convolutions of the low-channel layers with resident weights and a halo patch per tile
bool conv_res_supported(const ConvW&, const Act& in1, const Act& in2, const Act& out);
void conv_res_forward(Handle*, cudaStream_t, const ConvW&, const Act& in1, const Act& in2, Act& out, int flags, Act* pooled);

bool conv_res_cls_tail_supported(const ConvW& c2, const ConvW& c3, const ConvW& c4, const Act& in);
void conv_res_cls_tail(Handle*, cudaStream_t, const ConvW& c2, const ConvW& c3, const ConvW& c4, const Act& in, float* text, float* link);

// ---- weights.cu ------------------------------------------------------------------------------------------------------
void load_craft(Handle*, const bbocr_tensor* t, int n);
void load_crnn(Handle*, const bbocr_tensor* t, int n);
ConvW make_conv_raw(Handle*, const float* w, const float* bias, int cout, int cin, int kh, int kw, int pad, int dil);

// ---- craft.cu --------------------------------------------------------------------------------------------------------
struct CanvasGeom {
    int H, W;            // source image
    int th, tw;          // after resize_aspect_ratio
    int H32, W32;        // padded canvas
    double ratio;
};
CanvasGeom canvas_geom(int H, int W, int canvas_size, double mag_ratio);
// img_dev: HxWx3 u8 device.  text/link: device float maps (H32/2 x W32/2)
void craft_forward_dev(Handle*, cudaStream_t, const uint8_t* img_dev, const CanvasGeom&, float* text, float* link);
void craft_forward_batch_dev(Handle*, cudaStream_t, const uint8_t* const* imgs_dev, int nimg, const CanvasGeom&, float* text,
                             float* link);     // nimg same-size images as one NHWC batch; maps [nimg][H32/2][W32/2]
void resize_bilinear_u8(Handle*, cudaStream_t, const uint8_t* src, int sH, int sW, int sstride, int C, uint8_t* dst,
                        int dH, int dW);

// ---- postproc.cu + boxes.cpp -----------------------------------------------------------------------------------------
struct DetComponents {             // kept components of one page, host side
    int n_labels = 0;              // all foreground components (OpenCV nLabels - 1)
    std::vector<int> comp_x, comp_y, comp_w, comp_h, comp_area;
    std::vector<int> row_off;      // per kept comp: offset into row_min/row_max (comp_h entries each)
    std::vector<int> row_min, row_max;   // per source row: min/max x of (component minus link-only) pixels; min > max = empty
};
void det_components_dev(Handle*, Lane&, const float* text, const float* link, int mh, int mw, float text_threshold,
                        float link_threshold, float low_text, DetComponents& out);
void boxes_from_components(const DetComponents&, int mh, int mw, std::vector<float>& boxes /*n*8*/);
// the same on the device incl. hull + min-area rectangle (one D2H of the finished boxes); false = capacity exceeded, use the two above
bool det_boxes_dev(Handle*, Lane&, const float* text, const float* link, int mh, int mw, float text_threshold,
                   float link_threshold, float low_text, std::vector<float>& boxes, int* n_labels);
void min_area_box(const int32_t* xy, int n, float* out8);
void debug_convex_hull(const int32_t* xy, int n, int clockwise, std::vector<int>& hull);
void group_boxes(const float* boxes, int n, double ratio, const bbocr_group_params& p, std::vector<int32_t>& hlist,
                 std::vector<double>& flist);
void free_box_transform(const double* quad, int* max_w, int* max_h, double* Minv);

// ---- recog.cu --------------------------------------------------------------------------------------------------------
struct CropDesc {                  // one crop job (device-visible POD)
    int x0, y0, w, h;              // horizontal: source rect in the gray page.  free-form: x0 = scratch offset of the
                                   // warped patch (w x h), y0 unused
    int ow, oh;                    // compute_ratio_and_resize output size
    int model_w;                   // ceil(ratio)*64
    int off;                       // byte offset of the (oh x ow) crop in the packed crop buffer
    int aoff;                      // byte offset of the AlignCollate'd (64 x resized_w) crop (== off for wide crops)
    int resized_w;                 // AlignCollate resized width (<= model_w)
    int slot, bucket_off;          // slot inside its width bucket / float offset of the bucket's input tensor
    int free_idx;                  // >= 0 : free-form box (index into the 3x3 warp matrices), else -1
};
void crops_dev(Handle*, cudaStream_t, const uint8_t* gray, int H, int W, const CropDesc* descs_dev, int n,
               const CropDesc* descs_host, const double* warp_dev, uint8_t* scratch, uint8_t* crops);
void crop_hist_dev(Handle*, cudaStream_t, const uint8_t* crops, const CropDesc* descs_dev, int n, unsigned int* hist);
void crop_contrast_dev(Handle*, cudaStream_t, const uint8_t* crops, const CropDesc* descs_dev, int n, const double* low,
                       const double* ratio, const int* apply, uint8_t* out);
void pil_resize_bicubic_dev(Handle*, cudaStream_t, const uint8_t* src, int sH, int sW, uint8_t* dst, int dH, int dW,
                            uint8_t* scratch, float box_w = 0.f, float box_h = 0.f);
void pil_reduce_dev(Handle*, cudaStream_t, const uint8_t* src, int sH, int sW, int fx, int fy, uint8_t* dst);   // Image.reduce
void crops_to_input_dev(Handle*, cudaStream_t, const uint8_t* aligned, const CropDesc* descs_dev, int n, int max_model_w,
                        float* inputs);
struct SeqDesc {                   // one crop's feature sequence inside the flat [rows][channels] tensors
    int row0, T;
};
// conv stack + row mean of one width bucket: x [N][64][Wm] FP32 -> seq rows [row0, row0 + N*(Wm/4-1)) of `seq` ([rows][256])
Act crnn_alloc_seq(Handle*, cudaStream_t, DevBuf& buf, int rows);
void crnn_features_dev(Handle*, cudaStream_t, const float* x, int N, int Wm, const Act& seq, int row0);
// ragged variant (throughput mode): all crops side by side in ONE strip image [64][Wtot] with >= 16 zero columns between
// neighbours; mask1 [Wtot/2] / mask2 [Wtot/4] mark the crop columns at the two pooled resolutions; meta as above
bool crnn_ragged(const Handle*);
void crnn_features_strip_dev(Handle*, cudaStream_t, const float* strip, int Wtot, const uint8_t* mask1, const uint8_t* mask2,
                             const int* meta_dev, int n_crops, int t_max, const Act& seq);
void crops_to_strip_dev(Handle*, cudaStream_t, const uint8_t* aligned, const CropDesc* descs_dev, const int* xoff_dev, int n,
                        int max_model_w, int Wtot, float* strip);
// both BiLSTM blocks + Prediction over all sequences at once: seq [rows][256] -> logits [rows][num_class] FP32
void crnn_sequence_dev(Handle*, Lane&, const Act& seq, const std::vector<SeqDesc>& seqs, float* logits);
void crnn_forward_dev(Handle*, Lane&, const float* x, int N, int Wm, float* logits);
void lstm_sequences(Handle*, Lane&, const float* gates_in, const float* w_hh, const SeqDesc* seqs_host, int n_seq,
                    const SeqDesc* seqs_dev, const int* groups_dev, int n_groups, void* out, void* out_lo = nullptr);
int lstm_group_size(const Handle*);
// lstm_mma.cu: 128 crops per 16-CTA cluster, split-precision tcgen05 recurrence (throughput mode default)
void lstm_sequences_mma(Handle*, Lane&, const float* gates_in, const float* w_hh, int n_seq, const SeqDesc* seqs_dev,
                        const int* groups_dev, int n_groups, void* out, void* out_lo, int out_mode);
int lstm_mma_group_size();
// greedy CTC over all rows: per-row argmax / renormalised max probability, then per-sequence collapse
void ctc_decode_dev(Handle*, cudaStream_t, const float* logits, int rows, int C, const uint8_t* ignore_dev,
                    const SeqDesc* seqs_dev, int n_seq, int32_t* text_idx /*[rows]*/, int32_t* text_len /*[n_seq]*/,
                    float* step_prob /*[rows]*/, int32_t* step_idx /*[rows]*/);
// recognizer_predict's probability matrix (softmax, ignored classes zeroed, renormalised): probs [rows][C] FP32
void row_probs_dev(Handle*, cudaStream_t, const float* logits, int rows, int C, const uint8_t* ignore_dev, float* probs);
// rotation_info: out = np.rot90(src, k) for every job (k = 1, 2, 3), u8 crops inside one packed buffer
struct RotDesc { int src_off, dst_off, sh, sw, k; };
void rotate_crops_dev(Handle*, cudaStream_t, uint8_t* crops, const RotDesc* descs_dev, int n, int max_pixels);
// jpeg.cu: baseline JPEG -> BGR / luma planes on the device, bit-exact with cv2.imdecode (libjpeg-turbo defaults + EXIF orientation)
void jpeg_info(const uint8_t* data, size_t n, int* H, int* W, int* channels, int* orientation);
long long jpeg_coefficients_host(const uint8_t* data, size_t n, int16_t* out, long long cap_blocks);
struct JpegJob { const uint8_t* data; size_t n; uint8_t* out_bgr; uint8_t* out_gray; int H, W; };   // H, W: set by the decoder
void jpeg_decode_group_dev(Handle*, Lane&, JpegJob* jobs, int nj, int ignore_orientation);
void jpeg_decode_dev(Handle*, Lane&, const uint8_t* data, size_t n, int ignore_orientation, uint8_t* out_bgr, uint8_t* out_gray,
                     int* outH, int* outW);
// beam.cpp: CTCLabelConverter.decode_beamsearch (decoder 1) / decode_wordbeamsearch (decoder 2) of one crop, host side
void decode_beam(const float* probs, int T, int C, int decoder, int beam_width, int space_idx,
                 const std::set<std::vector<int32_t>>* dict, std::vector<int32_t>& text);

}  // namespace bbocr
