/* bbocr.h -- C ABI of libbbocr.so, the B200-native OCR stage of BB-OCR.
 *
 * BB-OCR itself has no FFI: its OCR stage is a duck-typed Python boundary
 *     pipeline_demo/extractor/enhanced_extractor.py:153   easyocr.Reader(["en"], gpu=use_gpu)
 *     pipeline_demo/extractor/enhanced_extractor.py:520   reader.readtext(path, paragraph=False, batch_size=1, workers=0)
 *     pipeline_demo/extractor/enhanced_extractor.py:39,431  preprocess_for_book_cover(image_path, output_path)
 * and the arithmetic behind it is EasyOCR 1.7.2 + OpenCV + Pillow.  This header is what a ctypes binding of that
 * boundary binds (INTEGRATION.md shows the stub); every entry point cites the reference / upstream function it
 * replaces.  Plain C: pointers + sizes, no C++ or torch types.  Every int-returning function returns 0 on success
 * and a negative BBOCR_E_* code on failure, with text available from bbocr_last_error().  Nothing here ever aborts
 * the process (the caller swallows exceptions: enhanced_extractor.py:529-531).
 *
 * Memory kinds: pointers are HOST pointers unless the parameter is documented "device" or guarded by an
 * `on_device` flag, in which case they are raw CUDA device pointers (e.g. torch.Tensor.data_ptr()).
 */
#ifndef BBOCR_H
#define BBOCR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BBOCR_OK 0
#define BBOCR_E_ARG (-1)      /* bad argument */
#define BBOCR_E_CUDA (-2)     /* CUDA runtime / driver error (message holds cudaGetErrorString) */
#define BBOCR_E_STATE (-3)    /* weights not loaded, wrong precision mode, ... */
#define BBOCR_E_NOMEM (-4)
#define BBOCR_E_UNSUPPORTED (-5)

#define BBOCR_PREC_FP32 0     /* CUDA-core FP32 implicit GEMM: the <=1e-3 parity mode */
#define BBOCR_PREC_BF16 1     /* tcgen05, detector in plain BF16 operands (FP32 accumulation in TMEM), recogniser in split
                                 precision: fastest mode; score maps within a stated 6e-2, boxes may move by a pixel */
#define BBOCR_PREC_BF16X3 2   /* tcgen05, detector AND recogniser in split precision: every operand is bf16 hi + bf16 lo and
                                 every product x_hi*w_hi + x_lo*w_hi + x_hi*w_lo accumulates in FP32 (TMEM): FP32-class score
                                 maps and logits on the tensor cores -- the parity-grade throughput mode */

typedef struct bbocr_handle bbocr_handle;

/* ---- lifecycle ------------------------------------------------------------------------------------------------ */
/* easyocr.Reader.__init__ (device selection).  Fails (BBOCR_E_CUDA) if `device` is not a CUDA device: there is no
 * CPU fallback. */
int bbocr_create(int device, bbocr_handle** out);
void bbocr_destroy(bbocr_handle* h);
const char* bbocr_last_error(const bbocr_handle* h);   /* h may be NULL: error of the last failed bbocr_create */
const char* bbocr_version(void);

/* A named FP32 host tensor in PyTorch state-dict layout (conv weight [Cout,Cin,kh,kw], LSTM weight_ih [4H,I], ...). */
typedef struct {
    const char* name;
    const float* data;
    int32_t ndim;
    int64_t shape[4];
} bbocr_tensor;

/* easyocr.detection.get_detector: load craft_mlt_25k state dict (keys basenet.slice*, upconv*, conv_cls.*). */
int bbocr_load_craft(bbocr_handle* h, const bbocr_tensor* tensors, int n);
/* easyocr.recognition.get_recognizer: load english_g2 state dict (FeatureExtraction.*, SequenceModeling.*, Prediction.*). */
int bbocr_load_crnn(bbocr_handle* h, const bbocr_tensor* tensors, int n);
int bbocr_set_precision(bbocr_handle* h, int prec);
int bbocr_get_precision(const bbocr_handle* h);

/* ---- stage 1: preprocessing (pipeline_demo/ocr_testing/preprocessing/image_preprocessor.py) ------------------------ */
typedef struct {
    float scale;            /* resize(scale_factor)      :125-132  (1.5)                       */
    float sigma;            /* denoise(strength)         :32-37    (3; legacy 5)               */
    float contrast;         /* increase_contrast(factor) :70-84    (1.9; legacy 1.3)           */
    float brightness;       /* increase_brightness       :86-100   (1.2; <=0 skips the step)   */
    float clahe_clip;       /* clahe(clip_limit)         :48-56    (2.5; legacy 2.0)           */
    int32_t sharpen_percent;/* sharpen(amount) -> int(amount*100)  :102-115 (30; legacy 20)    */
    int32_t resize_mode;    /* 0 = T1: OpenCV generic fixed-point INTER_CUBIC (bit-exact, cv2.ipp off)
                               1 = T2: real-arithmetic FP32 cubic (what IPP-enabled x86 wheels approximate) */
} bbocr_pp_params;

/* preprocess_for_book_cover :147-160 minus file I/O: BGR u8 -> gray, x1.5 cubic, 3x3 Gaussian, contrast, brightness,
 * CLAHE(8x8), unsharp.  out must hold int(H*scale)*int(W*scale) bytes.  *_on_device select device pointers. */
int bbocr_preprocess_u8(bbocr_handle* h, const uint8_t* bgr, int H, int W, int stride_bytes, int in_on_device,
                        const bbocr_pp_params* p, uint8_t* out, int out_on_device, int* outH, int* outW);
/* The same chain over n same-size photos (BASELINE config 3): bgr[i] / out[i] are n pointers (all host or all device);
 * photos are spread over the handle's streams so their launches overlap, one synchronisation per stream. */
int bbocr_preprocess_batch_u8(bbocr_handle* h, int n, const uint8_t* const* bgr, int H, int W, int stride_bytes,
                              int in_on_device, const bbocr_pp_params* p, uint8_t* const* out, int out_on_device, int* outH,
                              int* outW);
/* Number of kernels the chain launches per image. */
int bbocr_preprocess_launches_per_image(void);

/* single steps, host buffers (parity-test surface; each mirrors one ImagePreprocessor method) */
int bbocr_pp_gray(bbocr_handle* h, const uint8_t* bgr, int H, int W, uint8_t* out);                       /* :25-30  */
int bbocr_pp_resize_cubic(bbocr_handle* h, const uint8_t* src, int H, int W, int dstH, int dstW, int mode,
                          uint8_t* out);                                                                   /* :125-132 */
int bbocr_pp_gaussian3(bbocr_handle* h, const uint8_t* src, int H, int W, float sigma, uint8_t* out);      /* :32-37  */
int bbocr_pp_contrast(bbocr_handle* h, const uint8_t* src, int H, int W, float factor, uint8_t* out);      /* :70-84  */
int bbocr_pp_brightness(bbocr_handle* h, const uint8_t* src, int H, int W, float factor, uint8_t* out);    /* :86-100 */
int bbocr_pp_clahe(bbocr_handle* h, const uint8_t* src, int H, int W, float clip, uint8_t* out);           /* :48-56  */
int bbocr_pp_equalize_hist(bbocr_handle* h, const uint8_t* src, int H, int W, uint8_t* out);             /* :39-46  */
int bbocr_pp_unsharp(bbocr_handle* h, const uint8_t* src, int H, int W, int percent, int threshold,
                     uint8_t* out);                                                                        /* :102-115 */
/* gentle_threshold :58-68 and enhanced_extractor.py:258-259.  method 0 = MEAN_C, 1 = GAUSSIAN_C; inv = BINARY_INV */
int bbocr_pp_adaptive_threshold(bbocr_handle* h, const uint8_t* src, int H, int W, int method, int inv, int block,
                                float delta, uint8_t* out);
/* Deskew does not exist in the reference; defined here as: Otsu-free foreground = adaptive
 * Gaussian 31/5 INV, angle = projection-profile variance peak over [-max_deg, max_deg] in 0.1 deg steps, then a
 * bilinear replicate-border rotation about the centre.  angle_out (degrees, may be NULL). */
int bbocr_pp_deskew(bbocr_handle* h, const uint8_t* src, int H, int W, float max_deg, uint8_t* out, float* angle_out);

/* BASELINE.json config[2] as one device-resident chain ("gray, CLAHE, adaptive threshold, deskew" on phone photos), composed of
 * the steps above in the order a scan pipeline applies them: BGR -> gray (:25-30) -> CLAHE(clahe_clip) (:48-56) -> deskew
 * (max_deg; this repository's definition, see bbocr_pp_deskew) -> gentle_threshold(block, delta) (:58-68).  out: HxW u8 in
 * {0, 255}; angle_out (degrees, may be NULL). */
int bbocr_preprocess_scan_u8(bbocr_handle* h, const uint8_t* bgr, int H, int W, int stride_bytes, int in_on_device,
                             float clahe_clip, int block, float delta, float max_deg, uint8_t* out, int out_on_device,
                             float* angle_out);

/* ---- stage 2: detector (easyocr/detection.py, craft.py, imgproc.py) -------------------------------------------------- */
/* test_net up to the score maps: resize_aspect_ratio(canvas_size, INTER_LINEAR, mag_ratio) -> normalizeMeanVariance
 * -> CRAFT.forward.  img is HxWx3 u8 in the channel order EasyOCR feeds the net.  score_text/score_link: host float
 * buffers of capacity >= (ceil32(H')/2)*(ceil32(W')/2) (query sizes first with NULL buffers). */
int bbocr_craft_forward(bbocr_handle* h, const uint8_t* img, int H, int W, int on_device, int canvas_size,
                        double mag_ratio, float* score_text, float* score_link, int* mapH, int* mapW, double* ratio);

/* craft_utils.getDetBoxes_core(textmap, linkmap, text_threshold, link_threshold, low_text) on given maps (host
 * float, mapH x mapW).  boxes: capacity cap*8 floats, 4 (x,y) corners per box in upstream order. */
int bbocr_det_boxes(bbocr_handle* h, const float* textmap, const float* linkmap, int mapH, int mapW,
                    double text_threshold, double link_threshold, double low_text, float* boxes, int cap, int* n);

/* cv2.minAreaRect + cv2.boxPoints on integer points (x,y pairs), restated; pure host, no CUDA call. out: 8 floats */
int bbocr_min_area_box(const int32_t* xy, int npoints, float* out8);

typedef struct {            /* doubles: upstream compares Python floats */
    double slope_ths, ycenter_ths, height_ths, width_ths, add_margin;  /* readtext defaults .1 .5 .5 .5 .1 */
    int32_t min_size;                                                  /* 20 */
} bbocr_group_params;

/* craft_utils.adjustResultCoordinates + detection.get_textbox (scale by 2/ratio, truncate to int32) followed by
 * utils.group_text_box and Reader.detect's min_size filter.  Pure host.  hlist: cap*4 int32 [xmin,xmax,ymin,ymax];
 * flist: cap*8 doubles. */
int bbocr_group_boxes(const float* boxes, int n, double ratio, const bbocr_group_params* p, int32_t* hlist, int* nh,
                      double* flist, int* nf, int cap);

/* ---- stage 3-5: recogniser (easyocr/utils.py, recognition.py, model/vgg_model.py) ------------------------------------ */
/* utils.get_image_list for ONE horizontal box [xmin,xmax,ymin,ymax] on a gray page: clamp, slice,
 * compute_ratio_and_resize (cv2.resize INTER_LINEAR u8).  out capacity >= 64*max(outW) ; returns crop size and the
 * model width max_width = ceil(ratio)*64.  Returns outW = 0 when upstream skips the box. */
int bbocr_crop_horizontal(bbocr_handle* h, const uint8_t* gray, int H, int W, const int32_t box[4], uint8_t* out,
                          int cap, int* outH, int* outW, int* model_w);
/* utils.four_point_transform + compute_ratio_and_resize for ONE free-form box (4 (x,y) doubles). */
int bbocr_crop_free(bbocr_handle* h, const uint8_t* gray, int H, int W, const double quad[8], uint8_t* out, int cap,
                    int* outH, int* outW, int* model_w);
/* AlignCollate+NormalizePAD (x/255-0.5)/0.5 with replicate right pad, then vgg_model.Model.forward.
 * x: N x 64 x Wm normalised float (host); logits: N x (Wm/4-1) x 97 float (host). */
int bbocr_crnn_forward(bbocr_handle* h, const float* x, int N, int Wm, float* logits);
/* recognition.recognizer_predict greedy branch + CTCLabelConverter.decode_greedy + custom_mean.
 * logits: N x T x C host floats; ignore: C bytes (1 = masked class) or NULL.  text_idx: N*T int32 (collapsed class
 * indices, 1-based into the alphabet), text_len: N, conf: N doubles. */
int bbocr_ctc_decode(bbocr_handle* h, const float* logits, int N, int T, int C, const uint8_t* ignore,
                     int32_t* text_idx, int32_t* text_len, double* conf);

/* ---- whole stage ------------------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t* color;   /* H x W x 3 u8, channel order as EasyOCR's reformat_input hands `img` to the detector */
    const uint8_t* gray;    /* H x W u8 (`img_cv_grey`); NULL => (c0*3735 + c1*19235 + c2*9798 + 16384) >> 15 on device */
    int32_t H, W;
    int32_t on_device;      /* 0: host pointers (copied inside the call), 1: device pointers */
} bbocr_image;

typedef struct {            /* the subset of readtext kwargs that changes arithmetic; defaults in comments */
    int32_t min_size;       /* 20 */
    int32_t canvas_size;    /* 2560 */
    double contrast_ths;    /* 0.1 */
    double adjust_contrast; /* 0.5 */
    double text_threshold;  /* 0.7 */
    double low_text;        /* 0.4 */
    double link_threshold;  /* 0.4 */
    double mag_ratio;       /* 1.0 */
    double slope_ths, ycenter_ths, height_ths, width_ths, add_margin;  /* .1 .5 .5 .5 .1 */
    const uint8_t* ignore;  /* num_class bytes or NULL (allowlist / blocklist mask; 1 = masked) */
    /* SURVEY.md §8f-3 -- the remaining readtext options (easyocr/easyocr.py::Reader.recognize, easyocr/utils.py) */
    int32_t decoder;        /* 0 'greedy' | 1 'beamsearch' | 2 'wordbeamsearch' (dictionary: bbocr_set_dictionary) */
    int32_t beam_width;     /* 5 (beamWidth) */
    int32_t batch_mode;     /* 0: batch_size == 1 branch (every box has its own max_width; horizontal then free boxes).
                               1: upstream's batched branch (batch_size > 1 on a GPU Reader, or rotation_info): ONE max_width
                               for the page's boxes, free boxes first, results stably ordered by the first corner's y */
    int32_t n_rotations;    /* rotation_info: number of extra orientations (0..3); implies batch_mode */
    int32_t rotation[3];    /* each 90, 180 or 270 (the eligible values upstream documents) */
    int32_t space_idx;      /* 43: class index of ' ' in english_g2 (CTCLabelConverter.dict[' ']); wordbeamsearch cuts there */
} bbocr_params;

void bbocr_default_params(bbocr_params* p);

typedef struct {
    int32_t n;              /* number of (box, text, confidence) tuples, in EasyOCR's order */
    double* box;            /* n x 8 : 4 (x,y) corners; integral values for horizontal boxes */
    uint8_t* is_free;       /* n     : 1 = free-form box (floats), 0 = horizontal (ints) */
    int32_t* text_off;      /* n + 1 : offsets into text_idx */
    int32_t* text_idx;      /* class indices (1..96) of the decoded characters */
    double* conf;           /* n */
    int32_t n_crops;        /* recogniser invocations, second passes included (work accounting) */
    int32_t n_components;   /* connected components seen by getDetBoxes */
} bbocr_results;

/* Reader.readtext for ndarray input (easyocr/easyocr.py; call site enhanced_extractor.py:520). */
int bbocr_readtext(bbocr_handle* h, const bbocr_image* img, const bbocr_params* p, bbocr_results** out);
/* readtext over n independent pages, pipelined over the handle's streams; out[i] filled per page. */
int bbocr_readtext_batch(bbocr_handle* h, int n, const bbocr_image* imgs, const bbocr_params* p, bbocr_results** out);
/* Reader.recognize (easyocr/easyocr.py): crops (utils.get_image_list), AlignCollate, CRNN, greedy CTC and the
 * contrast-retry pass for given boxes of one gray page.  hlist: nh x [x_min, x_max, y_min, y_max]; flist: nf x 4 (x,y)
 * corners.  Results in upstream order: horizontal boxes, then free boxes. */
int bbocr_recognize(bbocr_handle* h, const uint8_t* gray, int H, int W, int on_device, const int32_t* hlist, int nh,
                    const double* flist, int nf, const bbocr_params* p, bbocr_results** out);
void bbocr_results_free(bbocr_results* r);
/* Dictionary of decoder 'wordbeamsearch' (upstream: easyocr/dict/en.txt read by CTCLabelConverter.__init__): n words given as
 * class-index sequences idx[off[i] .. off[i+1]).  n = 0 clears it (every word is then plain beam search). */
int bbocr_set_dictionary(bbocr_handle* h, const int32_t* idx, const int32_t* off, int n);
/* CTCLabelConverter.decode_beamsearch / decode_wordbeamsearch of ONE crop on the host (upstream runs them in Python on the
 * CPU as well): probs = T x C float32 as recognizer_predict hands them over; decoder 1 | 2; space_idx = class index of ' '.
 * Writes at most cap class indices to text_out, *len = the text's length.  Needs no device (parity-test surface). */
int bbocr_ctc_beam_decode(const float* probs, int T, int C, int decoder, int beam_width, int space_idx, const int32_t* dict_idx,
                          const int32_t* dict_off, int n_dict, int32_t* text_out, int cap, int* len);

/* ---- extractor glue ---------------------------------------------------------------------------- */
/* The OCR-input cap of extract_text_with_ocr (pipeline_demo/extractor/enhanced_extractor.py:486-512): PIL
 * Image.thumbnail((max_dim, max_dim)) = aspect-preserving BICUBIC down-scale (Pillow's fixed-point two-pass convolution),
 * for one gray u8 plane, without the lossy JPEG round trip the reference adds.  out == NULL: size query only.  Shrinks
 * by 4x or more (the reference's 5712x4284 photos: x1.5 -> 8568x6426 -> 1600) take Pillow's Image.reduce() box pre-pass
 * followed by the bicubic pass over the fractional box, bit-exact like the plain regime. */
int bbocr_thumbnail_u8(bbocr_handle* h, const uint8_t* src, int H, int W, int in_on_device, int max_dim, uint8_t* out,
                       int out_on_device, int* outH, int* outW);

/* ---- image decode in front of the path ------------------------------------------------------------- */
/* cv2.imread / cv2.imdecode of a baseline JPEG (pipeline_demo/ocr_testing/preprocessing/image_preprocessor.py:18;
 * easyocr/utils.py::reformat_input reads the file as IMREAD_GRAYSCALE and as colour), bit-exact with OpenCV's libjpeg-turbo
 * (islow IDCT, fancy up-sampling, fixed-point YCbCr->BGR) including the EXIF orientation step.  Entropy decoding runs on the
 * device with one thread per restart interval (phone cameras write one per MCU row); files without restart markers are
 * Huffman-decoded by host threads and take the device from the IDCT on.  Progressive / arithmetic / 12-bit / CMYK /
 * multi-scan files return BBOCR_E_UNSUPPORTED (callers keep their host decoder for those). */
int bbocr_jpeg_info(const uint8_t* data, size_t n, int* H, int* W, int* channels, int* orientation);   /* oriented H, W; no device */
/* Host only (parity-test surface): the quantised DCT coefficients (jdhuff.c output) of every block -- component after
 * component, blocks in raster order padded to whole MCUs, 64 int16 each in natural order -- through the product's parser and
 * entropy decoder.  out == NULL or cap_blocks too small: only *n_blocks is set. */
int bbocr_jpeg_coefficients(const uint8_t* data, size_t n, int16_t* out, int64_t cap_blocks, int64_t* n_blocks);
/* out_bgr: H x W x 3 (cv2.IMREAD_COLOR) and / or out_gray: H x W (cv2.IMREAD_GRAYSCALE = the luma plane); either may be NULL.
 * out_on_device = 1: device pointers, the call returns once the work is enqueued and complete on return of
 * bbocr_jpeg_decode_batch / any later synchronous call of the handle. */
int bbocr_jpeg_decode(bbocr_handle* h, const uint8_t* data, size_t n, int ignore_orientation, uint8_t* out_bgr, uint8_t* out_gray,
                      int out_on_device, int* H, int* W);
/* n files -> n device images, pipelined over the handle's streams; returns when all are complete.  out_bgr[i] / out_gray[i]
 * are device pointers (NULL entries or NULL arrays skip that output). */
int bbocr_jpeg_decode_batch(bbocr_handle* h, int n, const uint8_t* const* data, const size_t* sizes, int ignore_orientation,
                            uint8_t* const* out_bgr, uint8_t* const* out_gray);

/* ---- page crops in front of the OCR stage ---------------------------------------------------------- */
/* _auto_crop_text_region (pipeline_demo/extractor/enhanced_extractor.py:239-372) up to the slice it writes: the crop
 * rectangle rect = (x0, y0, x1, y1) of a page for the given margin; *found = 0 is the reference's `return None`
 * ("no crop").  channels = 3: BGR u8 (what cv2.imread gives the reference); channels = 1: a gray plane, e.g. the
 * preprocessing output (cv2.imread of a gray PNG yields three equal channels, whose BGR2GRAY is the plane itself).  The composite text mask, the rectangle morphology, the external-contour bounding boxes all run on the
 * device, bit-exact against cv2; the caller slices img[y0:y1, x0:x1] itself (no PNG round trip).
 * Optional parity outputs (host pointers, each may be NULL): mask_out / merged_out HxW u8 (0/255) = the reference's `mask`
 * and `merged`; boxes_out = up to boxes_cap rows (x, y, w, h) = boundingRect of every RETR_EXTERNAL contour of `merged`
 * sorted by (y, x, w, h), *nboxes = their total number; otsu_out[2] = the two Otsu thresholds. */
int bbocr_autocrop_rect(bbocr_handle* h, const uint8_t* img, int H, int W, int channels, int stride_bytes, int in_on_device, int margin,
                        int32_t rect[4], int* found, uint8_t* mask_out, uint8_t* merged_out, int32_t* boxes_out,
                        int boxes_cap, int* nboxes, int32_t* otsu_out);
/* Building blocks of the above on host buffers (parity-test surface): cv2.findContours(RETR_EXTERNAL) + boundingRect of a
 * binary u8 image (non-zero = foreground; rows (x, y, w, h) sorted by (y, x, w, h)), and cv2.dilate / cv2.erode with a
 * kw x kh rectangle (odd sizes, kw <= 63, centre anchor, default border). */
int bbocr_external_boxes(bbocr_handle* h, const uint8_t* binary, int H, int W, int32_t* boxes_out, int boxes_cap, int* nboxes);
int bbocr_rect_morph(bbocr_handle* h, const uint8_t* binary, int H, int W, int kw, int kh, int erode, uint8_t* out);

/* ---- instrumentation -------------------------------------------------------------------------------------------- */
/* Kernels launched by this handle since the last reset (the bench's gpu_launches claim). */
int64_t bbocr_launch_count(const bbocr_handle* h);
void bbocr_reset_launch_count(bbocr_handle* h);
/* Device time (ms, CUDA events on the launching stream) and launch count of the dominant kernel family
 * (implicit-GEMM convolution) since the last reset; algorithmic FLOPs of those launches. */
int bbocr_conv_stats(bbocr_handle* h, double* ms, int64_t* launches, double* flops);
int bbocr_enable_conv_timing(bbocr_handle* h, int on);

#ifdef __cplusplus
}
#endif
#endif /* BBOCR_H */
