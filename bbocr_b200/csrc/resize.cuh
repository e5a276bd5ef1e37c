// resize.cuh -- cv2.resize(u8, INTER_LINEAR) arithmetic shared by the canvas and crop kernels
#pragma once
#include "engine.h"

namespace bbocr {

// ------------------------------------------------------------------------------------------------------------------
// cv2.resize(u8, INTER_LINEAR): 11-bit fixed-point bilinear (used for the canvas and the crops)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lin_coeff_x(int d, double scale, int n, int* s, int* a0, int* a1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int i = (int)floorf(f);
    f -= (float)i;
    if (i < 0) { i = 0; f = 0.f; }
    if (i >= n - 1) { i = n - 1; f = 0.f; }
    *s = i;
    *a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    *a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}
__device__ __forceinline__ void lin_coeff_y(int d, double scale, int* s, int* b0, int* b1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int i = (int)floorf(f);
    f -= (float)i;
    *s = i;
    *b0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    *b1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

__device__ __forceinline__ uint8_t bilinear_u8_px(const uint8_t* __restrict__ src, int sH, int sW, int sstride, int C,
                                                  int c, int dx, int dy, double scale_x, double scale_y) {
    int sx, a0, a1, sy, b0, b1;
    lin_coeff_x(dx, scale_x, sW, &sx, &a0, &a1);
    lin_coeff_y(dy, scale_y, &sy, &b0, &b1);
    int x1 = min(sx + 1, sW - 1);
    int y0 = min(max(sy, 0), sH - 1), y1 = min(max(sy + 1, 0), sH - 1);
    const uint8_t* r0 = src + (int64_t)y0 * sstride;
    const uint8_t* r1 = src + (int64_t)y1 * sstride;
    int H0 = r0[sx * C + c] * a0 + r0[x1 * C + c] * a1;
    int H1 = r1[sx * C + c] * a0 + r1[x1 * C + c] * a1;
    int v = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2;
    return (uint8_t)min(max(v, 0), 255);
}

}  // namespace bbocr
