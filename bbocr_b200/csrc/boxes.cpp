// boxes.cpp -- host-side geometry of the detector tail (O(boxes) work; no CUDA calls in this file):
//   * min_area_box      : cv2.minAreaRect + cv2.boxPoints on integer points, restated in float32 exactly
//                         (OpenCV imgproc convhull.cpp Sklansky hull + rotcalipers.cpp rotatingCalipers)
//   * boxes_from_components : the per-label part of craft_utils.getDetBoxes_core after the per-pixel work
//                         (square dilation of the row extents, min-area box, "diamond" fix-up, clockwise start)
//   * group_boxes       : adjustResultCoordinates + get_textbox int32 truncation + utils.group_text_box + min_size filter
// SURVEY.md §8a B5-B8.  x86-64 host code: no FMA contraction, float math as written.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "engine.h"
#include "geom.cuh"

namespace bbocr {

namespace {

using geom::P2i;
using geom::P2f;

// sort order of cv::convexHull's pointer array: (x, y, original position); sp = the points in that order
void sort_points(const P2i* pts, int n, std::vector<int>& sorted, std::vector<P2i>& sp) {
    sorted.resize(n);
    for (int i = 0; i < n; ++i) sorted[i] = i;
    std::sort(sorted.begin(), sorted.end(), [&](int a, int b) {
        if (pts[a].x != pts[b].x) return pts[a].x < pts[b].x;
        if (pts[a].y != pts[b].y) return pts[a].y < pts[b].y;
        return a < b;
    });
    sp.resize(std::max(n, 1));
    for (int i = 0; i < n; ++i) sp[i] = pts[sorted[i]];
}

}  // namespace

void debug_convex_hull(const int32_t* xy, int n, int clockwise, std::vector<int>& hull) {
    std::vector<P2i> pts(n);
    for (int i = 0; i < n; ++i) { pts[i].x = xy[2 * i]; pts[i].y = xy[2 * i + 1]; }
    std::vector<int> sorted, stack(n + 2), hullbuf(std::max(n, 1));
    std::vector<P2i> sp;
    sort_points(pts.data(), n, sorted, sp);
    const int hn = geom::convex_hull(sp.data(), sorted.data(), n, stack.data(), hullbuf.data(), clockwise != 0);
    hull.assign(hullbuf.begin(), hullbuf.begin() + hn);
}

// cv2.boxPoints(cv2.minAreaRect(points)) for int32 (x,y) points (the arithmetic lives in geom.cuh, shared with the device)
void min_area_box(const int32_t* xy, int n, float* out8) {
    std::vector<P2i> pts(n);
    for (int i = 0; i < n; ++i) { pts[i].x = xy[2 * i]; pts[i].y = xy[2 * i + 1]; }
    std::vector<int> sorted, stack(n + 2), hullbuf(std::max(n, 1));
    std::vector<P2i> sp;
    sort_points(pts.data(), n, sorted, sp);
    // minAreaRect: convexHull(points, hull, clockwise=false, returnPoints=true)
    const int hn = geom::convex_hull(sp.data(), sorted.data(), n, stack.data(), hullbuf.data(), false);
    std::vector<P2f> hp(std::max(hn, 1)), vect(std::max(hn, 1));
    std::vector<float> inv_len(std::max(hn, 1));
    geom::min_area_box_from_hull(pts.data(), hullbuf.data(), hn, hp.data(), inv_len.data(), vect.data(), out8);
}

// getDetBoxes_core per kept label, from the per-row extents of (label minus link-only pixels)
void boxes_from_components(const DetComponents& dc, int mh, int mw, std::vector<float>& boxes) {
    boxes.clear();
    const int nk = (int)dc.row_off.size();
    std::vector<int32_t> pts;
    std::vector<int> dmin, dmax;
    for (int c = 0; c < nk; ++c) {
        const int x = dc.comp_x[c], y = dc.comp_y[c], w = dc.comp_w[c], h = dc.comp_h[c], size = dc.comp_area[c];
        int niter = (int)(std::sqrt((double)((int64_t)size * std::min(w, h)) / (double)((int64_t)w * h)) * 2);
        int sx = x - niter, ex = x + w + niter + 1, sy = y - niter, ey = y + h + niter + 1;
        if (sx < 0) sx = 0;
        if (sy < 0) sy = 0;
        if (ex >= mw) ex = mw;
        if (ey >= mh) ey = mh;
        const int ks = 1 + niter, anchor = ks / 2;
        const int lo = ks - 1 - anchor, hi = anchor;          // a source pixel q covers [q - lo, q + hi] on each axis
        const int* rmin = dc.row_min.data() + dc.row_off[c];
        const int* rmax = dc.row_max.data() + dc.row_off[c];
        dmin.assign(ey - sy, INT_MAX);
        dmax.assign(ey - sy, -1);
        for (int r = 0; r < h; ++r) {
            if (rmin[r] > rmax[r]) continue;                   // row holds only link-only pixels
            int x0 = std::max(rmin[r] - lo, sx), x1 = std::min(rmax[r] + hi, ex - 1);
            int y0 = std::max(y + r - lo, sy), y1 = std::min(y + r + hi, ey - 1);
            for (int yy = y0; yy <= y1; ++yy) {
                dmin[yy - sy] = std::min(dmin[yy - sy], x0);
                dmax[yy - sy] = std::max(dmax[yy - sy], x1);
            }
        }
        pts.clear();
        int l = INT_MAX, r_ = -1, t = INT_MAX, b = -1;
        for (int yy = sy; yy < ey; ++yy) {
            int a = dmin[yy - sy], e = dmax[yy - sy];
            if (a > e) continue;
            pts.push_back(a); pts.push_back(yy);
            if (e != a) { pts.push_back(e); pts.push_back(yy); }
            l = std::min(l, a); r_ = std::max(r_, e);
            t = std::min(t, yy); b = std::max(b, yy);
        }
        float box[8], rolled[8];
        if (pts.empty()) {
            // cv2.minAreaRect on an empty point set returns a zero rect; upstream would then produce a zero box
            for (float& v : box) v = 0.f;
            for (float& v : rolled) v = 0.f;
        } else {
            min_area_box(pts.data(), (int)pts.size() / 2, box);
            geom::finish_det_box(box, l, t, r_, b, rolled);          // "diamond" rule + clockwise start
        }
        boxes.insert(boxes.end(), rolled, rolled + 8);
    }
}

// ---- adjustResultCoordinates + get_textbox + group_text_box + min_size filter -----------------------------------------
namespace {
struct HBox { int64_t xmin, xmax, ymin, ymax; double yc; int64_t h; };
double mean_d(const std::vector<double>& v) {     // np.mean of a Python list of float64/int -> pairwise-free for len < 8
    double s = 0;
    for (double x : v) s += x;
    return s / (double)v.size();
}
}  // namespace

void group_boxes(const float* boxes, int n, double ratio, const bbocr_group_params& p, std::vector<int32_t>& hlist,
                 std::vector<double>& flist) {
    hlist.clear();
    flist.clear();
    // adjustResultCoordinates: box(float32) *= (ratio_w*2, ratio_h*2) computed in float64, stored back to float32;
    // get_textbox: astype(int32) truncation
    const double rw = (1.0 / ratio) * 2, rh = (1.0 / ratio) * 2;
    std::vector<HBox> horiz;
    const double slope_ths = p.slope_ths, ycenter_ths = p.ycenter_ths, height_ths = p.height_ths,
                 width_ths = p.width_ths, add_margin = p.add_margin;
    for (int i = 0; i < n; ++i) {
        int32_t q[8];
        for (int j = 0; j < 8; ++j) {
            float v = (float)((double)boxes[i * 8 + j] * ((j & 1) ? rh : rw));
            q[j] = (int32_t)v;
        }
        double slope_up = (double)(q[3] - q[1]) / (double)std::max(10, q[2] - q[0]);
        double slope_down = (double)(q[5] - q[7]) / (double)std::max(10, q[4] - q[6]);
        if (std::max(std::fabs(slope_up), std::fabs(slope_down)) < slope_ths) {
            HBox b;
            b.xmax = std::max(std::max(q[0], q[2]), std::max(q[4], q[6]));
            b.xmin = std::min(std::min(q[0], q[2]), std::min(q[4], q[6]));
            b.ymax = std::max(std::max(q[1], q[3]), std::max(q[5], q[7]));
            b.ymin = std::min(std::min(q[1], q[3]), std::min(q[5], q[7]));
            b.yc = 0.5 * (double)(b.ymin + b.ymax);
            b.h = b.ymax - b.ymin;
            horiz.push_back(b);
        } else {
            double hx = q[6] - q[0], hy = q[7] - q[1], wx = q[2] - q[0], wy = q[3] - q[1];
            double height = std::sqrt(hx * hx + hy * hy), width = std::sqrt(wx * wx + wy * wy);
            int margin = (int)(1.44 * add_margin * std::min(width, height));
            double theta13 = std::fabs(std::atan((double)(q[1] - q[5]) / (double)std::max(10, q[0] - q[4])));
            double theta24 = std::fabs(std::atan((double)(q[3] - q[7]) / (double)std::max(10, q[2] - q[6])));
            double f[8];
            f[0] = q[0] - std::cos(theta13) * margin; f[1] = q[1] - std::sin(theta13) * margin;
            f[2] = q[2] + std::cos(theta24) * margin; f[3] = q[3] - std::sin(theta24) * margin;
            f[4] = q[4] + std::cos(theta13) * margin; f[5] = q[5] + std::sin(theta13) * margin;
            f[6] = q[6] - std::cos(theta24) * margin; f[7] = q[7] + std::sin(theta24) * margin;
            for (double v : f) flist.push_back(v);
        }
    }
    std::stable_sort(horiz.begin(), horiz.end(), [](const HBox& a, const HBox& b) { return a.yc < b.yc; });
    std::vector<std::vector<HBox>> combined;
    std::vector<HBox> new_box;
    std::vector<double> b_height, b_ycenter;
    for (const HBox& poly : horiz) {
        if (new_box.empty()) {
            b_height = {(double)poly.h};
            b_ycenter = {poly.yc};
            new_box.push_back(poly);
        } else if (std::fabs(mean_d(b_ycenter) - poly.yc) < ycenter_ths * mean_d(b_height)) {
            b_height.push_back((double)poly.h);
            b_ycenter.push_back(poly.yc);
            new_box.push_back(poly);
        } else {
            b_height = {(double)poly.h};
            b_ycenter = {poly.yc};
            combined.push_back(new_box);
            new_box = {poly};
        }
    }
    combined.push_back(new_box);
    std::vector<std::array<int64_t, 4>> merged;
    for (auto& boxes_l : combined) {
        if (boxes_l.size() == 1) {
            const HBox& box = boxes_l[0];
            int64_t margin = (int64_t)(add_margin * (double)std::min(box.xmax - box.xmin, box.h));
            merged.push_back({box.xmin - margin, box.xmax + margin, box.ymin - margin, box.ymax + margin});
        } else {
            std::stable_sort(boxes_l.begin(), boxes_l.end(), [](const HBox& a, const HBox& b) { return a.xmin < b.xmin; });
            std::vector<std::vector<HBox>> merged_box;
            std::vector<HBox> nb;
            std::vector<double> bh;
            int64_t x_max = 0;
            for (const HBox& box : boxes_l) {
                if (nb.empty()) {
                    bh = {(double)box.h};
                    x_max = box.xmax;
                    nb.push_back(box);
                } else if (std::fabs(mean_d(bh) - (double)box.h) < height_ths * mean_d(bh) &&
                           (double)(box.xmin - x_max) < width_ths * (double)(box.ymax - box.ymin)) {
                    bh.push_back((double)box.h);
                    x_max = box.xmax;
                    nb.push_back(box);
                } else {
                    bh = {(double)box.h};
                    x_max = box.xmax;
                    merged_box.push_back(nb);
                    nb = {box};
                }
            }
            if (!nb.empty()) merged_box.push_back(nb);
            for (auto& mbox : merged_box) {
                if (mbox.size() != 1) {
                    int64_t xmin = mbox[0].xmin, xmax = mbox[0].xmax, ymin = mbox[0].ymin, ymax = mbox[0].ymax;
                    for (const HBox& b : mbox) {
                        xmin = std::min(xmin, b.xmin); xmax = std::max(xmax, b.xmax);
                        ymin = std::min(ymin, b.ymin); ymax = std::max(ymax, b.ymax);
                    }
                    int64_t margin = (int64_t)(add_margin * (double)std::min(xmax - xmin, ymax - ymin));
                    merged.push_back({xmin - margin, xmax + margin, ymin - margin, ymax + margin});
                } else {
                    const HBox& box = mbox[0];
                    int64_t margin = (int64_t)(add_margin * (double)std::min(box.xmax - box.xmin, box.ymax - box.ymin));
                    merged.push_back({box.xmin - margin, box.xmax + margin, box.ymin - margin, box.ymax + margin});
                }
            }
        }
    }
    // Reader.detect: min_size filter
    for (auto& m : merged)
        if (!p.min_size || std::max(m[1] - m[0], m[3] - m[2]) > p.min_size)
            for (int j = 0; j < 4; ++j) hlist.push_back((int32_t)m[j]);
    if (p.min_size) {
        std::vector<double> kept;
        for (size_t i = 0; i + 8 <= flist.size(); i += 8) {
            double xs0 = flist[i], xs1 = flist[i], ys0 = flist[i + 1], ys1 = flist[i + 1];
            for (int j = 1; j < 4; ++j) {
                xs0 = std::min(xs0, flist[i + 2 * j]); xs1 = std::max(xs1, flist[i + 2 * j]);
                ys0 = std::min(ys0, flist[i + 2 * j + 1]); ys1 = std::max(ys1, flist[i + 2 * j + 1]);
            }
            if (std::max(xs1 - xs0, ys1 - ys0) > p.min_size) kept.insert(kept.end(), flist.begin() + i, flist.begin() + i + 8);
        }
        flist.swap(kept);
    }
}

// ---- utils.four_point_transform geometry: cv2.getPerspectiveTransform (LU solve, float64) and the 3x3 inverse that
// cv2.warpPerspective applies when WARP_INVERSE_MAP is not set ----------------------------------------------------------
namespace {
// OpenCV hal::LU64f (partial pivoting), m x m system with one right-hand side
bool lu_solve(double* A, int m, double* b) {
    const double eps = DBL_EPSILON * 100;
    for (int i = 0; i < m; ++i) {
        int k = i;
        for (int j = i + 1; j < m; ++j)
            if (std::abs(A[j * m + i]) > std::abs(A[k * m + i])) k = j;
        if (std::abs(A[k * m + i]) < eps) return false;
        if (k != i) {
            for (int j = i; j < m; ++j) std::swap(A[i * m + j], A[k * m + j]);
            std::swap(b[i], b[k]);
        }
        double d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; ++j) {
            double alpha = A[j * m + i] * d;
            for (int q = i + 1; q < m; ++q) A[j * m + q] += alpha * A[i * m + q];
            b[j] += alpha * b[i];
        }
    }
    for (int i = m - 1; i >= 0; --i) {
        double s = b[i];
        for (int q = i + 1; q < m; ++q) s -= A[i * m + q] * b[q];
        b[i] = s / A[i * m + i];
    }
    return true;
}
}  // namespace

// quad: 4 (x,y) doubles (tl,tr,br,bl) as group_text_box emits them.  Outputs the warped patch size and the inverse
// perspective matrix (dst pixel -> source coordinates) used by cv2.warpPerspective.
void free_box_transform(const double* quad, int* max_w, int* max_h, double* Minv) {
    float r[8];
    for (int i = 0; i < 8; ++i) r[i] = (float)quad[i];          // np.array(box, dtype="float32")
    const float *tl = r, *tr = r + 2, *br = r + 4, *bl = r + 6;
    auto dist = [](const float* a, const float* b) {
        float dx = a[0] - b[0], dy = a[1] - b[1];
        float s = dx * dx + dy * dy;                              // float32 arithmetic, as NumPy scalars
        return std::sqrt(s);
    };
    int widthA = (int)dist(br, bl), widthB = (int)dist(tr, tl);
    int mw = std::max(widthA, widthB);
    int heightA = (int)dist(tr, br), heightB = (int)dist(tl, bl);
    int mh = std::max(heightA, heightB);
    *max_w = mw;
    *max_h = mh;
    if (mw <= 0 || mh <= 0) return;
    float dst[8] = {0, 0, (float)(mw - 1), 0, (float)(mw - 1), (float)(mh - 1), 0, (float)(mh - 1)};
    double a[64], b[8];
    for (int i = 0; i < 4; ++i) {
        double sx = r[2 * i], sy = r[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
        double* r0 = a + i * 8;
        double* r1 = a + (i + 4) * 8;
        r0[0] = r1[3] = sx;
        r0[1] = r1[4] = sy;
        r0[2] = r1[5] = 1;
        r0[3] = r0[4] = r0[5] = r1[0] = r1[1] = r1[2] = 0;
        // cv::getPerspectiveTransform multiplies the Point2f coordinates in FLOAT and widens the product:
        // a[i][6] = -src[i].x*dst[i].x  (found by diffing against cv2: a double product moves M's low bits and, through a
        // tie in cvRound, 1 pixel in ~10^6)
        r0[6] = (double)(-r[2 * i] * dst[2 * i]);
        r0[7] = (double)(-r[2 * i + 1] * dst[2 * i]);
        r1[6] = (double)(-r[2 * i] * dst[2 * i + 1]);
        r1[7] = (double)(-r[2 * i + 1] * dst[2 * i + 1]);
        b[i] = dx;
        b[i + 4] = dy;
    }
    double M[9];
    if (!lu_solve(a, 8, b)) { for (double& v : b) v = 0; }
    for (int i = 0; i < 8; ++i) M[i] = b[i];
    M[8] = 1.0;
    // cv::invert 3x3 (DECOMP_LU closed form)
    double d = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
    if (d != 0.) {
        d = 1. / d;
        Minv[0] = (M[4] * M[8] - M[5] * M[7]) * d;
        Minv[1] = (M[2] * M[7] - M[1] * M[8]) * d;
        Minv[2] = (M[1] * M[5] - M[2] * M[4]) * d;
        Minv[3] = (M[5] * M[6] - M[3] * M[8]) * d;
        Minv[4] = (M[0] * M[8] - M[2] * M[6]) * d;
        Minv[5] = (M[2] * M[3] - M[0] * M[5]) * d;
        Minv[6] = (M[3] * M[7] - M[4] * M[6]) * d;
        Minv[7] = (M[1] * M[6] - M[0] * M[7]) * d;
        Minv[8] = (M[0] * M[4] - M[1] * M[3]) * d;
    } else {
        for (int i = 0; i < 9; ++i) Minv[i] = 0;
    }
}

}  // namespace bbocr
