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

namespace bbocr {

namespace {

struct Pt { int x, y; };
struct Pt2f { float x, y; };

inline int sgn(int64_t v) { return (v > 0) - (v < 0); }

// OpenCV convhull.cpp::Sklansky_<int,int64>
int sklansky(const std::vector<const Pt*>& array, int start, int end, int* stack, int nsign, int sign2) {
    int incr = end > start ? 1 : -1;
    int pprev = start, pcur = pprev + incr, pnext = pcur + incr;
    int stacksize = 3;
    if (start == end || (array[start]->x == array[end]->x && array[start]->y == array[end]->y)) {
        stack[0] = start;
        return 1;
    }
    stack[0] = pprev;
    stack[1] = pcur;
    stack[2] = pnext;
    end += incr;
    while (pnext != end) {
        int cury = array[pcur]->y;
        int nexty = array[pnext]->y;
        int by = nexty - cury;
        if (sgn(by) != nsign) {
            int ax = array[pcur]->x - array[pprev]->x;
            int bx = array[pnext]->x - array[pcur]->x;
            int ay = cury - array[pprev]->y;
            int64_t convexity = (int64_t)ay * bx - (int64_t)ax * by;
            if (sgn(convexity) == sign2 && (ax != 0 || ay != 0)) {
                pprev = pcur;
                pcur = pnext;
                pnext += incr;
                stack[stacksize] = pnext;
                stacksize++;
            } else {
                if (pprev == start) {
                    pcur = pnext;
                    stack[1] = pcur;
                    pnext += incr;
                    stack[2] = pnext;
                } else {
                    stack[stacksize - 2] = pnext;
                    pcur = pprev;
                    pprev = stack[stacksize - 4];
                    stacksize--;
                }
            }
        } else {
            pnext += incr;
            stack[stacksize - 1] = pnext;
        }
    }
    return --stacksize;
}

// cv::convexHull(points, hull, clockwise=false, returnPoints=true) for CV_32S points -> hull vertex indices
void convex_hull(const Pt* data0, int total, std::vector<int>& hull, bool clockwise = false) {
    hull.clear();
    if (total == 0) return;
    std::vector<const Pt*> pointer(total);
    std::vector<int> stackv(total + 2), hullbuf(total);
    int* stack = stackv.data();
    for (int i = 0; i < total; ++i) pointer[i] = &data0[i];
    std::sort(pointer.begin(), pointer.end(), [](const Pt* a, const Pt* b) {
        if (a->x != b->x) return a->x < b->x;
        if (a->y != b->y) return a->y < b->y;
        return a < b;
    });
    int miny_ind = 0, maxy_ind = 0, nout = 0;
    for (int i = 1; i < total; ++i) {
        int y = pointer[i]->y;
        if (pointer[miny_ind]->y > y) miny_ind = i;
        if (pointer[maxy_ind]->y < y) maxy_ind = i;
    }
    if (pointer[0]->x == pointer[total - 1]->x && pointer[0]->y == pointer[total - 1]->y) {
        hullbuf[nout++] = 0;
    } else {
        int* tl_stack = stack;
        int tl_count = sklansky(pointer, 0, maxy_ind, tl_stack, -1, 1);
        int* tr_stack = stack + tl_count;
        int tr_count = sklansky(pointer, total - 1, maxy_ind, tr_stack, -1, -1);
        if (!clockwise) {
            std::swap(tl_stack, tr_stack);
            std::swap(tl_count, tr_count);
        }
        for (int i = 0; i < tl_count - 1; ++i) hullbuf[nout++] = (int)(pointer[tl_stack[i]] - data0);
        for (int i = tr_count - 1; i > 0; --i) hullbuf[nout++] = (int)(pointer[tr_stack[i]] - data0);
        int stop_idx = tr_count > 2 ? tr_stack[1] : tl_count > 2 ? tl_stack[tl_count - 2] : -1;

        int* bl_stack = stack;
        int bl_count = sklansky(pointer, 0, miny_ind, bl_stack, 1, -1);
        int* br_stack = stack + bl_count;
        int br_count = sklansky(pointer, total - 1, miny_ind, br_stack, 1, 1);
        if (clockwise) {
            std::swap(bl_stack, br_stack);
            std::swap(bl_count, br_count);
        }
        if (stop_idx >= 0) {
            int check_idx = bl_count > 2 ? bl_stack[1] : bl_count + br_count > 2 ? br_stack[2 - bl_count] : -1;
            if (check_idx == stop_idx ||
                (check_idx >= 0 && pointer[check_idx]->x == pointer[stop_idx]->x && pointer[check_idx]->y == pointer[stop_idx]->y)) {
                bl_count = std::min(bl_count, 2);
                br_count = std::min(br_count, 2);
            }
        }
        for (int i = 0; i < bl_count - 1; ++i) hullbuf[nout++] = (int)(pointer[bl_stack[i]] - data0);
        for (int i = br_count - 1; i > 0; --i) hullbuf[nout++] = (int)(pointer[br_stack[i]] - data0);

        if (nout >= 3) {
            int min_idx = 0, max_idx = 0, lt = 0, i;
            for (i = 1; i < nout; ++i) {
                int idx = hullbuf[i];
                lt += hullbuf[i - 1] < idx;
                if (lt > 1 && lt <= i - 2) break;
                if (idx < hullbuf[min_idx]) min_idx = i;
                if (idx > hullbuf[max_idx]) max_idx = i;
            }
            int mmdist = std::abs(max_idx - min_idx);
            if ((mmdist == 1 || mmdist == nout - 1) && (lt <= 1 || lt >= nout - 2)) {
                int ascending = (max_idx + 1) % nout == min_idx;
                int i0 = ascending ? min_idx : max_idx, j = i0;
                if (i0 > 0) {
                    for (i = 0; i < nout; ++i) {
                        int curr_idx = stack[i] = hullbuf[j];
                        int next_j = j + 1 < nout ? j + 1 : 0;
                        int next_idx = hullbuf[next_j];
                        if (i < nout - 1 && (ascending != (curr_idx < next_idx))) break;
                        j = next_j;
                    }
                    if (i == nout) memcpy(hullbuf.data(), stack, nout * sizeof(int));
                }
            }
        }
    }
    hull.assign(hullbuf.begin(), hullbuf.begin() + nout);
}

inline void rot90cw(const Pt2f& in, Pt2f& out) { out.x = in.y; out.y = -in.x; }
inline void rot90ccw(const Pt2f& in, Pt2f& out) { out.x = -in.y; out.y = in.x; }
inline void rot180(const Pt2f& in, Pt2f& out) { out.x = -in.x; out.y = -in.y; }
inline bool first_vec_is_right(const Pt2f& v1, const Pt2f& v2) {
    Pt2f t;
    rot90cw(v1, t);
    return t.x * v2.x + t.y * v2.y < 0;
}

// OpenCV rotcalipers.cpp::rotatingCalipers(points, n, CALIPERS_MINAREARECT, out[6])
void rotating_calipers(const Pt2f* points, int n, float* out) {
    float minarea = FLT_MAX;
    float buf[7] = {0, 0, 0, 0, 0, 0, 0};
    int ibuf0 = 0, ibuf5 = 0;
    std::vector<float> inv_vect_length(n);
    std::vector<Pt2f> vect(n);
    int left = 0, bottom = 0, right = 0, top = 0;
    int seq[4] = {-1, -1, -1, -1};
    Pt2f rot_vect[4];
    float orientation = 0;
    float base_a;
    float base_b = 0;
    float left_x, right_x, top_y, bottom_y;
    Pt2f pt0 = points[0];
    left_x = right_x = pt0.x;
    top_y = bottom_y = pt0.y;
    for (int i = 0; i < n; ++i) {
        double dx, dy;
        if (pt0.x < left_x) left_x = pt0.x, left = i;
        if (pt0.x > right_x) right_x = pt0.x, right = i;
        if (pt0.y > top_y) top_y = pt0.y, top = i;
        if (pt0.y < bottom_y) bottom_y = pt0.y, bottom = i;
        Pt2f pt = points[(i + 1) & (i + 1 < n ? -1 : 0)];
        dx = pt.x - pt0.x;
        dy = pt.y - pt0.y;
        vect[i].x = (float)dx;
        vect[i].y = (float)dy;
        inv_vect_length[i] = (float)(1. / std::sqrt(dx * dx + dy * dy));
        pt0 = pt;
    }
    {
        double ax = vect[n - 1].x;
        double ay = vect[n - 1].y;
        for (int i = 0; i < n; ++i) {
            double bx = vect[i].x;
            double by = vect[i].y;
            double convexity = ax * by - ay * bx;
            if (convexity != 0) {
                orientation = (convexity > 0) ? 1.f : (-1.f);
                break;
            }
            ax = bx;
            ay = by;
        }
    }
    base_a = orientation;
    seq[0] = bottom;
    seq[1] = right;
    seq[2] = top;
    seq[3] = left;
    for (int k = 0; k < n; ++k) {
        int main_element = 0;
        rot_vect[0] = vect[seq[0]];
        rot90cw(vect[seq[1]], rot_vect[1]);
        rot180(vect[seq[2]], rot_vect[2]);
        rot90ccw(vect[seq[3]], rot_vect[3]);
        for (int i = 1; i < 4; ++i)
            if (first_vec_is_right(rot_vect[i], rot_vect[main_element])) main_element = i;
        {
            int pindex = seq[main_element];
            float lead_x = vect[pindex].x * inv_vect_length[pindex];
            float lead_y = vect[pindex].y * inv_vect_length[pindex];
            switch (main_element) {
                case 0: base_a = lead_x; base_b = lead_y; break;
                case 1: base_a = lead_y; base_b = -lead_x; break;
                case 2: base_a = -lead_x; base_b = -lead_y; break;
                case 3: base_a = -lead_y; base_b = lead_x; break;
            }
        }
        seq[main_element] += 1;
        seq[main_element] = (seq[main_element] == n) ? 0 : seq[main_element];
        {
            float dx = points[seq[1]].x - points[seq[3]].x;
            float dy = points[seq[1]].y - points[seq[3]].y;
            float width = dx * base_a + dy * base_b;
            dx = points[seq[2]].x - points[seq[0]].x;
            dy = points[seq[2]].y - points[seq[0]].y;
            float height = -dx * base_b + dy * base_a;
            float area = width * height;
            if (area <= minarea) {
                minarea = area;
                ibuf0 = seq[3];
                buf[1] = base_a;
                buf[2] = width;
                buf[3] = base_b;
                buf[4] = height;
                ibuf5 = seq[0];
                buf[6] = area;
            }
        }
    }
    float A1 = buf[1];
    float B1 = buf[3];
    float A2 = -buf[3];
    float B2 = buf[1];
    float C1 = A1 * points[ibuf0].x + points[ibuf0].y * B1;
    float C2 = A2 * points[ibuf5].x + points[ibuf5].y * B2;
    float idet = 1.f / (A1 * B2 - A2 * B1);
    float px = (C1 * B2 - C2 * B1) * idet;
    float py = (A1 * C2 - A2 * C1) * idet;
    out[0] = px;
    out[1] = py;
    out[2] = A1 * buf[2];
    out[3] = B1 * buf[2];
    out[4] = A2 * buf[4];
    out[5] = B2 * buf[4];
}

}  // namespace

void debug_convex_hull(const int32_t* xy, int n, int clockwise, std::vector<int>& hull) {
    std::vector<Pt> pts(n);
    for (int i = 0; i < n; ++i) { pts[i].x = xy[2 * i]; pts[i].y = xy[2 * i + 1]; }
    convex_hull(pts.data(), n, hull, clockwise != 0);
}

// cv2.boxPoints(cv2.minAreaRect(points)) for int32 (x,y) points
void min_area_box(const int32_t* xy, int n, float* out8) {
    std::vector<Pt> pts(n);
    for (int i = 0; i < n; ++i) { pts[i].x = xy[2 * i]; pts[i].y = xy[2 * i + 1]; }
    std::vector<int> hull;
    convex_hull(pts.data(), n, hull, false);         // minAreaRect: convexHull(points, hull, clockwise=false, returnPoints=true)
    int hn = (int)hull.size();
    std::vector<Pt2f> hp(hn);
    for (int i = 0; i < hn; ++i) { hp[i].x = (float)pts[hull[i]].x; hp[i].y = (float)pts[hull[i]].y; }
    float cx = 0, cy = 0, bw = 0, bh = 0, angle = 0;
    double ad = 0;
    if (hn > 2) {
        float o[6];
        rotating_calipers(hp.data(), hn, o);
        cx = o[0] + (o[2] + o[4]) * 0.5f;
        cy = o[1] + (o[3] + o[5]) * 0.5f;
        bw = (float)std::sqrt((double)o[2] * o[2] + (double)o[3] * o[3]);
        bh = (float)std::sqrt((double)o[4] * o[4] + (double)o[5] * o[5]);
        ad = atan2((double)o[3], (double)o[2]);
    } else if (hn == 2) {
        cx = (hp[0].x + hp[1].x) * 0.5f;
        cy = (hp[0].y + hp[1].y) * 0.5f;
        double dx = hp[1].x - hp[0].x;
        double dy = hp[1].y - hp[0].y;
        bw = (float)std::sqrt(dx * dx + dy * dy);
        bh = 0;
        ad = atan2(dy, dx);
    } else if (hn == 1) {
        cx = hp[0].x;
        cy = hp[0].y;
    }
    // OpenCV >= 4.5.1 reports the angle in [-90, 0).  Measured against cv2 4.13: the caliper angle (degrees, kept in
    // double) in [0, 90) is shifted by -90 with width/height exchanged; 90 becomes -90 without the exchange; the
    // single cast to float happens after that.
    ad = ad * 180 / 3.1415926535897932384626433832795;
    if (hn >= 2) {
        while (ad >= 0.0) { ad -= 90.0; std::swap(bw, bh); }
        while (ad < -90.0) { ad += 90.0; std::swap(bw, bh); }
    }
    angle = (float)ad;
    // RotatedRect::points
    double _angle = angle * 3.1415926535897932384626433832795 / 180.;
    float b = (float)cos(_angle) * 0.5f;
    float a = (float)sin(_angle) * 0.5f;
    float p0x = cx - a * bh - b * bw, p0y = cy + b * bh - a * bw;
    float p1x = cx + a * bh - b * bw, p1y = cy - b * bh - a * bw;
    out8[0] = p0x; out8[1] = p0y;
    out8[2] = p1x; out8[3] = p1y;
    out8[4] = 2 * cx - p0x; out8[5] = 2 * cy - p0y;
    out8[6] = 2 * cx - p1x; out8[7] = 2 * cy - p1y;
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
        float box[8];
        if (pts.empty()) {
            // cv2.minAreaRect on an empty point set returns a zero rect; upstream would then produce a zero box
            for (float& v : box) v = 0.f;
        } else {
            min_area_box(pts.data(), (int)pts.size() / 2, box);
            float dx = box[0] - box[2], dy = box[1] - box[3];
            float bw = std::sqrt(dx * dx + dy * dy);
            dx = box[2] - box[4]; dy = box[3] - box[5];
            float bh = std::sqrt(dx * dx + dy * dy);
            float box_ratio = std::max(bw, bh) / (std::min(bw, bh) + 1e-5f);
            if (std::fabs(1.f - box_ratio) <= 0.1f) {
                box[0] = (float)l; box[1] = (float)t; box[2] = (float)r_; box[3] = (float)t;
                box[4] = (float)r_; box[5] = (float)b; box[6] = (float)l; box[7] = (float)b;
            }
        }
        int start = 0;
        float best = box[0] + box[1];
        for (int i = 1; i < 4; ++i) {
            float s = box[2 * i] + box[2 * i + 1];
            if (s < best) { best = s; start = i; }
        }
        for (int i = 0; i < 4; ++i) {                          // np.roll(box, 4 - start, 0): out[i] = box[(i + start) % 4]
            boxes.push_back(box[2 * ((i + start) % 4)]);
            boxes.push_back(box[2 * ((i + start) % 4) + 1]);
        }
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
        r0[6] = -sx * dx;
        r0[7] = -sy * dx;
        r1[6] = -sx * dy;
        r1[7] = -sy * dy;
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
