// geom.cuh -- cv2.minAreaRect + cv2.boxPoints restated once, for the host (boxes.cpp, fuzzed against cv2 by the CPU tests)
// AND the device (postproc.cu::k_det_boxes): OpenCV imgproc convhull.cpp (Sklansky hull on integer points) and
// rotcalipers.cpp (rotating calipers in float32) with every operation in OpenCV's order.  No std:: containers: the caller
// provides the work arrays.  Both translation units are compiled with --fmad=false, so a * b + c is never contracted and
// the float / double arithmetic is IEEE-identical on x86-64 and on sm_100a; only atan2 / sin / cos come from different
// math libraries (their double results are rounded to float right away).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#else
#define BB_HD inline
#endif

namespace bbocr {
namespace geom {

struct P2i { int x, y; };
struct P2f { float x, y; };

BB_HD int sgn64(int64_t v) { return (v > 0) - (v < 0); }

// OpenCV convhull.cpp::Sklansky_<int, int64>.  `sp[i]` = the i-th point in (x, y, index) order.  The three points of the
// running triple live in registers (on the device the walk is one thread's dependent chain: a shared-memory round trip per
// operand made a 1 400-point component cost ~0.4 ms).
BB_HD int sklansky(const P2i* sp, int start, int end, int* stack, int nsign, int sign2) {
    int incr = end > start ? 1 : -1;
    int pprev = start, pcur = pprev + incr, pnext = pcur + incr;
    int stacksize = 3;
    if (start == end || (sp[start].x == sp[end].x && sp[start].y == sp[end].y)) {
        stack[0] = start;
        return 1;
    }
    stack[0] = pprev;
    stack[1] = pcur;
    stack[2] = pnext;
    end += incr;
    P2i Pprev = sp[pprev], Pcur = sp[pcur];
    P2i Pnext = pnext != end ? sp[pnext] : Pcur;
    while (pnext != end) {
        // the point after pnext is requested one iteration early (pnext either advances by incr or stays): on the device the
        // walk is a single thread's dependent chain and this load's latency was the whole iteration
        const P2i Pahead = (pnext + incr != end) ? sp[pnext + incr] : Pnext;
        int by = Pnext.y - Pcur.y;
        if (sgn64(by) != nsign) {
            int ax = Pcur.x - Pprev.x;
            int bx = Pnext.x - Pcur.x;
            int ay = Pcur.y - Pprev.y;
            int64_t convexity = (int64_t)ay * bx - (int64_t)ax * by;
            if (sgn64(convexity) == sign2 && (ax != 0 || ay != 0)) {
                pprev = pcur; Pprev = Pcur;
                pcur = pnext; Pcur = Pnext;
                pnext += incr; Pnext = Pahead;
                stack[stacksize] = pnext;
                stacksize++;
            } else {
                if (pprev == start) {
                    pcur = pnext; Pcur = Pnext;
                    stack[1] = pcur;
                    pnext += incr; Pnext = Pahead;
                    stack[2] = pnext;
                } else {
                    stack[stacksize - 2] = pnext;
                    pcur = pprev; Pcur = Pprev;
                    pprev = stack[stacksize - 4];
                    Pprev = sp[pprev];
                    stacksize--;
                }
            }
        } else {
            pnext += incr; Pnext = Pahead;
            stack[stacksize - 1] = pnext;
        }
    }
    return --stacksize;
}

BB_HD void swap_ip(int*& a, int*& b) { int* t = a; a = b; b = t; }
BB_HD void swap_i(int& a, int& b) { int t = a; a = b; b = t; }
BB_HD void swap_f(float& a, float& b) { float t = a; a = b; b = t; }

// cv::convexHull(points, hull, clockwise, returnPoints) for CV_32S points -> hull vertex indices (into the ORIGINAL point
// array) in hullbuf; returns their number.  sp[i] = the i-th point in (x, y, index) order, sorted[i] = its original index.
// stack: total + 2 ints, hullbuf: total ints.
BB_HD int convex_hull(const P2i* sp, const int* sorted, int total, int* stack, int* hullbuf, bool clockwise) {
    if (total == 0) return 0;
    int miny_ind = 0, maxy_ind = 0, nout = 0;
    {
        int miny = sp[0].y, maxy = sp[0].y;
        for (int i = 1; i < total; ++i) {
            int y = sp[i].y;
            if (miny > y) { miny = y; miny_ind = i; }
            if (maxy < y) { maxy = y; maxy_ind = i; }
        }
    }
    if (sp[0].x == sp[total - 1].x && sp[0].y == sp[total - 1].y) {
        hullbuf[nout++] = 0;
    } else {
        int* tl_stack = stack;
        int tl_count = sklansky(sp, 0, maxy_ind, tl_stack, -1, 1);
        int* tr_stack = stack + tl_count;
        int tr_count = sklansky(sp, total - 1, maxy_ind, tr_stack, -1, -1);
        if (!clockwise) {
            swap_ip(tl_stack, tr_stack);
            swap_i(tl_count, tr_count);
        }
        for (int i = 0; i < tl_count - 1; ++i) hullbuf[nout++] = sorted[tl_stack[i]];
        for (int i = tr_count - 1; i > 0; --i) hullbuf[nout++] = sorted[tr_stack[i]];
        int stop_idx = tr_count > 2 ? tr_stack[1] : tl_count > 2 ? tl_stack[tl_count - 2] : -1;

        int* bl_stack = stack;
        int bl_count = sklansky(sp, 0, miny_ind, bl_stack, 1, -1);
        int* br_stack = stack + bl_count;
        int br_count = sklansky(sp, total - 1, miny_ind, br_stack, 1, 1);
        if (clockwise) {
            swap_ip(bl_stack, br_stack);
            swap_i(bl_count, br_count);
        }
        if (stop_idx >= 0) {
            int check_idx = bl_count > 2 ? bl_stack[1] : bl_count + br_count > 2 ? br_stack[2 - bl_count] : -1;
            if (check_idx == stop_idx || (check_idx >= 0 && sp[check_idx].x == sp[stop_idx].x && sp[check_idx].y == sp[stop_idx].y)) {
                bl_count = bl_count < 2 ? bl_count : 2;
                br_count = br_count < 2 ? br_count : 2;
            }
        }
        for (int i = 0; i < bl_count - 1; ++i) hullbuf[nout++] = sorted[bl_stack[i]];
        for (int i = br_count - 1; i > 0; --i) hullbuf[nout++] = sorted[br_stack[i]];

        if (nout >= 3) {
            int min_idx = 0, max_idx = 0, lt = 0, i;
            for (i = 1; i < nout; ++i) {
                int idx = hullbuf[i];
                lt += hullbuf[i - 1] < idx;
                if (lt > 1 && lt <= i - 2) break;
                if (idx < hullbuf[min_idx]) min_idx = i;
                if (idx > hullbuf[max_idx]) max_idx = i;
            }
            int mmdist = max_idx - min_idx;
            if (mmdist < 0) mmdist = -mmdist;
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
                    if (i == nout)
                        for (int q = 0; q < nout; ++q) hullbuf[q] = stack[q];
                }
            }
        }
    }
    return nout;
}

BB_HD void rot90cw(const P2f& in, P2f& out) { out.x = in.y; out.y = -in.x; }
BB_HD void rot90ccw(const P2f& in, P2f& out) { out.x = -in.y; out.y = in.x; }
BB_HD void rot180(const P2f& in, P2f& out) { out.x = -in.x; out.y = -in.y; }
BB_HD bool first_vec_is_right(const P2f& v1, const P2f& v2) {
    P2f t;
    rot90cw(v1, t);
    return t.x * v2.x + t.y * v2.y < 0;
}

// OpenCV rotcalipers.cpp::rotatingCalipers(points, n, CALIPERS_MINAREARECT, out[6]); work arrays: n floats, n P2f
BB_HD void rotating_calipers(const P2f* points, int n, float* inv_vect_length, P2f* vect, float* out) {
    float minarea = FLT_MAX;
    float buf[7] = {0, 0, 0, 0, 0, 0, 0};
    int ibuf0 = 0, ibuf5 = 0;
    int left = 0, bottom = 0, right = 0, top = 0;
    int seq[4] = {-1, -1, -1, -1};
    P2f rot_vect[4];
    float orientation = 0;
    float base_a;
    float base_b = 0;
    float left_x, right_x, top_y, bottom_y;
    P2f pt0 = points[0];
    left_x = right_x = pt0.x;
    top_y = bottom_y = pt0.y;
    for (int i = 0; i < n; ++i) {
        double dx, dy;
        if (pt0.x < left_x) left_x = pt0.x, left = i;
        if (pt0.x > right_x) right_x = pt0.x, right = i;
        if (pt0.y > top_y) top_y = pt0.y, top = i;
        if (pt0.y < bottom_y) bottom_y = pt0.y, bottom = i;
        P2f pt = points[(i + 1) & (i + 1 < n ? -1 : 0)];
        dx = pt.x - pt0.x;
        dy = pt.y - pt0.y;
        vect[i].x = (float)dx;
        vect[i].y = (float)dy;
        inv_vect_length[i] = (float)(1. / sqrt(dx * dx + dy * dy));
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

// cv2.boxPoints(cv2.minAreaRect(points)) given the convex hull (vertex indices into pts).  hp: hn P2f; inv_len: hn floats;
// vect: hn P2f.
BB_HD void min_area_box_from_hull(const P2i* pts, const int* hull, int hn, P2f* hp, float* inv_len, P2f* vect, float* out8) {
    for (int i = 0; i < hn; ++i) { hp[i].x = (float)pts[hull[i]].x; hp[i].y = (float)pts[hull[i]].y; }
    float cx = 0, cy = 0, bw = 0, bh = 0, angle = 0;
    double ad = 0;
    if (hn > 2) {
        float o[6];
        rotating_calipers(hp, hn, inv_len, vect, o);
        cx = o[0] + (o[2] + o[4]) * 0.5f;
        cy = o[1] + (o[3] + o[5]) * 0.5f;
        bw = (float)sqrt((double)o[2] * o[2] + (double)o[3] * o[3]);
        bh = (float)sqrt((double)o[4] * o[4] + (double)o[5] * o[5]);
        ad = atan2((double)o[3], (double)o[2]);
    } else if (hn == 2) {
        cx = (hp[0].x + hp[1].x) * 0.5f;
        cy = (hp[0].y + hp[1].y) * 0.5f;
        double dx = hp[1].x - hp[0].x;
        double dy = hp[1].y - hp[0].y;
        bw = (float)sqrt(dx * dx + dy * dy);
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
        while (ad >= 0.0) { ad -= 90.0; swap_f(bw, bh); }
        while (ad < -90.0) { ad += 90.0; swap_f(bw, bh); }
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

// craft_utils.getDetBoxes_core after minAreaRect: the 0.1 "diamond" rule (replace a near-square box by the axis-aligned
// bounding box l, t, r, b of the points) and the clockwise start at argmin(x + y) (np.roll); box -> out (8 floats)
BB_HD void finish_det_box(float* box, int l, int t, int r, int b, float* out) {
    float dx = box[0] - box[2], dy = box[1] - box[3];
    float bw = sqrtf(dx * dx + dy * dy);
    dx = box[2] - box[4]; dy = box[3] - box[5];
    float bh = sqrtf(dx * dx + dy * dy);
    float mx = bw > bh ? bw : bh, mn = bw < bh ? bw : bh;
    float box_ratio = mx / (mn + 1e-5f);
    if (fabsf(1.f - box_ratio) <= 0.1f) {
        box[0] = (float)l; box[1] = (float)t; box[2] = (float)r; box[3] = (float)t;
        box[4] = (float)r; box[5] = (float)b; box[6] = (float)l; box[7] = (float)b;
    }
    int start = 0;
    float best = box[0] + box[1];
    for (int i = 1; i < 4; ++i) {
        float s = box[2 * i] + box[2 * i + 1];
        if (s < best) { best = s; start = i; }
    }
    for (int i = 0; i < 4; ++i) {                          // np.roll(box, 4 - start, 0): out[i] = box[(i + start) % 4]
        out[2 * i] = box[2 * ((i + start) % 4)];
        out[2 * i + 1] = box[2 * ((i + start) % 4) + 1];
    }
}

}  // namespace geom
}  // namespace bbocr
