// jpeg.cu -- baseline JPEG decode on the device, bit-exact with cv2.imread / cv2.imdecode (OpenCV's bundled libjpeg-turbo
// with its default settings), SURVEY.md §8f-4.  Reference call sites:
//   pipeline_demo/ocr_testing/preprocessing/image_preprocessor.py:18   cv2.imread(image_path)          (175 ms per 12 MP photo)
//   easyocr/utils.py::reformat_input                                    cv2.imread(path, IMREAD_GRAYSCALE) + colour read
// What libjpeg-turbo does for such a file, and where it happens here:
//   jdmarker.c   marker segments (SOF0/1, DQT, DHT, DRI, SOS, APP1 Exif)            host   parse_jpeg
//   jdhuff.c     Huffman decoding of the interleaved scan                            device k_jpeg_huff: ONE THREAD PER RESTART
//                INTERVAL (phone cameras write one per MCU row: all 63 iPhone photos under the reference's books/ carry
//                DRI = MCUs per row, i.e. 189-268 independent segments per photo); files without restart markers are
//                bit-serial by construction and take the same routine on a host thread (decode_segment is host/device code)
//   jidctint.c   dequantisation + jpeg_idct_islow (13-bit constants, two passes)     device k_jpeg_idct, one thread per block
//   jdsample.c   h2v1 / h2v2 / h1v2 "fancy" triangle up-sampling, jdmainct.c's replicated context rows
//   jdcolor.c    YCbCr -> RGB through the 16-bit fixed-point tables                  device k_jpeg_color (fused, 32x32 tiles)
//   loadsave.cpp ExifTransform (orientation 1..8)                                    device k_jpeg_color: tiles are written
//                through shared memory in destination order, so rotated outputs stay coalesced
// IMREAD_GRAYSCALE makes libjpeg emit the luma plane only (out_color_space = JCS_GRAYSCALE): the gray output is Y.
// Oracle: oracle/jpeg_np.py (pinned against cv2.imdecode).  Progressive, arithmetic, 12-bit, CMYK and multi-scan files are
// refused with BBOCR_E_UNSUPPORTED -- the Python surface then takes the host decoder it always used.
#include <thread>

#include "engine.h"

namespace bbocr {

namespace {

constexpr int LA = 9;                        // look-ahead bits of the Huffman fast path

struct HuffTab {                             // jdhuff.c::jpeg_make_d_derived_tbl
    uint16_t look[1 << LA];                  // (code length << 8) | symbol, 0 = longer than LA bits
    int32_t maxcode[18];                     // largest code of each length (-1 = none); [17] = sentinel
    int32_t valoff[18];
    uint8_t vals[256];
    int16_t fast[1 << LA];                   // AC tables: (value << 8) | (run << 4) | (code + magnitude bits) when both fit into LA
                                             // bits and the value into a signed byte, else 0 (the stb_image "fast AC" trick)
};

struct JpegComp {
    int id, h, v, tq, td, ta;
    int bx, by;                              // blocks per row / column as stored (padded to whole MCUs when interleaved)
    int dw, dh;                              // downsampled_width / height: the real samples
    long long coef_off;                      // first block of this component in the coefficient buffer
    long long plane_off;                     // first byte of its sample plane
};

struct JpegInfo {
    int H = 0, W = 0, ncomp = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0, restart = 0, orientation = 1;
    JpegComp c[3];
    uint16_t q[4][64];                       // natural (row-major) order
    bool has_q[4] = {false, false, false, false}, has_dc[4] = {false, false, false, false}, has_ac[4] = {false, false, false, false};
    HuffTab dc[4], ac[4];
    size_t scan_start = 0, scan_end = 0;
    long long n_blocks = 0, plane_bytes = 0;
};

struct ScanDesc {                            // kernel parameter of the entropy decoder
    int ncomp, mcux, nmcu, restart;
    int h[3], v[3], bx[3], dc[3], ac[3], store[3];
    long long coef_off[3];
};

__constant__ uint8_t c_zigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                                     28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                                     54, 47, 55, 62, 63};
const uint8_t h_zigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                              28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                              54, 47, 55, 62, 63};

[[noreturn]] void unsupported(const char* what) { fail(BBOCR_E_UNSUPPORTED, "JPEG: %s", what); }

void build_huff(const uint8_t* counts, const uint8_t* vals, int nvals, HuffTab& t) {
    memset(&t, 0, sizeof t);
    memcpy(t.vals, vals, nvals);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        t.valoff[len] = k - code;
        for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
            if (len <= LA) {
                const int first = code << (LA - len);
                for (int j = 0; j < (1 << (LA - len)); ++j) t.look[first + j] = (uint16_t)((len << 8) | vals[k]);
            }
        }
        t.maxcode[len] = counts[len - 1] ? code - 1 : -1;
        code <<= 1;
    }
    t.maxcode[17] = 0x7fffffff;
    for (int i = 0; i < (1 << LA); ++i) {
        const int e = t.look[i];
        if (!e) continue;
        const int len = e >> 8, run = (e >> 4) & 15, mag = e & 15;
        if (mag == 0 || len + mag > LA) continue;
        int v = ((i << len) & ((1 << LA) - 1)) >> (LA - mag);
        if (v < (1 << (mag - 1))) v += (int)((~0u) << mag) + 1;
        if (v >= -128 && v <= 127) t.fast[i] = (int16_t)(v * 256 + run * 16 + len + mag);
    }
}

int exif_orientation(const uint8_t* t, size_t n) {          // TIFF header of an APP1 "Exif\0\0" segment -> tag 0x0112 of IFD0
    if (n < 8) return 0;
    const bool le = t[0] == 'I' && t[1] == 'I';
    if (!le && !(t[0] == 'M' && t[1] == 'M')) return 0;
    auto rd16 = [&](size_t o) { return le ? (int)(t[o] | (t[o + 1] << 8)) : (int)((t[o] << 8) | t[o + 1]); };
    auto rd32 = [&](size_t o) { return le ? (uint32_t)(t[o] | (t[o + 1] << 8) | (t[o + 2] << 16) | ((uint32_t)t[o + 3] << 24))
                                          : (uint32_t)(((uint32_t)t[o] << 24) | (t[o + 1] << 16) | (t[o + 2] << 8) | t[o + 3]); };
    const size_t off = rd32(4);
    if (off + 2 > n) return 0;
    const int cnt = rd16(off);
    for (int k = 0; k < cnt; ++k) {
        const size_t e = off + 2 + (size_t)12 * k;
        if (e + 12 > n) return 0;
        if (rd16(e) == 0x0112) return rd16(e + 8);
    }
    return 0;
}

void parse_jpeg(const uint8_t* d, size_t n, JpegInfo& J) {
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) fail(BBOCR_E_ARG, "not a JPEG stream");
    size_t pos = 2;
    bool have_frame = false, have_scan = false;
    int adobe_transform = -1;
    while (pos + 4 <= n) {
        if (d[pos] != 0xFF) fail(BBOCR_E_ARG, "JPEG: marker expected at byte %zu", pos);
        while (pos < n && d[pos] == 0xFF) ++pos;
        if (pos >= n) break;
        const int m = d[pos++];
        if (m == 0xD9) break;
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (pos + 2 > n) break;
        const size_t L = ((size_t)d[pos] << 8) | d[pos + 1];
        if (L < 2 || pos + L > n) fail(BBOCR_E_ARG, "JPEG: truncated segment");
        const uint8_t* s = d + pos + 2;
        const size_t sl = L - 2;
        pos += L;
        if (m == 0xDB) {
            for (size_t i = 0; i < sl;) {
                const int pq = s[i] >> 4, tq = s[i] & 15;
                ++i;
                if (tq > 3 || i + (pq ? 128 : 64) > sl) fail(BBOCR_E_ARG, "JPEG: bad DQT");
                for (int k = 0; k < 64; ++k) {
                    J.q[tq][h_zigzag[k]] = pq ? (uint16_t)((s[i + 2 * k] << 8) | s[i + 2 * k + 1]) : s[i + k];
                }
                i += pq ? 128 : 64;
                J.has_q[tq] = true;
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (sl < 6 || s[0] != 8) unsupported("only 8-bit samples");
            J.H = (s[1] << 8) | s[2];
            J.W = (s[3] << 8) | s[4];
            J.ncomp = s[5];
            if (J.ncomp != 1 && J.ncomp != 3) unsupported("CMYK / YCCK or two-component file");
            if (sl < (size_t)6 + 3 * J.ncomp || J.H <= 0 || J.W <= 0) fail(BBOCR_E_ARG, "JPEG: bad SOF");
            for (int k = 0; k < J.ncomp; ++k) {
                J.c[k].id = s[6 + 3 * k];
                J.c[k].h = s[7 + 3 * k] >> 4;
                J.c[k].v = s[7 + 3 * k] & 15;
                J.c[k].tq = s[8 + 3 * k];
                if (J.c[k].tq > 3) fail(BBOCR_E_ARG, "JPEG: bad quantisation table index");
            }
            have_frame = true;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            unsupported("progressive / lossless / arithmetic coding");
        } else if (m == 0xC4) {
            for (size_t i = 0; i < sl;) {
                if (i + 17 > sl) fail(BBOCR_E_ARG, "JPEG: bad DHT");
                const int tc = s[i] >> 4, th = s[i] & 15;
                int nv = 0;
                for (int k = 0; k < 16; ++k) nv += s[i + 1 + k];
                if (tc > 1 || th > 3 || nv > 256 || i + 17 + nv > sl) fail(BBOCR_E_ARG, "JPEG: bad DHT");
                build_huff(s + i + 1, s + i + 17, nv, tc ? J.ac[th] : J.dc[th]);
                (tc ? J.has_ac : J.has_dc)[th] = true;
                i += 17 + nv;
            }
        } else if (m == 0xDD) {
            if (sl < 2) fail(BBOCR_E_ARG, "JPEG: bad DRI");
            J.restart = (s[0] << 8) | s[1];
        } else if (m == 0xE1 && sl > 6 && memcmp(s, "Exif\0\0", 6) == 0) {
            const int o = exif_orientation(s + 6, sl - 6);
            if (o >= 1 && o <= 8) J.orientation = o;
        } else if (m == 0xEE && sl >= 12 && memcmp(s, "Adobe", 5) == 0) {
            adobe_transform = s[11];
        } else if (m == 0xDA) {
            if (!have_frame) fail(BBOCR_E_ARG, "JPEG: SOS before SOF");
            if (sl < 1 || s[0] != J.ncomp) unsupported("multi-scan (non-interleaved) file");
            if (sl < (size_t)1 + 2 * J.ncomp + 3) fail(BBOCR_E_ARG, "JPEG: bad SOS");
            for (int k = 0; k < J.ncomp; ++k) {
                if (s[1 + 2 * k] != J.c[k].id) unsupported("scan components out of frame order");
                J.c[k].td = s[2 + 2 * k] >> 4;
                J.c[k].ta = s[2 + 2 * k] & 15;
                if (J.c[k].td > 3 || J.c[k].ta > 3 || !J.has_dc[J.c[k].td] || !J.has_ac[J.c[k].ta] || !J.has_q[J.c[k].tq])
                    fail(BBOCR_E_ARG, "JPEG: scan refers to a missing table");
            }
            J.scan_start = pos;
            have_scan = true;
            break;
        }
    }
    if (!have_frame || !have_scan) fail(BBOCR_E_ARG, "JPEG: no frame / scan");
    if (J.ncomp == 3) {
        if (adobe_transform == 0 || (J.c[0].id == 'R' && J.c[1].id == 'G' && J.c[2].id == 'B')) unsupported("RGB-coded file");
        if (J.c[1].h != 1 || J.c[1].v != 1 || J.c[2].h != 1 || J.c[2].v != 1 || J.c[0].h > 2 || J.c[0].v > 2 || J.c[0].h < 1 || J.c[0].v < 1)
            unsupported("sampling factors other than 4:4:4, 4:2:2, 4:4:0, 4:2:0");
    } else {
        J.c[0].h = J.c[0].v = 1;                                  // a single-component scan is never interleaved: MCU = one block
    }
    J.hmax = J.c[0].h;
    J.vmax = J.c[0].v;
    J.mcux = cdiv(J.W, 8 * J.hmax);
    J.mcuy = cdiv(J.H, 8 * J.vmax);
    long long blocks = 0, bytes = 0;
    for (int k = 0; k < J.ncomp; ++k) {
        JpegComp& c = J.c[k];
        c.bx = J.mcux * c.h;
        c.by = J.mcuy * c.v;
        c.dw = cdiv(J.W * c.h, J.hmax);
        c.dh = cdiv(J.H * c.v, J.vmax);
        c.coef_off = blocks;
        c.plane_off = bytes;
        blocks += (long long)c.bx * c.by;
        bytes += (long long)c.bx * c.by * 64;
    }
    J.n_blocks = blocks;
    J.plane_bytes = bytes;
    // the entropy-coded segment ends at the first marker that is neither a stuffed zero nor RSTn
    size_t e = J.scan_start;
    while (e + 1 < n) {
        const uint8_t* f = (const uint8_t*)memchr(d + e, 0xFF, n - 1 - e);
        if (!f) { e = n; break; }
        e = (size_t)(f - d);
        const int nx = d[e + 1];
        if (nx == 0x00 || (nx >= 0xD0 && nx <= 0xD7) || nx == 0xFF) { e += (nx == 0xFF) ? 1 : 2; continue; }
        break;
    }
    J.scan_end = e < n ? e : n;
}

// restart-interval boundaries inside [scan_start, scan_end): seg[i] = first byte of interval i, seg[nseg] = scan_end;
// ends[i] = one past the last entropy byte of interval i (the RSTn marker is excluded)
void find_segments(const uint8_t* d, const JpegInfo& J, std::vector<uint32_t>& begin, std::vector<uint32_t>& end) {
    begin.clear();
    end.clear();
    begin.push_back((uint32_t)J.scan_start);
    if (J.restart > 0) {
        size_t e = J.scan_start;
        while (e + 1 < J.scan_end) {
            const uint8_t* f = (const uint8_t*)memchr(d + e, 0xFF, J.scan_end - 1 - e);
            if (!f) break;
            e = (size_t)(f - d);
            const int nx = d[e + 1];
            if (nx >= 0xD0 && nx <= 0xD7) {
                end.push_back((uint32_t)e);
                begin.push_back((uint32_t)(e + 2));
                e += 2;
            } else {
                e += (nx == 0xFF) ? 1 : 2;
            }
        }
    }
    end.push_back((uint32_t)J.scan_end);
}

// ---- entropy decoding: jdhuff.c::decode_mcu over the MCUs [mcu0, mcu1) of one restart interval ---------------------------
// The bit reservoir of jdhuff.c (64 bits), fed from the interval's bytes with the stuffed zero after every 0xFF removed.
// Host: byte by byte.  Device: a dependent chain of single-byte loads and branches per symbol is what a SIMT lane is worst at,
// so the lane keeps 4-7 stream bytes in a register (aligned 32-bit loads) and moves FOUR bytes into the reservoir at once
// whenever they contain no 0xFF (98.5 % of the time); only words with an 0xFF take the byte-wise path.
struct BitReader {
    uint32_t left;             // bytes of the interval not yet moved into the reservoir
    uint64_t acc;
    int n;
#ifdef __CUDA_ARCH__
    const uint32_t* wp;        // next aligned word
    uint64_t word;             // stream bytes fetched but not yet consumed, lowest first
    int wn;
    __device__ BitReader(const uint8_t* begin, const uint8_t* end) : left((uint32_t)(end - begin)), acc(0), n(0) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(begin);
        const int mis = (int)(a & 3);
        wp = reinterpret_cast<const uint32_t*>(a - mis);       // the bytes in front of `begin` belong to the same upload
        word = (uint64_t)(__ldg(wp++) >> (8 * mis));
        wn = 4 - mis;
    }
    __device__ __forceinline__ uint32_t next_byte() {
        if (left == 0) return 0;                               // past the interval: zeros, like libjpeg at a marker
        --left;
        if (wn == 0) { word = __ldg(wp++); wn = 4; }           // the upload is padded: whole words are readable
        const uint32_t b = (uint32_t)(word & 0xff);
        word >>= 8;
        --wn;
        return b;
    }
    __device__ __forceinline__ void fill() {
        if (n > 32) return;
        if (wn < 4) { word |= (uint64_t)__ldg(wp++) << (8 * wn); wn += 4; }
        const uint32_t b4 = (uint32_t)word;
        if (left >= 4 && !((~b4 - 0x01010101u) & b4 & 0x80808080u)) {      // no byte of b4 is 0xFF
            acc |= (uint64_t)__byte_perm(b4, 0, 0x0123) << (32 - n);
            n += 32;
            word >>= 32;
            wn -= 4;
            left -= 4;
            return;
        }
        while (n <= 56) {
            const uint32_t b = next_byte();
            if (b == 0xFF) next_byte();                        // FF 00: the stuffed zero (a marker never lies inside an interval)
            acc |= (uint64_t)b << (56 - n);
            n += 8;
        }
    }
#else
    const uint8_t* p;
    BitReader(const uint8_t* begin, const uint8_t* end) : left((uint32_t)(end - begin)), acc(0), n(0), p(begin) {}
    uint32_t next_byte() {
        if (left == 0) return 0;
        --left;
        return *p++;
    }
    void fill() {
        while (n <= 56) {
            const uint32_t b = next_byte();
            if (b == 0xFF) next_byte();
            acc |= (uint64_t)b << (56 - n);
            n += 8;
        }
    }
#endif
    __host__ __device__ __forceinline__ uint32_t peek(int k) const { return (uint32_t)(acc >> (64 - k)); }
    __host__ __device__ __forceinline__ void skip(int k) { acc <<= k; n -= k; }
};

__host__ __device__ __forceinline__ int huff_symbol(BitReader& br, const HuffTab& t) {
    const uint32_t e = t.look[br.peek(LA)];
    if (e) {
        br.skip((int)(e >> 8));
        return (int)(e & 255);
    }
    int len = LA + 1;
    int code = (int)br.peek(len);
    while (len < 17 && code > t.maxcode[len]) {
        ++len;
        code = (int)br.peek(len);
    }
    if (len > 16) { br.skip(16); return 0; }              // corrupt data: libjpeg warns and returns zero
    br.skip(len);
    return t.vals[(code + t.valoff[len]) & 255];
}

__host__ __device__ __forceinline__ int receive_extend(BitReader& br, int s) {
    if (s == 0) return 0;
    const int v = (int)br.peek(s);
    br.skip(s);
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

template <typename ZZ>
__host__ __device__ __forceinline__ void decode_segment(const uint8_t* data, uint32_t begin, uint32_t end, const ScanDesc& S, const HuffTab* tabs,
                                               int mcu0, int mcu1, int16_t* coef, ZZ zigzag) {
    BitReader br(data + begin, data + end);
    int pred0 = 0, pred1 = 0, pred2 = 0;
    for (int m = mcu0; m < mcu1; ++m) {
        const int my = m / S.mcux, mx = m - my * S.mcux;
        for (int ci = 0; ci < S.ncomp; ++ci) {
            const HuffTab& dct = tabs[S.dc[ci]];
            const HuffTab& act = tabs[4 + S.ac[ci]];
            for (int by = 0; by < S.v[ci]; ++by)
                for (int bx = 0; bx < S.h[ci]; ++bx) {
                    // the coefficient buffer was zeroed up front (one memset per image): only the non-zero values are stored
                    int16_t* blk = nullptr;
                    if (S.store[ci]) blk = coef + (S.coef_off[ci] + (long long)(my * S.v[ci] + by) * S.bx[ci] + (mx * S.h[ci] + bx)) * 64;
                    br.fill();
                    int s = huff_symbol(br, dct);
                    const int diff = receive_extend(br, s & 15);
                    int dcv;
                    if (ci == 0) dcv = (pred0 += diff);
                    else if (ci == 1) dcv = (pred1 += diff);
                    else dcv = (pred2 += diff);
                    if (blk) blk[0] = (int16_t)dcv;
                    for (int k = 1; k < 64;) {
                        br.fill();
                        const int f = act.fast[br.peek(LA)];
                        if (f) {                                   // short code + small value: one lookup
                            k += (f >> 4) & 15;
                            br.skip(f & 15);
                            if (blk) blk[zigzag(k & 63)] = (int16_t)(f >> 8);
                            ++k;
                            continue;
                        }
                        const int rs = huff_symbol(br, act);
                        const int r = rs >> 4;
                        s = rs & 15;
                        if (s == 0) {
                            if (r != 15) break;
                            k += 16;
                            continue;
                        }
                        k += r;
                        const int v = receive_extend(br, s);
                        if (blk) blk[zigzag(k & 63)] = (int16_t)v;
                        ++k;
                    }
                }
        }
    }
}

struct ZigDev { __device__ int operator()(int k) const { return c_zigzag[k]; } };
struct ZigHost { int operator()(int k) const { return h_zigzag[k]; } };

// One thread per restart interval, `lpw` of them per warp: the threads of a warp walk different bit streams, so every
// data-dependent branch (reservoir refill, end of block, long codes) serialises the warp; with few active lanes per warp the
// intervals run close to their own latency chain, and the idle lanes cost nothing (a photo has ~200-270 intervals).
// (kernel: k_jpeg_huff_group below)

// ---- jidctint.c::jpeg_idct_islow: one thread per 8x8 block ----------------------------------------------------------------
__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

__device__ __forceinline__ void idct_1d(int& v0, int& v1, int& v2, int& v3, int& v4, int& v5, int& v6, int& v7, int shift) {
    int z1 = (v2 + v6) * 4433;
    const int tmp2 = z1 + v6 * (-15137);
    const int tmp3 = z1 + v2 * 6270;
    const int tmp0 = (v0 + v4) << 13;
    const int tmp1 = (v0 - v4) << 13;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    int t0 = v7, t1 = v5, t2 = v3, t3 = v1;
    z1 = t0 + t3;
    int z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
    const int z5 = (z3 + z4) * 9633;
    t0 *= 2446; t1 *= 16819; t2 *= 25172; t3 *= 12299;
    z1 *= -7373; z2 *= -20995;
    z3 = z3 * -16069 + z5;
    z4 = z4 * -3196 + z5;
    t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
    v0 = descale(tmp10 + t3, shift); v7 = descale(tmp10 - t3, shift);
    v1 = descale(tmp11 + t2, shift); v6 = descale(tmp11 - t2, shift);
    v2 = descale(tmp12 + t1, shift); v5 = descale(tmp12 - t1, shift);
    v3 = descale(tmp13 + t0, shift); v4 = descale(tmp13 - t0, shift);
}

__device__ __forceinline__ uint32_t range_limit(int x) {          // sample_range_limit + CENTERJSAMPLE, index & RANGE_MASK
    const int i = x & 1023;
    return i < 128 ? i + 128 : i < 512 ? 255 : i < 896 ? 0 : i - 896;
}

__global__ void __launch_bounds__(128)
    k_jpeg_idct(int16_t* __restrict__ coef, const uint16_t* __restrict__ qt /*[64]*/, int bx, int n_blocks, uint8_t* __restrict__ plane) {
    __shared__ int q[64];
    if (threadIdx.x < 64) q[threadIdx.x] = qt[threadIdx.x];
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    int v[64];
    // the block is handed back zeroed: the entropy decoder of the NEXT image on this lane stores only non-zero coefficients
    // (a cudaMemsetAsync of the 36 MB buffer per photo runs at copy-engine speed and serialised the lanes)
    int4* src = reinterpret_cast<int4*>(coef + (size_t)b * 64);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int4 w = src[r];
        src[r] = int4{0, 0, 0, 0};
        const int wi[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[r * 8 + 2 * j] = (int)(int16_t)(wi[j] & 0xffff) * q[r * 8 + 2 * j];
            v[r * 8 + 2 * j + 1] = (wi[j] >> 16) * q[r * 8 + 2 * j + 1];
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) idct_1d(v[c], v[8 + c], v[16 + c], v[24 + c], v[32 + c], v[40 + c], v[48 + c], v[56 + c], 13 - 2);
    const int by = b / bx, bxx = b - by * bx;
    uint8_t* dst = plane + ((size_t)by * 8) * ((size_t)bx * 8) + (size_t)bxx * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        idct_1d(v[r * 8], v[r * 8 + 1], v[r * 8 + 2], v[r * 8 + 3], v[r * 8 + 4], v[r * 8 + 5], v[r * 8 + 6], v[r * 8 + 7], 13 + 2 + 3);
        uint2 o;
        o.x = range_limit(v[r * 8]) | (range_limit(v[r * 8 + 1]) << 8) | (range_limit(v[r * 8 + 2]) << 16) | (range_limit(v[r * 8 + 3]) << 24);
        o.y = range_limit(v[r * 8 + 4]) | (range_limit(v[r * 8 + 5]) << 8) | (range_limit(v[r * 8 + 6]) << 16) | (range_limit(v[r * 8 + 7]) << 24);
        *reinterpret_cast<uint2*>(dst + (size_t)r * bx * 8) = o;
    }
}

// ---- jdsample.c + jdcolor.c + ExifTransform ---------------------------------------------------------------------------------
struct ColorDesc {
    const uint8_t* y;
    const uint8_t* cb;
    const uint8_t* cr;
    int ypitch, cpitch, cdw, cdh;            // chroma downsampled_width / height
    int H, W, hx, vx;                        // up-sampling factors of the chroma planes (1 or 2 each)
    int ncomp, orientation;
    uint8_t* bgr;                            // dest: (o < 5 ? H x W : W x H) x 3, or null
    uint8_t* gray;                           // dest luma plane, or null
};

__device__ __forceinline__ int chroma_at(const uint8_t* __restrict__ p, const ColorDesc& D, int x, int y) {
    const bool fancy = D.cdw > 2;
    if (D.hx == 2 && D.vx == 2) {
        const int c = x >> 1, r = y >> 1;
        if (!fancy) return p[(size_t)r * D.cpitch + c];
        const int o = (y & 1) ? min(r + 1, D.cdh - 1) : max(r - 1, 0);
        const uint8_t* p0 = p + (size_t)r * D.cpitch;
        const uint8_t* p1 = p + (size_t)o * D.cpitch;
        const int cs = 3 * p0[c] + p1[c];
        if (x & 1) {
            if (c == D.cdw - 1) return (4 * cs + 7) >> 4;
            return (3 * cs + 3 * p0[c + 1] + p1[c + 1] + 7) >> 4;
        }
        if (c == 0) return (4 * cs + 8) >> 4;
        return (3 * cs + 3 * p0[c - 1] + p1[c - 1] + 8) >> 4;
    }
    if (D.hx == 2) {                                               // h2v1
        const int c = x >> 1;
        const uint8_t* p0 = p + (size_t)y * D.cpitch;
        if (!fancy) return p0[c];
        if (x & 1) return c == D.cdw - 1 ? p0[c] : (3 * p0[c] + p0[c + 1] + 2) >> 2;
        return c == 0 ? p0[c] : (3 * p0[c] + p0[c - 1] + 1) >> 2;
    }
    if (D.vx == 2) {                                               // h1v2
        const int r = y >> 1;
        if (!fancy) return p[(size_t)r * D.cpitch + x];
        const int o = (y & 1) ? min(r + 1, D.cdh - 1) : max(r - 1, 0);
        return (3 * p[(size_t)r * D.cpitch + x] + p[(size_t)o * D.cpitch + x] + ((y & 1) ? 2 : 1)) >> 2;
    }
    return p[(size_t)y * D.cpitch + x];
}

__device__ __forceinline__ void ycc_bgr(int yy, int cb, int cr, int& b, int& g, int& r) {       // jdcolor.c tables, inlined
    cb -= 128; cr -= 128;
    r = min(max(yy + ((91881 * cr + 32768) >> 16), 0), 255);
    g = min(max(yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16), 0), 255);
    b = min(max(yy + ((116130 * cb + 32768) >> 16), 0), 255);
}

// h2v2_fancy_upsample for the four pixels x .. x+3 (x % 4 == 0) of row y: chroma columns c0 = x / 2 and c0 + 1, their outer
// neighbours c0 - 1 and c0 + 2, from the nearer row r and the further row o (clamped reads; unused values are never selected)
__device__ __forceinline__ void chroma4_h2v2(const uint8_t* __restrict__ p, const ColorDesc& D, int x, int y, int (&out)[4]) {
    const int c0 = x >> 1, r = y >> 1;
    const int o = (y & 1) ? min(r + 1, D.cdh - 1) : max(r - 1, 0);
    const uint8_t* p0 = p + (size_t)r * D.cpitch;
    const uint8_t* p1 = p + (size_t)o * D.cpitch;
    int cs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = min(max(c0 - 1 + k, 0), D.cdw - 1);
        cs[k] = 3 * p0[c] + p1[c];
    }
    out[0] = c0 == 0 ? (4 * cs[1] + 8) >> 4 : (3 * cs[1] + cs[0] + 8) >> 4;
    out[1] = c0 == D.cdw - 1 ? (4 * cs[1] + 7) >> 4 : (3 * cs[1] + cs[2] + 7) >> 4;
    out[2] = (3 * cs[2] + cs[1] + 8) >> 4;
    out[3] = c0 + 1 == D.cdw - 1 ? (4 * cs[2] + 7) >> 4 : (3 * cs[2] + cs[3] + 7) >> 4;
}

__global__ void __launch_bounds__(256) k_jpeg_color(ColorDesc D) {
    __shared__ __align__(16) uint8_t sb[32][32 * 3 + 4];      // full tiles: indexed in DESTINATION order [row][col * 3]
    __shared__ __align__(16) uint8_t sg[32][32 + 4];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const bool want_bgr = D.bgr != nullptr, want_gray = D.gray != nullptr;
    const int o = D.orientation;
    const int dW = o < 5 ? D.W : D.H;
    // ---- fast path: a full 32 x 32 tile whose destination rows are word-aligned: four pixels per thread in, 32-bit words out ----
    if (x0 + 32 <= D.W && y0 + 32 <= D.H && (dW & 3) == 0 && ((D.ypitch | D.cpitch) & 3) == 0) {
        const int ly = threadIdx.x >> 3, lx = (threadIdx.x & 7) * 4;
        const int x = x0 + lx, y = y0 + ly;
        const uint32_t y4 = *reinterpret_cast<const uint32_t*>(D.y + (size_t)y * D.ypitch + x);
        int cb[4], cr[4];
        if (want_bgr && D.ncomp == 3) {
            if (D.hx == 2 && D.vx == 2 && D.cdw > 2) {
                chroma4_h2v2(D.cb, D, x, y, cb);
                chroma4_h2v2(D.cr, D, x, y, cr);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) { cb[k] = chroma_at(D.cb, D, x + k, y); cr[k] = chroma_at(D.cr, D, x + k, y); }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int yy = (y4 >> (8 * k)) & 255;
            const int sx = lx + k;
            int drow, dcol;                                       // destination-local position of source pixel (sx, ly)
            switch (o) {
                case 2: drow = ly; dcol = 31 - sx; break;
                case 3: drow = 31 - ly; dcol = 31 - sx; break;
                case 4: drow = 31 - ly; dcol = sx; break;
                case 5: drow = sx; dcol = ly; break;
                case 6: drow = sx; dcol = 31 - ly; break;
                case 7: drow = 31 - sx; dcol = 31 - ly; break;
                case 8: drow = 31 - sx; dcol = ly; break;
                default: drow = ly; dcol = sx; break;
            }
            sg[drow][dcol] = (uint8_t)yy;
            if (want_bgr) {
                int b = yy, g = yy, r = yy;
                if (D.ncomp == 3) ycc_bgr(yy, cb[k], cr[k], b, g, r);
                sb[drow][3 * dcol] = (uint8_t)b; sb[drow][3 * dcol + 1] = (uint8_t)g; sb[drow][3 * dcol + 2] = (uint8_t)r;
            }
        }
        __syncthreads();
        int dx0, dy0;                                             // destination origin of the tile
        switch (o) {
            case 2: dx0 = D.W - x0 - 32; dy0 = y0; break;
            case 3: dx0 = D.W - x0 - 32; dy0 = D.H - y0 - 32; break;
            case 4: dx0 = x0; dy0 = D.H - y0 - 32; break;
            case 5: dx0 = y0; dy0 = x0; break;
            case 6: dx0 = D.H - y0 - 32; dy0 = x0; break;
            case 7: dx0 = D.H - y0 - 32; dy0 = D.W - x0 - 32; break;
            case 8: dx0 = y0; dy0 = D.W - x0 - 32; break;
            default: dx0 = x0; dy0 = y0; break;
        }
        if (want_bgr && (reinterpret_cast<uintptr_t>(D.bgr) & 3) == 0) {
            for (int w = threadIdx.x; w < 32 * 24; w += 256) {    // 96 bytes = 24 words per destination row
                const int row = w / 24, col = w - row * 24;
                reinterpret_cast<uint32_t*>(D.bgr + ((size_t)(dy0 + row) * dW + dx0) * 3)[col] = reinterpret_cast<const uint32_t*>(sb[row])[col];
            }
        } else if (want_bgr) {
            for (int i = threadIdx.x; i < 32 * 96; i += 256) {
                const int row = i / 96, col = i - row * 96;
                D.bgr[((size_t)(dy0 + row) * dW + dx0) * 3 + col] = sb[row][col];
            }
        }
        if (want_gray && (reinterpret_cast<uintptr_t>(D.gray) & 3) == 0) {
            const int row = threadIdx.x >> 3, col = threadIdx.x & 7;
            reinterpret_cast<uint32_t*>(D.gray + (size_t)(dy0 + row) * dW + dx0)[col] = reinterpret_cast<const uint32_t*>(sg[row])[col];
        } else if (want_gray) {
            for (int i = threadIdx.x; i < 1024; i += 256) D.gray[(size_t)(dy0 + (i >> 5)) * dW + dx0 + (i & 31)] = sg[i >> 5][i & 31];
        }
        return;
    }
    // ---- general path (edge tiles, odd widths): pixel by pixel, source order in shared memory --------------------------------
    for (int i = threadIdx.x; i < 1024; i += 256) {
        const int ly = i >> 5, lx = i & 31;
        const int x = x0 + lx, y = y0 + ly;
        if (x >= D.W || y >= D.H) continue;
        const int yy = D.y[(size_t)y * D.ypitch + x];
        sg[ly][lx] = (uint8_t)yy;
        if (want_bgr) {
            int b = yy, g = yy, r = yy;
            if (D.ncomp == 3) ycc_bgr(yy, chroma_at(D.cb, D, x, y), chroma_at(D.cr, D, x, y), b, g, r);
            sb[ly][3 * lx] = (uint8_t)b; sb[ly][3 * lx + 1] = (uint8_t)g; sb[ly][3 * lx + 2] = (uint8_t)r;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += 256) {
        // the fast index runs along the destination row: source x for orientations 1-4, source y for 5-8
        const int a = i >> 5, f = i & 31;
        const int lx = o < 5 ? f : a, ly = o < 5 ? a : f;
        const int x = x0 + lx, y = y0 + ly;
        if (x >= D.W || y >= D.H) continue;
        int dx, dy;
        switch (o) {
            case 2: dx = D.W - 1 - x; dy = y; break;
            case 3: dx = D.W - 1 - x; dy = D.H - 1 - y; break;
            case 4: dx = x; dy = D.H - 1 - y; break;
            case 5: dx = y; dy = x; break;
            case 6: dx = D.H - 1 - y; dy = x; break;
            case 7: dx = D.H - 1 - y; dy = D.W - 1 - x; break;
            case 8: dx = y; dy = D.W - 1 - x; break;
            default: dx = x; dy = y; break;
        }
        const size_t di = (size_t)dy * dW + dx;
        if (want_gray) D.gray[di] = sg[ly][lx];
        if (want_bgr) {
            D.bgr[3 * di] = sb[ly][3 * lx];
            D.bgr[3 * di + 1] = sb[ly][3 * lx + 1];
            D.bgr[3 * di + 2] = sb[ly][3 * lx + 2];
        }
    }
}

}  // namespace

void jpeg_info(const uint8_t* data, size_t n, int* H, int* W, int* channels, int* orientation) {
    JpegInfo J;
    parse_jpeg(data, n, J);
    const bool swap = J.orientation >= 5;
    *H = swap ? J.W : J.H;
    *W = swap ? J.H : J.W;
    *channels = J.ncomp;
    *orientation = J.orientation;
}

// Host-only: the quantised DCT coefficients of every block (component after component, blocks in raster order, natural
// coefficient order) through the same parser / segment finder / decode_segment the device path uses.  Parity-test surface.
long long jpeg_coefficients_host(const uint8_t* data, size_t n, int16_t* out, long long cap_blocks) {
    JpegInfo J;
    parse_jpeg(data, n, J);
    if (!out || cap_blocks < J.n_blocks) return J.n_blocks;
    ScanDesc S{};
    S.ncomp = J.ncomp; S.mcux = J.mcux; S.nmcu = J.mcux * J.mcuy; S.restart = J.restart;
    for (int k = 0; k < J.ncomp; ++k) {
        S.h[k] = J.c[k].h; S.v[k] = J.c[k].v; S.bx[k] = J.c[k].bx; S.dc[k] = J.c[k].td; S.ac[k] = J.c[k].ta;
        S.store[k] = 1;
        S.coef_off[k] = J.c[k].coef_off;
    }
    std::vector<uint32_t> sb, se;
    find_segments(data, J, sb, se);
    const int nseg = (int)sb.size();
    if (nseg != (J.restart > 0 ? cdiv(S.nmcu, J.restart) : 1)) fail(BBOCR_E_ARG, "JPEG: restart interval count mismatch");
    std::vector<HuffTab> tabs(8);
    memset(tabs.data(), 0, sizeof(HuffTab) * 8);
    for (int k = 0; k < 4; ++k) {
        if (J.has_dc[k]) tabs[k] = J.dc[k];
        if (J.has_ac[k]) tabs[4 + k] = J.ac[k];
    }
    std::vector<int16_t> tmp((size_t)J.n_blocks * 64 + 8);
    int16_t* base = reinterpret_cast<int16_t*>(((uintptr_t)tmp.data() + 15) & ~(uintptr_t)15);
    memset(base, 0, (size_t)J.n_blocks * 128);
    for (int sgi = 0; sgi < nseg; ++sgi) {
        const int mcu0 = J.restart > 0 ? sgi * J.restart : 0;
        const int mcu1 = J.restart > 0 ? std::min(S.nmcu, mcu0 + J.restart) : S.nmcu;
        decode_segment(data, sb[sgi], se[sgi], S, tabs.data(), mcu0, mcu1, base, ZigHost());
    }
    memcpy(out, base, (size_t)J.n_blocks * 128);
    return J.n_blocks;
}

// ---- a GROUP of files on one lane: one upload, ONE entropy-decoder launch over the restart intervals of all of them ----------
// Entropy decoding is a latency chain per interval (~2 ms per 12 MP photo whatever else the GPU does), so the lane-level unit
// of work is a group: 8 photos = ~1 500-2 100 intervals in one grid.  (Running one photo per stream instead needs 30+ streams
// that really overlap, i.e. CUDA_DEVICE_MAX_CONNECTIONS = 32 -- which costs the detector lanes of readtext 8 %.)
struct ImgDesc {                             // device-visible descriptor of one image of the group
    ScanDesc S;
    uint32_t tabs_off, seg_off, data_off;    // byte offsets into the group blob
    int nseg;
    long long coef_base;                     // first int16 of this image's blocks in the lane's coefficient buffer
};

__global__ void __launch_bounds__(128)
    k_jpeg_huff_group(const uint8_t* __restrict__ blob, const ImgDesc* __restrict__ imgs, const int2* __restrict__ blockmap /*(image, first interval)*/,
                      int16_t* __restrict__ coef, int lpw) {
    __shared__ HuffTab tabs[8];                                   // dc0..3 | ac0..3 of this block's image
    __shared__ ImgDesc I;
    const int2 bm = blockmap[blockIdx.x];
    if (threadIdx.x < (int)(sizeof(ImgDesc) / 4))
        reinterpret_cast<uint32_t*>(&I)[threadIdx.x] = reinterpret_cast<const uint32_t*>(imgs + bm.x)[threadIdx.x];
    __syncthreads();
    const uint32_t* tg = reinterpret_cast<const uint32_t*>(blob + I.tabs_off);
    for (int i = threadIdx.x; i < (int)(sizeof(HuffTab) * 8 / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(tabs)[i] = tg[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    if (lane >= lpw) return;
    const int seg = bm.y + (threadIdx.x >> 5) * lpw + lane;
    if (seg >= I.nseg) return;
    const uint32_t* sg = reinterpret_cast<const uint32_t*>(blob + I.seg_off);
    const int mcu0 = I.S.restart > 0 ? seg * I.S.restart : 0;
    const int mcu1 = I.S.restart > 0 ? min(I.S.nmcu, mcu0 + I.S.restart) : I.S.nmcu;
    decode_segment(blob + I.data_off, sg[seg], sg[I.nseg + seg], I.S, tabs, mcu0, mcu1, coef + I.coef_base, ZigDev());
}

struct GroupImage {
    JpegInfo J;
    std::vector<uint32_t> sb, se;
    bool need_chroma = false, device_huffman = false;
    long long n_blocks = 0, coef_base = 0, plane_base = 0;
    size_t o_tabs = 0, o_q = 0, o_seg = 0, o_data = 0, o_end = 0, o_hostcoef = 0;
};

void jpeg_decode_group_dev(Handle* h, Lane& lane, JpegJob* jobs, int nj, int ignore_orientation) {
    cudaStream_t st = lane.stream;
    if (nj <= 0) return;
    static const int lpw = getenv("BBOCR_JPEG_LPW") ? std::min(32, std::max(1, atoi(getenv("BBOCR_JPEG_LPW")))) : 2;
    std::vector<GroupImage> g(nj);
    long long coef_total = 0, plane_total = 0;
    size_t blob = ((size_t)nj * sizeof(ImgDesc) + 15) & ~(size_t)15;
    int total_blocks = 0;
    for (int i = 0; i < nj; ++i) {
        GroupImage& G = g[i];
        JpegJob& jb = jobs[i];
        parse_jpeg(jb.data, jb.n, G.J);
        JpegInfo& J = G.J;
        if (ignore_orientation) J.orientation = 1;
        ARG_CHECK((long long)J.H * J.W <= (1ll << 30), "JPEG: image too large");
        const bool swap = J.orientation >= 5;
        jb.H = swap ? J.W : J.H;
        jb.W = swap ? J.H : J.W;
        if (!jb.out_bgr && !jb.out_gray) continue;
        G.need_chroma = jb.out_bgr != nullptr && J.ncomp == 3;
        find_segments(jb.data, J, G.sb, G.se);
        const int nseg = (int)G.sb.size();
        const int want_seg = J.restart > 0 ? cdiv(J.mcux * J.mcuy, J.restart) : 1;
        if (nseg != want_seg) fail(BBOCR_E_ARG, "JPEG: %d restart intervals found, %d expected (corrupt stream)", nseg, want_seg);
        G.device_huffman = J.restart > 0 && nseg >= 8;
        G.n_blocks = G.need_chroma ? J.n_blocks : (long long)J.c[0].bx * J.c[0].by;
        G.coef_base = coef_total * 64;
        G.plane_base = plane_total;
        coef_total += G.n_blocks;
        plane_total += G.need_chroma ? J.plane_bytes : G.n_blocks * 64;
        G.o_tabs = blob;
        G.o_q = G.o_tabs + sizeof(HuffTab) * 8;
        G.o_seg = G.o_q + 4 * 64 * 2;
        G.o_data = (G.o_seg + (size_t)nseg * 8 + 15) & ~(size_t)15;
        G.o_end = G.device_huffman ? ((G.o_data + (J.scan_end - J.scan_start) + 16 + 15) & ~(size_t)15) : G.o_data;
        blob = G.o_end;
        if (G.device_huffman) total_blocks += cdiv(nseg, 4 * lpw);
    }
    if (coef_total == 0) return;
    const size_t o_map = blob;
    blob = (o_map + (size_t)std::max(total_blocks, 1) * sizeof(int2) + 15) & ~(size_t)15;
    const size_t blob_bytes = blob;
    size_t pin_bytes = blob_bytes;
    for (auto& G : g)
        if (!G.device_huffman && G.n_blocks) { G.o_hostcoef = pin_bytes; pin_bytes += (size_t)G.n_blocks * 128; }

    if (lane.in_busy) { CUDA_CHECK(stream_sync(st)); lane.in_busy = false; }   // also orders a cudaFree of a growing scratch buffer
    // invariant: a lane's coefficient buffer is all zeros between groups (k_jpeg_idct clears what the decoder wrote)
    const size_t coef_cap_before = lane.scr[0].cap;
    int16_t* dcoef = reinterpret_cast<int16_t*>(lane.scr[0].get((size_t)coef_total * 128));
    if (lane.scr[0].cap != coef_cap_before || lane.scr_dirty) CUDA_CHECK(cudaMemsetAsync(dcoef, 0, lane.scr[0].cap, st));
    lane.scr_dirty = true;                                         // until every IDCT launch of the group is enqueued
    uint8_t* dplanes = reinterpret_cast<uint8_t*>(lane.scr[1].get((size_t)plane_total));
    uint8_t* dblob = reinterpret_cast<uint8_t*>(lane.scr[2].get(blob_bytes));
    uint8_t* pin = (uint8_t*)lane.pin_in.get(pin_bytes);
    ImgDesc* descs = reinterpret_cast<ImgDesc*>(pin);
    int2* bmap = reinterpret_cast<int2*>(pin + o_map);
    int nb_map = 0;
    for (int i = 0; i < nj; ++i) {
        GroupImage& G = g[i];
        ImgDesc& D = descs[i];
        memset(&D, 0, sizeof D);
        if (!G.n_blocks) continue;
        const JpegInfo& J = G.J;
        ScanDesc& S = D.S;
        S.ncomp = J.ncomp; S.mcux = J.mcux; S.nmcu = J.mcux * J.mcuy; S.restart = J.restart;
        for (int k = 0; k < J.ncomp; ++k) {
            S.h[k] = J.c[k].h; S.v[k] = J.c[k].v; S.bx[k] = J.c[k].bx; S.dc[k] = J.c[k].td; S.ac[k] = J.c[k].ta;
            S.store[k] = (k == 0 || G.need_chroma) ? 1 : 0;
            S.coef_off[k] = J.c[k].coef_off;
        }
        const int nseg = (int)G.sb.size();
        D.tabs_off = (uint32_t)G.o_tabs; D.seg_off = (uint32_t)G.o_seg; D.data_off = (uint32_t)G.o_data;
        D.nseg = nseg;
        D.coef_base = G.coef_base;
        HuffTab* tabs = reinterpret_cast<HuffTab*>(pin + G.o_tabs);
        memset(tabs, 0, sizeof(HuffTab) * 8);
        for (int k = 0; k < 4; ++k) {
            if (J.has_dc[k]) tabs[k] = J.dc[k];
            if (J.has_ac[k]) tabs[4 + k] = J.ac[k];
        }
        memcpy(pin + G.o_q, J.q, sizeof J.q);
        if (G.device_huffman) {
            uint32_t* segs = reinterpret_cast<uint32_t*>(pin + G.o_seg);
            for (int q = 0; q < nseg; ++q) {
                segs[q] = G.sb[q] - (uint32_t)J.scan_start;          // relative to the uploaded entropy bytes
                segs[nseg + q] = G.se[q] - (uint32_t)J.scan_start;
            }
            memcpy(pin + G.o_data, jobs[i].data + J.scan_start, J.scan_end - J.scan_start);
            for (int q = 0; q < nseg; q += 4 * lpw) bmap[nb_map++] = int2{i, q};
        } else {
            // no (or too few) restart intervals: the scan is one bit-serial chain -> the same routine on host threads
            int16_t* hc = reinterpret_cast<int16_t*>(pin + G.o_hostcoef);
            memset(hc, 0, (size_t)G.n_blocks * 128);
            const int nt = std::max(1, std::min(nseg, 8));
            std::vector<std::thread> th;
            auto work = [&](int t) {
                for (int sgi = t; sgi < nseg; sgi += nt) {
                    const int mcu0 = J.restart > 0 ? sgi * J.restart : 0;
                    const int mcu1 = J.restart > 0 ? std::min(S.nmcu, mcu0 + J.restart) : S.nmcu;
                    decode_segment(jobs[i].data, G.sb[sgi], G.se[sgi], S, tabs, mcu0, mcu1, hc, ZigHost());
                }
            };
            for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
            work(0);
            for (auto& t : th) t.join();
        }
    }
    if ((size_t)blob_bytes >= (1ull << 32)) fail(BBOCR_E_ARG, "JPEG: group too large");
    CUDA_CHECK(cudaMemcpyAsync(dblob, pin, blob_bytes, cudaMemcpyHostToDevice, st));
    for (auto& G : g)
        if (!G.device_huffman && G.n_blocks)
            CUDA_CHECK(cudaMemcpyAsync(dcoef + G.coef_base, pin + G.o_hostcoef, (size_t)G.n_blocks * 128, cudaMemcpyHostToDevice, st));
    lane.in_busy = true;
    if (nb_map > 0) {
        k_jpeg_huff_group<<<nb_map, 128, 0, st>>>(dblob, reinterpret_cast<const ImgDesc*>(dblob), reinterpret_cast<const int2*>(dblob + o_map),
                                                  dcoef, lpw);
        count_launch(h);
    }
    for (int i = 0; i < nj; ++i) {
        GroupImage& G = g[i];
        if (!G.n_blocks) continue;
        const JpegInfo& J = G.J;
        const int ncomp_run = G.need_chroma ? J.ncomp : 1;
        uint8_t* planes = dplanes + G.plane_base;
        for (int k = 0; k < ncomp_run; ++k) {
            const JpegComp& c = J.c[k];
            const int nb = c.bx * c.by;
            k_jpeg_idct<<<cdiv(nb, 128), 128, 0, st>>>(dcoef + G.coef_base + c.coef_off * 64,
                                                       reinterpret_cast<const uint16_t*>(dblob + G.o_q) + c.tq * 64, c.bx, nb,
                                                       planes + c.plane_off);
            count_launch(h);
        }
        ColorDesc D{};
        D.y = planes + J.c[0].plane_off;
        D.ypitch = J.c[0].bx * 8;
        D.ncomp = G.need_chroma ? 3 : 1;
        D.hx = D.vx = 1;
        if (G.need_chroma) {
            D.cb = planes + J.c[1].plane_off;
            D.cr = planes + J.c[2].plane_off;
            D.cpitch = J.c[1].bx * 8;
            D.cdw = J.c[1].dw;
            D.cdh = J.c[1].dh;
            D.hx = J.hmax / J.c[1].h;
            D.vx = J.vmax / J.c[1].v;
        }
        D.H = J.H; D.W = J.W;
        D.orientation = J.orientation;
        D.bgr = jobs[i].out_bgr;
        D.gray = jobs[i].out_gray;
        k_jpeg_color<<<dim3(cdiv(J.W, 32), cdiv(J.H, 32)), 256, 0, st>>>(D);
        count_launch(h);
    }
    lane.scr_dirty = false;
    CUDA_CHECK(cudaGetLastError());
}

// Decode one JPEG on `lane`.  out_bgr / out_gray: device pointers (either may be null), sized for the oriented image.
// Returns after enqueueing the work on the lane's stream (the pinned staging stays busy until the stream drains).
void jpeg_decode_dev(Handle* h, Lane& lane, const uint8_t* data, size_t n, int ignore_orientation, uint8_t* out_bgr, uint8_t* out_gray,
                     int* outH, int* outW) {
    JpegJob jb{data, n, out_bgr, out_gray, 0, 0};
    jpeg_decode_group_dev(h, lane, &jb, 1, ignore_orientation);
    if (outH) *outH = jb.H;
    if (outW) *outW = jb.W;
}

}  // namespace bbocr
