/*
 * orc_prims.c -- oracle restatement of the un-vendored OpenCV primitives the
 * reference's hot path calls.  TEST INFRASTRUCTURE ONLY
 * (see plf_oracle.h).  Each function names the reference call site it stands for;
 * every one is pinned bit-for-bit against cv2 4.13.0 in tests/test_oracle_vs_cv2.py.
 */
#include "plf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

/* cv::resize(..., INTER_LINEAR) on 8UC1 -- call site src/ORBextractor.cc:1120.
 * 11-bit fixed-point coefficients, int32 horizontal pass, vertical pass with the
 * (>>4, *b >>16, +2 >>2) rounding of OpenCV's 8u path. */
static void linear_coeffs(int ssize, int dsize, int* ofs, short* coef /* 2 per dst */)
{
    double scale = 1.0 / ((double)dsize / (double)ssize);
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        ofs[d] = s;
        coef[2 * d]     = (short)lrintf((1.f - f) * 2048.f);
        coef[2 * d + 1] = (short)lrintf(f * 2048.f);
    }
}

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                          uint8_t* dst, int dw, int dh, size_t dstride)
{
    int* xofs = (int*)malloc(sizeof(int) * (size_t)(dw + dh));
    int* yofs = xofs + dw;
    short* xc = (short*)malloc(sizeof(short) * 2 * (size_t)(dw + dh));
    short* yc = xc + 2 * dw;
    linear_coeffs(sw, dw, xofs, xc);
    linear_coeffs(sh, dh, yofs, yc);
    int* row0 = (int*)malloc(sizeof(int) * 2 * (size_t)dw);
    int* row1 = row0 + dw;
    int have0 = -1, have1 = -1;
    for (int y = 0; y < dh; y++) {
        int sy0 = yofs[y], sy1 = sy0 + 1 < sh ? sy0 + 1 : sh - 1;
        /* horizontal pass of the two needed source rows (cached) */
        if (have1 == sy0) { int* t = row0; row0 = row1; row1 = t; have0 = sy0; have1 = -1; }
        if (have0 != sy0) {
            const uint8_t* s = src + (size_t)sy0 * sstride;
            for (int x = 0; x < dw; x++) {
                int sx = xofs[x], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
                row0[x] = s[sx] * xc[2 * x] + s[sx1] * xc[2 * x + 1];
            }
            have0 = sy0;
        }
        if (have1 != sy1) {
            const uint8_t* s = src + (size_t)sy1 * sstride;
            for (int x = 0; x < dw; x++) {
                int sx = xofs[x], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
                row1[x] = s[sx] * xc[2 * x] + s[sx1] * xc[2 * x + 1];
            }
            have1 = sy1;
        }
        int b0 = yc[2 * y], b1 = yc[2 * y + 1];
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < dw; x++)
            d[x] = (uint8_t)((((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2);
    }
    free(row0 < row1 ? row0 : row1);
    free(xc);
    free(xofs);
}

/* cv::resize(src, dst, Size(), fx, fy, INTER_LINEAR_EXACT) on 8UC1 -- used inside
 * cv::LineSegmentDetector when SCALE != 1 (LSDDetector_custom.cpp:246-262 -> OpenCV).
 * Q8 coefficients, (V + 32768) >> 16. dw,dh must be lrint(sw*fx), lrint(sh*fy). */
static void exact_coeffs(int ssize, int dsize, double inv_scale, int* ofs, int* c1)
{
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; d++) {
        double f = (d + 0.5) * scale - 0.5;
        int s = (int)floor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0; }
        ofs[d] = s;
        c1[d] = (int)floor(f * 256.0 + 0.5);
    }
}

void orc_resize_linear_exact_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                                uint8_t* dst, int dw, int dh, size_t dstride, double fx, double fy)
{
    int* xofs = (int*)malloc(sizeof(int) * 2 * (size_t)(dw + dh));
    int* yofs = xofs + dw;
    int* xc = yofs + dh;
    int* yc = xc + dw;
    exact_coeffs(sw, dw, fx, xofs, xc);
    exact_coeffs(sh, dh, fy, yofs, yc);
    for (int y = 0; y < dh; y++) {
        int sy0 = yofs[y], sy1 = sy0 + 1 < sh ? sy0 + 1 : sh - 1;
        const uint8_t* s0 = src + (size_t)sy0 * sstride;
        const uint8_t* s1 = src + (size_t)sy1 * sstride;
        int cy1 = yc[y], cy0 = 256 - cy1;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < dw; x++) {
            int sx = xofs[x], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            int cx1 = xc[x], cx0 = 256 - cx1;
            int h0 = cx0 * s0[sx] + cx1 * s0[sx1];
            int h1 = cx0 * s1[sx] + cx1 * s1[sx1];
            d[x] = (uint8_t)((cy0 * h0 + cy1 * h1 + 32768) >> 16);
        }
    }
    free(xofs);
}

/* cv::copyMakeBorder(..., BORDER_REFLECT_101) -- src/ORBextractor.cc:1122-1128.
 * dst is (w+2b) x (h+2b). */
void orc_border_reflect101_u8(const uint8_t* src, int w, int h, size_t sstride,
                              uint8_t* dst, int border, size_t dstride)
{
    for (int y = -border; y < h + border; y++) {
        const uint8_t* s = src + (size_t)reflect101(y, h) * sstride;
        uint8_t* d = dst + (size_t)(y + border) * dstride;
        for (int x = -border; x < w + border; x++)
            d[x + border] = s[reflect101(x, w)];
    }
}

/* Q8 Gaussian kernel with error diffusion from the edge to the centre
 * (OpenCV's bit-exact 8U GaussianBlur path). Returns 0 on success. */
int orc_gauss_kernel_q8(int ksize, double sigma, int* q)
{
    if (ksize < 1 || !(ksize & 1)) return -1;
    int n2 = ksize / 2;
    double w[64];
    if (ksize > 63) return -1;
    if (sigma <= 0) sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    double scale2x = -0.125 / (sigma * sigma);
    double sum = 0;
    for (int i = 0, x = 1 - ksize; i < n2; i++, x += 2) {
        double t = exp((double)(x * x) * scale2x);
        w[i] = t;
        sum += t;
    }
    sum = sum * 2 + 1;
    sum = 1.0 / sum;
    double err = 0;
    long tot = 0;
    for (int i = 0; i < n2; i++) {
        double v = w[i] * sum * 256.0 + err;
        long v0 = lrint(v);
        err = v - (double)v0;
        q[i] = (int)v0;
        q[ksize - 1 - i] = (int)v0;
        tot += v0;
    }
    q[n2] = (int)(256 - 2 * tot);
    return 0;
}

/* cv::GaussianBlur on 8UC1, BORDER_REFLECT_101 -- src/ORBextractor.cc:1086 (7x7 s2),
 * binary_descriptor_custom.cpp:358 (5x5 s1), and inside LSD.  Separable Q8,
 * dst = (sum_j q_j * (sum_i q_i * p) + 32768) >> 16. */
void orc_gauss_blur_u8(const uint8_t* src, int w, int h, size_t sstride,
                       uint8_t* dst, size_t dstride, int ksize, double sigma)
{
    int q[64];
    if (orc_gauss_kernel_q8(ksize, sigma, q)) return;
    int r = ksize / 2;
    uint16_t* H = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)w * (size_t)h);
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (size_t)y * sstride;
        uint16_t* hr = H + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            unsigned acc = 0;
            if (x >= r && x + r < w) {
                for (int i = 0; i < ksize; i++) acc += (unsigned)q[i] * s[x + i - r];
            } else {
                for (int i = 0; i < ksize; i++) acc += (unsigned)q[i] * s[reflect101(x + i - r, w)];
            }
            hr[x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; y++) {
        uint8_t* d = dst + (size_t)y * dstride;
        const uint16_t* rows[64];
        for (int j = 0; j < ksize; j++) rows[j] = H + (size_t)reflect101(y + j - r, h) * w;
        for (int x = 0; x < w; x++) {
            unsigned acc = 32768u;
            for (int j = 0; j < ksize; j++) acc += (unsigned)q[j] * rows[j][x];
            d[x] = (uint8_t)(acc >> 16);
        }
    }
    free(H);
}

/* cv::pyrDown(src, dst, Size(w/2, h/2)) -- LSDDetector_custom.cpp:70,
 * binary_descriptor_custom.cpp:366. [1 4 6 4 1]^2, REFLECT_101, (V+128)>>8. */
void orc_pyrdown_u8(const uint8_t* src, int w, int h, size_t sstride,
                    uint8_t* dst, size_t dstride)
{
    int dw = w / 2, dh = h / 2;
    static const int k[5] = {1, 4, 6, 4, 1};
    int* H = (int*)malloc(sizeof(int) * (size_t)dw * (size_t)h);
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (size_t)y * sstride;
        for (int x = 0; x < dw; x++) {
            int acc = 0;
            for (int i = 0; i < 5; i++) acc += k[i] * s[reflect101(2 * x + i - 2, w)];
            H[(size_t)y * dw + x] = acc;
        }
    }
    for (int y = 0; y < dh; y++) {
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < dw; x++) {
            int acc = 128;
            for (int j = 0; j < 5; j++) acc += k[j] * H[(size_t)reflect101(2 * y + j - 2, h) * dw + x];
            d[x] = (uint8_t)(acc >> 8);
        }
    }
    free(H);
}

/* cv::Sobel(img, dst, CV_16SC1, 1,0,3) and (0,1,3) -- binary_descriptor_custom.cpp:395-396 */
void orc_sobel3_s16(const uint8_t* src, int w, int h, size_t sstride, int16_t* dx, int16_t* dy)
{
    for (int y = 0; y < h; y++) {
        const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * sstride;
        const uint8_t* r1 = src + (size_t)y * sstride;
        const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * sstride;
        for (int x = 0; x < w; x++) {
            int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
            dx[(size_t)y * w + x] = (int16_t)((r0[xr] - r0[xl]) + 2 * (r1[xr] - r1[xl]) + (r2[xr] - r2[xl]));
            dy[(size_t)y * w + x] = (int16_t)((r2[xl] + 2 * r2[x] + r2[xr]) - (r0[xl] + 2 * r0[x] + r0[xr]));
        }
    }
}

/* cv::fastAtan2 scalar path (degrees, float32, no FMA) -- src/ORBextractor.cc:103 and
 * inside LSD. */
float orc_fast_atan2(float y, float x)
{
    static const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* cv::FAST(img, kps, th, true) TYPE_9_16 -- src/ORBextractor.cc:809-815.
 * Output in raster order. */
static const int ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static inline int fast_score(const uint8_t* p, const int* off, int th)
{
    /* returns max(A,B) where A = max over 9-arcs of min(d), B = max over arcs of min(-d);
     * corner iff result > th */
    int v = p[0], d[25];
    int any_hi = 0, any_lo = 0;
    /* quick reject (speed only): a 9-long arc contains one pixel of every opposite pair (k, k+8) */
    {
        int cls = 3;
        for (int k = 0; k < 8 && cls; k++) {
            int a = v - p[off[k]], b = v - p[off[k + 8]];
            cls &= ((a > th) | (b > th)) | (((a < -th) | (b < -th)) << 1);
        }
        if (!cls) return 0;
    }
    for (int k = 0; k < 16; k++) {
        d[k] = v - p[off[k]];
        any_hi |= d[k] > th;
        any_lo |= d[k] < -th;
    }
    if (!any_hi && !any_lo) return 0;
    for (int k = 0; k < 9; k++) d[16 + k] = d[k];
    int A = -1000, B = -1000;
    for (int s = 0; s < 16; s++) {
        int mn = d[s], mx = d[s];
        for (int k = 1; k < 9; k++) {
            int t = d[s + k];
            if (t < mn) mn = t;
            if (t > mx) mx = t;
        }
        if (mn > A) A = mn;
        if (-mx > B) B = -mx;
    }
    return A > B ? A : B;
}

int orc_fast9(const uint8_t* img, int w, int h, size_t stride, int th,
              int* xs, int* ys, int* score, int cap)
{
    if (w < 7 || h < 7) return 0;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = ring_dy[k] * (int)stride + ring_dx[k];
    int* sc = (int*)calloc((size_t)w * (size_t)h, sizeof(int));
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int b = fast_score(img + (size_t)y * stride + x, off, th);
            sc[(size_t)y * w + x] = b > th ? b - 1 : 0;
        }
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = sc[(size_t)y * w + x];
            if (s == 0) continue;
            const int* r = sc + (size_t)y * w + x;
            if (s > r[-1] && s > r[1] && s > r[-w - 1] && s > r[-w] && s > r[-w + 1] &&
                s > r[w - 1] && s > r[w] && s > r[w + 1]) {
                if (n < cap) { xs[n] = x; ys[n] = y; score[n] = s; }
                n++;
            }
        }
    free(sc);
    return n;
}
