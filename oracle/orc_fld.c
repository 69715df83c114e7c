/*
 * orc_fld.c -- CPU oracle, FLD branch: the OpenCV primitives it needs (cv::Canny, cv::fitLine DIST_L2).
 *
 * TEST INFRASTRUCTURE ONLY (see plf_oracle.h).  Both models are pinned bit-for-bit against cv2 4.13 by
 * tests/test_oracle_vs_cv2.py; the reference's own FLD logic (src/Lineextractor.cc:242-336, 413-980) is NOT restated
 * here -- it is compiled unmodified into oracle/_ref/libref.so on top of these primitives.
 */
#include "plf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

/* cv::Canny(src, edges, th1, th2, apertureSize = 3, L2gradient = false) on 8-bit single-channel images:
 * Sobel 3x3 with BORDER_REPLICATE, L1 magnitude, integer tan(22.5) sector test (Q15), hysteresis over 8 neighbours.
 * edges: dense w*h, 255 / 0. */
void orc_canny_u8(const uint8_t* src, int w, int h, size_t stride, double th1, double th2, uint8_t* edges)
{
    int low = (int)floor(th1), high = (int)floor(th2);
    if (low > high) { int t = low; low = high; high = t; }
    size_t n = (size_t)w * (size_t)h;
    int16_t* dx = (int16_t*)malloc(n * sizeof(int16_t));
    int16_t* dy = (int16_t*)malloc(n * sizeof(int16_t));
    /* magnitude with a zero frame of one pixel */
    int mw = w + 2;
    int* mag = (int*)calloc((size_t)mw * (size_t)(h + 2), sizeof(int));
    for (int y = 0; y < h; y++) {
        const uint8_t* r0 = src + (size_t)clampi(y - 1, 0, h - 1) * stride;
        const uint8_t* r1 = src + (size_t)y * stride;
        const uint8_t* r2 = src + (size_t)clampi(y + 1, 0, h - 1) * stride;
        for (int x = 0; x < w; x++) {
            int xm = clampi(x - 1, 0, w - 1), xp = clampi(x + 1, 0, w - 1);
            int gx = (r0[xp] + 2 * r1[xp] + r2[xp]) - (r0[xm] + 2 * r1[xm] + r2[xm]);
            int gy = (r2[xm] + 2 * r2[x] + r2[xp]) - (r0[xm] + 2 * r0[x] + r0[xp]);
            dx[(size_t)y * w + x] = (int16_t)gx;
            dy[(size_t)y * w + x] = (int16_t)gy;
            mag[(size_t)(y + 1) * mw + x + 1] = abs(gx) + abs(gy);
        }
    }
    /* map: 1 = not an edge, 0 = candidate, 2 = edge; framed with 1 */
    uint8_t* map = (uint8_t*)malloc((size_t)mw * (size_t)(h + 2));
    memset(map, 1, (size_t)mw * (size_t)(h + 2));
    size_t* stack = (size_t*)malloc(n * sizeof(size_t) + sizeof(size_t));
    size_t sp = 0;
    const int TG22 = (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5);
    for (int y = 0; y < h; y++) {
        const int* ma = mag + (size_t)(y + 1) * mw + 1;
        const int* mp = ma - mw;
        const int* mn = ma + mw;
        uint8_t* pm = map + (size_t)(y + 1) * mw + 1;
        for (int x = 0; x < w; x++) {
            int m = ma[x];
            if (m <= low) continue;
            int xs = dx[(size_t)y * w + x], ys = dy[(size_t)y * w + x];
            int ax = abs(xs), ay = abs(ys) << 15;
            int tg22x = ax * TG22;
            int keep;
            if (ay < tg22x)
                keep = m > ma[x - 1] && m >= ma[x + 1];
            else {
                int tg67x = tg22x + (ax << 16);
                if (ay > tg67x)
                    keep = m > mp[x] && m >= mn[x];
                else {
                    int s = (xs ^ ys) < 0 ? -1 : 1;
                    keep = m > mp[x - s] && m > mn[x + s];
                }
            }
            if (!keep) continue;
            if (m > high) { pm[x] = 2; stack[sp++] = (size_t)(y + 1) * mw + x + 1; }
            else pm[x] = 0;
        }
    }
    static const int nb[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
    while (sp) {
        size_t p = stack[--sp];
        for (int k = 0; k < 8; k++) {
            size_t q = p + (size_t)((long)nb[k][1] * mw + nb[k][0]);
            if (map[q] == 0) { map[q] = 2; stack[sp++] = q; }
        }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) edges[(size_t)y * w + x] = map[(size_t)(y + 1) * mw + x + 1] == 2 ? 255 : 0;
    free(dx); free(dy); free(mag); free(map); free(stack);
}

/* cv::fitLine(points (int x, y pairs), line, DIST_L2, 0, reps, aeps): imgproc/linefit.cpp fitLine2D_wods --
 * moments in double (the float products are exact for pixel coordinates), principal axis angle in float. */
void orc_fit_line_l2(const int32_t* pts, int count, float* line)
{
    double x = 0, y = 0, x2 = 0, y2 = 0, xy = 0, w;
    for (int i = 0; i < count; i++) {
        float px = (float)pts[2 * i], py = (float)pts[2 * i + 1];
        x += px;
        y += py;
        x2 += px * px;
        y2 += py * py;
        xy += px * py;
    }
    w = (float)count;
    x /= w; y /= w; x2 /= w; y2 /= w; xy /= w;
    double dx2 = x2 - x * x, dy2 = y2 - y * y, dxy = xy - x * y;
    float t = (float)atan2(2 * dxy, dx2 - dy2) / 2;
    line[0] = cosf(t); /* cv2 4.13 evaluates cos / sin of the float angle in single precision (pinned: tests/test_oracle_vs_cv2.py) */
    line[1] = sinf(t);
    line[2] = (float)x;
    line[3] = (float)y;
}
