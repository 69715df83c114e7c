/*
 * orc_lsd.c -- oracle restatement of cv::LineSegmentDetector::detect with refine = 0
 * (LSD_REFINE_NONE), which the reference reaches through
 * Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:246-264, and of
 * LSDDetectorC::computeGaussianPyramid/detectImpl (:56-73, :227-324).
 * TEST INFRASTRUCTURE ONLY (see plf_oracle.h).  The LSD arithmetic itself lives in
 * un-vendored OpenCV; this restates its observable behaviour
 * and is pinned against cv2 4.13.0's createLineSegmentDetector in
 * tests/test_oracle_vs_cv2.py (identical ordered Vec4f sequences).
 */
#include "plf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NOTDEF (-1024.0)
#define PI_D 3.14159265358979323846
#define M_3_2_PI (3 * PI_D / 2)
#define M_2__PI (2 * PI_D)
#define DEG_TO_RADS (PI_D / 180)

typedef struct { int x, y; } ipt;

static inline int is_aligned(const double* angles, int w, int h, int x, int y, double theta, double prec)
{
    if (x < 0 || y < 0 || x >= w || y >= h) return 0;
    double a = angles[(size_t)y * w + x];
    if (a == NOTDEF) return 0;
    double n_theta = theta - a;
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > M_3_2_PI) {
        n_theta -= M_2__PI;
        if (n_theta < 0) n_theta = -n_theta;
    }
    return n_theta <= prec;
}

static inline double angle_diff(double a, double b)
{
    double diff = a - b;
    while (diff <= -PI_D) diff += M_2__PI;
    while (diff > PI_D) diff -= M_2__PI;
    return fabs(diff);
}

int orc_lsd_detect(const uint8_t* img, int iw, int ih, size_t stride,
                   double SCALE, double SIGMA_SCALE, double QUANT, double ANG_TH, int N_BINS,
                   float* lines, int cap)
{
    const double prec = PI_D * ANG_TH / 180;
    const double p = ANG_TH / 180;
    const double rho = QUANT / sin(prec);
    int w = iw, h = ih;
    uint8_t* scaled = NULL;
    const uint8_t* S; size_t sstride;
    if (SCALE != 1) {
        const double sigma = (SCALE < 1) ? (SIGMA_SCALE / SCALE) : SIGMA_SCALE;
        const double sprec = 3;
        const unsigned int hh = (unsigned int)(ceil(sigma * sqrt(2 * sprec * log(10.0))));
        int ksize = 1 + 2 * (int)hh;
        uint8_t* g = (uint8_t*)malloc((size_t)iw * (size_t)ih);
        orc_gauss_blur_u8(img, iw, ih, stride, g, (size_t)iw, ksize, sigma);
        w = (int)lrint(iw * SCALE);
        h = (int)lrint(ih * SCALE);
        scaled = (uint8_t*)malloc((size_t)w * (size_t)h);
        orc_resize_linear_exact_u8(g, iw, ih, (size_t)iw, scaled, w, h, (size_t)w, SCALE, SCALE);
        free(g);
        S = scaled; sstride = (size_t)w;
    } else {
        S = img; sstride = stride;
    }
    size_t npx = (size_t)w * (size_t)h;
    double* angles = (double*)malloc(sizeof(double) * npx);
    double* modgrad = (double*)calloc(npx, sizeof(double));
    float* cosf_ = NULL; (void)cosf_;
    /* ll_angle */
    for (int x = 0; x < w; x++) angles[(size_t)(h - 1) * w + x] = NOTDEF;
    for (int y = 0; y < h; y++) angles[(size_t)y * w + (w - 1)] = NOTDEF;
    double max_grad = -1;
    for (int y = 0; y < h - 1; y++) {
        const uint8_t* r0 = S + (size_t)y * sstride;
        const uint8_t* r1 = S + (size_t)(y + 1) * sstride;
        for (int x = 0; x < w - 1; x++) {
            int DA = r1[x + 1] - r0[x];
            int BC = r0[x + 1] - r1[x];
            int gx = DA + BC, gy = DA - BC;
            double norm = sqrt((gx * gx + gy * gy) / 4.0);
            modgrad[(size_t)y * w + x] = norm;
            if (norm <= rho) {
                angles[(size_t)y * w + x] = NOTDEF;
            } else {
                angles[(size_t)y * w + x] = (double)orc_fast_atan2((float)gx, (float)(-gy)) * DEG_TO_RADS;
                if (norm > max_grad) max_grad = norm;
            }
        }
    }
    /* stable descending-bin order (counting sort), raster order inside a bin */
    double bin_coef = (max_grad > 0) ? (double)(N_BINS - 1) / max_grad : 0;
    size_t nord = (size_t)(w - 1) * (size_t)(h - 1);
    int* bins = (int*)malloc(sizeof(int) * (nord ? nord : 1));
    size_t* hist = (size_t*)calloc((size_t)N_BINS + 1, sizeof(size_t));
    {
        size_t k = 0;
        for (int y = 0; y < h - 1; y++)
            for (int x = 0; x < w - 1; x++, k++) {
                int b = (int)(modgrad[(size_t)y * w + x] * bin_coef);
                if (b < 0) b = 0;
                if (b >= N_BINS) b = N_BINS - 1;
                bins[k] = b;
                hist[b]++;
            }
    }
    size_t* start = (size_t*)malloc(sizeof(size_t) * (size_t)N_BINS);
    {
        size_t acc = 0;
        for (int b = N_BINS - 1; b >= 0; b--) { start[b] = acc; acc += hist[b]; }
    }
    ipt* ordered = (ipt*)malloc(sizeof(ipt) * (nord ? nord : 1));
    {
        size_t k = 0;
        for (int y = 0; y < h - 1; y++)
            for (int x = 0; x < w - 1; x++, k++) {
                ipt q = {x, y};
                ordered[start[bins[k]]++] = q;
            }
    }
    free(bins); free(hist); free(start);

    const double LOG_NT = 5 * (log10((double)w) + log10((double)h)) / 2 + log10(11.0);
    const size_t min_reg_size = (size_t)(-LOG_NT / log10(p));
    uint8_t* used = (uint8_t*)calloc(npx, 1);
    ipt* reg = (ipt*)malloc(sizeof(ipt) * npx);
    int nlines = 0;

    for (size_t i = 0; i < nord; i++) {
        ipt s = ordered[i];
        size_t si = (size_t)s.y * w + s.x;
        if (used[si] || angles[si] == NOTDEF) continue;
        /* region_grow */
        double reg_angle = angles[si];
        size_t nreg = 0;
        reg[nreg++] = s;
        float sumdx = (float)cos(reg_angle);
        float sumdy = (float)sin(reg_angle);
        used[si] = 1;
        for (size_t r = 0; r < nreg; r++) {
            ipt rp = reg[r];
            int xx_min = rp.x - 1 > 0 ? rp.x - 1 : 0, xx_max = rp.x + 1 < w - 1 ? rp.x + 1 : w - 1;
            int yy_min = rp.y - 1 > 0 ? rp.y - 1 : 0, yy_max = rp.y + 1 < h - 1 ? rp.y + 1 : h - 1;
            for (int yy = yy_min; yy <= yy_max; ++yy)
                for (int xx = xx_min; xx <= xx_max; ++xx) {
                    size_t qi = (size_t)yy * w + xx;
                    if (!used[qi] && is_aligned(angles, w, h, xx, yy, reg_angle, prec)) {
                        double angle = angles[qi];
                        used[qi] = 1;
                        ipt q = {xx, yy};
                        reg[nreg++] = q;
                        /* float cosine / sine of the float-cast angle: correctly rounded float of
                         * the double result (equals cosf/sinf except in ~1e-9 of inputs) */
                        sumdx += (float)cos((double)(float)angle);
                        sumdy += (float)sin((double)(float)angle);
                        reg_angle = (double)orc_fast_atan2(sumdy, sumdx) * DEG_TO_RADS;
                    }
                }
        }
        if (nreg < min_reg_size) continue;
        /* region2rect */
        double x = 0, y = 0, sum = 0;
        for (size_t r = 0; r < nreg; r++) {
            double weight = modgrad[(size_t)reg[r].y * w + reg[r].x];
            x += (double)reg[r].x * weight;
            y += (double)reg[r].y * weight;
            sum += weight;
        }
        x /= sum; y /= sum;
        double Ixx = 0, Iyy = 0, Ixy = 0;
        for (size_t r = 0; r < nreg; r++) {
            double weight = modgrad[(size_t)reg[r].y * w + reg[r].x];
            double dx = (double)reg[r].x - x, dy = (double)reg[r].y - y;
            Ixx += dy * dy * weight;
            Iyy += dx * dx * weight;
            Ixy -= dx * dy * weight;
        }
        double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
        double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)orc_fast_atan2((float)(lambda - Ixx), (float)Ixy)
                                               : (double)orc_fast_atan2((float)Ixy, (float)(lambda - Iyy));
        theta *= DEG_TO_RADS;
        if (angle_diff(theta, reg_angle) > prec) theta += PI_D;
        double dx = cos(theta), dy = sin(theta);
        double l_min = 0, l_max = 0;
        for (size_t r = 0; r < nreg; r++) {
            double regdx = (double)reg[r].x - x, regdy = (double)reg[r].y - y;
            double l = regdx * dx + regdy * dy;
            if (l > l_max) l_max = l;
            else if (l < l_min) l_min = l;
        }
        double x1 = x + l_min * dx, y1 = y + l_min * dy, x2 = x + l_max * dx, y2 = y + l_max * dy;
        x1 += 0.5; y1 += 0.5; x2 += 0.5; y2 += 0.5;
        if (SCALE != 1) { x1 /= SCALE; y1 /= SCALE; x2 /= SCALE; y2 /= SCALE; }
        if (nlines < cap) {
            lines[4 * nlines + 0] = (float)x1; lines[4 * nlines + 1] = (float)y1;
            lines[4 * nlines + 2] = (float)x2; lines[4 * nlines + 3] = (float)y2;
        }
        nlines++;
    }
    free(angles); free(modgrad); free(ordered); free(used); free(reg); free(scaled);
    return nlines;
}

/* Lineextractor ctor feature split, src/Lineextractor.cc:54-66 */
int orc_line_features_per_level(const orc_line_params* p, int level)
{
    float factor = (float)(1.0f / p->scale);
    float nDesired = (float)(p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)p->nlevels)));
    int sum = 0;
    for (int l = 0; l < p->nlevels - 1; l++) {
        int v = (int)lrintf(nDesired);
        if (l == level) return v;
        sum += v;
        nDesired *= factor;
    }
    return p->nfeatures - sum > 0 ? p->nfeatures - sum : 0;
}

/* LSDDetectorC::detect(image, keylines, scale=2, numOctaves, opts),
 * LSDDetector_custom.cpp:56-73 (pyramid), :227-324 (detectImpl) */
int orc_lsd_detect_keylines(const orc_line_params* P, const uint8_t* img, int w, int h, size_t stride,
                            orc_keyline* kl, int cap)
{
    int n = 0, class_counter = -1;
    uint8_t* cur = (uint8_t*)malloc((size_t)w * (size_t)h);
    for (int y = 0; y < h; y++) memcpy(cur + (size_t)y * w, img + (size_t)y * stride, (size_t)w);
    int cw = w, ch = h;
    int lcap = 1 << 16;
    float* lines = (float*)malloc(sizeof(float) * 4 * (size_t)lcap);
    for (int oct = 0; oct < P->nlevels; oct++) {
        if (oct > 0) {
            int nw = cw / 2, nh = ch / 2;
            uint8_t* nx = (uint8_t*)malloc((size_t)(nw > 0 ? nw : 1) * (size_t)(nh > 0 ? nh : 1));
            orc_pyrdown_u8(cur, cw, ch, (size_t)cw, nx, (size_t)nw);
            free(cur); cur = nx; cw = nw; ch = nh;
        }
        int nl = orc_lsd_detect(cur, cw, ch, (size_t)cw, P->scale, P->sigma_scale, P->quant, P->ang_th,
                                P->n_bins, lines, lcap);
        if (nl > lcap) nl = lcap;
        float octaveScale = (float)pow((double)2.0f, (double)oct);
        for (int k = 0; k < nl; k++) {
            float e[4] = {lines[4 * k], lines[4 * k + 1], lines[4 * k + 2], lines[4 * k + 3]};
            /* checkLineExtremes (:76-102) */
            if (e[0] < 0) e[0] = 0;
            if (e[0] >= cw) e[0] = (float)cw - 1.0f;
            if (e[2] < 0) e[2] = 0;
            if (e[2] >= cw) e[2] = (float)cw - 1.0f;
            if (e[1] < 0) e[1] = 0;
            if (e[1] >= ch) e[1] = (float)ch - 1.0f;
            if (e[3] < 0) e[3] = 0;
            if (e[3] >= ch) e[3] = (float)ch - 1.0f;
            float d02 = e[0] - e[2], d13 = e[1] - e[3];
            double length = (float)sqrt((double)d02 * (double)d02 + (double)d13 * (double)d13);
            if (!(length > P->min_line_length)) continue;
            orc_keyline K;
            K.startPointX = e[0] * octaveScale; K.startPointY = e[1] * octaveScale;
            K.endPointX = e[2] * octaveScale;   K.endPointY = e[3] * octaveScale;
            K.sPointInOctaveX = e[0]; K.sPointInOctaveY = e[1];
            K.ePointInOctaveX = e[2]; K.ePointInOctaveY = e[3];
            K.lineLength = (float)length;
            /* cv::LineIterator(img, Point(cvRound), Point(cvRound)).count, 8-connected */
            int ax = (int)lrintf(e[0]), ay = (int)lrintf(e[1]), bx = (int)lrintf(e[2]), by = (int)lrintf(e[3]);
            int adx = abs(bx - ax), ady = abs(by - ay);
            K.numOfPixels = (adx > ady ? adx : ady) + 1;
            K.angle = atan2f(K.endPointY - K.startPointY, K.endPointX - K.startPointX); /* atan2(float, float) = atan2f */
            K.class_id = ++class_counter;
            K.octave = oct;
            K.size = (K.endPointX - K.startPointX) * (K.endPointY - K.startPointY);
            K.response = K.lineLength / (float)(cw > ch ? cw : ch);
            K.pt_x = (K.endPointX + K.startPointX) / 2;
            K.pt_y = (K.endPointY + K.startPointY) / 2;
            if (n < cap) kl[n] = K;
            n++;
        }
    }
    free(lines); free(cur);
    return n;
}
