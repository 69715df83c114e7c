/*
 * orc_lbd.c -- oracle restatement of BinaryDescriptor::compute (LBD) from
 * Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp (:74-116, :217-259,
 * :350-412, :524-687, :1026-1372), of Lineextractor::ComputeLsdWithLbd
 * (src/Lineextractor.cc:112-212) and of the brute-force matcher pieces
 * (src/Linematcher.cc:50-66, :520-541).  TEST INFRASTRUCTURE ONLY (see plf_oracle.h).
 *
 * Float semantics, as the reference's code resolves under g++ / libstdc++ (observed by compiling it: oracle/_ref):
 * unqualified cos / sin / sqrt / round / atan2 on float arguments pick the FLOAT overloads (cosf, sinf, sqrtf, roundf,
 * atan2f: libstdc++'s <math.h> puts std::cos(float) etc. into the global namespace and cvstd.hpp has `using std::sqrt`
 * inside namespace cv), `1 / sqrt(x)` is an int / float division; no FMA contraction.
 * Pinned against the reference's own binary_descriptor_custom.cpp compiled into oracle/_ref: the 72-float
 * descriptor is bit-identical to the -O3 build; the reference's CMake flags (-O3 -march=native) let GCC contract
 * multiply-adds into FMAs, which moves 45 % of the floats by one ulp and none of the binary descriptors on the test images
 * (tests/test_oracle_vs_ref.py asserts both).
 * The std::sort by response (Lineextractor.cc:175) is checked against the compiled reference as well.
 */
#include "plf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NUM_OF_BANDS 9
#define WIDTH_OF_BAND 7

static const int combinations[32][2] = {
    {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
    {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
    {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

static void lbd_one(const orc_keyline* K, const int16_t* pdx, const int16_t* pdy, int realWidth, int realHeight,
                    const double* gaussCoefL, const double* gaussCoefG, float* desVec)
{
    float pgdLBandSum[NUM_OF_BANDS] = {0}, ngdLBandSum[NUM_OF_BANDS] = {0};
    float pgdL2BandSum[NUM_OF_BANDS] = {0}, ngdL2BandSum[NUM_OF_BANDS] = {0};
    float pgdOBandSum[NUM_OF_BANDS] = {0}, ngdOBandSum[NUM_OF_BANDS] = {0};
    float pgdO2BandSum[NUM_OF_BANDS] = {0}, ngdO2BandSum[NUM_OF_BANDS] = {0};
    const short heightOfLSP = WIDTH_OF_BAND * NUM_OF_BANDS;
    const short halfHeight = (short)((heightOfLSP - 1) / 2);
    const short imageWidth = (short)(realWidth - 1), imageHeight = (short)(realHeight - 1);
    const short lengthOfLSP = (short)K->numOfPixels;
    const short halfWidth = (short)((lengthOfLSP - 1) / 2);
    float lineMiddlePointX = (float)(0.5 * (K->sPointInOctaveX + K->ePointInOctaveX));
    float lineMiddlePointY = (float)(0.5 * (K->sPointInOctaveY + K->ePointInOctaveY));
    float dL[2], dO[2];
    dL[0] = cosf(K->angle); /* cos(float) resolves to std::cos(float) = cosf (libstdc++ <math.h> overloads) */
    dL[1] = sinf(K->angle);
    dO[0] = -dL[1];
    dO[1] = dL[0];
    float t0 = -dL[0] * halfWidth, t1 = dL[1] * halfHeight;
    float sCorX0 = t0 + t1 + lineMiddlePointX;
    t0 = -dL[1] * halfWidth; t1 = dL[0] * halfHeight;
    float sCorY0 = t0 - t1 + lineMiddlePointY;

    for (short hID = 0; hID < heightOfLSP; hID++) {
        float sCorX = sCorX0, sCorY = sCorY0;
        float pgdLRowSum = 0, ngdLRowSum = 0, pgdORowSum = 0, ngdORowSum = 0;
        for (short wID = 0; wID < lengthOfLSP; wID++) {
            short tempCor = (short)roundf(sCorX);
            short xCor = (tempCor < 0) ? 0 : (tempCor > imageWidth) ? imageWidth : tempCor;
            tempCor = (short)roundf(sCorY);
            short yCor = (tempCor < 0) ? 0 : (tempCor > imageHeight) ? imageHeight : tempCor;
            short dx = pdx[yCor * realWidth + xCor];
            short dy = pdy[yCor * realWidth + xCor];
            float a0 = dx * dL[0], a1 = dy * dL[1];
            float gDL = a0 + a1;
            a0 = dx * dO[0]; a1 = dy * dO[1];
            float gDO = a0 + a1;
            if (gDL > 0) pgdLRowSum += gDL; else ngdLRowSum -= gDL;
            if (gDO > 0) pgdORowSum += gDO; else ngdORowSum -= gDO;
            sCorX += dL[0];
            sCorY += dL[1];
        }
        sCorX0 -= dL[1];
        sCorY0 += dL[0];
        float coef = (float)gaussCoefG[hID];
        pgdLRowSum = coef * pgdLRowSum;
        ngdLRowSum = coef * ngdLRowSum;
        float pgdL2RowSum = pgdLRowSum * pgdLRowSum;
        float ngdL2RowSum = ngdLRowSum * ngdLRowSum;
        pgdORowSum = coef * pgdORowSum;
        ngdORowSum = coef * ngdORowSum;
        float pgdO2RowSum = pgdORowSum * pgdORowSum;
        float ngdO2RowSum = ngdORowSum * ngdORowSum;

        short bandID = (short)(hID / WIDTH_OF_BAND);
        for (int pass = 0; pass < 3; pass++) {
            int b; float c;
            if (pass == 0) { b = bandID; c = (float)gaussCoefL[hID % WIDTH_OF_BAND + WIDTH_OF_BAND]; }
            else if (pass == 1) { b = bandID - 1; if (b < 0) continue; c = (float)gaussCoefL[hID % WIDTH_OF_BAND + 2 * WIDTH_OF_BAND]; }
            else { b = bandID + 1; if (b >= NUM_OF_BANDS) continue; c = (float)gaussCoefL[hID % WIDTH_OF_BAND]; }
            float cc = c * c, m;
            m = c * pgdLRowSum;   pgdLBandSum[b] += m;
            m = c * ngdLRowSum;   ngdLBandSum[b] += m;
            m = cc * pgdL2RowSum; pgdL2BandSum[b] += m;
            m = cc * ngdL2RowSum; ngdL2BandSum[b] += m;
            m = c * pgdORowSum;   pgdOBandSum[b] += m;
            m = c * ngdORowSum;   ngdOBandSum[b] += m;
            m = cc * pgdO2RowSum; pgdO2BandSum[b] += m;
            m = cc * ngdO2RowSum; ngdO2BandSum[b] += m;
        }
    }

    float invN2 = (float)(1.0 / (WIDTH_OF_BAND * 2.0));
    float invN3 = (float)(1.0 / (WIDTH_OF_BAND * 3.0));
    for (int bandID = 0; bandID < NUM_OF_BANDS; bandID++) {
        float invN = (bandID == 0 || bandID == NUM_OF_BANDS - 1) ? invN2 : invN3;
        int desID = bandID * 8;
        float temp, u, v;
        temp = pgdLBandSum[bandID] * invN; desVec[desID] = temp;
        u = pgdL2BandSum[bandID] * invN; v = temp * temp; desVec[desID + 4] = sqrtf(u - v);
        temp = ngdLBandSum[bandID] * invN; desVec[desID + 1] = temp;
        u = ngdL2BandSum[bandID] * invN; v = temp * temp; desVec[desID + 5] = sqrtf(u - v);
        temp = pgdOBandSum[bandID] * invN; desVec[desID + 2] = temp;
        u = pgdO2BandSum[bandID] * invN; v = temp * temp; desVec[desID + 6] = sqrtf(u - v);
        temp = ngdOBandSum[bandID] * invN; desVec[desID + 3] = temp;
        u = ngdO2BandSum[bandID] * invN; v = temp * temp; desVec[desID + 7] = sqrtf(u - v);
    }
    float tempM = 0, tempS = 0, m;
    for (int i = 0; i < NUM_OF_BANDS * 8; i += 8) {
        for (int k = 0; k < 4; k++) { m = desVec[i + k] * desVec[i + k]; tempM += m; }
        for (int k = 4; k < 8; k++) { m = desVec[i + k] * desVec[i + k]; tempS += m; }
    }
    tempM = 1 / sqrtf(tempM); /* namespace cv: using std::sqrt -> sqrt(float), then int / float */
    tempS = 1 / sqrtf(tempS);
    for (int i = 0; i < NUM_OF_BANDS * 8; i += 8) {
        for (int k = 0; k < 4; k++) desVec[i + k] = desVec[i + k] * tempM;
        for (int k = 4; k < 8; k++) desVec[i + k] = desVec[i + k] * tempS;
    }
    for (int i = 0; i < NUM_OF_BANDS * 8; i++)
        if ((double)desVec[i] > 0.4) desVec[i] = (float)0.4;
    float temp = 0;
    for (int i = 0; i < NUM_OF_BANDS * 8; i++) { m = desVec[i] * desVec[i]; temp += m; }
    temp = 1 / sqrtf(temp);
    for (int i = 0; i < NUM_OF_BANDS * 8; i++) desVec[i] = desVec[i] * temp;
}

void orc_lbd_compute(const uint8_t* img, int w, int h, size_t stride,
                     const orc_keyline* kl, int n, uint8_t* desc, float* fdesc)
{
    if (n <= 0) return;
    double gaussCoefL[WIDTH_OF_BAND * 3], gaussCoefG[NUM_OF_BANDS * WIDTH_OF_BAND];
    {
        double u = (WIDTH_OF_BAND * 3 - 1) / 2;
        double sigma = (WIDTH_OF_BAND * 2 + 1) / 2;
        double invsigma2 = -1 / (2 * sigma * sigma);
        for (int i = 0; i < WIDTH_OF_BAND * 3; i++) { double dis = i - u; gaussCoefL[i] = exp(dis * dis * invsigma2); }
        u = (NUM_OF_BANDS * WIDTH_OF_BAND - 1) / 2;
        sigma = u;
        invsigma2 = -1 / (2 * sigma * sigma);
        for (int i = 0; i < NUM_OF_BANDS * WIDTH_OF_BAND; i++) { double dis = i - u; gaussCoefG[i] = exp(dis * dis * invsigma2); }
    }
    int maxOct = -1;
    for (int i = 0; i < n; i++) if (kl[i].octave > maxOct) maxOct = kl[i].octave;
    int noct = maxOct + 1;
    /* computeGaussianPyramid (:350-370) + computeSobel (:373-398) */
    int16_t** dxs = (int16_t**)calloc((size_t)noct, sizeof(int16_t*));
    int16_t** dys = (int16_t**)calloc((size_t)noct, sizeof(int16_t*));
    int* ws = (int*)calloc((size_t)noct, sizeof(int));
    int* hs = (int*)calloc((size_t)noct, sizeof(int));
    uint8_t* cur = (uint8_t*)malloc((size_t)w * (size_t)h);
    orc_gauss_blur_u8(img, w, h, stride, cur, (size_t)w, 5, 1.0);
    int cw = w, ch = h;
    for (int o = 0; o < noct; o++) {
        if (o > 0) {
            int nw = cw / 2, nh = ch / 2;
            uint8_t* nx = (uint8_t*)malloc((size_t)(nw > 0 ? nw : 1) * (size_t)(nh > 0 ? nh : 1));
            orc_pyrdown_u8(cur, cw, ch, (size_t)cw, nx, (size_t)nw);
            free(cur); cur = nx; cw = nw; ch = nh;
        }
        ws[o] = cw; hs[o] = ch;
        dxs[o] = (int16_t*)malloc(sizeof(int16_t) * (size_t)cw * (size_t)ch);
        dys[o] = (int16_t*)malloc(sizeof(int16_t) * (size_t)cw * (size_t)ch);
        orc_sobel3_s16(cur, cw, ch, (size_t)cw, dxs[o], dys[o]);
    }
    free(cur);
    for (int i = 0; i < n; i++) {
        float dv[NUM_OF_BANDS * 8];
        int o = kl[i].octave;
        lbd_one(&kl[i], dxs[o], dys[o], ws[o], hs[o], gaussCoefL, gaussCoefG, dv);
        if (fdesc) memcpy(fdesc + (size_t)i * 72, dv, sizeof(dv));
        if (desc) {
            for (int c = 0; c < 32; c++) {
                const float* f1 = &dv[8 * combinations[c][0]];
                const float* f2 = &dv[8 * combinations[c][1]];
                unsigned r = 0;
                for (int b = 0; b < 8; b++) if (f1[b] > f2[b]) r += 1u << b;
                desc[(size_t)i * 32 + c] = (uint8_t)r;
            }
        }
    }
    for (int o = 0; o < noct; o++) { free(dxs[o]); free(dys[o]); }
    free(dxs); free(dys); free(ws); free(hs);
}

/* ---- std::sort as libstdc++ runs it (bits/stl_algo.h: introsort with median-of-3 to first, unguarded partition, threshold
 * 16, heapsort past depth 2 * floor(log2 n), final insertion sort).  Lineextractor.cc:175 sorts the lines of an octave by
 * `a.response > b.response` with std::sort, which is NOT stable: lines of equal response come out in an order that only
 * this exact sequence of comparisons and moves defines.  Restated on a permutation p[] of indices with keys k[] (the
 * algorithm only ever compares and moves elements); checked against the real std::sort inside oracle/_ref
 * (ref_std_sort_desc) on tie-heavy and adversarial inputs by tests/test_oracle_vs_ref.py. */
#define SS_LT(a, b) (k[a] > k[b]) /* comp(a, b) of sort_lines_by_response */
static void ss_unguarded_linear_insert(const float* k, int* p, int last)
{
    int val = p[last], next = last - 1;
    while (SS_LT(val, p[next])) { p[last] = p[next]; last = next; --next; }
    p[last] = val;
}
static void ss_insertion_sort(const float* k, int* p, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (SS_LT(p[i], p[first])) {
            int val = p[i];
            memmove(p + first + 1, p + first, sizeof(int) * (size_t)(i - first));
            p[first] = val;
        } else
            ss_unguarded_linear_insert(k, p, i);
    }
}
static void ss_push_heap(const float* k, int* p, int first, int hole, int top, int value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && SS_LT(p[first + parent], value)) { p[first + hole] = p[first + parent]; hole = parent; parent = (hole - 1) / 2; }
    p[first + hole] = value;
}
static void ss_adjust_heap(const float* k, int* p, int first, int hole, int len, int value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (SS_LT(p[first + child], p[first + child - 1])) child--;
        p[first + hole] = p[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        p[first + hole] = p[first + child - 1];
        hole = child - 1;
    }
    ss_push_heap(k, p, first, hole, top, value);
}
static int ss_heap_calls = 0; /* test hook: how often the depth limit sent a range to heapsort */
int orc_std_sort_heap_calls(void) { return ss_heap_calls; }
static void ss_heap_sort(const float* k, int* p, int first, int last) /* __partial_sort(first, last, last) */
{
    int len = last - first;
    ss_heap_calls++;
    if (len >= 2)
        for (int parent = (len - 2) / 2;; parent--) {
            ss_adjust_heap(k, p, first, parent, len, p[first + parent]);
            if (parent == 0) break;
        }
    while (last - first > 1) {
        --last;
        int value = p[last];
        p[last] = p[first];
        ss_adjust_heap(k, p, first, 0, last - first, value);
    }
}
static void ss_introsort_loop(const float* k, int* p, int first, int last, int depth)
{
    while (last - first > 16) {
        if (depth == 0) { ss_heap_sort(k, p, first, last); return; }
        --depth;
        /* __move_median_to_first(first, first + 1, mid, last - 1) */
        int a = first + 1, b = first + (last - first) / 2, c = last - 1, t, m;
        if (SS_LT(p[a], p[b])) m = SS_LT(p[b], p[c]) ? b : SS_LT(p[a], p[c]) ? c : a;
        else m = SS_LT(p[a], p[c]) ? a : SS_LT(p[b], p[c]) ? c : b;
        t = p[first]; p[first] = p[m]; p[m] = t;
        /* __unguarded_partition(first + 1, last, pivot = first) */
        int lo = first + 1, hi = last;
        for (;;) {
            while (SS_LT(p[lo], p[first])) ++lo;
            --hi;
            while (SS_LT(p[first], p[hi])) --hi;
            if (!(lo < hi)) break;
            t = p[lo]; p[lo] = p[hi]; p[hi] = t;
            ++lo;
        }
        ss_introsort_loop(k, p, lo, last, depth);
        last = lo;
    }
}
void orc_std_sort_desc(const float* k, int n, int* p)
{
    for (int i = 0; i < n; i++) p[i] = i;
    if (n == 0) return;
    int lg = 0;
    for (int v = n; v > 1; v >>= 1) lg++;
    ss_introsort_loop(k, p, 0, n, 2 * lg);
    if (n > 16) {
        ss_insertion_sort(k, p, 0, 16);
        for (int i = 16; i != n; ++i) ss_unguarded_linear_insert(k, p, i);
    } else
        ss_insertion_sort(k, p, 0, n);
}

/* Lineextractor::ComputeLsdWithLbd, src/Lineextractor.cc:112-212 */
int orc_line_extract(const orc_line_params* P, const uint8_t* img, int w, int h, size_t stride,
                     orc_keyline* kl, orc_keypoint* mid, uint8_t* desc, int cap)
{
    if (!img || w <= 0 || h <= 0) return 0;
    int dcap = 1 << 16;
    orc_keyline* det = (orc_keyline*)malloc(sizeof(orc_keyline) * (size_t)dcap);
    int nd = orc_lsd_detect_keylines(P, img, w, h, stride, det, dcap);
    if (nd > dcap) nd = dcap;
    /* bucket by octave exactly like :139-159 */
    int* bstart = (int*)calloc((size_t)nd + 2, sizeof(int));
    int nb = 0, octaveIdx = 0;
    bstart[0] = 0;
    for (int i = 0; i < nd; i++) {
        if (det[i].octave != octaveIdx) { bstart[++nb] = i; octaveIdx++; }
    }
    bstart[++nb] = nd;
    int n = 0;
    float* keys = (float*)malloc(sizeof(float) * (size_t)(nd > 0 ? nd : 1));
    int* perm = (int*)malloc(sizeof(int) * (size_t)(nd > 0 ? nd : 1));
    for (int b = 0; b < nb; b++) {
        int cnt = bstart[b + 1] - bstart[b];
        int quota = b < P->nlevels ? orc_line_features_per_level(P, b) : 0;
        if (cnt <= quota) {
            for (int k = 0; k < cnt; k++) { if (n < cap) kl[n] = det[bstart[b] + k]; n++; }
        } else {
            for (int k = 0; k < cnt; k++) keys[k] = det[bstart[b] + k].response;
            orc_std_sort_desc(keys, cnt, perm); /* std::sort(..., sort_lines_by_response()), :175 */
            for (int k = 0; k < quota; k++) { if (n < cap) kl[n] = det[bstart[b] + perm[k]]; n++; }
        }
    }
    free(keys); free(perm); free(bstart); free(det);
    if (n > cap) return -1;
    for (int k = 0; k < n; k++) {
        kl[k].class_id = k;
        if (mid) {
            memset(&mid[k], 0, sizeof(orc_keypoint));
            mid[k].x = (kl[k].startPointX + kl[k].endPointX) / 2;
            mid[k].y = (kl[k].startPointY + kl[k].endPointY) / 2;
            mid[k].octave = kl[k].octave;
            mid[k].angle = -1; mid[k].class_id = -1; /* cv::KeyPoint() defaults */
        }
    }
    if (n > 0 && desc) orc_lbd_compute(img, w, h, stride, kl, n, desc, NULL);
    return n;
}

/* ---- matching ---- */
/* Linematcher::DescriptorDistance (src/Linematcher.cc:50-66) == ORBmatcher's (:1656-1672) */
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t pa, pb;
        memcpy(&pa, a + 4 * i, 4); memcpy(&pb, b + 4 * i, 4);
        unsigned int v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

/* cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2): ascending distance, ties -> lowest trainIdx */
void orc_knn2(const uint8_t* q, int nq, const uint8_t* t, long nt, int32_t* idx, int32_t* dist)
{
    for (int i = 0; i < nq; i++) {
        const uint64_t* a = (const uint64_t*)(q + (size_t)i * 32);
        uint64_t a0, a1, a2, a3;
        memcpy(&a0, a, 8); memcpy(&a1, a + 1, 8); memcpy(&a2, a + 2, 8); memcpy(&a3, a + 3, 8);
        int d0 = 1 << 30, d1 = 1 << 30; long i0 = -1, i1 = -1;
        for (long j = 0; j < nt; j++) {
            uint64_t b0, b1, b2, b3;
            const uint8_t* tb = t + (size_t)j * 32;
            memcpy(&b0, tb, 8); memcpy(&b1, tb + 8, 8); memcpy(&b2, tb + 16, 8); memcpy(&b3, tb + 24, 8);
            int d = __builtin_popcountll(a0 ^ b0) + __builtin_popcountll(a1 ^ b1) +
                    __builtin_popcountll(a2 ^ b2) + __builtin_popcountll(a3 ^ b3);
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
            else if (d < d1) { d1 = d; i1 = j; }
        }
        idx[2 * i] = (int32_t)i0; idx[2 * i + 1] = (int32_t)i1;
        dist[2 * i] = i0 >= 0 ? d0 : -1; dist[2 * i + 1] = i1 >= 0 ? d1 : -1;
    }
}

/* Linematcher::matchNNR (src/Linematcher.cc:520-541). The reference reads matches_[idx][1]
 * even when the train set has < 2 rows (UB); defined here as "no match". */
int orc_match_nnr(const uint8_t* q, int nq, const uint8_t* t, long nt, float nnr, int32_t* matches12)
{
    int nm = 0;
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * 4 * (size_t)(nq > 0 ? nq : 1));
    int32_t* dist = idx + 2 * (size_t)(nq > 0 ? nq : 1);
    orc_knn2(q, nq, t, nt, idx, dist);
    for (int i = 0; i < nq; i++) {
        matches12[i] = -1;
        if (idx[2 * i + 1] < 0) continue;
        if ((float)dist[2 * i] < (float)dist[2 * i + 1] * nnr) { matches12[i] = idx[2 * i]; nm++; }
    }
    free(idx);
    return nm;
}

/* top-2 over a candidate list with the reference's update rule (ORBmatcher::SearchForInitialization,
 * src/ORBmatcher.cc:430-456: `dist < bestDist` -> shift, `else if (dist < bestDist2)`); also tracks which
 * candidate holds the second place (bestLevel2 bookkeeping of SearchByProjection, :88-104). */
void orc_hamming_candidates(const uint8_t* q, int nq, const uint8_t* t, const int32_t* off, const int32_t* cidx,
                            int32_t* bidx, int32_t* bdist, int32_t* cdist)
{
    for (int i = 0; i < nq; i++) {
        int bestDist = 2147483647, bestDist2 = 2147483647, bestIdx = -1, bestIdx2 = -1;
        for (int c = off[i]; c < off[i + 1]; c++) {
            const int dist = orc_descriptor_distance(q + (size_t)i * 32, t + (size_t)cidx[c] * 32);
            if (cdist) cdist[c] = dist;
            if (dist < bestDist) { bestDist2 = bestDist; bestIdx2 = bestIdx; bestDist = dist; bestIdx = cidx[c]; }
            else if (dist < bestDist2) { bestDist2 = dist; bestIdx2 = cidx[c]; }
        }
        bidx[2 * i] = bestIdx; bidx[2 * i + 1] = bestIdx2;
        bdist[2 * i] = bestIdx >= 0 ? bestDist : -1; bdist[2 * i + 1] = bestIdx2 >= 0 ? bestDist2 : -1;
    }
}
