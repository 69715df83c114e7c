// plf_line_kernels.cuh -- device kernels of the line path:
// LSD (cv::LineSegmentDetector, refine = 0, as reached from LSDDetector_custom.cpp:246-264) and LBD
// (binary_descriptor_custom.cpp:350-398, :1026-1372).
//
// LSD region growing is sequential in the reference (seeds in descending gradient-bin order, a shared
// `used` map).  Regions can only contain pixels whose gradient angle is defined, so the 8-connected
// components of the "defined" mask never interact: the kernels label those components, sort the seeds by
// (component, bin descending, raster) and grow each component with one thread, as-if-sequentially, which
// reproduces the reference's regions exactly while thousands of components (x frames) run in parallel.
#pragma once
#include "plf_common.cuh"
#include "plf_stdsort.cuh"

#define LSD_NOTDEF (-1024.0f)
#define LSD_USED (-2048.0f)
#define LSD_PI 3.14159265358979323846
#define LSD_D2R (LSD_PI / 180)
#define LSD_3_2_PI (3 * LSD_PI / 2)
#define LSD_2PI (2 * LSD_PI)

// key layout (64 bit).  Every kernel receives one packed int `kb` = KB | BB << 8: KB = bits of a raster index of the
// scaled octave (<= 22), BB = bits of the bin field (ceil(log2(n_bins)), <= 12):
//   frame | component root [KB bits] | n_bins-1-bin [BB bits] | raster index [KB bits]
// Keeping the fields as narrow as the problem allows saves radix-sort passes (56 instead of 58 bits at 752x480 with
// 1024 bins and 256 frames: 7 passes instead of 8).
#define LSD_KB (kb & 0xff)
#define LSD_BB (kb >> 8)
#define LSD_KEY_IDX(k) ((int)((k) & ((1ull << LSD_KB) - 1)))
#define LSD_KEY_TAG(k) ((k) >> (LSD_KB + LSD_BB))
#define LSD_KEY_FRAME(k) ((int)((k) >> (2 * LSD_KB + LSD_BB)))
#define LSD_KEY_BIN(k) ((int)(((k) >> LSD_KB) & ((1ull << LSD_BB) - 1)))

struct GaussQ8 { int ksize; int q[15]; };

// ---------------- generic separable Q8 Gaussian (cv::GaussianBlur 8U, REFLECT_101) ----------------
#define GB_TW 128
#define GB_TH 32
__global__ void __launch_bounds__(256)
k_gauss_q8(const uint8_t* __restrict__ src, size_t sframe, int spitch, uint8_t* __restrict__ dst, size_t dframe, int dpitch,
           int w, int h, GaussQ8 k)
{
    __shared__ uint8_t tile[(GB_TH + 14) * (GB_TW + 16)];
    __shared__ unsigned short hbuf[(GB_TH + 14) * GB_TW];
    const int r = k.ksize >> 1;
    const int tx0 = blockIdx.x * GB_TW, ty0 = blockIdx.y * GB_TH;
    const uint8_t* s = src + (size_t)blockIdx.z * sframe;
    uint8_t* d = dst + (size_t)blockIdx.z * dframe;
    const int tid = threadIdx.x;
    const int TWP = GB_TW + 16, cols = GB_TW + 2 * r, rows = GB_TH + 2 * r;
    for (int i = tid; i < rows * cols; i += 256) {
        int ry = i / cols, rx = i - ry * cols;
        int sy = plf_reflect101(ty0 + ry - r, h), sx = plf_reflect101(tx0 + rx - r, w);
        tile[ry * TWP + rx] = s[(size_t)sy * spitch + sx];
    }
    __syncthreads();
    for (int i = tid; i < rows * GB_TW; i += 256) {
        int ry = i / GB_TW, rx = i - ry * GB_TW;
        const uint8_t* t = &tile[ry * TWP + rx];
        unsigned v = 0;
        for (int j = 0; j < k.ksize; j++) v += (unsigned)k.q[j] * t[j];
        hbuf[i] = (unsigned short)v;
    }
    __syncthreads();
    for (int i = tid; i < GB_TH * GB_TW; i += 256) {
        int ry = i / GB_TW, rx = i - ry * GB_TW;
        int x = tx0 + rx, y = ty0 + ry;
        if (x < w && y < h) {
            unsigned v = 32768u;
            for (int j = 0; j < k.ksize; j++) v += (unsigned)k.q[j] * hbuf[(ry + j) * GB_TW + rx];
            d[(size_t)y * dpitch + x] = (uint8_t)(v >> 16);
        }
    }
}

// 5- and 7-tap kernels (everything the reference's settings reach): register sliding window, see plf_blur_strip.
// block = (32, 4): lane = 4-px column strip, threadIdx.y = GS_ROWS-row band.
#define GS_ROWS 35   // a multiple of both window lengths (5 and 7): no partially used ring turns
template <int R>
__global__ void __launch_bounds__(128)
k_gauss_strip(const uint8_t* __restrict__ src, size_t sframe, int spitch, uint8_t* __restrict__ dst, size_t dframe, int dpitch,
              int w, int h, BlurTaps taps, int F, int ncx_int)
{
    const int strip = plf_strip_of(blockIdx.x, threadIdx.x, w, 1, F, ncx_int);   // interior strips first, edge strips in the last CTA column
    const int y0 = (blockIdx.y * 4 + threadIdx.y) * GS_ROWS;
    if (strip < 0 || y0 >= h) return;
    plf_blur_strip<R>(src + (size_t)blockIdx.z * sframe, spitch, dst + (size_t)blockIdx.z * dframe, dpitch, w, h, 4 * strip, y0, GS_ROWS, taps);
}

// ---------------- cv::resize INTER_LINEAR_EXACT; tab: .x = offset, .y = c1 (Q8) ----------------
// block = (32, 8): a thread produces four adjacent pixels of RX_ROWS rows (8 apart), one packed 32-bit store per row.  The
// column table entries are read once per thread.  Fast path (aligned source rows, the <= 8 source bytes of the four pixels
// inside two aligned words -- always the case when up-scaling): two 32-bit loads per source row, the pair of neighbouring
// bytes of every pixel by one PRMT, the horizontal blend cx0 * a + cx1 * b by one DP2A.  The byte gathers this replaces were
// 21 % of the kernel's stall samples (16 single-byte loads per thread and row).
#define RX_ROWS 4
__global__ void __launch_bounds__(256)
k_resize_exact(const uint8_t* __restrict__ src, size_t sframe, int spitch, int sw, int sh,
               uint8_t* __restrict__ dst, size_t dframe, int dpitch, int dw, int dh,
               const int2* __restrict__ xtab, const int2* __restrict__ ytab)
{
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (x0 >= dw) return;
    const uint8_t* s = src + (size_t)blockIdx.z * sframe;
    int sx0[4], sx1[4];
    unsigned coef[4], sel[4];
    int base = 0;
    bool fast = ((((size_t)s) | (size_t)spitch) & 3) == 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int2 tx = xtab[min(x0 + k, dw - 1)];
        sx0[k] = tx.x; sx1[k] = min(tx.x + 1, sw - 1);
        coef[k] = (unsigned)(256 - tx.y) | ((unsigned)tx.y << 16);          // cx0 | cx1 << 16
        if (k == 0) base = sx0[0] & ~3;
        fast = fast && sx0[k] >= base && sx1[k] - base <= 7;
        sel[k] = (unsigned)((sx0[k] - base) & 7) | ((unsigned)((sx1[k] - base) & 7) << 4);
    }
    fast = fast && base + 8 <= ((sw + 3) & ~3) && base + 8 <= spitch;       // the second word lies inside the row's storage
    const bool wide = x0 + 3 < dw && ((((size_t)dst) | (size_t)dpitch | dframe) & 3) == 0;
#pragma unroll
    for (int rr = 0; rr < RX_ROWS; rr++) {
        const int y = (blockIdx.y * RX_ROWS + rr) * 8 + threadIdx.y;
        if (y >= dh) break;
        const int2 ty = ytab[y];
        const int sy0 = ty.x, sy1 = min(sy0 + 1, sh - 1);
        const int cy1 = ty.y, cy0 = 256 - cy1;
        const uint8_t* r0 = s + (size_t)sy0 * spitch;
        const uint8_t* r1 = s + (size_t)sy1 * spitch;
        unsigned out = 0;
        if (fast) {
            const unsigned* p0 = (const unsigned*)(r0 + base);      // base, the row pitch and the frame base are multiples of 4
            const unsigned* p1 = (const unsigned*)(r1 + base);
            const uint2 a = make_uint2(p0[0], p0[1]), b = make_uint2(p1[0], p1[1]);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int h0 = (int)__dp2a_lo(coef[k], __byte_perm(a.x, a.y, sel[k]), 0u);
                const int h1 = (int)__dp2a_lo(coef[k], __byte_perm(b.x, b.y, sel[k]), 0u);
                out |= (unsigned)((cy0 * h0 + cy1 * h1 + 32768) >> 16) << (8 * k);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int cx1 = (int)(coef[k] >> 16), cx0 = (int)(coef[k] & 0xffffu);
                const int h0 = cx0 * r0[sx0[k]] + cx1 * r0[sx1[k]];
                const int h1 = cx0 * r1[sx0[k]] + cx1 * r1[sx1[k]];
                out |= (unsigned)((cy0 * h0 + cy1 * h1 + 32768) >> 16) << (8 * k);
            }
        }
        uint8_t* d = dst + (size_t)blockIdx.z * dframe + (size_t)y * dpitch + x0;
        if (wide) *(unsigned*)d = out;
        else
            for (int k = 0; k < 4 && x0 + k < dw; k++) d[k] = (uint8_t)(out >> (8 * k));
    }
}

// 12 bytes x0-4 .. x0+7 of an image row as three words (aligned fast path, or gathered with REFLECT_101)
__device__ __forceinline__ void row_words3(const uint8_t* __restrict__ rp, int x0, int w, bool fastx, unsigned& w0, unsigned& w1, unsigned& w2)
{
    if (fastx) {
        const unsigned* p = (const unsigned*)(rp + x0 - 4);
        w0 = p[0]; w1 = p[1]; w2 = p[2];
    } else {
        unsigned b[12];
#pragma unroll
        for (int i = 0; i < 12; i++) b[i] = rp[plf_reflect101(x0 - 4 + i, w)];
        w0 = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
        w1 = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
        w2 = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
    }
}
__device__ __forceinline__ int reflect_once(int p, int n)   // valid for -n < p < 2n - 1
{
    p = p < 0 ? -p : p;
    return p >= n ? 2 * (n - 1) - p : p;
}

// ---------------- cv::pyrDown to (w/2, h/2) ----------------
// [1 4 6 4 1]^2 / 256, REFLECT_101, rounding (v + 128) >> 8.  A thread produces a 4-px wide strip of PD_ROWS output
// rows: per input row it loads the 16 bytes 2*x0-4 .. 2*x0+11 (four aligned words), forms the four horizontal sums
// and keeps the last five rows of sums in registers (two new input rows per output row).
#define PD_ROWS 15
__device__ __forceinline__ void pyrdown_hrow(const uint8_t* __restrict__ rp, int sx0, int w, bool fastx, int (&out)[4])
{
    unsigned b[16];
    if (fastx) {
        const unsigned* pw = (const unsigned*)(rp + sx0 - 4);
        const unsigned ww[4] = {pw[0], pw[1], pw[2], pw[3]};
#pragma unroll
        for (int i = 0; i < 16; i++) b[i] = (ww[i >> 2] >> (8 * (i & 3))) & 0xffu;
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) b[i] = (i >= 2 && i <= 12) ? rp[plf_reflect101(sx0 - 4 + i, w)] : 0u;
    }
#pragma unroll
    for (int j = 0; j < 4; j++)   // source column 2 * (x0 + j) is byte 4 + 2j
        out[j] = (int)(b[2 + 2 * j] + b[6 + 2 * j] + 4u * (b[3 + 2 * j] + b[5 + 2 * j]) + 6u * b[4 + 2 * j]);
}
__global__ void __launch_bounds__(128)
k_pyrdown(const uint8_t* __restrict__ src, size_t sframe, int spitch, int w, int h,
          uint8_t* __restrict__ dst, size_t dframe, int dpitch, int F, int ncx_int)
{
    const int dw = w >> 1, dh = h >> 1;
    const int strip = plf_strip_of(blockIdx.x, threadIdx.x, dw, 1, F, ncx_int);   // interior strips first, edge strips in the last CTA column
    const int x0 = 4 * strip, y0 = (blockIdx.y * 4 + threadIdx.y) * PD_ROWS;
    if (strip < 0 || x0 >= dw || y0 >= dh) return;
    const uint8_t* s = src + (size_t)blockIdx.z * sframe;
    uint8_t* d = dst + (size_t)blockIdx.z * dframe;
    const int sx0 = 2 * x0;
    const bool fastx = ((((size_t)s) | (size_t)spitch) & 3) == 0 && sx0 >= 4 && sx0 + 12 <= w;   // four aligned words
    const bool fullw = ((((size_t)d) | (size_t)dpitch) & 3) == 0 && x0 + 4 <= dw;
    int ring[5][4];
#pragma unroll
    for (int t = 0; t < 3; t++) pyrdown_hrow(s + (size_t)reflect_once(2 * y0 - 2 + t, h) * spitch, sx0, w, fastx, ring[t]);
    const int yend = min(y0 + PD_ROWS, dh);
    // ring slot of input row r (relative to 2*y0 - 2) is r % 5; output row y0 + i uses relative rows 2i .. 2i + 4
    for (int ib = 0; ib < PD_ROWS; ib += 5) {
#pragma unroll
        for (int u = 0; u < 5; u++) {
            const int i = ib + u, y = y0 + i;
            const int ya = reflect_once(min(2 * y + 1, h + 1), h), yb = reflect_once(min(2 * y + 2, h + 1), h);
            pyrdown_hrow(s + (size_t)ya * spitch, sx0, w, fastx, ring[(2 * u + 3) % 5]);
            pyrdown_hrow(s + (size_t)yb * spitch, sx0, w, fastx, ring[(2 * u + 4) % 5]);
            unsigned o = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int v = ring[(2 * u) % 5][j] + ring[(2 * u + 4) % 5][j] + 4 * (ring[(2 * u + 1) % 5][j] + ring[(2 * u + 3) % 5][j]) +
                              6 * ring[(2 * u + 2) % 5][j] + 128;
                o |= (unsigned)(v >> 8) << (8 * j);
            }
            if (y < yend) {
                uint8_t* dp = d + (size_t)y * dpitch + x0;
                if (fullw) *(unsigned*)dp = o;
                else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (x0 + j < dw) dp[j] = (uint8_t)(o >> (8 * j));
                }
            }
        }
    }
}

// ---------------- cv::Sobel 3x3 -> CV_16S dx, dy (binary_descriptor_custom.cpp:395-396), REFLECT_101 ----------------
// Same strip scheme: per input row the horizontal difference d[x] = r[x+1] - r[x-1] and smoothing
// s[x] = r[x-1] + 2 r[x] + r[x+1] of four pixels; dx = d0 + 2 d1 + d2, dy = s2 - s0 over a three-row register window.
#define SB_ROWS 33
__device__ __forceinline__ void sobel_hrow(const uint8_t* __restrict__ rp, int x0, int w, bool fastx, int (&dd)[4], int (&ss)[4])
{
    unsigned w0, w1, w2;
    row_words3(rp, x0, w, fastx, w0, w1, w2);
    int b[6];
    b[0] = (int)(w0 >> 24);
    b[1] = (int)(w1 & 0xffu); b[2] = (int)((w1 >> 8) & 0xffu); b[3] = (int)((w1 >> 16) & 0xffu); b[4] = (int)(w1 >> 24);
    b[5] = (int)(w2 & 0xffu);
#pragma unroll
    for (int j = 0; j < 4; j++) { dd[j] = b[j + 2] - b[j]; ss[j] = b[j] + 2 * b[j + 1] + b[j + 2]; }
}
__global__ void __launch_bounds__(128)
k_sobel3(const uint8_t* __restrict__ src, size_t sframe, int spitch, int w, int h,
         short2* __restrict__ dxy, size_t dframe, int F, int ncx_int)
{
    const int strip = plf_strip_of(blockIdx.x, threadIdx.x, w, 1, F, ncx_int);
    const int x0 = 4 * strip, y0 = (blockIdx.y * 4 + threadIdx.y) * SB_ROWS;
    if (strip < 0 || x0 >= w || y0 >= h) return;
    const uint8_t* s = src + (size_t)blockIdx.z * sframe;
    short2* oxy = dxy + (size_t)blockIdx.z * dframe;      // (dx, dy) interleaved: k_lbd fetches both with one 4-byte gather
    const bool fastx = ((((size_t)s) | (size_t)spitch) & 3) == 0 && x0 >= 4 && x0 + 8 <= w;
    const bool fullw = (w & 3) == 0 && (((size_t)oxy) & 15) == 0 && x0 + 4 <= w;
    int D[3][4], S[3][4];
    sobel_hrow(s + (size_t)reflect_once(y0 - 1, h) * spitch, x0, w, fastx, D[0], S[0]);
    sobel_hrow(s + (size_t)y0 * spitch, x0, w, fastx, D[1], S[1]);
    const int yend = min(y0 + SB_ROWS, h);
    for (int yb = y0; yb < yend; yb += 3) {
#pragma unroll
        for (int u = 0; u < 3; u++) {       // rows y-1, y, y+1 live in slots u, u+1, u+2 (mod 3)
            const int y = yb + u;
            sobel_hrow(s + (size_t)reflect_once(min(y, h - 1) + 1, h) * spitch, x0, w, fastx, D[(u + 2) % 3], S[(u + 2) % 3]);
            int gx[4], gy[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                gx[j] = D[u % 3][j] + 2 * D[(u + 1) % 3][j] + D[(u + 2) % 3][j];
                gy[j] = S[(u + 2) % 3][j] - S[u % 3][j];
            }
            if (y < yend) {
                const size_t o = (size_t)y * w + x0;
                if (fullw) {
                    uint4 v;
                    v.x = (unsigned)(gx[0] & 0xffff) | ((unsigned)gy[0] << 16); v.y = (unsigned)(gx[1] & 0xffff) | ((unsigned)gy[1] << 16);
                    v.z = (unsigned)(gx[2] & 0xffff) | ((unsigned)gy[2] << 16); v.w = (unsigned)(gx[3] & 0xffff) | ((unsigned)gy[3] << 16);
                    *(uint4*)(oxy + o) = v;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (x0 + j < w) oxy[o + j] = make_short2((short)gx[j], (short)gy[j]);
                }
            }
        }
    }
}

// ---------------- LSD ll_angle: gradient, level-line angle, component label init ----------------
// A pixel is "defined" iff norm = sqrt(q / 4.0) > rho, q = gx^2 + gy^2: the host turns that into the integer
// test q > qthr (qthr = largest q whose correctly rounded double norm is <= rho), so only the defined pixels
// (a few percent) do any floating point.  Per pixel: fa = fastAtan2(gx, -gy) in degrees or NOTDEF (all pixels);
// for defined pixels also q, cs = float cos / sin of the float-cast radian angle (what region_grow
// accumulates) and label = first pixel of the pixel's horizontal run inside its 32-px row segment.
// mask holds one bit per pixel (one word per warp: 32 consecutive pixels of a row); it is what the CCL and key
// kernels test, so they skip empty segments without touching the per-pixel arrays.
// Layout: every per-pixel array has P = w rounded up to 4 columns per row (the pad columns are NOTDEF), so a thread
// handles four adjacent pixels with one 32-bit load per image row and one 16-byte store of the angles.
// block = (32, 4): a warp covers 128 pixels (four mask words) of a row.
// A warp walks GRAD_ROWS consecutive rows: the GRAD_ROWS + 1 image words it needs are requested up front (one memory
// round trip per thread instead of one per row, and every image row is fetched 1.25 instead of 2 times).
#define GRAD_ROWS 4
__global__ void __launch_bounds__(128)
k_lsd_grad(const uint8_t* __restrict__ img, size_t iframe, int ipitch, int w, int P, int h, int qthr,
           int* __restrict__ q, float* __restrict__ fa, int* __restrict__ label,
           unsigned* __restrict__ mask, int mw, int* __restrict__ maxq)
{
    __shared__ unsigned s_ang[4][256];                    // per warp: 128 packed (gx, gy) in, 128 angles out
    const int lane = threadIdx.x;
    const int x0 = (blockIdx.x * 32 + lane) * 4, y0 = (blockIdx.y * 4 + threadIdx.y) * GRAD_ROWS;
    const int f = blockIdx.z;
    const unsigned FULL = 0xffffffffu;
    if (y0 >= h) return;                                  // warp-uniform
    const uint8_t* fr = img + (size_t)f * iframe;
    const bool aligned = ((((size_t)img) | (size_t)ipitch | iframe) & 3) == 0 && x0 + 4 <= ipitch;
    // five pixels of every row: x0 .. x0 + 4 (the fifth comes from the next lane's word)
    unsigned rw[GRAD_ROWS + 1], r4[GRAD_ROWS + 1];
#pragma unroll
    for (int k = 0; k <= GRAD_ROWS; k++) {
        rw[k] = 0; r4[k] = 0;
        const uint8_t* r = fr + (size_t)(y0 + k) * ipitch;
        if (y0 + k < h && x0 < w) {
            if (aligned) rw[k] = *(const unsigned*)(r + x0);
            else {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x0 + j < w) rw[k] |= (unsigned)r[x0 + j] << (8 * j);
            }
            if (lane == 31 && x0 + 4 < w) r4[k] = r[x0 + 4];
        }
    }
#pragma unroll
    for (int k = 0; k <= GRAD_ROWS; k++) {
        const unsigned nx = __shfl_down_sync(FULL, rw[k], 1) & 0xffu;
        if (lane != 31) r4[k] = nx;
    }
    int myq = -1;
    bool anydef = false;
#pragma unroll
    for (int k = 0; k < GRAD_ROWS; k++) {
        const int y = y0 + k;
        if (y >= h) break;                                // warp-uniform
        const bool rows_ok = y < h - 1;
        const unsigned a = rw[k], b = rw[k + 1];
        int pa[5], pb[5];
#pragma unroll
        for (int j = 0; j < 4; j++) { pa[j] = (int)((a >> (8 * j)) & 0xffu); pb[j] = (int)((b >> (8 * j)) & 0xffu); }
        pa[4] = (int)r4[k]; pb[4] = (int)r4[k + 1];
        int gx[4], gy[4], qq[4];
        unsigned nib = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int DA = pb[j + 1] - pa[j], BC = pa[j + 1] - pb[j];
            gx[j] = DA + BC; gy[j] = DA - BC;
            qq[j] = gx[j] * gx[j] + gy[j] * gy[j];
            const bool def = rows_ok && x0 + j < w - 1 && qq[j] > qthr;
            if (def) { nib |= 1u << j; myq = max(myq, qq[j]); }
        }
        anydef |= nib != 0;
        // the mask word of this lane's 32-px segment: nibbles of its eight lanes
        unsigned m = nib << (4 * (lane & 7));
        m |= __shfl_xor_sync(FULL, m, 1);
        m |= __shfl_xor_sync(FULL, m, 2);
        m |= __shfl_xor_sync(FULL, m, 4);
        const int seg = blockIdx.x * 4 + (lane >> 3);
        if ((lane & 7) == 0 && seg < mw) mask[((size_t)f * h + y) * mw + seg] = m;
        // Level-line angles.  Defined pixels are ~7 % of an image, but almost every warp row has a few, so evaluating fastAtan2
        // per pixel slot (j = 0 .. 3) ran the arctangent four times per row with one or two active lanes each.  Instead the
        // defined pixels of the warp's row are compacted through shared memory ((j, lane) order by four ballots), the arctangent
        // runs ONCE over the compacted list (<= 128 entries, typically ~9), and every lane picks its own results up again.
        float4 ang = make_float4(LSD_NOTDEF, LSD_NOTDEF, LSD_NOTDEF, LSD_NOTDEF);
        {
            const unsigned LT = (1u << lane) - 1u;
            unsigned bal[4];
            int idx[4], total = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                bal[j] = __ballot_sync(FULL, (nib >> j) & 1u);
                idx[j] = total + __popc(bal[j] & LT);
                total += __popc(bal[j]);
            }
            if (total) {                                   // warp-uniform
                unsigned* sin_ = s_ang[threadIdx.y];
                float* sout = (float*)(s_ang[threadIdx.y] + 128);
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if ((nib >> j) & 1u) sin_[idx[j]] = ((unsigned)gx[j] & 0xffffu) | ((unsigned)gy[j] << 16);
                __syncwarp();
                for (int i = lane; i < total; i += 32) {
                    const unsigned pk = sin_[i];
                    const int ggx = (int)(short)(pk & 0xffffu), ggy = (int)(short)(pk >> 16);
                    sout[i] = plf_fast_atan2((float)ggx, (float)(-ggy));
                }
                __syncwarp();
                float* av = &ang.x;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if ((nib >> j) & 1u) av[j] = sout[idx[j]];
                __syncwarp();                              // the buffers are reused by the next row
            }
        }
        if (x0 < P) {
            const size_t o = (size_t)f * P * h + (size_t)y * P + x0;
            if (nib) {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if ((nib >> j) & 1u) {
                        q[o + j] = qq[j];     // (cos / sin of the angle are filled in densely by k_lsd_cid after the sort)
                        // head of the run of defined pixels this pixel belongs to (inside its 32-px segment)
                        const int bit = 4 * (lane & 7) + j;
                        const unsigned below = ~m & ((1u << bit) - 1u);
                        const int head = below ? 32 - __clz((int)below) : 0;
                        label[o + j] = y * P + seg * 32 + head;
                    }
            }
            *(float4*)(fa + o) = ang;
        }
    }
    // frame maximum of q over defined pixels -> one atomic per warp that has any
    if (__any_sync(FULL, anydef)) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) myq = max(myq, __shfl_xor_sync(FULL, myq, s));
        if (lane == 0) atomicMax(&maxq[f], myq);
    }
}

// per frame: bin_coef = (n_bins - 1) / max_grad (OpenCV lsd.cpp ll_angle), once instead of per pixel
__global__ void k_lsd_bincoef(const int* __restrict__ maxq, int nframes, int n_bins, double* __restrict__ coef)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const int mq = maxq[f];
    const double max_grad = mq >= 0 ? sqrt((double)mq / 4.0) : -1.0;
    coef[f] = max_grad > 0 ? (double)(n_bins - 1) / max_grad : 0.0;
}

__device__ __forceinline__ int ccl_find(const int* L, int a)
{
    int p = ((volatile const int*)L)[a];
    while (p != a) { a = p; p = ((volatile const int*)L)[a]; }
    return a;
}
__device__ __forceinline__ void ccl_union(int* L, int a, int b)
{
    bool done;
    do {
        a = ccl_find(L, a);
        b = ccl_find(L, b);
        if (a < b) { int old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { int old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}
// 8-connected components of the defined mask: union-find in global memory (atomicMin links the larger root to
// the smaller).  A warp owns one 32-px row segment and reads the definedness of its neighbours from the bit
// mask (five warp-uniform words), so empty segments leave at once.  Pixels of one run inside a segment already
// share a label (k_lsd_grad), so only these links remain, each made once per pair of touching runs:
//   W  : the run continues from the previous segment (lane 0 only),
//   N  : first pixel of the overlap with a run above (skipped when the W neighbour made the same link via its N),
//   NE : N undefined, NE defined (a run above starting one to the right),
//   NW : N and W undefined, NW defined.
// One WARP per image row (block = (32, 8) = 8 rows): the lanes fetch the row's mask words and those of the row above
// with two coalesced loads, then the warp walks only the non-empty 32-px segments (the others cost nothing).
__global__ void __launch_bounds__(256)
k_ccl_merge(int* __restrict__ label, const unsigned* __restrict__ mask, int mw, int w, int h)
{
    const int lane = threadIdx.x, y = blockIdx.x * 8 + threadIdx.y;
    if (y >= h) return;
    const unsigned FULL = 0xffffffffu;
    const unsigned* M = mask + ((size_t)blockIdx.z * h + y) * mw;
    int* L = label + (size_t)blockIdx.z * w * h;
    for (int s0 = 0; s0 < mw; s0 += 32) {
        // words s0-1 .. s0+32 of this row and the row above, held one per lane (+ the two neighbours of the chunk)
        const int sg = s0 + lane;
        const unsigned cur = sg < mw ? M[sg] : 0u;
        const unsigned up = (y > 0 && sg < mw) ? M[sg - mw] : 0u;
        const unsigned curPrev = s0 > 0 ? M[s0 - 1] : 0u;                       // word left of the chunk (warp-uniform)
        const unsigned upPrev = (y > 0 && s0 > 0) ? M[s0 - 1 - mw] : 0u;
        const unsigned upNext = (y > 0 && s0 + 32 < mw) ? M[s0 + 32 - mw] : 0u;  // word right of the chunk in the row above
        unsigned todo = __ballot_sync(FULL, cur != 0u);
        while (todo) {
            const int s = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const unsigned m1 = __shfl_sync(FULL, cur, s);
            const unsigned m0 = __shfl_sync(FULL, up, s);
            const unsigned m1l = s > 0 ? __shfl_sync(FULL, cur, s - 1) : curPrev;
            const unsigned m0l = s > 0 ? __shfl_sync(FULL, up, s - 1) : upPrev;
            const unsigned m0r = s < 31 ? __shfl_sync(FULL, up, s + 1) : upNext;
            if (!((m1 >> lane) & 1u)) continue;
            const bool dW = lane > 0 ? (m1 >> (lane - 1)) & 1u : (m1l >> 31) & 1u;
            const bool dN = (m0 >> lane) & 1u;
            const bool dNW = lane > 0 ? (m0 >> (lane - 1)) & 1u : (m0l >> 31) & 1u;
            const bool dNE = lane < 31 ? (m0 >> (lane + 1)) & 1u : m0r & 1u;
            const int p = y * w + (s0 + s) * 32 + lane;
            if (lane == 0 && dW) ccl_union(L, p, p - 1);
            if (dN) { if (!(dW && dNW)) ccl_union(L, p, p - w); }
            else {
                if (dNE) ccl_union(L, p, p - w + 1);
                if (!dW && dNW) ccl_union(L, p, p - w - 1);
            }
        }
    }
}

// emit one sort key per defined pixel (root label, bin, raster index).  Positions come from an inclusive prefix sum of
// the mask popcounts (offs[i + 1] = number of defined pixels in mask words 0 .. i): no atomics, keys in raster order.
__global__ void __launch_bounds__(256)
k_lsd_keys(int* __restrict__ label, const int* __restrict__ q, const unsigned* __restrict__ mask, int mw,
           const int* __restrict__ offs, const double* __restrict__ coef, int w, int h, int n_bins,
           unsigned long long* __restrict__ keys, int keycap, int kb)
{
    // one WARP per image row (block = (32, 8)): the mask words of the row are fetched with one coalesced load and only
    // the non-empty segments are walked
    const int lane = threadIdx.x, y = blockIdx.x * 8 + threadIdx.y;
    const int f = blockIdx.z;
    if (y >= h) return;
    const unsigned FULL = 0xffffffffu;
    const size_t rowbase = ((size_t)f * h + y) * mw;
    int* Lf = label + (size_t)f * w * h;
    const double bin_coef = coef[f];
    for (int s0 = 0; s0 < mw; s0 += 32) {
        const int sg = s0 + lane;
        const unsigned cur = sg < mw ? mask[rowbase + sg] : 0u;
        const int myoff = sg < mw ? offs[rowbase + sg] : 0;
        unsigned todo = __ballot_sync(FULL, cur != 0u);
        while (todo) {
            const int s = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const unsigned m = __shfl_sync(FULL, cur, s);
            const int base = __shfl_sync(FULL, myoff, s);
            const bool have = (m >> lane) & 1u;
            // pixels of one run share a label (k_lsd_grad): only the head of each run chases its root, the others take it by shuffle
            const unsigned below = ~m & ((1u << lane) - 1u);
            const int head = below ? 32 - __clz((int)below) : 0;
            const int p = y * w + (s0 + s) * 32 + lane;
            int root = -1;
            if (have && head == lane) {
                root = ccl_find(Lf, p);
                if (Lf[p] != root) Lf[p] = root;     // shorten the chain for the run heads that hang below this one
            }
            root = __shfl_sync(FULL, root, head);
            if (have) {
                int bin = (int)(sqrt((double)q[(size_t)f * w * h + p] / 4.0) * bin_coef);
                if (bin < 0) bin = 0;
                if (bin > n_bins - 1) bin = n_bins - 1;
                const unsigned long long key = ((unsigned long long)f << (2 * LSD_KB + LSD_BB)) | ((unsigned long long)root << (LSD_KB + LSD_BB)) |
                                               ((unsigned long long)(n_bins - 1 - bin) << LSD_KB) | (unsigned long long)p;
                const int oo = base + __popc(m & ((1u << lane) - 1));
                if (oo < keycap) keys[oo] = key;
            }
        }
    }
}

// component heads in the sorted key array.  Two passes (count, fill) bucket the components by
// floor(log2(size)) so that the grow kernel starts the largest components first (longest-chain-first).
#define LSD_NBUCKET 24
#ifndef LSD_BIG_BUCKET
#define LSD_BIG_BUCKET 8    // components with >= 256 seeds get a warp of their own (k_lsd_grow_warp)
#endif
__global__ void __launch_bounds__(256)
k_lsd_heads(const unsigned long long* __restrict__ keys, int n, int2* __restrict__ comp, int* __restrict__ bcount,
            int* __restrict__ bfill, int* __restrict__ maxsize, int pass, int kb)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const unsigned long long tag = LSD_KEY_TAG(keys[i]);
    if (i != 0 && LSD_KEY_TAG(keys[i - 1]) == tag) return;
    int lo = i + 1, hi = n;   // first index whose tag differs
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (LSD_KEY_TAG(keys[mid]) == tag) lo = mid + 1; else hi = mid;
    }
    const int size = lo - i;
    int b = 31 - __clz(size);
    if (b >= LSD_NBUCKET) b = LSD_NBUCKET - 1;
    if (pass == 0) { atomicAdd(&bcount[b], 1); atomicMax(maxsize, size); return; }
    int base = 0;
    for (int k = LSD_NBUCKET - 1; k > b; k--) base += bcount[k];
    comp[base + atomicAdd(&bfill[b], 1)] = make_int2(i, size);
}

struct LsdRegion {
    int start, n;                 // slice of the region point arena
    double reg_angle;
    unsigned long long seedkey;   // key of the seed pixel (frame, bin, raster index)
};

// region_grow for every seed of one component, one thread per component, dynamic work distribution.
// Region points are stored packed (x | y << 16).  The 3x3 neighbourhood of a popped point is loaded in one
// batch (the only writes in between are this thread's own USED marks on distinct pixels), so a pop costs
// one memory round trip instead of nine.
__global__ void __launch_bounds__(128)
k_lsd_grow(const unsigned long long* __restrict__ keys, int n, const int2* __restrict__ comp, const int* __restrict__ bcount,
           int* __restrict__ next, float* __restrict__ fa, const float2* __restrict__ cs, int w, int h, double prec,
           int min_reg_size, int* __restrict__ regpts, LsdRegion* __restrict__ regions, int* __restrict__ nregions, int regcap, int skip_big, int spec_maxc, int kb)
{
    int ncomp = 0;
    for (int k = 0; k < LSD_NBUCKET; k++) ncomp += bcount[k];
    const size_t px = (size_t)w * h;
    // components are listed largest first.  Lanes of one warp execute divergent chains one after another, so
    // the first round gives every WARP one of the largest components (lane 0), then the next largest to lane 1,
    // ...; later rounds draw from the shared counter.
    const int lane = threadIdx.x & 31;
    const int totalWarps = (gridDim.x * blockDim.x) >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    bool first = true;
    for (;;) {
        int c;
        if (first) { c = lane * totalWarps + gwarp; first = false; }
        else c = 32 * totalWarps + atomicAdd(next, 1);
        if (c >= ncomp) { if (c >= 32 * totalWarps) break; else continue; }
        const int start = comp[c].x, end = start + comp[c].y;
        if (skip_big && comp[c].y >= (1 << LSD_BIG_BUCKET) && comp[c].y <= spec_maxc) continue;   // k_lsd_grow_warp does these
        float* F = fa + (size_t)LSD_KEY_FRAME(keys[start]) * px;
        const float2* CS = cs + (size_t)LSD_KEY_FRAME(keys[start]) * px;
        int arena = start;
        for (int i = start; i < end; i++) {
            const unsigned long long key = keys[i];
            const int p = LSD_KEY_IDX(key);
            const float v0 = F[p];
            if (v0 < -500.f) continue;   // already used by an earlier region
            const int r0 = arena;
            {
                const int sy = p / w;
                regpts[arena++] = (p - sy * w) | (sy << 16);
            }
            double reg_angle = (double)v0 * LSD_D2R;
            float sumdx = (float)cos(reg_angle), sumdy = (float)sin(reg_angle);
            F[p] = LSD_USED;
            int pp = regpts[r0];
            for (int r = r0; r < arena; r++) {
                const int x = pp & 0xffff, y = pp >> 16;
                float v[9];
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    const int xx = x + (k % 3) - 1, yy = y + (k / 3) - 1;
                    const bool inb = xx >= 0 && yy >= 0 && xx < w && yy < h;
                    v[k] = inb ? F[yy * w + xx] : LSD_NOTDEF;
                }
                if (r + 1 < arena) pp = regpts[r + 1];   // overlap the next pop with this point's tests
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    if (v[k] < -500.f) continue;   // NOTDEF, USED or outside
                    double n_theta = reg_angle - (double)v[k] * LSD_D2R;
                    if (n_theta < 0) n_theta = -n_theta;
                    if (n_theta > LSD_3_2_PI) {
                        n_theta -= LSD_2PI;
                        if (n_theta < 0) n_theta = -n_theta;
                    }
                    if (n_theta <= prec) {
                        const int xx = x + (k % 3) - 1, yy = y + (k / 3) - 1;
                        const int qi = yy * w + xx;
                        F[qi] = LSD_USED;
                        if (arena == r + 1) pp = xx | (yy << 16);   // it is the next point to pop
                        regpts[arena++] = xx | (yy << 16);
                        const float2 c2 = CS[qi];
                        sumdx += c2.x;
                        sumdy += c2.y;
                        reg_angle = (double)plf_fast_atan2(sumdy, sumdx) * LSD_D2R;
                    }
                }
            }
            const int nreg = arena - r0;
            if (nreg >= min_reg_size) {
                const int rf = LSD_KEY_FRAME(key);
                const int rr = atomicAdd(nregions + rf, 1);      // slot inside the frame's block of `regcap` regions
                if (rr < regcap) {
                    LsdRegion R;
                    R.start = r0; R.n = nreg; R.reg_angle = reg_angle; R.seedkey = key;
                    regions[(size_t)rf * regcap + rr] = R;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Large components: one WARP per component, still strictly in the reference's order.
// A single GPU thread needs ~250 dependent instructions per region pixel; here the nine neighbours of the
// popped point are handled by nine lanes at once (address, component index, used bit, angle, cos/sin and the
// alignment test), and only the part that is sequential by definition stays sequential: neighbours are
// accepted in k order, and after each acceptance the region angle is updated and the neighbours AFTER it are
// re-tested (the reference's loop never goes back to an earlier neighbour of the same pop).
// `used` is a shared-memory bitmap indexed by the pixel's position in the sorted seed list (`cid`, written
// into the label array after the sort), so the angle / cos-sin arrays stay read-only and L1-resident.
// ------------------------------------------------------------------------------------------------
#define WARPGROW_MAXC (256 * 1024)          // component pixels one warp can track (32 KB of used bits; four warps per CTA)
#define WG_RING 128                         // queue entries kept in shared memory (the frontier is a handful of entries; older ones come from the global arena)

// after the sort: the sorted position of every defined pixel (its compact index inside the component), and -- one
// thread per defined pixel, no divergence -- cs = float cos / sin of the float-cast radian angle, which is what
// region_grow accumulates (OpenCV lsd.cpp: sumdx += cos(float(angle)))
__global__ void __launch_bounds__(256)
k_lsd_cid(const unsigned long long* __restrict__ keys, int n, int* __restrict__ label, const float* __restrict__ fa,
          float2* __restrict__ cs, size_t px, int kb)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    const size_t o = (size_t)LSD_KEY_FRAME(k) * px + LSD_KEY_IDX(k);
    label[o] = i;
    const float af = (float)((double)fa[o] * LSD_D2R);
    cs[o] = make_float2((float)cos((double)af), (float)sin((double)af));
}

// One candidate: neighbour `nb` of queue entry q.  The queue tail lives in a shared-memory ring; older entries
// come from the global arena.  The three loads are independent of each other (one memory round trip).
struct WgCand { int xx, yy, ci, pp; float fv; float2 cv; bool inb; };
__device__ __forceinline__ WgCand wg_load(int q, int arena, const int* ring, const int* __restrict__ regpts, int gdx, int gdy, int w, int h,
                                          const int* __restrict__ CID, const float* __restrict__ F, const float2* __restrict__ CS)
{
    WgCand c;
    const int pp = (arena - q <= WG_RING) ? ring[q & (WG_RING - 1)] : regpts[q];
    c.pp = pp;
    c.xx = (pp & 0xffff) + gdx; c.yy = (pp >> 16) + gdy;
    c.inb = c.xx >= 0 && c.yy >= 0 && c.xx < w && c.yy < h;
    c.ci = -1; c.fv = 0.f; c.cv = make_float2(0.f, 0.f);
    if (c.inb) {
        const int qi = c.yy * w + c.xx;
        c.ci = __ldg(&CID[qi]);
        c.fv = __ldg(&F[qi]);
        c.cv = __ldg(&CS[qi]);
        // the entries accepted from here will look one pixel further: pull the 5x5 ring towards L1 now
        const int x2 = c.xx + gdx, y2 = c.yy + gdy;
        if (x2 >= 0 && y2 >= 0 && x2 < w && y2 < h) {
            const int q2 = qi + gdy * w + gdx;
            PLF_PREFETCH_L1(&CID[q2]);
            PLF_PREFETCH_L1(&F[q2]);
            PLF_PREFETCH_L1(&CS[q2]);
        }
    }
    return c;
}

__device__ __forceinline__ double wg_ntheta(double reg_angle, double a)
{
    double n_theta = reg_angle - a;
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > LSD_3_2_PI) {
        n_theta -= LSD_2PI;
        if (n_theta < 0) n_theta = -n_theta;
    }
    return n_theta;
}

// Four queue entries are expanded per iteration: lane 8g + k handles the k-th neighbour (centre skipped) of
// entry r + g, so lane order equals the reference's test order (entry, then neighbour in raster order).
//
// Sequential rule: neighbours are accepted in that order, and after every acceptance the region angle
// reg_angle = fastAtan2(sumdy, sumdx) changes, so every LATER test sees the new angle.
// Batched rule (the common case): with m candidates passing the test at the iteration's starting angle, the
// angle can move by at most D = 1.02 m sin(prec + 0.1) / |sum| + 1e-3 rad while they are accepted (each
// accepted unit vector lies within prec + D + E of the current direction and |sum| only grows; E = 1.7e-4 rad
// is the fastAtan2 model error, 2E < 1e-3).  If every candidate's |n_theta - prec| exceeds D, no decision can
// flip, so all passing candidates are accepted at once (first lane per pixel), the float sums are added in
// lane order, and one fastAtan2 closes the iteration -- bit-identical to the sequential rule.  Otherwise the
// iteration falls back to one acceptance per round.
// The loads of the NEXT iteration's candidates are issued before this iteration's tests (their used bits are
// checked when they are consumed), which hides the L2 round trip behind the arithmetic.
// CTAs hold WG_WARPS independent warps (one component each): an SM has only 32 CTA slots, and one-warp CTAs of a few
// concurrent launches (one per line context) would take them all and starve every other kernel on the device.
#define WG_E 4
#define WG_WARPS 4
__global__ void __launch_bounds__(32 * WG_WARPS)
k_lsd_grow_warp(const unsigned long long* __restrict__ keys, const int2* __restrict__ comp, const int* __restrict__ bcount,
                const float* __restrict__ fa, const float2* __restrict__ cs, const int* __restrict__ cid, int w, int h,
                double prec, int min_reg_size, int* __restrict__ regpts, LsdRegion* __restrict__ regions,
                int* __restrict__ nregions, int regcap, int kb, int maxc, int giant_bucket)
{
    PLF_DYN_SMEM(smem);
    __shared__ int s_ring[WG_WARPS][WG_RING];
    __shared__ float2 s_acc[WG_WARPS][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned* used = (unsigned*)smem + (size_t)wid * (maxc >> 5);
    int* ring = s_ring[wid];
    float2* acc = s_acc[wid];
    const unsigned FULL = 0xffffffffu;
    int nbig = 0;
    for (int k = LSD_BIG_BUCKET; k < LSD_NBUCKET; k++) nbig += bcount[k];
    const size_t px = (size_t)w * h;
    const int grp = lane >> 3, kk = lane & 7;
    const int nbr = kk < 4 ? kk : kk + 1;
    const int gdx = (nbr % 3) - 1, gdy = (nbr / 3) - 1;
    const bool fast_ok = prec < 1.4;
    const float sphi = (float)sin(prec + 0.1) * 1.05f;   // 1.05 covers the float rounding of this product and of rsqrtf below
    int ngiant = 0;      // the first components of the list (largest first) belong to k_lsd_grow_cta
    for (int k = giant_bucket; k < LSD_NBUCKET; k++) ngiant += bcount[k];
    for (int c = ngiant + blockIdx.x * WG_WARPS + wid; c < nbig; c += gridDim.x * WG_WARPS) {
        const int start = comp[c].x, C = comp[c].y, end = start + C;
        if (C > maxc) continue;   // handled by k_lsd_grow
        const size_t foff = (size_t)LSD_KEY_FRAME(keys[start]) * px;
        const float* F = fa + foff;
        const float2* CS = cs + foff;
        const int* CID = cid + foff;
        __syncwarp();
        for (int i = lane; i < (C + 31) / 32; i += 32) used[i] = 0u;
        __syncwarp();
        int arena = start;
        for (int i0 = start; i0 < end; i0 += 32) {
            // 32 seeds at a time; only the ones still unused at this point can start a region
            const int ii = i0 + lane;
            unsigned long long mykey = 0;
            bool mine = false;
            if (ii < end) {
                mykey = keys[ii];
                const int sc = ii - start;
                mine = !((used[sc >> 5] >> (sc & 31)) & 1u);
            }
            unsigned todo = __ballot_sync(FULL, mine);
            while (todo) {
                const int s = __ffs((int)todo) - 1;
                todo &= todo - 1;
                const int sc = i0 + s - start;
                if ((used[sc >> 5] >> (sc & 31)) & 1u) continue;            // taken by a region grown meanwhile (warp-uniform)
                const unsigned long long key = __shfl_sync(FULL, mykey, s);
                const int p = LSD_KEY_IDX(key);
                const int r0 = arena;
                const int sy = p / w, sx = p - sy * w;
                if (lane == 0) {
                    regpts[arena] = sx | (sy << 16);
                    ring[arena & (WG_RING - 1)] = sx | (sy << 16);
                    used[sc >> 5] |= 1u << (sc & 31);
                }
                arena++;
                double reg_angle = (double)__ldg(&F[p]) * LSD_D2R;
                float sumdx = (float)cos(reg_angle), sumdy = (float)sin(reg_angle);
                __syncwarp();
                int npf = 0;              // entries whose candidates were prefetched for the coming iteration
                WgCand pf;
                pf.xx = pf.yy = pf.pp = 0; pf.ci = -1; pf.fv = 0.f; pf.cv = make_float2(0.f, 0.f); pf.inb = false;
                for (int r = r0; r < arena;) {
                    const int ng = min(WG_E, arena - r);
                    WgCand cd = pf;
                    if (grp >= npf) {
                        cd.inb = false; cd.ci = -1; cd.pp = -(4 << 16);   // far outside: never adjacent to a pixel
                        if (grp < ng) cd = wg_load(r + grp, arena, ring, regpts, gdx, gdy, w, h, CID, F, CS);
                    }
                    const int rn = r + ng;
                    npf = min(WG_E, arena - rn);
                    if (grp < npf) pf = wg_load(rn + grp, arena, ring, regpts, gdx, gdy, w, h, CID, F, CS);
                    int cc = -1 - lane;       // position of the pixel in the component's seed list (negative: none)
                    bool cand = cd.inb && cd.fv > -500.f;     // defined (labels of undefined pixels are stale)
                    if (cand) {
                        cc = cd.ci - start;
                        cand = !((used[cc >> 5] >> (cc & 31)) & 1u);
                        if (!cand) cc = -1 - lane;
                    }
                    const double a = (double)cd.fv * LSD_D2R;
                    double n_theta = wg_ntheta(reg_angle, a);
                    bool pass = cand && n_theta <= prec;
                    unsigned m = __ballot_sync(FULL, pass);
                    if (m) {
                        bool fast = false;
                        if (fast_ok) {
                            const float L2 = sumdx * sumdx + sumdy * sumdy;
                            if (L2 >= 16.f) {
                                const float D = (float)__popc(m) * sphi * rsqrtf(L2) + 1e-3f;
                                const double Dd = (double)D;
                                const bool risky = cand && fabs(n_theta - prec) <= Dd;
                                fast = D <= 0.09f && !__any_sync(FULL, risky);
                            }
                        }
                        if (fast) {
                            // first lane per pixel: two queue entries can have the same neighbour (needed only when
                            // more than one lane passes)
                            bool win = pass;
                            if (__popc(m) > 1) {
                                const unsigned same = __match_any_sync(FULL, cc);     // non-candidates hold unique negative values
                                win = pass && (__ffs((int)same) - 1 == lane);
                            }
                            const unsigned mw = __ballot_sync(FULL, win);
                            const int nw = __popc(mw);
                            if (win) {
                                const int rank = __popc(mw & ((1u << lane) - 1u));
                                atomicOr(&used[cc >> 5], 1u << (cc & 31));
                                regpts[arena + rank] = cd.xx | (cd.yy << 16);
                                ring[(arena + rank) & (WG_RING - 1)] = cd.xx | (cd.yy << 16);
                                acc[rank] = cd.cv;
                            }
                            __syncwarp();
                            // float sums in acceptance (lane) order, read back as broadcasts
                            for (int j = 0; j < nw; j += 4) {
                                const float2 v0 = acc[j], v1 = acc[(j + 1) & 31], v2 = acc[(j + 2) & 31], v3 = acc[(j + 3) & 31];
                                sumdx += v0.x; sumdy += v0.y;
                                if (j + 1 < nw) { sumdx += v1.x; sumdy += v1.y; }
                                if (j + 2 < nw) { sumdx += v2.x; sumdy += v2.y; }
                                if (j + 3 < nw) { sumdx += v3.x; sumdy += v3.y; }
                            }
                            arena += nw;
                            reg_angle = (double)plf_fast_atan2(sumdy, sumdx) * LSD_D2R;
                        } else {
                            for (;;) {
                                const int k0 = __ffs((int)m) - 1;                  // next acceptance in the reference's order
                                if (lane == k0) {
                                    used[cc >> 5] |= 1u << (cc & 31);              // only this lane writes in this round
                                    regpts[arena] = cd.xx | (cd.yy << 16);
                                    ring[arena & (WG_RING - 1)] = cd.xx | (cd.yy << 16);
                                }
                                arena++;
                                const int cc0 = __shfl_sync(FULL, cc, k0);
                                sumdx += __shfl_sync(FULL, cd.cv.x, k0);
                                sumdy += __shfl_sync(FULL, cd.cv.y, k0);
                                reg_angle = (double)plf_fast_atan2(sumdy, sumdx) * LSD_D2R;
                                cand = cand && lane > k0 && cc != cc0;             // earlier tests stand; the same pixel seen from another entry is now used
                                pass = cand && wg_ntheta(reg_angle, a) <= prec;
                                m = __ballot_sync(FULL, pass);
                                if (!m) break;
                            }
                        }
                    }
                    __syncwarp();
                    r = rn;
                }
                const int nreg = arena - r0;
                if (lane == 0 && nreg >= min_reg_size) {
                    const int rf = LSD_KEY_FRAME(key);
                    const int rr = atomicAdd(nregions + rf, 1);
                    if (rr < regcap) {
                        LsdRegion R;
                        R.start = r0; R.n = nreg; R.reg_angle = reg_angle; R.seedkey = key;
                        regions[(size_t)rf * regcap + rr] = R;
                    }
                }
                __syncwarp();
            }
        }
    }
}

#include "plf_lsd_grow_cta.cuh"

// region2rect + get_theta (OpenCV lsd.cpp, refine = 0): one WARP per region.  The double sums are order dependent,
// so they are accumulated strictly in region order -- but the per-point work (gather, double sqrt, products) is done
// by 32 lanes at once and only the additions are replayed in order (every lane replays them from shuffles, which
// keeps the warp converged): ~10 cycles per point instead of a dependent gather + sqrt chain per point.
// Output: line end points (float, +0.5, / SCALE) and an order key (frame, bin descending, raster).
#define RECT_WARPS 8
__global__ void __launch_bounds__(32 * RECT_WARPS)
k_lsd_rect(const LsdRegion* __restrict__ regions, const int* __restrict__ nregions, int regcap, const int* __restrict__ regpts,
           const int* __restrict__ q, int w, int h, double prec, double scale, float4* __restrict__ lines,
           unsigned long long* __restrict__ linekey, int* __restrict__ errflag, int kb)
{
    // grid (split, frames): the warps of a frame's CTAs stride over the frame's region slots
    __shared__ double s_buf[RECT_WARPS][32][3];
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    double (*buf)[3] = s_buf[threadIdx.x >> 5];
    const unsigned FULL = 0xffffffffu;
    int nr = nregions[f];
    if (nr > regcap) { nr = regcap; if (blockIdx.x == 0 && threadIdx.x == 0) *errflag = 1; }
    const int* Q = q + (size_t)f * w * h;
    for (int slot = blockIdx.x * RECT_WARPS + (threadIdx.x >> 5); slot < nr; slot += gridDim.x * RECT_WARPS) {
    const size_t rr = (size_t)f * regcap + slot;
    const LsdRegion R = regions[rr];
    const int* pts = regpts + R.start;
    double x = 0, y = 0, sum = 0;
    for (int i0 = 0; i0 < R.n; i0 += 32) {
        const int cnt = min(32, R.n - i0);
        double wgt = 0, xw = 0, yw = 0;
        if (lane < cnt) {
            const int pk = pts[i0 + lane];
            const int py = pk >> 16, pxx = pk & 0xffff;
            wgt = sqrt((double)Q[py * w + pxx] / 4.0);
            xw = (double)pxx * wgt;
            yw = (double)py * wgt;
        }
        __syncwarp();
        buf[lane][0] = xw; buf[lane][1] = yw; buf[lane][2] = wgt;
        __syncwarp();
        for (int j = 0; j < cnt; j++) {      // ordered replay from shared-memory broadcasts
            x += buf[j][0];
            y += buf[j][1];
            sum += buf[j][2];
        }
    }
    x /= sum; y /= sum;
    double Ixx = 0, Iyy = 0, Ixy = 0;
    for (int i0 = 0; i0 < R.n; i0 += 32) {
        const int cnt = min(32, R.n - i0);
        double txx = 0, tyy = 0, txy = 0;
        if (lane < cnt) {
            const int pk = pts[i0 + lane];
            const int py = pk >> 16, pxx = pk & 0xffff;
            const double wgt = sqrt((double)Q[py * w + pxx] / 4.0);
            const double dx = (double)pxx - x, dy = (double)py - y;
            txx = dy * dy * wgt;
            tyy = dx * dx * wgt;
            txy = dx * dy * wgt;
        }
        __syncwarp();
        buf[lane][0] = txx; buf[lane][1] = tyy; buf[lane][2] = txy;
        __syncwarp();
        for (int j = 0; j < cnt; j++) {
            Ixx += buf[j][0];
            Iyy += buf[j][1];
            Ixy -= buf[j][2];
        }
    }
    const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)plf_fast_atan2((float)(lambda - Ixx), (float)Ixy)
                                           : (double)plf_fast_atan2((float)Ixy, (float)(lambda - Iyy));
    theta *= LSD_D2R;
    double diff = theta - R.reg_angle;
    while (diff <= -LSD_PI) diff += LSD_2PI;
    while (diff > LSD_PI) diff -= LSD_2PI;
    if (fabs(diff) > prec) theta += LSD_PI;
    const double dx = cos(theta), dy = sin(theta);
    // l_min = min(0, min l), l_max = max(0, max l): the reference's `if (l > l_max) .. else if (l < l_min) ..` with both
    // starting at 0 never lets one value update both, so the order of the points does not matter here
    double l_min = 0, l_max = 0;
    for (int i = lane; i < R.n; i += 32) {
        const int pk = pts[i];
        const int py = pk >> 16, pxx = pk & 0xffff;
        const double regdx = (double)pxx - x, regdy = (double)py - y;
        const double l = regdx * dx + regdy * dy;
        if (l > l_max) l_max = l;
        if (l < l_min) l_min = l;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const double a = __shfl_xor_sync(FULL, l_max, s), b = __shfl_xor_sync(FULL, l_min, s);
        if (a > l_max) l_max = a;
        if (b < l_min) l_min = b;
    }
    __syncwarp();
    if (lane != 0) continue;
    double x1 = x + l_min * dx, y1 = y + l_min * dy, x2 = x + l_max * dx, y2 = y + l_max * dy;
    x1 += 0.5; y1 += 0.5; x2 += 0.5; y2 += 0.5;
    if (scale != 1) { x1 /= scale; y1 /= scale; x2 /= scale; y2 /= scale; }
    float4 L;
    L.x = (float)x1; L.y = (float)y1; L.z = (float)x2; L.w = (float)y2;
    lines[rr] = L;
    linekey[rr] = R.seedkey & ((1ull << (LSD_KB + LSD_BB)) - 1);     // (bin descending, raster) order inside the frame
    }
}

// KeyLine assembly for one octave (LSDDetector_custom.cpp:266-308): one CTA per frame.  The frame's regions sit in its own
// block of slots in arrival order; their order keys are unique, so the position of a line in the reference's seed order is
// the number of keys of the frame below its own (counted against a shared-memory copy).
// det layout: [frame][octave][detcap] keylines in seed order, with a per (frame, octave) count.
#define KL_T 256
#define KL_MAXREG 4096
__global__ void __launch_bounds__(KL_T)
k_lsd_keylines(const unsigned long long* __restrict__ linekey, const int* __restrict__ nregions, int regcap,
               const float4* __restrict__ lines, int octave, int noct, int ow, int oh, double min_length,
               plf_keyline* __restrict__ det, int* __restrict__ detcount, int detcap)
{
    __shared__ unsigned long long s_key[KL_MAXREG];
    const int f = blockIdx.x;
    int n = nregions[f];
    if (n > regcap) n = regcap;
    const size_t base = (size_t)f * regcap;
    for (int i = threadIdx.x; i < n; i += KL_T) s_key[i] = linekey[base + i];
    if (threadIdx.x == 0) detcount[f * noct + octave] = n;
    __syncthreads();
    plf_keyline* out = det + ((size_t)f * noct + octave) * detcap;
    for (int i = threadIdx.x; i < n; i += KL_T) {
    const unsigned long long key = s_key[i];
    int pos = 0;
    for (int j = 0; j < n; j++) pos += s_key[j] < key;
    if (pos >= detcap) continue;
    const float4 L = lines[base + i];
    float e0 = L.x, e1 = L.y, e2 = L.z, e3 = L.w;
    // checkLineExtremes (:76-102)
    if (e0 < 0) e0 = 0;
    if (e0 >= ow) e0 = (float)ow - 1.0f;
    if (e2 < 0) e2 = 0;
    if (e2 >= ow) e2 = (float)ow - 1.0f;
    if (e1 < 0) e1 = 0;
    if (e1 >= oh) e1 = (float)oh - 1.0f;
    if (e3 < 0) e3 = 0;
    if (e3 >= oh) e3 = (float)oh - 1.0f;
    const float d02 = e0 - e2, d13 = e1 - e3;
    const double length = (double)(float)sqrt((double)d02 * (double)d02 + (double)d13 * (double)d13);
    plf_keyline K;
    const float os = (float)(1 << octave);
    K.startPointX = e0 * os; K.startPointY = e1 * os; K.endPointX = e2 * os; K.endPointY = e3 * os;
    K.sPointInOctaveX = e0; K.sPointInOctaveY = e1; K.ePointInOctaveX = e2; K.ePointInOctaveY = e3;
    K.lineLength = (float)length;
    const int ax = __float2int_rn(e0), ay = __float2int_rn(e1), bx = __float2int_rn(e2), by = __float2int_rn(e3);
    const int adx = ax > bx ? ax - bx : bx - ax, ady = ay > by ? ay - by : by - ay;
    K.numOfPixels = (adx > ady ? adx : ady) + 1;
    K.angle = plf_libm::atan2f_glibc(K.endPointY - K.startPointY, K.endPointX - K.startPointX);   // atan2(float, float) = atan2f (:298)
    K.class_id = (length > min_length) ? 0 : -1;   // -1 marks "dropped by the min_length filter" for the next stage
    K.octave = octave;
    K.size = (K.endPointX - K.startPointX) * (K.endPointY - K.startPointY);
    K.response = K.lineLength / (float)(ow > oh ? ow : oh);
    K.pt_x = (K.endPointX + K.startPointX) / 2;
    K.pt_y = (K.endPointY + K.startPointY) / 2;
    out[pos] = K;
    }
}

// min_length filter + class ids (detectImpl) and, when `select`, the per-octave response quota with
// mid-point keypoints (Lineextractor::ComputeLsdWithLbd, src/Lineextractor.cc:138-207).  One CTA per frame.
#define SEL_T 256
#define SEL_SMEM_BYTES(detcap) ((size_t)(detcap) * (sizeof(int) + sizeof(float) + sizeof(unsigned short)))
__global__ void __launch_bounds__(SEL_T)
k_line_select(const plf_keyline* __restrict__ det, const int* __restrict__ detcount, int detcap, int noct,
              int select, int quota0, int quota1, plf_keyline* __restrict__ out_kl, plf_keypoint* __restrict__ out_mid,
              int cap, int* __restrict__ n_out, const int* __restrict__ regerr0, const int* __restrict__ regerr1)
{
    __shared__ int s_pos[2][2];   // [octave]: {valid count, kept count}
    __shared__ int s_err, s_tie;
    PLF_DYN_SMEM(smem);
    int* vidx = (int*)smem;                                   // valid line indices of the current octave (order preserved)
    float* skey = (float*)(vidx + detcap);                    // their responses
    unsigned short* sperm = (unsigned short*)(skey + detcap); // sorted position -> index into vidx
    const int f = blockIdx.x, tid = threadIdx.x;
    int outbase = 0;
    if (tid == 0) s_err = 0;
    __syncthreads();
    for (int o = 0; o < noct; o++) {
        const plf_keyline* D = det + ((size_t)f * noct + o) * detcap;
        int cnt = detcount[f * noct + o];
        if (cnt > detcap) { if (tid == 0) s_err = 1; cnt = detcap; }
        // order-preserving compaction of lines that passed min_length (serial scan by one thread: cnt is small)
        if (tid == 0) {
            int m = 0;
            for (int i = 0; i < cnt; i++) if (D[i].class_id == 0) vidx[m++] = i;
            s_pos[o][0] = m;
            s_tie = 0;
        }
        __syncthreads();
        const int m = s_pos[o][0];
        const int quota = o == 0 ? quota0 : quota1;
        const bool sorted = select && m > quota;
        const int keep = sorted ? quota : m;
        if (sorted) {
            for (int i = tid; i < m; i += SEL_T) skey[i] = D[vidx[i]].response;
            __syncthreads();
            // Lineextractor.cc:175: std::sort by descending response.  With all responses distinct the order is unique:
            // a line's place is the number of larger responses.  Equal responses make the result depend on libstdc++'s
            // introsort itself (std::sort is unstable), which one thread then replays exactly (plf_stdsort.cuh).
            for (int i = tid; i < m; i += SEL_T) {
                const float r = skey[i];
                int rank = 0, tie = 0;
                for (int j = 0; j < m; j++) {
                    const float rj = skey[j];
                    rank += (rj > r || (rj == r && j < i)) ? 1 : 0;
                    tie |= (rj == r && j != i) ? 1 : 0;
                }
                sperm[rank] = (unsigned short)i;
                if (tie) s_tie = 1;
            }
            __syncthreads();
            if (s_tie && tid == 0) {
                for (int i = 0; i < m; i++) sperm[i] = (unsigned short)i;
                plf_stdsort::sort_desc(skey, m, sperm);
            }
            __syncthreads();
        }
        for (int pos = tid; pos < keep; pos += SEL_T) {
            const int i = sorted ? (int)sperm[pos] : pos;
            const int op = outbase + pos;
            if (op < cap) {
                plf_keyline K = D[vidx[i]];
                K.class_id = op;
                out_kl[(size_t)f * cap + op] = K;
                if (out_mid) {
                    plf_keypoint P;
                    P.x = (K.startPointX + K.endPointX) / 2;
                    P.y = (K.startPointY + K.endPointY) / 2;
                    P.size = 0; P.angle = -1; P.response = 0; P.octave = K.octave; P.class_id = -1;
                    out_mid[(size_t)f * cap + op] = P;
                }
            } else s_err = 2;
        }
        outbase += keep;
        __syncthreads();
    }
    // -1: more than detcap lines in one octave, -2: output capacity, -3: the LSD region buffer of an octave overflowed somewhere in
    // this batch (sticky per call: regions were dropped, so no frame of the batch is trustworthy)
    if (tid == 0) {
        const bool regerr = (regerr0 && *regerr0) || (regerr1 && *regerr1);
        n_out[f] = regerr ? -3 : (s_err ? -s_err : outbase);
    }
}

// ---------------- LBD (binary_descriptor_custom.cpp:1026-1372 + :645-667) ----------------
// One CTA of 64 threads per line: thread hID < 63 walks one row of the line support region sequentially
// (float sums in the reference's order), threads < 9 accumulate their band in row order, thread 0 does the
// mean/std, normalisations and clamp; then 32 threads binarise.
struct LbdImages {
    const short2* dxy[2];   // (dx, dy) per pixel
    size_t frame[2];   // pixels per frame
    int w[2], h[2];
};
struct LbdCoefs { float g[63]; float l[21]; };

__constant__ int c_lbd_comb[32][2] = {
    {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
    {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
    {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

// bands, mean / std, normalisations, clamp and the 32 comparison bytes from the 63 row sums (all 64 threads of the CTA call it)
__device__ __forceinline__ void lbd_finish(const int tid, float (*rows)[4], float (*band)[8], float* dv, const LbdCoefs& cf,
                                           uint8_t* __restrict__ desc, float* __restrict__ fdesc, const int f, const int cap, const int li)
{
    __syncthreads();
    if (tid < 9) {
        float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // pgdL, ngdL, pgdL2, ngdL2, pgdO, ngdO, pgdO2, ngdO2
        const int h0 = tid * 7 - 7 < 0 ? 0 : tid * 7 - 7, h1 = tid * 7 + 14 > 63 ? 63 : tid * 7 + 14;
        for (int hID = h0; hID < h1; hID++) {
            const int b = hID / 7;
            float c;
            if (b == tid) c = cf.l[hID % 7 + 7];
            else if (b == tid + 1) c = cf.l[hID % 7 + 14];   // row's band-1 == this band
            else c = cf.l[hID % 7];                            // row's band+1 == this band
            const float cc = c * c;
            const float r0 = rows[hID][0], r1 = rows[hID][1], r2 = rows[hID][2], r3 = rows[hID][3];
            float m;
            m = c * r0; s[0] += m;
            m = c * r1; s[1] += m;
            m = cc * (r0 * r0); s[2] += m;
            m = cc * (r1 * r1); s[3] += m;
            m = c * r2; s[4] += m;
            m = c * r3; s[5] += m;
            m = cc * (r2 * r2); s[6] += m;
            m = cc * (r3 * r3); s[7] += m;
        }
        // mean / standard deviation of this thread's band (:1262-1290): the same operations thread by thread instead of one
        // thread walking the nine bands
        const float invN = (tid == 0 || tid == 8) ? (float)(1.0 / 14.0) : (float)(1.0 / 21.0);
        float temp, u, v;
        temp = s[0] * invN; dv[tid * 8] = temp;
        u = s[2] * invN; v = temp * temp; dv[tid * 8 + 4] = sqrtf(u - v);
        temp = s[1] * invN; dv[tid * 8 + 1] = temp;
        u = s[3] * invN; v = temp * temp; dv[tid * 8 + 5] = sqrtf(u - v);
        temp = s[4] * invN; dv[tid * 8 + 2] = temp;
        u = s[6] * invN; v = temp * temp; dv[tid * 8 + 6] = sqrtf(u - v);
        temp = s[5] * invN; dv[tid * 8 + 3] = temp;
        u = s[7] * invN; v = temp * temp; dv[tid * 8 + 7] = sqrtf(u - v);
    }
    __syncthreads();
    // the three order-dependent float sums stay with one thread (the reference's order); the element-wise scalings and the
    // clamp are spread over the CTA (same operation per element, so the same bits)
    float* scale = &band[0][0];       // band[][] is dead from here: two floats of it carry the factors
    if (tid == 0) {
        float tempM = 0, tempS = 0, m;
        for (int i = 0; i < 72; i += 8) {
            for (int k = 0; k < 4; k++) { m = dv[i + k] * dv[i + k]; tempM += m; }
            for (int k = 4; k < 8; k++) { m = dv[i + k] * dv[i + k]; tempS += m; }
        }
        scale[0] = 1.0f / sqrtf(tempM);   // namespace cv: using std::sqrt -> sqrt(float); int / float (:1302-1303)
        scale[1] = 1.0f / sqrtf(tempS);
    }
    __syncthreads();
    for (int i = tid; i < 72; i += 64) {
        float x = dv[i] * ((i & 4) ? scale[1] : scale[0]);
        if ((double)x > 0.4) x = (float)0.4;
        dv[i] = x;
    }
    __syncthreads();
    if (tid == 0) {
        float temp = 0, m;
        for (int i = 0; i < 72; i++) { m = dv[i] * dv[i]; temp += m; }
        scale[2] = 1.0f / sqrtf(temp);
    }
    __syncthreads();
    for (int i = tid; i < 72; i += 64) dv[i] = dv[i] * scale[2];
    __syncthreads();
    if (tid < 32 && desc) {
        const float* f1 = &dv[8 * c_lbd_comb[tid][0]];
        const float* f2 = &dv[8 * c_lbd_comb[tid][1]];
        unsigned r = 0;
        for (int b = 0; b < 8; b++) if (f1[b] > f2[b]) r += 1u << b;
        desc[((size_t)f * cap + li) * 32 + tid] = (uint8_t)r;
    }
    if (fdesc) for (int i = tid; i < 72; i += 64) fdesc[((size_t)f * cap + li) * 72 + i] = dv[i];
}

__global__ void __launch_bounds__(64, 16)
k_lbd(const plf_keyline* __restrict__ kl, const int* __restrict__ nlines, int cap, LbdImages im, LbdCoefs cf,
      uint8_t* __restrict__ desc, float* __restrict__ fdesc)
{
    __shared__ float rows[63][4];
    __shared__ float band[9][8];
    __shared__ float dv[72];
    const int f = blockIdx.y, li = blockIdx.x, tid = threadIdx.x;
    int nl = nlines[f];
    if (li >= nl) return;
    const plf_keyline K = kl[(size_t)f * cap + li];
    const int o = K.octave;
    const int realWidth = im.w[o], imageWidth = realWidth - 1, imageHeight = im.h[o] - 1;
    const short2* pdxy = im.dxy[o] + (size_t)f * im.frame[o];
    const short lengthOfLSP = (short)K.numOfPixels;
    const short halfWidth = (short)((lengthOfLSP - 1) / 2);
    const short halfHeight = 31;
    const float midX = (float)(0.5 * (double)(K.sPointInOctaveX + K.ePointInOctaveX));
    const float midY = (float)(0.5 * (double)(K.sPointInOctaveY + K.ePointInOctaveY));
    const float dL0 = plf_libm::cosf_glibc(K.angle), dL1 = plf_libm::sinf_glibc(K.angle);   // cos(float) = cosf (:1130-1131)
    const float dO0 = -dL1, dO1 = dL0;
    if (tid < 63) {
        float t0 = -dL0 * (float)halfWidth, t1 = dL1 * (float)halfHeight;
        float sCorX0 = t0 + t1 + midX;
        t0 = -dL1 * (float)halfWidth; t1 = dL0 * (float)halfHeight;
        float sCorY0 = t0 - t1 + midY;
        for (int hh = 0; hh < tid; hh++) { sCorX0 -= dL1; sCorY0 += dL0; }   // same repeated float updates as the row loop
        float sCorX = sCorX0, sCorY = sCorY0;
        float pgdL = 0, ngdL = 0, pgdO = 0, ngdO = 0;
        for (int wID = 0; wID < lengthOfLSP; wID++) {
            // round() of the reference (double, half away from zero) == roundf(): float -> double is exact and so is the rounding
            short tc = (short)(int)roundf(sCorX);
            const int xCor = tc < 0 ? 0 : (tc > imageWidth ? imageWidth : tc);
            tc = (short)(int)roundf(sCorY);
            const int yCor = tc < 0 ? 0 : (tc > imageHeight ? imageHeight : tc);
            const short2 g2 = __ldg(&pdxy[yCor * realWidth + xCor]);
            const float dxv = (float)g2.x, dyv = (float)g2.y;
            float a0 = dxv * dL0, a1 = dyv * dL1;
            const float gDL = a0 + a1;
            a0 = dxv * dO0; a1 = dyv * dO1;
            const float gDO = a0 + a1;
            if (gDL > 0) pgdL += gDL; else ngdL -= gDL;
            if (gDO > 0) pgdO += gDO; else ngdO -= gDO;
            sCorX += dL0;
            sCorY += dL1;
        }
        const float coef = cf.g[tid];
        rows[tid][0] = coef * pgdL; rows[tid][1] = coef * ngdL; rows[tid][2] = coef * pgdO; rows[tid][3] = coef * ngdO;
    }
    lbd_finish(tid, rows, band, dv, cf, desc, fdesc, f, cap, li);
}

#include "plf_fld_kernels.cuh"
