// plf_common.cuh -- shared declarations for the libplf.so translation units.
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (float/double results
// are specified as IEEE ops without FMA contraction, see DESIGN.md "float exactness").
// The PLF_EMU branch exists only for tests/emu (functional emulation on a GPU-less box).
#pragma once
#ifdef PLF_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
// every launch site has a `plf_ctx* ctx` in scope; with profiling on, each launch is bracketed by CUDA events
#define PLF_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
    do {                                                                   \
        plf_prof_begin(ctx, #kernel);                                      \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);        \
        plf_prof_end(ctx);                                                 \
    } while (0)
#define PLF_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define PLF_PREFETCH_L1(ptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr))
#endif
#ifdef PLF_EMU
#define PLF_PREFETCH_L1(ptr) ((void)(ptr))
#endif
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "plf.h"
#include "plf_libm.cuh"

struct plf_ctx {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    uint64_t launches;
    char err[512];
    // small reusable device scratch (grown on demand)
    void* scratch;
    size_t scratch_bytes;
    // host waits: spinning (lowest latency, the default) or sleep-and-poll.  Batch calls switch to sleeping:
    // several ranks x several extractor threads spinning in cudaStreamSynchronize starve each other on the host cores
    // (8 ranks x 5 threads on 32 cores: end-to-end throughput at 8 GPUs dropped to 0.80 of the device-resident figure)
    int blocking;
    // second region for the device copies of host-buffer arguments (the kernels' own temporaries live in `scratch`)
    void* ioscratch;
    size_t ioscratch_bytes;
    // pinned host staging (grown on demand)
    void* pinned;
    size_t pinned_bytes;
    // optional per-kernel profiling (plf_profile_enable): event pairs per launch, summed per kernel name
    int prof_on, prof_n;
    struct plf_prof_state* prof;
};

#ifdef PLF_EMU
static inline void plf_prof_begin(plf_ctx*, const char*) {}
static inline void plf_prof_end(plf_ctx*) {}
#else
void plf_prof_begin(plf_ctx* ctx, const char* name);
void plf_prof_end(plf_ctx* ctx);
#endif

static inline plf_status plf_fail(plf_ctx* ctx, plf_status st, const char* fmt, ...)
{
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return st;
}

#define PLF_CUDA(ctx, call)                                                                             \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return plf_fail((ctx), PLF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                            __FILE__, __LINE__);                                                        \
    } while (0)

#ifndef PLF_EMU
// Opt a kernel in to the device's MAXIMUM dynamic shared memory, once per (kernel, device).  The attribute is per-function
// state shared by every extractor and host thread of the process: setting it to one caller's requirement would LOWER it
// for another (two ORB extractors with different nfeatures), so it is only ever set to the maximum.
#include <mutex>
#include <map>
static inline cudaError_t plf_smem_optin(const void* kernel, int device)
{
    static std::mutex mu;
    static std::map<const void*, unsigned long long> done;   // kernel -> bit per device
    std::lock_guard<std::mutex> lk(mu);
    const unsigned long long bit = 1ull << (device & 63);
    if (done[kernel] & bit) return cudaSuccess;
    int mx = 0;
    cudaError_t e = cudaDeviceGetAttribute(&mx, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx - (int)fa.sharedSizeBytes);   // dynamic + static <= opt-in maximum
    if (e == cudaSuccess) done[kernel] |= bit;
    return e;
}
#define PLF_SMEM_OPTIN(ctx, kernel) PLF_CUDA(ctx, plf_smem_optin((const void*)(kernel), (ctx)->device))
#else
#define PLF_SMEM_OPTIN(ctx, kernel) do { } while (0)
#endif

#define PLF_CHECK_LAUNCH(ctx)                                                                     \
    do {                                                                                          \
        (ctx)->launches++;                                                                        \
        cudaError_t e_ = cudaGetLastError();                                                      \
        if (e_ != cudaSuccess)                                                                    \
            return plf_fail((ctx), PLF_ERR_CUDA, "kernel launch failed: %s (%s:%d)",              \
                            cudaGetErrorString(e_), __FILE__, __LINE__);                          \
    } while (0)

plf_status plf_ctx_scratch(plf_ctx* ctx, size_t bytes, void** out);
plf_status plf_sync(plf_ctx* ctx, cudaStream_t st);      // wait for a stream the way ctx->blocking says
plf_status plf_ctx_ioscratch(plf_ctx* ctx, size_t bytes, void** out);
plf_status plf_ctx_pinned(plf_ctx* ctx, size_t bytes, void** out);

static inline size_t plf_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int plf_div_up(int a, int b) { return (a + b - 1) / b; }

// cv::fastAtan2 scalar model (degrees); float32, no FMA (file is built with -fmad=false).
__host__ __device__ __forceinline__ float plf_fast_atan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = 2.220446049250313e-16f;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// REFLECT_101 for an index that is at most n - 1 beyond either end (every strip kernel: halo <= 8, rows >= 16 px): one
// reflection, no loop -- the edge strips execute this 12 times per row and were a tenth of k_blur7's instructions with the loop
__host__ __device__ __forceinline__ int plf_reflect101_near(int p, int n)
{
    p = p < 0 ? -p : p;
    return p >= n ? 2 * (n - 1) - p : p;
}

__host__ __device__ __forceinline__ int plf_reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

// ------------------------------------------------------------------------------------------------
// Separable Q8 Gaussian with 2R+1 taps (cv::GaussianBlur on 8U, BORDER_REFLECT_101):
//   H = sum_t k[t] * src[x + t - R]   (<= 65280),   dst = (sum_t k[t] * H[y + t - R] + 32768) >> 16.
// A thread produces a 4-px wide column strip of `rows` rows with a register sliding window: per input row it
// loads three aligned 32-bit words (bytes x0-4 .. x0+7; the overlap between neighbouring threads is served
// by L1), forms the horizontal sums for its four pixels two at a time (stride-2 byte pairs packed in 16-bit
// halves: b[i] | b[i+2] << 16, products <= 65280 never carry between the halves), keeps the last 2R+1 rows of
// sums in registers and emits one packed 32-bit store per output row.  No shared memory, no barriers.
// Strips that touch the left/right image edge (or unaligned buffers) gather their 12 bytes with reflection.
// ------------------------------------------------------------------------------------------------
struct BlurTaps { int k[8]; };   // k[0..2R]

// the 12 bytes x0-4 .. x0+7 of a row as three words (aligned fast path, or gathered with REFLECT_101)
template <int R>
__device__ __forceinline__ void plf_blur_load3(const uint8_t* __restrict__ rp, int x0, int w, bool fastx, unsigned& w0, unsigned& w1, unsigned& w2)
{
    if (fastx) {
        const unsigned* p = (const unsigned*)(rp + x0 - 4);
        w0 = p[0]; w1 = p[1]; w2 = p[2];
    } else {
        unsigned b[12];
        if (w >= 16) {
#pragma unroll
            for (int i = 0; i < 12; i++) b[i] = (i >= 4 - R && i < 8 + R) ? rp[plf_reflect101_near(x0 - 4 + i, w)] : 0u;
        } else {
#pragma unroll
            for (int i = 0; i < 12; i++) b[i] = (i >= 4 - R && i < 8 + R) ? rp[plf_reflect101(x0 - 4 + i, w)] : 0u;
        }
        w0 = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
        w1 = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
        w2 = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
    }
}

// horizontal sums of the four pixels of a strip from the three words of its row
template <int R>
__device__ __forceinline__ void plf_blur_hsum(unsigned w0, unsigned w1, unsigned w2, const int (&kc)[2 * R + 1], int (&out)[4])
{
    // q[i] = byte(i) | byte(i + 2) << 16, byte index relative to x0 - 4
    unsigned q[10];
    q[0] = __byte_perm(w0, 0u, 0x4240); q[1] = __byte_perm(w0, 0u, 0x4341);
    q[4] = __byte_perm(w1, 0u, 0x4240); q[5] = __byte_perm(w1, 0u, 0x4341);
    q[8] = __byte_perm(w2, 0u, 0x4240); q[9] = __byte_perm(w2, 0u, 0x4341);
    q[2] = __funnelshift_r(q[0], q[4], 16); q[3] = __funnelshift_r(q[1], q[5], 16);
    q[6] = __funnelshift_r(q[4], q[8], 16); q[7] = __funnelshift_r(q[5], q[9], 16);
    unsigned h02 = (unsigned)kc[R] * q[4], h13 = (unsigned)kc[R] * q[5];
#pragma unroll
    for (int t = 0; t < R; t++) {
        h02 += (unsigned)kc[t] * (q[4 - R + t] + q[4 + R - t]);
        h13 += (unsigned)kc[t] * (q[5 - R + t] + q[5 + R - t]);
    }
    out[0] = (int)(h02 & 0xffffu); out[2] = (int)(h02 >> 16);
    out[1] = (int)(h13 & 0xffffu); out[3] = (int)(h13 >> 16);
}

template <int R>
__device__ __forceinline__ void plf_blur_hrow(const uint8_t* __restrict__ rp, int x0, int w, bool fastx, const int (&kc)[2 * R + 1], int (&out)[4])
{
    unsigned w0, w1, w2;
    plf_blur_load3<R>(rp, x0, w, fastx, w0, w1, w2);
    plf_blur_hsum<R>(w0, w1, w2, kc, out);
}

template <int R>
__device__ __forceinline__ void plf_blur_strip(const uint8_t* __restrict__ src, int spitch, uint8_t* __restrict__ dst, int dpitch,
                                               int w, int h, int x0, int y0, int rows, const BlurTaps& taps)
{
    constexpr int K = 2 * R + 1;
    if (x0 >= w) return;
    int kc[K];
#pragma unroll
    for (int t = 0; t < K; t++) kc[t] = taps.k[t];
    const bool aligned = ((((size_t)src) | ((size_t)dst) | (size_t)spitch | (size_t)dpitch) & 3) == 0;
    const bool fastx = aligned && x0 >= 4 && x0 + 8 <= w;
    const bool fullw = aligned && x0 + 4 <= w;
    int ring[K][4];
    // prime the window with rows y0 - R .. y0 + R - 1 (slots 0 .. K-2); h > R, so one reflection is enough
#pragma unroll
    for (int t = 0; t < K - 1; t++) {
        int yy = y0 - R + t;
        yy = yy < 0 ? -yy : yy;
        yy = yy >= h ? 2 * (h - 1) - yy : yy;
        plf_blur_hrow<R>(src + (size_t)yy * spitch, x0, w, fastx, kc, ring[t]);
    }
    const int yend = min(y0 + rows, h);
    // the words of the NEXT input row are requested before the current row is worked on: the loads were what the kernel
    // waited for at the first byte permute of every row (14 % of k_blur7's stall samples)
    auto row_of = [&](int y) { int yy = min(y, h - 1) + R; return yy >= h ? 2 * (h - 1) - yy : yy; };
    unsigned n0, n1, n2;
    plf_blur_load3<R>(src + (size_t)row_of(y0) * spitch, x0, w, fastx, n0, n1, n2);
    // rows are processed K at a time so that the ring slots are compile-time; rows past the image bottom are
    // computed from clamped addresses and simply not stored (no branch around the loads)
    for (int yb = y0; yb < yend; yb += K) {
#pragma unroll
        for (int s = 0; s < K; s++) {        // slot (s + K - 1) % K receives row y + R
            const int y = yb + s;
            const unsigned c0 = n0, c1 = n1, c2 = n2;
            plf_blur_load3<R>(src + (size_t)row_of(y + 1) * spitch, x0, w, fastx, n0, n1, n2);
            plf_blur_hsum<R>(c0, c1, c2, kc, ring[(s + K - 1) % K]);
            unsigned o = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                unsigned v = 32768u + (unsigned)kc[R] * (unsigned)ring[(s + R) % K][j];
#pragma unroll
                for (int t = 0; t < R; t++)
                    v += (unsigned)kc[t] * (unsigned)(ring[(s + t) % K][j] + ring[(s + K - 1 - t) % K][j]);
                o |= (v >> 16) << (8 * j);
            }
            if (y < yend) {
                uint8_t* dp = dst + (size_t)y * dpitch + x0;
                if (fullw) *(unsigned*)dp = o;
                else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (x0 + j < w) dp[j] = (uint8_t)(o >> (8 * j));
                }
            }
        }
    }
}

// Strip scheduling for the 4-px column-strip kernels: strips whose 12-byte (or 16-byte) window stays inside the row take
// the aligned fast path, the strips at the left / right image edge gather with reflection and are ~3x more expensive.
// Mixing both in one warp makes the whole warp pay for both, so the interior strips 1 .. F are packed densely into
// the first `ncx - 1` CTA columns and ONE extra CTA column holds the edge strips (strip 0 and F+1 .. S-1).
// F = number of interior strips, S = total strips; returns the strip index of (CTA column, lane) or -1.
__host__ __device__ __forceinline__ int plf_strip_interior(int w, int halo_lo, int halo_hi)
{
    // strip s (pixels 4s .. 4s+3) is interior iff 4s - halo_lo >= 0 and 4s + halo_hi <= w, halo_lo a multiple of 4
    const int first = halo_lo / 4;                 // first interior strip
    const int last = (w - halo_hi) / 4;            // last interior strip (may be < first)
    return last >= first ? last - first + 1 : 0;
}
__device__ __forceinline__ int plf_strip_of(int bx, int lane, int w, int first, int F, int ncx_int)
{
    const int S = (w + 3) >> 2;
    if (bx < ncx_int) {
        const int i = bx * 32 + lane;
        return i < F ? first + i : -1;
    }
    // edge column: lanes 0 .. first-1 -> left strips, then the right strips
    if (lane < first) return lane < S ? lane : -1;
    const int s = first + F + (lane - first);
    return s < S ? s : -1;
}
