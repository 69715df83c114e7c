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

struct plf_ctx {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    uint64_t launches;
    char err[512];
    // small reusable device scratch (grown on demand)
    void* scratch;
    size_t scratch_bytes;
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

#define PLF_CHECK_LAUNCH(ctx)                                                                     \
    do {                                                                                          \
        (ctx)->launches++;                                                                        \
        cudaError_t e_ = cudaGetLastError();                                                      \
        if (e_ != cudaSuccess)                                                                    \
            return plf_fail((ctx), PLF_ERR_CUDA, "kernel launch failed: %s (%s:%d)",              \
                            cudaGetErrorString(e_), __FILE__, __LINE__);                          \
    } while (0)

plf_status plf_ctx_scratch(plf_ctx* ctx, size_t bytes, void** out);
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

__host__ __device__ __forceinline__ int plf_reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}
