// cuda_emu.h -- TEST-ONLY functional emulation of the small CUDA subset the kernels in
// spl_slam_b200/csrc use, so their logic can be exercised by `pytest -m "not gpu"` on a
// box without a GPU.  The product library (libplf.so) is built by nvcc and never sees this
// file; libplf_emu.so is built only by tests/emu/build_emu.py and loaded only by tests.
//
// Model: blocks run one after another; the threads of a block are real pthreads from a
// pool; __syncthreads / warp collectives are implemented with barriers.  Shared memory is
// a function-local static (blocks are sequential, so one copy is enough).
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <functional>
#include <vector>
#include <algorithm>

#define PLF_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3_ { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uchar4 { unsigned char x, y, z, w; };
struct int2 { int x, y; };
struct short2 { short x, y; };
static inline short2 make_short2(short a, short b) { short2 r = {a, b}; return r; }
struct uint2 { unsigned x, y; };
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline uint2 make_uint2(unsigned a, unsigned b) { uint2 r = {a, b}; return r; }
static inline int2 make_int2(int a, int b) { int2 r = {a, b}; return r; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r = {a, b, c, d}; return r; }
static inline float2 make_float2(float a, float b) { float2 r = {a, b}; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r = {a, b, c, d}; return r; }

extern thread_local uint3_ threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
extern thread_local int emu_lane_linear;  // linear thread id inside the block
extern thread_local unsigned char* emu_dyn_smem_ptr;

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaHostAllocDefault = 0 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emu error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = 0)
{
    for (size_t y = 0; y < h; y++) memcpy((char*)d + y * dp, (const char*)s + y * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaFuncSetAttribute(T, int, int) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
#define cudaMemcpyToSymbol(sym, src, n) (memcpy(&(sym), (src), (n)), cudaSuccess)
#define cudaMemcpyToSymbolAsync(sym, src, n, off, kind, st) (memcpy((char*)&(sym) + (off), (src), (n)), cudaSuccess)

// ---- launch machinery ----
void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void emu_syncthreads();
unsigned emu_ballot(unsigned mask, int pred);
unsigned long long emu_shfl64(unsigned mask, unsigned long long v, int srcLane);
void emu_syncwarp(unsigned mask);

#define PLF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu_launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define PLF_DYN_SMEM(name) unsigned char* name = emu_dyn_smem_ptr

static inline void __syncthreads() { emu_syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu_syncwarp(mask); }
static inline unsigned __ballot_sync(unsigned mask, int pred) { return emu_ballot(mask, pred); }
static inline int __any_sync(unsigned mask, int pred) { return emu_ballot(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return emu_ballot(mask, pred) == emu_ballot(mask, 1); }
static inline unsigned __activemask() { return 0xffffffffu; }

template <class T> static inline T emu_shfl(unsigned mask, T v, int src)
{
    static_assert(sizeof(T) <= 8, "shfl size");
    unsigned long long u = 0;
    memcpy(&u, &v, sizeof(T));
    u = emu_shfl64(mask, u, src);
    T r;
    memcpy(&r, &u, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    int lane = emu_lane_linear & 31;
    int base = lane & ~(width - 1);
    return emu_shfl(mask, v, base + (src & (width - 1)));
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    int lane = emu_lane_linear & 31;
    int src = lane + (int)d;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
    return emu_shfl(mask, v, src);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    int lane = emu_lane_linear & 31;
    int src = lane - (int)d;
    if (src < (lane & ~(width - 1))) src = lane;
    return emu_shfl(mask, v, src);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32)
{
    int lane = emu_lane_linear & 31;
    return emu_shfl(mask, v, lane ^ x);
}
template <class T> static inline unsigned __match_any_sync(unsigned mask, T v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; i++) {
        T o = emu_shfl(mask, v, i);
        if (o == v) r |= 1u << i;
    }
    return r;
}

// ---- atomics ----
template <class T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float* p, float v)
{
    float old, nw;
    do { old = *p; nw = old + v; } while (!__atomic_compare_exchange(p, &old, &nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return old;
}
template <class T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicCAS(T* p, T cmp, T v)
{
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
template <class T> static inline T atomicMax(T* p, T v)
{
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
template <class T> static inline T atomicMin(T* p, T v)
{
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- math / bit intrinsics ----
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __brev(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline int __float2int_rn(float v) { return (int)lrintf(v); }
static inline int __float2int_rd(float v) { return (int)floorf(v); }
static inline int __double2int_rn(double v) { return (int)lrint(v); }
static inline int __double2int_rd(double v) { return (int)floor(v); }
static inline float __double2float_rn(double v) { return (float)v; }
static inline float __int2float_rn(int v) { return (float)v; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline unsigned __vabsdiffu4(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        int x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff;
        r |= (unsigned)(x > y ? x - y : y - x) << (8 * i);
    }
    return r;
}
static inline unsigned __vcmpgtu4(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 0xff) > ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i);
    return r;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned n = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)((v >> (8 * (n & 7))) & 0xff);
        if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __dp2a_lo(unsigned a, unsigned b, unsigned c) { return c + (a & 0xffffu) * (b & 0xffu) + (a >> 16) * ((b >> 8) & 0xffu); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh)
{
    unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
template <class T> static inline T __ldg(const T* p) { return *p; }
using std::max;
using std::min;
static inline int __vabsdiffu4_sum_placeholder() { return 0; }
