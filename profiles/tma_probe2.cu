// tma_probe2.cu -- second round: the CUDA programming guide's own TMA example (libcu++ wrappers), a plain bulk copy, descriptor dump.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <dlfcn.h>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

__global__ void k_guide(const __grid_constant__ CUtensorMap tm, int x, int y, int* out)
{
    __shared__ alignas(128) int tile[64][64];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token tok;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&tile, &tm, x, y, bar);
        tok = cuda::device::barrier_arrive_tx(bar, 1, sizeof(tile));
    } else tok = bar.arrive();
    bar.wait(std::move(tok));
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) out[i] = tile[i / 64][i % 64];
}
__global__ void k_bulk(const int* src, int* out)
{
    __shared__ alignas(128) int tile[1024];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token tok;
    if (threadIdx.x == 0) {
        cuda::memcpy_async(tile, src, cuda::aligned_size_t<16>(sizeof(tile)), bar);
        tok = bar.arrive();
    } else tok = bar.arrive();
    bar.wait(std::move(tok));
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = tile[i];
}
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 256, H = 256;
    int* src; cudaMalloc(&src, W * H * 4);
    int* h = (int*)malloc(W * H * 4); for (int i = 0; i < W * H; i++) h[i] = i * 7 + 1;
    cudaMemcpy(src, h, W * H * 4, cudaMemcpyHostToDevice);
    int* out; cudaMalloc(&out, 64 * 64 * 4); cudaMemset(out, 0, 64 * 64 * 4);
    int ho[64 * 64];
    if (mode == 0) {
        k_bulk<<<1, 128>>>(src, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("bulk copy: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(ho, out, 4096, cudaMemcpyDeviceToHost); int bad = 0; for (int i = 0; i < 1024; i++) bad += ho[i] != h[i];
        printf("bulk copy ok, %d bad\n", bad); return 0;
    }
    enc_fn enc = nullptr;
    if (mode == 1) { void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); enc = (enc_fn)p; }
    else { void* lib = dlopen("libcuda.so.1", RTLD_NOW); enc = lib ? (enc_fn)dlsym(lib, "cuTensorMapEncodeTiled") : nullptr; }
    if (!enc) { printf("no encode\n"); return 2; }
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[2] = {W, H}, strides[1] = {W * 4}; cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc %d; descriptor:", (int)r);
    for (int i = 0; i < 16; i++) printf(" %016llx", (unsigned long long)((uint64_t*)&tm)[i]);
    printf("\n");
    k_guide<<<1, 128>>>(tm, 64, 32, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d guide example: %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(ho, out, 64 * 64 * 4, cudaMemcpyDeviceToHost); int bad = 0;
    for (int i = 0; i < 64 * 64; i++) bad += ho[i] != h[(32 + i / 64) * W + 64 + i % 64];
    printf("mode %d guide example ok, %d bad\n", mode, bad);
    return 0;
}
