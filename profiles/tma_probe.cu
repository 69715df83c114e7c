// tma_probe.cu -- which way of handing a CUtensorMap to cp.async.bulk.tensor works on this driver (one variant per process:
// a fault poisons the context).   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcudart
//   variant 0: one map, __grid_constant__ parameter           1: array of 16 maps in a >4 KB parameter block, dynamic index
//   variant 2: array of maps in global memory, dynamic index  3: like 1 but a 2-D map            4: like 1, small parameter block
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

struct Maps { CUtensorMap m[16]; };
struct Pad { char c[2176]; };
struct One { CUtensorMap m; };

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__constant__ int g_fence;
template <int RANK>
__device__ void load_and_dump(const CUtensorMap* tm, int x, int y, int z, int bytes, uint8_t* out)
{
    __shared__ __align__(128) uint8_t tile[4096];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = s32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (g_fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (g_fence) __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(s32(tile)), "l"(tm), "r"(b), "r"(x), "r"(y), "r"(z) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(tile)), "l"(tm), "r"(b), "r"(x), "r"(y) : "memory");
    }
    __syncwarp();
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }"
                     : "=r"(done) : "r"(b), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += 32) out[i] = tile[i];
}

__global__ void k0(const __grid_constant__ One t, int x, int y, int z, int bytes, uint8_t* out) { load_and_dump<3>(&t.m, x, y, z, bytes, out); }
__global__ void k1(Pad pad, const __grid_constant__ Maps t, int l, int x, int y, int z, int bytes, uint8_t* out)
{
    if (pad.c[5] == 77) out[0] = 1;
    load_and_dump<3>(&t.m[l], x, y, z, bytes, out);
}
__global__ void k2(const CUtensorMap* t, int l, int x, int y, int z, int bytes, uint8_t* out) { load_and_dump<3>(t + l, x, y, z, bytes, out); }
__global__ void k3(Pad pad, const __grid_constant__ Maps t, int l, int x, int y, int bytes, uint8_t* out)
{
    if (pad.c[5] == 77) out[0] = 1;
    load_and_dump<2>(&t.m[l], x, y, 0, bytes, out);
}
__global__ void k4(const __grid_constant__ Maps t, int l, int x, int y, int z, int bytes, uint8_t* out) { load_and_dump<3>(&t.m[l], x, y, z, bytes, out); }

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int BW = argc > 2 ? atoi(argv[2]) : 32, BH = argc > 3 ? atoi(argv[3]) : 31;
    const int W = 752, H = 480, F = 4, pitch = 768;
    uint8_t* img; cudaMalloc(&img, (size_t)pitch * H * F);
    uint8_t* h = (uint8_t*)malloc((size_t)pitch * H * F);
    for (size_t i = 0; i < (size_t)pitch * H * F; i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    cudaMemcpy(img, h, (size_t)pitch * H * F, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encode fn\n"); return 2; }
    enc_fn enc = (enc_fn)p;
    const int rank = variant == 3 ? 2 : 3;
    Maps maps; memset(&maps, 0, sizeof(maps));
    for (int l = 0; l < 16; l++) {
        cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
        cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * H};
        cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 3; }
    }
    uint8_t* out; cudaMalloc(&out, 4096); cudaMemset(out, 0, 4096);
    const int x = argc > 4 ? atoi(argv[4]) : 100, y = 50; { int fz = argc > 5 ? atoi(argv[5]) : 0; cudaMemcpyToSymbol(g_fence, &fz, 4); } const int z = rank == 3 ? 2 : 0, bytes = BW * BH, l = 5;
    Pad pad; memset(&pad, 0, sizeof(pad));
    if (variant == 0) { One o; o.m = maps.m[0]; k0<<<1, 32>>>(o, x, y, z, bytes, out); }
    else if (variant == 1) k1<<<1, 32>>>(pad, maps, l, x, y, z, bytes, out);
    else if (variant == 2) { CUtensorMap* d; cudaMalloc(&d, sizeof(maps)); cudaMemcpy(d, &maps, sizeof(maps), cudaMemcpyHostToDevice); k2<<<1, 32>>>(d, l, x, y, z, bytes, out); }
    else if (variant == 3) k3<<<1, 32>>>(pad, maps, l, x, y, bytes, out);
    else k4<<<1, 32>>>(maps, l, x, y, z, bytes, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d box %dx%d x=%d: %s\n", variant, BW, BH, x, cudaGetErrorString(e)); return 1; }
    uint8_t o[4096]; cudaMemcpy(o, out, 4096, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < BH; r++) for (int c = 0; c < BW; c++) {
        uint8_t want = (x + c < W) ? h[((size_t)z * H + y + r) * pitch + x + c] : 0;
        bad += o[r * BW + c] != want;
    }
    printf("variant %d box %dx%d x=%d: ok, %d mismatching bytes\n", variant, BW, BH, x, bad);
    return 0;
}
