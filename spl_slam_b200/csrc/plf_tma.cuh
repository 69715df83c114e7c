// plf_tma.cuh -- the few TMA / mbarrier pieces the kernels use (sm_100a PTX, no library): tensor-map encoding through the
// driver entry point (no link-time dependency on libcuda), mbarrier init / expect-tx / wait, cp.async.bulk.tensor.3d loads.
// Hardware rule found the hard way (profiles/tma_probe.cu): the first byte of a box must be 16-byte aligned in global memory,
// i.e. the innermost coordinate times the element size must be a multiple of 16, or the load faults ("illegal instruction").
#pragma once
#ifndef PLF_EMU
#include <cuda.h>

typedef CUresult (*plf_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static plf_encode_tiled_fn plf_get_encode_tiled()
{
    static plf_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (plf_encode_tiled_fn)p;
    }
    return fn;
}

// (x, y, frame) map over a batch of images of `esize`-byte elements (1 or 4); false when the driver call is missing or the
// base / pitch / frame stride are not 16-byte aligned.  Out-of-range box parts are zero-filled.
static bool plf_tma_map_images(CUtensorMap* tm, const void* base, int esize, int w, int h, int nframes, size_t pitch_bytes, size_t frame_bytes,
                               int box_w, int box_h)
{
    plf_encode_tiled_fn enc = plf_get_encode_tiled();
    if (!enc) return false;
    if (((uintptr_t)base & 15) || (pitch_bytes & 15) || (frame_bytes & 15) || ((size_t)box_w * esize & 15) || box_w > 256 || box_h > 256) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nframes};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch_bytes, (cuuint64_t)frame_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ unsigned plf_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void plf_mbar_init(unsigned bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void plf_mbar_expect_tx(unsigned bar, int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void plf_mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    // a wait that cannot complete (wrong parity, wrong byte count, faulted copy) must become an error, not a hung GPU
    for (unsigned spins = 0; !done; spins++) {
        asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 22)) asm volatile("trap;");
    }
}
// box at (x, y, z) of the map -> shared memory (128-byte aligned), completion counted in bytes on the barrier
__device__ __forceinline__ void plf_tma_load_3d(unsigned dst, const CUtensorMap* tm, unsigned bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}
#endif
