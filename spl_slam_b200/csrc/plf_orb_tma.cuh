// plf_orb_tma.cuh -- IC_Angle + steered BRIEF with the two patches of a keypoint staged in shared memory by TMA.
//
// k_describe gathers 749 + 512 single bytes per keypoint straight from global memory (60 % of its stall samples were long
// scoreboard waits).  Here one lane per warp arms an mbarrier and issues two cp.async.bulk.tensor.3d tile loads -- the 31 x 31
// intensity-centroid patch of the level image (box 48 x 31) and the 39 x 39 neighbourhood of the blurred level that the
// rotated pattern can reach (box 64 x 39; +-19 = EDGE_THRESHOLD, src/ORBextractor.cc:73) -- the warp waits on the barrier and
// does all its reads from shared memory.  The hardware wants the first byte of a box 16-byte aligned in global memory (measured:
// profiles/tma_probe.cu faults with "illegal instruction" at x = 77 or 100 and works at 96 / 128), so the boxes start at the column
// rounded down to 16 and are 15 columns wider than the patches.  The tensor maps (one per level and image kind: x = column, y = row, z = frame;
// out-of-range coordinates are zero-filled and never read) are encoded on the host per call.  Same arithmetic, same results.
// Needs 16-byte aligned level bases and pitches (always true for the library's own buffers; a caller-owned level 0 with
// another pitch takes k_describe).
#pragma once
#ifndef PLF_EMU
#include <cuda.h>

struct OrbTensorMaps {
    CUtensorMap raw[ORB_MAX_LEVELS];
    CUtensorMap blr[ORB_MAX_LEVELS];
};

#define DT_RAW_W 48
#define DT_RAW_H 31
#define DT_BLR_W 64
#define DT_BLR_H 39
#define DT_RAW_BYTES (DT_RAW_W * DT_RAW_H)      // 1488
#define DT_BLR_BYTES (DT_BLR_W * DT_BLR_H)      // 2496
#define DT_BLR_OFF 1536                         // per warp: raw tile at 0, blurred tile at 1536 (both 128-byte aligned)
#define DT_SLOT 4096

__device__ __forceinline__ unsigned dt_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256)
k_describe_tma(OrbGeom g, OrbPtrs p, const __grid_constant__ OrbTensorMaps tm, plf_keypoint* __restrict__ kps, uint8_t* __restrict__ desc,
               int cap, int* __restrict__ n_out)
{
    __shared__ __align__(128) uint8_t tiles[8][DT_SLOT];
    __shared__ __align__(8) unsigned long long bars[8];
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int wid = blockIdx.x * 8 + wl;
    const int* kc = p.keptcount + (size_t)f * g.nlevels;
    int total = 0, l = -1, pos = 0;
    bool bad = false;
    for (int i = 0; i < g.nlevels; i++) {
        int c = kc[i];
        if (c < 0) { bad = true; c = 0; }
        if (l < 0 && wid < total + c) { l = i; pos = wid - total; }
        total += c;
    }
    if (wid == 0 && lane == 0) n_out[f] = bad ? -1 : (total <= cap ? total : -2);
    if (bad || l < 0 || wid >= cap) return;          // warp-uniform
    const OrbLevelGeom& L = g.lv[l];
    const int kidx = p.kept[(size_t)f * g.keptPerFrame + L.keptOff + pos];
    const unsigned key = p.rawkeys[(size_t)f * g.rawPerFrame + L.rawOff + kidx];
    const int X = (int)(key & 0xfff) + ORB_MINB, Y = (int)((key >> 12) & 0xfff) + ORB_MINB;
    const int resp = key >> 24;
    uint8_t* rawT = tiles[wl];
    uint8_t* blrT = tiles[wl] + DT_BLR_OFF;
    const int xr = (X - 15) & ~15, xb = (X - 19) & ~15;      // box origins; X >= 19 always (FAST starts 19 px inside)
    const unsigned bar = dt_smem_u32(&bars[wl]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(DT_RAW_BYTES + DT_BLR_BYTES) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dt_smem_u32(rawT)), "l"(&tm.raw[l]), "r"(bar), "r"(xr), "r"(Y - 15), "r"(f) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dt_smem_u32(blrT)), "l"(&tm.blr[l]), "r"(bar), "r"(xb), "r"(Y - 19), "r"(f) : "memory");
    }
    __syncwarp();
    // the lane's 16 point pairs (32 signed bytes) as two 16-byte loads, while the tiles are in flight
    signed char pat[32];
    {
        const int4* pp = (const int4*)(d_orb_pattern + lane * 32);
        const int4 q0 = pp[0], q1 = pp[1];
        const int wds[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int i = 0; i < 32; i++) pat[i] = (signed char)((wds[i >> 2] >> (8 * (i & 3))) & 0xff);
    }
    {
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }"
                         : "=r"(done) : "r"(bar), "r"(0) : "memory");
    }
    // orientation on the un-blurred level: row v of the disc, columns -umax[v] .. umax[v]
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - ORB_HALF_PATCH;
        const int d = g.umax[v < 0 ? -v : v];
        const uint8_t* row = rawT + lane * DT_RAW_W + (X - xr);
        int s = 0;
        for (int u = -d; u <= d; u++) {
            const int val = row[u];
            s += val;
            m10 += u * val;
        }
        m01 = v * s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = plf_fast_atan2((float)m01, (float)m10);
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = angle * factorPI;
    const float a = plf_libm::cosf_glibc(ang), b = plf_libm::sinf_glibc(ang);
    const uint8_t* center = blrT + 19 * DT_BLR_W + (X - xb);
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float x0 = (float)pat[4 * k], y0 = (float)pat[4 * k + 1], x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
        float r0 = x0 * b, r1 = y0 * a, r2 = x0 * a, r3 = y0 * b;
        int t0 = center[__float2int_rn(r0 + r1) * DT_BLR_W + __float2int_rn(r2 - r3)];
        r0 = x1 * b; r1 = y1 * a; r2 = x1 * a; r3 = y1 * b;
        int t1 = center[__float2int_rn(r0 + r1) * DT_BLR_W + __float2int_rn(r2 - r3)];
        val |= (t0 < t1) << k;
    }
    desc[((size_t)f * cap + wid) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        plf_keypoint kp;
        kp.x = (float)X; kp.y = (float)Y;
        if (l != 0) { kp.x = kp.x * L.scale; kp.y = kp.y * L.scale; }
        kp.size = (float)L.sizeval;
        kp.angle = angle;
        kp.response = (float)resp;
        kp.octave = l;
        kp.class_id = -1;
        kps[(size_t)f * cap + wid] = kp;
    }
}

// host: encode the per-level maps for this call (x, y, frame); false when a base or pitch is not 16-byte aligned
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
static bool orb_make_tensor_maps(const OrbGeom& g, const OrbPtrs& P, int nframes, OrbTensorMaps* tm)
{
    plf_encode_tiled_fn enc = plf_get_encode_tiled();
    if (!enc) return false;
    for (int l = 0; l < g.nlevels; l++) {
        const OrbLevelGeom& L = g.lv[l];
        for (int kind = 0; kind < 2; kind++) {
            const void* base = kind ? (const void*)P.blr[l] : (const void*)P.lvl[l];
            const size_t pitch = kind ? (size_t)L.pitch : (size_t)P.pitch[l];
            const size_t fstride = kind ? L.frameBytes : P.frameStride[l];
            if (((uintptr_t)base & 15) || (pitch & 15) || (fstride & 15)) return false;
            const cuuint64_t dims[3] = {(cuuint64_t)L.w, (cuuint64_t)L.h, (cuuint64_t)nframes};
            const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fstride};
            const cuuint32_t box[3] = {(cuuint32_t)(kind ? DT_BLR_W : DT_RAW_W), (cuuint32_t)(kind ? DT_BLR_H : DT_RAW_H), 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            if (enc(kind ? &tm->blr[l] : &tm->raw[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return false;
        }
    }
    return true;
}
#endif
