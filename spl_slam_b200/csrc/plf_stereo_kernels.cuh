// plf_stereo_kernels.cuh -- Frame::ComputeStereoMatches (src/Frame.cc:881-1055) on the device.
// One warp per left keypoint: row-band best-1 Hamming over the right keypoints (octave +-1, disparity
// window, init TH_HIGH, first best wins), 11x11 SAD slide over +-5 px on the two extractors' pyramid
// levels, parabola sub-pixel, depth; then one CTA per stereo pair removes matches whose SAD is at least
// 1.5 * 1.4 * median (the median is the element at position size/2 of the sorted (SAD, iL) list).
#pragma once
#include "plf_orb.cuh"

#define STEREO_TH_HIGH 100   // ORBmatcher::TH_HIGH, src/ORBmatcher.cc:38
#define STEREO_TH_LOW 50     // ORBmatcher::TH_LOW,  src/ORBmatcher.cc:39

struct StereoSide {
    const uint8_t* lvl[ORB_MAX_LEVELS];   // pyramid levels of the batch
    size_t frameStride[ORB_MAX_LEVELS];
    int pitch[ORB_MAX_LEVELS];
    int w[ORB_MAX_LEVELS], h[ORB_MAX_LEVELS];
    const plf_keypoint* kps;              // [frame][cap]
    const uint8_t* desc;                  // [frame][cap][32]
    const int* n;                         // [frame]
    int first, step;                      // frame of pair p = first + p * step
};

struct StereoTables { float scale[ORB_MAX_LEVELS], inv_scale[ORB_MAX_LEVELS]; int nlevels; };

__global__ void __launch_bounds__(256)
k_stereo_match(StereoSide SL, StereoSide SR, StereoTables T, int cap, float mb, float mbf,
               float* __restrict__ uRight, float* __restrict__ depth, int* __restrict__ sad)
{
    const int pair = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int iL = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int fL = SL.first + pair * SL.step, fR = SR.first + pair * SR.step;
    int nL = SL.n[fL], nR = SR.n[fR];
    if (nL > cap) nL = cap;
    if (nR > cap) nR = cap;
    if (iL >= cap) return;
    const size_t o = (size_t)pair * cap + iL;
    if (lane == 0) { uRight[o] = -1.0f; depth[o] = -1.0f; sad[o] = -1; }
    if (iL >= nL) return;
    const plf_keypoint kl = SL.kps[(size_t)fL * cap + iL];
    const int levelL = kl.octave;
    const float uL = kl.x, vL = kl.y;
    const int nRows = SL.h[0];
    const int row = (int)vL;
    if (row < 0 || row >= nRows || levelL < 0 || levelL >= T.nlevels) return;
    const float minZ = mb, minD = 0.f, maxD = mbf / minZ;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) return;
    const uint4* dl = (const uint4*)(SL.desc + ((size_t)fL * cap + iL) * 32);
    const uint4 a0 = dl[0], a1 = dl[1];
    const plf_keypoint* KR = SR.kps + (size_t)fR * cap;
    const uint4* DR = (const uint4*)(SR.desc + (size_t)fR * cap * 32);
    // best = lexicographic minimum of (distance, iR) among candidates with distance < TH_HIGH
    int best = (STEREO_TH_HIGH << 16) | 0xffff;
    for (int iR = lane; iR < nR; iR += 32) {
        const plf_keypoint kr = KR[iR];
        if (kr.octave < 0 || kr.octave >= T.nlevels) continue;
        const float r = 2.0f * T.scale[kr.octave];
        const int maxr = (int)ceilf(kr.y + r), minr = (int)floorf(kr.y - r);   // vRowIndices band, :898-908
        if (row < minr || row > maxr) continue;
        if (kr.octave < levelL - 1 || kr.octave > levelL + 1) continue;
        if (kr.x >= minU && kr.x <= maxU) {
            const uint4 b0 = DR[2 * iR], b1 = DR[2 * iR + 1];
            const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                          __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            if (d < STEREO_TH_HIGH) best = min(best, (d << 16) | iR);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, s));
    const int bestDist = best >> 16, bestIdxR = best & 0xffff;
    const int thOrbDist = (STEREO_TH_HIGH + STEREO_TH_LOW) / 2;
    if (!(bestDist < thOrbDist) || bestIdxR >= nR) return;
    // sub-pixel match by correlation, :967-1020
    const float uR0 = KR[bestIdxR].x;
    const float scaleFactor = T.inv_scale[levelL];
    const float scaleduL = roundf(kl.x * scaleFactor), scaledvL = roundf(kl.y * scaleFactor);
    const float scaleduR0 = roundf(uR0 * scaleFactor);
    const int w = 5, L = 5;
    const int lw = SL.w[levelL], lh = SL.h[levelL], rw = SR.w[levelL], rh = SR.h[levelL];
    const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
    if (iniu < 0 || endu >= rw) return;
    const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
    // windows leaving the level image make the reference throw (cv::Mat::colRange); skipped here
    if (cv - w < 0 || cv + w >= lh || cu - w < 0 || cu + w >= lw || cr - L - w < 0 || cr + L + w >= rw || cv + w >= rh) return;
    const uint8_t* imL = SL.lvl[levelL] + (size_t)fL * SL.frameStride[levelL];
    const uint8_t* imR = SR.lvl[levelL] + (size_t)fR * SR.frameStride[levelL];
    const int pl = SL.pitch[levelL], pr = SR.pitch[levelL];
    const int cL = imL[(size_t)cv * pl + cu];
    int cR[11], acc[11];
#pragma unroll
    for (int k = 0; k < 11; k++) { cR[k] = imR[(size_t)cv * pr + cr + k - L]; acc[k] = 0; }
    for (int i = lane; i < 121; i += 32) {
        const int dy = i / 11 - w, dx = i % 11 - w;
        const int a = imL[(size_t)(cv + dy) * pl + cu + dx] - cL;
        const uint8_t* rr = imR + (size_t)(cv + dy) * pr + cr + dx - L;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const int b = rr[k] - cR[k];
            const int d = a - b;
            acc[k] += d < 0 ? -d : d;
        }
    }
#pragma unroll
    for (int k = 0; k < 11; k++)
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], s);
    if (lane != 0) return;
    int bestSad = 2147483647, bestincR = 0;
#pragma unroll
    for (int k = 0; k < 11; k++)
        if (acc[k] < bestSad) { bestSad = acc[k]; bestincR = k - L; }
    if (bestincR == -L || bestincR == L) return;
    float dist1 = 0, dist2 = 0, dist3 = 0;
#pragma unroll
    for (int k = 1; k < 10; k++)
        if (k == bestincR + L) { dist1 = (float)acc[k - 1]; dist2 = (float)acc[k]; dist3 = (float)acc[k + 1]; }
    const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
    if (deltaR < -1 || deltaR > 1) return;
    float bestuR = T.scale[levelL] * (scaleduR0 + (float)bestincR + deltaR);
    float disparity = uL - bestuR;
    if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) { disparity = (float)0.01; bestuR = (float)((double)uL - 0.01); }
        depth[o] = mbf / disparity;
        uRight[o] = bestuR;
        sad[o] = bestSad;
    }
}

// median SAD filter (:1041-1054): one CTA per pair.  The median is the SAD of the entry at sorted position
// np/2; ties do not change its value, so a rank by (SAD, iL) selects it.
__global__ void __launch_bounds__(256)
k_stereo_filter(int cap, float* __restrict__ uRight, float* __restrict__ depth, const int* __restrict__ sad)
{
    PLF_DYN_SMEM(smem);
    int* s = (int*)smem;
    __shared__ int s_np, s_med;
    const int pair = blockIdx.x, tid = threadIdx.x;
    const int* S = sad + (size_t)pair * cap;
    if (tid == 0) { s_np = 0; s_med = -1; }
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < cap; i += 256) {
        const int v = S[i];
        s[i] = v;
        if (v >= 0) mine++;
    }
    if (mine) atomicAdd(&s_np, mine);
    __syncthreads();
    const int np = s_np;
    if (np == 0) return;   // the reference indexes an empty vector here (undefined); defined as a no-op
    const int target = np / 2;
    for (int i = tid; i < cap; i += 256) {
        const int v = s[i];
        if (v < 0) continue;
        int rank = 0;
        for (int j = 0; j < cap; j++) {
            const int u = s[j];
            if (u >= 0 && (u < v || (u == v && j < i))) rank++;
        }
        if (rank == target) s_med = v;
    }
    __syncthreads();
    const float median = (float)s_med;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = tid; i < cap; i += 256) {
        const int v = s[i];
        if (v >= 0 && !((float)v < thDist)) {
            uRight[(size_t)pair * cap + i] = -1.0f;
            depth[(size_t)pair * cap + i] = -1.0f;
        }
    }
}
