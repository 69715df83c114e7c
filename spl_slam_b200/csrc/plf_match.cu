// plf_match.cu -- brute-force 256-bit Hamming top-2 + ratio test.
// Replaces cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) as called by Linematcher::matchNNR
// (src/Linematcher.cc:520-541) and the DescriptorDistance loops (src/Linematcher.cc:50-66,
// src/ORBmatcher.cc:1656-1672).  Integer XOR + POPC on the CUDA cores; no tensor cores (this is not a
// float contraction).  Roofline: POPC issue rate, 8 popc32 per descriptor pair.
#include "plf_common.cuh"

#define KNN_THREADS 128
#define KNN_QPT 4           // queries per thread
#define KNN_TILE 256        // train descriptors staged in shared memory per step
#define KNN_INF 0x7fffffff

// 256-bit Hamming distance with FOUR population counts instead of eight.  POPC issues at a quarter of the rate of the
// logic / add instructions (16 vs 64 lanes per clock per SM), so the eight XOR words are first compressed with carry-save
// adders -- sum = a ^ b ^ c and carry = maj(a, b, c) are one LOP3 each:
//   x0..x6 -> (s_c; c_a, c_b, c_c),  (c_a, c_b, c_c) -> (t_s; t_c),   d = popc(s_c) + popc(x7) + 2 popc(t_s) + 4 popc(t_c)
// 8 XOR + 8 LOP3 + 3 adds + 4 POPC per pair: the POPC pipe (32 issue cycles per warp and sub-partition) and the integer pipe
// (38) are balanced, against 64 POPC cycles for the plain form.  Exact (an identity on bit counts).
__device__ __forceinline__ void csa(unsigned a, unsigned b, unsigned c, unsigned& sum, unsigned& carry)
{
    sum = a ^ b ^ c;
    carry = (a & b) | (c & (a ^ b));
}
__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1)
{
    const unsigned x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const unsigned x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
    unsigned sa, ca, sb, cb, sc, cc, ts, tc;
    csa(x0, x1, x2, sa, ca);
    csa(x3, x4, x5, sb, cb);
    csa(sa, sb, x6, sc, cc);
    csa(ca, cb, cc, ts, tc);
    return __popc(sc) + __popc(x7) + 2 * __popc(ts) + 4 * __popc(tc);
}

// grid = (query tiles, train splits).  Each thread keeps KNN_QPT queries in registers and scans the
// split's train rows, staged through shared memory and read back as warp-wide broadcasts.
// Scanning in ascending train index with strict '<' keeps the lowest index on distance ties
// (the BFMatcher ordering).
__global__ void __launch_bounds__(KNN_THREADS)
knn2_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, long long nt, long long chunk,
            long long index_base, int* __restrict__ pidx, int* __restrict__ pdist)
{
    __shared__ uint4 tile[KNN_TILE * 2];
    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * (KNN_THREADS * KNN_QPT);
    const long long tbeg = (long long)blockIdx.y * chunk;
    long long tend = tbeg + chunk;
    if (tend > nt) tend = nt;

    uint4 qa[KNN_QPT], qb[KNN_QPT];
    int d0[KNN_QPT], d1[KNN_QPT], i0[KNN_QPT], i1[KNN_QPT];
#pragma unroll
    for (int j = 0; j < KNN_QPT; j++) {
        int qi = q0 + tid + j * KNN_THREADS;
        if (qi < nq) { qa[j] = q[2 * (size_t)qi]; qb[j] = q[2 * (size_t)qi + 1]; }
        else { qa[j] = make_uint4(0, 0, 0, 0); qb[j] = qa[j]; }
        d0[j] = d1[j] = KNN_INF;
        i0[j] = i1[j] = -1;
    }
    for (long long base = tbeg; base < tend; base += KNN_TILE) {
        long long rem = tend - base;
        int n = rem < KNN_TILE ? (int)rem : KNN_TILE;
        __syncthreads();
        for (int k = tid; k < 2 * n; k += KNN_THREADS) tile[k] = t[2 * base + k];
        __syncthreads();
        int gidx = (int)(index_base + base);
#pragma unroll 4
        for (int e = 0; e < n; e++) {
            uint4 a = tile[2 * e], b = tile[2 * e + 1];
#pragma unroll
            for (int j = 0; j < KNN_QPT; j++) {
                int d = hamming256(qa[j], qb[j], a, b);
                if (d < d1[j]) {
                    if (d < d0[j]) { d1[j] = d0[j]; i1[j] = i0[j]; d0[j] = d; i0[j] = gidx + e; }
                    else { d1[j] = d; i1[j] = gidx + e; }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < KNN_QPT; j++) {
        int qi = q0 + tid + j * KNN_THREADS;
        if (qi < nq) {
            size_t o = ((size_t)blockIdx.y * nq + qi) * 2;
            pidx[o] = i0[j]; pidx[o + 1] = i1[j];
            pdist[o] = i0[j] >= 0 ? d0[j] : -1;
            pdist[o + 1] = i1[j] >= 0 ? d1[j] : -1;
        }
    }
}

// Merge nshards partial top-2 tables by (distance, index) lexicographic order: identical to a single
// scan over the whole train set, including tie-breaks.
__global__ void knn2_merge_kernel(const int* __restrict__ pidx, const int* __restrict__ pdist, int nshards, int nq,
                                  int* __restrict__ idx, int* __restrict__ dist)
{
    int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    int d0 = KNN_INF, d1 = KNN_INF, i0 = -1, i1 = -1;
    for (int s = 0; s < nshards; s++) {
        size_t o = ((size_t)s * nq + qi) * 2;
        for (int k = 0; k < 2; k++) {
            int i = pidx[o + k], d = pdist[o + k];
            if (i < 0) continue;
            bool lt0 = d < d0 || (d == d0 && (i0 < 0 || i < i0));
            bool lt1 = d < d1 || (d == d1 && (i1 < 0 || i < i1));
            if (lt0) { d1 = d0; i1 = i0; d0 = d; i0 = i; }
            else if (lt1) { d1 = d; i1 = i; }
        }
    }
    idx[2 * (size_t)qi] = i0; idx[2 * (size_t)qi + 1] = i1;
    dist[2 * (size_t)qi] = i0 >= 0 ? d0 : -1;
    dist[2 * (size_t)qi + 1] = i1 >= 0 ? d1 : -1;
}

// Linematcher::matchNNR ratio test (src/Linematcher.cc:534-538): float compare d0 < d1 * nnr.
__global__ void nnr_kernel(const int* __restrict__ idx, const int* __restrict__ dist, int nq, float nnr,
                           int* __restrict__ m12, int* __restrict__ nmatches)
{
    int qi = blockIdx.x * blockDim.x + threadIdx.x;
    int ok = 0;
    if (qi < nq) {
        int r = -1;
        if (idx[2 * qi + 1] >= 0 && (float)dist[2 * qi] < (float)dist[2 * qi + 1] * nnr) { r = idx[2 * qi]; ok = 1; }
        m12[qi] = r;
    }
    unsigned b = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(nmatches, __popc(b));
}

// mutual-consistency filter of Linematcher::SearchByKNN (src/Linematcher.cc:460-471)
__global__ void mutual_kernel(int* __restrict__ m12, const int* __restrict__ m21, int n1, int n2, int* __restrict__ nmatches)
{
    int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    int drop = 0;
    if (i1 < n1) {
        int i2 = m12[i1];
        if (i2 >= 0 && (i2 >= n2 || m21[i2] != i1)) { m12[i1] = -1; drop = 1; }
    }
    unsigned b = __ballot_sync(0xffffffffu, drop);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(nmatches, -__popc(b));
}

__global__ void pair_distance_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, int n, int* __restrict__ d)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = hamming256(a[2 * (size_t)i], a[2 * (size_t)i + 1], b[2 * (size_t)i], b[2 * (size_t)i + 1]);
}

// ---------------------------------------------------------------------------------------------
static int knn2_splits(int nq, long long nt)
{
    int qtiles = plf_div_up(nq, KNN_THREADS * KNN_QPT);
    if (qtiles < 1) qtiles = 1;
    // aim at >= 4 CTAs per SM on 148 SMs, but keep >= 4 tiles of train rows per split
    int want = plf_div_up(148 * 4, qtiles);
    long long maxs = nt / (4 * KNN_TILE);
    if (maxs < 1) maxs = 1;
    if (want > maxs) want = (int)maxs;
    if (want < 1) want = 1;
    if (want > 1024) want = 1024;
    return want;
}

static plf_status knn2_device(plf_ctx* ctx, const uint8_t* dq, int nq, const uint8_t* dt, long long nt, long long base,
                              int* didx, int* ddist)
{
    if (nq <= 0) return PLF_OK;
    if (base + nt > 0x7fffffffLL) return plf_fail(ctx, PLF_ERR_INVALID, "train index exceeds int32 range");
    int splits = knn2_splits(nq, nt);
    long long chunk = nt > 0 ? (nt + splits - 1) / splits : 1;
    chunk = (chunk + KNN_TILE - 1) / KNN_TILE * KNN_TILE;
    splits = nt > 0 ? (int)((nt + chunk - 1) / chunk) : 1;
    int qtiles = plf_div_up(nq, KNN_THREADS * KNN_QPT);
    int *pidx = didx, *pdist = ddist;
    if (splits > 1) {
        void* s;
        plf_status st = plf_ctx_scratch(ctx, (size_t)splits * nq * 2 * sizeof(int) * 2, &s);
        if (st) return st;
        pidx = (int*)s;
        pdist = pidx + (size_t)splits * nq * 2;
    }
    PLF_LAUNCH(knn2_kernel, dim3(qtiles, splits), dim3(KNN_THREADS), 0, ctx->stream, (const uint4*)dq, nq,
               (const uint4*)dt, nt, chunk, base, pidx, pdist);
    PLF_CHECK_LAUNCH(ctx);
    if (splits > 1) {
        PLF_LAUNCH(knn2_merge_kernel, dim3(plf_div_up(nq, 128)), dim3(128), 0, ctx->stream, pidx, pdist, splits, nq,
                   didx, ddist);
        PLF_CHECK_LAUNCH(ctx);
    }
    return PLF_OK;
}

extern "C" plf_status plf_hamming_knn2_device(plf_ctx* ctx, const uint8_t* dq, int nq, const uint8_t* dt, int64_t nt,
                                              int64_t train_index_base, int32_t* didx, int32_t* ddist)
{
    if (!ctx || nq < 0 || nt < 0 || (nq > 0 && (!dq || !didx || !ddist)) || (nt > 0 && !dt))
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_hamming_knn2_device: bad arguments");
    if (((uintptr_t)dq | (uintptr_t)dt) & 15) return plf_fail(ctx, PLF_ERR_INVALID, "descriptor pointers must be 16-byte aligned");
    return knn2_device(ctx, dq, nq, dt, nt, train_index_base, didx, ddist);
}

extern "C" plf_status plf_knn2_merge_device(plf_ctx* ctx, const int32_t* pidx, const int32_t* pdist, int nshards, int nq,
                                            int32_t* didx, int32_t* ddist)
{
    if (!ctx || nshards < 1 || nq < 0) return plf_fail(ctx, PLF_ERR_INVALID, "plf_knn2_merge_device: bad arguments");
    if (nq == 0) return PLF_OK;
    PLF_LAUNCH(knn2_merge_kernel, dim3(plf_div_up(nq, 128)), dim3(128), 0, ctx->stream, pidx, pdist, nshards, nq, didx, ddist);
    PLF_CHECK_LAUNCH(ctx);
    return PLF_OK;
}

extern "C" plf_status plf_nnr_from_knn2_device(plf_ctx* ctx, const int32_t* didx, const int32_t* ddist, int nq, float nnr,
                                               int32_t* dm12, int32_t* dnm)
{
    if (!ctx || nq < 0) return plf_fail(ctx, PLF_ERR_INVALID, "plf_nnr_from_knn2_device: bad arguments");
    PLF_CUDA(ctx, cudaMemsetAsync(dnm, 0, sizeof(int), ctx->stream));
    if (nq == 0) return PLF_OK;
    PLF_LAUNCH(nnr_kernel, dim3(plf_div_up(nq, 128)), dim3(128), 0, ctx->stream, didx, ddist, nq, nnr, dm12, dnm);
    PLF_CHECK_LAUNCH(ctx);
    return PLF_OK;
}

// device layout helper for the host-buffer entry points
struct match_bufs {
    uint8_t *q, *t;
    int *idx, *dist, *m12, *m21, *nm;
};

// device copies of host-buffer arguments come out of the context's I/O scratch (no cudaMalloc / cudaFree per call: both
// synchronise the whole device and would stall the other contexts' streams; matchNNR is a per-frame-pair call in SLAM)
static inline size_t desc_bytes(size_t rows) { return plf_align_up(rows ? rows * 32 : 32, 256); }
static plf_status upload_desc(plf_ctx* ctx, uint8_t* dev, const uint8_t* host, size_t rows)
{
    if (rows) PLF_CUDA(ctx, cudaMemcpyAsync(dev, host, rows * 32, cudaMemcpyHostToDevice, ctx->stream));
    return PLF_OK;
}

extern "C" plf_status plf_hamming_knn2(plf_ctx* ctx, const uint8_t* hq, int nq, const uint8_t* ht, int64_t nt,
                                       int32_t* hidx, int32_t* hdist)
{
    if (!ctx || nq < 0 || nt < 0 || (nq > 0 && (!hq || !hidx || !hdist)) || (nt > 0 && !ht))
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_hamming_knn2: bad arguments");
    if (nq == 0) return PLF_OK;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    void* io = nullptr;
    plf_status st = plf_ctx_ioscratch(ctx, desc_bytes((size_t)nq) + desc_bytes((size_t)nt) + (size_t)nq * 4 * sizeof(int), &io);
    if (st) return st;
    uint8_t *dq = (uint8_t*)io, *dt = dq + desc_bytes((size_t)nq);
    int* dout = (int*)(dt + desc_bytes((size_t)nt));
    st = upload_desc(ctx, dq, hq, (size_t)nq);
    if (!st) st = upload_desc(ctx, dt, ht, (size_t)nt);
    if (!st) st = knn2_device(ctx, dq, nq, dt, nt, 0, dout, dout + 2 * (size_t)nq);
    if (!st) {
        cudaError_t e = cudaMemcpyAsync(hidx, dout, (size_t)nq * 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hdist, dout + 2 * (size_t)nq, (size_t)nq * 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = plf_fail(ctx, PLF_ERR_CUDA, "knn2 result copy failed: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(ctx->stream);
    }
    return st;
}

static plf_status match_nnr_impl(plf_ctx* ctx, const uint8_t* h1, int n1, const uint8_t* h2, int64_t n2, float nnr,
                                 int32_t* hm12, int* nmatches, bool mutual)
{
    if (!ctx || n1 < 0 || n2 < 0 || (n1 > 0 && (!h1 || !hm12)) || (n2 > 0 && !h2) || !nmatches)
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_match_nnr: bad arguments");
    *nmatches = 0;
    if (n1 == 0) return PLF_OK;
    if (mutual && n2 > 0x7fffffff) return plf_fail(ctx, PLF_ERR_INVALID, "mutual matching needs n2 < 2^31");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t nmax = (size_t)(n1 > n2 ? n1 : n2);
    // layout: idx[2*nmax] dist[2*nmax] m12[n1] m21[n2] nm[2]
    size_t words = 4 * nmax + (size_t)n1 + (size_t)n2 + 2;
    void* io = nullptr;
    plf_status st = plf_ctx_ioscratch(ctx, desc_bytes((size_t)n1) + desc_bytes((size_t)n2) + words * sizeof(int), &io);
    if (st) return st;
    uint8_t *d1 = (uint8_t*)io, *d2 = d1 + desc_bytes((size_t)n1);
    int* w = (int*)(d2 + desc_bytes((size_t)n2));
    st = upload_desc(ctx, d1, h1, (size_t)n1);
    if (!st) st = upload_desc(ctx, d2, h2, (size_t)n2);
    int *idx = w, *dist = w + 2 * nmax, *m12 = w + 4 * nmax, *m21 = m12 + n1, *nm = m21 + n2;
    if (!st) st = knn2_device(ctx, d1, n1, d2, n2, 0, idx, dist);
    if (!st) st = plf_nnr_from_knn2_device(ctx, idx, dist, n1, nnr, m12, nm);
    if (!st && mutual && n2 > 0) {
        st = knn2_device(ctx, d2, (int)n2, d1, n1, 0, idx, dist);
        if (!st) st = plf_nnr_from_knn2_device(ctx, idx, dist, (int)n2, nnr, m21, nm + 1);
        if (!st) {
            PLF_LAUNCH(mutual_kernel, dim3(plf_div_up(n1, 128)), dim3(128), 0, ctx->stream, m12, (const int*)m21, n1, (int)n2, nm);
            ctx->launches++;
        }
    }
    if (!st) {
        cudaError_t e = cudaMemcpyAsync(hm12, m12, (size_t)n1 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nmatches, nm, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = plf_fail(ctx, PLF_ERR_CUDA, "match result copy failed: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(ctx->stream);
    }
    return st;
}

extern "C" plf_status plf_match_nnr(plf_ctx* ctx, const uint8_t* hq, int nq, const uint8_t* ht, int64_t nt, float nnr,
                                    int32_t* hm12, int* nmatches)
{
    return match_nnr_impl(ctx, hq, nq, ht, nt, nnr, hm12, nmatches, false);
}

extern "C" plf_status plf_match_nnr_mutual(plf_ctx* ctx, const uint8_t* h1, int n1, const uint8_t* h2, int n2, float nnr,
                                           int32_t* hm12, int* nmatches)
{
    return match_nnr_impl(ctx, h1, n1, h2, n2, nnr, hm12, nmatches, true);
}

extern "C" plf_status plf_descriptor_distance(plf_ctx* ctx, const uint8_t* ha, const uint8_t* hb, int n, int32_t* hd)
{
    if (!ctx || n < 0 || (n > 0 && (!ha || !hb || !hd))) return plf_fail(ctx, PLF_ERR_INVALID, "plf_descriptor_distance: bad arguments");
    if (n == 0) return PLF_OK;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    void* io = nullptr;
    plf_status st = plf_ctx_ioscratch(ctx, 2 * desc_bytes((size_t)n) + (size_t)n * sizeof(int), &io);
    if (st) return st;
    uint8_t *da = (uint8_t*)io, *db = da + desc_bytes((size_t)n);
    int* dd = (int*)(db + desc_bytes((size_t)n));
    st = upload_desc(ctx, da, ha, (size_t)n);
    if (!st) st = upload_desc(ctx, db, hb, (size_t)n);
    if (!st) {
        PLF_LAUNCH(pair_distance_kernel, dim3(plf_div_up(n, 128)), dim3(128), 0, ctx->stream, (const uint4*)da, (const uint4*)db, n, dd);
        ctx->launches++;
        cudaError_t e = cudaMemcpyAsync(hd, dd, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) st = plf_fail(ctx, PLF_ERR_CUDA, "distance copy failed: %s", cudaGetErrorString(e));
    }
    return st;
}

// ---- candidate-list top-2 (ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:430-456, and the other
// Search* loops): one warp per query; lane l takes list positions l, l+32, ...; the per-lane and cross-lane
// merges order entries by (distance, list position), which is what the reference's sequential
// `dist < bestDist` / `else if (dist < bestDist2)` scan produces. ----
__global__ void __launch_bounds__(256)
cand_top2_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, const int* __restrict__ off,
                 const int* __restrict__ cidx, int* __restrict__ bidx, int* __restrict__ bdist, int* __restrict__ cdist)
{
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const uint4 a0 = q[2 * (size_t)qi], a1 = q[2 * (size_t)qi + 1];
    const int beg = off[qi], end = off[qi + 1];
    // keys: distance << 22 | list position (lists up to 4M entries); KNN_INF = none
    long long k0 = 0x7fffffffffffffffLL, k1 = 0x7fffffffffffffffLL;
    for (int c = beg + lane; c < end; c += 32) {
        const int ti = cidx[c];
        int d = -1;
        if (ti >= 0 && ti < nt) {
            const uint4 b0 = t[2 * (size_t)ti], b1 = t[2 * (size_t)ti + 1];
            d = hamming256(a0, a1, b0, b1);
            const long long key = ((long long)d << 32) | (unsigned)(c - beg);
            if (key < k0) { k1 = k0; k0 = key; }
            else if (key < k1) k1 = key;
        }
        if (cdist) cdist[c] = d;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const long long o0 = __shfl_xor_sync(0xffffffffu, k0, s), o1 = __shfl_xor_sync(0xffffffffu, k1, s);
        // merge two sorted pairs
        const long long lo = k0 < o0 ? k0 : o0, hi = k0 < o0 ? o0 : k0;
        const long long m1 = k1 < o1 ? k1 : o1;
        k0 = lo;
        k1 = hi < m1 ? hi : m1;
    }
    if (lane == 0) {
        const bool h0 = k0 != 0x7fffffffffffffffLL, h1 = k1 != 0x7fffffffffffffffLL;
        bidx[2 * (size_t)qi] = h0 ? cidx[beg + (int)(k0 & 0xffffffffLL)] : -1;
        bidx[2 * (size_t)qi + 1] = h1 ? cidx[beg + (int)(k1 & 0xffffffffLL)] : -1;
        bdist[2 * (size_t)qi] = h0 ? (int)(k0 >> 32) : -1;
        bdist[2 * (size_t)qi + 1] = h1 ? (int)(k1 >> 32) : -1;
    }
}

extern "C" plf_status plf_hamming_candidates_device(plf_ctx* ctx, const uint8_t* dq, int nq, const uint8_t* dt, int nt, const int32_t* doff,
                                                    const int32_t* dcidx, int32_t* dbidx, int32_t* dbdist, int32_t* dcdist)
{
    if (!ctx || nq < 0 || nt < 0 || (nq > 0 && (!dq || !doff || !dbidx || !dbdist)))
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_hamming_candidates_device: bad arguments");
    if (((uintptr_t)dq | (uintptr_t)dt) & 15) return plf_fail(ctx, PLF_ERR_INVALID, "descriptor pointers must be 16-byte aligned");
    if (nq == 0) return PLF_OK;
    PLF_LAUNCH(cand_top2_kernel, dim3(plf_div_up(nq, 8)), dim3(256), 0, ctx->stream, (const uint4*)dq, nq, (const uint4*)dt, nt, doff, dcidx,
               dbidx, dbdist, dcdist);
    PLF_CHECK_LAUNCH(ctx);
    return PLF_OK;
}

extern "C" plf_status plf_hamming_candidates(plf_ctx* ctx, const uint8_t* hq, int nq, const uint8_t* ht, int nt, const int32_t* hoff,
                                             const int32_t* hcidx, int32_t* hbidx, int32_t* hbdist, int32_t* hcdist)
{
    if (!ctx || nq < 0 || nt < 0 || (nq > 0 && (!hq || !hoff || !hbidx || !hbdist)) || (nt > 0 && !ht))
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_hamming_candidates: bad arguments");
    if (nq == 0) return PLF_OK;
    for (int i = 0; i < nq; i++)
        if (hoff[i] < 0 || hoff[i + 1] < hoff[i]) return plf_fail(ctx, PLF_ERR_INVALID, "cand_off must be non-decreasing");
    const int ncand = hoff[nq];
    if (ncand > 0 && !hcidx) return plf_fail(ctx, PLF_ERR_INVALID, "plf_hamming_candidates: cand_idx is NULL");
    for (int c = hoff[0]; c < ncand; c++)
        if (hcidx[c] < 0 || hcidx[c] >= nt) return plf_fail(ctx, PLF_ERR_INVALID, "candidate %d: train index %d out of range", c, hcidx[c]);
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t qb = plf_align_up((size_t)nq * 32, 256), tb = plf_align_up((size_t)(nt > 0 ? nt : 1) * 32, 256);
    const size_t ob = plf_align_up((size_t)(nq + 1) * 4, 256), cb = plf_align_up((size_t)(ncand > 0 ? ncand : 1) * 4, 256);
    const size_t bb = plf_align_up((size_t)nq * 2 * 4, 256);
    void* s;
    plf_status st = plf_ctx_scratch(ctx, qb + tb + ob + 2 * cb + 2 * bb, &s);
    if (st) return st;
    uint8_t* p = (uint8_t*)s;
    uint8_t* dq = p; p += qb;
    uint8_t* dt = p; p += tb;
    int* doff = (int*)p; p += ob;
    int* dcidx = (int*)p; p += cb;
    int* dcdist = (int*)p; p += cb;
    int* dbidx = (int*)p; p += bb;
    int* dbdist = (int*)p;
    cudaStream_t sq = ctx->stream;
    PLF_CUDA(ctx, cudaMemcpyAsync(dq, hq, (size_t)nq * 32, cudaMemcpyHostToDevice, sq));
    if (nt > 0) PLF_CUDA(ctx, cudaMemcpyAsync(dt, ht, (size_t)nt * 32, cudaMemcpyHostToDevice, sq));
    PLF_CUDA(ctx, cudaMemcpyAsync(doff, hoff, (size_t)(nq + 1) * 4, cudaMemcpyHostToDevice, sq));
    if (ncand > 0) PLF_CUDA(ctx, cudaMemcpyAsync(dcidx, hcidx, (size_t)ncand * 4, cudaMemcpyHostToDevice, sq));
    st = plf_hamming_candidates_device(ctx, dq, nq, dt, nt, doff, dcidx, dbidx, dbdist, hcdist ? dcdist : nullptr);
    if (st) return st;
    PLF_CUDA(ctx, cudaMemcpyAsync(hbidx, dbidx, (size_t)nq * 2 * 4, cudaMemcpyDeviceToHost, sq));
    PLF_CUDA(ctx, cudaMemcpyAsync(hbdist, dbdist, (size_t)nq * 2 * 4, cudaMemcpyDeviceToHost, sq));
    if (hcdist && ncand > 0) PLF_CUDA(ctx, cudaMemcpyAsync(hcdist, dcdist, (size_t)ncand * 4, cudaMemcpyDeviceToHost, sq));
    PLF_CUDA(ctx, cudaStreamSynchronize(sq));
    return PLF_OK;
}

// ---- POPC issue-rate micro-benchmark: the roofline denominator for the matching kernels ----
__global__ void __launch_bounds__(256)
popc_peak_kernel(unsigned* __restrict__ out, int iters)
{
    unsigned x0 = threadIdx.x * 2654435761u + blockIdx.x, x1 = x0 ^ 0x9e3779b9u, x2 = x0 * 3u, x3 = x1 * 5u;
    unsigned x4 = x0 + 17u, x5 = x1 + 31u, x6 = x2 ^ 0x55555555u, x7 = x3 ^ 0xaaaaaaaau;
    unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        a0 += __popc(x0 ^ i); a1 += __popc(x1 ^ i); a2 += __popc(x2 ^ i); a3 += __popc(x3 ^ i);
        a4 += __popc(x4 ^ i); a5 += __popc(x5 ^ i); a6 += __popc(x6 ^ i); a7 += __popc(x7 ^ i);
    }
    out[blockIdx.x * 256 + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// measured popc32 results per second of this GPU (XOR + POPC + ADD per result, like the Hamming inner loop)
extern "C" plf_status plf_popc_peak(plf_ctx* ctx, double* popc_per_s)
{
    if (!ctx || !popc_per_s) return PLF_ERR_INVALID;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    const int blocks = 148 * 8, iters = 1 << 14;
    void* s;
    plf_status st = plf_ctx_scratch(ctx, (size_t)blocks * 256 * sizeof(unsigned), &s);
    if (st) return st;
    cudaEvent_t e0, e1;
    PLF_CUDA(ctx, cudaEventCreate(&e0));
    PLF_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        PLF_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        PLF_LAUNCH(popc_peak_kernel, dim3(blocks), dim3(256), 0, ctx->stream, (unsigned*)s, iters);
        PLF_CHECK_LAUNCH(ctx);
        PLF_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        PLF_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        PLF_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *popc_per_s = (double)blocks * 256 * 8.0 * iters / (best * 1e-3);
    return PLF_OK;
}
