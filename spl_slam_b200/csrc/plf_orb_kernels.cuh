// plf_orb_kernels.cuh -- device kernels of the ORB path.
// Integer work is bit-exact by construction; the float pieces (fastAtan2, rBRIEF rotation) use IEEE
// single ops without FMA (-fmad=false) and round-half-even conversions, like the oracle.
#pragma once
#include "plf_orb.cuh"

__device__ __align__(16) const signed char d_orb_pattern[1024] = {
#include "orb_pattern.inc"
};

// ------------------------------------------------------------------------------------------------
// ComputePyramid: cv::resize INTER_LINEAR 8U, level l from level l-1 (src/ORBextractor.cc:1120).
// tab entries: .x = source offset, .y = coef0 | coef1 << 16 (11-bit fixed point).
// block = (32, 8): each thread produces 4 adjacent pixels of RL_ROWS rows.
// ------------------------------------------------------------------------------------------------
#define RL_ROWS 2
__global__ void __launch_bounds__(256)
k_resize_linear(const uint8_t* __restrict__ src, size_t srcFrameStride, int spitch, int sw, int sh,
                uint8_t* __restrict__ dst, size_t dstFrameStride, int dpitch, int dw, int dh,
                const int2* __restrict__ xtab, const int2* __restrict__ ytab)
{
    // a thread produces 4 adjacent pixels of RL_ROWS rows (8 apart): the column table entries are loaded once
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (x0 >= dw) return;
    const uint8_t* s = src + (size_t)blockIdx.z * srcFrameStride;
    int sx[4], sx1[4], a0[4], a1[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int2 tx = xtab[min(x0 + k, dw - 1)];
        sx[k] = tx.x; sx1[k] = min(tx.x + 1, sw - 1);
        a0[k] = (short)(tx.y & 0xffff); a1[k] = (short)(tx.y >> 16);
    }
#pragma unroll
    for (int rr = 0; rr < RL_ROWS; rr++) {
        const int y = (blockIdx.y * RL_ROWS + rr) * 8 + threadIdx.y;
        if (y >= dh) break;
        uint8_t* d = dst + (size_t)blockIdx.z * dstFrameStride + (size_t)y * dpitch;
        const int2 ty = ytab[y];
        const int sy0 = ty.x, sy1 = min(sy0 + 1, sh - 1);
        const int b0 = (short)(ty.y & 0xffff), b1 = (short)(ty.y >> 16);
        const uint8_t* r0 = s + (size_t)sy0 * spitch;
        const uint8_t* r1 = s + (size_t)sy1 * spitch;
        unsigned out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int h0 = r0[sx[k]] * a0[k] + r0[sx1[k]] * a1[k];
            const int h1 = r1[sx[k]] * a0[k] + r1[sx1[k]] * a1[k];
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            out |= (unsigned)(v & 0xff) << (8 * k);
        }
        if (x0 + 3 < dw) {
            *(unsigned*)(d + x0) = out;  // dpitch and x0 are multiples of 4
        } else {
            for (int k = 0; k < 4 && x0 + k < dw; k++) d[x0 + k] = (uint8_t)(out >> (8 * k));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// GaussianBlur 7x7 sigma 2, BORDER_REFLECT_101, Q8 kernel [18 34 48 56 48 34 18]
// (src/ORBextractor.cc:1086).  One CTA = 128 x 32 output tile of one level of one frame.
// ------------------------------------------------------------------------------------------------
#define BLUR_TW 128
#define BLUR_TH 35           // a multiple of the 7-row register window
#define BLUR_WARPS 4
// one WARP per 32 strips x 35 rows (lane = 4-px column strip, see plf_blur_strip; interior strips packed densely,
// the edge strips of a row band in a warp of their own, see plf_strip_of); tiles of all levels in one launch
__global__ void __launch_bounds__(32 * BLUR_WARPS, 8)
k_blur7(OrbGeom g, OrbPtrs p)
{
    int tileId = blockIdx.x * BLUR_WARPS + (threadIdx.x >> 5), l = 0;
    if (tileId >= g.totalBlurTiles) return;
    while (l + 1 < g.nlevels && tileId >= g.lv[l + 1].blurTileBase) l++;
    tileId -= g.lv[l].blurTileBase;
    const OrbLevelGeom& L = g.lv[l];
    // tiles of a level: first all interior tiles (row-major), then the edge tiles of all row bands, so that the slow
    // edge warps share CTAs with each other instead of holding a CTA of fast warps resident
    const int ncx_int = L.blurTilesX - 1, nInt = ncx_int * ((L.h + BLUR_TH - 1) / BLUR_TH);
    int tx, ty;
    if (tileId < nInt) { ty = tileId / ncx_int; tx = tileId - ty * ncx_int; }
    else { ty = tileId - nInt; tx = ncx_int; }
    const int ty0 = ty * BLUR_TH;
    const int strip = plf_strip_of(tx, threadIdx.x & 31, L.w, 1, L.blurF, ncx_int);
    if (strip < 0) return;
    const uint8_t* src = p.lvl[l] + (size_t)blockIdx.y * p.frameStride[l];
    uint8_t* dst = p.blr[l] + (size_t)blockIdx.y * L.frameBytes;
    BlurTaps taps;
    taps.k[0] = 18; taps.k[1] = 34; taps.k[2] = 48; taps.k[3] = 56; taps.k[4] = 48; taps.k[5] = 34; taps.k[6] = 18; taps.k[7] = 0;
    plf_blur_strip<3>(src, p.pitch[l], dst, L.pitch, L.w, L.h, 4 * strip, ty0, BLUR_TH, taps);
}

// ------------------------------------------------------------------------------------------------
// Per-cell FAST-9/16 with threshold retry and per-cell NMS (src/ORBextractor.cc:789-829 over cv::FAST).
// One CTA per 30-px cell.  Keys are appended to the (frame, level) list with one atomic per cell; their
// order in the list is irrelevant because the octree breaks response ties with an explicit order key
// (cell row, cell column, y, x) that restates the reference's push_back order.
// ------------------------------------------------------------------------------------------------
#define FAST_MAXC 72          // max cell window edge (wCell + 6); geometry setup checks this
#define FAST_WARPS 4          // cells per CTA (one warp each)

// 16 differences centre - ring pixel (ring order of cv::FAST, k and k + 8 opposite)
__device__ __forceinline__ void fast_ring(const uint8_t* c, int pitch, int v, int (&d)[16])
{
    d[1] = v - c[3 * pitch + 1];  d[2] = v - c[2 * pitch + 2];  d[3] = v - c[pitch + 3];
    d[5] = v - c[-pitch + 3];     d[6] = v - c[-2 * pitch + 2]; d[7] = v - c[-3 * pitch + 1];
    d[9] = v - c[-3 * pitch - 1]; d[10] = v - c[-2 * pitch - 2]; d[11] = v - c[-pitch - 3];
    d[13] = v - c[pitch - 3];     d[14] = v - c[2 * pitch - 2]; d[15] = v - c[3 * pitch - 1];
}

// max over the 16 nine-long arcs of the arc minimum of d (three-input min / max trees)
__device__ __forceinline__ int fast_arc_maxmin(const int (&d)[16])
{
    int m3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m3[k] = min(min(d[k], d[(k + 1) & 15]), d[(k + 2) & 15]);
    int best = -1000;
#pragma unroll
    for (int k = 0; k < 16; k++) best = max(best, min(min(m3[k], m3[(k + 3) & 15]), m3[(k + 6) & 15]));
    return best;
}

// Quick reject on the four compass points: a 9-long arc contains one pixel of every opposite pair (k, k + 8), so both
// pixels of a pair inside [v - th, v + th] rules the corner out.
__device__ __forceinline__ bool fast_quick(const uint8_t* c, int pitch, int th)
{
    const int v = c[0];
    const int d0 = v - c[3 * pitch], d8 = v - c[-3 * pitch], d4 = v - c[3], d12 = v - c[-3];
    return (d0 > th || d8 > th || d0 < -th || d8 < -th) && (d4 > th || d12 > th || d4 < -th || d12 < -th);
}

__device__ __forceinline__ void fast_ring_all(const uint8_t* c, int pitch, int (&d)[16])
{
    const int v = c[0];
    d[0] = v - c[3 * pitch];  d[8] = v - c[-3 * pitch];  d[4] = v - c[3];  d[12] = v - c[-3];
    fast_ring(c, pitch, v, d);
}

// bit 0: a 9-run of ring pixels darker than centre - th exists (d > th), bit 1: a 9-run brighter than centre + th
__device__ __forceinline__ int fast_sides(const uint8_t* c, int pitch, int th)
{
    int d[16];
    fast_ring_all(c, pitch, d);
    unsigned hi = 0, lo = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        hi |= (unsigned)(d[k] > th) << k;
        lo |= (unsigned)(d[k] < -th) << k;
    }
    hi |= hi << 16; lo |= lo << 16;
    unsigned h = hi & (hi >> 1); h &= h >> 2; h &= h >> 4; h &= hi >> 8;  // runs of 9
    unsigned w = lo & (lo >> 1); w &= w >> 2; w &= w >> 4; w &= lo >> 8;
    return (h ? 1 : 0) | (w ? 2 : 0);
}

// max(A, B) = FAST score + 1.  Only a side that has a 9-run can reach a value above the threshold, so only that
// side's arcs are evaluated.
__device__ __forceinline__ int fast_score(const uint8_t* c, int pitch, int sides)
{
    int d[16];
    fast_ring_all(c, pitch, d);
    int best = 0;
    if (sides & 1) best = fast_arc_maxmin(d);
    if (sides & 2) {
#pragma unroll
        for (int k = 0; k < 16; k++) d[k] = -d[k];
        best = max(best, fast_arc_maxmin(d));
    }
    return best;
}

// One WARP per 30-px cell, no block barriers.  Pass A evaluates the cell at iniTh; only a cell that yields no
// keypoint there is evaluated again at minTh (src/ORBextractor.cc:809-816).  Inside a pass the work is compacted
// between stages so that the expensive stages run with full warps: (1) compass quick-reject over the interior ->
// list of survivors, (2) 16-pixel ring + 9-run test on the list -> list of corners, (3) arc scores of the corners
// into the score map, (4) per-cell NMS of the corners (strict maximum over the 8 neighbours).  Keys go straight to
// the (frame, level) list with warp-aggregated atomics; their order is irrelevant (the octree breaks response ties
// with an explicit order key that restates the reference's push_back order).
// Dynamic shared memory per warp: pixel map and score map of `rows` x `tp` bytes (tp = window width + 3 alignment
// bytes rounded up to a multiple of 4) and a 16-bit work list of `lcap` entries, sized from the largest cell.
// geometry of one FAST cell of the fused (all levels) cell grid; ok = false for the cells the reference skips
struct FastCell { int l, ci, cj, iniX, iniY, cw, ch; bool ok; };
__device__ __forceinline__ FastCell fast_cell_geom(const OrbGeom& g, int cell)
{
    FastCell c;
    int l = 0;
    while (l + 1 < g.nlevels && cell >= g.lv[l + 1].cellBase) l++;
    cell -= g.lv[l].cellBase;
    const OrbLevelGeom& L = g.lv[l];
    c.l = l;
    c.ci = cell / L.nCols; c.cj = cell - c.ci * L.nCols;
    const int maxBorderX = L.w - ORB_MINB, maxBorderY = L.h - ORB_MINB;
    c.iniY = ORB_MINB + c.ci * L.hCell; c.iniX = ORB_MINB + c.cj * L.wCell;
    c.ok = !(c.iniY >= maxBorderY - 3 || c.iniX >= maxBorderX - 6);   // skip rules, :794, :803
    const int maxY = min(c.iniY + L.hCell + 6, maxBorderY), maxX = min(c.iniX + L.wCell + 6, maxBorderX);
    c.cw = maxX - c.iniX; c.ch = maxY - c.iniY;
    if (c.cw < 7 || c.ch < 7) c.ok = false;
    return c;
}

// the two threshold passes over one cell window that sits in shared memory: `tile` = pixel map with row pitch tp (window column
// rx at map column rx + off), `best` = score map of the same shape (all zero on entry), `list` = 16-bit work list
__device__ __forceinline__ void fast_cell_passes(const OrbGeom& g, const OrbPtrs& p, const OrbLevelGeom& L, const int l, const int ci, const int cj,
                                                 const int cw, const int ch, const int off, const int tp, uint8_t* tile, uint8_t* best,
                                                 unsigned short* list, const int lane)
{
    const unsigned FULL = 0xffffffffu, LT = (1u << lane) - 1u;
    int* cnt = p.rawcount + (size_t)blockIdx.y * g.nlevels + l;
    unsigned* out = p.rawkeys + (size_t)blockIdx.y * g.rawPerFrame + L.rawOff;
    for (int pass = 0; pass < 2; pass++) {
        const int th = pass == 0 ? g.iniTh : g.minTh;
        // (1) compass quick-reject over the interior, four pixels per lane from 32-bit shared-memory words: a lane owns
        // one aligned word of a row, a warp covers 32 / (words per row) rows per step
        int n1 = 0;
        {
            const int tpw = tp >> 2;
            const int c0 = 3 + off, c1 = cw - 3 + off;          // interior map columns [c0, c1)
            const int w0 = c0 >> 2, nwl = ((c1 + 3) >> 2) - w0;  // words per row that hold interior pixels (<= 19)
            const int rpi = 32 / nwl;                            // rows per step
            const int lr = lane / nwl, wq = w0 + lane - lr * nwl;
            const unsigned* T = (const unsigned*)tile;
            const unsigned th4 = (unsigned)th * 0x01010101u;
            unsigned vmask = 0;       // bytes of this lane's word that are interior pixels
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (4 * wq + b >= c0 && 4 * wq + b < c1) vmask |= 0xffu << (8 * b);
            for (int rb = 3; rb < ch - 3; rb += rpi) {
                const int ry = rb + lr;
                unsigned go = 0;          // byte mask: 0xff where the pixel survives the compass test
                if (lr < rpi && ry < ch - 3) {
                    const unsigned C = T[ry * tpw + wq], Lw = T[ry * tpw + wq - 1], Rw = T[ry * tpw + wq + 1];
                    const unsigned U = T[(ry - 3) * tpw + wq], D = T[(ry + 3) * tpw + wq];
                    // four pixels at once with byte SIMD: |centre - ring pixel| > th for the four compass points
                    const unsigned Lf = __funnelshift_r(Lw, C, 8);      // the pixels three columns to the left of C's four
                    const unsigned Rt = __funnelshift_r(C, Rw, 24);     // ... three columns to the right
                    const unsigned gv = __vcmpgtu4(__vabsdiffu4(C, U), th4) | __vcmpgtu4(__vabsdiffu4(C, D), th4);
                    const unsigned gh = __vcmpgtu4(__vabsdiffu4(C, Lf), th4) | __vcmpgtu4(__vabsdiffu4(C, Rt), th4);
                    go = gv & gh & vmask;
                }
                if (!__any_sync(FULL, go != 0u)) continue;      // no survivor in these rows (flat areas): skip the compaction votes
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const bool g1 = (go >> (8 * b)) & 1u;
                    const unsigned m = __ballot_sync(FULL, g1);
                    if (g1) list[n1 + __popc(m & LT)] = (unsigned short)(ry * tp + 4 * wq + b);
                    n1 += __popc(m);
                }
            }
        }
        __syncwarp();
        // (2) ring + 9-run test, compacted in place (entry: map offset | sides << 14)
        int n2 = 0;
        for (int i0 = 0; i0 < n1; i0 += 32) {
            const int i = i0 + lane;
            int o = 0, sides = 0;
            if (i < n1) { o = list[i]; sides = fast_sides(&tile[o], tp, th); }
            __syncwarp();                                   // every lane has read its entry before slots are reused
            const unsigned m = __ballot_sync(FULL, sides != 0);
            if (sides) list[n2 + __popc(m & LT)] = (unsigned short)(o | (sides << 14));
            n2 += __popc(m);
        }
        __syncwarp();
        // (3) scores of the corners (independent of the threshold; <= 255)
        for (int i = lane; i < n2; i += 32) {
            const int e = list[i], o = e & 0x3fff;
            best[o] = (uint8_t)fast_score(&tile[o], tp, e >> 14);
        }
        __syncwarp();
        // (4) per-cell NMS: scores at or below the threshold (and the ring around the interior, which stays 0) count
        // as 0; score = best - 1 is monotone, so best values compare the same
        int found = 0;
        for (int i0 = 0; i0 < n2; i0 += 32) {
            const int i = i0 + lane;
            bool key = false;
            int sc = 0, o = 0;
            if (i < n2) {
                o = list[i] & 0x3fff;
                const uint8_t* b = &best[o];
                sc = b[0];
#define NB(d) ((int)b[d] > th ? (int)b[d] : 0)
                key = sc > th && sc > NB(-1) && sc > NB(1) && sc > NB(-tp - 1) && sc > NB(-tp) && sc > NB(-tp + 1) &&
                      sc > NB(tp - 1) && sc > NB(tp) && sc > NB(tp + 1);
#undef NB
            }
            const unsigned m = __ballot_sync(FULL, key);
            if (m) {
                int base = 0;
                const int leader = __ffs((int)m) - 1;
                if (lane == leader) base = atomicAdd(cnt, __popc(m));
                base = __shfl_sync(FULL, base, leader);
                if (key) {
                    const int oo = base + __popc(m & LT);
                    const int ry = o / tp, rx = o - ry * tp - off;
                    const int kx = rx + cj * L.wCell, ky = ry + ci * L.hCell;   // relative to minBorder, :822-823
                    if (oo < L.rawcap) out[oo] = (unsigned)kx | ((unsigned)ky << 12) | ((unsigned)(sc - 1) << 24);
                }
                found += __popc(m);
            }
        }
        if (found > 0) break;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32 * FAST_WARPS)
k_fast_cells(OrbGeom g, OrbPtrs p, int tp, int rows, int lcap)
{
    PLF_DYN_SMEM(smem);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell = blockIdx.x * FAST_WARPS + wid;
    if (cell >= g.totalCells) return;
    const FastCell C = fast_cell_geom(g, cell);
    if (!C.ok) return;
    const int l = C.l, ci = C.ci, cj = C.cj, iniX = C.iniX, iniY = C.iniY, cw = C.cw, ch = C.ch;
    const OrbLevelGeom& L = g.lv[l];
    const size_t per_warp = (size_t)2 * tp * rows + (size_t)2 * lcap;
    uint8_t* tile = smem + (size_t)wid * per_warp;
    uint8_t* best = tile + (size_t)tp * rows;
    unsigned short* list = (unsigned short*)(best + (size_t)tp * rows);
    const int spitch = p.pitch[l];
    const uint8_t* src = p.lvl[l] + (size_t)blockIdx.y * p.frameStride[l] + (size_t)iniY * spitch + iniX;
    // window -> shared memory; aligned buffers are read as 32-bit words (16 words x 2 rows per step)
    const int off = (int)((size_t)src & 3);            // window column rx lives at map column rx + off
    if (((size_t)spitch & 3) == 0) {
        const unsigned* s4 = (const unsigned*)(src - off);
        const int nW = (off + cw + 3) >> 2, wx = lane & 15;
        for (int ry = lane >> 4; ry < ch; ry += 2)
            if (wx < nW) {
                ((unsigned*)tile)[ry * (tp >> 2) + wx] = s4[(size_t)ry * (spitch >> 2) + wx];
                ((unsigned*)best)[ry * (tp >> 2) + wx] = 0u;
            }
    } else {
        for (int ry = 0; ry < ch; ry++)
            for (int rx = lane; rx < cw; rx += 32) {
                tile[ry * tp + rx + off] = src[(size_t)ry * spitch + rx];
                best[ry * tp + rx + off] = 0;
            }
    }
    __syncwarp();
    fast_cell_passes(g, p, L, l, ci, cj, cw, ch, off, tp, tile, best, list, lane);
}

// ------------------------------------------------------------------------------------------------
// DistributeOctTree (src/ORBextractor.cc:539-763), one CTA per (level, frame).
//
// The reference walks a std::list; this kernel keeps the list as an array and rebuilds it once per pass.
// A pass (full or "refinement") is: pick the nodes to divide and their processing order, count the keys
// of every child, find where the pass stops (refinement breaks as soon as size >= N), then emit
// [children of the last processed parent n4..n1, ..., children of the first processed parent n4..n1,
//  old list without the divided parents] -- exactly the push_front / erase result.
// Equal-size ties in the refinement sort are broken by creation order, later-created first (the oracle's
// documented stand-in for the reference's heap-address order).
// ------------------------------------------------------------------------------------------------
#define OCT_T 256
struct OctSmem {
    int cap;
    short* nb[2];      // bounds: 4 shorts per node (ulx, uly, brx, bry)
    int* ncount[2];
    int* ncid[2];
    int* ord;          // list pos -> processing order o (or -1)
    int* proc;         // o -> list pos
    int* cc;           // 4 per candidate: child key counts
    int* cpos;         // 4 per candidate: child new position (or -1)
    int* spos;         // list pos -> survivor new position
    int* tmp;          // scan buffer, 4*cap
    unsigned long long* best;
    int* partial;      // OCT_T + 1
};

__device__ __forceinline__ void oct_carve(unsigned char* base, int cap, OctSmem& s)
{
    size_t o = 0;
    s.cap = cap;
    s.best = (unsigned long long*)(base + o); o += sizeof(unsigned long long) * cap;
    for (int b = 0; b < 2; b++) { s.ncount[b] = (int*)(base + o); o += sizeof(int) * cap; }
    for (int b = 0; b < 2; b++) { s.ncid[b] = (int*)(base + o); o += sizeof(int) * cap; }
    s.ord = (int*)(base + o); o += sizeof(int) * cap;
    s.proc = (int*)(base + o); o += sizeof(int) * cap;
    s.spos = (int*)(base + o); o += sizeof(int) * cap;
    s.cc = (int*)(base + o); o += sizeof(int) * 4 * cap;
    s.cpos = (int*)(base + o); o += sizeof(int) * 4 * cap;
    s.tmp = (int*)(base + o); o += sizeof(int) * 4 * cap;
    s.partial = (int*)(base + o); o += sizeof(int) * (OCT_T + 1);
    for (int b = 0; b < 2; b++) { s.nb[b] = (short*)(base + o); o += sizeof(short) * 4 * cap; }
}
static inline size_t oct_smem_bytes(int cap)
{
    return sizeof(unsigned long long) * cap + sizeof(int) * cap * (2 + 2 + 3 + 12) + sizeof(int) * (OCT_T + 1) +
           sizeof(short) * 8 * cap + 16;
}

// exclusive scan of data[0..n) in place; returns the total. All threads must call.
__device__ __forceinline__ int oct_exscan(int* data, int n, int* partial)
{
    const int tid = threadIdx.x;
    const int per = (n + OCT_T - 1) / OCT_T;
    const int b = tid * per, e = min(b + per, n);
    int sum = 0;
    for (int i = b; i < e; i++) sum += data[i];
    partial[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < OCT_T; i++) { int v = partial[i]; partial[i] = acc; acc += v; }
        partial[OCT_T] = acc;
    }
    __syncthreads();
    int acc = partial[tid];
    for (int i = b; i < e; i++) { int v = data[i]; data[i] = acc; acc += v; }
    int total = partial[OCT_T];
    __syncthreads();
    return total;
}

__device__ __forceinline__ int oct_quadrant(unsigned key, const short* nb)
{
    const int x = key & 0xfff, y = (key >> 12) & 0xfff;
    const int mx = nb[0] + ((nb[2] - nb[0] + 1) >> 1);   // UL.x + ceil((UR.x-UL.x)/2)
    const int my = nb[1] + ((nb[3] - nb[1] + 1) >> 1);
    return (x < mx ? 0 : 1) + (y < my ? 0 : 2);
}

__device__ void oct_distribute(const unsigned* __restrict__ keys, int K, unsigned short* __restrict__ knode,
                               int W, int H, int N, int wCell, int hCell, int* __restrict__ out, int outcap,
                               int* __restrict__ outcount, unsigned char* smem, int cap)
{
    __shared__ int s_n, s_mode, s_finish, s_ostar, s_cidbase;
    OctSmem S;
    oct_carve(smem, cap, S);
    const int tid = threadIdx.x;
    if (K <= 0) {
        if (tid == 0) *outcount = 0;
        return;
    }
    int nIni = (int)roundf((float)W / (float)H);
    if (nIni < 1) nIni = 1;
    if (nIni > cap) nIni = cap;
    const float hX = (float)W / (float)nIni;
    int cur = 0;
    for (int i = tid; i < nIni; i += OCT_T) {
        S.nb[0][4 * i + 0] = (short)(int)(hX * (float)i);
        S.nb[0][4 * i + 1] = 0;
        S.nb[0][4 * i + 2] = (short)(int)(hX * (float)(i + 1));
        S.nb[0][4 * i + 3] = (short)H;
        S.ncount[0][i] = 0;
    }
    __syncthreads();
    for (int k = tid; k < K; k += OCT_T) {
        int i = (int)((float)(keys[k] & 0xfff) / hX);
        if (i >= nIni) i = nIni - 1;
        atomicAdd(&S.ncount[0][i], 1);
        knode[k] = (unsigned short)i;
    }
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int i = 0; i < nIni; i++) {
            if (S.ncount[0][i] > 0) {
                S.spos[i] = m;
                for (int c = 0; c < 4; c++) S.nb[0][4 * m + c] = S.nb[0][4 * i + c];
                S.ncount[0][m] = S.ncount[0][i];
                S.ncid[0][m] = i;
                m++;
            } else S.spos[i] = -1;
        }
        s_n = m; s_mode = 0; s_finish = 0; s_cidbase = nIni;
    }
    __syncthreads();
    for (int k = tid; k < K; k += OCT_T) knode[k] = (unsigned short)S.spos[knode[k]];
    __syncthreads();

    while (true) {
        const int n = s_n, mode = s_mode;
        const int* cnt = S.ncount[cur];
        const int* cid = S.ncid[cur];
        const short* nb = S.nb[cur];
        // 1. candidates and processing order
        for (int p = tid; p < n; p += OCT_T) S.tmp[p] = cnt[p] > 1 ? 1 : 0;
        __syncthreads();
        int E;
        if (mode == 0) {
            E = oct_exscan(S.tmp, n, S.partial);
            for (int p = tid; p < n; p += OCT_T) S.ord[p] = cnt[p] > 1 ? S.tmp[p] : -1;
        } else {
            for (int p = tid; p < n; p += OCT_T) {
                int o = -1;
                if (cnt[p] > 1) {
                    o = 0;
                    const int c0 = cnt[p], i0 = cid[p];
                    for (int r = 0; r < n; r++) {
                        int c1 = cnt[r];
                        if (c1 > 1 && (c1 > c0 || (c1 == c0 && cid[r] > i0))) o++;
                    }
                }
                S.ord[p] = o;
            }
            __syncthreads();
            E = oct_exscan(S.tmp, n, S.partial);
        }
        __syncthreads();
        if (E == 0) break;   // nothing to divide: size == prevSize -> finish (uniform)
        for (int p = tid; p < n; p += OCT_T) if (S.ord[p] >= 0) S.proc[S.ord[p]] = p;
        for (int j = tid; j < 4 * E; j += OCT_T) S.cc[j] = 0;
        __syncthreads();
        // 2. child key counts
        for (int k = tid; k < K; k += OCT_T) {
            int p = knode[k], o = S.ord[p];
            if (o >= 0) atomicAdd(&S.cc[4 * o + oct_quadrant(keys[k], &nb[4 * p])], 1);
        }
        __syncthreads();
        // 3. where the pass stops
        if (mode == 0) {
            if (tid == 0) s_ostar = E - 1;
        } else {
            if (tid == 0) {
                int sz = n, os = E - 1;
                for (int o = 0; o < E; o++) {
                    int nc = (S.cc[4 * o] > 0) + (S.cc[4 * o + 1] > 0) + (S.cc[4 * o + 2] > 0) + (S.cc[4 * o + 3] > 0);
                    sz += nc - 1;
                    if (sz >= N) { os = o; break; }
                }
                s_ostar = os;
            }
        }
        __syncthreads();
        const int ostar = s_ostar, J = 4 * (ostar + 1);
        // 4. child positions (o descending, q descending)
        for (int j = tid; j < J; j += OCT_T) S.tmp[j] = S.cc[j] > 0 ? 1 : 0;
        __syncthreads();
        const int C = oct_exscan(S.tmp, J, S.partial);
        for (int j = tid; j < J; j += OCT_T) S.cpos[j] = S.cc[j] > 0 ? C - 1 - S.tmp[j] : -1;
        __syncthreads();
        // 5. survivors keep their relative order behind the new children
        for (int p = tid; p < n; p += OCT_T) S.tmp[p] = (S.ord[p] >= 0 && S.ord[p] <= ostar) ? 0 : 1;
        __syncthreads();
        const int NS = oct_exscan(S.tmp, n, S.partial);
        for (int p = tid; p < n; p += OCT_T) S.spos[p] = (S.ord[p] >= 0 && S.ord[p] <= ostar) ? -1 : C + S.tmp[p];
        __syncthreads();
        const int nn = C + NS;
        if (nn > cap) {   // cannot happen for a consistent geometry; fail loudly through the count
            if (tid == 0) *outcount = -1;
            return;
        }
        // 6. build the new list
        const int nxt = cur ^ 1;
        const int cidbase = s_cidbase;
        for (int j = tid; j < J; j += OCT_T) {
            int np = S.cpos[j];
            if (np < 0) continue;
            const int pp = S.proc[j >> 2], q = j & 3;
            const short* b = &nb[4 * pp];
            const short mx = (short)(b[0] + ((b[2] - b[0] + 1) >> 1)), my = (short)(b[1] + ((b[3] - b[1] + 1) >> 1));
            short* d = &S.nb[nxt][4 * np];
            d[0] = (q & 1) ? mx : b[0];
            d[1] = (q & 2) ? my : b[1];
            d[2] = (q & 1) ? b[2] : mx;
            d[3] = (q & 2) ? b[3] : my;
            S.ncount[nxt][np] = S.cc[j];
            S.ncid[nxt][np] = cidbase + j;
        }
        for (int p = tid; p < n; p += OCT_T) {
            int np = S.spos[p];
            if (np < 0) continue;
            for (int c = 0; c < 4; c++) S.nb[nxt][4 * np + c] = nb[4 * p + c];
            S.ncount[nxt][np] = cnt[p];
            S.ncid[nxt][np] = cid[p];
        }
        // 7. move the keys
        for (int k = tid; k < K; k += OCT_T) {
            int p = knode[k], o = S.ord[p];
            int np = (o >= 0 && o <= ostar) ? S.cpos[4 * o + oct_quadrant(keys[k], &nb[4 * p])] : S.spos[p];
            knode[k] = (unsigned short)np;
        }
        __syncthreads();
        // 8. loop control (:669-673, :734-735)
        if (tid == 0) {
            int nexp = 0;
            for (int i = 0; i < nn; i++) nexp += S.ncount[nxt][i] > 1;
            s_cidbase = cidbase + J;
            s_n = nn;
            if (nn >= N || nn == n) s_finish = 1;
            else if (mode == 0 && nn + nexp * 3 > N) s_mode = 1;
        }
        cur = nxt;
        __syncthreads();
        if (s_finish) break;
    }
    // best key per node: maximal response, first in the reference's push_back order
    const int n = s_n;
    for (int p = tid; p < n; p += OCT_T) S.best[p] = 0ull;
    __syncthreads();
    for (int k = tid; k < K; k += OCT_T) {
        unsigned key = keys[k];
        int x = key & 0xfff, y = (key >> 12) & 0xfff, r = key >> 24;
        unsigned long long prio = ((unsigned long long)r << 44) | ((unsigned long long)(1023 - (y - 3) / hCell) << 34) |
                                  ((unsigned long long)(1023 - (x - 3) / wCell) << 24) |
                                  ((unsigned long long)(4095 - y) << 12) | (unsigned long long)(4095 - x);
        atomicMax(&S.best[knode[k]], prio);
    }
    __syncthreads();
    for (int k = tid; k < K; k += OCT_T) {
        unsigned key = keys[k];
        int x = key & 0xfff, y = (key >> 12) & 0xfff, r = key >> 24;
        unsigned long long prio = ((unsigned long long)r << 44) | ((unsigned long long)(1023 - (y - 3) / hCell) << 34) |
                                  ((unsigned long long)(1023 - (x - 3) / wCell) << 24) |
                                  ((unsigned long long)(4095 - y) << 12) | (unsigned long long)(4095 - x);
        int p = knode[k];
        if (S.best[p] == prio && p < outcap) out[p] = k;
    }
    if (tid == 0) *outcount = n <= outcap ? n : -1;
}

__global__ void __launch_bounds__(OCT_T)
k_octree(OrbGeom g, OrbPtrs p, int maxcap)
{
    PLF_DYN_SMEM(smem);
    const int l = blockIdx.x, f = blockIdx.y;
    const OrbLevelGeom& L = g.lv[l];
    int K = p.rawcount[(size_t)f * g.nlevels + l];
    int* outcount = p.keptcount + (size_t)f * g.nlevels + l;
    if (K > L.rawcap) {   // raw key list overflowed: report instead of truncating silently
        if (threadIdx.x == 0) { *outcount = -1; }
        return;
    }
    oct_distribute(p.rawkeys + (size_t)f * g.rawPerFrame + L.rawOff, K, p.knode + (size_t)f * g.rawPerFrame + L.rawOff,
                   L.w - 2 * ORB_MINB, L.h - 2 * ORB_MINB, L.nfeat, L.wCell, L.hCell,
                   p.kept + (size_t)f * g.keptPerFrame + L.keptOff, L.keptcap, outcount, smem, L.nodecap < maxcap ? L.nodecap : maxcap);
}

// standalone entry (plf_orb_distribute_octree): keys given explicitly
__global__ void __launch_bounds__(OCT_T)
k_octree_single(const unsigned* keys, int K, unsigned short* knode, int W, int H, int N, int wCell, int hCell,
                int* out, int outcap, int* outcount, int cap)
{
    PLF_DYN_SMEM(smem);
    oct_distribute(keys, K, knode, W, H, N, wCell, hCell, out, outcap, outcount, smem, cap);
}

// ------------------------------------------------------------------------------------------------
// IC_Angle (:77-104) + computeOrbDescriptor (:107-147) + keypoint assembly (:837-847, :1095-1101).
// One warp per retained keypoint; output position = level offset + list position (levels concatenated).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_describe(OrbGeom g, OrbPtrs p, plf_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, int cap,
           int* __restrict__ n_out)
{
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * 8 + (threadIdx.x >> 5);
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
    if (bad || l < 0 || wid >= cap) return;
    const OrbLevelGeom& L = g.lv[l];
    const int kidx = p.kept[(size_t)f * g.keptPerFrame + L.keptOff + pos];
    const unsigned key = p.rawkeys[(size_t)f * g.rawPerFrame + L.rawOff + kidx];
    const int X = (int)(key & 0xfff) + ORB_MINB, Y = (int)((key >> 12) & 0xfff) + ORB_MINB;
    const int resp = key >> 24;
    // orientation on the un-blurred level
    const uint8_t* img = p.lvl[l] + (size_t)f * p.frameStride[l];
    const int pitch = p.pitch[l];
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - ORB_HALF_PATCH;
        const int d = g.umax[v < 0 ? -v : v];
        const uint8_t* row = img + (size_t)(Y + v) * pitch + X;
        int s = 0;
        for (int u = -d; u <= d; u++) {
            int val = row[u];
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
    // steered BRIEF on the blurred level; lane i produces descriptor byte i
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = angle * factorPI;
    // cos(float) under `using namespace std` is cosf (src/ORBextractor.cc:114): glibc's cosf / sinf, bit for bit (plf_libm.cuh)
    const float a = plf_libm::cosf_glibc(ang), b = plf_libm::sinf_glibc(ang);
    const uint8_t* center = p.blr[l] + (size_t)f * L.frameBytes + (size_t)Y * L.pitch + X;
    // the lane's 16 point pairs (32 signed bytes) as two 16-byte loads
    signed char pat[32];
    {
        const int4* pp = (const int4*)(d_orb_pattern + lane * 32);
        const int4 q0 = pp[0], q1 = pp[1];
        const int wds[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int i = 0; i < 32; i++) pat[i] = (signed char)((wds[i >> 2] >> (8 * (i & 3))) & 0xff);
    }
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float x0 = (float)pat[4 * k], y0 = (float)pat[4 * k + 1], x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
        float r0 = x0 * b, r1 = y0 * a, r2 = x0 * a, r3 = y0 * b;
        int t0 = center[__float2int_rn(r0 + r1) * L.pitch + __float2int_rn(r2 - r3)];
        r0 = x1 * b; r1 = y1 * a; r2 = x1 * a; r3 = y1 * b;
        int t1 = center[__float2int_rn(r0 + r1) * L.pitch + __float2int_rn(r2 - r3)];
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
