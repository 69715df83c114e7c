// plf_sort.cuh -- the two data-parallel primitives of the LSD pre-phase, written for this path (no library kernels):
//
//  * plf_scan_popc: offs[i + 1] = number of set bits in mask words 0 .. i (inclusive scan of the popcounts; offs[0] = 0 is the
//    caller's).  Three kernels: per-tile sums, one CTA scans the tile sums, per-tile rescan + base.  A tile = SC_TILE words.
//
//  * plf_sort_frame_keys: stable LSD radix sort of the 64-bit seed keys, 8-bit digits, restricted to the bits that are not yet
//    in order.  k_lsd_keys emits the keys frame by frame in raster order, so (a) the frame field never has to be sorted -- every
//    frame is its own segment [fo[f], fo[f + 1]) and keys only move inside it -- and (b) the raster field is already the
//    tie-break of a stable sort: only the (root, bin) bits above it are sorted, ceil((KB + BB) / 8) passes (4 at 1080p instead
//    of the 5 a whole-array sort of (frame, root, bin) needs).  A pass is three kernels over tiles of RS_TILE keys, a tile never
//    straddling two frames:  k_rs_hist (digit histogram of every tile) -> k_rs_offsets (one CTA per frame: exclusive scan in
//    (digit, tile) order = the stable destination of every tile's digit groups) -> k_rs_scatter (every warp ranks its 512
//    consecutive keys 32 at a time: peers of equal digit by __match_any_sync, rank = peers below the lane, per-warp running
//    digit counters in shared memory that start at the tile's offset plus the counts of the warps before it).
//    Keys with runs of equal digits (pixels of one component follow each other in raster order) scatter to consecutive
//    addresses, so the writes of a warp mostly coalesce without a shared-memory reorder stage.
#pragma once

#define SC_T 256
#define SC_PER 16
#define SC_TILE (SC_T * SC_PER)      // mask words per scan tile

// block-wide exclusive scan of one int per thread (SC_T threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int plf_block_exscan(int v, int* s_warp, int* total)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    int base = 0, tot = 0;
    for (int i = 0; i < nw; i++) {
        const int c = s_warp[i];
        if (i < wid) base += c;
        tot += c;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

static __global__ void __launch_bounds__(SC_T)
k_scan_tile_sums(const unsigned* __restrict__ mask, int n, int* __restrict__ tsum)
{
    __shared__ int s_warp[SC_T / 32];
    const int base = blockIdx.x * SC_TILE;
    int v = 0;
#pragma unroll 4
    for (int j = 0; j < SC_PER; j++) {
        const int i = base + j * SC_T + threadIdx.x;       // coalesced: the order inside a tile does not matter for its sum
        if (i < n) v += __popc(mask[i]);
    }
    int tot;
    plf_block_exscan(v, s_warp, &tot);
    if (threadIdx.x == 0) tsum[blockIdx.x] = tot;
}

// one CTA: exclusive scan of the tile sums in place
static __global__ void __launch_bounds__(SC_T)
k_scan_top(int* __restrict__ tsum, int ntiles)
{
    __shared__ int s_warp[SC_T / 32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b = 0; b < ntiles; b += SC_T) {
        const int i = b + threadIdx.x;
        const int v = i < ntiles ? tsum[i] : 0;
        int tot;
        const int ex = plf_block_exscan(v, s_warp, &tot);
        const int carry = s_carry;
        if (i < ntiles) tsum[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
}

// every thread owns SC_PER consecutive words of the tile
static __global__ void __launch_bounds__(SC_T)
k_scan_apply(const unsigned* __restrict__ mask, int n, const int* __restrict__ tbase, int* __restrict__ out)
{
    __shared__ int s_warp[SC_T / 32];
    const int i0 = blockIdx.x * SC_TILE + threadIdx.x * SC_PER;
    int c[SC_PER];
    int v = 0;
    if (i0 + SC_PER <= n) {
        const uint4* m4 = (const uint4*)(mask + i0);         // i0 is a multiple of 16 words and the mask array is 256-byte aligned
#pragma unroll
        for (int j = 0; j < SC_PER / 4; j++) {
            const uint4 q = m4[j];
            c[4 * j] = __popc(q.x); c[4 * j + 1] = __popc(q.y); c[4 * j + 2] = __popc(q.z); c[4 * j + 3] = __popc(q.w);
        }
    } else {
#pragma unroll
        for (int j = 0; j < SC_PER; j++) c[j] = i0 + j < n ? __popc(mask[i0 + j]) : 0;
    }
#pragma unroll
    for (int j = 0; j < SC_PER; j++) v += c[j];
    int tot;
    int run = tbase[blockIdx.x] + plf_block_exscan(v, s_warp, &tot);
#pragma unroll
    for (int j = 0; j < SC_PER; j++) {
        run += c[j];
        if (i0 + j < n) out[i0 + j] = run;
    }
}

// ------------------------------------------------------------------------------------------------ radix sort
#define RS_T 256
#define RS_PER 16
#define RS_TILE (RS_T * RS_PER)
#define RS_WARPS (RS_T / 32)

// per frame: number of tiles and their first index; tile_frame[t] = frame of tile t; tbase[nframes] = tiles in use
// fo(f) = offs[f * wpf] (offs has the leading 0): first key of frame f.  One CTA.
static __global__ void __launch_bounds__(SC_T)
k_rs_frames(const int* __restrict__ offs, int wpf, int nframes, int* __restrict__ tbase, int* __restrict__ tile_frame, int tile_cap)
{
    __shared__ int s_warp[SC_T / 32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b = 0; b < nframes; b += SC_T) {
        const int f = b + threadIdx.x;
        int nt = 0;
        if (f < nframes) nt = (offs[(size_t)(f + 1) * wpf] - offs[(size_t)f * wpf] + RS_TILE - 1) / RS_TILE;
        int tot;
        const int ex = plf_block_exscan(nt, s_warp, &tot);
        const int first = s_carry + ex;
        if (f < nframes) {
            tbase[f] = first;
            for (int t = 0; t < nt; t++)
                if (first + t < tile_cap) tile_frame[first + t] = f;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) tbase[nframes] = s_carry;
}

struct RsTile { int f, begin, end, tile; };      // keys [begin, end) of frame f; tile = global tile index
__device__ __forceinline__ bool rs_tile(const int* __restrict__ offs, int wpf, int nframes, const int* __restrict__ tbase,
                                        const int* __restrict__ tile_frame, RsTile* T)
{
    const int t = blockIdx.x;
    if (t >= tbase[nframes]) return false;
    const int f = tile_frame[t];
    const int fo = offs[(size_t)f * wpf], fe = offs[(size_t)(f + 1) * wpf];
    T->f = f; T->tile = t;
    T->begin = fo + (t - tbase[f]) * RS_TILE;
    T->end = min(T->begin + RS_TILE, fe);
    return true;
}

// digit counts of one warp's keys into cnt[256] (shared, this warp's own): one shared-memory add per group of equal digits
__device__ __forceinline__ void rs_warp_count(const unsigned long long (&key)[RS_PER], const int nvalid_base, const int nkeys, int shift, int* cnt)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < RS_PER; r++) {
        if (nvalid_base + r * 32 >= nkeys) break;                                // warp-uniform: nothing left in a partial tile
        const bool valid = nvalid_base + r * 32 + lane < nkeys;
        const int d = valid ? (int)((key[r] >> shift) & 0xffu) : 256 + lane;     // invalid lanes: groups of their own
        const unsigned peers = __match_any_sync(FULL, d);
        if (valid && (__ffs((int)peers) - 1) == lane) cnt[d] += __popc(peers);   // one lane per digit: no two lanes write the same counter
        __syncwarp();
    }
}

static __global__ void __launch_bounds__(RS_T)
k_rs_hist(const unsigned long long* __restrict__ in, const int* __restrict__ offs, int wpf, int nframes, const int* __restrict__ tbase,
          const int* __restrict__ tile_frame, int shift, int* __restrict__ hist)
{
    __shared__ int cnt[RS_WARPS][256];
    RsTile T;
    if (!rs_tile(offs, wpf, nframes, tbase, tile_frame, &T)) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_T) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int nkeys = T.end - T.begin, wbase = wid * (RS_PER * 32);
    unsigned long long key[RS_PER];
#pragma unroll
    for (int r = 0; r < RS_PER; r++) {
        const int i = wbase + r * 32 + lane;
        key[r] = i < nkeys ? in[T.begin + i] : 0ull;
    }
    rs_warp_count(key, wbase, nkeys, shift, cnt[wid]);
    __syncthreads();
    int s = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) s += cnt[w][threadIdx.x];
    hist[(size_t)T.tile * 256 + threadIdx.x] = s;
}

// one CTA per frame, thread d = digit d: destination of (digit, tile) = frame start + keys of smaller digits + keys of this
// digit in earlier tiles (in place: hist becomes the offset table)
static __global__ void __launch_bounds__(256)
k_rs_offsets(int* __restrict__ hist, const int* __restrict__ offs, int wpf, const int* __restrict__ tbase)
{
    __shared__ int s_warp[8];
    const int f = blockIdx.x, d = threadIdx.x;
    const int t0 = tbase[f], t1 = tbase[f + 1];
    int tot = 0;
    for (int t = t0; t < t1; t++) tot += hist[(size_t)t * 256 + d];
    int all;
    int run = offs[(size_t)f * wpf] + plf_block_exscan(tot, s_warp, &all);
    for (int t = t0; t < t1; t++) {
        const int c = hist[(size_t)t * 256 + d];
        hist[(size_t)t * 256 + d] = run;
        run += c;
    }
}

static __global__ void __launch_bounds__(RS_T)
k_rs_scatter(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, const int* __restrict__ offs, int wpf, int nframes,
             const int* __restrict__ tbase, const int* __restrict__ tile_frame, int shift, const int* __restrict__ toff)
{
    __shared__ int cnt[RS_WARPS][256];
    RsTile T;
    if (!rs_tile(offs, wpf, nframes, tbase, tile_frame, &T)) return;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_T) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int nkeys = T.end - T.begin, wbase = wid * (RS_PER * 32);
    unsigned long long key[RS_PER];
#pragma unroll
    for (int r = 0; r < RS_PER; r++) {
        const int i = wbase + r * 32 + lane;
        key[r] = i < nkeys ? in[T.begin + i] : 0ull;
    }
    rs_warp_count(key, wbase, nkeys, shift, cnt[wid]);
    __syncthreads();
    {   // counters -> first destination of every (warp, digit): tile offset + the warps before
        int run = toff[(size_t)T.tile * 256 + threadIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const int c = cnt[w][threadIdx.x];
            cnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
    int* my = cnt[wid];
#pragma unroll
    for (int r = 0; r < RS_PER; r++) {
        if (wbase + r * 32 >= nkeys) break;          // warp-uniform
        const bool valid = wbase + r * 32 + lane < nkeys;
        const int d = valid ? (int)((key[r] >> shift) & 0xffu) : 256 + lane;
        const unsigned peers = __match_any_sync(FULL, d);
        int pos = 0;
        if (valid) pos = my[d] + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
        if (valid && (__ffs((int)peers) - 1) == lane) my[d] += __popc(peers);
        __syncwarp();
        if (valid) out[pos] = key[r];
    }
}
