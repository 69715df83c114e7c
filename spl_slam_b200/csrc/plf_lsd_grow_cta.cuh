// plf_lsd_grow_cta.cuh -- LSD region growing for GIANT components: in-order commit, speculative growth ahead of it.
//
// At 1080p one 8-connected component of the "gradient defined" mask holds ~95 % of a frame's defined pixels (the whole edge
// network hangs together), so "one warp per component" (k_lsd_grow_warp) degenerates into ONE dependent chain per frame:
// ~130 k pixels x ~370 ns.  The reference's order cannot be given up (regions are coupled through the `used` map, and the
// output depends on it), but most of the chain is not really sequential:
//
//   * Region j's growth reads `used` only to skip pixels.  Let it run EARLY, against whatever is committed at that moment.
//     Every pixel it saw as used is used in the true execution too (commits are monotone and all come from seeds before j).
//     The only way the early run can differ from the true one is by ACCEPTING a pixel that an earlier seed takes later.  So a
//     speculative region T_j is exactly the reference's region iff, when j's turn comes, no pixel of T_j is used -- a check of
//     |T_j| bits, no halo, no snapshots.  (Proof sketch: walk both executions test by test; the first differing decision is an
//     acceptance of a pixel that is used in the true execution, which puts that pixel into T_j.)
//   * Which seeds will start a region?  Mostly the local maxima of the seed order: a seed with an angle-aligned 8-neighbour
//     that comes EARLIER in the order is almost always swallowed by that neighbour's region before its turn ("likely" = no such
//     neighbour; k_lsd_likely).  Speculating on every unused seed would grow the same region dozens of times.
//
// One CTA per giant component: warp 0 is the COMMITTER, it walks the seeds in order exactly like k_lsd_grow_warp; for an unused
// seed it looks for a finished speculative region (validate |T| bits, copy it into the arena, set the used bits) and otherwise
// grows the region itself (seeds nobody speculated on, failed validations).  Warps 1.. are SPECULATORS: each takes the next
// likely unused seed beyond the committer, grows it with a private "in my region" bitmap into one of its two scratch buffers
// and publishes it.  The committer never waits for a speculator that is not actively growing and a speculator never waits in
// the middle of a region, so there is no cycle.  Results are bit-identical to the sequential order by construction; the parity
// tests run this path on every image (the emulated build lowers the giant threshold to 128 pixels to stress it).
#pragma once

#ifndef LSD_GIANT_BUCKET
#define LSD_GIANT_BUCKET 10                 // smallest components (>= 1024 seeds) that may get a CTA of their own; the host raises it for big batches
#endif
#ifndef GC_MAXWARPS
#define GC_MAXWARPS 16                      // committer + up to 15 speculators
#endif
#ifndef GC_BUF
#define GC_BUF 4                            // scratch buffers (outstanding results) per speculator
#endif
#ifndef GC_CHUNK
#define GC_CHUNK 8                          // seed positions a speculator claims at a time: consecutive likely seeds go to DIFFERENT warps,
#endif                                      // so the regions the committer needs next are grown side by side, not one after the other
#define GC_SLOTS ((GC_MAXWARPS - 1) * GC_BUF)
#define GC_RMAX 2048                        // points per speculative region; larger regions are grown by the committer
#define GC_SMEM_BUDGET (200 * 1024)
enum { GC_FREE = 0, GC_BUSY = 1, GC_DONE = 2, GC_FAIL = 3 };

#ifdef PLF_EMU
#include <sched.h>
static inline void gc_pause() { sched_yield(); }
#else
__device__ __forceinline__ void gc_pause() { __nanosleep(64); }
#endif

// likely[i] = 1 iff seed i (position in the sorted key array) has no angle-aligned 8-neighbour that precedes it in the order
__global__ void __launch_bounds__(256)
k_lsd_likely(const unsigned long long* __restrict__ keys, int n, const int* __restrict__ cid, const float* __restrict__ fa, int w, int h,
             double prec, unsigned char* __restrict__ likely, int kb)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    const size_t foff = (size_t)LSD_KEY_FRAME(k) * ((size_t)w * h);
    const int p = LSD_KEY_IDX(k);
    const int y = p / w, x = p - y * w;
    const double a0 = (double)fa[foff + p] * LSD_D2R;
    int lk = 1;
#pragma unroll
    for (int nb = 0; nb < 9; nb++) {
        if (nb == 4) continue;
        const int xx = x + (nb % 3) - 1, yy = y + (nb / 3) - 1;
        if (xx < 0 || yy < 0 || xx >= w || yy >= h) continue;
        const float fv = __ldg(&fa[foff + yy * w + xx]);
        if (fv > -500.f && __ldg(&cid[foff + yy * w + xx]) < i && wg_ntheta(a0, (double)fv * LSD_D2R) <= prec) lk = 0;
    }
    likely[i] = (unsigned char)lk;
}

struct GcLane { int lane, grp, gdx, gdy; };

__device__ __forceinline__ WgCand gc_load(int q, int cnt, const int* ring, const int* pts, int gdx, int gdy, int w, int h,
                                          const int* __restrict__ CID, const float* __restrict__ F, const float2* __restrict__ CS)
{
    WgCand c;
    const int pp = (cnt - q <= WG_RING) ? ring[q & (WG_RING - 1)] : pts[q];
    c.pp = pp;
    c.xx = (pp & 0xffff) + gdx; c.yy = (pp >> 16) + gdy;
    c.inb = c.xx >= 0 && c.yy >= 0 && c.xx < w && c.yy < h;
    c.ci = -1; c.fv = 0.f; c.cv = make_float2(0.f, 0.f);
    if (c.inb) {
        const int qi = c.yy * w + c.xx;
        c.ci = __ldg(&CID[qi]);
        c.fv = __ldg(&F[qi]);
        c.cv = __ldg(&CS[qi]);
        const int x2 = c.xx + gdx, y2 = c.yy + gdy;
        if (x2 >= 0 && y2 >= 0 && x2 < w && y2 < h) {
            const int q2 = qi + gdy * w + gdx;
            PLF_PREFETCH_L1(&CID[q2]);
            PLF_PREFETCH_L1(&F[q2]);
            PLF_PREFETCH_L1(&CS[q2]);
        }
    }
    return c;
}

// region_grow of one seed by one warp (the loop of k_lsd_grow_warp, see the comments there for the sequential / batched
// acceptance rules).  SPEC = false: the committer's own growth, points go straight into the arena (pts = regpts + arena) and
// `used` is marked.  SPEC = true: points and their component indices go to scratch (pts / ccs, capacity cap), "in my region"
// is the private bitmap `mine`, `used` is only read; returns -1 when the region outgrows the buffer or the committer has
// passed the seed (*frontier > mypos).  Returns the number of points; *angle_out = final region angle.
template <bool SPEC>
__device__ __forceinline__ int gc_grow(const int p, const int sc, const int start, const float* __restrict__ F, const float2* __restrict__ CS,
                                       const int* __restrict__ CID, const int w, const int h, const double prec, const bool fast_ok,
                                       const float sphi, unsigned* used, unsigned* mine, int* ring, float2* acc, int* pts, int* ccs,
                                       const int cap, volatile int* frontier, const int mypos, double* angle_out, const GcLane L)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = L.lane, grp = L.grp, gdx = L.gdx, gdy = L.gdy;
    const int sy = p / w, sx = p - sy * w;
    if (lane == 0) {
        pts[0] = sx | (sy << 16);
        ring[0] = sx | (sy << 16);
        if (SPEC) { ccs[0] = sc; mine[sc >> 5] |= 1u << (sc & 31); }
        else used[sc >> 5] |= 1u << (sc & 31);
    }
    int cnt = 1;
    double reg_angle = (double)__ldg(&F[p]) * LSD_D2R;
    float sumdx = (float)cos(reg_angle), sumdy = (float)sin(reg_angle);
    __syncwarp();
    int npf = 0;
    WgCand pf;
    pf.xx = pf.yy = pf.pp = 0; pf.ci = -1; pf.fv = 0.f; pf.cv = make_float2(0.f, 0.f); pf.inb = false;
    bool failed = false;
    for (int r = 0; r < cnt;) {
        if (SPEC) {
            const int fr = *frontier;          // lanes may read it at different moments: decide together
            if (cnt + 32 > cap || __any_sync(FULL, fr > mypos)) { failed = true; break; }
        }
        const int ng = min(WG_E, cnt - r);
        WgCand cd = pf;
        if (grp >= npf) {
            cd.inb = false; cd.ci = -1; cd.pp = -(4 << 16);
            if (grp < ng) cd = gc_load(r + grp, cnt, ring, pts, gdx, gdy, w, h, CID, F, CS);
        }
        const int rn = r + ng;
        npf = min(WG_E, cnt - rn);
        if (grp < npf) pf = gc_load(rn + grp, cnt, ring, pts, gdx, gdy, w, h, CID, F, CS);
        int cc = -1 - lane;
        bool cand = cd.inb && cd.fv > -500.f;
        if (cand) {
            cc = cd.ci - start;
            unsigned u = (used[cc >> 5] >> (cc & 31)) & 1u;
            if (SPEC) u |= (mine[cc >> 5] >> (cc & 31)) & 1u;
            cand = !u;
            if (!cand) cc = -1 - lane;
        }
        const double a = (double)cd.fv * LSD_D2R;
        double n_theta = wg_ntheta(reg_angle, a);
        bool pass = cand && n_theta <= prec;
        unsigned m = __ballot_sync(FULL, pass);
        if (m) {
            bool fast = false;
            if (fast_ok) {
                const float L2 = sumdx * sumdx + sumdy * sumdy;
                if (L2 >= 16.f) {
                    const float D = (float)__popc(m) * sphi * rsqrtf(L2) + 1e-3f;
                    const double Dd = (double)D;
                    const bool risky = cand && fabs(n_theta - prec) <= Dd;
                    fast = D <= 0.09f && !__any_sync(FULL, risky);
                }
            }
            if (fast) {
                bool win = pass;
                if (__popc(m) > 1) {
                    const unsigned same = __match_any_sync(FULL, cc);
                    win = pass && (__ffs((int)same) - 1 == lane);
                }
                const unsigned mw = __ballot_sync(FULL, win);
                const int nw = __popc(mw);
                if (win) {
                    const int rank = __popc(mw & ((1u << lane) - 1u));
                    if (SPEC) { atomicOr(&mine[cc >> 5], 1u << (cc & 31)); ccs[cnt + rank] = cc; }
                    else atomicOr(&used[cc >> 5], 1u << (cc & 31));
                    pts[cnt + rank] = cd.xx | (cd.yy << 16);
                    ring[(cnt + rank) & (WG_RING - 1)] = cd.xx | (cd.yy << 16);
                    acc[rank] = cd.cv;
                }
                __syncwarp();
                for (int j = 0; j < nw; j += 4) {
                    const float2 v0 = acc[j], v1 = acc[(j + 1) & 31], v2 = acc[(j + 2) & 31], v3 = acc[(j + 3) & 31];
                    sumdx += v0.x; sumdy += v0.y;
                    if (j + 1 < nw) { sumdx += v1.x; sumdy += v1.y; }
                    if (j + 2 < nw) { sumdx += v2.x; sumdy += v2.y; }
                    if (j + 3 < nw) { sumdx += v3.x; sumdy += v3.y; }
                }
                cnt += nw;
                reg_angle = (double)plf_fast_atan2(sumdy, sumdx) * LSD_D2R;
            } else {
                for (;;) {
                    const int k0 = __ffs((int)m) - 1;
                    if (lane == k0) {
                        if (SPEC) { mine[cc >> 5] |= 1u << (cc & 31); ccs[cnt] = cc; }
                        else used[cc >> 5] |= 1u << (cc & 31);
                        pts[cnt] = cd.xx | (cd.yy << 16);
                        ring[cnt & (WG_RING - 1)] = cd.xx | (cd.yy << 16);
                    }
                    cnt++;
                    const int cc0 = __shfl_sync(FULL, cc, k0);
                    sumdx += __shfl_sync(FULL, cd.cv.x, k0);
                    sumdy += __shfl_sync(FULL, cd.cv.y, k0);
                    reg_angle = (double)plf_fast_atan2(sumdy, sumdx) * LSD_D2R;
                    cand = cand && lane > k0 && cc != cc0;
                    pass = cand && wg_ntheta(reg_angle, a) <= prec;
                    m = __ballot_sync(FULL, pass);
                    if (!m) break;
                }
            }
        }
        __syncwarp();
        r = rn;
    }
    if (SPEC) {
        // the private bitmap goes back to all-zero: clear exactly the bits this region set
        __syncwarp();
        for (int i = lane; i < cnt; i += 32) {
            const int c = ccs[i];
            atomicAnd(&mine[c >> 5], ~(1u << (c & 31)));
        }
        __syncwarp();
        if (failed) return -1;
    }
    *angle_out = reg_angle;
    return cnt;
}

struct GcShared {
    int frontier;                 // seed position the committer is working on (positions before it are resolved)
    int cursor;                   // next chunk of GC_CHUNK seed positions nobody has scanned for speculation yet
    int st[GC_SLOTS], pos[GC_SLOTS], n[GC_SLOTS];   // slot table: slot = (speculator - 1) * GC_BUF + buffer
    double ang[GC_SLOTS];
    int ring[GC_MAXWARPS][WG_RING];
    float2 acc[GC_MAXWARPS][32];
};

__device__ __forceinline__ void gc_emit_region(LsdRegion* regions, int* nregions, int regcap, unsigned long long key, int r0, int nreg,
                                               double reg_angle, int kb)
{
    const int rf = LSD_KEY_FRAME(key);
    const int rr = atomicAdd(nregions + rf, 1);
    if (rr < regcap) {
        LsdRegion R;
        R.start = r0; R.n = nreg; R.reg_angle = reg_angle; R.seedkey = key;
        regions[(size_t)rf * regcap + rr] = R;
    }
}

__global__ void __launch_bounds__(32 * GC_MAXWARPS)
k_lsd_grow_cta(const unsigned long long* __restrict__ keys, const int2* __restrict__ comp, const int* __restrict__ bcount,
               const unsigned char* __restrict__ likely, const float* __restrict__ fa, const float2* __restrict__ cs,
               const int* __restrict__ cid, int w, int h, double prec, int min_reg_size, int* __restrict__ regpts,
               LsdRegion* __restrict__ regions, int* __restrict__ nregions, int regcap, int kb, int maxc, int2* __restrict__ scratch,
               int giant_bucket, unsigned long long* __restrict__ dbg)
{
    PLF_DYN_SMEM(smem);
    __shared__ GcShared S;
    // dbg (optional): [0] regions taken from speculation [1] their pixels [2] failed validations [3] regions grown by the committer
    // [4] their pixels [5] committer wait polls [6] speculative regions grown [7] their pixels [8] speculations aborted / overflowed
    unsigned long long d_ok = 0, d_okpx = 0, d_bad = 0, d_inl = 0, d_inlpx = 0, d_wait = 0, d_spec = 0, d_specpx = 0, d_fail = 0;
    const unsigned FULL = 0xffffffffu;
    const int wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5, nspec = nwarps - 1;
    GcLane L;
    L.lane = threadIdx.x & 31;
    L.grp = L.lane >> 3;
    {
        const int kk = L.lane & 7, nbr = kk < 4 ? kk : kk + 1;
        L.gdx = (nbr % 3) - 1; L.gdy = (nbr / 3) - 1;
    }
    const int lane = L.lane;
    const int words = maxc >> 5;
    unsigned* used = (unsigned*)smem;
    unsigned* mine = (unsigned*)smem + (size_t)wid * words;        // wid >= 1: private bitmap of this speculator
    int* ring = S.ring[wid];
    float2* acc = S.acc[wid];
    int ngiant = 0;
    for (int k = giant_bucket; k < LSD_NBUCKET; k++) ngiant += bcount[k];
    const size_t px = (size_t)w * h;
    const bool fast_ok = prec < 1.4;
    const float sphi = (float)sin(prec + 0.1) * 1.05f;
    // scratch of this CTA: [speculator][buffer][GC_RMAX] of (xy, component index)
    int2* cta_scr = scratch + (size_t)blockIdx.x * nspec * GC_BUF * GC_RMAX;

    for (int c = blockIdx.x; c < ngiant; c += gridDim.x) {
        const int start = comp[c].x, C = comp[c].y, end = start + C;
        if (C > maxc) continue;          // beyond the bitmap budget: k_lsd_grow takes it (uniform for the CTA)
        const size_t foff = (size_t)LSD_KEY_FRAME(keys[start]) * px;
        const float* F = fa + foff;
        const float2* CS = cs + foff;
        const int* CID = cid + foff;
        __syncthreads();
        for (int i = threadIdx.x; i < nwarps * words; i += blockDim.x) ((unsigned*)smem)[i] = 0u;
        for (int i = threadIdx.x; i < GC_SLOTS; i += blockDim.x) { S.st[i] = GC_FREE; S.pos[i] = -1; S.n[i] = 0; }
        if (threadIdx.x == 0) { S.frontier = -1; S.cursor = 0; }
        __syncthreads();
        volatile int* vfront = &S.frontier;
        volatile int* vst = S.st;
        volatile int* vpos = S.pos;

        if (wid == 0) {
            // ------------------------------------------------ committer: the reference's order, one seed after the other
            int arena = start;
            for (int i0 = start; i0 < end; i0 += 32) {
                const int ii = i0 + lane;
                unsigned long long mykey = 0;
                bool mineb = false;
                if (ii < end) {
                    mykey = keys[ii];
                    const int sc = ii - start;
                    mineb = !((used[sc >> 5] >> (sc & 31)) & 1u);
                }
                unsigned todo = __ballot_sync(FULL, mineb);
                while (todo) {
                    const int s = __ffs((int)todo) - 1;
                    todo &= todo - 1;
                    const int q = i0 + s - start;                       // seed position inside the component
                    if ((used[q >> 5] >> (q & 31)) & 1u) continue;      // taken by a region committed meanwhile
                    const unsigned long long key = __shfl_sync(FULL, mykey, s);
                    const int p = LSD_KEY_IDX(key);
                    if (lane == 0) *vfront = q;
                    __threadfence_block();
                    __syncwarp();
                    // a speculative region for this seed?  (one lane per slot)
                    int slot = -1;
                    for (int s0 = 0; s0 < nspec * GC_BUF && slot < 0; s0 += 32) {
                        const int sl = s0 + lane;
                        const bool hit = sl < nspec * GC_BUF && vst[sl] != GC_FREE && vpos[sl] == q;
                        const unsigned hm = __ballot_sync(FULL, hit);
                        if (hm) slot = s0 + __ffs((int)hm) - 1;
                    }
                    int nreg = -1;
                    double reg_angle = 0;
                    const int r0 = arena;
                    if (slot >= 0) {
                        while (vst[slot] == GC_BUSY) { gc_pause(); d_wait++; }        // its speculator is growing right now and never blocks
                        __syncwarp();
                        __threadfence_block();
                        if (vst[slot] == GC_DONE) {                      // final until this warp frees the slot: uniform
                            const int n = S.n[slot];
                            const int* sxy = (const int*)(cta_scr + (size_t)slot * GC_RMAX);   // [GC_RMAX] xy, then [GC_RMAX] component index
                            const int* scc = sxy + GC_RMAX;
                            bool bad = false;
                            if (n <= 128) {
                                // the usual case: the region's points and component indices are fetched ONCE (both loads in flight
                                // together), validated from registers and written from registers -- one round trip to the scratch
                                // instead of a validation pass followed by a copy pass
                                int rcc[4], rxy[4];
#pragma unroll
                                for (int t = 0; t < 4; t++) {
                                    const int i = lane + 32 * t;
                                    rcc[t] = i < n ? scc[i] : -1;
                                    rxy[t] = i < n ? sxy[i] : 0;
                                }
#pragma unroll
                                for (int t = 0; t < 4; t++)
                                    if (rcc[t] >= 0) bad |= ((used[rcc[t] >> 5] >> (rcc[t] & 31)) & 1u) != 0;
                                bad = __any_sync(FULL, bad);
                                if (!bad) {
#pragma unroll
                                    for (int t = 0; t < 4; t++)
                                        if (rcc[t] >= 0) {
                                            regpts[arena + lane + 32 * t] = rxy[t];
                                            atomicOr(&used[rcc[t] >> 5], 1u << (rcc[t] & 31));
                                        }
                                }
                            } else {
                                for (int i = lane; i < n; i += 32) {
                                    const int cc = scc[i];
                                    bad |= ((used[cc >> 5] >> (cc & 31)) & 1u) != 0;
                                }
                                bad = __any_sync(FULL, bad);
                                if (!bad) {
                                    for (int i = lane; i < n; i += 32) {
                                        const int cc = scc[i];
                                        regpts[arena + i] = sxy[i];
                                        atomicOr(&used[cc >> 5], 1u << (cc & 31));
                                    }
                                }
                            }
                            if (!bad) {
                                nreg = n;
                                reg_angle = S.ang[slot];
                                d_ok++; d_okpx += n;
                            } else d_bad++;
                        }
                        __syncwarp();
                        if (lane == 0) { vpos[slot] = -1; __threadfence_block(); vst[slot] = GC_FREE; }
                    }
                    if (nreg < 0) {   // nobody speculated on it, the speculation failed, or an earlier region took one of its pixels
                        nreg = gc_grow<false>(p, q, start, F, CS, CID, w, h, prec, fast_ok, sphi, used, nullptr, ring, acc, regpts + arena, nullptr,
                                              0, vfront, q, &reg_angle, L);
                        d_inl++; d_inlpx += nreg;
                    }
                    arena += nreg;
                    if (lane == 0 && nreg >= min_reg_size) gc_emit_region(regions, nregions, regcap, key, r0, nreg, reg_angle, kb);
                    __syncwarp();
                }
            }
            if (lane == 0) *vfront = C;       // everything is resolved: speculators stop
            __threadfence_block();
        } else {
            // ------------------------------------------------ speculator
            int chunk = -1;
            unsigned cand_mask = 0;       // warp-uniform: candidates of the current chunk not tried yet
            for (;;) {
                // decisions that read state another warp is changing are taken by lane 0 and broadcast
                int code = 0, b = -1;
                if (lane == 0) {
                    const int fr = *vfront;
                    if (fr < C) {
                        for (int bb = 0; bb < GC_BUF && b < 0; bb++) {
                            const int sl = (wid - 1) * GC_BUF + bb;
                            const int st = vst[sl];
                            // free, or a finished result the committer has already passed (dead)
                            if (st == GC_FREE || ((st == GC_DONE || st == GC_FAIL) && vpos[sl] < fr)) b = bb;
                        }
                        code = b < 0 ? 1 : 2;
                    }
                }
                code = __shfl_sync(FULL, code, 0);
                b = __shfl_sync(FULL, b, 0);
                if (code == 0) break;
                if (code == 1) { gc_pause(); continue; }
                const int sl = (wid - 1) * GC_BUF + b;
                if (!cand_mask) {
                    int cstart = 0;
                    if (lane == 0) cstart = atomicAdd(&S.cursor, GC_CHUNK);
                    cstart = __shfl_sync(FULL, cstart, 0);
                    if (cstart >= C) break;
                    chunk = cstart;
                    const int qq = chunk + lane;
                    const bool ok = lane < GC_CHUNK && qq < C && likely[start + qq] && !((used[qq >> 5] >> (qq & 31)) & 1u);
                    cand_mask = __ballot_sync(FULL, ok);
                    if (!cand_mask) continue;
                }
                const int s = __ffs((int)cand_mask) - 1;
                cand_mask &= cand_mask - 1;
                const int q = chunk + s;
                int go = 0;
                if (lane == 0) {
                    go = q > *vfront && !((used[q >> 5] >> (q & 31)) & 1u);
                    if (go) { vpos[sl] = q; __threadfence_block(); vst[sl] = GC_BUSY; }
                }
                go = __shfl_sync(FULL, go, 0);
                if (!go) continue;
                int* pxy = (int*)(cta_scr + (size_t)sl * GC_RMAX);   // [GC_RMAX] xy, then [GC_RMAX] component index
                int* pcc = pxy + GC_RMAX;
                const int p = LSD_KEY_IDX(keys[start + q]);
                double ang = 0;
                const int n = gc_grow<true>(p, q, start, F, CS, CID, w, h, prec, fast_ok, sphi, used, mine, ring, acc, pxy, pcc, GC_RMAX, vfront, q, &ang, L);
                __syncwarp();
                if (lane == 0) {
                    S.n[sl] = n;
                    S.ang[sl] = ang;
                    __threadfence();
                    vst[sl] = n >= 0 ? GC_DONE : GC_FAIL;
                }
                if (n >= 0) { d_spec++; d_specpx += n; } else d_fail++;
                __syncwarp();
            }
        }
        __syncthreads();
    }
    if (dbg && lane == 0) {
        atomicAdd(&dbg[0], d_ok); atomicAdd(&dbg[1], d_okpx); atomicAdd(&dbg[2], d_bad); atomicAdd(&dbg[3], d_inl); atomicAdd(&dbg[4], d_inlpx);
        atomicAdd(&dbg[5], d_wait); atomicAdd(&dbg[6], d_spec); atomicAdd(&dbg[7], d_specpx); atomicAdd(&dbg[8], d_fail);
    }
}
