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
#include "plf_tma.cuh"

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
    const unsigned bar = plf_smem_u32(&bars[wl]);
    if (lane == 0) {
        plf_mbar_init(bar, 1);
        plf_mbar_expect_tx(bar, DT_RAW_BYTES + DT_BLR_BYTES);
        plf_tma_load_3d(plf_smem_u32(rawT), &tm.raw[l], bar, xr, Y - 15, f);
        plf_tma_load_3d(plf_smem_u32(blrT), &tm.blr[l], bar, xb, Y - 19, f);
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
    plf_mbar_wait(bar, 0);
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

// ------------------------------------------------------------------------------------------------
// FAST with the cell windows fed by TMA.  k_fast_cells spends a third of its stall samples between the global loads of a cell
// window and the shared-memory stores that stage it (one warp per cell: nothing else to do while the loads are in flight), and
// a tenth of its instructions on that staging.  Here a warp walks `cpw` cells; lane 0 issues ONE cp.async.bulk.tensor.3d per
// cell window (box tp x rows at a 16-byte aligned column, zero fill outside the level) into one of two buffers and the window of
// the NEXT cell is already in flight while the warp runs the threshold passes on the current one (fast_cell_passes, shared with
// k_fast_cells, which remains for level-0 buffers whose pitch is not a multiple of 16 bytes).  Consecutive cells go to
// consecutive warps, so neighbouring windows (which overlap by 6 pixels) are fetched at about the same time and meet in L2.
// Dynamic shared memory per warp: two pixel maps + score map (tp x rows each, tp a multiple of 16) + work list.
struct OrbFastMaps { CUtensorMap m[ORB_MAX_LEVELS]; };

__global__ void __launch_bounds__(32 * FAST_WARPS)
k_fast_cells_tma(OrbGeom g, OrbPtrs p, const __grid_constant__ OrbFastMaps fm, int tp, int rows, int lcap, int cpw)
{
    extern __shared__ __align__(128) unsigned char smem_f[];
    __shared__ __align__(8) unsigned long long bars[FAST_WARPS][2];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int map_bytes = tp * rows;                                     // a multiple of 128 (host)
    const size_t per_warp = (size_t)3 * map_bytes + (((size_t)2 * lcap + 127) & ~(size_t)127);
    uint8_t* buf0 = smem_f + (size_t)wid * per_warp;
    uint8_t* best = buf0 + 2 * map_bytes;
    unsigned short* list = (unsigned short*)(best + map_bytes);
    const int nwarps = gridDim.x * FAST_WARPS, gw = blockIdx.x * FAST_WARPS + wid;
    const int f = blockIdx.y;
    const unsigned bar0 = plf_smem_u32(&bars[wid][0]), bar1 = plf_smem_u32(&bars[wid][1]);
    if (lane == 0) { plf_mbar_init(bar0, 1); plf_mbar_init(bar1, 1); }
    __syncwarp();
    // cell k of this warp = gw + k * nwarps; buffer k & 1.  A skipped cell issues no load, so the phase of a buffer's barrier
    // advances only with the loads actually issued into it: ph[b] = parity of the NEXT load's phase (warp-uniform registers)
    FastCell cur, nxt;
    cur.ok = false; nxt.ok = false;
    unsigned ph0 = 0, ph1 = 0, cur_par = 0, nxt_par = 0;
    int c = gw;
    if (c < g.totalCells) {
        cur = fast_cell_geom(g, c);
        if (cur.ok) {
            cur_par = ph0; ph0 ^= 1u;
            if (lane == 0) {
                plf_mbar_expect_tx(bar0, map_bytes);
                plf_tma_load_3d(plf_smem_u32(buf0), &fm.m[cur.l], bar0, cur.iniX & ~15, cur.iniY, f);
            }
        }
    }
    for (int k = 0; k < cpw && c < g.totalCells; k++, c += nwarps) {
        const int cn = c + nwarps;
        nxt.ok = false;
        __syncwarp();                    // every lane is done with the buffer the next load overwrites
        if (k + 1 < cpw && cn < g.totalCells) {
            nxt = fast_cell_geom(g, cn);
            if (nxt.ok) {
                const int nb = (k + 1) & 1;
                if (nb) { nxt_par = ph1; ph1 ^= 1u; } else { nxt_par = ph0; ph0 ^= 1u; }
                if (lane == 0) {
                    const unsigned b = nb ? bar1 : bar0;
                    plf_mbar_expect_tx(b, map_bytes);
                    plf_tma_load_3d(plf_smem_u32(buf0 + nb * map_bytes), &fm.m[nxt.l], b, nxt.iniX & ~15, nxt.iniY, f);
                }
            }
        }
        if (cur.ok) {
            for (int i = lane; i < (map_bytes >> 4); i += 32) ((uint4*)best)[i] = make_uint4(0u, 0u, 0u, 0u);
            plf_mbar_wait((k & 1) ? bar1 : bar0, cur_par);
            __syncwarp();
            fast_cell_passes(g, p, g.lv[cur.l], cur.l, cur.ci, cur.cj, cur.cw, cur.ch, cur.iniX & 15, tp, buf0 + (k & 1) * map_bytes, best, list, lane);
        }
        cur = nxt; cur_par = nxt_par;
    }
}

// host: encode the per-level maps for this call (x, y, frame); false when a base or pitch is not 16-byte aligned
static bool orb_make_tensor_maps(const OrbGeom& g, const OrbPtrs& P, int nframes, OrbTensorMaps* tm)
{
    for (int l = 0; l < g.nlevels; l++) {
        const OrbLevelGeom& L = g.lv[l];
        if (!plf_tma_map_images(&tm->raw[l], P.lvl[l], 1, L.w, L.h, nframes, (size_t)P.pitch[l], P.frameStride[l], DT_RAW_W, DT_RAW_H)) return false;
        if (!plf_tma_map_images(&tm->blr[l], P.blr[l], 1, L.w, L.h, nframes, (size_t)L.pitch, L.frameBytes, DT_BLR_W, DT_BLR_H)) return false;
    }
    return true;
}
// FAST window maps: one box shape (tp x rows) for every level
static bool orb_make_fast_maps(const OrbGeom& g, const OrbPtrs& P, int nframes, int tp, int rows, OrbFastMaps* fm)
{
    for (int l = 0; l < g.nlevels; l++) {
        const OrbLevelGeom& L = g.lv[l];
        if (!plf_tma_map_images(&fm->m[l], P.lvl[l], 1, L.w, L.h, nframes, (size_t)P.pitch[l], P.frameStride[l], tp, rows)) return false;
    }
    return true;
}
#endif
