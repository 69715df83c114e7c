// plf_fld_kernels.cuh -- the FLD branch of Lineextractor (System.usingLsdFeature: 0): Lineextractor::detect /
// lineDetection / getPointChain / extractSegments / incidentPoint / additionalOperationsOnSegment
// (src/Lineextractor.cc:443-460, 614-905) over cv::Canny and cv::fitLine.
//
//   k_fld_sobel   Sobel 3x3 (BORDER_REPLICATE) + L1 magnitude                            -- cv::Canny, aperture 3
//   k_fld_nms     non-maximum suppression (Q15 tan 22.5 sector test) -> candidate mask bits, strong bits, run-head labels
//   k_ccl_merge   (shared with LSD) 8-connected components of the candidate mask
//   k_fld_flag    components that contain a strong pixel
//   k_fld_edges   hysteresis result = all candidates of flagged components; the two corner squares cleared (:747-748)
//   k_fld_chains  one warp per frame: the raster scan for seeds is warp-wide, chain following + segment fitting follow the
//                 reference's sequential order on lane 0 (a chain erases the pixels it visits, later seeds depend on it)
//
// Arithmetic as the reference compiles it: distPointLine re-normalises the line on EVERY call (the Mat is modified in
// place, :546-557), so the distances of a scan are a dependent chain and are replayed one by one; cross / dot products in
// double without contraction; cv::fitLine DIST_L2 = fitLine2D_wods (moments in double, angle in float, cosf / sinf).
#pragma once

struct FldSeg { float x1, y1, x2, y2, angle; };

__global__ void __launch_bounds__(256)
k_fld_sobel(const uint8_t* __restrict__ img, int w, int h, short* __restrict__ dx, short* __restrict__ dy, int* __restrict__ mag)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t fo = (size_t)blockIdx.z * w * h;
    const uint8_t* I = img + fo;
    const int xm = x > 0 ? x - 1 : 0, xp = x < w - 1 ? x + 1 : w - 1;
    const int ym = y > 0 ? y - 1 : 0, yp = y < h - 1 ? y + 1 : h - 1;
    const uint8_t *r0 = I + (size_t)ym * w, *r1 = I + (size_t)y * w, *r2 = I + (size_t)yp * w;
    const int gx = (r0[xp] + 2 * r1[xp] + r2[xp]) - (r0[xm] + 2 * r1[xm] + r2[xm]);
    const int gy = (r2[xm] + 2 * r2[x] + r2[xp]) - (r0[xm] + 2 * r0[x] + r0[xp]);
    const size_t p = fo + (size_t)y * w + x;
    dx[p] = (short)gx;
    dy[p] = (short)gy;
    mag[p] = (gx < 0 ? -gx : gx) + (gy < 0 ? -gy : gy);
}

// one warp per 32-px row segment: mask word (candidate or strong), strong word, labels = index of the run head inside the word
__global__ void __launch_bounds__(256)
k_fld_nms(const short* __restrict__ dx, const short* __restrict__ dy, const int* __restrict__ mag, int w, int h, int low, int high,
          unsigned* __restrict__ mask, unsigned* __restrict__ strong, int mw, int* __restrict__ label)
{
    const int lane = threadIdx.x, seg = blockIdx.x * 8 + threadIdx.y, y = blockIdx.y, f = blockIdx.z;
    if (seg >= mw) return;
    const int x = seg * 32 + lane;
    const size_t fo = (size_t)f * w * h;
    const int* M = mag + fo;
    bool keep = false, str = false;
    if (x < w) {
        const size_t p = (size_t)y * w + x;
        const int m = M[p];
        if (m > low) {
            const int xs = dx[fo + p], ys = dy[fo + p];
            const int ax = xs < 0 ? -xs : xs, ay = (ys < 0 ? -ys : ys) << 15;
            const int TG22 = 13573;   // (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)
            const int tg22x = ax * TG22;
            auto at = [&](int xx, int yy) { return (xx < 0 || yy < 0 || xx >= w || yy >= h) ? 0 : M[(size_t)yy * w + xx]; };
            if (ay < tg22x) keep = m > at(x - 1, y) && m >= at(x + 1, y);
            else {
                const int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) keep = m > at(x, y - 1) && m >= at(x, y + 1);
                else {
                    const int s = (xs ^ ys) < 0 ? -1 : 1;
                    keep = m > at(x - s, y - 1) && m > at(x + s, y + 1);
                }
            }
            str = keep && m > high;
        }
    }
    const unsigned mk = __ballot_sync(0xffffffffu, keep), ms = __ballot_sync(0xffffffffu, str);
    const size_t wo = ((size_t)f * h + y) * mw + seg;
    if (lane == 0) { mask[wo] = mk; strong[wo] = ms; }
    if (keep) {
        // first pixel of this pixel's run of set bits inside the word
        const unsigned below = ~mk & ((1u << lane) - 1u);            // clear bits below the lane
        const int head = below ? 32 - __clz(below) : 0;
        label[fo + (size_t)y * w + seg * 32 + head] = (int)((size_t)y * w + seg * 32 + head);   // the head labels itself ...
        label[fo + (size_t)y * w + x] = (int)((size_t)y * w + seg * 32 + head);                  // ... and its run
    }
}

__global__ void __launch_bounds__(256)
k_fld_flag(const unsigned* __restrict__ strong, int mw, int w, int h, const int* __restrict__ label, int* __restrict__ flag)
{
    const int lane = threadIdx.x, seg = blockIdx.x * 8 + threadIdx.y, y = blockIdx.y, f = blockIdx.z;
    if (seg >= mw) return;
    const unsigned ms = strong[((size_t)f * h + y) * mw + seg];
    if (!((ms >> lane) & 1u)) return;
    const size_t fo = (size_t)f * w * h;
    const int root = ccl_find(label + fo, y * w + seg * 32 + lane);
    flag[fo + root] = 1;
}

__global__ void __launch_bounds__(256)
k_fld_edges(const unsigned* __restrict__ mask, int mw, int w, int h, const int* __restrict__ label, const int* __restrict__ flag,
            uint8_t* __restrict__ edge)
{
    const int lane = threadIdx.x, seg = blockIdx.x * 8 + threadIdx.y, y = blockIdx.y, f = blockIdx.z;
    if (seg >= mw) return;
    const int x = seg * 32 + lane;
    if (x >= w) return;
    const size_t fo = (size_t)f * w * h;
    const unsigned mk = mask[((size_t)f * h + y) * mw + seg];
    uint8_t e = 0;
    if ((mk >> lane) & 1u) e = flag[fo + ccl_find(label + fo, y * w + x)] ? 255 : 0;
    // canny.colRange(0,6).rowRange(0,6) = 0;  canny.colRange(cols-5,cols).rowRange(rows-5,rows) = 0   (:747-748)
    if ((x < 6 && y < 6) || (x >= w - 5 && y >= h - 5)) e = 0;
    edge[fo + (size_t)y * w + x] = e;
}

// ---------------------------------------------------------------------------------------------------------------------
struct FldLine3 { double a, b, c; };

// Lineextractor::distPointLine (:546-557): normalises l IN PLACE, then l . p
__device__ __forceinline__ double fld_dist(double px, double py, double pz, FldLine3& l)
{
    const double x = l.a, y = l.b;
    const double ww = sqrt(x * x + y * y);
    l.a = x / ww;
    l.b = y / ww;
    l.c = l.c / ww;
    double r = 0;
    r += l.a * px;
    r += l.b * py;
    r += l.c * pz;
    return r;
}
__device__ __forceinline__ FldLine3 fld_cross(double a0, double a1, double a2, double b0, double b1, double b2)
{
    FldLine3 c;
    c.a = a1 * b2 - a2 * b1;
    c.b = a2 * b0 - a0 * b2;
    c.c = a0 * b1 - a1 * b0;
    return c;
}
// cv::fitLine(points, line, DIST_L2, 0, 0.01, 0.01) on integer points -> (vx, vy, x0, y0)
__device__ __forceinline__ void fld_fit_line(const int* pts, int n, float* line)
{
    double x = 0, y = 0, x2 = 0, y2 = 0, xy = 0;
    for (int i = 0; i < n; i++) {
        const float px = (float)(pts[i] & 0xffff), py = (float)(pts[i] >> 16);
        x += px; y += py;
        x2 += px * px; y2 += py * py; xy += px * py;
    }
    const double wn = (float)n;
    x /= wn; y /= wn; x2 /= wn; y2 /= wn; xy /= wn;
    const double dx2 = x2 - x * x, dy2 = y2 - y * y, dxy = xy - x * y;
    const float t = (float)atan2(2 * dxy, dx2 - dy2) / 2;
    line[0] = plf_libm::cosf_glibc(t);
    line[1] = plf_libm::sinf_glibc(t);
    line[2] = (float)x;
    line[3] = (float)y;
}
// Lineextractor::incidentPoint (:595-612): foot of the perpendicular from pt to l, clamped into the image
__device__ __forceinline__ void fld_incident(const FldLine3& l, float ptx, float pty, int iw, int ih, float& ox, float& oy)
{
    const double a0 = (double)ptx, a1 = (double)pty, a2 = 1.0;
    const double b0 = l.a, b1 = l.b, b2 = 0.0;
    const FldLine3 lk = fld_cross(a0, a1, a2, b0, b1, b2);
    FldLine3 xk = fld_cross(lk.a, lk.b, lk.c, l.a, l.b, l.c);
    const double alpha = 1.0 / xk.c;
    xk.a = xk.a * alpha + 0.0; xk.b = xk.b * alpha + 0.0;     // Mat::convertTo(xk, -1, 1 / xk(2))
    const float fx = (float)xk.a, fy = (float)xk.b;
    ox = fx < 0.0f ? 0.0f : (fx >= ((float)iw - 1.0f) ? ((float)iw - 1.0f) : fx);
    oy = fy < 0.0f ? 0.0f : (fy >= ((float)ih - 1.0f) ? ((float)ih - 1.0f) : fy);
}

// one warp per frame.  points / lpts: [frame][w * h] ints (x | y << 16); segs: [frame][segcap]
__global__ void __launch_bounds__(32)
k_fld_chains(uint8_t* __restrict__ edge, const uint8_t* __restrict__ img, int w, int h, int threshold_length, float threshold_dist,
             int* __restrict__ points_all, int* __restrict__ lpts_all, FldSeg* __restrict__ segs_all, int segcap, int* __restrict__ nsegs)
{
    const int f = blockIdx.x, lane = threadIdx.x;
    const size_t fo = (size_t)f * w * h;
    volatile uint8_t* E = edge + fo;
    const uint8_t* I = img + fo;
    int* points = points_all + fo;
    int* lpts = lpts_all + fo;
    FldSeg* out = segs_all + (size_t)f * segcap;
    int nout = 0;
    const unsigned FULL = 0xffffffffu;
    for (int r = 0; r < h; r++) {
        for (int c0 = 0; c0 < w; c0 += 128) {
            // 128 pixels per step: four bytes per lane, requested together
            unsigned nz = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = c0 + 4 * lane + j;
                if (c < w && E[(size_t)r * w + c] != 0) nz |= 1u << j;
            }
            unsigned m = __ballot_sync(FULL, nz != 0);
            while (m) {
                const int sl = __ffs((int)m) - 1;
                const unsigned snz = __shfl_sync(FULL, nz, sl);
                const int cc = c0 + 4 * sl + (__ffs((int)snz) - 1);
                int np = 0;
                {
                    // ---- a seed: follow the chain (lineDetection :760-775, getPointChain :689-740).  The eight neighbours of the
                    // current point are fetched by eight lanes at once; the choice among them is the reference's
                    int px = cc, py = r;
                    if (lane == 0) { points[np] = px | (py << 16); E[(size_t)py * w + px] = 0; }
                    np++;
                    __syncwarp();
                    float direction = 0.0f;
                    int step = 0;
                    const int di = lane & 7;
                    // indices[8][2] = {1,1},{1,0},{1,-1},{0,-1},{-1,-1},{-1,0},{-1,1},{0,1}  (row offset, column offset)
                    const int drow = di < 3 ? 1 : (di == 3 || di == 7) ? 0 : -1;
                    const int dcol = (di == 0 || di == 6 || di == 7) ? 1 : (di == 1 || di == 5) ? 0 : -1;
                    for (;;) {
                        const int ci = px + dcol, ri = py + drow;
                        const bool on = lane < 8 && !(ri < 0 || ri == h || ci < 0 || ci == w) && E[(size_t)ri * w + ci] != 0;
                        const unsigned nb = __ballot_sync(FULL, on);
                        if (!nb) break;
                        int pick;
                        if (step == 0) {
                            pick = __ffs((int)nb) - 1;                         // the first neighbour in table order
                            direction = pick > 4 ? (float)(pick - 8) : (float)pick;
                        } else {
                            const float curr_dir = di > 4 ? (float)(di - 8) : (float)di;
                            float dir_diff = fabsf(curr_dir - direction);
                            dir_diff = dir_diff > 4.0f ? 8.0f - dir_diff : dir_diff;
                            float best = on ? dir_diff : 100.0f;
#pragma unroll
                            for (int o = 4; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(FULL, best, o));
                            best = __shfl_sync(FULL, best, 0);                  // lanes 8..31 reduced over idle lanes: take the real one
                            // `dir_diff <= min_dir_diff` in table order: the LAST neighbour with the smallest difference wins
                            const unsigned eq = __ballot_sync(FULL, on && dir_diff == best) & 0xffu;
                            if (!(best < 2.0f)) break;
                            pick = 31 - __clz((int)eq);
                            const int cdir = pick > 4 ? pick - 8 : pick;
                            direction = (direction * (float)step + (float)cdir) / (float)(step + 1);
                        }
                        px = __shfl_sync(FULL, ci, pick);
                        py = __shfl_sync(FULL, ri, pick);
                        if (lane == 0) { points[np] = px | (py << 16); E[(size_t)py * w + px] = 0; }
                        np++;
                        step++;
                        __syncwarp();
                    }
                }
                if (lane == 0) {
                    if (np >= threshold_length + 1) {
                        // ---- extractSegments (:614-687) on this chain, then the per-segment checks of lineDetection (:789-805)
                        const int total = np;
                        for (int i = 0; i + threshold_length < total; i++) {
                            int ps = points[i], pe = points[i + threshold_length];
                            FldLine3 l = fld_cross((double)(ps & 0xffff), (double)(ps >> 16), 1.0, (double)(pe & 0xffff), (double)(pe >> 16), 1.0);
                            bool is_line = true;
                            int nl = 0;
                            lpts[nl++] = ps;
                            int j;
                            for (j = 1; j < threshold_length; j++) {
                                const int pt = points[i + j];
                                const double dist = fld_dist((double)(pt & 0xffff), (double)(pt >> 16), 1.0, l);
                                if (fabs(dist) > threshold_dist) { is_line = false; break; }
                                lpts[nl++] = pt;
                            }
                            if (!is_line) continue;
                            lpts[nl++] = pe;
                            float line[4];
                            fld_fit_line(lpts, nl, line);
                            l = fld_cross((double)line[2], (double)line[3], 1.0, (double)(line[2] + line[0]), (double)(line[3] + line[1]), 1.0);
                            // incidentPoint(l, ps) with ps an integer point: the result is rounded back to integers
                            float fxs, fys;
                            fld_incident(l, (float)(ps & 0xffff), (float)(ps >> 16), w, h, fxs, fys);
                            int psx = __float2int_rn(fxs), psy = __float2int_rn(fys);
                            int pex = pe & 0xffff, pey = pe >> 16;
                            for (j = threshold_length + 1; i + j < total; j++) {
                                const int pt = points[i + j];
                                double dist = fld_dist((double)(pt & 0xffff), (double)(pt >> 16), 1.0, l);
                                if (fabs(dist) > threshold_dist) {
                                    fld_fit_line(lpts, nl, line);
                                    l = fld_cross((double)line[2], (double)line[3], 1.0, (double)(line[2] + line[0]), (double)(line[3] + line[1]), 1.0);
                                    dist = fld_dist((double)(pt & 0xffff), (double)(pt >> 16), 1.0, l);
                                    if (fabs(dist) > threshold_dist) { j--; break; }
                                }
                                pex = pt & 0xffff; pey = pt >> 16;
                                lpts[nl++] = pt;
                            }
                            fld_fit_line(lpts, nl, line);
                            l = fld_cross((double)line[2], (double)line[3], 1.0, (double)(line[2] + line[0]), (double)(line[3] + line[1]), 1.0);
                            float e1x, e1y, e2x, e2y;
                            fld_incident(l, (float)psx, (float)psy, w, h, e1x, e1y);
                            fld_incident(l, (float)pex, (float)pey, w, h, e2x, e2y);
                            FldSeg seg;
                            seg.x1 = e1x; seg.y1 = e1y; seg.x2 = e2x; seg.y2 = e2y; seg.angle = 0.0f;
                            i = i + j;
                            // ---- lineDetection's checks on the new segment
                            const float length = sqrtf((seg.x1 - seg.x2) * (seg.x1 - seg.x2) + (seg.y1 - seg.y2) * (seg.y1 - seg.y2));
                            if (length < (float)threshold_length) continue;
                            if ((seg.x1 <= 5.0f && seg.x2 <= 5.0f) || (seg.y1 <= 5.0f && seg.y2 <= 5.0f) ||
                                (seg.x1 >= (float)w - 5.0f && seg.x2 >= (float)w - 5.0f) || (seg.y1 >= (float)h - 5.0f && seg.y2 >= (float)h - 5.0f))
                                continue;
                            // ---- additionalOperationsOnSegment (:846-905): orient the segment by the brighter side
                            if (!(seg.x1 == 0.0f && seg.x2 == 0.0f && seg.y1 == 0.0f && seg.y2 == 0.0f)) {
                                seg.angle = (float)((double)(plf_fast_atan2(seg.y2 - seg.y1, seg.x2 - seg.x1) / 180.0f) * 3.1415926535897932384626433832795);
                                const double ang = (double)seg.angle;
                                const double ddx = (double)seg.x2 - (double)seg.x1, ddy = (double)seg.y2 - (double)seg.y1;
                                const double cs = cos(90.0 * 3.1415926535897932384626433832795 / 180.0 + ang), sn = sin(90.0 * 3.1415926535897932384626433832795 / 180.0 + ang);
                                int iR = 0, iL = 0;
                                for (int k = 0; k < 10; k++) {
                                    float qx, qy;
                                    if (k == 0) { qx = seg.x1; qy = seg.y1; }
                                    else if (k == 9) { qx = seg.x2; qy = seg.y2; }
                                    else { qx = seg.x1 + ((float)ddx / 9.0f * (float)k); qy = seg.y1 + ((float)ddy / 9.0f * (float)k); }
                                    int rx = __double2int_rn((double)qx + 1.0 * cs), ry = __double2int_rn((double)qy + 1.0 * sn);
                                    int lx = __double2int_rn((double)qx - 1.0 * cs), ly = __double2int_rn((double)qy - 1.0 * sn);
                                    rx = rx <= 5 ? 5 : rx >= w - 5 ? w - 5 : rx;  ry = ry <= 5 ? 5 : ry >= h - 5 ? h - 5 : ry;
                                    lx = lx <= 5 ? 5 : lx >= w - 5 ? w - 5 : lx;  ly = ly <= 5 ? 5 : ly >= h - 5 ? h - 5 : ly;
                                    iR += I[(size_t)ry * w + rx];
                                    iL += I[(size_t)ly * w + lx];
                                }
                                if (iR > iL) {
                                    float t = seg.x1; seg.x1 = seg.x2; seg.x2 = t;
                                    t = seg.y1; seg.y1 = seg.y2; seg.y2 = t;
                                    seg.angle = (float)((double)(plf_fast_atan2(seg.y2 - seg.y1, seg.x2 - seg.x1) / 180.0f) * 3.1415926535897932384626433832795);
                                }
                            }
                            if (nout < segcap) out[nout] = seg;
                            nout++;
                        }
                    }
                }
                __syncwarp();
                // pixels of this 128-pixel group may have been erased by the chain (the seed itself has been): look again
                nz = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = c0 + 4 * lane + j;
                    if (c < w && E[(size_t)r * w + c] != 0) nz |= 1u << j;
                }
                m = __ballot_sync(FULL, nz != 0);
            }
        }
    }
    if (lane == 0) nsegs[f] = nout;
}
