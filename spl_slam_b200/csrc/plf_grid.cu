// plf_grid.cu -- the feature grid of a Frame and its area queries on the device:
// Frame::AssignFeaturesToGrid / AssignFeaturesToGridLines (src/Frame.cc:365-399) with PosInGrid / PosInGridLines
// (:677-722), and Frame::GetFeaturesInArea / GetFeaturesInAreaLines (:562-676).  The queries produce the candidate
// lists (CSR) that plf_hamming_candidates_device scans, in exactly the reference's order (cell column outer, cell row
// inner, insertion order inside a cell), because the matchers' "first best wins" rule depends on it.
#include "plf_common.cuh"
#include "plf_sort.cuh"      // k_scan_top: single-CTA exclusive scan (the candidate offsets are a few thousand entries)

// cell of a point; false when it falls outside the grid (undistorted coordinates may leave the image)
__device__ __forceinline__ bool grid_pos(float x, float y, const plf_grid_params& g, int& cx, int& cy)
{
    cx = (int)roundf((x - g.min_x) * g.inv_w);
    cy = (int)roundf((y - g.min_y) * g.inv_h);
    return !(cx < 0 || cx >= g.cols || cy < 0 || cy >= g.rows);
}

// one CTA per frame: cell of every feature, counts, exclusive scan, stable placement (ascending feature index per cell)
#define GRID_T 256
__global__ void __launch_bounds__(GRID_T)
k_grid_build(const plf_keypoint* __restrict__ kps, const plf_keyline* __restrict__ kls, const int* __restrict__ nfeat, int cap,
             plf_grid_params g, int* __restrict__ cell_start, int* __restrict__ cell_items, int* __restrict__ feat_cell)
{
    PLF_DYN_SMEM(smem);
    int* cnt = (int*)smem;                    // cols * rows + 1
    const int f = blockIdx.x, tid = threadIdx.x;
    const int ncell = g.cols * g.rows;
    int n = nfeat[f];
    if (n > cap) n = cap;
    const plf_keypoint* K = kps + (size_t)f * cap;
    int* FC = feat_cell + (size_t)f * cap;
    for (int c = tid; c <= ncell; c += GRID_T) cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += GRID_T) {
        int cx, cy;
        bool ok = grid_pos(K[i].x, K[i].y, g, cx, cy);
        if (ok && kls) {   // lines: both end points must fall inside the grid as well (PosInGridLines)
            const plf_keyline L = kls[(size_t)f * cap + i];
            int ax, ay;
            ok = grid_pos(L.startPointX, L.startPointY, g, ax, ay) && grid_pos(L.endPointX, L.endPointY, g, ax, ay);
        }
        const int c = ok ? cx * g.rows + cy : -1;
        FC[i] = c;
        if (c >= 0) atomicAdd(&cnt[c], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int c = 0; c < ncell; c++) { const int v = cnt[c]; cnt[c] = acc; acc += v; }
        cnt[ncell] = acc;
    }
    __syncthreads();
    int* CS = cell_start + (size_t)f * (ncell + 1);
    for (int c = tid; c <= ncell; c += GRID_T) CS[c] = cnt[c];
    // stable placement: rank of feature i inside its cell = number of earlier features of the same cell
    int* CI = cell_items + (size_t)f * cap;
    for (int i = tid; i < n; i += GRID_T) {
        const int c = FC[i];
        if (c < 0) continue;
        int rank = 0;
        for (int j = 0; j < i; j++) rank += (FC[j] == c);
        CI[cnt[c] + rank] = i;
    }
}

// Frame::GetFeaturesInArea for one query per thread; pass 0 counts, pass 1 writes the candidates
__global__ void __launch_bounds__(128)
k_grid_query(const plf_keypoint* __restrict__ kps, plf_grid_params g, const int* __restrict__ cell_start, const int* __restrict__ cell_items,
             const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qr, const int* __restrict__ qminl,
             const int* __restrict__ qmaxl, int nq, int* __restrict__ counts, const int* __restrict__ off, int* __restrict__ cand, int candcap, int pass)
{
    const int q = blockIdx.x * 128 + threadIdx.x;
    if (q >= nq) return;
    const float x = qx[q], y = qy[q], r = qr[q];
    const int minLevel = qminl ? qminl[q] : -1, maxLevel = qmaxl ? qmaxl[q] : -1;
    int n = 0;
    int o = pass ? off[q] : 0;
    const int nMinCellX = max(0, (int)floorf((x - g.min_x - r) * g.inv_w));
    const int nMaxCellX = min(g.cols - 1, (int)ceilf((x - g.min_x + r) * g.inv_w));
    const int nMinCellY = max(0, (int)floorf((y - g.min_y - r) * g.inv_h));
    const int nMaxCellY = min(g.rows - 1, (int)ceilf((y - g.min_y + r) * g.inv_h));
    if (!(nMinCellX >= g.cols || nMaxCellX < 0 || nMinCellY >= g.rows || nMaxCellY < 0)) {
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const int c = ix * g.rows + iy;
                for (int j = cell_start[c]; j < cell_start[c + 1]; j++) {
                    const int idx = cell_items[j];
                    const plf_keypoint kp = kps[idx];
                    if (bCheckLevels) {
                        if (kp.octave < minLevel) continue;
                        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                    }
                    const float distx = kp.x - x, disty = kp.y - y;
                    if (fabsf(distx) < r && fabsf(disty) < r) {
                        if (pass && o + n < candcap) cand[o + n] = idx;
                        n++;
                    }
                }
            }
    }
    if (!pass) counts[q] = n;
}

static plf_status grid_check(plf_ctx* ctx, const plf_grid_params* g)
{
    if (!g || g->cols < 1 || g->rows < 1 || (long)g->cols * g->rows > 11000 || !(g->inv_w > 0) || !(g->inv_h > 0))
        return plf_fail(ctx, PLF_ERR_INVALID, "bad grid parameters (at most 11000 cells)");
    return PLF_OK;
}

extern "C" plf_status plf_grid_build_device(plf_ctx* ctx, const plf_keypoint* dev_kps, const plf_keyline* dev_kls, const int32_t* dev_n,
                                            int nframes, int cap, const plf_grid_params* g, int32_t* dev_cell_start, int32_t* dev_cell_items,
                                            int32_t* dev_feat_cell)
{
    if (!ctx) return PLF_ERR_INVALID;
    if (!dev_kps || !dev_n || nframes < 1 || cap < 1 || !dev_cell_start || !dev_cell_items || !dev_feat_cell)
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_grid_build_device: bad arguments");
    plf_status st = grid_check(ctx, g);
    if (st) return st;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_LAUNCH(k_grid_build, dim3(nframes), dim3(GRID_T), (size_t)(g->cols * g->rows + 1) * sizeof(int), ctx->stream, dev_kps, dev_kls, dev_n, cap,
               *g, dev_cell_start, dev_cell_items, dev_feat_cell);
    PLF_CHECK_LAUNCH(ctx);
    return PLF_OK;
}

extern "C" plf_status plf_grid_query_device(plf_ctx* ctx, const plf_keypoint* dev_kps, const plf_grid_params* g, const int32_t* dev_cell_start,
                                            const int32_t* dev_cell_items, const float* dev_qx, const float* dev_qy, const float* dev_qr,
                                            const int32_t* dev_qminl, const int32_t* dev_qmaxl, int nq, int32_t* dev_cand_off,
                                            int32_t* dev_cand_idx, int cand_cap, int* total)
{
    if (!ctx) return PLF_ERR_INVALID;
    if (!dev_kps || !dev_cell_start || !dev_cell_items || nq < 0 || (nq > 0 && (!dev_qx || !dev_qy || !dev_qr)) || !dev_cand_off || !total ||
        cand_cap < 0 || (cand_cap > 0 && !dev_cand_idx))
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_grid_query_device: bad arguments");
    plf_status st = grid_check(ctx, g);
    if (st) return st;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    *total = 0;
    cudaStream_t s = ctx->stream;
    if (nq == 0) { PLF_CUDA(ctx, cudaMemsetAsync(dev_cand_off, 0, sizeof(int), s)); return PLF_OK; }
    // counts -> exclusive scan (nq + 1 entries) -> fill
    void* scr;
    st = plf_ctx_scratch(ctx, (size_t)(nq + 1) * sizeof(int) + 512, &scr);
    if (st) return st;
    int* counts = (int*)scr;
    PLF_CUDA(ctx, cudaMemsetAsync(counts + nq, 0, sizeof(int), s));
    PLF_LAUNCH(k_grid_query, dim3(plf_div_up(nq, 128)), dim3(128), 0, s, dev_kps, *g, dev_cell_start, dev_cell_items, dev_qx, dev_qy, dev_qr,
               dev_qminl, dev_qmaxl, nq, counts, (const int*)nullptr, (int*)nullptr, 0, 0);
    PLF_CHECK_LAUNCH(ctx);
    PLF_CUDA(ctx, cudaMemcpyAsync(dev_cand_off, counts, (size_t)(nq + 1) * sizeof(int), cudaMemcpyDeviceToDevice, s));
    PLF_LAUNCH(k_scan_top, dim3(1), dim3(SC_T), 0, s, dev_cand_off, nq + 1);       // exclusive scan in place
    PLF_CHECK_LAUNCH(ctx);
    void* pin;
    st = plf_ctx_pinned(ctx, 64, &pin);
    if (st) return st;
    PLF_CUDA(ctx, cudaMemcpyAsync(pin, dev_cand_off + nq, sizeof(int), cudaMemcpyDeviceToHost, s));
    PLF_CUDA(ctx, cudaStreamSynchronize(s));
    *total = *(int*)pin;
    if (*total > cand_cap) return plf_fail(ctx, PLF_ERR_CAPACITY, "candidate buffer too small: %d candidates, capacity %d", *total, cand_cap);
    if (*total > 0) {
        PLF_LAUNCH(k_grid_query, dim3(plf_div_up(nq, 128)), dim3(128), 0, s, dev_kps, *g, dev_cell_start, dev_cell_items, dev_qx, dev_qy, dev_qr,
                   dev_qminl, dev_qmaxl, nq, (int*)nullptr, (const int*)dev_cand_off, dev_cand_idx, cand_cap, 1);
        PLF_CHECK_LAUNCH(ctx);
    }
    return PLF_OK;
}

// host-buffer convenience: grid of one frame + area queries -> CSR candidate lists (what the Search* matchers iterate over)
extern "C" plf_status plf_grid_candidates(plf_ctx* ctx, const plf_keypoint* host_kps, const plf_keyline* host_kls, int n, const plf_grid_params* g,
                                          const float* host_qx, const float* host_qy, const float* host_qr, const int32_t* host_qminl,
                                          const int32_t* host_qmaxl, int nq, int32_t* host_cand_off, int32_t* host_cand_idx, int cand_cap, int* total)
{
    if (!ctx) return PLF_ERR_INVALID;
    if (n < 0 || nq < 0 || (n > 0 && !host_kps) || (nq > 0 && (!host_qx || !host_qy || !host_qr)) || !host_cand_off || !total)
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_grid_candidates: bad arguments");
    plf_status st = grid_check(ctx, g);
    if (st) return st;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    const int ncell = g->cols * g->rows, cap = n > 0 ? n : 1;
    const size_t kb = plf_align_up((size_t)cap * sizeof(plf_keypoint), 256), lb = plf_align_up((size_t)cap * sizeof(plf_keyline), 256);
    const size_t ib = plf_align_up((size_t)cap * 4, 256), cb = plf_align_up((size_t)(ncell + 1) * 4, 256), qb = plf_align_up((size_t)(nq + 1) * 4, 256);
    const size_t ob = plf_align_up((size_t)(cand_cap > 0 ? cand_cap : 1) * 4, 256);
    void* io = nullptr;
    st = plf_ctx_ioscratch(ctx, kb + lb + 2 * ib + cb + 256 + 6 * qb + ob, &io);     // no cudaMalloc / cudaFree per call
    if (st) return st;
    uint8_t* p = (uint8_t*)io;
    plf_keypoint* dk = (plf_keypoint*)p; p += kb;
    plf_keyline* dl = (plf_keyline*)p; p += lb;
    int* items = (int*)p; p += ib;
    int* fcell = (int*)p; p += ib;
    int* cstart = (int*)p; p += cb;
    int* dn = (int*)p; p += 256;
    float* dqx = (float*)p; p += qb;
    float* dqy = (float*)p; p += qb;
    float* dqr = (float*)p; p += qb;
    int* dmin = (int*)p; p += qb;
    int* dmax = (int*)p; p += qb;
    int* doff = (int*)p; p += qb;
    int* dcand = (int*)p;
    cudaStream_t s = ctx->stream;
    cudaError_t e = cudaSuccess;
    if (n > 0) e = cudaMemcpyAsync(dk, host_kps, (size_t)n * sizeof(plf_keypoint), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && n > 0 && host_kls) e = cudaMemcpyAsync(dl, host_kls, (size_t)n * sizeof(plf_keyline), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dn, &n, sizeof(int), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && nq > 0) {
        e = cudaMemcpyAsync(dqx, host_qx, (size_t)nq * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dqy, host_qy, (size_t)nq * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dqr, host_qr, (size_t)nq * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && host_qminl) e = cudaMemcpyAsync(dmin, host_qminl, (size_t)nq * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && host_qmaxl) e = cudaMemcpyAsync(dmax, host_qmaxl, (size_t)nq * 4, cudaMemcpyHostToDevice, s);
    }
    if (e != cudaSuccess) st = plf_fail(ctx, PLF_ERR_CUDA, "grid upload failed: %s", cudaGetErrorString(e));
    if (!st) st = plf_grid_build_device(ctx, dk, host_kls ? dl : nullptr, dn, 1, cap, g, cstart, items, fcell);
    if (!st) st = plf_grid_query_device(ctx, dk, g, cstart, items, dqx, dqy, dqr, host_qminl ? dmin : nullptr, host_qmaxl ? dmax : nullptr, nq, doff,
                                        dcand, cand_cap, total);
    if (!st) {
        e = cudaMemcpyAsync(host_cand_off, doff, (size_t)(nq + 1) * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && *total > 0) e = cudaMemcpyAsync(host_cand_idx, dcand, (size_t)*total * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) st = plf_fail(ctx, PLF_ERR_CUDA, "grid result copy failed: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(s);
    }
    return st;
}

// ---------------- Frame::UndistortKeyPoints / UndistortKeyLines (src/Frame.cc:733-826) ----------------
// cv::undistortPoints(pts, pts, K, distCoef, Mat(), K) on float points: normalise, five fixed-point iterations of the
// radial (k1, k2, k3) + tangential (p1, p2) model in double, re-project with K, round to float.  Same operation
// order as OpenCV's cvUndistortPointsInternal (verified bit-for-bit against cv2 4.13 through the oracle).
__device__ __forceinline__ void undistort_point(const plf_camera& c, float px, float py, float& ox, float& oy)
{
    const double fx = (double)c.fx, fy = (double)c.fy, cx = (double)c.cx, cy = (double)c.cy;
    const double k1 = (double)c.k[0], k2 = (double)c.k[1], p1 = (double)c.k[2], p2 = (double)c.k[3], k3 = c.nk > 4 ? (double)c.k[4] : 0.0;
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    const double u = (double)px, v = (double)py;
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((0.0 * r2 + 0.0) * r2 + 0.0) * r2) / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
        if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
        const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
        const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    ox = (float)(fx * x + cx);
    oy = (float)(fy * y + cy);
}

__global__ void k_undistort_keypoints(plf_camera c, const plf_keypoint* __restrict__ in, int n, plf_keypoint* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    plf_keypoint k = in[i];
    if (c.k[0] != 0.0f) undistort_point(c, k.x, k.y, k.x, k.y);     // mDistCoef.at<float>(0) == 0 -> plain copy (:735-739)
    out[i] = k;
}

__global__ void k_undistort_keylines(plf_camera c, const plf_keyline* __restrict__ in, int n, plf_keyline* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    plf_keyline k = in[i];
    if (c.k[0] != 0.0f) {
        undistort_point(c, k.startPointX, k.startPointY, k.startPointX, k.startPointY);
        undistort_point(c, k.endPointX, k.endPointY, k.endPointX, k.endPointY);
    }
    out[i] = k;
}

static plf_status cam_check(plf_ctx* ctx, const plf_camera* c)
{
    if (!c || !(c->fx != 0) || !(c->fy != 0) || c->nk < 4 || c->nk > 5) return plf_fail(ctx, PLF_ERR_INVALID, "bad camera parameters (4 or 5 distortion coefficients)");
    return PLF_OK;
}

extern "C" plf_status plf_undistort_keypoints_device(plf_ctx* ctx, const plf_camera* cam, const plf_keypoint* dev_in, int n, plf_keypoint* dev_out)
{
    if (!ctx) return PLF_ERR_INVALID;
    if (n < 0 || (n > 0 && (!dev_in || !dev_out))) return plf_fail(ctx, PLF_ERR_INVALID, "plf_undistort_keypoints_device: bad arguments");
    plf_status st = cam_check(ctx, cam);
    if (st || n == 0) return st;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_LAUNCH(k_undistort_keypoints, dim3(plf_div_up(n, 128)), dim3(128), 0, ctx->stream, *cam, dev_in, n, dev_out);
    PLF_CHECK_LAUNCH(ctx);
    return PLF_OK;
}

extern "C" plf_status plf_undistort_keylines_device(plf_ctx* ctx, const plf_camera* cam, const plf_keyline* dev_kl, const plf_keypoint* dev_mid, int n,
                                                    plf_keyline* dev_kl_out, plf_keypoint* dev_mid_out)
{
    if (!ctx) return PLF_ERR_INVALID;
    if (n < 0 || (n > 0 && (!dev_kl || !dev_mid || !dev_kl_out || !dev_mid_out))) return plf_fail(ctx, PLF_ERR_INVALID, "plf_undistort_keylines_device: bad arguments");
    plf_status st = cam_check(ctx, cam);
    if (st || n == 0) return st;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_LAUNCH(k_undistort_keypoints, dim3(plf_div_up(n, 128)), dim3(128), 0, ctx->stream, *cam, dev_mid, n, dev_mid_out);
    PLF_CHECK_LAUNCH(ctx);
    PLF_LAUNCH(k_undistort_keylines, dim3(plf_div_up(n, 128)), dim3(128), 0, ctx->stream, *cam, dev_kl, n, dev_kl_out);
    PLF_CHECK_LAUNCH(ctx);
    return PLF_OK;
}

extern "C" plf_status plf_undistort_keypoints(plf_ctx* ctx, const plf_camera* cam, const plf_keypoint* host_in, int n, plf_keypoint* host_out)
{
    if (!ctx) return PLF_ERR_INVALID;
    if (n < 0 || (n > 0 && (!host_in || !host_out))) return plf_fail(ctx, PLF_ERR_INVALID, "plf_undistort_keypoints: bad arguments");
    plf_status st = cam_check(ctx, cam);
    if (st || n == 0) return st;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    void* s;
    const size_t kb = plf_align_up((size_t)n * sizeof(plf_keypoint), 256);
    st = plf_ctx_scratch(ctx, 2 * kb, &s);
    if (st) return st;
    plf_keypoint* din = (plf_keypoint*)s;
    plf_keypoint* dout = (plf_keypoint*)((uint8_t*)s + kb);
    PLF_CUDA(ctx, cudaMemcpyAsync(din, host_in, (size_t)n * sizeof(plf_keypoint), cudaMemcpyHostToDevice, ctx->stream));
    st = plf_undistort_keypoints_device(ctx, cam, din, n, dout);
    if (st) return st;
    PLF_CUDA(ctx, cudaMemcpyAsync(host_out, dout, (size_t)n * sizeof(plf_keypoint), cudaMemcpyDeviceToHost, ctx->stream));
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

extern "C" plf_status plf_undistort_keylines(plf_ctx* ctx, const plf_camera* cam, const plf_keyline* host_kl, const plf_keypoint* host_mid, int n,
                                             plf_keyline* host_kl_out, plf_keypoint* host_mid_out)
{
    if (!ctx) return PLF_ERR_INVALID;
    if (n < 0 || (n > 0 && (!host_kl || !host_mid || !host_kl_out || !host_mid_out))) return plf_fail(ctx, PLF_ERR_INVALID, "plf_undistort_keylines: bad arguments");
    plf_status st = cam_check(ctx, cam);
    if (st || n == 0) return st;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    void* s;
    const size_t kb = plf_align_up((size_t)n * sizeof(plf_keypoint), 256), lb = plf_align_up((size_t)n * sizeof(plf_keyline), 256);
    st = plf_ctx_scratch(ctx, 2 * kb + 2 * lb, &s);
    if (st) return st;
    uint8_t* p = (uint8_t*)s;
    plf_keypoint* dmi = (plf_keypoint*)p; p += kb;
    plf_keypoint* dmo = (plf_keypoint*)p; p += kb;
    plf_keyline* dli = (plf_keyline*)p; p += lb;
    plf_keyline* dlo = (plf_keyline*)p;
    PLF_CUDA(ctx, cudaMemcpyAsync(dmi, host_mid, (size_t)n * sizeof(plf_keypoint), cudaMemcpyHostToDevice, ctx->stream));
    PLF_CUDA(ctx, cudaMemcpyAsync(dli, host_kl, (size_t)n * sizeof(plf_keyline), cudaMemcpyHostToDevice, ctx->stream));
    st = plf_undistort_keylines_device(ctx, cam, dli, dmi, n, dlo, dmo);
    if (st) return st;
    PLF_CUDA(ctx, cudaMemcpyAsync(host_mid_out, dmo, (size_t)n * sizeof(plf_keypoint), cudaMemcpyDeviceToHost, ctx->stream));
    PLF_CUDA(ctx, cudaMemcpyAsync(host_kl_out, dlo, (size_t)n * sizeof(plf_keyline), cudaMemcpyDeviceToHost, ctx->stream));
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}
