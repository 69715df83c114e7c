// plf_orb.cu -- host driver of the ORB path: replaces PL_SLAM::ORBextractor
// (include/ORBextractor.h:45-113, src/ORBextractor.cc:410-470, :765-853, :1043-1132).
#include "plf_orb_kernels.cuh"
#include "plf_orb_tma.cuh"
#include <math.h>
#include <vector>

struct plf_orb {
    plf_ctx* ctx;
    plf_orb_params prm;
    float scale[ORB_MAX_LEVELS], inv_scale[ORB_MAX_LEVELS], sigma2[ORB_MAX_LEVELS], inv_sigma2[ORB_MAX_LEVELS];
    int per_level[ORB_MAX_LEVELS];
    int umax[ORB_HALF_PATCH + 2];
    // workspace, valid for (ws_w, ws_h) and up to ws_frames frames
    int ws_w, ws_h, ws_frames;
    OrbGeom geom;
    OrbPtrs ptrs;
    uint8_t* d_levels;      // all level + blurred buffers
    uint8_t* d_lists;       // raw keys, knode, kept, counts
    int2* d_tabs;           // resize tables
    const int2* xtab[ORB_MAX_LEVELS];
    const int2* ytab[ORB_MAX_LEVELS];
    uint8_t* lvl_own[ORB_MAX_LEVELS];
    plf_keypoint* d_kps;    // output staging for the host-buffer entry points
    uint8_t* d_desc;
    int* d_nout;
    int out_frames, out_cap;
    int last_frames;        // frames of the last call (for pyramid/debug access)
    size_t oct_smem;
    int oct_cap;
};

static inline int cv_round_f(float v) { return (int)lrintf(v); }

extern "C" plf_status plf_orb_create(plf_ctx* ctx, const plf_orb_params* p, plf_orb** out)
{
    if (!ctx || !p || !out) return PLF_ERR_INVALID;
    if (p->nlevels < 1 || p->nlevels > ORB_MAX_LEVELS || p->nfeatures < 1 || !(p->scale_factor > 1.0f) ||
        p->min_th_fast < 1 || p->ini_th_fast < p->min_th_fast || p->ini_th_fast > 254)
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_orb_create: unsupported parameters");
    plf_orb* o = (plf_orb*)calloc(1, sizeof(plf_orb));
    o->ctx = ctx;
    o->prm = *p;
    // scale tables and feature split, src/ORBextractor.cc:415-446 (scaleFactor is a double member)
    const int n = p->nlevels;
    const double sf = (double)p->scale_factor;
    o->scale[0] = 1.0f; o->sigma2[0] = 1.0f;
    for (int i = 1; i < n; i++) {
        o->scale[i] = (float)((double)o->scale[i - 1] * sf);
        o->sigma2[i] = o->scale[i] * o->scale[i];
    }
    for (int i = 0; i < n; i++) { o->inv_scale[i] = 1.0f / o->scale[i]; o->inv_sigma2[i] = 1.0f / o->sigma2[i]; }
    float factor = (float)(1.0f / sf);
    float nDesired = (float)(p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n)));
    int sum = 0;
    for (int l = 0; l < n - 1; l++) {
        o->per_level[l] = cv_round_f(nDesired);
        sum += o->per_level[l];
        nDesired *= factor;
    }
    o->per_level[n - 1] = p->nfeatures - sum > 0 ? p->nfeatures - sum : 0;
    // umax, :454-469
    int v, v0, vmax = (int)floor(ORB_HALF_PATCH * sqrt(2.f) / 2 + 1);
    int vmin = (int)ceil(ORB_HALF_PATCH * sqrt(2.f) / 2);
    const double hp2 = ORB_HALF_PATCH * ORB_HALF_PATCH;
    for (v = 0; v <= vmax; ++v) o->umax[v] = (int)lrint(sqrt(hp2 - v * v));
    for (v = ORB_HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (o->umax[v0] == o->umax[v0 + 1]) ++v0;
        o->umax[v] = v0;
        ++v0;
    }
    *out = o;
    return PLF_OK;
}

static void orb_free_ws(plf_orb* o)
{
    if (o->d_levels) cudaFree(o->d_levels);
    if (o->d_lists) cudaFree(o->d_lists);
    if (o->d_tabs) cudaFree(o->d_tabs);
    o->d_levels = nullptr; o->d_lists = nullptr; o->d_tabs = nullptr;
    o->ws_w = o->ws_h = o->ws_frames = 0;
    // nothing of the last batch can be read any more (plf_orb_pyramid_level / plf_stereo_match* check last_frames): a failed
    // re-prepare must not leave them pointing into freed memory
    o->last_frames = 0;
    memset(&o->ptrs, 0, sizeof(o->ptrs));
}

extern "C" void plf_orb_destroy(plf_orb* o)
{
    if (!o) return;
    cudaSetDevice(o->ctx->device);
    cudaStreamSynchronize(o->ctx->stream);
    orb_free_ws(o);
    if (o->d_kps) cudaFree(o->d_kps);
    if (o->d_desc) cudaFree(o->d_desc);
    if (o->d_nout) cudaFree(o->d_nout);
    free(o);
}

extern "C" plf_status plf_orb_tables(const plf_orb* o, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                                     int32_t* per_level)
{
    if (!o) return PLF_ERR_INVALID;
    for (int i = 0; i < o->prm.nlevels; i++) {
        if (scale) scale[i] = o->scale[i];
        if (inv_scale) inv_scale[i] = o->inv_scale[i];
        if (sigma2) sigma2[i] = o->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = o->inv_sigma2[i];
        if (per_level) per_level[i] = o->per_level[i];
    }
    return PLF_OK;
}

static int orb_nodecap(int N) { return (N > 64 ? N : 64) + 16; }

extern "C" int plf_orb_max_keypoints(const plf_orb* o)
{
    if (!o) return 0;
    int s = 0;
    for (int i = 0; i < o->prm.nlevels; i++) s += orb_nodecap(o->per_level[i]);
    return s;
}

// cv::resize INTER_LINEAR coefficient table for one axis
static void linear_table(int ssize, int dsize, int2* tab)
{
    double scale = 1.0 / ((double)dsize / (double)ssize);
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        int a0 = (short)lrintf((1.f - f) * 2048.f), a1 = (short)lrintf(f * 2048.f);
        tab[d].x = s;
        tab[d].y = (a0 & 0xffff) | (a1 << 16);
    }
}

// (re)build geometry + device workspace for images of w x h and up to nframes frames
static plf_status orb_prepare(plf_orb* o, int w, int h, int nframes)
{
    plf_ctx* ctx = o->ctx;
    if (o->ws_w == w && o->ws_h == h && o->ws_frames >= nframes) return PLF_OK;
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    orb_free_ws(o);
    if (w > 4000 || h > 4000) return plf_fail(ctx, PLF_ERR_INVALID, "images larger than 4000 px are not supported");
    OrbGeom& g = o->geom;
    memset(&g, 0, sizeof(g));
    const int n = o->prm.nlevels;
    g.nlevels = n; g.iniTh = o->prm.ini_th_fast; g.minTh = o->prm.min_th_fast;
    for (int i = 0; i <= ORB_HALF_PATCH; i++) g.umax[i] = o->umax[i];
    size_t lvlBytes = 0, tabCount = 0;
    int maxnode = 0;
    for (int l = 0; l < n; l++) {
        OrbLevelGeom& L = g.lv[l];
        L.w = cv_round_f((float)w * o->inv_scale[l]);   // :1112
        L.h = cv_round_f((float)h * o->inv_scale[l]);
        if (L.w < 2 * ORB_EDGE + 7 || L.h < 2 * ORB_EDGE + 7)
            return plf_fail(ctx, PLF_ERR_INVALID, "image too small for %d pyramid levels (level %d is %dx%d)", n, l, L.w, L.h);
        L.pitch = (int)plf_align_up((size_t)L.w, 64);
        L.frameBytes = (size_t)L.pitch * L.h;
        const int maxBorderX = L.w - ORB_MINB, maxBorderY = L.h - ORB_MINB;
        const float width = (float)(maxBorderX - ORB_MINB), height = (float)(maxBorderY - ORB_MINB);
        L.nCols = (int)(width / ORB_CELL_W);
        L.nRows = (int)(height / ORB_CELL_W);
        if (L.nCols < 1 || L.nRows < 1) return plf_fail(ctx, PLF_ERR_INVALID, "level %d too small for a FAST cell", l);
        L.wCell = (int)ceilf(width / L.nCols);
        L.hCell = (int)ceilf(height / L.nRows);
        if (L.wCell + 6 > FAST_MAXC || L.hCell + 6 > FAST_MAXC) return plf_fail(ctx, PLF_ERR_INVALID, "FAST cell too large");
        L.cellBase = g.totalCells;
        g.totalCells += L.nCols * L.nRows;
        L.blurF = plf_strip_interior(L.w, 4, 8);
        L.blurTilesX = plf_div_up(L.blurF, 32) + 1;
        L.blurTileBase = g.totalBlurTiles;
        g.totalBlurTiles += L.blurTilesX * plf_div_up(L.h, BLUR_TH);
        L.nfeat = o->per_level[l];
        L.nodecap = orb_nodecap(L.nfeat);
        L.keptcap = L.nodecap;
        if (L.nodecap > maxnode) maxnode = L.nodecap;
        int px = L.w * L.h;
        L.rawcap = px / 5 > 4096 ? px / 5 : 4096;   // NMS keeps at most one key per 2x2 block; noise reaches ~px/10
        L.scale = o->scale[l];
        L.sizeval = (int)(31 * o->scale[l]);      // :837
        L.rawOff = g.rawPerFrame;
        g.rawPerFrame += (size_t)L.rawcap;
        L.keptOff = g.keptPerFrame;
        g.keptPerFrame += (size_t)L.keptcap;
        g.capPerFrame += L.keptcap;
        lvlBytes += L.frameBytes * 2;   // level + blurred
        if (l > 0) tabCount += (size_t)L.w + L.h;
    }
    o->oct_cap = maxnode;
    o->oct_smem = oct_smem_bytes(maxnode);
    if (o->oct_smem > 200 * 1024) return plf_fail(ctx, PLF_ERR_INVALID, "nfeatures too large for the octree kernel");
    PLF_SMEM_OPTIN(ctx, k_octree);
    // device buffers
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_levels, lvlBytes * nframes));
    size_t listBytes = (size_t)nframes * (g.rawPerFrame * (sizeof(unsigned) + sizeof(unsigned short)) +
                                          g.keptPerFrame * sizeof(int) + (size_t)n * 2 * sizeof(int) + sizeof(int)) + 1024;
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_lists, listBytes));
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_tabs, (tabCount + 1) * sizeof(int2)));
    OrbPtrs& P = o->ptrs;
    memset(&P, 0, sizeof(P));
    uint8_t* q = o->d_levels;
    for (int l = 0; l < n; l++) {
        o->lvl_own[l] = q; q += g.lv[l].frameBytes * nframes;
        P.blr[l] = q; q += g.lv[l].frameBytes * nframes;
        P.lvl[l] = o->lvl_own[l];
        P.frameStride[l] = g.lv[l].frameBytes;
        P.pitch[l] = g.lv[l].pitch;
    }
    uint8_t* r = o->d_lists;
    P.rawkeys = (unsigned*)r; r += (size_t)nframes * g.rawPerFrame * sizeof(unsigned);
    P.kept = (int*)r; r += (size_t)nframes * g.keptPerFrame * sizeof(int);
    P.rawcount = (int*)r; r += (size_t)nframes * n * sizeof(int);
    P.keptcount = (int*)r; r += (size_t)nframes * n * sizeof(int);
    P.status = (int*)r; r += (size_t)nframes * sizeof(int);
    P.knode = (unsigned short*)r;
    // resize tables
    std::vector<int2> tabs(tabCount + 1);
    size_t to = 0;
    for (int l = 1; l < n; l++) {
        linear_table(g.lv[l - 1].w, g.lv[l].w, &tabs[to]);
        o->xtab[l] = o->d_tabs + to; to += g.lv[l].w;
        linear_table(g.lv[l - 1].h, g.lv[l].h, &tabs[to]);
        o->ytab[l] = o->d_tabs + to; to += g.lv[l].h;
    }
    PLF_CUDA(ctx, cudaMemcpy(o->d_tabs, tabs.data(), tabCount * sizeof(int2), cudaMemcpyHostToDevice));
    o->ws_w = w; o->ws_h = h; o->ws_frames = nframes;
    return PLF_OK;
}

// all kernels of one batch; level 0 is read from (lvl0, stride0, frameStride0)
static plf_status orb_run(plf_orb* o, const uint8_t* lvl0, size_t stride0, size_t frameStride0, int nframes,
                          plf_keypoint* d_kps, uint8_t* d_desc, int cap, int* d_nout)
{
    plf_ctx* ctx = o->ctx;
    const OrbGeom& g = o->geom;
    OrbPtrs P = o->ptrs;
    P.lvl[0] = lvl0; P.frameStride[0] = frameStride0; P.pitch[0] = (int)stride0;
    o->ptrs.lvl[0] = lvl0; o->ptrs.frameStride[0] = frameStride0; o->ptrs.pitch[0] = (int)stride0;
    cudaStream_t st = ctx->stream;
    PLF_CUDA(ctx, cudaMemsetAsync(P.rawcount, 0, (size_t)nframes * g.nlevels * sizeof(int), st));
    for (int l = 1; l < g.nlevels; l++) {
        const OrbLevelGeom &S = g.lv[l - 1], &D = g.lv[l];
        // (a shared-memory tiled variant measured slower than these L1-served gathers: 2.3 vs 1.8 ms per 1024 frames)
        dim3 grid(plf_div_up(D.w, 128), plf_div_up(D.h, 8 * RL_ROWS), nframes);
        PLF_LAUNCH(k_resize_linear, grid, dim3(32, 8), 0, st, P.lvl[l - 1], P.frameStride[l - 1], P.pitch[l - 1], S.w, S.h,
                   o->lvl_own[l], D.frameBytes, D.pitch, D.w, D.h, o->xtab[l], o->ytab[l]);
        PLF_CHECK_LAUNCH(ctx);
    }
    int fw = 0, fh = 0;   // largest FAST cell window of the geometry -> shared-memory map size
    for (int l = 0; l < g.nlevels; l++) { fw = g.lv[l].wCell + 6 > fw ? g.lv[l].wCell + 6 : fw; fh = g.lv[l].hCell + 6 > fh ? g.lv[l].hCell + 6 : fh; }
    const int ftp = (fw + 3 + 3) & ~3;
    const int flcap = ((fw - 6) * (fh - 6) + 1) & ~1;   // work-list entries: one per interior pixel
    const size_t fsmem = (size_t)FAST_WARPS * (2 * ftp * fh + 2 * flcap);
#ifndef PLF_EMU
    static const bool no_tma = getenv("PLF_NO_TMA") != nullptr;
    // cell windows by TMA box loads, double buffered per warp; k_fast_cells (register-staged loads) when a caller-owned level 0
    // is not 16-byte aligned or the driver does not offer cuTensorMapEncodeTiled
    const int ttp = (fw + 15 + 15) & ~15;                 // box width: the window plus the <= 15 columns before it
    const int trows = (fh + 7) & ~7;                      // ttp * trows is a multiple of 128
    const size_t tsmem = (size_t)FAST_WARPS * (3 * (size_t)ttp * trows + (((size_t)2 * flcap + 127) & ~(size_t)127));
    OrbFastMaps fm;
    bool fast_tma = !no_tma && ttp <= 256 && trows <= 256 && orb_make_fast_maps(g, P, nframes, ttp, trows, &fm);
    if (fast_tma) {
        const int cpw = 8;                                // cells per warp: long enough to hide the first load, short enough to balance
        PLF_SMEM_OPTIN(ctx, k_fast_cells_tma);
        PLF_LAUNCH(k_fast_cells_tma, dim3(plf_div_up(plf_div_up(g.totalCells, cpw), FAST_WARPS), nframes), dim3(32 * FAST_WARPS), tsmem, st, g, P, fm,
                   ttp, trows, flcap, cpw);
        PLF_CHECK_LAUNCH(ctx);
    } else
#endif
    {
        PLF_SMEM_OPTIN(ctx, k_fast_cells);
        PLF_LAUNCH(k_fast_cells, dim3(plf_div_up(g.totalCells, FAST_WARPS), nframes), dim3(32 * FAST_WARPS), fsmem, st, g, P, ftp, fh, flcap);
        PLF_CHECK_LAUNCH(ctx);
    }
    PLF_LAUNCH(k_blur7, dim3(plf_div_up(g.totalBlurTiles, BLUR_WARPS), nframes), dim3(32 * BLUR_WARPS), 0, st, g, P);
    PLF_CHECK_LAUNCH(ctx);
    PLF_LAUNCH(k_octree, dim3(g.nlevels, nframes), dim3(OCT_T), o->oct_smem, st, g, P, o->oct_cap);
    PLF_CHECK_LAUNCH(ctx);
#ifndef PLF_EMU
    // the two patches of every keypoint staged in shared memory by TMA tile loads; k_describe (direct gathers) when a caller-owned
    // level 0 is not 16-byte aligned or the driver does not offer cuTensorMapEncodeTiled
    OrbTensorMaps tm;
    if (!no_tma && orb_make_tensor_maps(g, P, nframes, &tm)) {
        PLF_LAUNCH(k_describe_tma, dim3(plf_div_up(cap, 8), nframes), dim3(256), 0, st, g, P, tm, d_kps, d_desc, cap, d_nout);
        PLF_CHECK_LAUNCH(ctx);
    } else
#endif
    {
        PLF_LAUNCH(k_describe, dim3(plf_div_up(cap, 8), nframes), dim3(256), 0, st, g, P, d_kps, d_desc, cap, d_nout);
        PLF_CHECK_LAUNCH(ctx);
    }
    o->last_frames = nframes;
    return PLF_OK;
}

extern "C" plf_status plf_orb_extract_batch_device(plf_orb* o, const uint8_t* dev_imgs, int nframes, int w, int h,
                                                   size_t stride, size_t frame_stride, plf_keypoint* dev_kps,
                                                   uint8_t* dev_desc, int cap, int32_t* dev_n_out)
{
    if (!o) return PLF_ERR_INVALID;
    plf_ctx* ctx = o->ctx;
    if (!dev_imgs || nframes < 1 || w <= 0 || h <= 0 || stride < (size_t)w || !dev_kps || !dev_desc || !dev_n_out || cap < 1)
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_orb_extract_batch_device: bad arguments");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    plf_status st = orb_prepare(o, w, h, nframes);
    if (st) return st;
    return orb_run(o, dev_imgs, stride, frame_stride, nframes, dev_kps, dev_desc, cap, dev_n_out);
}

static plf_status orb_out_staging(plf_orb* o, int nframes, int cap)
{
    plf_ctx* ctx = o->ctx;
    if (o->out_frames >= nframes && o->out_cap == cap) return PLF_OK;
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (o->d_kps) cudaFree(o->d_kps);
    if (o->d_desc) cudaFree(o->d_desc);
    if (o->d_nout) cudaFree(o->d_nout);
    o->d_kps = nullptr; o->d_desc = nullptr; o->d_nout = nullptr; o->out_frames = 0;
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_kps, (size_t)nframes * cap * sizeof(plf_keypoint)));
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_desc, (size_t)nframes * cap * 32));
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_nout, (size_t)nframes * sizeof(int)));
    o->out_frames = nframes; o->out_cap = cap;
    return PLF_OK;
}

static plf_status orb_extract_to_host(plf_orb* o, const uint8_t* host_imgs, bool device_src, int nframes, int w, int h, size_t stride,
                                      size_t frame_stride, plf_keypoint* host_kps, uint8_t* host_desc, int cap, int32_t* n_out)
{
    if (!o) return PLF_ERR_INVALID;
    plf_ctx* ctx = o->ctx;
    if (nframes < 1 || !n_out) return plf_fail(ctx, PLF_ERR_INVALID, "plf_orb_extract_batch: bad arguments");
    if (!host_imgs || w <= 0 || h <= 0) {   // empty image: silent return, src/ORBextractor.cc:1046-1047
        for (int f = 0; f < nframes; f++) n_out[f] = 0;
        return PLF_OK;
    }
    if (stride < (size_t)w || !host_kps || !host_desc || cap < 1)
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_orb_extract_batch: bad arguments");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    plf_status st = orb_prepare(o, w, h, nframes);
    if (st) return st;
    st = orb_out_staging(o, nframes, cap);
    if (st) return st;
    const OrbLevelGeom& L0 = o->geom.lv[0];
    cudaStream_t s = ctx->stream;
    // level 0 is kept densely packed (pitch = width) when the width is a multiple of 4, so that a contiguous host batch
    // is ONE DMA transfer; other widths keep the 64-byte aligned pitch (the kernels' 32-bit fast paths need it).  The
    // level-0 buffer was sized for the aligned pitch, which is at least as large.
    const size_t p0 = (w & 3) ? (size_t)L0.pitch : (size_t)w;
    if (device_src) {
        // the images are on the device already (uploaded once for several extractors: plf_upload): level 0 is the caller's buffer
        st = orb_run(o, host_imgs, stride, frame_stride, nframes, o->d_kps, o->d_desc, cap, o->d_nout);
        if (st) return st;
    } else {
    if (stride == p0 && frame_stride == p0 * h) {
        // in pieces of 16 MB, so that transfers queued by other (higher-priority) contexts can slip in between
        const size_t total = (size_t)nframes * frame_stride, piece = (size_t)16 << 20;
        for (size_t off = 0; off < total; off += piece)
            PLF_CUDA(ctx, cudaMemcpyAsync(o->lvl_own[0] + off, host_imgs + off, total - off < piece ? total - off : piece, cudaMemcpyHostToDevice, s));
    } else {
        for (int f = 0; f < nframes; f++)
            PLF_CUDA(ctx, cudaMemcpy2DAsync(o->lvl_own[0] + (size_t)f * p0 * h, p0, host_imgs + (size_t)f * frame_stride,
                                            stride, w, h, cudaMemcpyHostToDevice, s));
    }
    st = orb_run(o, o->lvl_own[0], p0, p0 * h, nframes, o->d_kps, o->d_desc, cap, o->d_nout);
    if (st) return st;
    }
    void* pin;   // the counts go through pinned staging (a copy into pageable memory stalls the other contexts' host threads)
    st = plf_ctx_pinned(ctx, (size_t)nframes * sizeof(int), &pin);
    if (st) return st;
    PLF_CUDA(ctx, cudaMemcpyAsync(pin, o->d_nout, (size_t)nframes * sizeof(int), cudaMemcpyDeviceToHost, s));
    PLF_CUDA(ctx, cudaMemcpyAsync(host_kps, o->d_kps, (size_t)nframes * cap * sizeof(plf_keypoint), cudaMemcpyDeviceToHost, s));
    PLF_CUDA(ctx, cudaMemcpyAsync(host_desc, o->d_desc, (size_t)nframes * cap * 32, cudaMemcpyDeviceToHost, s));
    ctx->blocking = nframes >= 32;      // batches sleep in their waits, single frames spin (latency)
    { plf_status ss = plf_sync(ctx, s); if (ss) return ss; }
    memcpy(n_out, pin, (size_t)nframes * sizeof(int));
    for (int f = 0; f < nframes; f++) {
        if (n_out[f] == -1) return plf_fail(ctx, PLF_ERR_CAPACITY, "frame %d: internal key/node list overflow", f);
        if (n_out[f] == -2) return plf_fail(ctx, PLF_ERR_CAPACITY, "frame %d: output capacity %d too small (use plf_orb_max_keypoints)", f, cap);
    }
    return PLF_OK;
}

extern "C" plf_status plf_orb_extract_batch(plf_orb* o, const uint8_t* host_imgs, int nframes, int w, int h, size_t stride,
                                            size_t frame_stride, plf_keypoint* host_kps, uint8_t* host_desc, int cap,
                                            int32_t* n_out)
{
    return orb_extract_to_host(o, host_imgs, false, nframes, w, h, stride, frame_stride, host_kps, host_desc, cap, n_out);
}

extern "C" plf_status plf_orb_extract_batch_from_device(plf_orb* o, const uint8_t* dev_imgs, int nframes, int w, int h, size_t stride,
                                                        size_t frame_stride, plf_keypoint* host_kps, uint8_t* host_desc, int cap,
                                                        int32_t* n_out)
{
    return orb_extract_to_host(o, dev_imgs, true, nframes, w, h, stride, frame_stride, host_kps, host_desc, cap, n_out);
}

extern "C" plf_status plf_orb_extract(plf_orb* o, const uint8_t* host_img, int w, int h, size_t stride, plf_keypoint* host_kps,
                                      uint8_t* host_desc, int cap, int* n_out)
{
    if (!o || !n_out) return PLF_ERR_INVALID;
    int32_t n = 0;
    plf_status st = plf_orb_extract_batch(o, host_img, 1, w, h, stride, stride * (size_t)(h > 0 ? h : 0), host_kps, host_desc, cap, &n);
    *n_out = n;
    return st;
}

extern "C" plf_status plf_orb_pyramid_level(plf_orb* o, int frame, int level, uint8_t* host_dst, size_t dst_stride, int* w, int* h)
{
    if (!o) return PLF_ERR_INVALID;
    plf_ctx* ctx = o->ctx;
    if (o->last_frames < 1) return plf_fail(ctx, PLF_ERR_STATE, "no extraction has run yet");
    if (frame < 0 || frame >= o->last_frames || level < 0 || level >= o->geom.nlevels)
        return plf_fail(ctx, PLF_ERR_INVALID, "frame/level out of range");
    const OrbLevelGeom& L = o->geom.lv[level];
    if (w) *w = L.w;
    if (h) *h = L.h;
    if (!host_dst) return PLF_OK;
    if (dst_stride < (size_t)L.w) return plf_fail(ctx, PLF_ERR_INVALID, "dst_stride too small");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_CUDA(ctx, cudaMemcpy2DAsync(host_dst, dst_stride, o->ptrs.lvl[level] + (size_t)frame * o->ptrs.frameStride[level],
                                    o->ptrs.pitch[level], L.w, L.h, cudaMemcpyDeviceToHost, ctx->stream));
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

extern "C" plf_status plf_orb_debug_blurred(plf_orb* o, int frame, int level, uint8_t* host_dst, size_t dst_stride)
{
    if (!o) return PLF_ERR_INVALID;
    plf_ctx* ctx = o->ctx;
    if (o->last_frames < 1) return plf_fail(ctx, PLF_ERR_STATE, "no extraction has run yet");
    if (frame < 0 || frame >= o->last_frames || level < 0 || level >= o->geom.nlevels || !host_dst)
        return plf_fail(ctx, PLF_ERR_INVALID, "frame/level out of range");
    const OrbLevelGeom& L = o->geom.lv[level];
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_CUDA(ctx, cudaMemcpy2DAsync(host_dst, dst_stride, o->ptrs.blr[level] + (size_t)frame * L.frameBytes, L.pitch, L.w, L.h,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

extern "C" plf_status plf_orb_debug_raw_keys(plf_orb* o, int frame, int level, int32_t* xs, int32_t* ys, int32_t* resp, int cap, int* n_out)
{
    if (!o || !n_out) return PLF_ERR_INVALID;
    plf_ctx* ctx = o->ctx;
    if (o->last_frames < 1) return plf_fail(ctx, PLF_ERR_STATE, "no extraction has run yet");
    if (frame < 0 || frame >= o->last_frames || level < 0 || level >= o->geom.nlevels)
        return plf_fail(ctx, PLF_ERR_INVALID, "frame/level out of range");
    const OrbGeom& g = o->geom;
    const OrbLevelGeom& L = g.lv[level];
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    int n = 0;
    PLF_CUDA(ctx, cudaMemcpyAsync(&n, o->ptrs.rawcount + (size_t)frame * g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n > L.rawcap) return plf_fail(ctx, PLF_ERR_CAPACITY, "raw key list overflowed (%d > %d)", n, L.rawcap);
    if (n > cap) return plf_fail(ctx, PLF_ERR_CAPACITY, "cap too small for %d raw keys", n);
    std::vector<unsigned> keys((size_t)n + 1);
    PLF_CUDA(ctx, cudaMemcpyAsync(keys.data(), o->ptrs.rawkeys + (size_t)frame * g.rawPerFrame + L.rawOff, (size_t)n * sizeof(unsigned),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) {
        xs[i] = keys[i] & 0xfff; ys[i] = (keys[i] >> 12) & 0xfff; resp[i] = keys[i] >> 24;
    }
    *n_out = n;
    return PLF_OK;
}

extern "C" plf_status plf_orb_distribute_octree(plf_ctx* ctx, const int32_t* xs, const int32_t* ys, const int32_t* resp, int n,
                                                int minX, int maxX, int minY, int maxY, int N, int32_t* out_idx, int cap, int* n_out)
{
    if (!ctx || !n_out || n < 0 || N < 0 || maxX <= minX || maxY <= minY || (n > 0 && (!xs || !ys || !resp || !out_idx)))
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_orb_distribute_octree: bad arguments");
    *n_out = 0;
    if (n == 0) return PLF_OK;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    // the order key needs the FAST cell grid of the level (SURVEY.md hard part 4)
    const float width = (float)(maxX - minX), height = (float)(maxY - minY);
    int nCols = (int)(width / ORB_CELL_W), nRows = (int)(height / ORB_CELL_W);
    if (nCols < 1 || nRows < 1) return plf_fail(ctx, PLF_ERR_INVALID, "region too small");
    int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
    std::vector<unsigned> keys(n);
    for (int i = 0; i < n; i++) {
        if (xs[i] < 0 || xs[i] > 4095 || ys[i] < 0 || ys[i] > 4095 || resp[i] < 0 || resp[i] > 255)
            return plf_fail(ctx, PLF_ERR_INVALID, "key %d out of range", i);
        keys[i] = (unsigned)xs[i] | ((unsigned)ys[i] << 12) | ((unsigned)resp[i] << 24);
    }
    int nodecap = orb_nodecap(N);
    size_t smem = oct_smem_bytes(nodecap);
    if (smem > 200 * 1024) return plf_fail(ctx, PLF_ERR_INVALID, "N too large");
    PLF_SMEM_OPTIN(ctx, k_octree_single);
    size_t bytes = (size_t)n * 4 + (size_t)n * 2 + 16 + (size_t)nodecap * 4 + 16;
    void* s;
    plf_status st = plf_ctx_scratch(ctx, bytes + 64, &s);
    if (st) return st;
    unsigned* dk = (unsigned*)s;
    int* dout = (int*)(dk + n);
    int* dcnt = dout + nodecap;
    unsigned short* dn = (unsigned short*)(dcnt + 4);
    PLF_CUDA(ctx, cudaMemcpyAsync(dk, keys.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    PLF_LAUNCH(k_octree_single, dim3(1), dim3(OCT_T), smem, ctx->stream, (const unsigned*)dk, n, dn, maxX - minX, maxY - minY, N,
               wCell, hCell, dout, nodecap, dcnt, nodecap);
    PLF_CHECK_LAUNCH(ctx);
    int cnt = 0;
    PLF_CUDA(ctx, cudaMemcpyAsync(&cnt, dcnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (cnt < 0) return plf_fail(ctx, PLF_ERR_CAPACITY, "octree node list overflow");
    if (cnt > cap) return plf_fail(ctx, PLF_ERR_CAPACITY, "out_idx capacity too small");
    PLF_CUDA(ctx, cudaMemcpyAsync(out_idx, dout, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = cnt;
    return PLF_OK;
}

// ---------------- Frame::ComputeStereoMatches (src/Frame.cc:881-1055) ----------------
#include "plf_stereo_kernels.cuh"

static void stereo_side(const plf_orb* o, StereoSide* S)
{
    memset(S, 0, sizeof(*S));
    for (int l = 0; l < o->geom.nlevels; l++) {
        S->lvl[l] = o->ptrs.lvl[l]; S->frameStride[l] = o->ptrs.frameStride[l]; S->pitch[l] = o->ptrs.pitch[l];
        S->w[l] = o->geom.lv[l].w; S->h[l] = o->geom.lv[l].h;
    }
}

static plf_status stereo_run(plf_orb* left, plf_orb* right, int npairs, StereoSide SL, StereoSide SR, int cap, float mb, float mbf,
                             float* d_uright, float* d_depth, int* d_sad)
{
    plf_ctx* ctx = left->ctx;
    StereoTables T;
    memset(&T, 0, sizeof(T));
    T.nlevels = left->geom.nlevels;
    for (int l = 0; l < T.nlevels; l++) { T.scale[l] = left->scale[l]; T.inv_scale[l] = left->inv_scale[l]; }
    if (right->ctx != ctx) {
        plf_status st = plf_ctx_wait(ctx, right->ctx);   // the right extractor's pyramid must be complete
        if (st) return st;
    }
    PLF_LAUNCH(k_stereo_match, dim3(plf_div_up(cap, 8), npairs), dim3(256), 0, ctx->stream, SL, SR, T, cap, mb, mbf, d_uright, d_depth, d_sad);
    PLF_CHECK_LAUNCH(ctx);
    PLF_LAUNCH(k_stereo_filter, dim3(npairs), dim3(256), (size_t)cap * sizeof(int), ctx->stream, cap, d_uright, d_depth, (const int*)d_sad);
    PLF_CHECK_LAUNCH(ctx);
    return PLF_OK;
}

static plf_status stereo_check(plf_orb* left, plf_orb* right, int cap)
{
    plf_ctx* ctx = left->ctx;
    if (left->last_frames < 1 || right->last_frames < 1) return plf_fail(ctx, PLF_ERR_STATE, "stereo matching needs both extractors' pyramids (run the extraction first)");
    if (left->ctx->device != right->ctx->device) return plf_fail(ctx, PLF_ERR_INVALID, "both extractors must live on the same device");
    if (left->geom.nlevels != right->geom.nlevels || left->ws_w != right->ws_w || left->ws_h != right->ws_h)
        return plf_fail(ctx, PLF_ERR_INVALID, "left and right extractors differ in image size or pyramid levels");
    if (cap < 1 || cap > 0xffff || (size_t)cap * sizeof(int) > 200 * 1024) return plf_fail(ctx, PLF_ERR_INVALID, "stereo matching supports up to 51200 keypoints per image");
    return PLF_OK;
}

extern "C" plf_status plf_stereo_match_batch_device(plf_orb* left, plf_orb* right, int npairs, int left_first, int left_step,
                                                    int right_first, int right_step, const plf_keypoint* dev_kl, const uint8_t* dev_dl,
                                                    const int32_t* dev_nl, const plf_keypoint* dev_kr, const uint8_t* dev_dr,
                                                    const int32_t* dev_nr, int cap, float mb, float mbf, float* dev_uright, float* dev_depth)
{
    if (!left || !right) return PLF_ERR_INVALID;
    plf_ctx* ctx = left->ctx;
    if (npairs < 1 || !dev_kl || !dev_dl || !dev_nl || !dev_kr || !dev_dr || !dev_nr || !dev_uright || !dev_depth || !(mb > 0))
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_stereo_match_batch_device: bad arguments");
    plf_status st = stereo_check(left, right, cap);
    if (st) return st;
    const int lastL = left_first + (npairs - 1) * left_step, lastR = right_first + (npairs - 1) * right_step;
    if (left_first < 0 || lastL < 0 || left_first >= left->last_frames || lastL >= left->last_frames || right_first < 0 || lastR < 0 ||
        right_first >= right->last_frames || lastR >= right->last_frames)
        return plf_fail(ctx, PLF_ERR_INVALID, "stereo pair frames outside the extractors' last batch");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_SMEM_OPTIN(ctx, k_stereo_filter);
    void* s;
    st = plf_ctx_scratch(ctx, (size_t)npairs * cap * sizeof(int), &s);
    if (st) return st;
    StereoSide SL, SR;
    stereo_side(left, &SL); stereo_side(right, &SR);
    SL.kps = dev_kl; SL.desc = dev_dl; SL.n = dev_nl; SL.first = left_first; SL.step = left_step;
    SR.kps = dev_kr; SR.desc = dev_dr; SR.n = dev_nr; SR.first = right_first; SR.step = right_step;
    return stereo_run(left, right, npairs, SL, SR, cap, mb, mbf, dev_uright, dev_depth, (int*)s);
}

extern "C" plf_status plf_stereo_match(plf_orb* left, int frame_l, plf_orb* right, int frame_r, const plf_keypoint* host_kl,
                                       const uint8_t* host_dl, int nl, const plf_keypoint* host_kr, const uint8_t* host_dr, int nr,
                                       float mb, float mbf, float* host_uright, float* host_depth)
{
    if (!left || !right) return PLF_ERR_INVALID;
    plf_ctx* ctx = left->ctx;
    if (nl < 0 || nr < 0 || (nl > 0 && (!host_kl || !host_dl || !host_uright || !host_depth)) || (nr > 0 && (!host_kr || !host_dr)) || !(mb > 0))
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_stereo_match: bad arguments");
    if (nl == 0) return PLF_OK;
    const int cap = nl > nr ? nl : nr;
    plf_status st = stereo_check(left, right, cap);
    if (st) return st;
    if (frame_l < 0 || frame_l >= left->last_frames || frame_r < 0 || frame_r >= right->last_frames)
        return plf_fail(ctx, PLF_ERR_INVALID, "stereo frames outside the extractors' last batch");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_SMEM_OPTIN(ctx, k_stereo_filter);
    // scratch layout: kps L, kps R, desc L, desc R, n[2], uright, depth, sad
    const size_t kb = plf_align_up((size_t)cap * sizeof(plf_keypoint), 256), db = plf_align_up((size_t)cap * 32, 256), fb = plf_align_up((size_t)cap * 4, 256);
    void* s;
    st = plf_ctx_scratch(ctx, 2 * kb + 2 * db + 256 + 3 * fb, &s);
    if (st) return st;
    uint8_t* p = (uint8_t*)s;
    plf_keypoint* dkl = (plf_keypoint*)p; p += kb;
    plf_keypoint* dkr = (plf_keypoint*)p; p += kb;
    uint8_t* ddl = p; p += db;
    uint8_t* ddr = p; p += db;
    int* dn = (int*)p; p += 256;
    float* du = (float*)p; p += fb;
    float* dz = (float*)p; p += fb;
    int* dsad = (int*)p;
    cudaStream_t stq = ctx->stream;
    const int nn[2] = {nl, nr};
    PLF_CUDA(ctx, cudaMemcpyAsync(dkl, host_kl, (size_t)nl * sizeof(plf_keypoint), cudaMemcpyHostToDevice, stq));
    PLF_CUDA(ctx, cudaMemcpyAsync(ddl, host_dl, (size_t)nl * 32, cudaMemcpyHostToDevice, stq));
    if (nr > 0) {
        PLF_CUDA(ctx, cudaMemcpyAsync(dkr, host_kr, (size_t)nr * sizeof(plf_keypoint), cudaMemcpyHostToDevice, stq));
        PLF_CUDA(ctx, cudaMemcpyAsync(ddr, host_dr, (size_t)nr * 32, cudaMemcpyHostToDevice, stq));
    }
    PLF_CUDA(ctx, cudaMemcpyAsync(dn, nn, sizeof(nn), cudaMemcpyHostToDevice, stq));
    StereoSide SL, SR;
    stereo_side(left, &SL); stereo_side(right, &SR);
    // the keypoint tables of this call hold one frame each; the pyramid frames are frame_l / frame_r
    for (int l = 0; l < left->geom.nlevels; l++) {
        SL.lvl[l] += (size_t)frame_l * SL.frameStride[l];
        SR.lvl[l] += (size_t)frame_r * SR.frameStride[l];
    }
    SL.kps = dkl; SL.desc = ddl; SL.n = dn; SL.first = 0; SL.step = 0;
    SR.kps = dkr; SR.desc = ddr; SR.n = dn + 1; SR.first = 0; SR.step = 0;
    // frame index 0 for the tables, so the right count is read through n[0] of its own pointer
    st = stereo_run(left, right, 1, SL, SR, cap, mb, mbf, du, dz, dsad);
    if (st) return st;
    PLF_CUDA(ctx, cudaMemcpyAsync(host_uright, du, (size_t)nl * sizeof(float), cudaMemcpyDeviceToHost, stq));
    PLF_CUDA(ctx, cudaMemcpyAsync(host_depth, dz, (size_t)nl * sizeof(float), cudaMemcpyDeviceToHost, stq));
    PLF_CUDA(ctx, cudaStreamSynchronize(stq));
    return PLF_OK;
}
