// plf_fld_impl.cuh -- host side of the FLD branch (included at the end of plf_line.cu: it reuses k_pyrdown, k_ccl_merge and
// the LBD path of plf_line).  Replaces Lineextractor's FLD constructor and ComputeFldWithLbd (src/Lineextractor.cc:69-110,
// 242-336) with Lineextractor::ComputePyramid / detectFldWithPyramid / detect (:413-460) beneath it.
#pragma once
#include <algorithm>
#include <vector>

struct plf_fld {
    plf_ctx* ctx;
    plf_fld_params prm;
    int per_level[16];
    plf_line* lbd;                // BinaryDescriptor::compute
    int ws_w, ws_h;
    int lw[LINE_MAX_OCT], lh[LINE_MAX_OCT], mw[LINE_MAX_OCT];
    uint8_t* d_base;
    uint8_t *d_img[LINE_MAX_OCT], *d_edge[LINE_MAX_OCT];
    short *d_dx, *d_dy;
    int *d_mag, *d_label, *d_flag, *d_points, *d_lpts, *d_nsegs;
    unsigned *d_mask, *d_strong;
    FldSeg* d_segs;
    FldSeg* h_segs;               // pinned
    int* h_nsegs;
};
#define FLD_SEGCAP 8192

extern "C" plf_status plf_fld_create(plf_ctx* ctx, const plf_fld_params* p, plf_fld** out)
{
    if (!ctx || !p || !out) return PLF_ERR_INVALID;
    // CV_Assert(_length_threshold > 0 && _distance_threshold > 0 && _canny_th1 > 0 && _canny_th2 > 0 && _canny_aperture_size > 0) (:74-75)
    if (p->nlevels < 1 || p->nlevels > LINE_MAX_OCT || p->nfeatures < 1 || !(p->scale > 0) || p->length_threshold <= 0 ||
        !(p->distance_threshold > 0) || !(p->canny_th1 > 0) || !(p->canny_th2 > 0) || p->canny_aperture_size != 3 || p->do_merge)
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_fld_create: unsupported parameters (nlevels 1..2, canny aperture 3, do_merge 0)");
    plf_fld* o = (plf_fld*)calloc(1, sizeof(plf_fld));
    o->ctx = ctx; o->prm = *p;
    const int n = p->nlevels;
    float factor = (float)(1.0f / p->scale);
    float nDesired = (float)(p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n)));
    int sum = 0;
    for (int l = 0; l < n - 1; l++) { o->per_level[l] = (int)lrintf(nDesired); sum += o->per_level[l]; nDesired *= factor; }
    o->per_level[n - 1] = p->nfeatures - sum > 0 ? p->nfeatures - sum : 0;
    plf_line_params lp;
    memset(&lp, 0, sizeof lp);
    lp.nfeatures = p->nfeatures; lp.nlevels = n; lp.refine = 0; lp.scale = 1.0; lp.sigma_scale = 0.6; lp.quant = 2.0; lp.ang_th = 22.5;
    lp.log_eps = 0; lp.density_th = 0.7; lp.n_bins = 1024; lp.min_line_length = 0;
    plf_status st = plf_line_create(ctx, &lp, &o->lbd);
    if (st) { free(o); return st; }
    *out = o;
    return PLF_OK;
}

static void fld_free_ws(plf_fld* o)
{
    cudaFree(o->d_base); o->d_base = nullptr;
    if (o->h_segs) cudaFreeHost(o->h_segs);
    o->h_segs = nullptr; o->h_nsegs = nullptr;
    o->ws_w = o->ws_h = 0;
}

extern "C" void plf_fld_destroy(plf_fld* o)
{
    if (!o) return;
    cudaSetDevice(o->ctx->device);
    fld_free_ws(o);
    plf_line_destroy(o->lbd);
    free(o);
}

extern "C" plf_status plf_fld_features_per_level(const plf_fld* o, int32_t* per_level)
{
    if (!o || !per_level) return PLF_ERR_INVALID;
    for (int i = 0; i < o->prm.nlevels; i++) per_level[i] = o->per_level[i];
    return PLF_OK;
}

static plf_status fld_prepare(plf_fld* o, int w, int h)
{
    plf_ctx* ctx = o->ctx;
    if (o->ws_w == w && o->ws_h == h) return PLF_OK;
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    fld_free_ws(o);
    const int n = o->prm.nlevels;
    size_t bytes = 0;
    auto need = [&](size_t count, size_t elt) { bytes += plf_align_up(count * elt, 256); };
    for (int k = 0; k < n; k++) {
        o->lw[k] = w >> k; o->lh[k] = h >> k; o->mw[k] = plf_div_up(o->lw[k], 32);
        if (o->lw[k] < 16 || o->lh[k] < 16) return plf_fail(ctx, PLF_ERR_INVALID, "image too small for %d FLD levels", n);
        need((size_t)o->lw[k] * o->lh[k], 1); need((size_t)o->lw[k] * o->lh[k], 1);
    }
    const size_t px = (size_t)w * h, words = (size_t)o->mw[0] * h;
    need(px, 2); need(px, 2); need(px, 4); need(px, 4); need(px, 4); need(px, 4); need(px, 4);   // dx dy mag label flag points lpts
    need(words, 4); need(words, 4); need(FLD_SEGCAP, sizeof(FldSeg)); need(16, 4);
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_base, bytes + 4096));
    uint8_t* p = o->d_base;
    for (int k = 0; k < n; k++) {
        o->d_img[k] = carve<uint8_t>(p, (size_t)o->lw[k] * o->lh[k]);
        o->d_edge[k] = carve<uint8_t>(p, (size_t)o->lw[k] * o->lh[k]);
    }
    o->d_dx = carve<short>(p, px); o->d_dy = carve<short>(p, px); o->d_mag = carve<int>(p, px); o->d_label = carve<int>(p, px);
    o->d_flag = carve<int>(p, px); o->d_points = carve<int>(p, px); o->d_lpts = carve<int>(p, px);
    o->d_mask = carve<unsigned>(p, words); o->d_strong = carve<unsigned>(p, words);
    o->d_segs = carve<FldSeg>(p, FLD_SEGCAP); o->d_nsegs = carve<int>(p, 16);
    PLF_CUDA(ctx, cudaMallocHost((void**)&o->h_segs, FLD_SEGCAP * sizeof(FldSeg) + 64));
    o->h_nsegs = (int*)((uint8_t*)o->h_segs + FLD_SEGCAP * sizeof(FldSeg));
    o->ws_w = w; o->ws_h = h;
    return PLF_OK;
}

// Lineextractor::detect on every pyramid level (detectFldWithPyramid): segs[level] = the level's segments in detection order
static plf_status fld_detect_levels(plf_fld* o, const uint8_t* host_img, int w, int h, size_t stride, std::vector<std::vector<FldSeg> >& segs)
{
    plf_ctx* ctx = o->ctx;
    cudaStream_t st = ctx->stream;
    plf_status ps = fld_prepare(o, w, h);
    if (ps) return ps;
    const int n = o->prm.nlevels;
    PLF_CUDA(ctx, cudaMemcpy2DAsync(o->d_img[0], (size_t)w, host_img, stride, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, st));
    segs.assign(n, std::vector<FldSeg>());
    const int low0 = (int)floor(o->prm.canny_th1), high0 = (int)floor(o->prm.canny_th2);
    const int low = low0 < high0 ? low0 : high0, high = low0 < high0 ? high0 : low0;
    for (int k = 0; k < n; k++) {
        const int lw = o->lw[k], lh = o->lh[k], mw = o->mw[k];
        if (k > 0) {   // ComputePyramid (:413-431): pyrDown, no pre-blur
            const int pdF = pyrdown_interior(o->lw[k - 1]);
            PLF_LAUNCH(k_pyrdown, dim3(plf_div_up(pdF, 32) + 1, plf_div_up(lh, 4 * PD_ROWS), 1), dim3(32, 4), 0, st, (const uint8_t*)o->d_img[k - 1],
                       (size_t)o->lw[k - 1] * o->lh[k - 1], o->lw[k - 1], o->lw[k - 1], o->lh[k - 1], o->d_img[k], (size_t)lw * lh, lw, pdF,
                       plf_div_up(pdF, 32));
            PLF_CHECK_LAUNCH(ctx);
        }
        PLF_LAUNCH(k_fld_sobel, dim3(plf_div_up(lw, 32), plf_div_up(lh, 8), 1), dim3(32, 8), 0, st, (const uint8_t*)o->d_img[k], lw, lh, o->d_dx, o->d_dy, o->d_mag);
        PLF_CHECK_LAUNCH(ctx);
        PLF_CUDA(ctx, cudaMemsetAsync(o->d_flag, 0, (size_t)lw * lh * sizeof(int), st));
        PLF_LAUNCH(k_fld_nms, dim3(plf_div_up(mw, 8), lh, 1), dim3(32, 8), 0, st, (const short*)o->d_dx, (const short*)o->d_dy, (const int*)o->d_mag, lw, lh,
                   low, high, o->d_mask, o->d_strong, mw, o->d_label);
        PLF_CHECK_LAUNCH(ctx);
        // the CCL kernel addresses pixels as y * w + x with 32-pixel mask words per row: w must be the row pitch of the labels
        PLF_LAUNCH(k_ccl_merge, dim3(plf_div_up(lh, 8), 1, 1), dim3(32, 8), 0, st, o->d_label, (const unsigned*)o->d_mask, mw, lw, lh);
        PLF_CHECK_LAUNCH(ctx);
        PLF_LAUNCH(k_fld_flag, dim3(plf_div_up(mw, 8), lh, 1), dim3(32, 8), 0, st, (const unsigned*)o->d_strong, mw, lw, lh, (const int*)o->d_label, o->d_flag);
        PLF_CHECK_LAUNCH(ctx);
        PLF_LAUNCH(k_fld_edges, dim3(plf_div_up(mw, 8), lh, 1), dim3(32, 8), 0, st, (const unsigned*)o->d_mask, mw, lw, lh, (const int*)o->d_label,
                   (const int*)o->d_flag, o->d_edge[k]);
        PLF_CHECK_LAUNCH(ctx);
        PLF_LAUNCH(k_fld_chains, dim3(1), dim3(32), 0, st, o->d_edge[k], (const uint8_t*)o->d_img[k], lw, lh, o->prm.length_threshold,
                   o->prm.distance_threshold, o->d_points, o->d_lpts, o->d_segs, FLD_SEGCAP, o->d_nsegs);
        PLF_CHECK_LAUNCH(ctx);
        PLF_CUDA(ctx, cudaMemcpyAsync(o->h_nsegs, o->d_nsegs, sizeof(int), cudaMemcpyDeviceToHost, st));
        PLF_CUDA(ctx, cudaStreamSynchronize(st));
        const int ns = o->h_nsegs[0];
        if (ns > FLD_SEGCAP) return plf_fail(ctx, PLF_ERR_CAPACITY, "more than %d FLD segments on level %d", FLD_SEGCAP, k);
        if (ns > 0) {
            PLF_CUDA(ctx, cudaMemcpyAsync(o->h_segs, o->d_segs, (size_t)ns * sizeof(FldSeg), cudaMemcpyDeviceToHost, st));
            PLF_CUDA(ctx, cudaStreamSynchronize(st));
            segs[k].assign(o->h_segs, o->h_segs + ns);
        }
    }
    return PLF_OK;
}

// Lineextractor::detect(image, lines) of the single-level form: n x 4 floats
extern "C" plf_status plf_fld_detect(plf_fld* o, const uint8_t* host_img, int w, int h, size_t stride, float* host_lines, int cap, int* n_out)
{
    if (!o || !n_out) return PLF_ERR_INVALID;
    plf_ctx* ctx = o->ctx;
    *n_out = 0;
    if (!host_img || w <= 0 || h <= 0 || stride < (size_t)w || !host_lines) return plf_fail(ctx, PLF_ERR_INVALID, "plf_fld_detect: bad arguments");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<std::vector<FldSeg> > segs;
    plf_status st = fld_detect_levels(o, host_img, w, h, stride, segs);
    if (st) return st;
    const std::vector<FldSeg>& s0 = segs[0];
    if ((int)s0.size() > cap) return plf_fail(ctx, PLF_ERR_CAPACITY, "line capacity %d too small for %d segments", cap, (int)s0.size());
    for (size_t i = 0; i < s0.size(); i++) { host_lines[4 * i] = s0[i].x1; host_lines[4 * i + 1] = s0[i].y1; host_lines[4 * i + 2] = s0[i].x2; host_lines[4 * i + 3] = s0[i].y2; }
    *n_out = (int)s0.size();
    return PLF_OK;
}

namespace {
struct FldVec4 { float v[4]; };
// Lineextractor::sort_flines_by_length (include/Lineextractor.h:84-89): std::sort with this comparator is unstable; the host's
// libstdc++ is the implementation the reference itself would run with
struct FldByLength {
    inline bool operator()(const FldVec4& a, const FldVec4& b)
    {
        return (sqrt(pow(a.v[0] - a.v[2], 2.0) + pow(a.v[1] - a.v[3], 2.0)) > sqrt(pow(b.v[0] - b.v[2], 2.0) + pow(b.v[1] - b.v[3], 2.0)));
    }
};
}

// Lineextractor::ComputeFldWithLbd (:242-336).  KeyLine.pt is left unset by the reference (an uninitialised field of a local
// KeyLine); here it is 0.  `keypoints` of the reference is appended to, `keyLines` replaced: the caller passes fresh arrays.
extern "C" plf_status plf_fld_extract(plf_fld* o, const uint8_t* host_img, int w, int h, size_t stride, plf_keyline* host_kl,
                                      plf_keypoint* host_mid, uint8_t* host_desc, int cap, int* n_out)
{
    if (!o || !n_out) return PLF_ERR_INVALID;
    plf_ctx* ctx = o->ctx;
    *n_out = 0;
    if (!host_img || w <= 0 || h <= 0) return PLF_OK;     // empty image: silent return (:245-246)
    if (stride < (size_t)w || !host_kl || !host_desc || cap < 1) return plf_fail(ctx, PLF_ERR_INVALID, "plf_fld_extract: bad arguments");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<std::vector<FldSeg> > segs;
    plf_status st = fld_detect_levels(o, host_img, w, h, stride, segs);
    if (st) return st;
    int n = 0, class_counter = -1;
    for (int oct = 0; oct < o->prm.nlevels; oct++) {
        std::vector<FldVec4> lines(segs[oct].size());
        for (size_t i = 0; i < lines.size(); i++) { lines[i].v[0] = segs[oct][i].x1; lines[i].v[1] = segs[oct][i].y1; lines[i].v[2] = segs[oct][i].x2; lines[i].v[3] = segs[oct][i].y2; }
        if (lines.size() > (size_t)o->per_level[oct]) {
            std::sort(lines.begin(), lines.end(), FldByLength());
            lines.resize(o->per_level[oct]);
        }
        const int cw = o->lw[oct], ch = o->lh[oct];
        const float octaveScale = (float)pow((double)(float)o->prm.scale, (double)oct);     // pow((float)scale, octaveIdy), :296
        for (size_t k = 0; k < lines.size(); k++) {
            float e[4] = {lines[k].v[0], lines[k].v[1], lines[k].v[2], lines[k].v[3]};
            if (e[0] < 0) e[0] = 0;
            if (e[0] >= cw) e[0] = (float)cw - 1.0f;
            if (e[2] < 0) e[2] = 0;
            if (e[2] >= cw) e[2] = (float)cw - 1.0f;
            if (e[1] < 0) e[1] = 0;
            if (e[1] >= ch) e[1] = (float)ch - 1.0f;
            if (e[3] < 0) e[3] = 0;
            if (e[3] >= ch) e[3] = (float)ch - 1.0f;
            plf_keyline K;
            memset(&K, 0, sizeof K);
            K.startPointX = e[0] * octaveScale; K.startPointY = e[1] * octaveScale; K.endPointX = e[2] * octaveScale; K.endPointY = e[3] * octaveScale;
            K.sPointInOctaveX = e[0]; K.sPointInOctaveY = e[1]; K.ePointInOctaveX = e[2]; K.ePointInOctaveY = e[3];
            K.lineLength = (float)sqrt(pow((double)(e[0] - e[2]), 2.0) + pow((double)(e[1] - e[3]), 2.0));
            const int ax = (int)lrintf(e[0]), ay = (int)lrintf(e[1]), bx = (int)lrintf(e[2]), by = (int)lrintf(e[3]);
            const int adx = abs(bx - ax), ady = abs(by - ay);
            K.numOfPixels = (adx > ady ? adx : ady) + 1;                                  // cv::LineIterator(...).count
            K.angle = atan2f(K.endPointY - K.startPointY, K.endPointX - K.startPointX);   // atan2(float, float), host libm as in the reference
            K.class_id = ++class_counter;
            K.octave = oct;
            K.size = (K.endPointX - K.startPointX) * (K.endPointY - K.startPointY);
            K.response = K.lineLength / (float)(cw > ch ? cw : ch);
            if (n >= cap) return plf_fail(ctx, PLF_ERR_CAPACITY, "keyline capacity %d too small", cap);
            host_kl[n] = K;
            if (host_mid) {
                plf_keypoint P;
                P.x = (K.startPointX + K.endPointX) / 2; P.y = (K.startPointY + K.endPointY) / 2;
                P.size = 0; P.angle = -1; P.response = 0; P.octave = K.octave; P.class_id = -1;
                host_mid[n] = P;
            }
            n++;
        }
    }
    *n_out = n;
    if (n == 0) return PLF_OK;
    return plf_lbd_compute(o->lbd, host_img, w, h, stride, host_kl, n, host_desc, nullptr);
}
