// plf_line.cu -- host driver of the line path: replaces PL_SLAM::Lineextractor (LSD branch,
// src/Lineextractor.cc:32-67, :112-212) with LSDDetectorC::detect (LSDDetector_custom.cpp:56-73, :218-324)
// and BinaryDescriptor::compute (binary_descriptor_custom.cpp:350-398, :524-687, :1026-1372) beneath it.
#include "plf_line_kernels.cuh"
#include <math.h>
#include <vector>
#include <algorithm>
#include <mutex>
#include "plf_sort.cuh"      // the scan and the radix sort of the LSD pre-phase (own kernels)

#define LINE_MAX_OCT 2
#define LINE_DETCAP 4096       // detected lines per (frame, octave) before the response quota
#define LINE_REGCAP_PER_FRAME 4096

struct plf_line {
    plf_ctx* ctx;
    plf_line_params prm;
    float scale[16], inv_scale[16], sigma2[16], inv_sigma2[16];
    int per_level[16];
    // LSD constants
    double prec, rho;
    GaussQ8 lsd_gauss;      // pre-blur of cv::LineSegmentDetector when SCALE != 1
    GaussQ8 lbd_gauss;      // 5x5 sigma 1 (binary_descriptor_custom.cpp:358)
    LbdCoefs lbd_coefs;
    // workspace
    int ws_w, ws_h, ws_frames;
    int ow[LINE_MAX_OCT], oh[LINE_MAX_OCT], sw[LINE_MAX_OCT], sh[LINE_MAX_OCT], min_reg[LINE_MAX_OCT], kbits[LINE_MAX_OCT];
    int keybits[LINE_MAX_OCT];
    int sp[LINE_MAX_OCT];   // working width of the scaled octave: sw rounded up to 4 (row pitch of every per-pixel LSD array; the pad columns are NOTDEF)
    uint8_t* d_base;
    uint8_t *d_oct[LINE_MAX_OCT];                              // octave images (pitch == width)
    uint8_t *d_lbdimg[LINE_MAX_OCT];
    short2* d_dxy[LINE_MAX_OCT];            // Sobel (dx, dy) of the LBD octave images, interleaved
    // LSD workspace, one set per octave: the octaves of a batch are independent until the line selection, so they run
    // on two streams and their region-growing chains overlap
    uint8_t *d_tmp[LINE_MAX_OCT], *d_scaled[LINE_MAX_OCT];
    int2* d_comp[LINE_MAX_OCT];
    int *d_q[LINE_MAX_OCT], *d_label[LINE_MAX_OCT], *d_regpts[LINE_MAX_OCT], *d_cnt[LINE_MAX_OCT];
    unsigned* d_mask[LINE_MAX_OCT];       // one bit per scaled pixel: gradient defined
    int* d_offs[LINE_MAX_OCT];            // inclusive prefix sum of the mask popcounts (+ leading 0)
    double* d_bincoef[LINE_MAX_OCT];      // per frame
    float* d_fa[LINE_MAX_OCT];
    float2* d_cs[LINE_MAX_OCT];
    unsigned long long *d_keys[LINE_MAX_OCT], *d_keys2[LINE_MAX_OCT], *d_linekey[LINE_MAX_OCT];
    LsdRegion* d_regions[LINE_MAX_OCT];
    float4* d_lines[LINE_MAX_OCT];
    int *d_sctile[LINE_MAX_OCT], *d_rshist[LINE_MAX_OCT], *d_rstile[LINE_MAX_OCT], *d_rsbase[LINE_MAX_OCT];   // scan tile sums; radix: (tile, digit) table, frame of tile, first tile of frame
    size_t rs_tilecap[LINE_MAX_OCT];
    size_t keycap[LINE_MAX_OCT];
    int key_div;                            // seed-key capacity = scaled pixels / key_div (batches start at 4; halved on overflow, then the call is retried)
    size_t maskwords[LINE_MAX_OCT];
    int regcap;
    cudaStream_t st2;                      // stream of octave 1 (octave 0 uses the context stream)
    unsigned char* d_likely[LINE_MAX_OCT];  // per sorted seed: may start a region (k_lsd_likely)
    int2* d_gcscr[LINE_MAX_OCT];            // scratch of k_lsd_grow_cta (speculative regions)
    int gc_grid;                            // speculator warps the scratch is sized for
    unsigned long long* d_gcdbg;            // PLF_GC_DEBUG=1: counters of k_lsd_grow_cta (diagnosis only)
    cudaEvent_t ev_img, ev_join;
    // the three region-growing kernels of an octave (giant components, big components, the rest) work on disjoint components:
    // they run side by side on the octave's stream and two helpers
    cudaStream_t st_grow[LINE_MAX_OCT][2];
    cudaEvent_t ev_fork[LINE_MAX_OCT], ev_grow[LINE_MAX_OCT][2];
    int* h_pin;                            // pinned host staging for the per-octave counters (64 ints each)
    int *d_detcount;
    int qthr;               // defined <=> gx^2 + gy^2 > qthr
    plf_keyline* d_det;
    int2* d_tabs;
    const int2 *xtab[LINE_MAX_OCT], *ytab[LINE_MAX_OCT];
    // output staging for host entry points
    plf_keyline* d_okl;
    plf_keypoint* d_omid;
    uint8_t* d_odesc;
    float* d_ofdesc;
    int* d_onout;
    int out_frames, out_cap;
};

// d_cnt layout (ints): [0]=nkeys [1]=ncomp [2]=next [4]=sticky overflow flag [8..32) bucket counts [32..56) bucket fill [96..96+F) = maxq, then F region counts (one per frame)
enum { CNT_NKEYS = 0, CNT_NCOMP = 1, CNT_NEXT = 2, CNT_NREG = 3, CNT_ERR = 4, CNT_BCOUNT = 8, CNT_BFILL = 32, CNT_STATS = 64, CNT_MAXQ = 96 };

static int gauss_kernel_q8(int ksize, double sigma, int* q)
{
    if (ksize < 1 || !(ksize & 1) || ksize > 15) return -1;
    int n2 = ksize / 2;
    double w[16];
    if (sigma <= 0) sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    double scale2x = -0.125 / (sigma * sigma);
    double sum = 0;
    for (int i = 0, x = 1 - ksize; i < n2; i++, x += 2) {
        double t = exp((double)(x * x) * scale2x);
        w[i] = t;
        sum += t;
    }
    sum = sum * 2 + 1;
    sum = 1.0 / sum;
    double err = 0;
    long tot = 0;
    for (int i = 0; i < n2; i++) {
        double v = w[i] * sum * 256.0 + err;
        long v0 = lrint(v);
        err = v - (double)v0;
        q[i] = (int)v0;
        q[ksize - 1 - i] = (int)v0;
        tot += v0;
    }
    q[n2] = (int)(256 - 2 * tot);
    return 0;
}

extern "C" plf_status plf_line_create(plf_ctx* ctx, const plf_line_params* p, plf_line** out)
{
    if (!ctx || !p || !out) return PLF_ERR_INVALID;
    if (p->nlevels < 1 || p->nlevels > LINE_MAX_OCT || p->nfeatures < 1 || p->refine != 0 || !(p->scale > 0) ||
        p->n_bins < 2 || p->n_bins > 4096 || !(p->ang_th > 0) || !(p->ang_th < 180) || !(p->quant >= 0))
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_line_create: unsupported parameters (refine must be 0, nlevels 1..2, n_bins <= 4096)");
    plf_line* o = (plf_line*)calloc(1, sizeof(plf_line));
    o->ctx = ctx; o->prm = *p;
    const int n = p->nlevels;
    // Lineextractor::Lineextractor scale tables and feature split, src/Lineextractor.cc:36-66
    o->scale[0] = 1.0f; o->sigma2[0] = 1.0f;
    for (int i = 1; i < n; i++) { o->scale[i] = (float)((double)o->scale[i - 1] * p->scale); o->sigma2[i] = o->scale[i] * o->scale[i]; }
    for (int i = 0; i < n; i++) { o->inv_scale[i] = 1.0f / o->scale[i]; o->inv_sigma2[i] = 1.0f / o->sigma2[i]; }
    float factor = (float)(1.0f / p->scale);
    float nDesired = (float)(p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n)));
    int sum = 0;
    for (int l = 0; l < n - 1; l++) { o->per_level[l] = (int)lrintf(nDesired); sum += o->per_level[l]; nDesired *= factor; }
    o->per_level[n - 1] = p->nfeatures - sum > 0 ? p->nfeatures - sum : 0;
    // LSD constants (OpenCV lsd.cpp flsd)
    o->prec = LSD_PI * p->ang_th / 180;
    o->rho = p->quant / sin(o->prec);
    {   // largest integer q with sqrt(q / 4.0) <= rho in double arithmetic (the reference's NOTDEF test, monotone in q)
        long qq = (long)floor(4.0 * o->rho * o->rho);
        if (qq < 0) qq = 0;
        if (qq > 600000) qq = 600000;
        while (qq > 0 && !(sqrt((double)qq / 4.0) <= o->rho)) qq--;
        while (sqrt((double)(qq + 1) / 4.0) <= o->rho && qq < 600000) qq++;
        o->qthr = (sqrt((double)qq / 4.0) <= o->rho) ? (int)qq : -1;
    }
    if (p->scale != 1) {
        const double sigma = (p->scale < 1) ? (p->sigma_scale / p->scale) : p->sigma_scale;
        const unsigned hh = (unsigned)(ceil(sigma * sqrt(2 * 3.0 * log(10.0))));
        o->lsd_gauss.ksize = 1 + 2 * (int)hh;
        if (gauss_kernel_q8(o->lsd_gauss.ksize, sigma, o->lsd_gauss.q)) {
            free(o);
            return plf_fail(ctx, PLF_ERR_INVALID, "LSD pre-blur kernel larger than 15 taps (sigma_scale too large)");
        }
    }
    o->lbd_gauss.ksize = 5;
    gauss_kernel_q8(5, 1.0, o->lbd_gauss.q);
    {   // BinaryDescriptor ctor weights (binary_descriptor_custom.cpp:217-259), integer-division quirks kept
        double u = (7 * 3 - 1) / 2;
        double sigma = (7 * 2 + 1) / 2;
        double inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < 21; i++) { double d = i - u; o->lbd_coefs.l[i] = (float)exp(d * d * inv); }
        u = (9 * 7 - 1) / 2;
        sigma = u;
        inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < 63; i++) { double d = i - u; o->lbd_coefs.g[i] = (float)exp(d * d * inv); }
    }
    *out = o;
    return PLF_OK;
}

static void line_free_ws(plf_line* o)
{
    if (o->d_base) cudaFree(o->d_base);
    if (o->d_tabs) cudaFree(o->d_tabs);
    o->d_base = nullptr; o->d_tabs = nullptr;
    o->ws_w = o->ws_h = o->ws_frames = 0;
}

extern "C" void plf_line_destroy(plf_line* o)
{
    if (!o) return;
    cudaSetDevice(o->ctx->device);
    cudaStreamSynchronize(o->ctx->stream);
    line_free_ws(o);
    if (o->st2) cudaStreamDestroy(o->st2);
    if (o->ev_img) cudaEventDestroy(o->ev_img);
    if (o->ev_join) cudaEventDestroy(o->ev_join);
    for (int k = 0; k < LINE_MAX_OCT; k++) {
        if (o->ev_fork[k]) cudaEventDestroy(o->ev_fork[k]);
        for (int j = 0; j < 2; j++) {
            if (o->ev_grow[k][j]) cudaEventDestroy(o->ev_grow[k][j]);
            if (o->st_grow[k][j]) cudaStreamDestroy(o->st_grow[k][j]);
        }
    }
    if (o->h_pin) cudaFreeHost(o->h_pin);
    if (o->d_okl) cudaFree(o->d_okl);
    if (o->d_omid) cudaFree(o->d_omid);
    if (o->d_odesc) cudaFree(o->d_odesc);
    if (o->d_ofdesc) cudaFree(o->d_ofdesc);
    if (o->d_onout) cudaFree(o->d_onout);
    free(o);
}

extern "C" plf_status plf_line_tables(const plf_line* o, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* per_level)
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

extern "C" int plf_line_max_keylines(const plf_line* o) { return o ? o->prm.nfeatures + 16 : 0; }

static void exact_table(int ssize, int dsize, double inv_scale, int2* tab)
{
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; d++) {
        double f = (d + 0.5) * scale - 0.5;
        int s = (int)floor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0; }
        tab[d].x = s;
        tab[d].y = (int)floor(f * 256.0 + 0.5);
    }
}

template <class T> static T* carve(uint8_t*& p, size_t count)
{
    T* r = (T*)p;
    p += plf_align_up(count * sizeof(T), 256);
    return r;
}

static plf_status line_prepare(plf_line* o, int w, int h, int nframes)
{
    plf_ctx* ctx = o->ctx;
    if (o->ws_w == w && o->ws_h == h && o->ws_frames >= nframes) return PLF_OK;
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    line_free_ws(o);
    const int noct = o->prm.nlevels;
    const double S = o->prm.scale;
    size_t maxpx = 0, tabCount = 0;
    for (int k = 0; k < noct; k++) {
        o->ow[k] = w >> k; o->oh[k] = h >> k;
        if (o->ow[k] < 8 || o->oh[k] < 8) return plf_fail(ctx, PLF_ERR_INVALID, "image too small for %d octaves", noct);
        o->sw[k] = S != 1 ? (int)lrint(o->ow[k] * S) : o->ow[k];
        o->sh[k] = S != 1 ? (int)lrint(o->oh[k] * S) : o->oh[k];
        o->sp[k] = (o->sw[k] + 3) & ~3;
        if ((size_t)o->sp[k] * o->sh[k] > maxpx) maxpx = (size_t)o->sp[k] * o->sh[k];
        if ((size_t)o->sp[k] * o->sh[k] >= (1u << 22)) return plf_fail(ctx, PLF_ERR_INVALID, "scaled octave larger than 4M pixels");
        int kbv = 1, bbv = 1;
        while ((1u << kbv) < (unsigned)(o->sp[k] * o->sh[k])) kbv++;
        while ((1 << bbv) < o->prm.n_bins) bbv++;
        o->kbits[k] = kbv | (bbv << 8);     // packed key geometry (see plf_line_kernels.cuh)
        o->keybits[k] = 2 * kbv + bbv;      // key bits below the frame field
        if (64 - o->keybits[k] < 31 && nframes > (1 << (64 - o->keybits[k])))
            return plf_fail(ctx, PLF_ERR_INVALID, "at most %d frames of this size per line batch", 1 << (64 - o->keybits[k]));
        const double LOG_NT = 5 * (log10((double)o->sw[k]) + log10((double)o->sh[k])) / 2 + log10(11.0);
        o->min_reg[k] = (int)(size_t)(-LOG_NT / log10(o->prm.ang_th / 180));
        tabCount += (size_t)o->sw[k] + o->sh[k];
    }
    const size_t F = (size_t)nframes;
    // Seed keys, region points and component records exist only for pixels with a defined gradient (5-10 % of an image, 20 % on
    // busy ones): batches reserve a quarter of the pixel count (28 of the 49 bytes per scaled pixel are these arrays), which is
    // what lets 256 1080p frames share a workspace; an image set that needs more makes the call re-prepare with twice the room.
    if (o->key_div < 1) o->key_div = nframes >= 8 ? 4 : 1;
    if (nframes < 8) o->key_div = 1;
    o->regcap = nframes * LINE_REGCAP_PER_FRAME;
    if (!o->st2) {
#ifndef PLF_EMU
        int lo = 0, hi = 0, pr = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamGetPriority(ctx->stream, &pr);
        PLF_CUDA(ctx, cudaStreamCreateWithPriority(&o->st2, cudaStreamNonBlocking, pr));
#else
        PLF_CUDA(ctx, cudaStreamCreateWithFlags(&o->st2, cudaStreamNonBlocking));
#endif
        PLF_CUDA(ctx, cudaEventCreate(&o->ev_img));
        PLF_CUDA(ctx, cudaEventCreate(&o->ev_join));
        for (int k = 0; k < LINE_MAX_OCT; k++) {
            PLF_CUDA(ctx, cudaEventCreateWithFlags(&o->ev_fork[k], cudaEventDisableTiming));
            for (int j = 0; j < 2; j++) {
                PLF_CUDA(ctx, cudaEventCreateWithFlags(&o->ev_grow[k][j], cudaEventDisableTiming));
#ifndef PLF_EMU
                PLF_CUDA(ctx, cudaStreamCreateWithPriority(&o->st_grow[k][j], cudaStreamNonBlocking, pr));
#else
                PLF_CUDA(ctx, cudaStreamCreateWithFlags(&o->st_grow[k][j], cudaStreamNonBlocking));
#endif
            }
        }
        PLF_CUDA(ctx, cudaMallocHost((void**)&o->h_pin, LINE_MAX_OCT * 64 * sizeof(int)));
    }
    o->gc_grid = (int)((8 * F + 8) * (GC_MAXWARPS - 1) < 148 * 24 ? (8 * F + 8) * (GC_MAXWARPS - 1) : 148 * 24);   // most SPECULATOR WARPS of one k_lsd_grow_cta launch (its scratch is sized for them)
    size_t bytes = 0;
    auto need = [&](size_t count, size_t elt) { bytes += plf_align_up(count * elt, 256); };
    for (int k = 0; k < noct; k++) {
        const size_t px = (size_t)o->ow[k] * o->oh[k], spx = (size_t)o->sp[k] * o->sh[k];
        o->keycap[k] = F * spx / (size_t)o->key_div + 4096;
        o->maskwords[k] = F * (spx / 32 + (size_t)o->sh[k] + 64);
        need(F * px, 1); need(F * px, 1); need(F * px, 4);      // octave, LBD image, (dx, dy)
        need(F * px, 1); need(F * spx, 1);                                            // tmp, scaled
        need(F * spx, 4); need(F * spx, 4); need(F * spx, 4); need(F * spx, 8);   // q, label, fa, cs
        need(o->keycap[k], 8); need(o->keycap[k], 8); need(o->keycap[k], 4); need(o->keycap[k], 8);    // keys, keys2, regpts, comp
        need(o->regcap, sizeof(LsdRegion)); need(o->regcap, sizeof(float4));
        need(o->regcap, 8);
        need(CNT_MAXQ + 2 * F, 4);
        need(o->maskwords[k], 4); need(o->maskwords[k] + 64, 4); need(F, 8);   // mask, offsets, bin coefficients
        need(o->keycap[k], 1); need((size_t)o->gc_grid * GC_BUF * GC_RMAX, sizeof(int2));   // likely flags, speculation scratch
        o->rs_tilecap[k] = o->keycap[k] / RS_TILE + F + 1;
        need(o->maskwords[k] / SC_TILE + 2, 4); need(o->rs_tilecap[k] * 256, 4); need(o->rs_tilecap[k], 4); need(F + 1, 4);   // scan / sort tables
    }
    need(F * noct * LINE_DETCAP, sizeof(plf_keyline));
    need(F * noct, 4);

    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_base, bytes + 4096));
    uint8_t* p = o->d_base;
    for (int k = 0; k < noct; k++) {
        const size_t px = (size_t)o->ow[k] * o->oh[k], spx = (size_t)o->sp[k] * o->sh[k];
        o->d_oct[k] = carve<uint8_t>(p, F * px);
        o->d_lbdimg[k] = carve<uint8_t>(p, F * px);
        o->d_dxy[k] = carve<short2>(p, F * px);
        o->d_tmp[k] = carve<uint8_t>(p, F * px);
        o->d_scaled[k] = carve<uint8_t>(p, F * spx);
        o->d_q[k] = carve<int>(p, F * spx);
        o->d_label[k] = carve<int>(p, F * spx);
        o->d_fa[k] = carve<float>(p, F * spx);
        o->d_cs[k] = carve<float2>(p, F * spx);
        o->d_keys[k] = carve<unsigned long long>(p, o->keycap[k]);
        o->d_keys2[k] = carve<unsigned long long>(p, o->keycap[k]);
        o->d_regpts[k] = carve<int>(p, o->keycap[k]);
        o->d_comp[k] = carve<int2>(p, o->keycap[k]);
        o->d_regions[k] = carve<LsdRegion>(p, o->regcap);
        o->d_lines[k] = carve<float4>(p, o->regcap);
        o->d_linekey[k] = carve<unsigned long long>(p, o->regcap);
        o->d_cnt[k] = carve<int>(p, CNT_MAXQ + 2 * F);     // counters, maxq[F], regions per frame[F]
        o->d_mask[k] = carve<unsigned>(p, o->maskwords[k]);
        o->d_offs[k] = carve<int>(p, o->maskwords[k] + 64);
        o->d_bincoef[k] = carve<double>(p, F);
        o->d_likely[k] = carve<unsigned char>(p, o->keycap[k]);
        o->d_gcscr[k] = carve<int2>(p, (size_t)o->gc_grid * GC_BUF * GC_RMAX);
        o->d_sctile[k] = carve<int>(p, o->maskwords[k] / SC_TILE + 2);
        o->d_rshist[k] = carve<int>(p, o->rs_tilecap[k] * 256);
        o->d_rstile[k] = carve<int>(p, o->rs_tilecap[k]);
        o->d_rsbase[k] = carve<int>(p, F + 1);
    }
    o->d_det = carve<plf_keyline>(p, F * noct * LINE_DETCAP);
    o->d_detcount = carve<int>(p, F * noct);
    // INTER_LINEAR_EXACT tables
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_tabs, (tabCount + 1) * sizeof(int2)));
    if (S != 1) {
        std::vector<int2> tabs(tabCount + 1);
        size_t to = 0;
        for (int k = 0; k < noct; k++) {
            exact_table(o->ow[k], o->sw[k], S, &tabs[to]); o->xtab[k] = o->d_tabs + to; to += o->sw[k];
            exact_table(o->oh[k], o->sh[k], S, &tabs[to]); o->ytab[k] = o->d_tabs + to; to += o->sh[k];
        }
        PLF_CUDA(ctx, cudaMemcpy(o->d_tabs, tabs.data(), tabCount * sizeof(int2), cudaMemcpyHostToDevice));
    }
    for (int k = 0; k < noct; k++) PLF_CUDA(ctx, cudaMemset(o->d_offs[k], 0, sizeof(int)));
    o->ws_w = w; o->ws_h = h; o->ws_frames = nframes;
    return PLF_OK;
}

// k_lsd_keys emits the keys in (frame, raster index) order, so a STABLE sort of the (root, bin) bits above the raster index
// inside every frame's own range yields the fully sorted array (plf_sort.cuh).  The result ends up in d_keys2[k] (the two
// buffers trade places when the number of passes is even).
static plf_status sort_keys(plf_line* o, int k, int n, int nframes, int wpf, cudaStream_t st)
{
    plf_ctx* ctx = o->ctx;
    const int KB = o->kbits[k] & 0xff, BB = o->kbits[k] >> 8;
    const int passes = (KB + BB + 7) / 8;
    const int ntiles = n / RS_TILE + nframes;       // upper bound of the tiles in use (a frame's last tile may be partial)
    PLF_LAUNCH(k_rs_frames, dim3(1), dim3(SC_T), 0, st, (const int*)o->d_offs[k], wpf, nframes, o->d_rsbase[k], o->d_rstile[k], (int)o->rs_tilecap[k]);
    PLF_CHECK_LAUNCH(ctx);
    unsigned long long *in = o->d_keys[k], *out = o->d_keys2[k];
    for (int p = 0; p < passes; p++) {
        const int shift = KB + 8 * p;
        PLF_LAUNCH(k_rs_hist, dim3(ntiles), dim3(RS_T), 0, st, (const unsigned long long*)in, (const int*)o->d_offs[k], wpf, nframes, (const int*)o->d_rsbase[k],
                   (const int*)o->d_rstile[k], shift, o->d_rshist[k]);
        PLF_CHECK_LAUNCH(ctx);
        PLF_LAUNCH(k_rs_offsets, dim3(nframes), dim3(256), 0, st, o->d_rshist[k], (const int*)o->d_offs[k], wpf, (const int*)o->d_rsbase[k]);
        PLF_CHECK_LAUNCH(ctx);
        PLF_LAUNCH(k_rs_scatter, dim3(ntiles), dim3(RS_T), 0, st, (const unsigned long long*)in, out, (const int*)o->d_offs[k], wpf, nframes,
                   (const int*)o->d_rsbase[k], (const int*)o->d_rstile[k], shift, (const int*)o->d_rshist[k]);
        PLF_CHECK_LAUNCH(ctx);
        unsigned long long* t = in; in = out; out = t;
    }
    if (in != o->d_keys2[k]) { o->d_keys[k] = o->d_keys2[k]; o->d_keys2[k] = in; }      // `in` holds the sorted keys
    return PLF_OK;
}

// small device -> host reads go through pinned staging: a copy into pageable memory makes the driver wait for the stream
// inside the call (holding its lock), which stalls the launches of every other context's host thread
static plf_status read_ints(plf_ctx* ctx, const int* dev, int n, int* host)
{
    void* pin;
    plf_status st = plf_ctx_pinned(ctx, (size_t)(n > 64 ? n : 64) * sizeof(int), &pin);
    if (st) return st;
    PLF_CUDA(ctx, cudaMemcpyAsync(pin, dev, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    { plf_status ss = plf_sync(ctx, ctx->stream); if (ss) return ss; }
    memcpy(host, pin, (size_t)n * sizeof(int));
    return PLF_OK;
}

// interior output strips of pyrDown for a source of width w: strip s reads source bytes 8s - 4 .. 8s + 11
static int pyrdown_interior(int w)
{
    const int last = (w - 12) / 8;     // 8s + 12 <= w
    return last >= 1 ? last : 0;       // first interior strip is 1 (8s - 4 >= 0)
}

// cv::GaussianBlur 8U for a batch: 5 / 7 taps go to the register sliding-window kernel, other sizes to the generic one
static plf_status gauss_batch(plf_ctx* ctx, cudaStream_t st, const uint8_t* src, uint8_t* dst, int w, int h, int nframes, const GaussQ8& k)
{
    const size_t frame = (size_t)w * h;
    if (k.ksize == 5 || k.ksize == 7) {
        BlurTaps taps;
        memset(&taps, 0, sizeof(taps));
        for (int i = 0; i < k.ksize; i++) taps.k[i] = k.q[i];
        const int F = plf_strip_interior(w, 4, 8), ncx_int = plf_div_up(F, 32);   // interior strips / their CTA columns (+1 edge column)
        dim3 grid(ncx_int + 1, plf_div_up(h, 4 * GS_ROWS), nframes), block(32, 4);
        if (k.ksize == 5) PLF_LAUNCH(k_gauss_strip<2>, grid, block, 0, st, src, frame, w, dst, frame, w, w, h, taps, F, ncx_int);
        else PLF_LAUNCH(k_gauss_strip<3>, grid, block, 0, st, src, frame, w, dst, frame, w, w, h, taps, F, ncx_int);
    } else {
        PLF_LAUNCH(k_gauss_q8, dim3(plf_div_up(w, GB_TW), plf_div_up(h, GB_TH), nframes), dim3(256), 0, st, src, frame, w, dst, frame, w, w, h, k);
    }
    PLF_CHECK_LAUNCH(ctx);
    return PLF_OK;
}

// Several line contexts (host threads) may run at once.  The phase of an octave before region growing is
// bandwidth-bound and fills the GPU; region growing is a latency-bound dependent chain that leaves it mostly
// idle.  Contexts therefore take turns for the first phase (this mutex) and overlap their growing phases with
// the other contexts' first phases instead of marching in lockstep.
static std::mutex g_lsd_prephase[64];   // one per device: contexts of different GPUs in one process never wait for each other

// LSDDetectorC::detect for a batch resident in d_oct[0]: fills d_det / d_detcount.
// The octaves are independent until the line selection: each has its own workspace and (when there are two) its own
// stream, the host walks both through the same phases, so the two region-growing chains run side by side.
// With profiling on, everything stays on the context stream (the per-kernel event times must not overlap).
#define PLF_RETRY_INTERNAL ((plf_status)1000)   // never leaves this file
// host-side phase trace of a call (PLF_TRACE=1): microseconds since the start of lsd_detect_batch at each phase boundary
#include <chrono>
struct PhaseTrace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    PhaseTrace() : on(getenv("PLF_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what)
    {
        if (on) fprintf(stderr, "[trace] %-28s %8.1f us\n", what, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    }
};

static plf_status lsd_detect_batch(plf_line* o, int nframes)
{
    plf_ctx* ctx = o->ctx;
    PhaseTrace tr;
    const int noct = o->prm.nlevels;
    const double S = o->prm.scale;
    cudaStream_t st0 = ctx->stream;
    cudaStream_t stk[LINE_MAX_OCT];
    for (int k = 0; k < noct; k++) stk[k] = (k == 0 || ctx->prof_on) ? st0 : o->st2;
    PLF_CUDA(ctx, cudaMemsetAsync(o->d_detcount, 0, (size_t)nframes * noct * sizeof(int), st0));
    // whatever is still queued on this stream (the image upload of a host-buffer call) finishes before the turn is
    // taken: the lock must not be held across a PCIe transfer
    ctx->blocking = nframes >= 32;      // batches sleep in their waits, single frames spin (latency)
    { plf_status ss = plf_sync(ctx, st0); if (ss) return ss; }
    tr.mark("upload done");
    std::unique_lock<std::mutex> prephase(g_lsd_prephase[ctx->device & 63], std::defer_lock);
    if (!getenv("PLF_NO_PREPHASE_LOCK")) prephase.lock();
    // computeGaussianPyramid: pyrDown, no pre-blur (LSDDetector_custom.cpp:56-73)
    for (int k = 1; k < noct; k++) {
        const int pdF = pyrdown_interior(o->ow[k - 1]);
        PLF_LAUNCH(k_pyrdown, dim3(plf_div_up(pdF, 32) + 1, plf_div_up(o->oh[k], 4 * PD_ROWS), nframes), dim3(32, 4), 0, st0, (const uint8_t*)o->d_oct[k - 1],
                   (size_t)o->ow[k - 1] * o->oh[k - 1], o->ow[k - 1], o->ow[k - 1], o->oh[k - 1], o->d_oct[k], (size_t)o->ow[k] * o->oh[k], o->ow[k], pdF,
                   plf_div_up(pdF, 32));
        PLF_CHECK_LAUNCH(ctx);
    }
    if (noct > 1 && stk[1] != st0) {
        PLF_CUDA(ctx, cudaEventRecord(o->ev_img, st0));
        PLF_CUDA(ctx, cudaStreamWaitEvent(stk[1], o->ev_img, 0));
    }
    int nwords[LINE_MAX_OCT], nkeys[LINE_MAX_OCT], mwk[LINE_MAX_OCT];
    // ---- phase 1: scaled image, gradient, mask, components, keys ----
    for (int k = 0; k < noct; k++) {
        const int ow = o->ow[k], oh = o->oh[k], sw = o->sw[k], sh = o->sh[k], sp = o->sp[k];
        cudaStream_t st = stk[k];
        // the image the gradient is taken of: the octave itself (SCALE == 1, pitch = octave width) or its blurred and
        // resized copy (pitch = sp)
        const uint8_t* scaled = o->d_oct[k];
        int spitch = ow;
        size_t sframe = (size_t)ow * oh;
        if (S != 1) {
            { plf_status gs = gauss_batch(ctx, st, o->d_oct[k], o->d_tmp[k], ow, oh, nframes, o->lsd_gauss); if (gs) return gs; }
            PLF_LAUNCH(k_resize_exact, dim3(plf_div_up(sw, 128), plf_div_up(sh, 8 * RX_ROWS), nframes), dim3(32, 8), 0, st, (const uint8_t*)o->d_tmp[k],
                       (size_t)ow * oh, ow, ow, oh, o->d_scaled[k], (size_t)sp * sh, sp, sw, sh, o->xtab[k], o->ytab[k]);
            PLF_CHECK_LAUNCH(ctx);
            scaled = o->d_scaled[k]; spitch = sp; sframe = (size_t)sp * sh;
        }
        PLF_CUDA(ctx, cudaMemsetAsync(o->d_cnt[k], 0, 8 * sizeof(int), st));
        PLF_CUDA(ctx, cudaMemsetAsync(o->d_cnt[k] + CNT_BCOUNT, 0, (CNT_MAXQ - CNT_BCOUNT) * sizeof(int), st));
        PLF_CUDA(ctx, cudaMemsetAsync(o->d_cnt[k] + CNT_MAXQ, 0xff, (size_t)nframes * sizeof(int), st));   // maxq = -1
        PLF_CUDA(ctx, cudaMemsetAsync(o->d_cnt[k] + CNT_MAXQ + nframes, 0, (size_t)nframes * sizeof(int), st));   // regions per frame
        // from here on every per-pixel array has sp columns per row (sw real ones + NOTDEF padding)
        const int mw = plf_div_up(sp, 32);
        mwk[k] = mw;
        dim3 g2(plf_div_up(sh, 8), 1, nframes), b2(32, 8);      // one warp per image row (ccl, keys)
        PLF_LAUNCH(k_lsd_grad, dim3(plf_div_up(sp, 128), plf_div_up(sh, 4 * GRAD_ROWS), nframes), dim3(32, 4), 0, st, scaled, sframe, spitch, sw, sp, sh, o->qthr,
                   o->d_q[k], o->d_fa[k], o->d_label[k], o->d_mask[k], mw, o->d_cnt[k] + CNT_MAXQ);
        PLF_CHECK_LAUNCH(ctx);
        PLF_LAUNCH(k_lsd_bincoef, dim3(plf_div_up(nframes, 128)), dim3(128), 0, st, (const int*)(o->d_cnt[k] + CNT_MAXQ), nframes, o->prm.n_bins,
                   o->d_bincoef[k]);
        PLF_CHECK_LAUNCH(ctx);
        PLF_LAUNCH(k_ccl_merge, g2, b2, 0, st, o->d_label[k], (const unsigned*)o->d_mask[k], mw, sp, sh);
        PLF_CHECK_LAUNCH(ctx);
        // key positions: inclusive scan of the mask popcounts (d_offs[0] = 0 is set once per workspace)
        nwords[k] = nframes * sh * mw;
        {
            const int nt = plf_div_up(nwords[k], SC_TILE);
            PLF_LAUNCH(k_scan_tile_sums, dim3(nt), dim3(SC_T), 0, st, (const unsigned*)o->d_mask[k], nwords[k], o->d_sctile[k]);
            PLF_CHECK_LAUNCH(ctx);
            PLF_LAUNCH(k_scan_top, dim3(1), dim3(SC_T), 0, st, o->d_sctile[k], nt);
            PLF_CHECK_LAUNCH(ctx);
            PLF_LAUNCH(k_scan_apply, dim3(nt), dim3(SC_T), 0, st, (const unsigned*)o->d_mask[k], nwords[k], (const int*)o->d_sctile[k], o->d_offs[k] + 1);
            PLF_CHECK_LAUNCH(ctx);
        }
        PLF_LAUNCH(k_lsd_keys, g2, b2, 0, st, o->d_label[k], (const int*)o->d_q[k], (const unsigned*)o->d_mask[k], mw, (const int*)o->d_offs[k],
                   (const double*)o->d_bincoef[k], sp, sh, o->prm.n_bins, o->d_keys[k], (int)o->keycap[k], o->kbits[k]);
        PLF_CHECK_LAUNCH(ctx);
        PLF_CUDA(ctx, cudaMemcpyAsync(o->h_pin + 64 * k, o->d_offs[k] + nwords[k], sizeof(int), cudaMemcpyDeviceToHost, st));   // pinned staging
    }
    // ---- phase 2: sort, component heads, sorted positions ----
    tr.mark("phase 1 queued");
    for (int k = 0; k < noct; k++) {
        cudaStream_t st = stk[k];
        { plf_status ss = plf_sync(ctx, st); if (ss) return ss; }
        tr.mark("phase 1 done (octave)");
        nkeys[k] = o->h_pin[64 * k];
        if (nkeys[k] > (int)o->keycap[k]) {
            // more defined pixels than the workspace reserves: finish what is queued, then let the caller re-prepare and retry
            for (int j = 0; j < noct; j++) cudaStreamSynchronize(stk[j]);
            if (o->key_div <= 1) return plf_fail(ctx, PLF_ERR_CAPACITY, "LSD key buffer overflow");
            o->key_div /= 2;
            o->ws_frames = 0;
            return PLF_RETRY_INTERNAL;
        }
        if (nkeys[k] <= 0) continue;
        plf_status s = sort_keys(o, k, nkeys[k], nframes, o->sh[k] * mwk[k], st);
        if (s) return s;
        for (int pass = 0; pass < 2; pass++) {
            PLF_LAUNCH(k_lsd_heads, dim3(plf_div_up(nkeys[k], 256)), dim3(256), 0, st, (const unsigned long long*)o->d_keys2[k], nkeys[k], o->d_comp[k],
                       o->d_cnt[k] + CNT_BCOUNT, o->d_cnt[k] + CNT_BFILL, o->d_cnt[k] + CNT_NCOMP, pass, o->kbits[k]);
            PLF_CHECK_LAUNCH(ctx);
        }
        // the sorted position of every defined pixel (compact component index) and its cos / sin
        PLF_LAUNCH(k_lsd_cid, dim3(plf_div_up(nkeys[k], 256)), dim3(256), 0, st, (const unsigned long long*)o->d_keys2[k], nkeys[k], o->d_label[k],
                   (const float*)o->d_fa[k], o->d_cs[k], (size_t)o->sp[k] * o->sh[k], o->kbits[k]);
        PLF_CHECK_LAUNCH(ctx);
        PLF_CUDA(ctx, cudaMemcpyAsync(o->h_pin + 64 * k + 8, o->d_cnt[k] + CNT_BCOUNT, LSD_NBUCKET * sizeof(int), cudaMemcpyDeviceToHost, st));
        PLF_CUDA(ctx, cudaMemcpyAsync(o->h_pin + 64 * k + 1, o->d_cnt[k] + CNT_NCOMP, sizeof(int), cudaMemcpyDeviceToHost, st));   // largest component
    }
    tr.mark("phase 2 queued");
    for (int k = 0; k < noct; k++)
        if (nkeys[k] > 0) { plf_status ss = plf_sync(ctx, stk[k]); if (ss) return ss; }
    tr.mark("phase 2 done");
    if (prephase.owns_lock()) prephase.unlock();   // everything up to here has finished on the GPU; growing may overlap other contexts
    // ---- phase 3: region growing (the latency-bound chains of both octaves side by side), rectangles, keylines ----
    for (int k = 0; k < noct; k++) {
        cudaStream_t st = stk[k];
        const int sp = o->sp[k], sh = o->sh[k];
        if (nkeys[k] > 0) {
            // the big components with warp-cooperative ordered growth (one warp each), everything else with one thread per
            // component; the used-bitmap of the warp kernel is sized from the largest component present (bucket counts)
            const int* bc = o->h_pin + 64 * k + 8;
            int topb = 0, nbig = 0;
            for (int b = 0; b < LSD_NBUCKET; b++) { if (bc[b]) topb = b; if (b >= LSD_BIG_BUCKET) nbig += bc[b]; }
            (void)topb;
            int wg_maxc = (o->h_pin[64 * k + 1] + 1023) & ~1023;     // used-bitmap bits per warp: the largest component, rounded up
            if (wg_maxc > WARPGROW_MAXC) wg_maxc = WARPGROW_MAXC;
            if (wg_maxc < 1024) wg_maxc = 1024;
            const int wg_smem = WG_WARPS * (wg_maxc / 8);
            // giant components: one CTA each, in-order commit with speculative growth ahead of it (plf_lsd_grow_cta.cuh)
            // Which components get a CTA (speculators + committer), and how many speculators each: few components in the launch
            // -> everything from 1024 seeds up with as many speculators as the bitmaps allow (latency); many -> only the largest
            // size classes and fewer speculators, so that every CTA is resident at once and other kernels still find shared
            // memory (throughput).  The rest goes to k_lsd_grow_warp.
            const int bm = wg_maxc / 8;
            // CTAs of g warps that fit at once: by shared memory (160 KB per SM, the rest stays for other kernels), by warps (24 per SM)
            // and by the speculation scratch
            auto resident = [&](int g) {
                int per_sm = (int)((160 * 1024) / ((size_t)g * bm + sizeof(GcShared)));
                if (per_sm > 24 / g) per_sm = 24 / g;
                const int by_scratch = o->gc_grid / (g - 1);
                return 148 * per_sm < by_scratch ? 148 * per_sm : by_scratch;
            };
            int gthr = LSD_GIANT_BUCKET, ngiant = 0;
            for (;; gthr += 2) {
                ngiant = 0;
                for (int b = gthr; b < LSD_NBUCKET; b++) ngiant += bc[b];
                if (ngiant <= resident(4) || gthr + 2 >= LSD_NBUCKET) break;
            }
            // Speculation buys latency with extra work and shared memory, so it is used where the longest chain of the launch
            // (largest component x ~0.37 us per pixel with one warp) outweighs the launch's bandwidth-bound work (~34 ns per 1000
            // scaled pixels per frame, from the round-2 profiles): a single frame always qualifies, 1080p x 128 frames per call does
            // (ratio 4.3, measured 15-25 % faster with it), 752x480 x 512 frames per call does not (ratio 1.4-2.7, measured 20 % slower with it).
            {
                const double chain_us = 0.37 * (double)o->h_pin[64 * k + 1];
                const double parallel_us = 3.4e-5 * (double)nframes * (double)sp * (double)sh;
                if (ngiant > resident(4) || chain_us < 4.0 * parallel_us || getenv("PLF_GC_OFF")) { ngiant = 0; gthr = LSD_NBUCKET; }
            }
#ifndef PLF_EMU
            if (!o->d_gcdbg && getenv("PLF_GC_DEBUG")) { PLF_CUDA(ctx, cudaMalloc((void**)&o->d_gcdbg, 16 * 8)); PLF_CUDA(ctx, cudaMemset(o->d_gcdbg, 0, 16 * 8)); }
#endif
            // fork: the warp kernel and the thread kernel do not wait for the giant components' CTAs (disjoint components; regions
            // are appended with atomics and ordered later by their keys).  With the per-launch profiler on everything stays on one stream.
            const bool forked = !ctx->prof_on;
            cudaStream_t st_w = forked ? o->st_grow[k][0] : st, st_t = forked ? o->st_grow[k][1] : st;
            if (forked) {
                PLF_CUDA(ctx, cudaEventRecord(o->ev_fork[k], st));
                PLF_CUDA(ctx, cudaStreamWaitEvent(st_w, o->ev_fork[k], 0));
                PLF_CUDA(ctx, cudaStreamWaitEvent(st_t, o->ev_fork[k], 0));
            }
            if (ngiant > 0) {
                int gc_warps = GC_SMEM_BUDGET / bm;
                if (gc_warps > GC_MAXWARPS) gc_warps = GC_MAXWARPS;
                while (gc_warps > 2 && ngiant > resident(gc_warps)) gc_warps--;
                if (getenv("PLF_GC_WARPS") && atoi(getenv("PLF_GC_WARPS")) < gc_warps) gc_warps = atoi(getenv("PLF_GC_WARPS"));   // diagnosis
                if (gc_warps < 2) gc_warps = 2;
                // which seeds are worth growing ahead of their turn
                PLF_LAUNCH(k_lsd_likely, dim3(plf_div_up(nkeys[k], 256)), dim3(256), 0, st, (const unsigned long long*)o->d_keys2[k], nkeys[k], (const int*)o->d_label[k],
                           (const float*)o->d_fa[k], o->sp[k], o->sh[k], o->prec, o->d_likely[k], o->kbits[k]);
                PLF_CHECK_LAUNCH(ctx);
                PLF_SMEM_OPTIN(ctx, k_lsd_grow_cta);
                PLF_LAUNCH(k_lsd_grow_cta, dim3(ngiant), dim3(32 * gc_warps), (size_t)gc_warps * (wg_maxc / 8), st,
                           (const unsigned long long*)o->d_keys2[k], (const int2*)o->d_comp[k], (const int*)(o->d_cnt[k] + CNT_BCOUNT),
                           (const unsigned char*)o->d_likely[k], (const float*)o->d_fa[k], (const float2*)o->d_cs[k], (const int*)o->d_label[k], sp, sh,
                           o->prec, o->min_reg[k], o->d_regpts[k], o->d_regions[k], o->d_cnt[k] + CNT_MAXQ + nframes, LINE_REGCAP_PER_FRAME, o->kbits[k],
                           wg_maxc, o->d_gcscr[k], gthr, o->d_gcdbg);
                PLF_CHECK_LAUNCH(ctx);
            }
            PLF_SMEM_OPTIN(ctx, k_lsd_grow_warp);
            const int nmid = nbig - ngiant;
            const int wg_ctas = plf_div_up(nmid, WG_WARPS) < 148 * 4 ? plf_div_up(nmid, WG_WARPS) : 148 * 4;
            if (nmid > 0) PLF_LAUNCH(k_lsd_grow_warp, dim3(wg_ctas), dim3(32 * WG_WARPS), wg_smem, st_w, (const unsigned long long*)o->d_keys2[k],
                       (const int2*)o->d_comp[k], (const int*)(o->d_cnt[k] + CNT_BCOUNT), (const float*)o->d_fa[k], (const float2*)o->d_cs[k],
                       (const int*)o->d_label[k], sp, sh, o->prec, o->min_reg[k], o->d_regpts[k], o->d_regions[k], o->d_cnt[k] + CNT_MAXQ + nframes, LINE_REGCAP_PER_FRAME, o->kbits[k], wg_maxc, gthr);
            PLF_CHECK_LAUNCH(ctx);
            PLF_LAUNCH(k_lsd_grow, dim3(148 * 4), dim3(128), 0, st_t, (const unsigned long long*)o->d_keys2[k], nkeys[k], (const int2*)o->d_comp[k],
                       (const int*)(o->d_cnt[k] + CNT_BCOUNT), o->d_cnt[k] + CNT_NEXT, o->d_fa[k], (const float2*)o->d_cs[k], sp, sh, o->prec,
                       o->min_reg[k], o->d_regpts[k], o->d_regions[k], o->d_cnt[k] + CNT_MAXQ + nframes, LINE_REGCAP_PER_FRAME, 1, wg_maxc, o->kbits[k]);
            PLF_CHECK_LAUNCH(ctx);
            if (forked) {      // join before the rectangles
                PLF_CUDA(ctx, cudaEventRecord(o->ev_grow[k][0], st_w));
                PLF_CUDA(ctx, cudaEventRecord(o->ev_grow[k][1], st_t));
                PLF_CUDA(ctx, cudaStreamWaitEvent(st, o->ev_grow[k][0], 0));
                PLF_CUDA(ctx, cudaStreamWaitEvent(st, o->ev_grow[k][1], 0));
            }
        }
        // enough CTAs per frame to fill the GPU for small batches
        int rsplit = plf_div_up(148 * 8, nframes);
        if (rsplit < 8) rsplit = 8;      // region sizes vary a lot: shorter per-warp loops even out the tails
        if (rsplit > LINE_REGCAP_PER_FRAME / RECT_WARPS) rsplit = LINE_REGCAP_PER_FRAME / RECT_WARPS;
        PLF_LAUNCH(k_lsd_rect, dim3(rsplit, nframes), dim3(32 * RECT_WARPS), 0, st, (const LsdRegion*)o->d_regions[k],
                   (const int*)(o->d_cnt[k] + CNT_MAXQ + nframes), LINE_REGCAP_PER_FRAME, (const int*)o->d_regpts[k], (const int*)o->d_q[k], sp, sh,
                   o->prec, S, o->d_lines[k], o->d_linekey[k], o->d_cnt[k] + CNT_ERR, o->kbits[k]);
        PLF_CHECK_LAUNCH(ctx);
        PLF_LAUNCH(k_lsd_keylines, dim3(nframes), dim3(KL_T), 0, st, (const unsigned long long*)o->d_linekey[k],
                   (const int*)(o->d_cnt[k] + CNT_MAXQ + nframes), LINE_REGCAP_PER_FRAME, (const float4*)o->d_lines[k], k, noct, o->ow[k], o->oh[k],
                   o->prm.min_line_length, o->d_det, o->d_detcount, LINE_DETCAP);
        PLF_CHECK_LAUNCH(ctx);
    }
#ifndef PLF_EMU
    if (o->d_gcdbg) {
        unsigned long long hd[16];
        cudaDeviceSynchronize();
        cudaMemcpy(hd, o->d_gcdbg, sizeof hd, cudaMemcpyDeviceToHost);
        cudaMemset(o->d_gcdbg, 0, sizeof hd);
        fprintf(stderr, "[gc] frames %d: spec-consumed %llu regions / %llu px, failed validations %llu, committer-grown %llu regions / %llu px, wait polls %llu, speculated %llu regions / %llu px, aborted %llu\n",
                nframes, hd[0], hd[1], hd[2], hd[3], hd[4], hd[5], hd[6], hd[7], hd[8]);
    }
#endif
    tr.mark("phase 3 queued");
    if (tr.on) { for (int k = 0; k < noct; k++) cudaStreamSynchronize(stk[k]); tr.mark("phase 3 done (trace only sync)"); }
    // the selection / LBD that follow run on the context stream: join octave 1
    if (noct > 1 && stk[1] != st0) {
        PLF_CUDA(ctx, cudaEventRecord(o->ev_join, stk[1]));
        PLF_CUDA(ctx, cudaStreamWaitEvent(st0, o->ev_join, 0));
    }
    return PLF_OK;
}

// selection (or plain min_length filter) into out_kl / out_mid / n_out (device)
static plf_status line_select(plf_line* o, int nframes, bool select, plf_keyline* d_kl, plf_keypoint* d_mid, int cap, int* d_nout)
{
    plf_ctx* ctx = o->ctx;
    const int noct = o->prm.nlevels;
    PLF_LAUNCH(k_line_select, dim3(nframes), dim3(SEL_T), SEL_SMEM_BYTES(LINE_DETCAP), ctx->stream, (const plf_keyline*)o->d_det,
               (const int*)o->d_detcount, LINE_DETCAP, noct, select ? 1 : 0, o->per_level[0], noct > 1 ? o->per_level[1] : 0, d_kl, d_mid,
               cap, d_nout, (const int*)(o->d_cnt[0] + CNT_ERR), noct > 1 ? (const int*)(o->d_cnt[1] + CNT_ERR) : (const int*)nullptr);
    PLF_CHECK_LAUNCH(ctx);
    return PLF_OK;
}

// BinaryDescriptor::compute for keylines resident on the device; image batch in d_oct[0]
static plf_status lbd_batch(plf_line* o, int nframes, const plf_keyline* d_kl, const int* d_nlines, int cap, int maxlines,
                            uint8_t* d_desc, float* d_fdesc)
{
    plf_ctx* ctx = o->ctx;
    cudaStream_t st = ctx->stream;
    const int noct = o->prm.nlevels;
    LbdImages im;
    memset(&im, 0, sizeof(im));
    for (int k = 0; k < noct; k++) {
        const int ow = o->ow[k], oh = o->oh[k];
        if (k == 0) {   // computeGaussianPyramid (binary_descriptor_custom.cpp:350-370): blur 5x5 sigma 1, then pyrDown
            { plf_status gs = gauss_batch(ctx, st, o->d_oct[0], o->d_lbdimg[0], ow, oh, nframes, o->lbd_gauss); if (gs) return gs; }
        } else {
            const int pdF = pyrdown_interior(o->ow[k - 1]);
            PLF_LAUNCH(k_pyrdown, dim3(plf_div_up(pdF, 32) + 1, plf_div_up(oh, 4 * PD_ROWS), nframes), dim3(32, 4), 0, st, (const uint8_t*)o->d_lbdimg[k - 1],
                       (size_t)o->ow[k - 1] * o->oh[k - 1], o->ow[k - 1], o->ow[k - 1], o->oh[k - 1], o->d_lbdimg[k], (size_t)ow * oh, ow, pdF, plf_div_up(pdF, 32));
        }
        PLF_CHECK_LAUNCH(ctx);
        const int sbF = plf_strip_interior(ow, 4, 8);
        PLF_LAUNCH(k_sobel3, dim3(plf_div_up(sbF, 32) + 1, plf_div_up(oh, 4 * SB_ROWS), nframes), dim3(32, 4), 0, st, (const uint8_t*)o->d_lbdimg[k],
                   (size_t)ow * oh, ow, ow, oh, o->d_dxy[k], (size_t)ow * oh, sbF, plf_div_up(sbF, 32));
        PLF_CHECK_LAUNCH(ctx);
        im.dxy[k] = o->d_dxy[k]; im.frame[k] = (size_t)ow * oh; im.w[k] = ow; im.h[k] = oh;
    }
    if (maxlines > 0) {
        {
            PLF_LAUNCH(k_lbd, dim3(maxlines, nframes), dim3(64), 0, st, d_kl, d_nlines, cap, im, o->lbd_coefs, d_desc, d_fdesc);
            PLF_CHECK_LAUNCH(ctx);
        }
    }
    return PLF_OK;
}

static plf_status line_out_staging(plf_line* o, int nframes, int cap)
{
    plf_ctx* ctx = o->ctx;
    if (o->out_frames >= nframes && o->out_cap == cap) return PLF_OK;
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (o->d_okl) cudaFree(o->d_okl);
    if (o->d_omid) cudaFree(o->d_omid);
    if (o->d_odesc) cudaFree(o->d_odesc);
    if (o->d_ofdesc) cudaFree(o->d_ofdesc);
    if (o->d_onout) cudaFree(o->d_onout);
    o->d_okl = nullptr; o->d_omid = nullptr; o->d_odesc = nullptr; o->d_ofdesc = nullptr; o->d_onout = nullptr; o->out_frames = 0;
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_okl, (size_t)nframes * cap * sizeof(plf_keyline)));
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_omid, (size_t)nframes * cap * sizeof(plf_keypoint)));
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_odesc, (size_t)nframes * cap * 32));
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_ofdesc, (size_t)nframes * cap * 72 * sizeof(float)));
    PLF_CUDA(ctx, cudaMalloc((void**)&o->d_onout, (size_t)nframes * sizeof(int)));
    o->out_frames = nframes; o->out_cap = cap;
    return PLF_OK;
}

static plf_status upload_images(plf_line* o, const uint8_t* host_imgs, int nframes, int w, int h, size_t stride, size_t frame_stride, bool device_src)
{
    plf_ctx* ctx = o->ctx;
    const cudaMemcpyKind kind = device_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (stride == (size_t)w && frame_stride == (size_t)w * h) {
        PLF_CUDA(ctx, cudaMemcpyAsync(o->d_oct[0], host_imgs, (size_t)nframes * frame_stride, kind, ctx->stream));
    } else {
        for (int f = 0; f < nframes; f++)
            PLF_CUDA(ctx, cudaMemcpy2DAsync(o->d_oct[0] + (size_t)f * w * h, w, host_imgs + (size_t)f * frame_stride, stride, w, h, kind, ctx->stream));
    }
    return PLF_OK;
}

static plf_status line_extract_device_impl(plf_line* o, int nframes, plf_keyline* d_kl, plf_keypoint* d_mid, uint8_t* d_desc, int cap, int* d_nout)
{
    plf_status st = lsd_detect_batch(o, nframes);
    if (st) return st;
    st = line_select(o, nframes, true, d_kl, d_mid, cap, d_nout);
    if (st) return st;
    int maxlines = o->prm.nfeatures < cap ? o->prm.nfeatures : cap;
    return lbd_batch(o, nframes, d_kl, d_nout, cap, maxlines, d_desc, nullptr);
}

extern "C" plf_status plf_line_extract_batch_device(plf_line* o, const uint8_t* dev_imgs, int nframes, int w, int h, size_t stride,
                                                    size_t frame_stride, plf_keyline* dev_kl, plf_keypoint* dev_mid, uint8_t* dev_desc,
                                                    int cap, int32_t* dev_n_out)
{
    if (!o) return PLF_ERR_INVALID;
    plf_ctx* ctx = o->ctx;
    if (!dev_imgs || nframes < 1 || w <= 0 || h <= 0 || stride < (size_t)w || !dev_kl || !dev_desc || !dev_n_out || cap < 1)
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_line_extract_batch_device: bad arguments");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    plf_status st;
    do {
        st = line_prepare(o, w, h, nframes);
        if (st) return st;
        st = upload_images(o, dev_imgs, nframes, w, h, stride, frame_stride, true);
        if (st) return st;
        st = line_extract_device_impl(o, nframes, dev_kl, dev_mid, dev_desc, cap, dev_n_out);
    } while (st == PLF_RETRY_INTERNAL);
    return st;
}

static plf_status check_counts(plf_ctx* ctx, const int32_t* n_out, int nframes)
{
    for (int f = 0; f < nframes; f++) {
        if (n_out[f] == -1) return plf_fail(ctx, PLF_ERR_CAPACITY, "frame %d: more than %d lines detected in one octave", f, LINE_DETCAP);
        if (n_out[f] == -2) return plf_fail(ctx, PLF_ERR_CAPACITY, "frame %d: keyline output capacity too small", f);
        if (n_out[f] == -3) return plf_fail(ctx, PLF_ERR_CAPACITY, "LSD region buffer overflow (more than %d regions in one frame)", LINE_REGCAP_PER_FRAME);
    }
    return PLF_OK;
}

static plf_status check_regions(plf_line* o)
{
    plf_ctx* ctx = o->ctx;
    for (int k = 0; k < o->prm.nlevels; k++) {
        int err = 0;
        { plf_status rs = read_ints(ctx, o->d_cnt[k] + CNT_ERR, 1, &err); if (rs) return rs; }
        if (err) return plf_fail(ctx, PLF_ERR_CAPACITY, "LSD region buffer overflow (more than %d regions in one frame)", LINE_REGCAP_PER_FRAME);
    }
    return PLF_OK;
}

static plf_status line_extract_to_host(plf_line* o, const uint8_t* host_imgs, bool device_src, int nframes, int w, int h, size_t stride,
                                       size_t frame_stride, plf_keyline* host_kl, plf_keypoint* host_mid, uint8_t* host_desc, int cap,
                                       int32_t* n_out)
{
    if (!o) return PLF_ERR_INVALID;
    plf_ctx* ctx = o->ctx;
    if (nframes < 1 || !n_out) return plf_fail(ctx, PLF_ERR_INVALID, "plf_line_extract_batch: bad arguments");
    if (!host_imgs || w <= 0 || h <= 0) {   // empty image: silent return (src/Lineextractor.cc:115-116)
        for (int f = 0; f < nframes; f++) n_out[f] = 0;
        return PLF_OK;
    }
    if (stride < (size_t)w || !host_kl || !host_desc || cap < 1) return plf_fail(ctx, PLF_ERR_INVALID, "plf_line_extract_batch: bad arguments");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    plf_status st;
    do {
        st = line_prepare(o, w, h, nframes);
        if (st) return st;
        st = line_out_staging(o, nframes, cap);
        if (st) return st;
        st = upload_images(o, host_imgs, nframes, w, h, stride, frame_stride, device_src);
        if (st) return st;
        st = line_extract_device_impl(o, nframes, o->d_okl, o->d_omid, o->d_odesc, cap, o->d_onout);
    } while (st == PLF_RETRY_INTERNAL);
    if (st) return st;
    cudaStream_t s = ctx->stream;
    PLF_CUDA(ctx, cudaMemcpyAsync(host_kl, o->d_okl, (size_t)nframes * cap * sizeof(plf_keyline), cudaMemcpyDeviceToHost, s));
    if (host_mid) PLF_CUDA(ctx, cudaMemcpyAsync(host_mid, o->d_omid, (size_t)nframes * cap * sizeof(plf_keypoint), cudaMemcpyDeviceToHost, s));
    PLF_CUDA(ctx, cudaMemcpyAsync(host_desc, o->d_odesc, (size_t)nframes * cap * 32, cudaMemcpyDeviceToHost, s));
    st = read_ints(ctx, o->d_onout, nframes, n_out);       // also completes the result copies above
    if (st) return st;
    st = check_regions(o);
    if (st) return st;
    return check_counts(ctx, n_out, nframes);
}

extern "C" plf_status plf_line_extract_batch(plf_line* o, const uint8_t* host_imgs, int nframes, int w, int h, size_t stride,
                                             size_t frame_stride, plf_keyline* host_kl, plf_keypoint* host_mid, uint8_t* host_desc,
                                             int cap, int32_t* n_out)
{
    return line_extract_to_host(o, host_imgs, false, nframes, w, h, stride, frame_stride, host_kl, host_mid, host_desc, cap, n_out);
}

extern "C" plf_status plf_line_extract_batch_from_device(plf_line* o, const uint8_t* dev_imgs, int nframes, int w, int h, size_t stride,
                                                         size_t frame_stride, plf_keyline* host_kl, plf_keypoint* host_mid,
                                                         uint8_t* host_desc, int cap, int32_t* n_out)
{
    return line_extract_to_host(o, dev_imgs, true, nframes, w, h, stride, frame_stride, host_kl, host_mid, host_desc, cap, n_out);
}

extern "C" plf_status plf_line_extract(plf_line* o, const uint8_t* host_img, int w, int h, size_t stride, plf_keyline* host_kl,
                                       plf_keypoint* host_mid, uint8_t* host_desc, int cap, int* n_out)
{
    if (!o || !n_out) return PLF_ERR_INVALID;
    int32_t n = 0;
    plf_status st = plf_line_extract_batch(o, host_img, 1, w, h, stride, stride * (size_t)(h > 0 ? h : 0), host_kl, host_mid, host_desc, cap, &n);
    *n_out = n;
    return st;
}

extern "C" plf_status plf_lsd_detect(plf_line* o, const uint8_t* host_img, int w, int h, size_t stride, plf_keyline* host_kl, int cap, int* n_out)
{
    if (!o || !n_out) return PLF_ERR_INVALID;
    plf_ctx* ctx = o->ctx;
    *n_out = 0;
    if (!host_img || w <= 0 || h <= 0) return PLF_OK;
    if (stride < (size_t)w || !host_kl || cap < 1) return plf_fail(ctx, PLF_ERR_INVALID, "plf_lsd_detect: bad arguments");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    plf_status st = line_prepare(o, w, h, 1);
    if (st) return st;
    st = line_out_staging(o, 1, cap);
    if (st) return st;
    st = upload_images(o, host_img, 1, w, h, stride, stride * (size_t)h, false);
    if (st) return st;
    st = lsd_detect_batch(o, 1);
    if (st) return st;
    st = line_select(o, 1, false, o->d_okl, nullptr, cap, o->d_onout);
    if (st) return st;
    int32_t n = 0;
    st = read_ints(ctx, o->d_onout, 1, &n);
    if (st) return st;
    st = check_regions(o);
    if (st) return st;
    st = check_counts(ctx, &n, 1);
    if (st) return st;
    PLF_CUDA(ctx, cudaMemcpyAsync(host_kl, o->d_okl, (size_t)n * sizeof(plf_keyline), cudaMemcpyDeviceToHost, ctx->stream));
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    return PLF_OK;
}

extern "C" plf_status plf_lbd_compute(plf_line* o, const uint8_t* host_img, int w, int h, size_t stride, const plf_keyline* host_kl, int n,
                                      uint8_t* host_desc, float* host_fdesc)
{
    if (!o) return PLF_ERR_INVALID;
    plf_ctx* ctx = o->ctx;
    if (n <= 0) return PLF_OK;   // the reference prints an error and returns (binary_descriptor_custom.cpp:556-560)
    if (!host_img || w <= 0 || h <= 0 || stride < (size_t)w || !host_kl || (!host_desc && !host_fdesc))
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_lbd_compute: bad arguments");
    for (int i = 0; i < n; i++)
        if (host_kl[i].octave < 0 || host_kl[i].octave >= o->prm.nlevels || host_kl[i].numOfPixels < 0 || host_kl[i].numOfPixels > 32767)
            return plf_fail(ctx, PLF_ERR_INVALID, "keyline %d: octave/numOfPixels out of range", i);
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    plf_status st = line_prepare(o, w, h, 1);
    if (st) return st;
    st = line_out_staging(o, 1, n > o->out_cap ? n : (o->out_cap > 0 ? o->out_cap : n));
    if (st) return st;
    const int cap = o->out_cap;
    st = upload_images(o, host_img, 1, w, h, stride, stride * (size_t)h, false);
    if (st) return st;
    PLF_CUDA(ctx, cudaMemcpyAsync(o->d_okl, host_kl, (size_t)n * sizeof(plf_keyline), cudaMemcpyHostToDevice, ctx->stream));
    PLF_CUDA(ctx, cudaMemcpyAsync(o->d_onout, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    st = lbd_batch(o, 1, o->d_okl, o->d_onout, cap, n, o->d_odesc, o->d_ofdesc);
    if (st) return st;
    if (host_desc) PLF_CUDA(ctx, cudaMemcpyAsync(host_desc, o->d_odesc, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (host_fdesc) PLF_CUDA(ctx, cudaMemcpyAsync(host_fdesc, o->d_ofdesc, (size_t)n * 72 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PLF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

#include "plf_fld_impl.cuh"
