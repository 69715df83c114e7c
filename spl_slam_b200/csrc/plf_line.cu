// plf_line.cu -- host driver of the line path (LSD + LBD). (skeleton; kernels follow)
#include "plf_common.cuh"
#include <math.h>

struct plf_line {
    plf_ctx* ctx;
    plf_line_params prm;
    float scale[16], inv_scale[16], sigma2[16], inv_sigma2[16];
    int per_level[16];
};

extern "C" plf_status plf_line_create(plf_ctx* ctx, const plf_line_params* p, plf_line** out)
{
    if (!ctx || !p || !out) return PLF_ERR_INVALID;
    if (p->nlevels < 1 || p->nlevels > 2 || p->nfeatures < 1 || p->refine != 0 || !(p->scale > 0) || p->n_bins < 2 || p->n_bins > 65536)
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_line_create: unsupported parameters (refine must be 0, nlevels 1..2)");
    plf_line* o = (plf_line*)calloc(1, sizeof(plf_line));
    o->ctx = ctx; o->prm = *p;
    const int n = p->nlevels;
    o->scale[0] = 1.0f; o->sigma2[0] = 1.0f;
    for (int i = 1; i < n; i++) { o->scale[i] = (float)((double)o->scale[i - 1] * p->scale); o->sigma2[i] = o->scale[i] * o->scale[i]; }
    for (int i = 0; i < n; i++) { o->inv_scale[i] = 1.0f / o->scale[i]; o->inv_sigma2[i] = 1.0f / o->sigma2[i]; }
    float factor = (float)(1.0f / p->scale);
    float nDesired = (float)(p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n)));
    int sum = 0;
    for (int l = 0; l < n - 1; l++) { o->per_level[l] = (int)lrintf(nDesired); sum += o->per_level[l]; nDesired *= factor; }
    o->per_level[n - 1] = p->nfeatures - sum > 0 ? p->nfeatures - sum : 0;
    *out = o;
    return PLF_OK;
}
extern "C" void plf_line_destroy(plf_line* o) { free(o); }
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
#define NOTYET(ctx) return plf_fail((ctx), PLF_ERR_STATE, "line path not built yet")
extern "C" plf_status plf_line_extract(plf_line* le, const uint8_t*, int, int, size_t, plf_keyline*, plf_keypoint*, uint8_t*, int, int*) { NOTYET(le->ctx); }
extern "C" plf_status plf_line_extract_batch(plf_line* le, const uint8_t*, int, int, int, size_t, size_t, plf_keyline*, plf_keypoint*, uint8_t*, int, int32_t*) { NOTYET(le->ctx); }
extern "C" plf_status plf_line_extract_batch_device(plf_line* le, const uint8_t*, int, int, int, size_t, size_t, plf_keyline*, plf_keypoint*, uint8_t*, int, int32_t*) { NOTYET(le->ctx); }
extern "C" plf_status plf_lsd_detect(plf_line* le, const uint8_t*, int, int, size_t, plf_keyline*, int, int*) { NOTYET(le->ctx); }
extern "C" plf_status plf_lbd_compute(plf_line* le, const uint8_t*, int, int, size_t, const plf_keyline*, int, uint8_t*, float*) { NOTYET(le->ctx); }
