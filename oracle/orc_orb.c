/*
 * orc_orb.c -- oracle restatement of PL_SLAM::ORBextractor (src/ORBextractor.cc).
 * TEST INFRASTRUCTURE ONLY (see plf_oracle.h).
 *
 * Documented choices where the reference is not deterministic / toolchain dependent
 * (SURVEY.md section 7 "hard parts"):
 *  - DistributeOctTree refinement sort (ORBextractor.cc:684) orders equal-size nodes by
 *    heap address; the oracle orders them by creation index (ascending), iterated from
 *    the back, i.e. "later-created first".  This IS the reference's order when heap addresses
 *    grow with allocation order (oracle/_ref heap mode 1: identical keypoint lists, order included);
 *    under glibc malloc the reference itself returns different sets for the same input
 *    (tests/test_oracle_vs_ref.py).
 *  - rBRIEF rotation (ORBextractor.cc:112-120): `cos(angle)` on a float under `using namespace std`
 *    is cosf; products and sum in IEEE single without FMA; cvRound = round-half-even.
 * Pinned bit for bit (keypoints, order, angles, descriptors) against the reference's own
 * ORBextractor.cc compiled into oracle/_ref/libref.so.
 */
#include "plf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PATCH_SIZE 31
#define HALF_PATCH_SIZE 15
#define EDGE_THRESHOLD 19

static const int8_t bit_pattern_31[256 * 4] = {
#include "../spl_slam_b200/csrc/orb_pattern.inc"
};

typedef struct {
    int w, h;          /* level size */
    size_t stride;     /* stride of the bordered buffer */
    uint8_t* buf;      /* (w+38)x(h+38) bordered */
    uint8_t* blurred;  /* w x h, or NULL */
    int nraw, nkept;
    int *rx, *ry, *rr; /* raw FAST keys in distribute order */
} orb_level;

struct orc_orb {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;
    float* scale;
    float* inv_scale;
    int* per_level;
    int umax[HALF_PATCH_SIZE + 2];
    orb_level* lv;
};

static inline int cv_round_f(float v) { return (int)lrintf(v); }

/* ORBextractor::ORBextractor, src/ORBextractor.cc:410-470 */
orc_orb* orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh)
{
    orc_orb* o = (orc_orb*)calloc(1, sizeof(*o));
    o->nfeatures = nfeatures; o->nlevels = nlevels; o->iniTh = iniTh; o->minTh = minTh;
    o->scaleFactor = scaleFactor; /* member is double, assigned from float */
    o->scale = (float*)calloc((size_t)nlevels, sizeof(float));
    o->inv_scale = (float*)calloc((size_t)nlevels, sizeof(float));
    o->per_level = (int*)calloc((size_t)nlevels, sizeof(int));
    o->lv = (orb_level*)calloc((size_t)nlevels, sizeof(orb_level));
    o->scale[0] = 1.0f;
    /* mvScaleFactor[i-1]*scaleFactor: float * double -> double -> float */
    for (int i = 1; i < nlevels; i++) o->scale[i] = (float)((double)o->scale[i - 1] * o->scaleFactor);
    for (int i = 0; i < nlevels; i++) o->inv_scale[i] = 1.0f / o->scale[i];

    float factor = (float)(1.0f / o->scaleFactor);
    float nDesired = (float)(nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels)));
    int sum = 0;
    for (int level = 0; level < nlevels - 1; level++) {
        o->per_level[level] = cv_round_f(nDesired);
        sum += o->per_level[level];
        nDesired *= factor;
    }
    o->per_level[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;

    int v, v0, vmax = (int)floor(HALF_PATCH_SIZE * sqrt(2.f) / 2 + 1);
    int vmin = (int)ceil(HALF_PATCH_SIZE * sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) o->umax[v] = (int)lrint(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (o->umax[v0] == o->umax[v0 + 1]) ++v0;
        o->umax[v] = v0;
        ++v0;
    }
    return o;
}

static void level_free(orb_level* l)
{
    free(l->buf); free(l->blurred); free(l->rx); free(l->ry); free(l->rr);
    memset(l, 0, sizeof(*l));
}

void orc_orb_destroy(orc_orb* o)
{
    if (!o) return;
    for (int i = 0; i < o->nlevels; i++) level_free(&o->lv[i]);
    free(o->lv); free(o->scale); free(o->inv_scale); free(o->per_level); free(o);
}

int orc_orb_features_per_level(const orc_orb* o, int level) { return o->per_level[level]; }
float orc_orb_scale_factor(const orc_orb* o, int level) { return o->scale[level]; }
int orc_orb_umax(const orc_orb* o, int v) { return o->umax[v]; }

/* ---------------- DistributeOctTree, src/ORBextractor.cc:481-763 ---------------- */
typedef struct {
    int ulx, uly, brx, bry; /* UL and BR corners (UR.x == BR.x, BL.y == BR.y) */
    int kbeg, kcnt;         /* keys: slice of the key-index pool */
    int prev, next;         /* list links */
    int nomore;
} onode;

typedef struct {
    onode* nodes; int nnodes, capnodes;
    int* pool; int npool, cappool;
    int head, tail, size;
    const int *xs, *ys;
} otree;

static int node_new(otree* t)
{
    if (t->nnodes == t->capnodes) {
        t->capnodes = t->capnodes * 2 + 64;
        t->nodes = (onode*)realloc(t->nodes, sizeof(onode) * (size_t)t->capnodes);
    }
    onode* n = &t->nodes[t->nnodes];
    memset(n, 0, sizeof(*n));
    n->prev = n->next = -1;
    return t->nnodes++;
}
static int pool_reserve(otree* t, int cnt)
{
    if (t->npool + cnt > t->cappool) {
        t->cappool = (t->npool + cnt) * 2 + 1024;
        t->pool = (int*)realloc(t->pool, sizeof(int) * (size_t)t->cappool);
    }
    int b = t->npool;
    t->npool += cnt;
    return b;
}
static void list_push_front(otree* t, int id)
{
    onode* n = &t->nodes[id];
    n->prev = -1; n->next = t->head;
    if (t->head >= 0) t->nodes[t->head].prev = id; else t->tail = id;
    t->head = id; t->size++;
}
static void list_push_back(otree* t, int id)
{
    onode* n = &t->nodes[id];
    n->next = -1; n->prev = t->tail;
    if (t->tail >= 0) t->nodes[t->tail].next = id; else t->head = id;
    t->tail = id; t->size++;
}
static int list_erase(otree* t, int id) /* returns next */
{
    onode* n = &t->nodes[id];
    int nx = n->next;
    if (n->prev >= 0) t->nodes[n->prev].next = n->next; else t->head = n->next;
    if (n->next >= 0) t->nodes[n->next].prev = n->prev; else t->tail = n->prev;
    t->size--;
    return nx;
}

/* ExtractorNode::DivideNode (:481-537). Children ids are returned in c[0..3] (n1..n4). */
static void divide_node(otree* t, int id, int c[4])
{
    for (int i = 0; i < 4; i++) c[i] = node_new(t);
    onode P = t->nodes[id];
    int halfX = (int)ceilf((float)(P.brx - P.ulx) / 2);
    int halfY = (int)ceilf((float)(P.bry - P.uly) / 2);
    int mx = P.ulx + halfX, my = P.uly + halfY;
    onode* n1 = &t->nodes[c[0]]; onode* n2 = &t->nodes[c[1]];
    onode* n3 = &t->nodes[c[2]]; onode* n4 = &t->nodes[c[3]];
    n1->ulx = P.ulx; n1->uly = P.uly; n1->brx = mx;    n1->bry = my;
    n2->ulx = mx;    n2->uly = P.uly; n2->brx = P.brx; n2->bry = my;
    n3->ulx = P.ulx; n3->uly = my;    n3->brx = mx;    n3->bry = P.bry;
    n4->ulx = mx;    n4->uly = my;    n4->brx = P.brx; n4->bry = P.bry;
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < P.kcnt; i++) {
        int k = t->pool[P.kbeg + i];
        int q = (t->xs[k] < mx ? 0 : 1) + (t->ys[k] < my ? 0 : 2);
        cnt[q]++;
    }
    int beg[4];
    for (int q = 0; q < 4; q++) beg[q] = pool_reserve(t, cnt[q]);
    int fill[4] = {0, 0, 0, 0};
    for (int i = 0; i < P.kcnt; i++) {
        int k = t->pool[P.kbeg + i];
        int q = (t->xs[k] < mx ? 0 : 1) + (t->ys[k] < my ? 0 : 2);
        t->pool[beg[q] + fill[q]++] = k;
    }
    for (int q = 0; q < 4; q++) {
        onode* n = &t->nodes[c[q]];
        n->kbeg = beg[q]; n->kcnt = cnt[q]; n->nomore = (cnt[q] == 1);
    }
}

typedef struct { int size, id; } szid;
static int szid_cmp(const void* a, const void* b)
{
    const szid* x = (const szid*)a; const szid* y = (const szid*)b;
    if (x->size != y->size) return x->size < y->size ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id);
}

int orc_distribute_octree(const int* xs, const int* ys, const int* resp, int n,
                          int minX, int maxX, int minY, int maxY, int N, int* out_idx, int cap)
{
    otree T; memset(&T, 0, sizeof(T));
    T.head = T.tail = -1; T.xs = xs; T.ys = ys;
    otree* t = &T;
    int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) nIni = 1; /* reference divides by zero here (tall images); defined as one strip */
    const float hX = (float)(maxX - minX) / nIni;
    /* initial strips (:552-563) */
    int* cnt0 = (int*)calloc((size_t)(nIni > 0 ? nIni : 1), sizeof(int));
#define STRIP(i) ({ int s_ = (int)((float)xs[i] / hX); s_ >= nIni ? nIni - 1 : s_; })
    for (int i = 0; i < n; i++) cnt0[STRIP(i)]++;
    for (int i = 0; i < nIni; i++) {
        int id = node_new(t);
        onode* nd = &t->nodes[id];
        nd->ulx = (int)(hX * (float)i); nd->uly = 0;
        nd->brx = (int)(hX * (float)(i + 1)); nd->bry = maxY - minY;
        nd->kbeg = pool_reserve(t, cnt0[i]); nd->kcnt = 0;
        list_push_back(t, id);
    }
    for (int i = 0; i < n; i++) {
        onode* nd = &t->nodes[STRIP(i)];
        t->pool[nd->kbeg + nd->kcnt++] = i;
    }
#undef STRIP
    free(cnt0);
    for (int id = t->head; id >= 0;) {
        onode* nd = &t->nodes[id];
        if (nd->kcnt == 1) { nd->nomore = 1; id = nd->next; }
        else if (nd->kcnt == 0) id = list_erase(t, id);
        else id = nd->next;
    }

    int finish = 0;
    szid* vsz = NULL; int nvsz = 0, capvsz = 0;
    szid* vprev = NULL; int capprev = 0;
#define VSZ_PUSH(s, i) do { if (nvsz == capvsz) { capvsz = capvsz * 2 + 64; \
        vsz = (szid*)realloc(vsz, sizeof(szid) * (size_t)capvsz); } vsz[nvsz].size = (s); vsz[nvsz].id = (i); nvsz++; } while (0)
    while (!finish) {
        int prevSize = t->size;
        int nToExpand = 0;
        nvsz = 0;
        for (int id = t->head; id >= 0;) {
            if (t->nodes[id].nomore) { id = t->nodes[id].next; continue; }
            int c[4];
            divide_node(t, id, c);
            for (int q = 0; q < 4; q++) {
                int kc = t->nodes[c[q]].kcnt;
                if (kc > 0) {
                    list_push_front(t, c[q]);
                    if (kc > 1) { nToExpand++; VSZ_PUSH(kc, c[q]); }
                }
            }
            id = list_erase(t, id);
        }
        if (t->size >= N || t->size == prevSize) {
            finish = 1;
        } else if (t->size + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = t->size;
                if (nvsz > capprev) { capprev = nvsz; vprev = (szid*)realloc(vprev, sizeof(szid) * (size_t)capprev); }
                int np = nvsz;
                memcpy(vprev, vsz, sizeof(szid) * (size_t)np);
                nvsz = 0;
                qsort(vprev, (size_t)np, sizeof(szid), szid_cmp);
                for (int j = np - 1; j >= 0; j--) {
                    int c[4];
                    divide_node(t, vprev[j].id, c);
                    for (int q = 0; q < 4; q++) {
                        int kc = t->nodes[c[q]].kcnt;
                        if (kc > 0) {
                            list_push_front(t, c[q]);
                            if (kc > 1) VSZ_PUSH(kc, c[q]);
                        }
                    }
                    list_erase(t, vprev[j].id);
                    if (t->size >= N) break;
                }
                if (t->size >= N || t->size == prevSize) finish = 1;
            }
        }
    }
#undef VSZ_PUSH
    /* best key per node (:741-760): first key with maximal response */
    int m = 0;
    for (int id = t->head; id >= 0; id = t->nodes[id].next) {
        onode* nd = &t->nodes[id];
        int best = t->pool[nd->kbeg];
        for (int k = 1; k < nd->kcnt; k++) {
            int kk = t->pool[nd->kbeg + k];
            if (resp[kk] > resp[best]) best = kk;
        }
        if (m < cap) out_idx[m] = best;
        m++;
    }
    free(vsz); free(vprev); free(t->nodes); free(t->pool);
    return m;
}

/* IC_Angle, src/ORBextractor.cc:77-104 */
static float ic_angle(const uint8_t* center, int step, const int* umax)
{
    int m_01 = 0, m_10 = 0;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
}

/* computeOrbDescriptor, src/ORBextractor.cc:107-147 */
static void orb_descriptor(float angle_deg, const uint8_t* center, int step, uint8_t* desc)
{
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    float angle = angle_deg * factorPI;
    float a = cosf(angle), b = sinf(angle); /* cos(float) under `using namespace std` = std::cos(float) = cosf */
    const int8_t* p = bit_pattern_31;
    for (int i = 0; i < 32; i++) {
        int val = 0;
        for (int k = 0; k < 8; k++, p += 4) {
            float x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
            float r0 = x0 * b, r1 = y0 * a, r2 = x0 * a, r3 = y0 * b;
            int t0 = center[cv_round_f(r0 + r1) * step + cv_round_f(r2 - r3)];
            r0 = x1 * b; r1 = y1 * a; r2 = x1 * a; r3 = y1 * b;
            int t1 = center[cv_round_f(r0 + r1) * step + cv_round_f(r2 - r3)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ComputePyramid (:1107-1132) */
static void compute_pyramid(orc_orb* o, const uint8_t* img, int w, int h, size_t stride)
{
    for (int level = 0; level < o->nlevels; level++) {
        orb_level* L = &o->lv[level];
        level_free(L);
        float scale = o->inv_scale[level];
        L->w = cv_round_f((float)w * scale);
        L->h = cv_round_f((float)h * scale);
        L->stride = (size_t)(L->w + 2 * EDGE_THRESHOLD);
        L->buf = (uint8_t*)malloc(L->stride * (size_t)(L->h + 2 * EDGE_THRESHOLD));
        uint8_t* roi = L->buf + EDGE_THRESHOLD * L->stride + EDGE_THRESHOLD;
        if (level != 0) {
            orb_level* P = &o->lv[level - 1];
            const uint8_t* proi = P->buf + EDGE_THRESHOLD * P->stride + EDGE_THRESHOLD;
            uint8_t* tmp = (uint8_t*)malloc((size_t)L->w * (size_t)L->h);
            orc_resize_linear_u8(proi, P->w, P->h, P->stride, tmp, L->w, L->h, (size_t)L->w);
            orc_border_reflect101_u8(tmp, L->w, L->h, (size_t)L->w, L->buf, EDGE_THRESHOLD, L->stride);
            free(tmp);
        } else {
            orc_border_reflect101_u8(img, w, h, stride, L->buf, EDGE_THRESHOLD, L->stride);
        }
        (void)roi;
    }
}

int orc_orb_extract(orc_orb* o, const uint8_t* img, int w, int h, size_t stride,
                    orc_keypoint* kps, uint8_t* desc, int cap)
{
    if (!img || w <= 0 || h <= 0) return 0;
    compute_pyramid(o, img, w, h, stride);
    int total = 0;
    const int W = 30;
    int capraw = 0; int *fx = NULL, *fy = NULL, *fs = NULL;
    /* ComputeKeyPointsOctTree (:765-853) */
    int** kept = (int**)calloc((size_t)o->nlevels, sizeof(int*));
    for (int level = 0; level < o->nlevels; level++) {
        orb_level* L = &o->lv[level];
        const uint8_t* roi = L->buf + EDGE_THRESHOLD * L->stride + EDGE_THRESHOLD;
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
        const int maxBorderX = L->w - EDGE_THRESHOLD + 3, maxBorderY = L->h - EDGE_THRESHOLD + 3;
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        L->nraw = 0; L->nkept = 0;
        if (nCols <= 0 || nRows <= 0) continue;
        const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
        int rawcap = 4096, nraw = 0;
        L->rx = (int*)malloc(sizeof(int) * (size_t)rawcap);
        L->ry = (int*)malloc(sizeof(int) * (size_t)rawcap);
        L->rr = (int*)malloc(sizeof(int) * (size_t)rawcap);
        int cellcap = (wCell + 6) * (hCell + 6);
        if (cellcap > capraw) {
            capraw = cellcap;
            fx = (int*)realloc(fx, sizeof(int) * (size_t)capraw);
            fy = (int*)realloc(fy, sizeof(int) * (size_t)capraw);
            fs = (int*)realloc(fs, sizeof(int) * (size_t)capraw);
        }
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                int cw = (int)maxX - (int)iniX, ch = (int)maxY - (int)iniY;
                const uint8_t* cell = roi + (size_t)(int)iniY * L->stride + (int)iniX;
                int nc = orc_fast9(cell, cw, ch, L->stride, o->iniTh, fx, fy, fs, capraw);
                if (nc == 0) nc = orc_fast9(cell, cw, ch, L->stride, o->minTh, fx, fy, fs, capraw);
                if (nraw + nc > rawcap) {
                    rawcap = (nraw + nc) * 2;
                    L->rx = (int*)realloc(L->rx, sizeof(int) * (size_t)rawcap);
                    L->ry = (int*)realloc(L->ry, sizeof(int) * (size_t)rawcap);
                    L->rr = (int*)realloc(L->rr, sizeof(int) * (size_t)rawcap);
                }
                for (int k = 0; k < nc; k++) {
                    L->rx[nraw] = fx[k] + j * wCell;
                    L->ry[nraw] = fy[k] + i * hCell;
                    L->rr[nraw] = fs[k];
                    nraw++;
                }
            }
        }
        L->nraw = nraw;
        int kcap = nraw > 0 ? nraw : 1;
        kept[level] = (int*)malloc(sizeof(int) * (size_t)kcap);
        L->nkept = nraw ? orc_distribute_octree(L->rx, L->ry, L->rr, nraw, minBorderX, maxBorderX,
                                                minBorderY, maxBorderY, o->per_level[level], kept[level], kcap) : 0;
        total += L->nkept;
    }
    free(fx); free(fy); free(fs);
    if (total > cap) { for (int l = 0; l < o->nlevels; l++) free(kept[l]); free(kept); return -1; }

    /* orientation on the un-blurred level, then blur + descriptors, scale, concat (:851-852, :1076-1104) */
    int off = 0;
    for (int level = 0; level < o->nlevels; level++) {
        orb_level* L = &o->lv[level];
        if (L->nkept == 0) continue;
        const uint8_t* roi = L->buf + EDGE_THRESHOLD * L->stride + EDGE_THRESHOLD;
        L->blurred = (uint8_t*)malloc((size_t)L->w * (size_t)L->h);
        orc_gauss_blur_u8(roi, L->w, L->h, L->stride, L->blurred, (size_t)L->w, 7, 2.0);
        const int scaledPatchSize = (int)(PATCH_SIZE * o->scale[level]);
        float sc = o->scale[level];
        for (int k = 0; k < L->nkept; k++) {
            int r = kept[level][k];
            int x = L->rx[r] + (EDGE_THRESHOLD - 3), y = L->ry[r] + (EDGE_THRESHOLD - 3);
            orc_keypoint* kp = &kps[off + k];
            kp->angle = ic_angle(roi + (size_t)y * L->stride + x, (int)L->stride, o->umax);
            orb_descriptor(kp->angle, L->blurred + (size_t)y * L->w + x, L->w, desc + (size_t)(off + k) * 32);
            kp->x = (float)x; kp->y = (float)y;
            if (level != 0) { kp->x *= sc; kp->y *= sc; }
            kp->size = (float)scaledPatchSize;
            kp->response = (float)L->rr[r];
            kp->octave = level;
            kp->class_id = -1;
        }
        off += L->nkept;
    }
    for (int l = 0; l < o->nlevels; l++) free(kept[l]);
    free(kept);
    return total;
}

int orc_orb_level_size(const orc_orb* o, int level, int* w, int* h)
{
    if (level < 0 || level >= o->nlevels || !o->lv[level].buf) return -1;
    *w = o->lv[level].w; *h = o->lv[level].h;
    return 0;
}
const uint8_t* orc_orb_level_image(const orc_orb* o, int level, size_t* stride)
{
    const orb_level* L = &o->lv[level];
    *stride = L->stride;
    return L->buf + EDGE_THRESHOLD * L->stride + EDGE_THRESHOLD;
}
const uint8_t* orc_orb_level_blurred(const orc_orb* o, int level, size_t* stride)
{
    *stride = (size_t)o->lv[level].w;
    return o->lv[level].blurred;
}
int orc_orb_level_raw_count(const orc_orb* o, int level) { return o->lv[level].nraw; }
void orc_orb_level_raw(const orc_orb* o, int level, int* xs, int* ys, int* resp)
{
    const orb_level* L = &o->lv[level];
    memcpy(xs, L->rx, sizeof(int) * (size_t)L->nraw);
    memcpy(ys, L->ry, sizeof(int) * (size_t)L->nraw);
    memcpy(resp, L->rr, sizeof(int) * (size_t)L->nraw);
}
int orc_orb_level_kept_count(const orc_orb* o, int level) { return o->lv[level].nkept; }

/* ---------------- Frame::ComputeStereoMatches, src/Frame.cc:881-1055 ----------------
 * Row-band best-1 Hamming (octave +-1, disparity window, TH_HIGH init, first-best on ties) + 11x11 SAD slide
 * on the two extractors' pyramids + parabola sub-pixel + 2.1 x median SAD filter.  oL / oR must hold the
 * pyramids of the two images (state after orc_orb_extract).  Windows that would leave the level image are
 * skipped (the reference would throw in cv::Mat::colRange); an empty match list is a no-op (reference UB). */
static int cmp_pair(const void* a, const void* b)
{
    const int* x = (const int*)a; const int* y = (const int*)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1]);
}

void orc_stereo_match(const orc_orb* oL, const orc_orb* oR, const orc_keypoint* kL, const uint8_t* dL, int nL,
                      const orc_keypoint* kR, const uint8_t* dR, int nR, float mb, float mbf,
                      float* uRight, float* depth)
{
    const int TH_HIGH = 100, TH_LOW = 50, thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = oL->lv[0].h;
    for (int i = 0; i < nL; i++) { uRight[i] = -1.0f; depth[i] = -1.0f; }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    int (*pairs)[2] = (int (*)[2])malloc(sizeof(int[2]) * (size_t)(nL > 0 ? nL : 1));
    int np = 0;
    for (int iL = 0; iL < nL; iL++) {
        const int levelL = kL[iL].octave;
        const float vL = kL[iL].y, uL = kL[iL].x;
        const int row = (int)vL;
        if (row < 0 || row >= nRows) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH, bestIdxR = 0;
        for (int iR = 0; iR < nR; iR++) {
            /* iR is in vRowIndices[row] iff floor(y - r) <= row <= ceil(y + r) */
            const float r = 2.0f * oL->scale[kR[iR].octave];
            const int maxr = (int)ceilf(kR[iR].y + r), minr = (int)floorf(kR[iR].y - r);
            if (row < minr || row > maxr) continue;
            if (kR[iR].octave < levelL - 1 || kR[iR].octave > levelL + 1) continue;
            const float uR = kR[iR].x;
            if (uR >= minU && uR <= maxU) {
                int dist = orc_descriptor_distance(dL + (size_t)iL * 32, dR + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (!(bestDist < thOrbDist)) continue;
        const float uR0 = kR[bestIdxR].x;
        const float scaleFactor = oL->inv_scale[levelL];
        const float scaleduL = roundf(kL[iL].x * scaleFactor), scaledvL = roundf(kL[iL].y * scaleFactor);
        const float scaleduR0 = roundf(uR0 * scaleFactor);
        const int w = 5, L = 5;
        const orb_level* PL = &oL->lv[levelL];
        const orb_level* PR = &oR->lv[levelL];
        const uint8_t* imL = PL->buf + EDGE_THRESHOLD * PL->stride + EDGE_THRESHOLD;
        const uint8_t* imR = PR->buf + EDGE_THRESHOLD * PR->stride + EDGE_THRESHOLD;
        const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
        if (iniu < 0 || endu >= PR->w) continue;
        const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
        if (cv - w < 0 || cv + w >= PL->h || cu - w < 0 || cu + w >= PL->w || cr - L - w < 0 || cr + L + w >= PR->w || cv + w >= PR->h)
            continue;
        int bestSad = 2147483647, bestincR = 0;
        float vDists[11];
        const int cL = imL[(size_t)cv * PL->stride + cu];
        for (int incR = -L; incR <= L; incR++) {
            const int cR = imR[(size_t)cv * PR->stride + cr + incR];
            int sad = 0;
            for (int dy = -w; dy <= w; dy++)
                for (int dx = -w; dx <= w; dx++) {
                    int a = imL[(size_t)(cv + dy) * PL->stride + cu + dx] - cL;
                    int b = imR[(size_t)(cv + dy) * PR->stride + cr + incR + dx] - cR;
                    sad += abs(a - b);
                }
            float dist = (float)sad;
            if (dist < (float)bestSad) { bestSad = (int)dist; bestincR = incR; }
            vDists[L + incR] = dist;
        }
        if (bestincR == -L || bestincR == L) continue;
        const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
        const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
        if (deltaR < -1 || deltaR > 1) continue;
        float bestuR = oL->scale[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);
        float disparity = (uL - bestuR);
        if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) { disparity = (float)0.01; bestuR = (float)(uL - 0.01); }
            depth[iL] = mbf / disparity;
            uRight[iL] = bestuR;
            pairs[np][0] = bestSad; pairs[np][1] = iL; np++;
        }
    }
    if (np > 0) {
        qsort(pairs, (size_t)np, sizeof(int[2]), cmp_pair);
        const float median = (float)pairs[np / 2][0];
        const float thDist = 1.5f * 1.4f * median;
        for (int i = np - 1; i >= 0; i--) {
            if ((float)pairs[i][0] < thDist) break;
            uRight[pairs[i][1]] = -1; depth[pairs[i][1]] = -1;
        }
    }
    free(pairs);
}

/* ---------------- Frame grid + GetFeaturesInArea, src/Frame.cc:365-399, :562-722 ----------------
 * cell lists are built by pushing feature indices in ascending order (AssignFeaturesToGrid); the query walks the
 * cells column-major and keeps the insertion order, like the reference's vector<size_t> code. */
static int grid_pos(float x, float y, const orc_grid_params* g, int* cx, int* cy)
{
    *cx = (int)roundf((x - g->min_x) * g->inv_w);
    *cy = (int)roundf((y - g->min_y) * g->inv_h);
    return !(*cx < 0 || *cx >= g->cols || *cy < 0 || *cy >= g->rows);
}

int orc_grid_candidates(const orc_keypoint* kps, const orc_keyline* kls, int n, const orc_grid_params* g, const float* qx,
                        const float* qy, const float* qr, const int32_t* qminl, const int32_t* qmaxl, int nq, int32_t* cand_off,
                        int32_t* cand_idx, int cand_cap)
{
    const int ncell = g->cols * g->rows;
    int* cell = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* start = (int*)calloc((size_t)ncell + 2, sizeof(int));
    int* items = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        int cx, cy, ax, ay;
        int ok;
        if (kls) {   /* PosInGridLines: start point, end point, then mid point */
            ok = grid_pos(kls[i].startPointX, kls[i].startPointY, g, &ax, &ay) && grid_pos(kls[i].endPointX, kls[i].endPointY, g, &ax, &ay) &&
                 grid_pos(kps[i].x, kps[i].y, g, &cx, &cy);
        } else ok = grid_pos(kps[i].x, kps[i].y, g, &cx, &cy);
        cell[i] = ok ? cx * g->rows + cy : -1;
        if (ok) start[cell[i] + 1]++;
    }
    for (int c = 0; c < ncell; c++) start[c + 1] += start[c];
    int* fill = (int*)calloc((size_t)ncell + 1, sizeof(int));
    for (int i = 0; i < n; i++)
        if (cell[i] >= 0) items[start[cell[i]] + fill[cell[i]]++] = i;
    int total = 0;
    for (int q = 0; q < nq; q++) {
        cand_off[q] = total;
        const float x = qx[q], y = qy[q], r = qr[q];
        const int minLevel = qminl ? qminl[q] : -1, maxLevel = qmaxl ? qmaxl[q] : -1;
        int nMinCellX = (int)floorf((x - g->min_x - r) * g->inv_w); if (nMinCellX < 0) nMinCellX = 0;
        if (nMinCellX >= g->cols) continue;
        int nMaxCellX = (int)ceilf((x - g->min_x + r) * g->inv_w); if (nMaxCellX > g->cols - 1) nMaxCellX = g->cols - 1;
        if (nMaxCellX < 0) continue;
        int nMinCellY = (int)floorf((y - g->min_y - r) * g->inv_h); if (nMinCellY < 0) nMinCellY = 0;
        if (nMinCellY >= g->rows) continue;
        int nMaxCellY = (int)ceilf((y - g->min_y + r) * g->inv_h); if (nMaxCellY > g->rows - 1) nMaxCellY = g->rows - 1;
        if (nMaxCellY < 0) continue;
        const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const int c = ix * g->rows + iy;
                for (int j = start[c]; j < start[c + 1]; j++) {
                    const orc_keypoint* kp = &kps[items[j]];
                    if (bCheckLevels) {
                        if (kp->octave < minLevel) continue;
                        if (maxLevel >= 0 && kp->octave > maxLevel) continue;
                    }
                    const float distx = kp->x - x, disty = kp->y - y;
                    if (fabsf(distx) < r && fabsf(disty) < r) {
                        if (total < cand_cap) cand_idx[total] = items[j];
                        total++;
                    }
                }
            }
    }
    cand_off[nq] = total;
    free(cell); free(start); free(items); free(fill);
    return total;
}

/* ---------------- Frame::UndistortKeyPoints / UndistortKeyLines, src/Frame.cc:733-826 ----------------
 * cv::undistortPoints(pts, pts, K, D, Mat(), K) (OpenCV cvUndistortPointsInternal, default criteria = 5 iterations):
 * pinned bit-for-bit against cv2 4.13 in tests/test_oracle_vs_cv2.py.  k = k1, k2, p1, p2[, k3]. */
void orc_undistort_points(const float* cam /* fx fy cx cy */, const float* kd, int nk, const float* in, int n, float* out)
{
    const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
    const double k1 = kd[0], k2 = kd[1], p1 = kd[2], p2 = kd[3], k3 = nk > 4 ? kd[4] : 0.0;
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    for (int i = 0; i < n; i++) {
        const double u = in[2 * i], v = in[2 * i + 1];
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
        out[2 * i] = (float)(fx * x + cx);
        out[2 * i + 1] = (float)(fy * y + cy);
    }
}
