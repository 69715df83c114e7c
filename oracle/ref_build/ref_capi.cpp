/*
 * ref_capi.cpp -- C entry points over the REFERENCE'S OWN classes (oracle/_ref/libref.so).  TEST INFRASTRUCTURE ONLY.
 *
 * Everything computed here is computed by code compiled from /root/reference (see Makefile); this file only marshals
 * plain arrays in and out so that tests (ctypes) and bench.py's --impl reference leg can call it.
 */
#include "ORBextractor.h"
#include "Lineextractor.h"
#include "ref_matchers_decl.hpp"
#include "../plf_oracle.h"
#include <cstdio>
#include <mutex>
#include <new>
#include <sys/mman.h>

/*
 * Heap modes.  DistributeOctTree orders equal-size nodes by their HEAP ADDRESS (std::sort of pair<int, ExtractorNode*>,
 * src/ORBextractor.cc:684), so what the reference returns depends on the allocator's history: under glibc malloc two calls
 * on the same input in one process return different keypoint sets (tests/test_oracle_vs_ref.py records this).
 *   mode 0: operator new = malloc (the reference as it runs).
 *   mode 1: operator new = a bump arena whose addresses only grow and are never reused, reset at every call: the tie order
 *           of a fresh, unfragmented heap ("later allocated = higher address").  Deterministic; this is the mode the oracle
 *           and the CUDA path are pinned against.
 * The operators below replace the global ones for this library only (linked with -Bsymbolic).
 */
namespace {
struct Arena {
    char* base = 0;
    size_t cap = 0, used = 0;
} g_arena;
thread_local bool g_arena_on = false;
std::mutex g_arena_mu;
int g_heap_mode = 0;
const size_t ARENA_BYTES = (size_t)8 << 30;

struct ArenaScope {
    bool active;
    ArenaScope() : active(g_heap_mode == 1)
    {
        if (!active) return;
        g_arena_mu.lock();
        if (!g_arena.base) {
            void* p = mmap(0, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (p == MAP_FAILED) { g_arena_mu.unlock(); throw std::bad_alloc(); }
            g_arena.base = (char*)p;
            g_arena.cap = ARENA_BYTES;
        }
        g_arena.used = 0;
        g_arena_on = true;
    }
    ~ArenaScope()
    {
        if (!active) return;
        g_arena_on = false;
        if (g_arena.used) madvise(g_arena.base, g_arena.used, MADV_DONTNEED);
        g_arena_mu.unlock();
    }
};
inline bool in_arena(void* p) { return g_arena.base && (char*)p >= g_arena.base && (char*)p < g_arena.base + g_arena.cap; }
}

void* operator new(size_t n)
{
    if (g_arena_on) {
        size_t a = (n + 15) & ~(size_t)15;
        if (g_arena.used + a > g_arena.cap) throw std::bad_alloc();
        void* p = g_arena.base + g_arena.used;
        g_arena.used += a ? a : 16;
        return p;
    }
    void* p = malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { if (p && !in_arena(p)) free(p); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

static_assert(sizeof(cv::KeyPoint) == sizeof(orc_keypoint), "cv::KeyPoint layout");
static_assert(sizeof(cv::line_descriptor::KeyLine) == sizeof(orc_keyline), "KeyLine layout");

namespace {
/* DistributeOctTree and the per-level tables are protected members */
struct OrbProbe : public PL_SLAM::ORBextractor {
    OrbProbe(int n, float s, int l, int a, int b) : PL_SLAM::ORBextractor(n, s, l, a, b) {}
    using PL_SLAM::ORBextractor::DistributeOctTree;
    using PL_SLAM::ORBextractor::ComputePyramid;
    using PL_SLAM::ORBextractor::ComputeKeyPointsOctTree;
    int featuresPerLevel(int l) const { return mnFeaturesPerLevel[l]; }
    int umaxAt(int v) const { return umax[v]; }
};
struct LineProbe : public PL_SLAM::Lineextractor {
    using PL_SLAM::Lineextractor::Lineextractor;
    int featuresPerLevel(int l) const { return mnFeaturesPerLevel[l]; }
};
thread_local char g_err[512];
template <typename F> int guarded(F f)
{
    try { return f(); }
    catch (const std::exception& e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1000; }
}
}

extern "C" {

const char* ref_last_error(void) { return g_err; }
void ref_set_heap_mode(int mode) { g_heap_mode = mode; }
int ref_get_heap_mode(void) { return g_heap_mode; }

void* ref_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh)
{
    return new OrbProbe(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_orb_destroy(void* o) { delete (OrbProbe*)o; }
int ref_orb_features_per_level(void* o, int l) { return ((OrbProbe*)o)->featuresPerLevel(l); }
int ref_orb_umax(void* o, int v) { return ((OrbProbe*)o)->umaxAt(v); }
float ref_orb_scale_factor(void* o, int l) { return ((OrbProbe*)o)->GetScaleFactors()[l]; }

/* ORBextractor::operator() (src/ORBextractor.cc:1043-1105) */
int ref_orb_extract(void* o, const uint8_t* img, int w, int h, size_t stride, orc_keypoint* kps, uint8_t* desc, int cap)
{
    ArenaScope arena;
    return guarded([&]() {
        cv::Mat image(h, w, CV_8UC1, (void*)img, stride), d;
        std::vector<cv::KeyPoint> k;
        (*(OrbProbe*)o)(image, cv::Mat(), k, d);
        if ((int)k.size() > cap) return -1;
        if (!k.empty()) {
            memcpy(kps, k.data(), k.size() * sizeof(cv::KeyPoint));
            for (int i = 0; i < d.rows; i++) memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
        }
        return (int)k.size();
    });
}

/* mvImagePyramid[level] after extract: size and a copy without the border */
int ref_orb_level(void* o, int level, int* w, int* h, uint8_t* out /* may be NULL */)
{
    const cv::Mat& m = ((OrbProbe*)o)->mvImagePyramid[level];
    *w = m.cols; *h = m.rows;
    if (out) for (int y = 0; y < m.rows; y++) memcpy(out + (size_t)y * m.cols, m.ptr(y), m.cols);
    return 0;
}

/* raw per-level keypoints of ComputeKeyPointsOctTree (:765-853) after extract is not kept by the reference; this runs the
 * pyramid + per-level detection again and returns level `level`'s retained keypoints (level coordinates, angle set) */
int ref_orb_level_keypoints(void* o, const uint8_t* img, int w, int h, size_t stride, int level, orc_keypoint* kps, int cap)
{
    ArenaScope arena;
    return guarded([&]() {
        OrbProbe* p = (OrbProbe*)o;
        cv::Mat image(h, w, CV_8UC1, (void*)img, stride);
        p->ComputePyramid(image);
        std::vector<std::vector<cv::KeyPoint> > all;
        p->ComputeKeyPointsOctTree(all);
        int n = (int)all[level].size();
        if (n > cap) return -1;
        if (n) memcpy(kps, all[level].data(), n * sizeof(cv::KeyPoint));
        return n;
    });
}

/* ORBextractor::DistributeOctTree (:539-763) alone; keys relative to (minX, minY) */
int ref_distribute_octree(void* o, const int* xs, const int* ys, const int* resp, int n, int minX, int maxX, int minY, int maxY, int N,
                          int* out_idx, int cap)
{
    ArenaScope arena;
    return guarded([&]() {
        std::vector<cv::KeyPoint> in(n);
        for (int i = 0; i < n; i++) { in[i] = cv::KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)resp[i]); in[i].class_id = i; }
        int level = 0;
        std::vector<cv::KeyPoint> out = ((OrbProbe*)o)->DistributeOctTree(in, minX, maxX, minY, maxY, N, level);
        if ((int)out.size() > cap) return -1;
        for (size_t i = 0; i < out.size(); i++) out_idx[i] = out[i].class_id;
        return (int)out.size();
    });
}

static LineProbe* make_line(const orc_line_params* p)
{
    return new LineProbe(p->nfeatures, p->nlevels, p->refine, p->scale, p->sigma_scale, p->quant, p->ang_th, p->log_eps, p->density_th,
                         p->n_bins, p->min_line_length, true);
}
int ref_line_features_per_level(const orc_line_params* p, int level)
{
    LineProbe* L = make_line(p);
    int v = L->featuresPerLevel(level);
    delete L;
    return v;
}

/* LSDDetectorC::detect(image, keylines, 2, nlevels, opts) (LSDDetector_custom.cpp:218-324) */
int ref_lsd_detect_keylines(const orc_line_params* p, const uint8_t* img, int w, int h, size_t stride, orc_keyline* kl, int cap)
{
    return guarded([&]() {
        cv::Mat image(h, w, CV_8UC1, (void*)img, stride);
        cv::Ptr<cv::line_descriptor::LSDDetectorC> lsd = cv::line_descriptor::LSDDetectorC::createLSDDetectorC();
        cv::line_descriptor::LSDDetectorC::LSDOptions o;
        o.refine = p->refine; o.scale = p->scale; o.sigma_scale = p->sigma_scale; o.quant = p->quant; o.ang_th = p->ang_th;
        o.log_eps = p->log_eps; o.density_th = p->density_th; o.n_bins = p->n_bins; o.min_length = p->min_line_length;
        std::vector<cv::line_descriptor::KeyLine> k;
        lsd->detect(image, k, 2, p->nlevels, o);
        if ((int)k.size() > cap) return -1;
        if (!k.empty()) memcpy(kl, k.data(), k.size() * sizeof(orc_keyline));
        return (int)k.size();
    });
}

/* BinaryDescriptor::compute (binary_descriptor_custom.cpp:524-687, 1026-1372); fdesc (n x 72 floats) optional */
int ref_lbd_compute(const uint8_t* img, int w, int h, size_t stride, const orc_keyline* kl, int n, uint8_t* desc, float* fdesc)
{
    return guarded([&]() {
        if (n == 0) return 0;
        cv::Mat image(h, w, CV_8UC1, (void*)img, stride);
        std::vector<cv::line_descriptor::KeyLine> k(n);
        memcpy(k.data(), kl, n * sizeof(orc_keyline));
        cv::Ptr<cv::line_descriptor::BinaryDescriptor> lbd = cv::line_descriptor::BinaryDescriptor::createBinaryDescriptor();
        cv::Mat d;
        if (desc) {
            lbd->compute(image, k, d, false);
            for (int i = 0; i < n; i++) memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
        }
        if (fdesc) {
            cv::Mat f;
            lbd->compute(image, k, f, true);
            for (int i = 0; i < n; i++) memcpy(fdesc + 72 * (size_t)i, f.ptr(i), 72 * sizeof(float));
        }
        return n;
    });
}

/* Lineextractor::ComputeLsdWithLbd (src/Lineextractor.cc:112-212) */
int ref_line_extract(const orc_line_params* p, const uint8_t* img, int w, int h, size_t stride, orc_keyline* kl, orc_keypoint* mid,
                     uint8_t* desc, int cap)
{
    return guarded([&]() {
        cv::Mat image(h, w, CV_8UC1, (void*)img, stride), d;
        std::unique_ptr<LineProbe> L(make_line(p));
        std::vector<cv::line_descriptor::KeyLine> k;
        std::vector<cv::KeyPoint> m;
        L->ComputeLsdWithLbd(image, k, m, d);
        if ((int)k.size() > cap) return -1;
        if (!k.empty()) {
            memcpy(kl, k.data(), k.size() * sizeof(orc_keyline));
            memcpy(mid, m.data(), m.size() * sizeof(orc_keypoint));
            for (int i = 0; i < d.rows; i++) memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
        }
        return (int)k.size();
    });
}

/* Lineextractor FLD branch: ComputeFldWithLbd (src/Lineextractor.cc:242-336, 413-980) */
typedef struct {
    int nfeatures, nlevels;
    double scale;
    int length_threshold;
    float distance_threshold;
    double canny_th1, canny_th2;
    int canny_aperture_size, do_merge;
} ref_fld_params;
int ref_fld_extract(const ref_fld_params* p, const uint8_t* img, int w, int h, size_t stride, orc_keyline* kl, orc_keypoint* mid,
                    uint8_t* desc, int cap)
{
    return guarded([&]() {
        cv::Mat view(h, w, CV_8UC1, (void*)img, stride), d;
        cv::Mat image = view.clone();
        LineProbe L(p->nfeatures, p->nlevels, p->scale, p->length_threshold, p->distance_threshold, p->canny_th1, p->canny_th2,
                    p->canny_aperture_size, p->do_merge != 0, false);
        std::vector<cv::line_descriptor::KeyLine> k;
        std::vector<cv::KeyPoint> m;
        L.ComputeFldWithLbd(image, k, m, d);
        if ((int)k.size() > cap) return -1;
        if (!k.empty()) {
            memcpy(kl, k.data(), k.size() * sizeof(orc_keyline));
            memcpy(mid, m.data(), m.size() * sizeof(orc_keypoint));
            for (int i = 0; i < d.rows; i++) memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
        }
        return (int)k.size();
    });
}
/* Lineextractor::detect (single-level FLD, :443-460): n x 4 floats */
int ref_fld_detect(const ref_fld_params* p, const uint8_t* img, int w, int h, size_t stride, float* lines, int cap)
{
    return guarded([&]() {
        cv::Mat view(h, w, CV_8UC1, (void*)img, stride);
        LineProbe L(p->nfeatures, p->nlevels, p->scale, p->length_threshold, p->distance_threshold, p->canny_th1, p->canny_th2,
                    p->canny_aperture_size, p->do_merge != 0, false);
        std::vector<cv::Vec4f> v;
        L.detect(view, v);
        if ((int)v.size() > cap) return -1;
        for (size_t i = 0; i < v.size(); i++) for (int j = 0; j < 4; j++) lines[4 * i + j] = v[i][j];
        return (int)v.size();
    });
}

/* the real std::sort of this toolchain with the comparator of Lineextractor.h:66-71, on (response, index) records */
void ref_std_sort_desc(const float* k, int n, int* p)
{
    struct Rec { float response; int idx; };
    struct ByResponse { inline bool operator()(const Rec& a, const Rec& b) { return (a.response > b.response); } };
    std::vector<Rec> v(n);
    for (int i = 0; i < n; i++) { v[i].response = k[i]; v[i].idx = i; }
    std::sort(v.begin(), v.end(), ByResponse());
    for (int i = 0; i < n; i++) p[i] = v[i].idx;
}

/* ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1656-1672), Linematcher::DescriptorDistance (src/Linematcher.cc:50-66) */
int ref_orb_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    cv::Mat A(1, 32, CV_8UC1, (void*)a), B(1, 32, CV_8UC1, (void*)b);
    return PL_SLAM::ORBmatcher::DescriptorDistance(A, B);
}
int ref_line_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    cv::Mat A(1, 32, CV_8UC1, (void*)a), B(1, 32, CV_8UC1, (void*)b);
    return PL_SLAM::Linematcher::DescriptorDistance(A, B);
}
/* Linematcher::matchNNR (src/Linematcher.cc:520-541) */
int ref_match_nnr(const uint8_t* q, int nq, const uint8_t* t, int nt, float nnr, int32_t* matches12)
{
    return guarded([&]() {
        cv::Mat Q(nq, 32, CV_8UC1, (void*)q), T(nt, 32, CV_8UC1, (void*)t);
        std::vector<int> m;
        int n = 0;
        PL_SLAM::Linematcher lm;
        lm.matchNNR(Q, T, nnr, m, n);
        for (int i = 0; i < nq; i++) matches12[i] = m[i];
        return n;
    });
}
/* ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1610-1651) on a histogram given by its bin sizes */
void ref_three_maxima(const int* sizes, int L, int* ind)
{
    std::vector<std::vector<int> > h(L);
    for (int i = 0; i < L; i++) h[i].resize(sizes[i]);
    int a = -1, b = -1, c = -1;
    PL_SLAM::ORBmatcher m;
    m.ComputeThreeMaxima(h.data(), L, a, b, c);
    ind[0] = a; ind[1] = b; ind[2] = c;
}

/* ---- the reference's matcher entry points on mock SLAM objects (tests/shim/mock_slam.hpp) ---- */
static void fill_frame_points(PL_SLAM::Frame& F, const orc_keypoint* k, const uint8_t* d, int n, const float* bounds)
{
    F.mvKeysUn.resize(n);
    if (n) memcpy(F.mvKeysUn.data(), k, n * sizeof(cv::KeyPoint));
    F.mDescriptors = cv::Mat(n, 32, CV_8UC1, (void*)d);
    F.SetBoundsAndAssign(bounds[0], bounds[1], bounds[2], bounds[3]);
}
/* ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:406-521); prev = n1 x 2 floats (in / out) */
int ref_orb_search_for_initialization(const orc_keypoint* k1, const uint8_t* d1, int n1, const orc_keypoint* k2, const uint8_t* d2, int n2,
                                      const float* bounds, float* prev, int window, float nnratio, int check_ori, int32_t* matches12)
{
    return guarded([&]() {
        PL_SLAM::Frame F1, F2;
        fill_frame_points(F1, k1, d1, n1, bounds);
        fill_frame_points(F2, k2, d2, n2, bounds);
        std::vector<cv::Point2f> pm(n1);
        for (int i = 0; i < n1; i++) pm[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);
        std::vector<int> m;
        PL_SLAM::ORBmatcher om(nnratio, check_ori != 0);
        int n = om.SearchForInitialization(F1, F2, pm, m, window);
        for (int i = 0; i < n1; i++) { matches12[i] = m[i]; prev[2 * i] = pm[i].x; prev[2 * i + 1] = pm[i].y; }
        return n;
    });
}
/* Linematcher::SearchByKNN (src/Linematcher.cc:437-517).  ml_state per key-frame line: 0 no MapLine, 1 good, 2 bad;
 * out_ml[i2] = index of the key-frame line whose MapLine was assigned to frame line i2, or -1 */
int ref_line_search_by_knn(const uint8_t* dkf, int nkf, const uint8_t* df, int nf, const uint8_t* ml_state, const float* ml_len,
                           const float* f_linelen, float nnr, int checklen, float lengtherr, int32_t* out_ml)
{
    return guarded([&]() {
        std::vector<PL_SLAM::MapLine> mls(nkf);
        PL_SLAM::KeyFrame kf;
        PL_SLAM::Frame F;
        kf.mDescriptorLines = cv::Mat(nkf, 32, CV_8UC1, (void*)dkf);
        kf.mvpMapLines.resize(nkf);
        for (int i = 0; i < nkf; i++) {
            mls[i].mbBad = ml_state[i] == 2; mls[i].mfLen = ml_len[i];
            kf.mvpMapLines[i] = ml_state[i] ? &mls[i] : (PL_SLAM::MapLine*)NULL;
        }
        F.NL = nf;
        F.mDescriptorLines = cv::Mat(nf, 32, CV_8UC1, (void*)df);
        F.mvLinesUn.resize(nf);
        for (int i = 0; i < nf; i++) F.mvLinesUn[i].lineLength = f_linelen[i];
        std::vector<PL_SLAM::MapLine*> out;
        PL_SLAM::Linematcher lm(nnr, true, checklen != 0, lengtherr);
        int n = lm.SearchByKNN(&kf, F, out);
        for (int i = 0; i < nf; i++) out_ml[i] = out[i] ? (int)(out[i] - mls.data()) : -1;
        return n;
    });
}
/* Linematcher::SearchForTriangulation (src/Linematcher.cc:804-879).  cam = fx, fy, cx, cy of key frame 2; pose = Ow1[3], R2w[9],
 * t2w[3]; pairs = up to n1 (i1, i2); returns the number of pairs (the reference's own return value is decremented on
 * out-of-bounds reads and is not reported) */
int ref_line_search_for_triangulation(const uint8_t* d1, int n1, const uint8_t* d2, int n2, const uint8_t* has_ml1, const uint8_t* has_ml2,
                                      const orc_keypoint* mid1, const orc_keypoint* mid2, const float* scale, const float* sigma2, int nlev,
                                      const float* cam, const float* pose, const float* F12, float nnr, int32_t* pairs)
{
    return guarded([&]() {
        PL_SLAM::MapLine dummy;
        PL_SLAM::KeyFrame k1, k2;
        k1.mDescriptorLines = cv::Mat(n1, 32, CV_8UC1, (void*)d1);
        k2.mDescriptorLines = cv::Mat(n2, 32, CV_8UC1, (void*)d2);
        k1.mvpMapLines.resize(n1); k2.mvpMapLines.resize(n2);
        for (int i = 0; i < n1; i++) k1.mvpMapLines[i] = has_ml1[i] ? &dummy : (PL_SLAM::MapLine*)NULL;
        for (int i = 0; i < n2; i++) k2.mvpMapLines[i] = has_ml2[i] ? &dummy : (PL_SLAM::MapLine*)NULL;
        k1.mvMidPointsUn.resize(n1); k2.mvMidPointsUn.resize(n2);
        if (n1) memcpy(k1.mvMidPointsUn.data(), mid1, n1 * sizeof(cv::KeyPoint));
        if (n2) memcpy(k2.mvMidPointsUn.data(), mid2, n2 * sizeof(cv::KeyPoint));
        for (PL_SLAM::KeyFrame* k : {&k1, &k2}) {
            k->mvScaleFactorsLines.assign(scale, scale + nlev);
            k->mvLevelSigma2Lines.assign(sigma2, sigma2 + nlev);
            k->fx = cam[0]; k->fy = cam[1]; k->cx = cam[2]; k->cy = cam[3];
        }
        float ow[3], r[9], t[3], f[9];
        memcpy(ow, pose, 12); memcpy(r, pose + 3, 36); memcpy(t, pose + 12, 12); memcpy(f, F12, 36);
        k1.mOw = cv::Mat(3, 1, CV_32FC1, ow).clone();
        k2.mRcw = cv::Mat(3, 3, CV_32FC1, r).clone();
        k2.mtcw = cv::Mat(3, 1, CV_32FC1, t).clone();
        cv::Mat Fm = cv::Mat(3, 3, CV_32FC1, f).clone();
        std::vector<std::pair<size_t, size_t> > vp;
        PL_SLAM::Linematcher lm(nnr, true, true, 0.1f);
        lm.SearchForTriangulation(&k1, &k2, Fm, vp);
        for (size_t i = 0; i < vp.size(); i++) { pairs[2 * i] = (int)vp[i].first; pairs[2 * i + 1] = (int)vp[i].second; }
        return (int)vp.size();
    });
}

} // extern "C"

/* cv::BFMatcher model for matchNNR: Hamming, k = 2, ordering (distance, trainIdx) -- orc_knn2 */
namespace cv {
Ptr<BFMatcher> BFMatcher::create(int normType, bool crossCheck)
{
    CV_Assert(normType == NORM_HAMMING && !crossCheck);
    return Ptr<BFMatcher>(new BFMatcher());
}
void BFMatcher::knnMatch(InputArray queryDescriptors, InputArray trainDescriptors, std::vector<std::vector<DMatch> >& matches, int k)
{
    CV_Assert(k == 2);
    Mat q = queryDescriptors.getMat(), t = trainDescriptors.getMat();
    int nq = q.rows, nt = t.rows;
    matches.assign(nq, std::vector<DMatch>());
    if (nq == 0 || nt == 0) return;
    CV_Assert(q.isContinuous() && t.isContinuous() && q.cols == 32 && t.cols == 32);
    std::vector<int32_t> idx(2 * (size_t)nq), dist(2 * (size_t)nq);
    orc_knn2(q.data, nq, t.data, nt, idx.data(), dist.data());
    for (int i = 0; i < nq; i++)
        for (int j = 0; j < 2; j++)
            if (idx[2 * i + j] >= 0) matches[i].push_back(DMatch(i, idx[2 * i + j], (float)dist[2 * i + j]));
}
}
