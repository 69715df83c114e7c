/*
 * plf_oracle.h -- CPU oracle for the point-line feature front-end.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (spl_slam_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the checker
 * or the timed CPU baseline.
 *
 * It restates, in plain C, the reference's CPU algorithm for the hot path
 *  including the un-vendored OpenCV primitives it
 * calls.  Parity status: the reference ships no tests or golden vectors
 * ("parity unpinned" upstream); every primitive here is pinned bit-for-bit against
 * cv2 4.13.0 by tests/test_oracle_vs_cv2.py, and the glue follows the reference
 * source lines cited at each function.
 */
#ifndef PLF_ORACLE_H
#define PLF_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cv::KeyPoint layout (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint;

/* line_descriptor::KeyLine layout (68 bytes), descriptor_custom.hpp:105-174 */
typedef struct {
    float angle;
    int32_t class_id;
    int32_t octave;
    float pt_x, pt_y;
    float response;
    float size;
    float startPointX, startPointY, endPointX, endPointY;
    float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength;
    int32_t numOfPixels;
} orc_keyline;

/* ---- primitives (OpenCV behaviour) ---- */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                          uint8_t* dst, int dw, int dh, size_t dstride);
void orc_resize_linear_exact_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                                uint8_t* dst, int dw, int dh, size_t dstride, double fx, double fy);
void orc_border_reflect101_u8(const uint8_t* src, int w, int h, size_t sstride,
                              uint8_t* dst, int border, size_t dstride);
int  orc_gauss_kernel_q8(int ksize, double sigma, int* q /* ksize */);
void orc_gauss_blur_u8(const uint8_t* src, int w, int h, size_t sstride,
                       uint8_t* dst, size_t dstride, int ksize, double sigma);
void orc_pyrdown_u8(const uint8_t* src, int w, int h, size_t sstride,
                    uint8_t* dst, size_t dstride); /* dst is (w/2)x(h/2) */
void orc_sobel3_s16(const uint8_t* src, int w, int h, size_t sstride,
                    int16_t* dx, int16_t* dy); /* dense w*h */
float orc_fast_atan2(float y, float x);
/* FAST-9/16 + NMS on one cell; returns count; xs/ys/score sized cap */
int  orc_fast9(const uint8_t* img, int w, int h, size_t stride, int th,
               int* xs, int* ys, int* score, int cap);

/* FLD branch primitives (orc_fld.c): cv::Canny (aperture 3, L1 gradient) -> dense w*h 255/0; cv::fitLine DIST_L2 on int points */
void orc_canny_u8(const uint8_t* src, int w, int h, size_t stride, double th1, double th2, uint8_t* edges);
void orc_fit_line_l2(const int32_t* pts, int count, float* line /* vx, vy, x0, y0 */);

/* ---- ORB extractor (src/ORBextractor.cc) ---- */
typedef struct orc_orb orc_orb;
orc_orb* orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh);
void orc_orb_destroy(orc_orb*);
int  orc_orb_features_per_level(const orc_orb*, int level);
float orc_orb_scale_factor(const orc_orb*, int level);
int  orc_orb_umax(const orc_orb*, int v);
/* full operator(): returns number of keypoints (<= cap), or -1 if cap too small */
int  orc_orb_extract(orc_orb*, const uint8_t* img, int w, int h, size_t stride,
                     orc_keypoint* kps, uint8_t* desc, int cap);
/* intermediate access after extract (for stage-by-stage parity tests) */
int  orc_orb_level_size(const orc_orb*, int level, int* w, int* h);
const uint8_t* orc_orb_level_image(const orc_orb*, int level, size_t* stride);   /* ROI inside border */
const uint8_t* orc_orb_level_blurred(const orc_orb*, int level, size_t* stride); /* NULL if level had no kps */
int  orc_orb_level_raw_count(const orc_orb*, int level);
/* raw FAST keypoints of a level in distribute order: x,y (relative to minBorder), response */
void orc_orb_level_raw(const orc_orb*, int level, int* xs, int* ys, int* resp);
int  orc_orb_level_kept_count(const orc_orb*, int level);

/* Frame::ComputeStereoMatches (src/Frame.cc:881-1055); oL/oR hold the two images' pyramids (after extract) */
void orc_stereo_match(const orc_orb* oL, const orc_orb* oR, const orc_keypoint* kL, const uint8_t* dL, int nL,
                      const orc_keypoint* kR, const uint8_t* dR, int nR, float mb, float mbf,
                      float* uRight, float* depth);

/* Frame::AssignFeaturesToGrid[Lines] + GetFeaturesInArea[Lines] (src/Frame.cc:365-399, :562-722): CSR candidate lists
 * in the reference's order; kls == NULL for the point grid.  Returns the total candidate count. */
typedef struct { int32_t cols, rows; float min_x, min_y, inv_w, inv_h; } orc_grid_params;
int orc_grid_candidates(const orc_keypoint* kps, const orc_keyline* kls, int n, const orc_grid_params* g, const float* qx,
                        const float* qy, const float* qr, const int32_t* qminl, const int32_t* qmaxl, int nq, int32_t* cand_off,
                        int32_t* cand_idx, int cand_cap);

/* DBoW2 tree descent per feature (TemplatedVocabulary.h:1218-1258): word id, weight, node at level L - levelsup */
void orc_bow_transform(int L, int nnodes, const int32_t* parent, const uint8_t* ndesc, const double* nweight, const uint8_t* is_leaf,
                       const uint8_t* feat, int n, int levelsup, int32_t* word, double* weight, int32_t* node);

/* cv::undistortPoints(pts, pts, K, D, Mat(), K) as used by Frame::UndistortKeyPoints / UndistortKeyLines (src/Frame.cc:733-826) */
void orc_undistort_points(const float* cam, const float* kd, int nk, const float* in, int n, float* out);

/* DistributeOctTree alone (ORBextractor.cc:539-763); keys relative to (minX,minY).
 * out_idx receives indices into the input arrays in final list order; returns count. */
int  orc_distribute_octree(const int* xs, const int* ys, const int* resp, int n,
                           int minX, int maxX, int minY, int maxY, int N, int* out_idx, int cap);

/* ---- LSD (cv::LineSegmentDetector refine=0 behaviour) ---- */
/* returns number of lines; lines = n x 4 float (x1,y1,x2,y2) */
int  orc_lsd_detect(const uint8_t* img, int w, int h, size_t stride,
                    double scale, double sigma_scale, double quant, double ang_th, int n_bins,
                    float* lines, int cap);

/* ---- line extractor (src/Lineextractor.cc:32-212 + LSDDetector_custom.cpp + LBD) ---- */
typedef struct {
    int nfeatures, nlevels, refine;
    double scale, sigma_scale, quant, ang_th, log_eps, density_th;
    int n_bins;
    double min_line_length;
} orc_line_params;
int  orc_line_features_per_level(const orc_line_params* p, int level);
/* LSDDetectorC::detect(image, kl, 2, nlevels, opts): returns count */
int  orc_lsd_detect_keylines(const orc_line_params* p, const uint8_t* img, int w, int h, size_t stride,
                             orc_keyline* kl, int cap);
/* BinaryDescriptor::compute: desc = n x 32 u8, fdesc (optional) = n x 72 float */
void orc_lbd_compute(const uint8_t* img, int w, int h, size_t stride,
                     const orc_keyline* kl, int n, uint8_t* desc, float* fdesc);
/* ComputeLsdWithLbd: returns number of lines kept */
int  orc_line_extract(const orc_line_params* p, const uint8_t* img, int w, int h, size_t stride,
                      orc_keyline* kl, orc_keypoint* mid, uint8_t* desc, int cap);

/* libstdc++'s std::sort with comparator k[a] > k[b], as a permutation of 0..n-1 (unstable: the exact introsort sequence) */
void orc_std_sort_desc(const float* k, int n, int* p);

/* ---- matching (src/Linematcher.cc:50-66, 520-541; src/ORBmatcher.cc:1656-1672) ---- */
int  orc_descriptor_distance(const uint8_t* a, const uint8_t* b);
/* knnMatch k=2 semantics: idx/dist are nq x 2; missing entries -1 */
void orc_knn2(const uint8_t* q, int nq, const uint8_t* t, long nt, int32_t* idx, int32_t* dist);
/* matchNNR: matches12[q] = trainIdx or -1; returns nmatches. nt<2 -> all -1 (defined behaviour) */
/* candidate-list top-2, reference update rule (src/ORBmatcher.cc:430-456) */
void orc_hamming_candidates(const uint8_t* q, int nq, const uint8_t* t, const int32_t* off, const int32_t* cidx,
                            int32_t* bidx, int32_t* bdist, int32_t* cdist);
int  orc_match_nnr(const uint8_t* q, int nq, const uint8_t* t, long nt, float nnr, int32_t* matches12);

#ifdef __cplusplus
}
#endif
#endif
