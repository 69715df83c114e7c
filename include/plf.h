/*
 * plf.h -- C ABI of the B200-native point-line feature front-end (libplf.so).
 *
 * Drop-in boundary for the hot path of Hero941215/spl-slam (reference paths are
 * relative to the reference root).  Every entry point names the reference interface
 * it replaces.  Plain pointers and sizes only; no exceptions cross this boundary;
 * every function returns a plf_status and plf_last_error() explains failures.
 * There is NO CPU fallback behind this ABI: without a usable CUDA device every call
 * fails with PLF_ERR_CUDA.
 *
 * Buffers named host_* are host memory (pageable or pinned); buffers named dev_* are
 * device pointers on the context's device.  All images are 8-bit single channel
 * (CV_8UC1) with a row stride in bytes.
 */
#ifndef PLF_H
#define PLF_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PLF_OK = 0,
    PLF_ERR_INVALID = 1,  /* bad argument */
    PLF_ERR_CUDA = 2,     /* CUDA runtime error / no device */
    PLF_ERR_CAPACITY = 3, /* an output or internal buffer was too small */
    PLF_ERR_STATE = 4     /* call order error (e.g. pyramid requested before extract) */
} plf_status;

/* cv::KeyPoint memory layout (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} plf_keypoint;

/* cv::line_descriptor::KeyLine memory layout (68 bytes),
 * Thirdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp:105-174 */
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
} plf_keyline;

typedef struct plf_ctx plf_ctx;
typedef struct plf_orb plf_orb;
typedef struct plf_line plf_line;
typedef struct plf_vocab plf_vocab;
typedef struct plf_comm plf_comm;

/* ---- context: one per host thread / GPU stream.  The reference calls its extractors and
 * matchNNR from concurrent std::threads (src/Frame.cc:116-119, :301-304;
 * src/Linematcher.cc:454-457); give each such thread its own context. ---- */
plf_status plf_ctx_create(int device, plf_ctx** out);
/* the same with a stream priority: > 0 latency-critical (e.g. the line extractor, whose region-growing chains set the
 * frame time), < 0 filler work (e.g. ORB extraction of the next batch), 0 default */
plf_status plf_ctx_create_prio(int device, int priority, plf_ctx** out);
void plf_ctx_destroy(plf_ctx* ctx);
const char* plf_last_error(const plf_ctx* ctx);
plf_status plf_ctx_synchronize(plf_ctx* ctx);
/* cudaStream_t of the context (as void*) so callers can order their own work / events on it */
void* plf_ctx_stream(plf_ctx* ctx);
/* device-side stopwatch on the context's stream (CUDA events) */
plf_status plf_timer_start(plf_ctx* ctx);
plf_status plf_timer_stop(plf_ctx* ctx, float* elapsed_ms);
/* number of kernels this context has launched since creation (bench "gpu_launches") */
uint64_t plf_ctx_launch_count(const plf_ctx* ctx);
/* make ctx wait, on the device, for everything queued so far on `other` (the reference runs the ORB and the
 * line extractor in two threads, src/Frame.cc:301-304; here they are two contexts / streams) */
plf_status plf_ctx_wait(plf_ctx* ctx, plf_ctx* other);
/* optional per-kernel timing: every launch is bracketed by CUDA events on the context stream and summed per
 * kernel name; plf_profile_report writes "name total_ms launches" lines */
plf_status plf_profile_enable(plf_ctx* ctx, int on);
plf_status plf_profile_report(plf_ctx* ctx, char* buf, size_t bufsize);
/* diagnostic: "name start_ms end_ms" lines for the launches recorded since profiling was enabled, relative to the
 * last plf_timer_start of `ref` (same device); call before plf_profile_report */
plf_status plf_profile_timeline(plf_ctx* ctx, plf_ctx* ref, char* buf, size_t bufsize);

/* One upload for several consumers: the reference hands the same image to its ORB thread and its line thread
 * (src/Frame.cc:301-304).  plf_upload queues the host -> device copy on `ctx`'s stream (asynchronous for pinned memory);
 * a consumer context then calls plf_ctx_wait(consumer, ctx) and one of the *_from_device entry points below (device images,
 * host results).  plf_device_malloc / plf_device_free are for hosts that do not link the CUDA runtime themselves. */
plf_status plf_upload(plf_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes);
plf_status plf_device_malloc(plf_ctx* ctx, size_t bytes, void** out);
void plf_device_free(plf_ctx* ctx, void* dev_ptr);

/* ---- ORB extractor: replaces PL_SLAM::ORBextractor
 * (include/ORBextractor.h:45-113, src/ORBextractor.cc:410-470, :1043-1132) ---- */
typedef struct {
    int nfeatures;      /* ORBextractor.nFeatures */
    float scale_factor; /* ORBextractor.scaleFactor */
    int nlevels;        /* ORBextractor.nLevels */
    int ini_th_fast;    /* ORBextractor.iniThFAST */
    int min_th_fast;    /* ORBextractor.minThFAST */
} plf_orb_params;

plf_status plf_orb_create(plf_ctx* ctx, const plf_orb_params* p, plf_orb** out);
void plf_orb_destroy(plf_orb* orb);
/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:63-83) and mnFeaturesPerLevel; each array has nlevels entries, any may be NULL */
plf_status plf_orb_tables(const plf_orb* orb, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                          int32_t* features_per_level);
/* upper bound on keypoints one frame can return (the reference may exceed nfeatures by a few per level) */
int plf_orb_max_keypoints(const plf_orb* orb);
/* ORBextractor::operator()(image, mask, keypoints, descriptors) -- src/ORBextractor.cc:1043-1105.
 * host_kps/host_desc (cap x 28 B, cap x 32 B) receive *n_out entries.  An empty image
 * (w<=0 || h<=0 || !host_img) returns PLF_OK with *n_out = 0, like the reference's silent return. */
plf_status plf_orb_extract(plf_orb* orb, const uint8_t* host_img, int w, int h, size_t stride,
                           plf_keypoint* host_kps, uint8_t* host_desc, int cap, int* n_out);
/* Batched form: nframes images of identical size, frame f at host_imgs + f*frame_stride.
 * Outputs for frame f start at host_kps + f*cap and host_desc + f*cap*32; n_out[f] entries valid.
 * Host<->device copies are part of the call. */
plf_status plf_orb_extract_batch(plf_orb* orb, const uint8_t* host_imgs, int nframes, int w, int h, size_t stride,
                                 size_t frame_stride, plf_keypoint* host_kps, uint8_t* host_desc, int cap,
                                 int32_t* n_out);
/* Same, inputs and outputs resident in device memory (no copies; asynchronous on the context stream;
 * dev_n_out is an int32[nframes] device array).  The call cannot return a capacity error for work that is still queued, so
 * the *_device entry points report it per frame in dev_n_out[f] < 0 (the host-buffer entry points turn the same codes into
 * PLF_ERR_CAPACITY): ORB  -1 = a level's raw FAST list overflowed, -2 = more than cap keypoints;
 * lines -1 = more than the per-octave detection capacity, -2 = more than cap keylines, -3 = the LSD region buffer of an
 * octave overflowed somewhere in this batch (regions were dropped: no frame of the batch is valid).
 * dev_imgs must stay alive and unchanged until plf_stereo_match* / plf_orb_pyramid_level for this batch have run: level 0 of
 * the pyramid IS the caller's buffer (it is not copied). */
/* images already on the device (plf_upload), results to host buffers; dev_imgs must stay valid until the call returns */
plf_status plf_orb_extract_batch_from_device(plf_orb* orb, const uint8_t* dev_imgs, int nframes, int w, int h, size_t stride,
                                             size_t frame_stride, plf_keypoint* host_kps, uint8_t* host_desc, int cap, int32_t* n_out);
plf_status plf_orb_extract_batch_device(plf_orb* orb, const uint8_t* dev_imgs, int nframes, int w, int h,
                                        size_t stride, size_t frame_stride, plf_keypoint* dev_kps,
                                        uint8_t* dev_desc, int cap, int32_t* dev_n_out);
/* mvImagePyramid[level] of frame `frame` of the last extract call (include/ORBextractor.h:85; read by
 * Frame::ComputeStereoMatches, src/Frame.cc:967-1007).  Copies the level image (without the 19-px
 * border, which nothing on the hot path reads) to host_dst; host_dst may be NULL to query the size. */
plf_status plf_orb_pyramid_level(plf_orb* orb, int frame, int level, uint8_t* host_dst, size_t dst_stride,
                                 int* w, int* h);
/* stage-level access for parity tests: blurred level and raw FAST keys (x, y relative to the 16-px
 * border origin, response, order key) of one frame/level of the last call */
plf_status plf_orb_debug_blurred(plf_orb* orb, int frame, int level, uint8_t* host_dst, size_t dst_stride);
plf_status plf_orb_debug_raw_keys(plf_orb* orb, int frame, int level, int32_t* xs, int32_t* ys, int32_t* resp,
                                  int cap, int* n_out);
/* ORBextractor::DistributeOctTree alone (src/ORBextractor.cc:539-763) on host-provided keys of one level:
 * out_idx receives indices into xs/ys/resp in final list order */
plf_status plf_orb_distribute_octree(plf_ctx* ctx, const int32_t* xs, const int32_t* ys, const int32_t* resp, int n,
                                     int minX, int maxX, int minY, int maxY, int N, int32_t* out_idx, int cap,
                                     int* n_out);

/* ---- line extractor: replaces PL_SLAM::Lineextractor (LSD branch) with the vendored
 * LSDDetectorC::detect + BinaryDescriptor::compute beneath it
 * (include/Lineextractor.h:49-61, src/Lineextractor.cc:32-67, :112-212;
 * Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:218-324;
 * Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:524-687, :1026-1372) ---- */
typedef struct {
    int nfeatures;          /* Lineextractor.nFeatures */
    int nlevels;            /* Lineextractor.nLevels (octaves, ratio 2) */
    int refine;             /* must be 0 (LSD_REFINE_NONE): the only mode the shipped configs use */
    double scale;           /* LSD scale (also the feature-split factor, Lineextractor.cc:56) */
    double sigma_scale;
    double quant;
    double ang_th;
    double log_eps;         /* unused with refine = 0 */
    double density_th;      /* unused with refine = 0 */
    int n_bins;
    double min_line_length; /* LSDOptions.min_length */
} plf_line_params;

plf_status plf_line_create(plf_ctx* ctx, const plf_line_params* p, plf_line** out);
void plf_line_destroy(plf_line* le);
plf_status plf_line_tables(const plf_line* le, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                           int32_t* features_per_level);
int plf_line_max_keylines(const plf_line* le);
/* Lineextractor::ComputeLsdWithLbd(image, keyLines, keypoints, descriptors) -- src/Lineextractor.cc:112-212.
 * host_mid receives the mid-point KeyPoints (pt, octave; other fields cv::KeyPoint() defaults). */
plf_status plf_line_extract(plf_line* le, const uint8_t* host_img, int w, int h, size_t stride,
                            plf_keyline* host_kl, plf_keypoint* host_mid, uint8_t* host_desc, int cap, int* n_out);
plf_status plf_line_extract_batch(plf_line* le, const uint8_t* host_imgs, int nframes, int w, int h, size_t stride,
                                  size_t frame_stride, plf_keyline* host_kl, plf_keypoint* host_mid,
                                  uint8_t* host_desc, int cap, int32_t* n_out);
plf_status plf_line_extract_batch_from_device(plf_line* le, const uint8_t* dev_imgs, int nframes, int w, int h, size_t stride,
                                              size_t frame_stride, plf_keyline* host_kl, plf_keypoint* host_mid, uint8_t* host_desc,
                                              int cap, int32_t* n_out);
plf_status plf_line_extract_batch_device(plf_line* le, const uint8_t* dev_imgs, int nframes, int w, int h,
                                         size_t stride, size_t frame_stride, plf_keyline* dev_kl,
                                         plf_keypoint* dev_mid, uint8_t* dev_desc, int cap, int32_t* dev_n_out);
/* LSDDetectorC::detect(image, keylines, 2, nlevels, opts) alone (LSDDetector_custom.cpp:218-324) */
plf_status plf_lsd_detect(plf_line* le, const uint8_t* host_img, int w, int h, size_t stride,
                          plf_keyline* host_kl, int cap, int* n_out);
/* BinaryDescriptor::compute(image, keylines, descriptors, returnFloatDescr) alone
 * (binary_descriptor_custom.cpp:524-687); host_fdesc (n x 72 float) may be NULL */
plf_status plf_lbd_compute(plf_line* le, const uint8_t* host_img, int w, int h, size_t stride,
                           const plf_keyline* host_kl, int n, uint8_t* host_desc, float* host_fdesc);

/* ---- matching: replaces ORBmatcher::DescriptorDistance / Linematcher::DescriptorDistance
 * (src/ORBmatcher.cc:1656-1672, src/Linematcher.cc:50-66), cv::BFMatcher(NORM_HAMMING).knnMatch(k=2)
 * as used by Linematcher::matchNNR (src/Linematcher.cc:520-541) and the mutual-consistency step of
 * Linematcher::SearchByKNN / SearchForTriangulation (:454-471, :825-839).  Descriptors are n x 32 bytes. ---- */
/* ---- FLD line extractor: replaces PL_SLAM::Lineextractor's FLD branch (System.usingLsdFeature: 0): the FLD constructor
 * (include/Lineextractor.h:55-57, src/Lineextractor.cc:69-110), ComputeFldWithLbd (:242-336), ComputePyramid /
 * detectFldWithPyramid / detect (:413-460) and the detector itself (lineDetection, getPointChain, extractSegments,
 * incidentPoint, additionalOperationsOnSegment, :546-905) over cv::Canny and cv::fitLine; descriptors by the same LBD path as
 * the LSD branch.  do_merge (off in every shipped config) and Canny apertures other than 3 are rejected at create. ---- */
typedef struct {
    int nfeatures;             /* Lineextractor.nFeatures */
    int nlevels;               /* Lineextractor.nLevels (pyrDown pyramid) */
    double scale;              /* feature-split factor and the (reference's) octave -> image coordinate factor */
    int length_threshold;      /* Lineextractor.threshold_length */
    float distance_threshold;  /* Lineextractor.threshold_dist */
    double canny_th1, canny_th2;
    int canny_aperture_size;   /* must be 3 */
    int do_merge;              /* must be 0 */
} plf_fld_params;
typedef struct plf_fld plf_fld;
plf_status plf_fld_create(plf_ctx* ctx, const plf_fld_params* p, plf_fld** out);
void plf_fld_destroy(plf_fld* fld);
plf_status plf_fld_features_per_level(const plf_fld* fld, int32_t* per_level);
/* Lineextractor::detect (single level): n x 4 floats (x1, y1, x2, y2) in detection order */
plf_status plf_fld_detect(plf_fld* fld, const uint8_t* host_img, int w, int h, size_t stride, float* host_lines, int cap, int* n_out);
/* ComputeFldWithLbd: KeyLines (68 B), mid-point KeyPoints, N x 32 LBD descriptors */
plf_status plf_fld_extract(plf_fld* fld, const uint8_t* host_img, int w, int h, size_t stride, plf_keyline* host_kl,
                           plf_keypoint* host_mid, uint8_t* host_desc, int cap, int* n_out);

/* pairwise distances dist[i] = Hamming(a[i], b[i]) */
plf_status plf_descriptor_distance(plf_ctx* ctx, const uint8_t* host_a, const uint8_t* host_b, int n, int32_t* host_dist);
/* knnMatch(k=2): idx/dist are nq x 2; ascending distance, ties -> lowest train index; -1 where the
 * train set has fewer than 2 rows.  train_index_base is added to every returned index (shards). */
plf_status plf_hamming_knn2(plf_ctx* ctx, const uint8_t* host_q, int nq, const uint8_t* host_t, int64_t nt,
                            int32_t* host_idx, int32_t* host_dist);
plf_status plf_hamming_knn2_device(plf_ctx* ctx, const uint8_t* dev_q, int nq, const uint8_t* dev_t, int64_t nt,
                                   int64_t train_index_base, int32_t* dev_idx, int32_t* dev_dist);
/* merge nshards partial top-2 tables (each nq x 2, global indices) into one; exact same ordering rule.
 * dev_idx_parts/dev_dist_parts are [nshards][nq][2] contiguous (e.g. an NCCL all-gather result). */
plf_status plf_knn2_merge_device(plf_ctx* ctx, const int32_t* dev_idx_parts, const int32_t* dev_dist_parts,
                                 int nshards, int nq, int32_t* dev_idx, int32_t* dev_dist);
/* Linematcher::matchNNR: matches12[q] = train index if d0 < d1 * nnr (float compare) else -1.
 * Fewer than 2 train rows is undefined behaviour in the reference; here: no match. */
plf_status plf_match_nnr(plf_ctx* ctx, const uint8_t* host_q, int nq, const uint8_t* host_t, int64_t nt, float nnr,
                         int32_t* host_matches12, int* nmatches);
plf_status plf_nnr_from_knn2_device(plf_ctx* ctx, const int32_t* dev_idx, const int32_t* dev_dist, int nq, float nnr,
                                    int32_t* dev_matches12, int32_t* dev_nmatches);
/* both directions + mutual-consistency filter of SearchByKNN (:454-471): matches12[i1] = i2 kept only if
 * matches21[i2] == i1.  (The reference indexes matches_21[-1] when i2 == -1; defined here as "no match".) */
plf_status plf_match_nnr_mutual(plf_ctx* ctx, const uint8_t* host_d1, int n1, const uint8_t* host_d2, int n2,
                                float nnr, int32_t* host_matches12, int* nmatches);

/* ---- multi-GPU matching (BASELINE config 5): one process per GPU, the TRAIN set row-sharded across the ranks
 * (row j -> shard floor(j * world / nt)), queries replicated.  Replaces the same knnMatch / matchNNR calls as above when the
 * train set is spread over the GPUs of a node.  A communicator wraps one NCCL communicator (bound at run time from
 * libnccl.so.2): rank 0 calls plf_comm_unique_id and ships the 128 bytes to the other ranks by any means (MPI,
 * torch.distributed, a file); every rank then calls plf_comm_create (collective).  plf_hamming_knn2_sharded_device =
 * local top-2 with global indices + ONE ncclAllGather of 16 B per query per rank + merge by (distance, index), all queued
 * on the context stream; the result (identical on every rank) equals the single-GPU table, ties included. */
plf_status plf_comm_unique_id(uint8_t id[128]);
plf_status plf_comm_create(plf_ctx* ctx, const uint8_t id[128], int rank, int world, plf_comm** out);
void plf_comm_destroy(plf_comm* comm);
int plf_comm_rank(const plf_comm* comm);
int plf_comm_world(const plf_comm* comm);
plf_status plf_hamming_knn2_sharded_device(plf_ctx* ctx, plf_comm* comm, const uint8_t* dev_q, int nq, const uint8_t* dev_t_local,
                                           int64_t nt_local, int64_t train_index_base, int32_t* dev_idx, int32_t* dev_dist);
plf_status plf_match_nnr_sharded_device(plf_ctx* ctx, plf_comm* comm, const uint8_t* dev_q, int nq, const uint8_t* dev_t_local,
                                        int64_t nt_local, int64_t train_index_base, float nnr, int32_t* dev_idx, int32_t* dev_dist,
                                        int32_t* dev_matches12, int32_t* dev_nmatches);

/* Candidate-list matching: the data-parallel core of ORBmatcher::SearchForInitialization / SearchByProjection /
 * SearchByBoW and Linematcher::SearchForInitialization / SearchByProjection (src/ORBmatcher.cc:45-129, :406-521;
 * src/Linematcher.cc:146-286, :544-646).  Query q scans its candidate train rows cand_idx[cand_off[q] ..
 * cand_off[q+1]) in list order with the reference's update rule (`dist < bestDist` strict, else
 * `dist < bestDist2`): best[q] = {first minimum, second entry in (distance, list position) order} as train
 * indices (-1 where the list is shorter), best_dist likewise.  cand_dist (may be NULL) receives every
 * candidate's distance so the order-dependent greedy steps (vMatchedDistance, rotation histogram) can stay on
 * the host.  cand_off has nq + 1 entries. */
plf_status plf_hamming_candidates(plf_ctx* ctx, const uint8_t* host_q, int nq, const uint8_t* host_t, int nt,
                                  const int32_t* host_cand_off, const int32_t* host_cand_idx,
                                  int32_t* host_best_idx, int32_t* host_best_dist, int32_t* host_cand_dist);
plf_status plf_hamming_candidates_device(plf_ctx* ctx, const uint8_t* dev_q, int nq, const uint8_t* dev_t, int nt,
                                         const int32_t* dev_cand_off, const int32_t* dev_cand_idx,
                                         int32_t* dev_best_idx, int32_t* dev_best_dist, int32_t* dev_cand_dist);

/* ---- feature grid and area queries: Frame::AssignFeaturesToGrid / AssignFeaturesToGridLines with PosInGrid /
 * PosInGridLines (src/Frame.cc:365-399, :677-722) and Frame::GetFeaturesInArea / GetFeaturesInAreaLines (:562-676).
 * The queries return CSR candidate lists in the reference's order (cell column outer, cell row inner, insertion
 * order inside a cell) -- the input of plf_hamming_candidates[_device].  Reference grids: 64 x 48 for points,
 * 16 x 12 for lines; inv_w = cols / (mnMaxX - mnMinX), inv_h = rows / (mnMaxY - mnMinY) (src/Frame.cc:139-140). ---- */
typedef struct {
    int32_t cols, rows;
    float min_x, min_y, inv_w, inv_h;
} plf_grid_params;
/* one grid per frame of a batch: dev_kps is [frame][cap] (for lines: the mid-point keypoints, and dev_kls [frame][cap]
 * the keylines whose end points must lie inside the grid too; NULL for points), dev_n [frame]; outputs dev_cell_start
 * [frame][cols*rows+1] (cell = column * rows + row), dev_cell_items [frame][cap] (feature indices, ascending inside a
 * cell), dev_feat_cell [frame][cap] (cell of each feature, -1 = outside) */
plf_status plf_grid_build_device(plf_ctx* ctx, const plf_keypoint* dev_kps, const plf_keyline* dev_kls, const int32_t* dev_n,
                                 int nframes, int cap, const plf_grid_params* g, int32_t* dev_cell_start,
                                 int32_t* dev_cell_items, int32_t* dev_feat_cell);
/* GetFeaturesInArea(x, y, r, minLevel, maxLevel) for nq queries against ONE frame's grid (pass that frame's slices);
 * dev_qminl / dev_qmaxl may be NULL (= -1, no level check).  dev_cand_off gets nq + 1 offsets, dev_cand_idx the
 * feature indices; *total = number of candidates (PLF_ERR_CAPACITY if it exceeds cand_cap). */
plf_status plf_grid_query_device(plf_ctx* ctx, const plf_keypoint* dev_kps, const plf_grid_params* g,
                                 const int32_t* dev_cell_start, const int32_t* dev_cell_items, const float* dev_qx,
                                 const float* dev_qy, const float* dev_qr, const int32_t* dev_qminl, const int32_t* dev_qmaxl,
                                 int nq, int32_t* dev_cand_off, int32_t* dev_cand_idx, int cand_cap, int* total);
/* host buffers, one frame: build + query in one call */
plf_status plf_grid_candidates(plf_ctx* ctx, const plf_keypoint* host_kps, const plf_keyline* host_kls, int n,
                               const plf_grid_params* g, const float* host_qx, const float* host_qy, const float* host_qr,
                               const int32_t* host_qminl, const int32_t* host_qmaxl, int nq, int32_t* host_cand_off,
                               int32_t* host_cand_idx, int cand_cap, int* total);

/* ---- undistortion: Frame::UndistortKeyPoints / UndistortKeyLines (src/Frame.cc:733-826) =
 * cv::undistortPoints(pts, pts, mK, mDistCoef, Mat(), mK) on the key point positions, the line mid-points and both
 * line end points; everything else of the records is copied.  k[0] == 0 -> plain copy, as in the reference.
 * k = (k1, k2, p1, p2[, k3]) as in the settings files (Camera.k1 ... Camera.k3). ---- */
typedef struct {
    float fx, fy, cx, cy;
    float k[5];
    int32_t nk; /* 4 or 5 */
} plf_camera;
plf_status plf_undistort_keypoints(plf_ctx* ctx, const plf_camera* cam, const plf_keypoint* host_in, int n, plf_keypoint* host_out);
plf_status plf_undistort_keypoints_device(plf_ctx* ctx, const plf_camera* cam, const plf_keypoint* dev_in, int n, plf_keypoint* dev_out);
plf_status plf_undistort_keylines(plf_ctx* ctx, const plf_camera* cam, const plf_keyline* host_kl, const plf_keypoint* host_mid, int n,
                                  plf_keyline* host_kl_out, plf_keypoint* host_mid_out);
plf_status plf_undistort_keylines_device(plf_ctx* ctx, const plf_camera* cam, const plf_keyline* dev_kl, const plf_keypoint* dev_mid,
                                         int n, plf_keyline* dev_kl_out, plf_keypoint* dev_mid_out);

/* ---- bag of words: the per-feature vocabulary-tree descent of DBoW2's TemplatedVocabulary::transform(features, BowVector&,
 * FeatureVector&, levelsup) (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1190, :1218-1258) as called from
 * Frame::ComputeBoW (src/Frame.cc:724-731), FORB::distance = 256-bit Hamming.  For feature i: word[i] = leaf word id,
 * weight[i] = its weight (0 = stopped word), node[i] = ancestor at level L - levelsup (0 when that level is <= 0 or the
 * leaf is shallower -- the reference leaves it uninitialised then).  The ordered BowVector / FeatureVector insertions
 * (addWeight, addFeature, normalize) are replayed on the host from these arrays (plf_slam_shim.hpp). ---- */
/* vocabulary from flat arrays: node 0 is the root, parent[i] < i, children keep node-id order and words are numbered in
 * leaf order, exactly as loadFromTextFile builds them (:1377-1418); desc is nnodes x 32 bytes */
plf_status plf_vocab_create(plf_ctx* ctx, int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent,
                            const uint8_t* desc, const double* weight, const uint8_t* is_leaf, plf_vocab** out);
/* the ORBvoc.txt text format of TemplatedVocabulary::loadFromTextFile (:1338-1424) */
plf_status plf_vocab_load_text(plf_ctx* ctx, const char* path, plf_vocab** out);
void plf_vocab_destroy(plf_vocab* v);
plf_status plf_vocab_info(const plf_vocab* v, int* k, int* L, int* nnodes, int* nwords, int* scoring, int* weighting);
plf_status plf_bow_transform(plf_vocab* v, const uint8_t* host_desc, int n, int levelsup, int32_t* host_word,
                             double* host_weight, int32_t* host_node);
plf_status plf_bow_transform_device(plf_vocab* v, const uint8_t* dev_desc, int n, int levelsup, int32_t* dev_word,
                                    double* dev_weight, int32_t* dev_node);

/* ---- stereo: replaces Frame::ComputeStereoMatches (src/Frame.cc:881-1055).  Reads the pyramids the two
 * extractors hold after their last extraction (the reference reads mpORBextractorLeft/Right->mvImagePyramid),
 * so `left` / `right` must have processed the pair's images (the same extractor with two frames of one batch is
 * fine).  mb = baseline in metres, mbf = baseline * fx.  Outputs mvuRight / mvDepth (-1 where unmatched).
 * Defined behaviour where the reference has none: an SAD window that would leave the level image skips the
 * keypoint (the reference throws from cv::Mat::colRange); no match at all leaves everything at -1. ---- */
plf_status plf_stereo_match(plf_orb* left, int frame_l, plf_orb* right, int frame_r,
                            const plf_keypoint* host_kl, const uint8_t* host_dl, int nl,
                            const plf_keypoint* host_kr, const uint8_t* host_dr, int nr,
                            float mb, float mbf, float* host_uright, float* host_depth);
/* batched, device resident: pair p uses frame left_first + p * left_step of `left`'s last batch and
 * right_first + p * right_step of `right`'s; dev_k* / dev_d* / dev_n* are the [frame][cap] outputs of
 * plf_orb_extract_batch_device; dev_uright / dev_depth are [npairs][cap]. */
plf_status plf_stereo_match_batch_device(plf_orb* left, plf_orb* right, int npairs, int left_first, int left_step,
                                         int right_first, int right_step,
                                         const plf_keypoint* dev_kl, const uint8_t* dev_dl, const int32_t* dev_nl,
                                         const plf_keypoint* dev_kr, const uint8_t* dev_dr, const int32_t* dev_nr,
                                         int cap, float mb, float mbf, float* dev_uright, float* dev_depth);

/* measured POPC issue rate of the device (popc32 results per second; XOR + POPC + ADD per result): the roofline
 * denominator of the matching kernels */
plf_status plf_popc_peak(plf_ctx* ctx, double* popc_per_s);

#ifdef __cplusplus
}
#endif
#endif /* PLF_H */
