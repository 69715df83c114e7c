// plf_slam_shim.hpp -- C++ drop-in classes with the reference's signatures on top of the C ABI (plf.h).
//
// Include this INSTEAD of the reference's include/ORBextractor.h / include/Lineextractor.h and link libplf.so:
//   PL_SLAM::ORBextractor        include/ORBextractor.h:45-113      (ctor, operator(), six getters, mvImagePyramid)
//   PL_SLAM::Lineextractor       include/Lineextractor.h:44-181     (LSD ctor, ComputeLsdWithLbd, six getters)
//   PL_SLAM::PlfMatcher          Linematcher::DescriptorDistance / matchNNR / the SearchByKNN mutual step
//                                (include/Linematcher.h:41,68; src/Linematcher.cc:50-66, :454-471, :520-541) and
//                                ORBmatcher::DescriptorDistance (include/ORBmatcher.h:44)
// The translation to the reference's conventions happens here: silent return on empty images, assert on non
// CV_8UC1 input, std::runtime_error where line_descriptor / matchNNR throw, keypoints.clear() +
// descriptors.create() for ORB, append semantics for ComputeLsdWithLbd.
//
// Needs OpenCV's core headers (cv::Mat, cv::KeyPoint) and the vendored KeyLine type.  When built without OpenCV
// (unit test of this header) define PLF_SHIM_MOCK_OPENCV and provide layout-compatible stand-ins first.
#pragma once
#include <cassert>
#include <stdexcept>
#include <string>
#include <vector>
#include <map>
#include <cmath>
#include <cstring>
#include "plf.h"
#ifndef PLF_SHIM_MOCK_OPENCV
#include <opencv2/core/core.hpp>
#include <line_descriptor_custom.hpp>   // cv::line_descriptor::KeyLine (Thirdparty/line_descriptor)
#endif

namespace PL_SLAM {

using cv::line_descriptor::KeyLine;
static_assert(sizeof(cv::KeyPoint) == sizeof(plf_keypoint), "cv::KeyPoint layout");
static_assert(sizeof(KeyLine) == sizeof(plf_keyline), "KeyLine layout");

// one GPU context per extractor / matcher object: the reference calls distinct instances from concurrent
// std::threads (src/Frame.cc:116-119, :301-304), never one instance from two threads
class PlfContext {
public:
    explicit PlfContext(int device = 0) : ctx_(nullptr)
    {
        if (plf_ctx_create(device, &ctx_) != PLF_OK) throw std::runtime_error("plf: no usable CUDA device (there is no CPU fallback)");
    }
    ~PlfContext() { plf_ctx_destroy(ctx_); }
    PlfContext(const PlfContext&) = delete;
    PlfContext& operator=(const PlfContext&) = delete;
    plf_ctx* get() const { return ctx_; }
    void check(plf_status st) const
    {
        if (st != PLF_OK) throw std::runtime_error(std::string("plf: ") + plf_last_error(ctx_));
    }
private:
    plf_ctx* ctx_;
};

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0)
        : ctx_(device), orb_(nullptr), nlevels(nlevels), scaleFactor(scaleFactor)
    {
        plf_orb_params p = {nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST};
        ctx_.check(plf_orb_create(ctx_.get(), &p, &orb_));
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
        mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        plf_orb_tables(orb_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(), nullptr);
        mvImagePyramid.resize(nlevels);
        cap_ = plf_orb_max_keypoints(orb_);
    }
    ~ORBextractor() { plf_orb_destroy(orb_); }

    // Compute the ORB features and descriptors on an image; mask is ignored (as in the reference).
    void operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& _keypoints,
                    cv::OutputArray _descriptors)
    {
        if (_image.empty()) return;                         // src/ORBextractor.cc:1046-1047
        cv::Mat image = _image.getMat();
        assert(image.type() == CV_8UC1);                    // :1050
        kps_.resize(cap_);
        desc_.resize((size_t)cap_ * 32);
        int n = 0;
        ctx_.check(plf_orb_extract(orb_, image.data, image.cols, image.rows, image.step, (plf_keypoint*)kps_.data(),
                                   desc_.data(), cap_, &n));
        // mvImagePyramid (include/ORBextractor.h:85) is filled by ComputePyramid BEFORE the "no keypoints" return of the reference
        // (:1058 vs :1064), so it is refreshed first here too.  The shim's own ComputeStereoMatches reads the device-resident
        // pyramid; code that never touches mvImagePyramid can switch the download off (SetPyramidDownload(false): 8 copies and
        // about 1.1 MB per 752x480 frame less on the 0.3 ms ORB path).
        if (downloadPyramid_) FetchPyramid();
        else for (int l = 0; l < nlevels; l++) mvImagePyramid[l].release();     // never stale levels of an older frame
        _keypoints.clear();                                 // :1072
        if (n == 0) { _descriptors.release(); return; }     // :1064-1065
        _descriptors.create(n, 32, CV_8U);
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; i++) std::memcpy(d.ptr(i), &desc_[(size_t)i * 32], 32);
        _keypoints.assign(kps_.begin(), kps_.begin() + n);
    }

    // host copies of the pyramid levels of the last frame (what the reference keeps in mvImagePyramid)
    void FetchPyramid()
    {
        for (int l = 0; l < nlevels; l++) {
            int w = 0, h = 0;
            ctx_.check(plf_orb_pyramid_level(orb_, 0, l, nullptr, 0, &w, &h));
            mvImagePyramid[l].create(h, w, CV_8UC1);
            ctx_.check(plf_orb_pyramid_level(orb_, 0, l, mvImagePyramid[l].data, mvImagePyramid[l].step, &w, &h));
        }
    }
    void SetPyramidDownload(bool on) { downloadPyramid_ = on; }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    std::vector<cv::Mat> mvImagePyramid;

    plf_orb* handle() const { return orb_; }   // for ComputeStereoMatches below (reads the device-resident pyramid)

protected:
    PlfContext ctx_;
    plf_orb* orb_;
    int nlevels;
    double scaleFactor;
    int cap_;
    bool downloadPyramid_ = true;
    std::vector<cv::KeyPoint> kps_;
    std::vector<unsigned char> desc_;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

class Lineextractor {
public:
    // LSD-LBD constructor (include/Lineextractor.h:49-51)
    Lineextractor(int nfeatures = 240, int nlevels = 3, int refine = 0, double scale = 1.05, double sigma_scale = 0.6,
                  double quant = 2.0, double ang_th = 22.5, double log_eps = 1.0, double density_th = 0.7, int n_bins = 1024,
                  double min_line_length = 32.0, bool busingLSD = true, int device = 0)
        : busingLSD(busingLSD), ctx_(device), le_(nullptr), fld_(nullptr), nlevels(nlevels), scale(scale)
    {
        plf_line_params p = {nfeatures, nlevels, refine, scale, sigma_scale, quant, ang_th, log_eps, density_th, n_bins, min_line_length};
        ctx_.check(plf_line_create(ctx_.get(), &p, &le_));
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
        mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        plf_line_tables(le_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(), nullptr);
        cap_ = plf_line_max_keylines(le_);
    }
    // FLD-LBD constructor (include/Lineextractor.h:55-57, src/Lineextractor.cc:69-110)
    Lineextractor(int _nfeatures, int _nlevels, double _scale, int _length_threshold, float _distance_threshold, double _canny_th1,
                  double _canny_th2, int _canny_aperture_size, bool _do_merge, bool _busingLSD = false, int device = 0)
        : busingLSD(_busingLSD), ctx_(device), le_(nullptr), fld_(nullptr), nlevels(_nlevels), scale(_scale)
    {
        // CV_Assert(_length_threshold > 0 && _distance_threshold > 0 && _canny_th1 > 0 && _canny_th2 > 0 && _canny_aperture_size > 0)
        if (!(_length_threshold > 0 && _distance_threshold > 0 && _canny_th1 > 0 && _canny_th2 > 0 && _canny_aperture_size > 0))
            throw std::runtime_error("Lineextractor: FLD parameters must be positive");
        plf_fld_params p = {_nfeatures, _nlevels, _scale, _length_threshold, _distance_threshold, _canny_th1, _canny_th2, _canny_aperture_size,
                            _do_merge ? 1 : 0};
        ctx_.check(plf_fld_create(ctx_.get(), &p, &fld_));
        cap_ = 2 * _nfeatures + 16;
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
        mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
        for (int i = 1; i < nlevels; i++) { mvScaleFactor[i] = mvScaleFactor[i - 1] * scale; mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i]; }
        for (int i = 0; i < nlevels; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
    }
    ~Lineextractor() { plf_line_destroy(le_); plf_fld_destroy(fld_); }

    // src/Lineextractor.cc:242-336: replaces keyLines, appends to keypoints, overwrites descriptors
    void ComputeFldWithLbd(cv::Mat& image, std::vector<KeyLine>& keyLines, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors)
    {
        if (image.empty()) return;
        if (!fld_) throw std::runtime_error("Lineextractor: built with the LSD constructor");
        std::vector<KeyLine> kl(cap_);
        std::vector<cv::KeyPoint> mid(cap_);
        std::vector<unsigned char> desc((size_t)cap_ * 32);
        int n = 0;
        ctx_.check(plf_fld_extract(fld_, image.data, image.cols, image.rows, image.step, (plf_keyline*)kl.data(), (plf_keypoint*)mid.data(),
                                   desc.data(), cap_, &n));
        keyLines.assign(kl.begin(), kl.begin() + n);
        keypoints.insert(keypoints.end(), mid.begin(), mid.begin() + n);
        if (n == 0) return;
        descriptors.create(n, 32, CV_8UC1);
        for (int i = 0; i < n; i++) std::memcpy(descriptors.ptr(i), &desc[(size_t)i * 32], 32);
    }

    // src/Lineextractor.cc:112-212: appends to keyLines / keypoints, overwrites descriptors
    void ComputeLsdWithLbd(const cv::Mat& image, std::vector<KeyLine>& keyLines, std::vector<cv::KeyPoint>& keypoints,
                           cv::Mat& descriptors)
    {
        if (image.empty()) return;                                                       // :115-116
        if (!le_) throw std::runtime_error("Lineextractor: built with the FLD constructor");
        if (image.depth() != 0 || image.channels() != 1) throw std::runtime_error("Error, depth image!= 0");   // LSDDetector_custom.cpp:236-237
        std::vector<KeyLine> kl(cap_);
        std::vector<cv::KeyPoint> mid(cap_);
        std::vector<unsigned char> desc((size_t)cap_ * 32);
        int n = 0;
        ctx_.check(plf_line_extract(le_, image.data, image.cols, image.rows, image.step, (plf_keyline*)kl.data(),
                                    (plf_keypoint*)mid.data(), desc.data(), cap_, &n));
        // the reference clears keyLines (which held only this frame's detections) and refills it (:184-201)
        keyLines.assign(kl.begin(), kl.begin() + n);
        keypoints.insert(keypoints.end(), mid.begin(), mid.begin() + n);
        if (n == 0) return;   // BinaryDescriptor::compute prints an error and leaves descriptors untouched
        descriptors.create(n, 32, CV_8UC1);
        for (int i = 0; i < n; i++) std::memcpy(descriptors.ptr(i), &desc[(size_t)i * 32], 32);
    }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scale; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    bool busingLSD;
    std::vector<cv::Mat> mvImagePyramid;   // only filled by the FLD branch in the reference

protected:
    PlfContext ctx_;
    plf_line* le_;
    plf_fld* fld_;
    int nlevels;
    double scale;
    int cap_;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

// The brute-force pieces of ORBmatcher / Linematcher.  The candidate-list Search* routines stay on the host in the
// reference's own code and can call DescriptorDistance below unchanged.
class PlfMatcher {
public:
    explicit PlfMatcher(int device = 0) : ctx_(device) {}

    // ORBmatcher::DescriptorDistance / Linematcher::DescriptorDistance: one pair is cheaper on the CPU than a kernel
    // launch, so this keeps the reference's bit-twiddling popcount (src/ORBmatcher.cc:1656-1672) for single pairs.
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b)
    {
        const int* pa = a.ptr<int32_t>();
        const int* pb = b.ptr<int32_t>();
        int dist = 0;
        for (int i = 0; i < 8; i++, pa++, pb++) dist += __builtin_popcount((unsigned)(*pa ^ *pb));
        return dist;
    }

    // Linematcher::matchNNR (src/Linematcher.cc:520-541)
    void matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12, int& nmatches)
    {
        matches_12.resize(desc1.rows, -1);
        if (desc1.rows == 0) return;
        if (desc1.cols != 32 || (desc2.rows && desc2.cols != 32) || !desc1.isContinuous() || !desc2.isContinuous())
            throw std::runtime_error("[matchNNR] descriptors must be continuous N x 32 CV_8U");
        int n = 0;
        ctx_.check(plf_match_nnr(ctx_.get(), desc1.data, desc1.rows, desc2.data, desc2.rows, nnr, matches_12.data(), &n));
        nmatches += n;   // the reference increments the caller's counter
    }

    // both directions + mutual-consistency filter of SearchByKNN / SearchForTriangulation (:454-471, :825-839)
    int matchNNRMutual(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12)
    {
        matches_12.assign(desc1.rows, -1);
        if (desc1.rows == 0 || desc2.rows == 0) return 0;
        int n = 0;
        ctx_.check(plf_match_nnr_mutual(ctx_.get(), desc1.data, desc1.rows, desc2.data, desc2.rows, nnr, matches_12.data(), &n));
        return n;
    }

    // The data-parallel core of the candidate-list matchers (ORBmatcher::SearchForInitialization / SearchByProjection /
    // SearchByBoW, Linematcher::SearchForInitialization / SearchByProjection): query q scans candIdx[candOff[q] ..
    // candOff[q+1]) in order with the reference's `dist < bestDist` / `else if (dist < bestDist2)` rule
    // (src/ORBmatcher.cc:430-456).  bestIdx / bestDist are nq x 2; candDist (optional) gets every candidate's distance
    // for the order-dependent greedy steps, which stay in the reference's host code.
    void candidatesTop2(const cv::Mat& desc1, const cv::Mat& desc2, const std::vector<int>& candOff, const std::vector<int>& candIdx,
                        std::vector<int>& bestIdx, std::vector<int>& bestDist, std::vector<int>* candDist = nullptr)
    {
        bestIdx.assign((size_t)desc1.rows * 2, -1);
        bestDist.assign((size_t)desc1.rows * 2, -1);
        if (candDist) candDist->assign(candIdx.size(), -1);
        if (desc1.rows == 0) return;
        if ((int)candOff.size() != desc1.rows + 1) throw std::runtime_error("[candidatesTop2] candOff must have rows + 1 entries");
        ctx_.check(plf_hamming_candidates(ctx_.get(), desc1.data, desc1.rows, desc2.data, desc2.rows, candOff.data(), candIdx.data(),
                                          bestIdx.data(), bestDist.data(), candDist ? candDist->data() : nullptr));
    }

    // Frame::AssignFeaturesToGrid + GetFeaturesInArea (src/Frame.cc:365-378, :562-617) for many queries at once: the
    // candidate lists (CSR, the reference's vIndices order) that candidatesTop2 scans.  grid: 64 x 48 cells over
    // [mnMinX, mnMaxX) x [mnMinY, mnMaxY), inv_w = cols / (mnMaxX - mnMinX).  minLevel / maxLevel may be empty (= -1).
    void featuresInArea(const std::vector<cv::KeyPoint>& keysUn, const plf_grid_params& grid, const std::vector<float>& x,
                        const std::vector<float>& y, const std::vector<float>& r, const std::vector<int>& minLevel,
                        const std::vector<int>& maxLevel, std::vector<int>& candOff, std::vector<int>& candIdx)
    {
        const int nq = (int)x.size();
        candOff.assign((size_t)nq + 1, 0);
        if (candIdx.size() < (size_t)nq * 32 + 1) candIdx.resize((size_t)nq * 32 + 1);
        for (;;) {
            int total = 0;
            plf_status st = plf_grid_candidates(ctx_.get(), (const plf_keypoint*)keysUn.data(), nullptr, (int)keysUn.size(), &grid, x.data(),
                                                y.data(), r.data(), minLevel.empty() ? nullptr : minLevel.data(),
                                                maxLevel.empty() ? nullptr : maxLevel.data(), nq, candOff.data(), candIdx.data(),
                                                (int)candIdx.size(), &total);
            if (st == PLF_ERR_CAPACITY && total > (int)candIdx.size()) { candIdx.resize((size_t)total); continue; }
            ctx_.check(st);
            candIdx.resize((size_t)total);
            return;
        }
    }

private:
    PlfContext ctx_;
};

// Frame::UndistortKeyPoints / UndistortKeyLines (src/Frame.cc:733-826): cv::undistortPoints(pts, pts, mK, mDistCoef, Mat(), mK)
// on the key point positions, the line mid-points and both line end points; mDistCoef(0) == 0 -> plain copy.
// mK is the 3x3 CV_32F camera matrix, mDistCoef the 4x1 or 5x1 CV_32F coefficient vector of the Frame.
inline plf_camera MakeCamera(const cv::Mat& mK, const cv::Mat& mDistCoef)
{
    plf_camera c;
    std::memset(&c, 0, sizeof(c));
    c.fx = mK.at<float>(0, 0); c.fy = mK.at<float>(1, 1); c.cx = mK.at<float>(0, 2); c.cy = mK.at<float>(1, 2);
    c.nk = mDistCoef.rows * mDistCoef.cols > 4 ? 5 : 4;
    for (int i = 0; i < c.nk; i++) c.k[i] = mDistCoef.at<float>(i);
    return c;
}
inline void UndistortKeyPoints(PlfContext& ctx, const cv::Mat& mK, const cv::Mat& mDistCoef, const std::vector<cv::KeyPoint>& mvKeys,
                               std::vector<cv::KeyPoint>& mvKeysUn)
{
    mvKeysUn.resize(mvKeys.size());
    if (mvKeys.empty()) return;
    const plf_camera c = MakeCamera(mK, mDistCoef);
    ctx.check(plf_undistort_keypoints(ctx.get(), &c, (const plf_keypoint*)mvKeys.data(), (int)mvKeys.size(), (plf_keypoint*)mvKeysUn.data()));
}
inline void UndistortKeyLines(PlfContext& ctx, const cv::Mat& mK, const cv::Mat& mDistCoef, const std::vector<KeyLine>& mvLines,
                              const std::vector<cv::KeyPoint>& mvMidPoints, std::vector<KeyLine>& mvLinesUn,
                              std::vector<cv::KeyPoint>& mvMidPointsUn)
{
    mvLinesUn.resize(mvLines.size());
    mvMidPointsUn.resize(mvMidPoints.size());
    if (mvLines.empty()) return;
    const plf_camera c = MakeCamera(mK, mDistCoef);
    ctx.check(plf_undistort_keylines(ctx.get(), &c, (const plf_keyline*)mvLines.data(), (const plf_keypoint*)mvMidPoints.data(),
                                     (int)mvLines.size(), (plf_keyline*)mvLinesUn.data(), (plf_keypoint*)mvMidPointsUn.data()));
}

// ORBVocabulary::transform(features, BowVector&, FeatureVector&, levelsup) as used by Frame::ComputeBoW
// (src/Frame.cc:724-731; Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1190): the tree descent of every descriptor
// runs on the GPU, the ordered map insertions (BowVector::addWeight, FeatureVector::addFeature) and the L1
// normalisation are replayed here in feature order, so the two maps equal DBoW2's (TF_IDF / TF weighting, L1 scoring,
// what ORBvoc.txt uses).  BowVec / FeatVec are any std::map-like types: DBoW2::BowVector / DBoW2::FeatureVector fit.
class PlfVocabulary {
public:
    explicit PlfVocabulary(const std::string& orbvoc_txt, int device = 0) : ctx_(device), voc_(nullptr)
    {
        ctx_.check(plf_vocab_load_text(ctx_.get(), orbvoc_txt.c_str(), &voc_));
    }
    ~PlfVocabulary() { plf_vocab_destroy(voc_); }
    PlfVocabulary(const PlfVocabulary&) = delete;
    PlfVocabulary& operator=(const PlfVocabulary&) = delete;

    template <class BowVec, class FeatVec>
    void transform(const cv::Mat& descriptors, BowVec& v, FeatVec& fv, int levelsup)
    {
        v.clear();
        fv.clear();
        const int n = descriptors.rows;
        if (n == 0) return;
        word_.resize(n); weight_.resize(n); node_.resize(n);
        ctx_.check(plf_bow_transform(voc_, descriptors.data, n, levelsup, word_.data(), weight_.data(), node_.data()));
        for (int i = 0; i < n; i++)
            if (weight_[i] > 0) {                           // not stopped
                v[(unsigned)word_[i]] += weight_[i];        // BowVector::addWeight
                fv[(unsigned)node_[i]].push_back((unsigned)i);   // FeatureVector::addFeature
            }
        double norm = 0.0;                                  // BowVector::normalize(L1)
        for (typename BowVec::iterator it = v.begin(); it != v.end(); ++it) norm += std::fabs(it->second);
        if (norm > 0.0)
            for (typename BowVec::iterator it = v.begin(); it != v.end(); ++it) it->second /= norm;
    }

private:
    PlfContext ctx_;
    plf_vocab* voc_;
    std::vector<int> word_, node_;
    std::vector<double> weight_;
};

// Frame::ComputeStereoMatches (src/Frame.cc:881-1055) for the pair the two extractors processed last: fills mvuRight
// and mvDepth exactly like the reference (row-band Hamming best-1, 11x11 SAD slide on the extractors' pyramids,
// parabola sub-pixel, 2.1 x median SAD filter).  mb = baseline [m], mbf = baseline * fx.
inline void ComputeStereoMatches(const ORBextractor& left, const ORBextractor& right, const std::vector<cv::KeyPoint>& mvKeys,
                                 const cv::Mat& mDescriptors, const std::vector<cv::KeyPoint>& mvKeysRight,
                                 const cv::Mat& mDescriptorsRight, float mb, float mbf, std::vector<float>& mvuRight,
                                 std::vector<float>& mvDepth)
{
    const int N = (int)mvKeys.size(), Nr = (int)mvKeysRight.size();
    mvuRight.assign(N, -1.0f);
    mvDepth.assign(N, -1.0f);
    if (N == 0) return;
    plf_status st = plf_stereo_match(left.handle(), 0, right.handle(), 0, (const plf_keypoint*)mvKeys.data(), mDescriptors.data, N,
                                     (const plf_keypoint*)mvKeysRight.data(), mDescriptorsRight.data, Nr, mb, mbf, mvuRight.data(),
                                     mvDepth.data());
    if (st != PLF_OK) throw std::runtime_error("plf: ComputeStereoMatches failed");
}

}  // namespace PL_SLAM
