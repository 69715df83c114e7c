/*
 * ref_matchers_decl.hpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 * Declarations under which the line ranges cut from the reference's src/ORBmatcher.cc (:1609-1672) and
 * src/Linematcher.cc (:50-66, :520-541) compile on their own.  The real headers (include/ORBmatcher.h:44,
 * include/Linematcher.h:41,68) drag in Frame / MapPoint / KeyFrame (Eigen, g2o, DBoW2), none of which is on the path.
 * cv::BFMatcher::knnMatch is modelled by orc_knn2 (pinned against cv2.BFMatcher in tests/test_oracle_vs_cv2.py).
 */
#ifndef PLF_REF_MATCHERS_DECL_HPP
#define PLF_REF_MATCHERS_DECL_HPP
#include "cvshim.hpp"
#include "line_descriptor_custom.hpp"
#include <vector>
#include <stdexcept>
#include <thread>
#include <functional>
#include <climits>
#include <cmath>
#include <cassert>
#define MOCK_KEYLINE cv::line_descriptor::KeyLine
#include "../../tests/shim/mock_slam.hpp"

namespace cv {
class BFMatcher {
public:
    static Ptr<BFMatcher> create(int normType = NORM_L2, bool crossCheck = false);
    void knnMatch(InputArray queryDescriptors, InputArray trainDescriptors, std::vector<std::vector<DMatch> >& matches, int k);
};
}

using namespace std; /* both reference files say so at file scope */
using namespace cv::line_descriptor;

namespace PL_SLAM {
/* include/ORBmatcher.h:36-108 and include/Linematcher.h:33-78, reduced to the members the cut ranges define or use */
class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    static const int TH_LOW = 50;
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 30;
    float mfNNratio;
    bool mbCheckOrientation;
};
class Linematcher {
public:
    Linematcher(float nnratio = 0.6, bool checkOri = true, bool checklen = true, float lengtherr = 0.1)
        : mfNNratio(nnratio), mbCheckOrientation(checkOri), mbchecklen(checklen), mflengtherr(lengtherr) {}
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByKNN(KeyFrame* pKF, Frame& F, std::vector<MapLine*>& vpMapLineMatches);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<pair<size_t, size_t> >& vMatchedPairs);
    bool CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF);
    void matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12, int& nmatches);
    float mfNNratio;
    bool mbCheckOrientation;
    bool mbchecklen;
    float mflengtherr;
};
}
#endif
