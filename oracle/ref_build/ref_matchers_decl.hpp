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
#include <vector>
#include <stdexcept>

namespace cv {
class BFMatcher {
public:
    static Ptr<BFMatcher> create(int normType = NORM_L2, bool crossCheck = false);
    void knnMatch(InputArray queryDescriptors, InputArray trainDescriptors, std::vector<std::vector<DMatch> >& matches, int k);
};
}

using namespace std; /* both reference files say so at file scope */

namespace PL_SLAM {
class ORBmatcher {
public:
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
};
class Linematcher {
public:
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    void matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12, int& nmatches);
};
}
#endif
