// plf_matcher_shim.hpp -- drop-in PL_SLAM::ORBmatcher / PL_SLAM::Linematcher over the C ABI (include/plf.h).
//
// What a SPL-SLAM build includes instead of include/ORBmatcher.h / include/Linematcher.h for the matcher entry points that
// sit on the descriptor hot path.  Same class names, constructors, constants and member signatures as the reference
// (include/ORBmatcher.h:36-108, include/Linematcher.h:33-78); the SLAM-side types (Frame, KeyFrame, MapLine) are template
// parameters that are deduced at the call site, so the reference's own classes fit unchanged and nothing of the rest of
// SPL-SLAM is pulled in here.  The Hamming work runs on the GPU through the C ABI -- both matchNNR directions of
// SearchByKNN / SearchForTriangulation, all candidate distances of SearchForInitialization in one call -- and the
// order-dependent steps (greedy "already matched at a smaller distance" skips, rotation histogram + ComputeThreeMaxima,
// mutual filter, length and epipolar checks) are replayed on the host exactly as the reference writes them.
//
// Parity: tests/test_matcher_shim.py runs these templates and the REFERENCE'S OWN functions (cut by line range from
// src/ORBmatcher.cc / src/Linematcher.cc into oracle/_ref) on the same mock Frame / KeyFrame objects.
// Defined behaviour where the reference reads out of bounds: matches_21[i2] and mvMidPointsUn[i2] with i2 == -1
// (src/Linematcher.cc:462, :829, :846) -- unmatched entries are skipped; the matched pairs / assignments are the reference's,
// SearchForTriangulation's return value (which the reference decrements on garbage) is the number of surviving pairs.
#ifndef PLF_MATCHER_SHIM_HPP
#define PLF_MATCHER_SHIM_HPP

#include "plf_slam_shim.hpp"
#include <climits>
#include <cmath>
#include <utility>
#include <vector>

namespace PL_SLAM {

class ORBmatcher {
public:
    static const int TH_LOW = 50;
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 30;

    ORBmatcher(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri), m_(device) {}

    // src/ORBmatcher.cc:1656-1672 (a single pair: stays on the host, bit tricks as in the reference)
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return PlfMatcher::DescriptorDistance(a, b); }

    // src/ORBmatcher.cc:406-521.  The candidate lists come from F2.GetFeaturesInArea (the Frame's own grid, level 0 only), every
    // candidate distance from ONE plf_hamming_candidates call; the loop below is the reference's, reading those distances.
    template <class FrameT, class Point2fT>
    int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<Point2fT>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10)
    {
        int nmatches = 0;
        vnMatches12 = std::vector<int>(F1.mvKeysUn.size(), -1);
        std::vector<int> rotHist[HISTO_LENGTH];
        for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
        const float factor = 1.0f / HISTO_LENGTH;
        std::vector<int> vMatchedDistance(F2.mvKeysUn.size(), INT_MAX);
        std::vector<int> vnMatches21(F2.mvKeysUn.size(), -1);

        // candidate lists in the reference's order (CSR)
        const size_t n1 = F1.mvKeysUn.size();
        std::vector<int> off(n1 + 1, 0), idx;
        for (size_t i1 = 0; i1 < n1; i1++) {
            off[i1 + 1] = off[i1];
            const int level1 = F1.mvKeysUn[i1].octave;
            if (level1 > 0) continue;
            const std::vector<size_t> v = F2.GetFeaturesInArea(vbPrevMatched[i1].x, vbPrevMatched[i1].y, windowSize, level1, level1);
            for (size_t k = 0; k < v.size(); k++) idx.push_back((int)v[k]);
            off[i1 + 1] = (int)idx.size();
        }
        std::vector<int> bestIdx, bestDist, candDist;
        if (!idx.empty()) m_.candidatesTop2(F1.mDescriptors, F2.mDescriptors, off, idx, bestIdx, bestDist, &candDist);

        for (size_t i1 = 0; i1 < n1; i1++) {
            if (off[i1 + 1] == off[i1]) continue;      // level1 > 0 or no candidates
            int bestD = INT_MAX, bestD2 = INT_MAX, bestIdx2 = -1;
            for (int c = off[i1]; c < off[i1 + 1]; c++) {
                const int i2 = idx[c];
                const int dist = candDist[c];
                if (vMatchedDistance[i2] <= dist) continue;
                if (dist < bestD) { bestD2 = bestD; bestD = dist; bestIdx2 = i2; }
                else if (dist < bestD2) bestD2 = dist;
            }
            if (bestD <= TH_LOW) {
                if (bestD < (float)bestD2 * mfNNratio) {
                    if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                    vnMatches12[i1] = bestIdx2;
                    vnMatches21[bestIdx2] = (int)i1;
                    vMatchedDistance[bestIdx2] = bestD;
                    nmatches++;
                    if (mbCheckOrientation) {
                        float rot = F1.mvKeysUn[i1].angle - F2.mvKeysUn[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back((int)i1);
                    }
                }
            }
        }
        if (mbCheckOrientation) {
            int ind1 = -1, ind2 = -1, ind3 = -1;
            ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
            for (int i = 0; i < HISTO_LENGTH; i++) {
                if (i == ind1 || i == ind2 || i == ind3) continue;
                for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) {
                    const int idx1 = rotHist[i][j];
                    if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
                }
            }
        }
        for (size_t i1 = 0, iend1 = vnMatches12.size(); i1 < iend1; i1++)
            if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2.mvKeysUn[vnMatches12[i1]].pt;
        return nmatches;
    }

    // src/ORBmatcher.cc:1610-1651
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3)
    {
        int max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < L; i++) {
            const int s = (int)histo[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
    }

    PlfMatcher& matcher() { return m_; }

protected:
    float mfNNratio;
    bool mbCheckOrientation;
    PlfMatcher m_;
};

class Linematcher {
public:
    static const int TH_HIGH = 100;
    static const int TH_LOW = 50;
    static const int HISTO_LENGTH = 30;

    Linematcher(float nnratio = 0.6, bool checkOri = true, bool checklen = true, float lengtherr = 0.1, int device = 0)
        : mfNNratio(nnratio), mbCheckOrientation(checkOri), mbchecklen(checklen), mflengtherr(lengtherr), m_(device) {}

    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return PlfMatcher::DescriptorDistance(a, b); }

    // src/Linematcher.cc:520-541 (brute-force top-2 + ratio test on the GPU)
    void matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12, int& nmatches)
    {
        m_.matchNNR(desc1, desc2, nnr, matches_12, nmatches);
    }

    // src/Linematcher.cc:437-517
    template <class KeyFrameT, class FrameT, class MapLineT>
    int SearchByKNN(KeyFrameT* pKF, FrameT& F, std::vector<MapLineT*>& vpMapLineMatches)
    {
        int nmatches12 = 0, nmatches21 = 0;
        cv::Mat desc1 = pKF->mDescriptorLines;
        cv::Mat desc2 = F.mDescriptorLines;
        std::vector<int> matches_12, matches_21;
        const std::vector<MapLineT*> vpMapLinesKF = pKF->GetMapLineMatches();
        vpMapLineMatches = std::vector<MapLineT*>(F.NL, static_cast<MapLineT*>(NULL));
        matchNNR(desc1, desc2, mfNNratio, matches_12, nmatches12);      // the reference's two threads (:454-457): two GPU calls
        matchNNR(desc2, desc1, mfNNratio, matches_21, nmatches21);
        for (size_t i1 = 0; i1 < matches_12.size(); ++i1) {
            int& i2 = matches_12[i1];
            if (i2 >= 0 && (size_t)matches_21[i2] != i1) { i2 = -1; nmatches12--; }
        }
        if (mbchecklen) {
            for (size_t i1 = 0; i1 < matches_12.size(); ++i1) {
                int& i2 = matches_12[i1];
                if (i2 >= 0) {
                    MapLineT* pML = vpMapLinesKF[i1];
                    if (!pML) { i2 = -1; nmatches12--; continue; }
                    if (pML->isBad()) { i2 = -1; nmatches12--; continue; }
                    const float LineAverageLength = pML->Get2DLineLengthAverage();
                    if (((1 - mflengtherr) * LineAverageLength <= F.mvLinesUn[i2].lineLength) ||
                        ((1 + mflengtherr) * LineAverageLength >= F.mvLinesUn[i2].lineLength))
                        vpMapLineMatches[i2] = pML;
                    else { i2 = -1; nmatches12--; }
                }
            }
        }
        return nmatches12;
    }

    // src/Linematcher.cc:121-143
    template <class KeyFrameT>
    bool CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrameT* pKF2)
    {
        const float a = kp1.pt.x * F12.template at<float>(0, 0) + kp1.pt.y * F12.template at<float>(1, 0) + F12.template at<float>(2, 0);
        const float b = kp1.pt.x * F12.template at<float>(0, 1) + kp1.pt.y * F12.template at<float>(1, 1) + F12.template at<float>(2, 1);
        const float c = kp1.pt.x * F12.template at<float>(0, 2) + kp1.pt.y * F12.template at<float>(1, 2) + F12.template at<float>(2, 2);
        const float num = a * kp2.pt.x + b * kp2.pt.y + c;
        const float den = a * a + b * b;
        if (den == 0) return false;
        const float dsqr = num * num / den;
        return dsqr < 3.841 * pKF2->mvLevelSigma2Lines[kp2.octave];
    }

    // src/Linematcher.cc:804-879
    template <class KeyFrameT>
    int SearchForTriangulation(KeyFrameT* pKF1, KeyFrameT* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs)
    {
        // epipole in the second image: C2 = R2w * Cw + t2w, float products summed left to right (cv::gemm's small-matrix path)
        const cv::Mat Cw = pKF1->GetCameraCenter(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
        float C2[3];
        for (int i = 0; i < 3; i++) {
            float t = R2w.template at<float>(i, 0) * Cw.template at<float>(0);
            t = t + R2w.template at<float>(i, 1) * Cw.template at<float>(1);
            t = t + R2w.template at<float>(i, 2) * Cw.template at<float>(2);
            C2[i] = t + t2w.template at<float>(i);
        }
        const float invz = 1.0f / C2[2];
        const float ex = pKF2->fx * C2[0] * invz + pKF2->cx;
        const float ey = pKF2->fy * C2[1] * invz + pKF2->cy;

        cv::Mat desc1 = pKF1->mDescriptorLines, desc2 = pKF2->mDescriptorLines;
        std::vector<int> matches_12, matches_21;
        int nmatches12 = 0, nmatches21 = 0;
        matchNNR(desc1, desc2, mfNNratio, matches_12, nmatches12);
        matchNNR(desc2, desc1, mfNNratio, matches_21, nmatches21);
        const auto vpMapLinesKF1 = pKF1->GetMapLineMatches();
        const auto vpMapLinesKF2 = pKF2->GetMapLineMatches();
        for (size_t i1 = 0; i1 < matches_12.size(); ++i1) {
            int& i2 = matches_12[i1];
            if (i2 >= 0 && (size_t)matches_21[i2] != i1 && vpMapLinesKF1[i1] && vpMapLinesKF2[i2]) i2 = -1;
        }
        for (size_t i1 = 0; i1 < matches_12.size(); ++i1) {
            int& i2 = matches_12[i1];
            if (i2 < 0) continue;                 // the reference indexes mvMidPointsUn[-1] here; an unmatched line stays unmatched
            const cv::KeyPoint& kp1 = pKF1->mvMidPointsUn[i1];
            const cv::KeyPoint& kp2 = pKF2->mvMidPointsUn[i2];
            const float distex = ex - kp2.pt.x, distey = ey - kp2.pt.y;
            const bool near_epipole = distex * distex + distey * distey < 100 * pKF2->mvScaleFactorsLines[kp2.octave];
            const bool on_line = CheckDistEpipolarLine(kp1, kp2, F12, pKF2);
            if (near_epipole || !on_line) i2 = -1;
        }
        int n = 0;
        for (size_t i1 = 0; i1 < matches_12.size(); ++i1) {
            if (matches_12[i1] < 0) continue;
            vMatchedPairs.push_back(std::make_pair(i1, (size_t)matches_12[i1]));
            n++;
        }
        return n;
    }

    PlfMatcher& matcher() { return m_; }

protected:
    float mfNNratio;
    bool mbCheckOrientation;
    bool mbchecklen;
    float mflengtherr;
    PlfMatcher m_;
};

}  // namespace PL_SLAM
#endif
