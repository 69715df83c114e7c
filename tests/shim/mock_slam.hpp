// mock_slam.hpp -- TEST ONLY.  The smallest stand-ins for the SLAM-side types that the reference's matcher functions touch
// (Frame, KeyFrame, MapLine: include/Frame.h, include/KeyFrame.h, include/MapLine.h), with the same member names, so that
//   (a) the reference's OWN functions -- src/ORBmatcher.cc:406-521, src/Linematcher.cc:121-143, :437-517, :804-879 and
//       src/Frame.cc:562-617 -- compile against them by line range into oracle/_ref/libref.so, and
//   (b) the drop-in templates of include/plf_matcher_shim.hpp are run on exactly the same objects.
// The includer provides cv::Mat / cv::KeyPoint / cv::Point2f and a KeyLine type named MOCK_KEYLINE.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace PL_SLAM {

class MapLine {
public:
    bool mbBad;
    float mfLen;
    MapLine() : mbBad(false), mfLen(0) {}
    bool isBad() { return mbBad; }
    float Get2DLineLengthAverage() { return mfLen; }
};

class Frame {
public:
    int N, NL;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors, mDescriptorLines;
    std::vector<MOCK_KEYLINE> mvLinesUn;
    std::vector<cv::KeyPoint> mvMidPointsUn;
    float mnMinX, mnMinY, mnMaxX, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];

    Frame() : N(0), NL(0), mnMinX(0), mnMinY(0), mnMaxX(0), mnMaxY(0), mfGridElementWidthInv(0), mfGridElementHeightInv(0) {}

    // src/Frame.cc:139-140, :365-378, :678-690 restated for the mock (the grid is set up once per test frame)
    void SetBoundsAndAssign(float minX, float maxX, float minY, float maxY)
    {
        mnMinX = minX; mnMaxX = maxX; mnMinY = minY; mnMaxY = maxY;
        mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (mnMaxX - mnMinX);
        mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (mnMaxY - mnMinY);
        for (int i = 0; i < FRAME_GRID_COLS; i++)
            for (int j = 0; j < FRAME_GRID_ROWS; j++) mGrid[i][j].clear();
        N = (int)mvKeysUn.size();
        for (int i = 0; i < N; i++) {
            const cv::KeyPoint& kp = mvKeysUn[i];
            int posX = (int)std::round((kp.pt.x - mnMinX) * mfGridElementWidthInv);
            int posY = (int)std::round((kp.pt.y - mnMinY) * mfGridElementHeightInv);
            if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) continue;
            mGrid[posX][posY].push_back(i);
        }
    }
    std::vector<std::size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const;
};

// Linematcher::SearchForTriangulation indexes mvMidPointsUn with -1 for unmatched lines (src/Linematcher.cc:846: undefined
// behaviour on a std::vector).  The mock keeps one defined sentinel in front so that the reference's code can be run safely;
// everything it computes from the sentinel only ever touches entries that are already -1.
template <class T> struct MockPadded {
    std::vector<T> v;
    MockPadded() : v(1) {}
    void resize(std::size_t n) { v.resize(n + 1); }
    std::size_t size() const { return v.size() - 1; }
    T* data() { return v.data() + 1; }
    const T& operator[](std::ptrdiff_t i) const { return v[(std::size_t)(i + 1)]; }
    T& operator[](std::ptrdiff_t i) { return v[(std::size_t)(i + 1)]; }
};

class KeyFrame {
public:
    cv::Mat mDescriptorLines;
    std::vector<MapLine*> mvpMapLines;
    MockPadded<cv::KeyPoint> mvMidPointsUn;
    std::vector<float> mvScaleFactorsLines, mvLevelSigma2Lines;
    float fx, fy, cx, cy;
    cv::Mat mOw, mRcw, mtcw;       // 3x1, 3x3, 3x1 CV_32F
    std::vector<MapLine*> GetMapLineMatches() { return mvpMapLines; }
    cv::Mat GetCameraCenter() { return mOw; }
    cv::Mat GetRotation() { return mRcw; }
    cv::Mat GetTranslation() { return mtcw; }
};

} // namespace PL_SLAM
