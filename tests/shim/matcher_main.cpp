// Drives the drop-in matcher classes (include/plf_matcher_shim.hpp) on mock SLAM objects read from a binary file and dumps
// the results; tests/test_matcher_shim.py compares them with the reference's own functions (oracle/_ref) on the same data.
// usage: matcher_main in.bin out.bin
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "mock_opencv.hpp"
#define PLF_SHIM_MOCK_OPENCV
#include "plf_matcher_shim.hpp"
#define MOCK_KEYLINE PL_SLAM::KeyLine
#include "mock_slam.hpp"

// the one Frame member whose body the reference build cuts from src/Frame.cc:562-615; restated for this test binary
std::vector<std::size_t> PL_SLAM::Frame::GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel, const int maxLevel) const
{
    std::vector<std::size_t> vIndices;
    const int nMinCellX = std::max(0, (int)floor((x - mnMinX - r) * mfGridElementWidthInv));
    if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
    const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x - mnMinX + r) * mfGridElementWidthInv));
    if (nMaxCellX < 0) return vIndices;
    const int nMinCellY = std::max(0, (int)floor((y - mnMinY - r) * mfGridElementHeightInv));
    if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
    const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - mnMinY + r) * mfGridElementHeightInv));
    if (nMaxCellY < 0) return vIndices;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const std::vector<std::size_t>& vCell = mGrid[ix][iy];
            for (std::size_t j = 0; j < vCell.size(); j++) {
                const cv::KeyPoint& kpUn = mvKeysUn[vCell[j]];
                if (bCheckLevels) {
                    if (kpUn.octave < minLevel) continue;
                    if (maxLevel >= 0 && kpUn.octave > maxLevel) continue;
                }
                const float distx = kpUn.pt.x - x, disty = kpUn.pt.y - y;
                if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
            }
        }
    return vIndices;
}

struct Reader {
    std::vector<unsigned char> buf;
    size_t pos = 0;
    template <class T> T get() { T v; memcpy(&v, &buf[pos], sizeof(T)); pos += sizeof(T); return v; }
    const unsigned char* bytes(size_t n) { const unsigned char* p = buf.data() + pos; pos += n; return p; }
    std::vector<unsigned char> vec(size_t n) { const unsigned char* p = bytes(n); return std::vector<unsigned char>(p, p + n); }
};

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    Reader R;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    R.buf.resize(sz);
    if (fread(R.buf.data(), 1, sz, f) != (size_t)sz) return 3;
    fclose(f);
    FILE* o = fopen(argv[2], "wb");
    // ---- ORBmatcher::SearchForInitialization ----
    {
        const int n1 = R.get<int>(), n2 = R.get<int>(), window = R.get<int>(), checkOri = R.get<int>();
        const float nnr = R.get<float>();
        float bounds[4]; for (float& b : bounds) b = R.get<float>();
        PL_SLAM::Frame F1, F2;
        F1.mvKeysUn.resize(n1); memcpy(F1.mvKeysUn.data(), R.bytes((size_t)n1 * 28), (size_t)n1 * 28);
        std::vector<unsigned char> d1 = R.vec((size_t)n1 * 32);
        F2.mvKeysUn.resize(n2); memcpy(F2.mvKeysUn.data(), R.bytes((size_t)n2 * 28), (size_t)n2 * 28);
        std::vector<unsigned char> d2 = R.vec((size_t)n2 * 32);
        F1.mDescriptors = cv::Mat(n1, 32, CV_8UC1, d1.data());
        F2.mDescriptors = cv::Mat(n2, 32, CV_8UC1, d2.data());
        F1.SetBoundsAndAssign(bounds[0], bounds[1], bounds[2], bounds[3]);
        F2.SetBoundsAndAssign(bounds[0], bounds[1], bounds[2], bounds[3]);
        std::vector<cv::Point2f> prev(n1);
        memcpy(prev.data(), R.bytes((size_t)n1 * 8), (size_t)n1 * 8);
        std::vector<int> m12;
        PL_SLAM::ORBmatcher om(nnr, checkOri != 0);
        const int n = om.SearchForInitialization(F1, F2, prev, m12, window);
        fwrite(&n, 4, 1, o);
        fwrite(m12.data(), 4, n1, o);
        fwrite(prev.data(), 8, n1, o);
    }
    // ---- Linematcher::SearchByKNN ----
    {
        const int nkf = R.get<int>(), nf = R.get<int>(), checklen = R.get<int>();
        const float nnr = R.get<float>(), lengtherr = R.get<float>();
        std::vector<unsigned char> dkf = R.vec((size_t)nkf * 32); std::vector<unsigned char> df = R.vec((size_t)nf * 32);
        std::vector<unsigned char> state = R.vec(nkf);
        std::vector<float> mllen(nkf), flen(nf);
        memcpy(mllen.data(), R.bytes((size_t)nkf * 4), (size_t)nkf * 4);
        memcpy(flen.data(), R.bytes((size_t)nf * 4), (size_t)nf * 4);
        std::vector<PL_SLAM::MapLine> mls(nkf);
        PL_SLAM::KeyFrame kf;
        PL_SLAM::Frame F;
        kf.mDescriptorLines = cv::Mat(nkf, 32, CV_8UC1, dkf.data());
        kf.mvpMapLines.resize(nkf);
        for (int i = 0; i < nkf; i++) { mls[i].mbBad = state[i] == 2; mls[i].mfLen = mllen[i]; kf.mvpMapLines[i] = state[i] ? &mls[i] : nullptr; }
        F.NL = nf;
        F.mDescriptorLines = cv::Mat(nf, 32, CV_8UC1, df.data());
        F.mvLinesUn.resize(nf);
        for (int i = 0; i < nf; i++) F.mvLinesUn[i].lineLength = flen[i];
        std::vector<PL_SLAM::MapLine*> out;
        PL_SLAM::Linematcher lm(nnr, true, checklen != 0, lengtherr);
        const int n = lm.SearchByKNN(&kf, F, out);
        fwrite(&n, 4, 1, o);
        for (int i = 0; i < nf; i++) { int v = out[i] ? (int)(out[i] - mls.data()) : -1; fwrite(&v, 4, 1, o); }
    }
    // ---- Linematcher::SearchForTriangulation ----
    {
        const int n1 = R.get<int>(), n2 = R.get<int>(), nlev = R.get<int>();
        const float nnr = R.get<float>();
        std::vector<unsigned char> d1 = R.vec((size_t)n1 * 32); std::vector<unsigned char> d2 = R.vec((size_t)n2 * 32);
        std::vector<unsigned char> h1 = R.vec(n1); std::vector<unsigned char> h2 = R.vec(n2);
        PL_SLAM::MapLine dummy;
        PL_SLAM::KeyFrame k1, k2;
        k1.mvMidPointsUn.resize(n1); memcpy(k1.mvMidPointsUn.data(), R.bytes((size_t)n1 * 28), (size_t)n1 * 28);
        k2.mvMidPointsUn.resize(n2); memcpy(k2.mvMidPointsUn.data(), R.bytes((size_t)n2 * 28), (size_t)n2 * 28);
        std::vector<float> scale(nlev), sig(nlev);
        memcpy(scale.data(), R.bytes((size_t)nlev * 4), (size_t)nlev * 4);
        memcpy(sig.data(), R.bytes((size_t)nlev * 4), (size_t)nlev * 4);
        float cam[4], pose[15], f12[9];
        memcpy(cam, R.bytes(16), 16); memcpy(pose, R.bytes(60), 60); memcpy(f12, R.bytes(36), 36);
        k1.mDescriptorLines = cv::Mat(n1, 32, CV_8UC1, d1.data());
        k2.mDescriptorLines = cv::Mat(n2, 32, CV_8UC1, d2.data());
        k1.mvpMapLines.resize(n1); k2.mvpMapLines.resize(n2);
        for (int i = 0; i < n1; i++) k1.mvpMapLines[i] = h1[i] ? &dummy : nullptr;
        for (int i = 0; i < n2; i++) k2.mvpMapLines[i] = h2[i] ? &dummy : nullptr;
        for (PL_SLAM::KeyFrame* k : {&k1, &k2}) { k->mvScaleFactorsLines = scale; k->mvLevelSigma2Lines = sig; k->fx = cam[0]; k->fy = cam[1]; k->cx = cam[2]; k->cy = cam[3]; }
        k1.mOw = cv::Mat(3, 1, 5, pose, 4); k2.mRcw = cv::Mat(3, 3, 5, pose + 3, 12); k2.mtcw = cv::Mat(3, 1, 5, pose + 12, 4);
        cv::Mat Fm(3, 3, 5, f12, 12);
        std::vector<std::pair<size_t, size_t> > vp;
        PL_SLAM::Linematcher lm(nnr, true, true, 0.1f);
        const int n = lm.SearchForTriangulation(&k1, &k2, Fm, vp);
        fwrite(&n, 4, 1, o);
        for (auto& pr : vp) { int a = (int)pr.first, b = (int)pr.second; fwrite(&a, 4, 1, o); fwrite(&b, 4, 1, o); }
    }
    fclose(o);
    printf("matcher shim ok\n");
    return 0;
}
