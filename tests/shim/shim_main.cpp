// Drives the C++ drop-in classes (include/plf_slam_shim.hpp) on a raw 8-bit image and dumps the results.
// usage: shim_main in.raw W H out.bin
#include <cstdio>
#include "mock_opencv.hpp"
#define PLF_SHIM_MOCK_OPENCV
#include "plf_slam_shim.hpp"

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    int W = atoi(argv[2]), H = atoi(argv[3]);
    std::vector<unsigned char> img((size_t)W * H);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(img.data(), 1, img.size(), f) != img.size()) return 3;
    fclose(f);
    cv::Mat im(H, W, CV_8UC1, img.data());
    PL_SLAM::ORBextractor orb(500, 1.2f, 6, 20, 7);
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc, mask;
    orb(im, mask, kps, desc);
    PL_SLAM::Lineextractor le(100, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0, true);
    std::vector<PL_SLAM::KeyLine> kl;
    std::vector<cv::KeyPoint> mid;
    cv::Mat ld;
    le.ComputeLsdWithLbd(im, kl, mid, ld);
    PL_SLAM::PlfMatcher m;
    std::vector<int> m12;
    int nm = 0;
    m.matchNNR(desc, desc, 0.9f, m12, nm);
    int d01 = PL_SLAM::PlfMatcher::DescriptorDistance(cv::Mat(1, 32, 0, desc.ptr(0)), cv::Mat(1, 32, 0, desc.ptr(1)));
    // empty image: silent return, outputs untouched
    std::vector<cv::KeyPoint> k2(3);
    cv::Mat e, d2;
    orb(e, mask, k2, d2);
    // stereo: right image = left shifted by 9 px; a second extractor instance, as in Frame.cc:116-119
    std::vector<unsigned char> rimg((size_t)W * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) rimg[(size_t)y * W + x] = img[(size_t)y * W + (x + 9) % W];
    cv::Mat imR(H, W, CV_8UC1, rimg.data());
    PL_SLAM::ORBextractor orbR(500, 1.2f, 6, 20, 7);
    std::vector<cv::KeyPoint> kpsR;
    cv::Mat descR;
    orbR(imR, mask, kpsR, descR);
    std::vector<float> uRight, depth;
    PL_SLAM::ComputeStereoMatches(orb, orbR, kps, desc, kpsR, descR, 0.11f, 0.11f * 435.2f, uRight, depth);
    // FLD branch: the second constructor and ComputeFldWithLbd
    PL_SLAM::Lineextractor fe(240, 1, 1.05, 15, 1.732f, 50.0, 100.0, 3, false, false);
    std::vector<PL_SLAM::KeyLine> fkl;
    std::vector<cv::KeyPoint> fmid;
    cv::Mat fld_desc;
    fe.ComputeFldWithLbd(im, fkl, fmid, fld_desc);
    if (fkl.empty() || fkl.size() != fmid.size() || fld_desc.rows != (int)fkl.size()) return 5;
    FILE* o = fopen(argv[4], "wb");
    int hdr[8] = {(int)kps.size(), (int)kl.size(), nm, d01, (int)k2.size(), orb.GetLevels(), orb.mvImagePyramid[1].cols, orb.mvImagePyramid[1].rows};
    fwrite(hdr, sizeof(int), 8, o);
    fwrite(kps.data(), sizeof(cv::KeyPoint), kps.size(), o);
    fwrite(desc.data, 32, kps.size(), o);
    fwrite(kl.data(), sizeof(PL_SLAM::KeyLine), kl.size(), o);
    fwrite(ld.data, 32, kl.size(), o);
    fwrite(m12.data(), sizeof(int), m12.size(), o);
    fwrite(uRight.data(), sizeof(float), uRight.size(), o);
    fwrite(depth.data(), sizeof(float), depth.size(), o);
    fclose(o);
    printf("shim ok: %zu keypoints, %zu lines, %d self-matches\n", kps.size(), kl.size(), nm);
    return 0;
}
