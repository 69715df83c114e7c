/*
 * cvshim.cpp -- implementation of the test-only cv:: stand-in (see cvshim.hpp).
 *
 * TEST INFRASTRUCTURE ONLY.  The primitives forward to the C models of oracle/orc_prims.c, orc_lsd.c and orc_fld.c,
 * which tests/test_oracle_vs_cv2.py pins bit-for-bit against cv2 4.13; the container semantics (create() keeping a
 * matching buffer, in-place filters, ROI steps) follow OpenCV's so that the reference's code sees what it expects.
 */
#include "cvshim.hpp"
#include "../../plf_oracle.h"

namespace cv {

/* ------------------------------------------------------------------------------------------------ Mat */
void Mat::create(int r, int c, int t)
{
    t &= 4095;
    if (data && rows == r && cols == c && type() == t) return; /* OpenCV: same size and type -> keep the buffer */
    flags = t;
    rows = r;
    cols = c;
    size_t es = elemSize();
    step = (size_t)c * es;
    size_t bytes = (size_t)r * step.p;
    unref();
    /* one block: [refcount, padding to 64 bytes][pixels] */
    refcount = (int*)malloc(64 + (bytes ? bytes : 1));
    *refcount = 1;
    data = (uchar*)refcount + 64;
}

Mat Mat::clone() const
{
    Mat m;
    copyTo(m);
    return m;
}

void Mat::copyTo(Mat& dst) const
{
    if (empty()) { dst.release(); return; }
    if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.step.p == step.p) { dst.flags = flags; return; }
    dst.create(rows, cols, type());
    size_t rowbytes = (size_t)cols * elemSize();
    for (int y = 0; y < rows; y++) memmove(dst.ptr(y), ptr(y), rowbytes);
}

void Mat::copyTo(const _OutputArray& dst) const
{
    if (empty()) { dst.release(); return; }
    dst.create(rows, cols, type());
    Mat d = dst.getMat();
    size_t rowbytes = (size_t)cols * elemSize();
    for (int y = 0; y < rows; y++) memmove(d.ptr(y), ptr(y), rowbytes);
}

Mat& Mat::operator=(const Scalar& s)
{
    CV_Assert(channels() == 1);
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++)
            switch (depth()) {
            case CV_8U: at<uchar>(y, x) = saturate_cast<uchar>(cvRound(s[0])); break;
            case CV_16S: at<short>(y, x) = saturate_cast<short>(cvRound(s[0])); break;
            case CV_32S: at<int>(y, x) = cvRound(s[0]); break;
            case CV_32F: at<float>(y, x) = (float)s[0]; break;
            case CV_64F: at<double>(y, x) = s[0]; break;
            default: throw std::runtime_error("cvshim: Mat = Scalar for this depth");
            }
    return *this;
}

/* convert_scale: only the conversions the reference performs (8U -> 8U copy, 32S -> 32F, 64F scaled in place) */
void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const
{
    int ddepth = rtype < 0 ? depth() : CV_MAT_DEPTH(rtype);
    int dtype = CV_MAKETYPE(ddepth, channels());
    bool noscale = alpha == 1 && beta == 0;
    if (ddepth == depth() && noscale) { copyTo(dst); return; }
    Mat src = *this; /* keeps the source alive if dst aliases it and is re-created */
    Mat out;
    if (dst.data == data && dtype == type()) out = dst; else out.create(rows, cols, dtype);
    int n = cols * channels();
    for (int y = 0; y < rows; y++) {
        if (depth() == CV_64F && ddepth == CV_64F) {
            const double* s = src.ptr<double>(y); double* d = out.ptr<double>(y);
            for (int x = 0; x < n; x++) d[x] = s[x] * alpha + beta;
        } else if (depth() == CV_32S && ddepth == CV_32F && noscale) {
            const int* s = src.ptr<int>(y); float* d = out.ptr<float>(y);
            for (int x = 0; x < n; x++) d[x] = (float)s[x];
        } else
            throw std::runtime_error("cvshim: convertTo for these depths");
    }
    dst = out;
}

/* matmul.cpp: Mat::cross for 3-element vectors, Mat::dot = sequential sum of products */
Mat Mat::cross(const Mat& m) const
{
    CV_Assert(type() == CV_64FC1 && m.type() == CV_64FC1 && total() == 3 && m.total() == 3);
    Mat r(rows, cols, type());
    const double a0 = at<double>(0), a1 = at<double>(1), a2 = at<double>(2);
    const double b0 = m.at<double>(0), b1 = m.at<double>(1), b2 = m.at<double>(2);
    r.at<double>(0) = a1 * b2 - a2 * b1;
    r.at<double>(1) = a2 * b0 - a0 * b2;
    r.at<double>(2) = a0 * b1 - a1 * b0;
    return r;
}

double Mat::dot(const Mat& m) const
{
    CV_Assert(type() == CV_64FC1 && m.type() == CV_64FC1 && total() == m.total());
    double r = 0;
    int n = (int)total(), i = 0;
    for (; i <= n - 4; i += 4)
        r += at<double>(i) * m.at<double>(i) + at<double>(i + 1) * m.at<double>(i + 1) + at<double>(i + 2) * m.at<double>(i + 2) +
             at<double>(i + 3) * m.at<double>(i + 3);
    for (; i < n; i++) r += at<double>(i) * m.at<double>(i);
    return r;
}

/* core/src/matmul: for the small float matrices SPL-SLAM multiplies (3x3 by 3x1) OpenCV's gemm takes its unrolled path:
 * every output element is the float sum a0*b0 + a1*b1 + a2*b2 taken left to right */
Mat operator*(const Mat& a, const Mat& b)
{
    CV_Assert(a.type() == b.type() && (a.type() == CV_32FC1 || a.type() == CV_64FC1) && a.cols == b.rows);
    Mat r(a.rows, b.cols, a.type());
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            if (a.type() == CV_32FC1) {
                float t = 0;
                for (int k = 0; k < a.cols; k++) { float p = a.at<float>(i, k) * b.at<float>(k, j); t = k ? t + p : p; }
                r.at<float>(i, j) = t;
            } else {
                double t = 0;
                for (int k = 0; k < a.cols; k++) { double p = a.at<double>(i, k) * b.at<double>(k, j); t = k ? t + p : p; }
                r.at<double>(i, j) = t;
            }
        }
    return r;
}
Mat operator+(const Mat& a, const Mat& b)
{
    CV_Assert(a.type() == b.type() && (a.type() == CV_32FC1 || a.type() == CV_64FC1) && a.rows == b.rows && a.cols == b.cols);
    Mat r(a.rows, a.cols, a.type());
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) {
            if (a.type() == CV_32FC1) r.at<float>(i, j) = a.at<float>(i, j) + b.at<float>(i, j);
            else r.at<double>(i, j) = a.at<double>(i, j) + b.at<double>(i, j);
        }
    return r;
}

/* ------------------------------------------------------------------------------------------------ OutputArray */
void _OutputArray::create(int r, int c, int t) const
{
    if (mat) { mat->create(r, c, t); return; }
    if (fixed) { CV_Assert((t & 4095) == vtype && r * c == nfixed); return; }
    if (vec) {
        CV_Assert((t & 4095) == vtype && (r == 1 || c == 1));
        vec_resize(vec, (size_t)r * c);
        return;
    }
    throw std::runtime_error("cvshim: create() on an empty OutputArray");
}
void _OutputArray::release() const
{
    if (mat) mat->release();
    else if (vec) vec_resize(vec, 0);
}
Mat _OutputArray::getMat() const
{
    if (mat) return *mat;
    if (fixed) return Mat(nfixed, 1, vtype, fixed, esz);
    if (vec) {
        void* d = 0;
        size_t n = vec_size(vec, &d);
        return n ? Mat((int)n, 1, vtype, d, esz) : Mat();
    }
    return Mat();
}

void KeyPointsFilter::retainBest(std::vector<KeyPoint>&, int)
{
    throw std::runtime_error("cvshim: KeyPointsFilter::retainBest is not on the path (ComputeKeyPointsOld)");
}

/* ------------------------------------------------------------------------------------------------ primitives */
static Mat require_u8(InputArray a, const char* who)
{
    Mat m = a.getMat();
    if (m.empty() || m.type() != CV_8UC1) throw std::runtime_error(std::string("cvshim: ") + who + " expects a non-empty CV_8UC1 image");
    return m;
}

/* writes `tmp` (dense w x h bytes of element size es) into dst through create(), OpenCV style */
static void store(OutputArray dst, const void* tmp, int w, int h, int type)
{
    Mat hdr(h, w, type, (void*)tmp);
    hdr.copyTo(dst);
}

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression)
{
    CV_Assert(nonmaxSuppression);
    Mat img = require_u8(image, "FAST");
    keypoints.clear();
    int cap = img.rows * img.cols;
    std::vector<int> xs(cap), ys(cap), sc(cap);
    int n = orc_fast9(img.data, img.cols, img.rows, img.step, threshold, xs.data(), ys.data(), sc.data(), cap);
    for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));
}

void resize(InputArray src, OutputArray dst, Size dsize, double fx, double fy, int interpolation)
{
    CV_Assert(interpolation == INTER_LINEAR && dsize.width > 0 && dsize.height > 0 && fx == 0 && fy == 0);
    Mat s = require_u8(src, "resize");
    dst.create(dsize.height, dsize.width, CV_8UC1);
    Mat d = dst.getMat();
    CV_Assert(d.data != s.data);
    orc_resize_linear_u8(s.data, s.cols, s.rows, s.step, d.data, d.cols, d.rows, d.step);
}

void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar&)
{
    CV_Assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101 && top == bottom && left == right && top == left);
    Mat s = require_u8(src, "copyMakeBorder");
    dst.create(s.rows + top + bottom, s.cols + left + right, CV_8UC1);
    Mat d = dst.getMat();
    /* the source may be the interior ROI of the destination (ORBextractor.cc:1122): rows are then copied onto themselves */
    orc_border_reflect101_u8(s.data, s.cols, s.rows, s.step, d.data, top, d.step);
}

void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY, int borderType)
{
    CV_Assert(ksize.width == ksize.height && (sigmaY == 0 || sigmaY == sigmaX) && borderType == BORDER_REFLECT_101);
    Mat s = require_u8(src, "GaussianBlur");
    std::vector<uchar> tmp((size_t)s.rows * s.cols);
    orc_gauss_blur_u8(s.data, s.cols, s.rows, s.step, tmp.data(), (size_t)s.cols, ksize.width, sigmaX);
    store(dst, tmp.data(), s.cols, s.rows, CV_8UC1);
}

void pyrDown(InputArray src, OutputArray dst, const Size& dstsize, int borderType)
{
    Mat s = require_u8(src, "pyrDown");
    CV_Assert(borderType == BORDER_DEFAULT);
    Size ds = dstsize.width > 0 ? dstsize : Size((s.cols + 1) / 2, (s.rows + 1) / 2);
    /* the model covers the sizes the reference asks for: (w/2, h/2) */
    CV_Assert(ds.width == s.cols / 2 && ds.height == s.rows / 2);
    std::vector<uchar> tmp((size_t)std::max(ds.width, 1) * std::max(ds.height, 1));
    orc_pyrdown_u8(s.data, s.cols, s.rows, s.step, tmp.data(), (size_t)ds.width);
    store(dst, tmp.data(), ds.width, ds.height, CV_8UC1);
}

void Sobel(InputArray src, OutputArray dst, int ddepth, int dx, int dy, int ksize, double scale, double delta, int borderType)
{
    CV_Assert(CV_MAT_DEPTH(ddepth) == CV_16S && ksize == 3 && scale == 1 && delta == 0 && borderType == BORDER_DEFAULT &&
              ((dx == 1 && dy == 0) || (dx == 0 && dy == 1)));
    Mat s = require_u8(src, "Sobel");
    std::vector<int16_t> gx((size_t)s.rows * s.cols), gy((size_t)s.rows * s.cols);
    orc_sobel3_s16(s.data, s.cols, s.rows, s.step, gx.data(), gy.data());
    store(dst, dx ? gx.data() : gy.data(), s.cols, s.rows, CV_16SC1);
}

void Canny(InputArray image, OutputArray edges, double threshold1, double threshold2, int apertureSize, bool L2gradient)
{
    CV_Assert(apertureSize == 3 && !L2gradient);
    Mat s = require_u8(image, "Canny");
    std::vector<uchar> tmp((size_t)s.rows * s.cols);
    orc_canny_u8(s.data, s.cols, s.rows, s.step, threshold1, threshold2, tmp.data());
    store(edges, tmp.data(), s.cols, s.rows, CV_8UC1);
}

void fitLine(InputArray points, OutputArray line, int distType, double param, double, double)
{
    CV_Assert(distType == DIST_L2 && param == 0);
    Mat p = points.getMat();
    CV_Assert(p.type() == CV_32SC2 && p.cols == 1 && p.rows > 0 && p.step == 8);
    float l[4];
    orc_fit_line_l2((const int32_t*)p.data, p.rows, l);
    Mat(4, 1, CV_32FC1, l).copyTo(line);
}

void cvtColor(InputArray, OutputArray, int, int)
{
    throw std::runtime_error("cvshim: cvtColor is not on the path (the reference passes single-channel images)");
}

void line(Mat&, Point, Point, const Scalar&, int, int, int)
{
    throw std::runtime_error("cvshim: cv::line is not on the path (Lineextractor::drawSegment)");
}

LineIterator::LineIterator(const Mat& img, Point pt1, Point pt2, int connectivity, bool)
{
    CV_Assert(connectivity == 8);
    /* imgproc/drawing.cpp clips the segment to the image first; the reference clamps the end points into the image before
       it gets here, so clipping is the identity -- checked rather than modelled */
    CV_Assert(pt1.x >= 0 && pt1.y >= 0 && pt2.x >= 0 && pt2.y >= 0 && pt1.x < img.cols && pt2.x < img.cols && pt1.y < img.rows &&
              pt2.y < img.rows);
    int dx = std::abs(pt2.x - pt1.x), dy = std::abs(pt2.y - pt1.y);
    count = std::max(dx, dy) + 1;
}

namespace {
class LineSegmentDetectorShim : public LineSegmentDetector {
public:
    LineSegmentDetectorShim(int refine_, double scale_, double sigma_scale_, double quant_, double ang_th_, int n_bins_)
        : refine(refine_), scale(scale_), sigma_scale(sigma_scale_), quant(quant_), ang_th(ang_th_), n_bins(n_bins_) {}
    void detect(InputArray image, OutputArray lines, OutputArray, OutputArray, OutputArray) override
    {
        if (refine != LSD_REFINE_NONE) throw std::runtime_error("cvshim: LineSegmentDetector is modelled for refine = LSD_REFINE_NONE only");
        Mat img = require_u8(image, "LineSegmentDetector::detect");
        int cap = 1 << 16;
        std::vector<float> buf((size_t)cap * 4);
        int n = orc_lsd_detect(img.data, img.cols, img.rows, img.step, scale, sigma_scale, quant, ang_th, n_bins, buf.data(), cap);
        if (n > cap) throw std::runtime_error("cvshim: more than 65536 line segments");
        if (n == 0) { lines.release(); return; }
        std::vector<Vec4f> v(n);
        for (int i = 0; i < n; i++) v[i] = Vec4f(buf[4 * i], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3]);
        Mat(v).copyTo(lines);
    }
private:
    int refine;
    double scale, sigma_scale, quant, ang_th;
    int n_bins;
};
} // namespace

Ptr<LineSegmentDetector> createLineSegmentDetector(int refine, double scale, double sigma_scale, double quant, double ang_th, double,
                                                   double, int n_bins)
{
    return Ptr<LineSegmentDetector>(new LineSegmentDetectorShim(refine, scale, sigma_scale, quant, ang_th, n_bins));
}

} // namespace cv
