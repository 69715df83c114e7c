/*
 * cvshim.hpp -- a minimal stand-in for the OpenCV C++ API surface that the reference's hot-path sources use.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/).  Purpose: compile the reference's OWN, UNMODIFIED sources
 *   /root/reference/src/ORBextractor.cc, src/Lineextractor.cc,
 *   Thirdparty/line_descriptor/src/LSDDetector_custom.cpp and the compute path of binary_descriptor_custom.cpp
 * into oracle/_ref/libref.so without an OpenCV C++ installation (there is none in this container), so that the C
 * restatement in oracle/*.c and the CUDA kernels can be compared with what the reference's code really computes
 * under this toolchain (overload resolution, float/double promotion, std::sort and heap-address tie orders ...).
 *
 * What is real and what is modelled:
 *   - every line of the reference's own logic is the reference's (compiled from where it lies, never copied);
 *   - the OpenCV *primitives* it calls (resize, copyMakeBorder, GaussianBlur, FAST, fastAtan2, pyrDown, Sobel,
 *     LineSegmentDetector, LineIterator, Canny, fitLine) are implemented in cvshim.cpp over the C models of
 *     oracle/orc_prims.c / orc_lsd.c / orc_fld.c, each of which is pinned bit-for-bit against cv2 4.13 by
 *     tests/test_oracle_vs_cv2.py;
 *   - the container classes below follow OpenCV's observable semantics where the reference depends on them:
 *     reference-counted Mat headers with ROIs, create() that keeps a buffer of matching size and type,
 *     saturate_cast rounding in Point_ conversions, KeyPoint's default field values, and cvstd.hpp's
 *     `using std::sqrt/exp/pow/log/min/max/abs/swap` inside namespace cv (this decides float-vs-double overloads
 *     of unqualified calls made inside namespace cv).
 */
#ifndef PLF_REF_CVSHIM_HPP
#define PLF_REF_CVSHIM_HPP

#include <cstddef>
#include <cstring>
#include <cctype>
#include <string>
#include <algorithm>
#include <utility>
#include <cstdlib>
#include <cmath>
#include <cfloat>
#include <climits>
#include <cassert>
#include <vector>
#include <memory>
#include <stdexcept>
#include <stdint.h>

typedef unsigned char uchar;
typedef signed char schar;
typedef unsigned short ushort;

#define CV_PI 3.1415926535897932384626433832795
#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_WRAP
#define CV_OUT
#define CV_IN_OUT
#define CV_INLINE static inline

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH(flags) ((flags) & 7)
#define CV_MAT_CN(flags) ((((flags) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8SC1 CV_MAKETYPE(CV_8S, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32SC2 CV_MAKETYPE(CV_32S, 2)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr); } while (0)

/* fast_math.hpp: round half to even (SSE cvtsd2si / lrint) */
CV_INLINE int cvRound(double v) { return (int)lrint(v); }
CV_INLINE int cvRound(float v) { return (int)lrintf(v); }
CV_INLINE int cvRound(int v) { return v; }
CV_INLINE int cvFloor(double v) { int i = (int)v; return i - (i > v); }
CV_INLINE int cvFloor(float v) { int i = (int)v; return i - (i > v); }
CV_INLINE int cvFloor(int v) { return v; }
CV_INLINE int cvCeil(double v) { int i = (int)v; return i + (i < v); }
CV_INLINE int cvCeil(float v) { int i = (int)v; return i + (i < v); }
CV_INLINE int cvCeil(int v) { return v; }

namespace cv {

/* cvstd.hpp */
using std::min;
using std::max;
using std::abs;
using std::swap;
using std::sqrt;
using std::exp;
using std::pow;
using std::log;

typedef std::string String;

template <typename T> static inline T saturate_cast(int v) { return T(v); }
template <typename T> static inline T saturate_cast(float v) { return T(v); }
template <typename T> static inline T saturate_cast(double v) { return T(v); }
template <> inline uchar saturate_cast<uchar>(int v) { return (uchar)((unsigned)v <= 255 ? v : v > 0 ? 255 : 0); }
template <> inline int saturate_cast<int>(float v) { return cvRound(v); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <> inline short saturate_cast<short>(int v) { return (short)((unsigned)(v + 32768) <= 65535 ? v : v > 0 ? 32767 : -32768); }

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3, INTER_LINEAR_EXACT = 5 };
enum { COLOR_BGR2GRAY = 6 };
enum { NORM_L2 = 4, NORM_HAMMING = 6 };
enum { DIST_L1 = 1, DIST_L2 = 2 };
enum { LSD_REFINE_NONE = 0, LSD_REFINE_STD = 1, LSD_REFINE_ADV = 2 };

/* ---------------------------------------------------------------- small value types (types.hpp, matx.hpp) */
template <typename T> class Point_ {
public:
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    Point_(const Point_& p) = default;
    Point_& operator=(const Point_& p) = default;
    /* types.hpp: conversion to another coordinate type goes through saturate_cast (float -> int rounds) */
    template <typename T2> operator Point_<T2>() const { return Point_<T2>(saturate_cast<T2>(x), saturate_cast<T2>(y)); }
};
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, float b)
{ a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, double b)
{ a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, int b)
{ a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator+=(Point_<T>& a, const Point_<T>& b) { a.x += b.x; a.y += b.y; return a; }
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b)
{ return Point_<T>(saturate_cast<T>(a.x + b.x), saturate_cast<T>(a.y + b.y)); }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b)
{ return Point_<T>(saturate_cast<T>(a.x - b.x), saturate_cast<T>(a.y - b.y)); }
template <typename T> static inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> static inline bool operator!=(const Point_<T>& a, const Point_<T>& b) { return a.x != b.x || a.y != b.y; }
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point2i Point;

template <typename T> class Size_ {
public:
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
template <typename T> static inline bool operator==(const Size_<T>& a, const Size_<T>& b) { return a.width == b.width && a.height == b.height; }
template <typename T> static inline bool operator!=(const Size_<T>& a, const Size_<T>& b) { return !(a == b); }
typedef Size_<int> Size2i;
typedef Size2i Size;

template <typename T> class Rect_ {
public:
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T, int n> class Vec {
public:
    T val[n];
    Vec() { for (int i = 0; i < n; i++) val[i] = T(0); }
    Vec(T v0, T v1, T v2, T v3) { static_assert(n >= 4, "Vec size"); val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; for (int i = 4; i < n; i++) val[i] = T(0); }
    const T& operator[](int i) const { return val[i]; }
    T& operator[](int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
    T& operator()(int i) { return val[i]; }
};
typedef Vec<float, 4> Vec4f;
typedef Vec<int, 4> Vec4i;

template <typename T> class Scalar_ : public Vec<T, 4> {
public:
    Scalar_() {}
    Scalar_(T v0, T v1, T v2 = 0, T v3 = 0) : Vec<T, 4>(v0, v1, v2, v3) {}
    Scalar_(T v0) : Vec<T, 4>(v0, 0, 0, 0) {}
    static Scalar_<T> all(T v0) { return Scalar_<T>(v0, v0, v0, v0); }
};
typedef Scalar_<double> Scalar;

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType<short> { enum { type = CV_16SC1 }; };
template <> struct DataType<int> { enum { type = CV_32SC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<double> { enum { type = CV_64FC1 }; };
template <> struct DataType<Point2i> { enum { type = CV_32SC2 }; };
template <> struct DataType<Point2f> { enum { type = CV_32FC2 }; };
template <> struct DataType<Vec4f> { enum { type = CV_32FC4 }; };

/* ---------------------------------------------------------------- Mat */
class _OutputArray;
/* Mat::zeros returns an expression; assigning it to a Mat runs create() (which keeps a buffer of matching size and type)
   and then fills -- ORBextractor.cc:1037 relies on this to write descriptors into a row range of the output matrix */
struct MatExpr {
    int rows, cols, type;
    double value;
};
class Mat {
public:
    struct MStep {
        size_t p;
        MStep() : p(0) {}
        operator size_t() const { return p; }
        MStep& operator=(size_t s) { p = s; return *this; }
    };
    int flags;          /* the type (depth + channels) */
    int dims;
    int rows, cols;
    uchar* data;
    MStep step;

    Mat() : flags(0), dims(2), rows(0), cols(0), data(0) {}
    Mat(int r, int c, int t) : flags(0), dims(2), rows(0), cols(0), data(0) { create(r, c, t); }
    Mat(Size s, int t) : flags(0), dims(2), rows(0), cols(0), data(0) { create(s.height, s.width, t); }
    /* user data: a header over the caller's memory, no copy, no ownership */
    Mat(int r, int c, int t, void* d, size_t st = 0) : flags(t), dims(2), rows(r), cols(c), data((uchar*)d)
    { step = st ? st : (size_t)c * elemSize(); }
    /* mat.inl.hpp: Mat(const std::vector<T>&, copyData=false): an N x 1 header over the vector's storage */
    template <typename T> explicit Mat(const std::vector<T>& v) : flags(DataType<T>::type), dims(2), rows((int)v.size()), cols(1),
        data(v.empty() ? 0 : (uchar*)&v[0]) { step = sizeof(T); if (v.empty()) cols = 0; }
    /* reference counting as in OpenCV: the counter lives in the malloc'ed block of the pixels (never operator new) */
    Mat(const Mat& m) : flags(m.flags), dims(m.dims), rows(m.rows), cols(m.cols), data(m.data), step(m.step), refcount(m.refcount)
    { if (refcount) ++*refcount; }
    Mat& operator=(const Mat& m)
    {
        if (this != &m) {
            if (m.refcount) ++*m.refcount;
            unref();
            flags = m.flags; dims = m.dims; rows = m.rows; cols = m.cols; data = m.data; step = m.step; refcount = m.refcount;
        }
        return *this;
    }
    ~Mat() { unref(); }
    Mat& operator=(const Scalar& s);
    Mat(const MatExpr& e) : flags(0), dims(2), rows(0), cols(0), data(0) { *this = e; }
    Mat& operator=(const MatExpr& e) { create(e.rows, e.cols, e.type); *this = Scalar(e.value); return *this; }

    void create(int r, int c, int t);
    void create(Size s, int t) { create(s.height, s.width, t); }
    void release() { unref(); data = 0; rows = cols = 0; step = 0; }
    Mat clone() const;
    void copyTo(Mat& dst) const;
    void copyTo(const _OutputArray& dst) const;
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
    static MatExpr zeros(int r, int c, int t) { MatExpr e = {r, c, t, 0.0}; return e; }

    Mat operator()(const Rect& r) const { Mat m(*this); m.data = data + (size_t)r.y * step.p + (size_t)r.x * elemSize(); m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }

    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    int type() const { return flags & 4095; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize1() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return sz[depth()]; }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t step1() const { return step.p / elemSize1(); }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return rows <= 1 || step.p == (size_t)cols * elemSize(); }

    uchar* ptr(int r = 0) { return data + (size_t)r * step.p; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step.p; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.p); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.p); }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step.p))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step.p))[c]; }
    template <typename T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step.p); }
    template <typename T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step.p); }

    Mat cross(const Mat& m) const;
    double dot(const Mat& m) const;

private:
    void unref() { if (refcount && --*refcount == 0) free(refcount); refcount = 0; }
    int* refcount = 0;
};

Mat operator*(const Mat& a, const Mat& b);   // CV_32F / CV_64F matrix product
Mat operator+(const Mat& a, const Mat& b);

template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    T* operator[](int r) { return (T*)ptr(r); }
    const T* operator[](int r) const { return (const T*)ptr(r); }
};

/* ---------------------------------------------------------------- InputArray / OutputArray proxies */
class _InputArray {
public:
    _InputArray() : m(0) {}
    _InputArray(const Mat& mat) : m(&mat) {}
    template <typename T> _InputArray(const std::vector<T>& v) : own(new Mat(v)), m(own.get()) {}
    Mat getMat() const { return m ? *m : Mat(); }
    bool empty() const { return !m || m->empty(); }
protected:
    std::shared_ptr<Mat> own;
    const Mat* m;
};
class _OutputArray {
public:
    _OutputArray() : mat(0), vec(0), vec_resize(0), vec_size(0), esz(0), vtype(0) {}
    _OutputArray(Mat& m_) : mat(&m_), vec(0), vec_resize(0), vec_size(0), esz(0), vtype(0) {}
    template <typename T> _OutputArray(std::vector<T>& v) : mat(0), vec(&v), vec_resize(&resize_vec<T>), vec_size(&size_vec<T>), esz(sizeof(T)), vtype(DataType<T>::type) {}
    /* a fixed-size Vec<T, n> destination (matx): create() only checks the size */
    template <typename T, int n> _OutputArray(Vec<T, n>& v) : mat(0), vec(0), vec_resize(0), vec_size(0), esz(sizeof(T)), vtype(DataType<T>::type), fixed(v.val), nfixed(n) {}
    void create(int r, int c, int t) const;
    void create(Size s, int t) const { create(s.height, s.width, t); }
    void release() const;
    Mat getMat() const;
    bool needed() const { return mat || vec || fixed; }
private:
    template <typename T> static void* resize_vec(void* v, size_t n) { std::vector<T>* p = (std::vector<T>*)v; p->resize(n); return n ? (void*)&(*p)[0] : 0; }
    template <typename T> static size_t size_vec(void* v, void** d) { std::vector<T>* p = (std::vector<T>*)v; *d = p->empty() ? 0 : (void*)&(*p)[0]; return p->size(); }
    Mat* mat;
    void* vec;
    void* (*vec_resize)(void*, size_t);
    size_t (*vec_size)(void*, void**);
    size_t esz;
    int vtype;
    void* fixed = 0;
    int nfixed = 0;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
static inline InputArray noArray() { static _InputArray none; return none; }

/* ---------------------------------------------------------------- features2d value types */
class KeyPoint {
public:
    Point2f pt;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f _pt, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(_pt), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};
class KeyPointsFilter {
public:
    /* only reachable from ORBextractor::ComputeKeyPointsOld, which operator() never calls */
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

/* ---------------------------------------------------------------- Ptr, Algorithm, persistence (declarations only) */
template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> static inline Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

class FileNode {
public:
    FileNode operator[](const char*) const { return FileNode(); }
    operator int() const { return 0; }
};
class FileStorage {};
template <typename T> static inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

class Algorithm {
public:
    virtual ~Algorithm() {}
    virtual void read(const FileNode&) {}
    virtual void write(FileStorage&) const {}
};

/* ---------------------------------------------------------------- imgproc / features2d functions (cvshim.cpp) */
float fastAtan2(float y, float x);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar& value = Scalar());
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void pyrDown(InputArray src, OutputArray dst, const Size& dstsize = Size(), int borderType = BORDER_DEFAULT);
void Sobel(InputArray src, OutputArray dst, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0,
           int borderType = BORDER_DEFAULT);
void cvtColor(InputArray src, OutputArray dst, int code, int dstCn = 0);
void Canny(InputArray image, OutputArray edges, double threshold1, double threshold2, int apertureSize = 3, bool L2gradient = false);
void fitLine(InputArray points, OutputArray line, int distType, double param, double reps, double aeps);
void line(Mat& img, Point pt1, Point pt2, const Scalar& color, int thickness = 1, int lineType = 8, int shift = 0);

class LineIterator {
public:
    /* 8-connected: count = max(|dx|, |dy|) + 1 on the (rounded, clipped) end points */
    LineIterator(const Mat& img, Point pt1, Point pt2, int connectivity = 8, bool leftToRight = false);
    int count;
};

class LineSegmentDetector : public Algorithm {
public:
    virtual void detect(InputArray image, OutputArray lines, OutputArray width = _OutputArray(), OutputArray prec = _OutputArray(),
                        OutputArray nfa = _OutputArray()) = 0;
    virtual ~LineSegmentDetector() {}
};
Ptr<LineSegmentDetector> createLineSegmentDetector(int refine = LSD_REFINE_STD, double scale = 0.8, double sigma_scale = 0.6,
                                                   double quant = 2.0, double ang_th = 22.5, double log_eps = 0,
                                                   double density_th = 0.7, int n_bins = 1024);

class DescriptorMatcher;
class BFMatcher;

} // namespace cv

#endif
