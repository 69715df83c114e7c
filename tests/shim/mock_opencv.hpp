// Minimal layout-compatible stand-ins for the OpenCV types the shim touches, so the drop-in header can be
// compiled and exercised in an image without OpenCV C++ headers.  TEST ONLY.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CV_8U 0
#define CV_8UC1 0
namespace cv {
struct Point2f { float x, y; };
struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : pt{0, 0}, size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
    std::vector<unsigned char> store;
    Mat() {}
    Mat(int r, int c, int, void* d, size_t s = 0) : rows(r), cols(c), step(s ? s : (size_t)c), data((unsigned char*)d) {}
    void create(int r, int c, int) { rows = r; cols = c; step = (size_t)c; store.assign((size_t)r * c, 0); data = store.data(); }
    void release() { rows = cols = 0; data = nullptr; store.clear(); }
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    int type() const { return CV_8UC1; }
    int depth() const { return 0; }
    int channels() const { return 1; }
    bool isContinuous() const { return step == (size_t)cols; }
    unsigned char* ptr(int r = 0) { return data + (size_t)r * step; }
    const unsigned char* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <class T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    // float element access for the 3x3 camera matrix / the distortion vector (the mock stores them as rows x cols floats, step in bytes)
    template <class T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    template <class T> const T& at(int i) const { return ((const T*)data)[i]; }
    Mat getMat() const { return *this; }
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
namespace line_descriptor {
struct KeyLine {
    float angle; int class_id; int octave; Point2f pt; float response; float size;
    float startPointX, startPointY, endPointX, endPointY;
    float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength; int numOfPixels;
};
}
}
