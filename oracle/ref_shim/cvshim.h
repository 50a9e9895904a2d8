// Minimal stand-in for the few OpenCV / MatP types that two reference toolbox
// translation units (toolbox/rgbConvertMex.cpp, toolbox/imResampleMex.cpp) mention.
// TEST INFRASTRUCTURE ONLY: lets oracle/Makefile compile the reference's own
// arithmetic from /root/reference into oracle/_ref without OpenCV installed.
// Only the raw-pointer template functions (resample<float>, rgbConvert<float,float>)
// are ever called through oracle/ref_glue.cpp; the MatP-taking wrappers just need to parse.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <vector>
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_Assert(x) do { if (!(x)) throw std::runtime_error("CV_Assert: " #x); } while (0)
#ifndef ACF_EXPORT
#define ACF_EXPORT
#endif
typedef uint32_t uint32;
namespace cv {
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} int area() const { return width * height; } };
struct Mat {
    unsigned char* data = nullptr; int rows = 0, cols = 0, tp = CV_32F;
    int type() const { return tp; } int depth() const { return tp; }
    void convertTo(Mat&, int) const { throw std::runtime_error("cvshim: convertTo unavailable"); }
};
}
class MatP {
public:
    MatP() {}
    MatP(const cv::Size&, int, int) {}
    void create(const cv::Size&, int, int) {}
    cv::Size size() const { return {}; }
    int rows() const { return 0; } int cols() const { return 0; }
    int depth() const { return CV_32F; } int channels() const { return 0; }
    template <class T> T* ptr() const { return nullptr; }
    void* ptr() const { return nullptr; }
    cv::Mat& base() { return b; } const cv::Mat& base() const { return b; }
    cv::Mat& operator[](int) { return b; } const cv::Mat& operator[](int) const { return b; }
    void pop_back() {}
private:
    cv::Mat b;
};
