// cvgs_opencv_compat.hpp -- the handful of OpenCV types the cvGS interface touches.
//
// With OpenCV present (<opencv2/core/cuda.hpp> found) this header only includes it.  Without it (this image
// has no OpenCV C++ headers or libraries) it provides a minimal stand-in with the same names and members, so
// that code written against the reference's include/cvGPUSpeedup.cuh compiles unchanged: the reference
// wrapper only ever reads GpuMat::data/cols/rows/step/type() and unwraps the cudaStream_t of a
// cv::cuda::Stream (reference include/cvGPUSpeedup.cuh:36,42,69,466,615).
#pragma once

#if defined(__has_include)
#if __has_include(<opencv2/core/cuda.hpp>) && !defined(CVGS_FORCE_OPENCV_DOUBLE)
#define CVGS_HAVE_OPENCV 1
#endif
#endif

#ifdef CVGS_HAVE_OPENCV
#include <opencv2/core.hpp>
#include <opencv2/core/cuda.hpp>
#include <opencv2/core/cuda_stream_accessor.hpp>
#include <opencv2/imgproc.hpp>
#else
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned int uint;

#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH_MASK (CV_DEPTH_MAX - 1)
#define CV_MAT_DEPTH(flags) ((flags) & CV_MAT_DEPTH_MASK)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAT_CN(flags) ((((flags) >> CV_CN_SHIFT) & 511) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC3 CV_MAKETYPE(CV_16U, 3)
#define CV_16SC3 CV_MAKETYPE(CV_16S, 3)
#ifndef CV_8UC4
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#endif
#define CV_16UC4 CV_MAKETYPE(CV_16U, 4)
#define CV_16SC4 CV_MAKETYPE(CV_16S, 4)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)

namespace cv {

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2 };
enum ColorConversionCodes { COLOR_BGR2BGRA = 0, COLOR_BGRA2BGR = 1, COLOR_BGR2RGBA = 2, COLOR_RGBA2BGR = 3,
                            COLOR_BGR2RGB = 4, COLOR_RGB2BGR = COLOR_BGR2RGB, COLOR_BGRA2RGBA = 5,
                            COLOR_RGBA2BGRA = COLOR_BGRA2RGBA, COLOR_RGB2RGBA = COLOR_BGR2BGRA,
                            COLOR_RGBA2RGB = COLOR_BGRA2BGR, COLOR_RGB2BGRA = COLOR_BGR2RGBA, COLOR_BGRA2RGB = COLOR_RGBA2BGR,
                            COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7, COLOR_BGRA2GRAY = 10, COLOR_RGBA2GRAY = 11 };

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() = default;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Rect2d {
    double x = 0, y = 0, width = 0, height = 0;
    Rect2d() = default;
    Rect2d(double x_, double y_, double w, double h) : x(x_), y(y_), width(w), height(h) {}
};
struct Scalar {
    double val[4] = {0, 0, 0, 0};
    Scalar() = default;
    Scalar(double v0, double v1 = 0, double v2 = 0, double v3 = 0) : val{v0, v1, v2, v3} {}
    double operator[](int i) const { return val[i]; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
};

struct Point2f {
    float x = 0, y = 0;
    Point2f() = default;
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};

// The transform matrices of cvGS::warp: small CV_64FC1 host matrices (cv::Mat_<double>(2, 3) << ..., Mat::inv(),
// ptr<double>()).  Nothing else of cv::Mat is provided.
class Mat {
public:
    int rows = 0, cols = 0;
    Mat() = default;
    Mat(int r, int c, int type = CV_64FC1) : rows(r), cols(c), v_(static_cast<size_t>(r) * c, 0.0) {
        if (type != CV_64FC1) throw std::runtime_error("cv::Mat stand-in: CV_64FC1 only");
    }
    int type() const { return CV_64FC1; }
    bool empty() const { return v_.empty(); }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(v_.data() + static_cast<size_t>(r) * cols); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(v_.data() + static_cast<size_t>(r) * cols); }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    // 3x3 closed form (what cv::invert does for 3x3 CV_64F with DECOMP_LU); a singular matrix gives zeros
    Mat inv() const {
        if (rows != 3 || cols != 3) throw std::runtime_error("cv::Mat stand-in: inv() of 3x3 matrices only");
        const double* s = v_.data();
        Mat r(3, 3);
        double d = s[0] * (s[4] * s[8] - s[5] * s[7]) - s[1] * (s[3] * s[8] - s[5] * s[6]) + s[2] * (s[3] * s[7] - s[4] * s[6]);
        if (d == 0.) return r;
        d = 1. / d;
        double* t = r.v_.data();
        t[0] = (s[4] * s[8] - s[5] * s[7]) * d;
        t[1] = (s[2] * s[7] - s[1] * s[8]) * d;
        t[2] = (s[1] * s[5] - s[2] * s[4]) * d;
        t[3] = (s[5] * s[6] - s[3] * s[8]) * d;
        t[4] = (s[0] * s[8] - s[2] * s[6]) * d;
        t[5] = (s[2] * s[3] - s[0] * s[5]) * d;
        t[6] = (s[3] * s[7] - s[4] * s[6]) * d;
        t[7] = (s[1] * s[6] - s[0] * s[7]) * d;
        t[8] = (s[0] * s[4] - s[1] * s[3]) * d;
        return r;
    }

protected:
    std::vector<double> v_;
};
template <typename T>
class Mat_ : public Mat {
    static_assert(sizeof(T) == sizeof(double), "cv::Mat_ stand-in: double only");
    struct Filler {  // cv::MatCommaInitializer_
        Mat_* m;
        int i;
        Filler& operator,(double v) {
            if (i < m->rows * m->cols) m->template ptr<double>()[i++] = v;
            return *this;
        }
        operator Mat() const { return *m; }
    };

public:
    Mat_(int r, int c) : Mat(r, c) {}
    Filler operator<<(double v) {
        this->template ptr<double>()[0] = v;
        return Filler{this, 1};
    }
};

// cv::invertAffineTransform (imgproc/src/imgwarp.cpp): closed form in double; D == 0 gives a zero matrix
inline void invertAffineTransform(const Mat& m, Mat& im) {
    if (m.rows != 2 || m.cols != 3) throw std::runtime_error("invertAffineTransform: 2x3 matrix required");
    const double* M = m.ptr<double>();
    im = Mat(2, 3);
    double* iM = im.ptr<double>();
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D, A12 = -M[1] * D, A21 = -M[3] * D;
    const double b1 = -A11 * M[2] - A12 * M[5];
    const double b2 = -A21 * M[2] - A22 * M[5];
    iM[0] = A11; iM[1] = A12; iM[2] = b1;
    iM[3] = A21; iM[4] = A22; iM[5] = b2;
}

// cv::getPerspectiveTransform: the 8x8 system of the four point pairs, solved in double (partial pivoting)
inline Mat getPerspectiveTransform(const Point2f src[], const Point2f dst[]) {
    double a[8][9];
    for (int i = 0; i < 4; ++i) {
        const double x = src[i].x, y = src[i].y, u = dst[i].x, v = dst[i].y;
        const double r0[9] = {x, y, 1, 0, 0, 0, -x * u, -y * u, u};
        const double r1[9] = {0, 0, 0, x, y, 1, -x * v, -y * v, v};
        for (int k = 0; k < 9; ++k) { a[i][k] = r0[k]; a[i + 4][k] = r1[k]; }
    }
    for (int c = 0; c < 8; ++c) {
        int piv = c;
        for (int r = c + 1; r < 8; ++r)
            if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
        if (a[piv][c] == 0.) throw std::runtime_error("getPerspectiveTransform: degenerate points");
        for (int k = 0; k < 9; ++k) std::swap(a[c][k], a[piv][k]);
        for (int r = c + 1; r < 8; ++r) {
            const double f = a[r][c] / a[c][c];
            for (int k = c; k < 9; ++k) a[r][k] -= f * a[c][k];
        }
    }
    double h[9];
    for (int r = 7; r >= 0; --r) {
        double acc = a[r][8];
        for (int k = r + 1; k < 8; ++k) acc -= a[r][k] * h[k];
        h[r] = acc / a[r][r];
    }
    h[8] = 1.;
    Mat m(3, 3);
    for (int k = 0; k < 9; ++k) m.ptr<double>()[k] = h[k];
    return m;
}

namespace cuda {

class GpuMat {
public:
    int flags = 0, rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;
    uchar* datastart = nullptr;
    uchar* dataend = nullptr;

    GpuMat() = default;
    GpuMat(int rows_, int cols_, int type_) { create(rows_, cols_, type_); }
    GpuMat(Size s, int type_) { create(s.height, s.width, type_); }
    GpuMat(int rows_, int cols_, int type_, Scalar s) {
        create(rows_, cols_, type_);
        setTo(s);
    }
    // user-allocated memory, like the OpenCV constructor of the same shape
    GpuMat(int rows_, int cols_, int type_, void* data_, size_t step_)
        : flags(type_), rows(rows_), cols(cols_), step(step_), data(static_cast<uchar*>(data_)),
          datastart(static_cast<uchar*>(data_)), dataend(static_cast<uchar*>(data_) + step_ * rows_) {}
    // ROI
    GpuMat(const GpuMat& m, Rect r)
        : flags(m.flags), rows(r.height), cols(r.width), step(m.step), data(m.data + r.y * m.step + r.x * m.elemSize()),
          datastart(m.datastart), dataend(m.dataend), owner_(m.owner_) {
        if (r.x < 0 || r.y < 0 || r.width < 0 || r.height < 0 || r.x + r.width > m.cols || r.y + r.height > m.rows)
            throw std::runtime_error("GpuMat ROI outside the matrix");
    }
    GpuMat operator()(Rect r) const { return GpuMat(*this, r); }

    void create(int rows_, int cols_, int type_) {
        flags = type_;
        rows = rows_;
        cols = cols_;
        void* p = nullptr;
        size_t pitch = 0;
        if (cudaMallocPitch(&p, &pitch, elemSize() * static_cast<size_t>(cols_), static_cast<size_t>(rows_)) != cudaSuccess)
            throw std::runtime_error("GpuMat: cudaMallocPitch failed");
        owner_ = std::shared_ptr<void>(p, [](void* q) { cudaFree(q); });
        data = datastart = static_cast<uchar*>(p);
        step = pitch;
        dataend = data + step * rows;
    }
    int type() const { return flags & 4095; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize1() const {
        const int d = depth();
        return d <= CV_8S ? 1 : d <= CV_16S ? 2 : d <= CV_32F ? 4 : 8;
    }
    size_t elemSize() const { return elemSize1() * channels(); }
    bool empty() const { return data == nullptr; }
    Size size() const { return Size(cols, rows); }

    // 8U and 32F only: enough for the tests of this path
    void setTo(Scalar s) {
        const size_t n = static_cast<size_t>(cols) * channels();
        std::unique_ptr<uchar[]> row(new uchar[n * elemSize1()]);
        for (size_t i = 0; i < n; ++i) {
            const double v = s[static_cast<int>(i % channels())];
            if (depth() == CV_8U) row[i] = static_cast<uchar>(v);
            else if (depth() == CV_32F) reinterpret_cast<float*>(row.get())[i] = static_cast<float>(v);
            else throw std::runtime_error("GpuMat::setTo: depth not supported by the stand-in");
        }
        for (int y = 0; y < rows; ++y)
            if (cudaMemcpy(data + y * step, row.get(), n * elemSize1(), cudaMemcpyHostToDevice) != cudaSuccess)
                throw std::runtime_error("GpuMat::setTo: cudaMemcpy failed");
    }
    void upload(const void* host, size_t host_step) {
        if (cudaMemcpy2D(data, step, host, host_step, elemSize() * cols, rows, cudaMemcpyHostToDevice) != cudaSuccess)
            throw std::runtime_error("GpuMat::upload failed");
    }
    void download(void* host, size_t host_step) const {
        if (cudaMemcpy2D(host, host_step, data, step, elemSize() * cols, rows, cudaMemcpyDeviceToHost) != cudaSuccess)
            throw std::runtime_error("GpuMat::download failed");
    }

private:
    std::shared_ptr<void> owner_;
};

class Stream {
public:
    Stream() = default;
    explicit Stream(cudaStream_t s) : s_(s) {}
    void waitForCompletion() const {
        if (cudaStreamSynchronize(s_) != cudaSuccess) throw std::runtime_error("Stream::waitForCompletion failed");
    }
    cudaStream_t raw() const { return s_; }

private:
    cudaStream_t s_ = nullptr;
};
struct StreamAccessor {
    static cudaStream_t getStream(const Stream& s) { return s.raw(); }
    static Stream wrapStream(cudaStream_t s) { return Stream(s); }
};

}  // namespace cuda
}  // namespace cv
#endif  // CVGS_HAVE_OPENCV
