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
#include <memory>
#include <stdexcept>
#include <string>

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
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)

namespace cv {

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2 };
enum ColorConversionCodes { COLOR_BGR2BGRA = 0, COLOR_BGRA2BGR = 1, COLOR_BGR2RGBA = 2, COLOR_RGBA2BGR = 3,
                            COLOR_BGR2RGB = 4, COLOR_RGB2BGR = COLOR_BGR2RGB, COLOR_BGRA2RGBA = 5,
                            COLOR_RGBA2BGRA = COLOR_BGRA2RGBA };

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
struct Scalar {
    double val[4] = {0, 0, 0, 0};
    Scalar() = default;
    Scalar(double v0, double v1 = 0, double v2 = 0, double v3 = 0) : val{v0, v1, v2, v3} {}
    double operator[](int i) const { return val[i]; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
};

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
