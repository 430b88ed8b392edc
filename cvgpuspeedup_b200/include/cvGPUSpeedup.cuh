// cvGPUSpeedup.cuh -- source-compatible shim of the reference's cvGS interface for ONE hot path, on top of the
// C-ABI of libcvgs_b200.so (include/cvgs_b200.h).
//
// A pipeline written against the reference (include/cvGPUSpeedup.cuh of cvGPUSpeedup 0.21.0),
//
//     cvGS::executeOperations(cv_stream,
//         cvGS::resize<CV_8UC3, cv::INTER_LINEAR, BATCH>(crops, dsize, usedPlanes[, background]),
//         cvGS::cvtColor<cv::COLOR_RGB2BGR, CV_32FC3>(), cvGS::multiply<CV_32FC3>(alpha),
//         cvGS::subtract<CV_32FC3>(sub), cvGS::divide<CV_32FC3>(div), cvGS::split<CV_32FC3>(d_tensor, dsize));
//
// compiles unchanged against this header and becomes ONE call of cvgs_b200_preproc_launch (= one sm_100a kernel
// launch).  Where the reference builds nested operation-struct types and lets templates fuse them, the
// operation structs here are small runtime descriptors; chains outside the hot path fail to compile with a
// static_assert naming the unsupported operation.  Plain host C++17: usable from g++ as well as nvcc.
//
// Reference lines mirrored: resize :209-245, convertTo :74-129, multiply/subtract/divide/add :131-149,
// cvtColor :151-161, split :185-202, write :449-457, executeOperations :464-473, CircularTensor :600-627,
// AspectRatio :32; error convention fkl/include/fused_kernel/core/utils/utils.h:42-60 (std::runtime_error).
#pragma once
#include <algorithm>
#include <array>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/cvgs_b200.h"
#include "cvgs_opencv_compat.hpp"

namespace fk {  // the few fk names that appear in user code of this path
enum AspectRatio { PRESERVE_AR = 0, IGNORE_AR = 1, PRESERVE_AR_RN_EVEN = 2, PRESERVE_AR_LEFT = 3 };
enum class CircularTensorOrder { NewestFirst, OldestFirst };
enum class ColorPlanes { Standard, Transposed };
enum ND { _1D = 1, _2D = 2, _3D = 3, T3D = 4 };
struct PtrDims3D {
    unsigned width = 0, height = 0, planes = 0, color_planes = 0;
};
template <ND D, typename T>
struct RawPtr {
    T* data = nullptr;
    PtrDims3D dims;
};
enum WarpType { Affine = 0, Perspective = 1 };  // warping.cuh:25
enum PixelFormat { NV12, NV21, YV12, P010, P016, P216, P210, Y216, Y210, Y416 };  // color_conversion.cuh:89
}  // namespace fk

namespace cvGS {

enum AspectRatio { PRESERVE_AR = 0, IGNORE_AR = 1, PRESERVE_AR_RN_EVEN = 2, PRESERVE_AR_LEFT = 3 };

namespace detail {
inline void check(int rc, const char* what) {
    if (rc != 0)  // same convention as gpuErrchk in the reference: throw with the message
        throw std::runtime_error(std::string(what) + ": " + cvgs_b200_last_error() + " (code " + std::to_string(rc) + ")");
}
struct ReadBatch {  // cvGS::resize(...) result
    cv::cuda::GpuMat frame;            // crop(read, rects) form: the one image ...
    std::vector<cvgs_rect_t> rects;    // ... and the rectangles cut from it (then crops / parents stay empty)
    int yuv_standard = 0;  // CVGS_NV12 sources
    std::vector<cvgs_crop_t> crops;
    std::vector<cvgs_parent_t> parents;  // the image each ROI was cut from (GpuMat::datastart / dataend)
    int n_planes = 0, used = 0, dst_w = 0, dst_h = 0, aspect = CVGS_IGNORE_AR, src_type = 0;
    float bg[4] = {0, 0, 0, 0};
};
struct ChainOp {  // multiply / subtract / divide / add / cvtColor / convertTo pieces
    cvgs_op_t op[2];
    int n = 0;
    int to_u8 = 0;       // CVGS_8UC3 / CVGS_8UC4: convertTo<CV_32FCn, CV_8UCn>(), the last operation, followed by write<CV_8UCn>
    bool cast = false;   // fk::Cast<float3, uchar3>::build(): to_u8 with static_cast (truncation) instead of SaturateCast
};
struct WarpBatch {  // cvGS::warp(...) result
    std::vector<cvgs_crop_t> images;
    std::vector<cvgs_warp_t> warps;
    int n_planes = 0, used = 0, dst_w = 0, dst_h = 0, src_type = CVGS_8UC3;
    float bg[4] = {0, 0, 0, 0};
};
struct WriteOp {  // split / splitT / write
    void* out = nullptr;
    int layout = CVGS_OUT_NCHW;
    long long plane_stride = 0;
    std::vector<cvgs_plane_t> planes;  // split(vector<GpuMat>): one destination image per (crop, channel)
    int dst_type = 0;                  // CV_32FC1/3/4 as named by the write op; CV_8UC3: write<CV_8UC3>(...) after convertTo<CV_32FC3, CV_8UC3>()
    long long row_pitch = 0;           // 8-bit destination: GpuMat::step
};
template <typename T>
struct is_op : std::false_type {};
template <>
struct is_op<ChainOp> : std::true_type {};

inline cvgs_op_t scalar_op(int kind, const cv::Scalar& s) {
    cvgs_op_t o{};
    o.kind = kind;
    for (int c = 0; c < 4; ++c) o.v[c] = static_cast<float>(s[c]);  // cvGPUSpeedupHelpers.cuh:38-54: static_cast per channel
    return o;
}
inline void append(cvgs_pipeline_t& p, const ChainOp& c) {
    if (p.dst_type == CVGS_8UC3 || p.dst_type == CVGS_8UC4)
        throw std::runtime_error("cvGS: convertTo<CV_32FCn, CV_8UCn>() must be the last operation before the write");
    if (c.to_u8) p.dst_type = c.to_u8;
    if (c.cast) p.u8_cast = 1;
    for (int i = 0; i < c.n; ++i) {
        if (p.n_ops >= CVGS_MAX_OPS) throw std::runtime_error("cvGS: more than CVGS_MAX_OPS operations in the chain");
        p.ops[p.n_ops++] = c.op[i];
    }
}
inline void append(cvgs_pipeline_t& p, const WriteOp& w) {
    p.out = w.layout == CVGS_OUT_PLANES ? const_cast<cvgs_plane_t*>(w.planes.data()) : w.out;
    p.out_layout = w.layout;
    p.out_plane_stride = w.plane_stride;
    const bool chain_u8 = p.dst_type == CVGS_8UC3 || p.dst_type == CVGS_8UC4;
    const bool write_u8 = w.dst_type == CVGS_8UC3 || w.dst_type == CVGS_8UC4;
    if (chain_u8 != write_u8 || (chain_u8 && p.dst_type != w.dst_type))
        throw std::runtime_error("cvGS: write<CV_8UCn> and convertTo<CV_32FCn, CV_8UCn>() go together");
    p.dst_type = w.dst_type;
    p.out_row_pitch = w.row_pitch;
}
inline cvgs_plane_t plane_of(const cv::cuda::GpuMat& m) {  // gpuMat2RawPtr2D<float>, reference :40-44
    cvgs_plane_t q{};
    q.data = m.data;
    q.pitch_bytes = static_cast<int64_t>(m.step);
    return q;
}
template <typename T>
inline void append(cvgs_pipeline_t&, const T&) {
    static_assert(std::is_same<T, ChainOp>::value || std::is_same<T, WriteOp>::value,
                  "cvGS (B200 build): this operation is not on the fused resize/normalise/split path "
                  "(supported: convertTo, multiply, subtract, divide, add, cvtColor<RGB2BGR/BGR2RGB>, split, splitT, write)");
}
inline cvgs_crop_t crop_of(const cv::cuda::GpuMat& m) {  // gpuMat2RawPtr2D, reference :40-44
    cvgs_crop_t c{};
    c.data = m.data;
    c.width = m.cols;
    c.height = m.rows;
    c.pitch = static_cast<int32_t>(m.step);
    return c;
}
// The image a GpuMat header was cut from, from the members cv::cuda::GpuMat::locateROI uses (datastart, dataend,
// step): lets the library stage all crops of a frame through one cached tensor map (include/cvgs_b200.h,
// cvgs_b200_preproc_launch_ex).  An all-zero parent means "unknown" and is always safe.
inline cvgs_parent_t parent_of(const cv::cuda::GpuMat& m) {
    cvgs_parent_t p{};
    const size_t esz = m.elemSize();
    if (!m.datastart || !m.dataend || m.step == 0 || esz == 0 || m.data < m.datastart) return p;
    const ptrdiff_t d1 = m.data - m.datastart, d2 = m.dataend - m.datastart;
    const ptrdiff_t step = static_cast<ptrdiff_t>(m.step), minstep = static_cast<ptrdiff_t>(m.cols * esz);
    const ptrdiff_t oy = d1 / step, ox = (d1 - oy * step) / static_cast<ptrdiff_t>(esz);
    const ptrdiff_t h = std::max<ptrdiff_t>((d2 - minstep) / step + 1, oy + m.rows);
    const ptrdiff_t w = std::max<ptrdiff_t>((d2 - step * (h - 1)) / static_cast<ptrdiff_t>(esz), ox + m.cols);
    if (h <= 0 || w <= 0 || h > 0x7fffffff || w > 0x7fffffff) return p;
    p.datastart = m.datastart;
    p.whole_width = static_cast<int32_t>(std::min<ptrdiff_t>(w, step / static_cast<ptrdiff_t>(esz)));
    p.whole_height = static_cast<int32_t>(h);
    return p;
}
}  // namespace detail

// Floating-point contract and resize-output mode of subsequent executeOperations calls on this thread
// (DESIGN.md section 2).  Defaults reproduce the reference's fused kernel bit for bit.
inline int& fpContract() {
    thread_local int v = CVGS_FP_REFERENCE_FUSED;
    return v;
}
inline int& interpMode() {
    thread_local int v = CVGS_INTERP_FLOAT;
    return v;
}

// ---- resize (reference :209-245) -------------------------------------------------------------------------
template <int T, int INTER_F, int NPtr, AspectRatio AR_ = IGNORE_AR>
inline detail::ReadBatch resize(const std::array<cv::cuda::GpuMat, NPtr>& input, const cv::Size& dsize, const int& usedPlanes,
                                const cv::Scalar& backgroundValue = cv::Scalar()) {
    static_assert(T == CV_8UC3 || T == CV_16UC3 || T == CV_16SC3 || T == CV_8UC4 || T == CV_16UC4 || T == CV_16SC4,
                  "cvGS (B200 build): CV_8U / CV_16U / CV_16S sources with 3 or 4 channels are on the hot path in this build");
    static_assert(INTER_F == cv::INTER_LINEAR, "cvGS (B200 build): only INTER_LINEAR is implemented (as in the reference)");
    detail::ReadBatch r;
    r.n_planes = NPtr;
    r.used = usedPlanes;
    r.dst_w = dsize.width;
    r.dst_h = dsize.height;
    r.aspect = static_cast<int>(AR_);
    r.src_type = T;
    for (int c = 0; c < 4; ++c) r.bg[c] = static_cast<float>(backgroundValue[c]);
    r.crops.resize(NPtr);
    r.parents.resize(NPtr);
    for (int i = 0; i < NPtr && i < usedPlanes; ++i) {
        r.crops[i] = detail::crop_of(input[i]);
        r.parents[i] = detail::parent_of(input[i]);
    }
    return r;
}
template <int T, int INTER_F>
inline detail::ReadBatch resize(const cv::cuda::GpuMat& input, const cv::Size& dsize, double fx = 0., double fy = 0.) {
    cv::Size d = dsize;
    if (d.width == 0 && d.height == 0) {  // reference :209-216 -> Resize::build with fx, fy (resize.cuh:84-98)
        d.width = static_cast<int>(fx * input.cols + 0.5);  // cxp::round of a positive value
        d.height = static_cast<int>(fy * input.rows + 0.5);
    }
    return resize<T, INTER_F, 1>(std::array<cv::cuda::GpuMat, 1>{input}, d, 1);
}

// ---- crop (reference :247-265, :444 -> fk::Crop<BackIOp>, fkl/.../image_processing/crop.cuh:23-55) ---------------
// The reference's crop() overloads build fk::Crop operations that are composed with fk-level reads and resizes
// (fkl/tests/algorithm/test_crop.cu:24-45): readIOp.then(Crop<>::build(rects)).then(Resize<INTER_LINEAR>::build(size)).
// The same composition here: cvGS::read<T>(frame) is the read of a whole image (what fk::PerThreadRead<_2D, CUDA_T(T)>::
// build(gpuMat2RawPtr2D(frame)) is in the reference), crop(...) the rectangles, cvGS::resize<INTER>(size[, used, bg]) the
// incomplete resize; the result is the batch read executeOperations takes, backed by ONE
// cvgs_b200_preproc_launch_rects call (device frame + rectangle list).
namespace detail {
struct CropOp {  // crop(rect) / crop<BATCH>(rects): incomplete until it follows a read
    std::vector<cvgs_rect_t> rects;
};
struct CroppedRead;
struct ImageRead {  // read of a whole image
    cv::cuda::GpuMat image;
    int type = 0;
    inline CroppedRead then(const CropOp& c) const;  // readIOp.then(crop(rects))
};
struct ResizeOp {  // resize<INTER, AR>(size, used, bg): incomplete until it follows a read
    int dst_w = 0, dst_h = 0, aspect = CVGS_IGNORE_AR, used = -1;
    float bg[4] = {0, 0, 0, 0};
};
struct CroppedRead {  // read.then(crop(...)) / crop(read, ...)
    ImageRead src;
    std::vector<cvgs_rect_t> rects;
    // .then(resize(...)): the batch read of the hot path, one plane per rectangle
    ReadBatch then(const ResizeOp& rs) const {
        ReadBatch r;
        r.n_planes = static_cast<int>(rects.size());
        r.used = rs.used < 0 ? r.n_planes : std::min(rs.used, r.n_planes);
        r.dst_w = rs.dst_w;
        r.dst_h = rs.dst_h;
        r.aspect = rs.aspect;
        r.src_type = src.type;
        for (int c = 0; c < 4; ++c) r.bg[c] = rs.bg[c];
        r.frame = src.image;
        r.rects = rects;
        return r;
    }
    // without a resize: BatchRead of equally sized rectangles, pixels bit for bit (scale exactly 1)
    operator ReadBatch() const {
        if (rects.empty()) throw std::runtime_error("cvGS::crop: no rectangles");
        for (const cvgs_rect_t& q : rects)
            if (q.width != rects[0].width || q.height != rects[0].height)
                throw std::runtime_error("cvGS::crop without a resize: the rectangles of a batch must have one size");
        ResizeOp rs;
        rs.dst_w = rects[0].width;
        rs.dst_h = rects[0].height;
        return then(rs);
    }
};
inline CroppedRead ImageRead::then(const CropOp& c) const { return CroppedRead{*this, c.rects}; }
inline cvgs_rect_t rect_of(const cv::Rect2d& r) {  // the casts of the reference (:248)
    return cvgs_rect_t{static_cast<int32_t>(static_cast<unsigned>(r.x)), static_cast<int32_t>(static_cast<unsigned>(r.y)),
                       static_cast<int32_t>(r.width), static_cast<int32_t>(r.height)};
}
}  // namespace detail
template <int T>
inline detail::ImageRead read(const cv::cuda::GpuMat& image) {
    static_assert(T == CV_8UC3 || T == CV_16UC3 || T == CV_16SC3 || T == CV_8UC4 || T == CV_16UC4 || T == CV_16SC4,
                  "cvGS (B200 build): CV_8U / CV_16U / CV_16S images with 3 or 4 channels");
    return detail::ImageRead{image, T};
}
inline detail::CropOp crop(const cv::Rect2d& rect) { return detail::CropOp{{detail::rect_of(rect)}}; }
template <int BATCH>
inline detail::CropOp crop(const std::array<cv::Rect2d, BATCH>& rects) {
    detail::CropOp c;
    for (const cv::Rect2d& r : rects) c.rects.push_back(detail::rect_of(r));
    return c;
}
inline detail::CroppedRead crop(const detail::ImageRead& backIOp, const cv::Rect2d& rect) {
    return detail::CroppedRead{backIOp, {detail::rect_of(rect)}};
}
template <size_t BATCH>
inline detail::CroppedRead crop(const detail::ImageRead& backIOp, const std::array<cv::Rect2d, BATCH>& rects) {
    detail::CroppedRead c{backIOp, {}};
    for (const cv::Rect2d& r : rects) c.rects.push_back(detail::rect_of(r));
    return c;
}
// crop(backIOp, rects) of the reference is backIOp.then(crop(rects)) (:444)
// the incomplete resize that follows a read (fk::Resize<INTER_LINEAR, AR>::build(size), resize.cuh:84-98)
template <int INTER_F, AspectRatio AR_ = IGNORE_AR>
inline detail::ResizeOp resize(const cv::Size& dsize, int usedPlanes = -1, const cv::Scalar& backgroundValue = cv::Scalar()) {
    static_assert(INTER_F == cv::INTER_LINEAR, "cvGS (B200 build): only INTER_LINEAR is implemented (as in the reference)");
    detail::ResizeOp r;
    r.dst_w = dsize.width;
    r.dst_h = dsize.height;
    r.aspect = static_cast<int>(AR_);
    r.used = usedPlanes;
    for (int c = 0; c < 4; ++c) r.bg[c] = static_cast<float>(backgroundValue[c]);
    return r;
}

// NV12 frames (decoder output) as sources.  The reference reaches this through its fk:: layer only
// (fk::Resize<INTER_LINEAR>::build(fk::fuse(fk::Read<fk::ReadYUV<fk::NV12>>{frame}, fk::Unary<fk::ConvertYUVToRGB<
// fk::NV12, range, primaries, false, float3>>{}), dsize), tests/resize/test_fused_resize.cu:73-76); here it is one
// more read of the same pipeline.  Each GpuMat is the CV_8UC1 luma plane; the interleaved UV plane follows it at
// data + step * rows.  standard: enum cvgs_yuv_standard.
// resizeYUV<fk::NV21 / fk::P010 / fk::P210 / fk::Y210, N>: the other fk::ReadYUV formats (color_conversion.cuh:296-345);
// the GpuMat describes the luma plane in pixels (cols x rows, step in bytes) -- CV_8UC1, CV_16UC1, or for Y210 the
// packed image with cols = pixel width.
template <int NPtr>
inline detail::ReadBatch resizeNV12(const std::array<cv::cuda::GpuMat, NPtr>& frames, const cv::Size& dsize, int standard,
                                    int usedPlanes = NPtr, const cv::Scalar& backgroundValue = cv::Scalar());
template <fk::PixelFormat PF, int NPtr>
inline detail::ReadBatch resizeYUV(const std::array<cv::cuda::GpuMat, NPtr>& frames, const cv::Size& dsize, int standard,
                                   int usedPlanes = NPtr, const cv::Scalar& backgroundValue = cv::Scalar()) {
    static_assert(PF == fk::NV12 || PF == fk::NV21 || PF == fk::P010 || PF == fk::P210 || PF == fk::Y210,
                  "cvGS (B200 build): NV12, NV21, P010, P210 and Y210 frames are the YUV sources on this path");
    detail::ReadBatch r = resizeNV12<NPtr>(frames, dsize, standard, usedPlanes, backgroundValue);
    r.src_type = PF == fk::NV12 ? CVGS_NV12 : PF == fk::NV21 ? CVGS_NV21 : PF == fk::P010 ? CVGS_P010 : PF == fk::P210 ? CVGS_P210 : CVGS_Y210;
    for (int i = 0; i < NPtr && i < usedPlanes; ++i) {  // crop_of takes the width in pixels of the GpuMat's own type
        r.crops[i].width = frames[i].cols;
        r.crops[i].height = frames[i].rows;
    }
    return r;
}
template <int NPtr>
inline detail::ReadBatch resizeNV12(const std::array<cv::cuda::GpuMat, NPtr>& frames, const cv::Size& dsize, int standard,
                                    int usedPlanes, const cv::Scalar& backgroundValue) {
    detail::ReadBatch r;
    r.n_planes = NPtr;
    r.used = usedPlanes;
    r.dst_w = dsize.width;
    r.dst_h = dsize.height;
    r.aspect = CVGS_IGNORE_AR;
    r.src_type = CVGS_NV12;
    r.yuv_standard = standard;
    for (int c = 0; c < 4; ++c) r.bg[c] = static_cast<float>(backgroundValue[c]);
    r.crops.resize(NPtr);
    for (int i = 0; i < NPtr && i < usedPlanes; ++i) r.crops[i] = detail::crop_of(frames[i]);
    return r;
}

// ---- element-wise operations (reference :74-161) ---------------------------------------------------------
template <int I, int O>
inline detail::ChainOp convertTo() {  // SaturateCast<u8 -> f32>: the resize already yields float
    static_assert(CV_MAT_DEPTH(O) == CV_32F || (I == CV_32FC3 && O == CV_8UC3) || (I == CV_32FC4 && O == CV_8UC4),
                  "cvGS (B200 build): convertTo produces CV_32F, or CV_8UCn from CV_32FCn as the last operation before write<CV_8UCn>");
    detail::ChainOp c;
    c.to_u8 = (O == CV_8UC3 || O == CV_8UC4) ? O : 0;  // SaturateCast<float, uchar>: applied by the kernel when it writes the 8-bit pixels
    return c;
}
template <int I, int O>
inline detail::ChainOp convertTo(float alpha) {
    static_assert(CV_MAT_DEPTH(O) == CV_32F, "cvGS (B200 build): the fused path produces CV_32F");
    detail::ChainOp c;
    c.op[c.n++] = detail::scalar_op(CVGS_OP_MUL, cv::Scalar::all(alpha));
    return c;
}
template <int I, int O>
inline detail::ChainOp convertTo(float alpha, float beta) {
    detail::ChainOp c = convertTo<I, O>(alpha);
    c.op[c.n++] = detail::scalar_op(CVGS_OP_ADD, cv::Scalar::all(beta));
    return c;
}
#define CVGS_SCALAR_OP(NAME, KIND)                              \
    template <int I>                                            \
    inline detail::ChainOp NAME(const cv::Scalar& src2) {       \
        detail::ChainOp c;                                      \
        c.op[c.n++] = detail::scalar_op(KIND, src2);            \
        return c;                                               \
    }
CVGS_SCALAR_OP(multiply, CVGS_OP_MUL)
CVGS_SCALAR_OP(subtract, CVGS_OP_SUB)
CVGS_SCALAR_OP(divide, CVGS_OP_DIV)
CVGS_SCALAR_OP(add, CVGS_OP_ADD)
#undef CVGS_SCALAR_OP

// fk::ColorConversion<CODE, I, O> (color_conversion.cuh:364-461; supported set cv2cuda_types.cuh:77-86): the R<->B swaps
// are a channel reorder; the other codes add an opaque alpha (AddOpaqueAlpha<I, p8bit>: 255), drop the alpha or reduce
// to gray (RGB2Gray), each optionally behind the swap.
template <cv::ColorConversionCodes CODE, int I, int O = I>
inline detail::ChainOp cvtColor() {
    static_assert(CV_MAT_DEPTH(I) == CV_32F && CV_MAT_DEPTH(O) == CV_32F,
                  "cvGS (B200 build): colour conversions run on the float pixels behind the resize");
    constexpr int ci = CV_MAT_CN(I), co = CV_MAT_CN(O);
    constexpr bool swap = CODE == cv::COLOR_BGR2RGB || CODE == cv::COLOR_BGRA2RGBA || CODE == cv::COLOR_BGR2RGBA ||
                          CODE == cv::COLOR_RGBA2BGR || CODE == cv::COLOR_BGR2GRAY || CODE == cv::COLOR_BGRA2GRAY;
    constexpr bool add_alpha = CODE == cv::COLOR_BGR2BGRA || CODE == cv::COLOR_BGR2RGBA;
    constexpr bool drop_alpha = CODE == cv::COLOR_BGRA2BGR || CODE == cv::COLOR_RGBA2BGR;
    constexpr bool gray = CODE == cv::COLOR_BGR2GRAY || CODE == cv::COLOR_RGB2GRAY || CODE == cv::COLOR_BGRA2GRAY ||
                          CODE == cv::COLOR_RGBA2GRAY;
    static_assert(swap || add_alpha || drop_alpha || gray, "Color conversion type not supported yet.");
    constexpr int want_ci = (CODE == cv::COLOR_BGRA2RGBA || drop_alpha || CODE == cv::COLOR_BGRA2GRAY || CODE == cv::COLOR_RGBA2GRAY) ? 4 : 3;
    constexpr int want_co = gray ? 1 : (add_alpha ? 4 : (drop_alpha ? 3 : want_ci));
    static_assert(ci == want_ci && co == want_co, "cvGS: colour code and channel counts disagree");
    detail::ChainOp c;
    if (swap) {
        cvgs_op_t& r = c.op[c.n++];
        r.kind = CVGS_OP_REORDER;
        r.perm[0] = 2; r.perm[1] = 1; r.perm[2] = 0; r.perm[3] = 3;
    }
    if (add_alpha) {
        cvgs_op_t& a = c.op[c.n++];
        a.kind = CVGS_OP_ADD_ALPHA;
        a.v[0] = 255.f;  // maxDepthValue<p8bit>, the ColorDepth cvGS::cvtColor leaves at its default
    } else if (drop_alpha) {
        c.op[c.n++].kind = CVGS_OP_DROP_ALPHA;
    } else if (gray) {
        cvgs_op_t& g = c.op[c.n++];
        g.kind = CVGS_OP_GRAY;
        g.perm[0] = swap ? 0 : 1;  // the stand-alone FMUL of that instantiation (include/cvgs_b200.h, CVGS_OP_GRAY)
    }
    return c;
}

namespace detail {
template <int O>
inline WriteOp typed(WriteOp w) {  // the write op names the pixel type the chain must end with
    w.dst_type = O;                // CV_32FC1 / CV_32FC3 / CV_32FC4 / CV_8UC3 have the values of CVGS_32FC1 ... CVGS_8UC3
    return w;
}
}  // namespace detail
// ---- writes (reference :185-202, :449-457) ---------------------------------------------------------------
template <int O>
inline detail::WriteOp split(const cv::cuda::GpuMat& output, const cv::Size& /*planeDims*/) {
    static_assert(O == CV_32FC1 || O == CV_32FC3 || O == CV_32FC4, "cvGS (B200 build): CV_32FC1 / CV_32FC3 / CV_32FC4 output");
    return detail::typed<O>({output.data, CVGS_OUT_NCHW, 0, {}});  // the reference builds a tight Tensor and ignores GpuMat::step (:67-71)
}
// fk::SplitWrite: the channels of a crop go to separate CV_32FC1 images (reference :163-183)
template <int O>
inline detail::WriteOp split(const std::vector<cv::cuda::GpuMat>& output) {
    static_assert(O == CV_32FC1 || O == CV_32FC3 || O == CV_32FC4, "cvGS (B200 build): CV_32FC1 / CV_32FC3 / CV_32FC4 output");
    if (output.size() != static_cast<size_t>(CV_MAT_CN(O))) throw std::runtime_error("cvGS::split: one destination image per channel is required");
    detail::WriteOp w;
    w.layout = CVGS_OUT_PLANES;
    for (const auto& m : output) w.planes.push_back(detail::plane_of(m));
    return w;
}
template <int O, int N>
inline detail::WriteOp split(const std::array<std::vector<cv::cuda::GpuMat>, N>& output) {
    static_assert(O == CV_32FC1 || O == CV_32FC3 || O == CV_32FC4, "cvGS (B200 build): CV_32FC1 / CV_32FC3 / CV_32FC4 output");
    detail::WriteOp w;
    w.layout = CVGS_OUT_PLANES;
    for (const auto& crop : output) {
        if (crop.size() != static_cast<size_t>(CV_MAT_CN(O))) throw std::runtime_error("cvGS::split: one destination image per channel and crop is required");
        for (const auto& m : crop) w.planes.push_back(detail::plane_of(m));
    }
    return w;
}
template <int O>
inline detail::WriteOp split(const fk::RawPtr<fk::_3D, float>& output) {
    return detail::typed<O>({output.data, CVGS_OUT_NCHW, 0, {}});
}
template <int O>
inline detail::WriteOp splitT(const fk::RawPtr<fk::T3D, float>& output) {
    return detail::typed<O>({output.data, CVGS_OUT_CNHW, 0, {}});
}
template <int O>
inline detail::WriteOp write(const cv::cuda::GpuMat& output, const cv::Size& plane) {
    static_assert(O == CV_32FC1 || O == CV_32FC3 || O == CV_32FC4 || O == CV_8UC3 || O == CV_8UC4, "cvGS (B200 build): CV_32FC1/3/4 or CV_8UC3/4 output");
    detail::WriteOp w = detail::typed<O>({output.data, CVGS_OUT_NHWC, 0, {}});
    if (CV_MAT_DEPTH(O) == CV_8U) w.row_pitch = static_cast<long long>(CV_MAT_CN(O)) * plane.width;  // gpuMat2Tensor builds a tight tensor of `plane`-sized images (reference :67-71)
    return w;
}
// PerThreadWrite<_2D, O>: one image with the GpuMat's own pitch (reference :449-452; tests/resize/test_resize_write.cu)
template <int O>
inline detail::WriteOp write(const cv::cuda::GpuMat& output) {
    static_assert(O == CV_32FC1 || O == CV_32FC3 || O == CV_32FC4 || O == CV_8UC3 || O == CV_8UC4, "cvGS (B200 build): CV_32FC1/3/4 or CV_8UC3/4 output");
    detail::WriteOp w = detail::typed<O>({output.data, CVGS_OUT_NHWC, 0, {}});
    w.row_pitch = static_cast<long long>(output.step);  // the image's own pitch (cudaMallocPitch pads rows)
    return w;
}

// ---- executeOperations (reference :464-473) --------------------------------------------------------------
template <typename... IOpTypes>
inline void executeOperations(const cv::cuda::Stream& stream, const detail::ReadBatch& read, const IOpTypes&... iops) {
    cvgs_pipeline_t p{};
    p.src_type = read.src_type;
    p.dst_width = read.dst_w;
    p.dst_height = read.dst_h;
    p.aspect_mode = read.aspect;
    p.interp_mode = interpMode();
    p.fp_contract = fpContract();
    p.yuv_standard = read.yuv_standard;
    for (int c = 0; c < 4; ++c) p.background[c] = read.bg[c];
    (detail::append(p, iops), ...);
    if (!read.rects.empty()) {  // crop(read, rects)[.then(resize)]: one device frame + its rectangles
        detail::check(cvgs_b200_preproc_launch_rects(read.frame.data, read.frame.cols, read.frame.rows, static_cast<int32_t>(read.frame.step),
                                                     read.rects.data(), read.n_planes, read.used, &p,
                                                     cv::cuda::StreamAccessor::getStream(stream)),
                      "cvGS::executeOperations");
        return;
    }
    detail::check(cvgs_b200_preproc_launch_ex(read.crops.data(), read.parents.empty() ? nullptr : read.parents.data(),
                                              read.n_planes, read.used, &p, cv::cuda::StreamAccessor::getStream(stream)),
                  "cvGS::executeOperations");
}
template <bool ENABLE_THREAD_FUSION, typename... IOpTypes>
inline void executeOperations(const cv::cuda::Stream& stream, const detail::ReadBatch& read, const IOpTypes&... iops) {
    executeOperations(stream, read, iops...);  // thread fusion is a property of the hand-written kernel here
}

// ---- warp (reference :266-442 -> fk::Warping, warping.cuh:43-91) ------------------------------------------------------
// The wrapper inverts the user's matrix on the host (cv::invertAffineTransform / Mat::inv(), double) and hands the
// kernel its float cast; so does this one.
namespace detail {
template <fk::WarpType WT>
inline cvgs_warp_t inverse_of(const cv::Mat& m) {
    if (m.type() != CV_64FC1) throw std::runtime_error("Transform matrix type should be CV_64FC1.");
    cvgs_warp_t w{};
    cv::Mat inv;
    if (WT == fk::WarpType::Affine) {
        cv::invertAffineTransform(m, inv);
        w.type = CVGS_WARP_AFFINE;
    } else {
        inv = m.inv();
        w.type = CVGS_WARP_PERSPECTIVE;
    }
    const double* t = inv.ptr<double>();
    for (int k = 0; k < (WT == fk::WarpType::Affine ? 6 : 9); ++k) w.m[k] = static_cast<float>(t[k]);
    return w;
}
}  // namespace detail
template <fk::WarpType WT, int InputType = CV_8UC3>
inline detail::WarpBatch warp(const cv::cuda::GpuMat& input, const cv::Mat& transform_matrix, const cv::Size& dstSize) {
    static_assert((CV_MAT_DEPTH(InputType) == CV_8U || CV_MAT_DEPTH(InputType) == CV_16U || CV_MAT_DEPTH(InputType) == CV_16S) &&
                      (CV_MAT_CN(InputType) == 3 || CV_MAT_CN(InputType) == 4),
                  "cvGS (B200 build): warp takes CV_8U / CV_16U / CV_16S images with 3 or 4 channels");
    if (InputType != input.type()) throw std::runtime_error("Input type does not match the input type of the operation.");
    detail::WarpBatch b;
    b.images.push_back(detail::crop_of(input));
    b.warps.push_back(detail::inverse_of<WT>(transform_matrix));
    b.n_planes = b.used = 1;
    b.src_type = InputType;  // CV type codes and CVGS_* source codes coincide
    b.dst_w = dstSize.width;
    b.dst_h = dstSize.height;
    return b;
}
template <fk::WarpType WT, int InputType, size_t BATCH>
inline detail::WarpBatch warp(const std::array<cv::cuda::GpuMat, BATCH>& inputs, const std::array<cv::Mat, BATCH>& transform_matrices,
                              const std::array<cv::Size, BATCH>& dstSize, const int& usedPlanes, const cv::Scalar& defaultValue) {
    static_assert((CV_MAT_DEPTH(InputType) == CV_8U || CV_MAT_DEPTH(InputType) == CV_16U || CV_MAT_DEPTH(InputType) == CV_16S) &&
                      (CV_MAT_CN(InputType) == 3 || CV_MAT_CN(InputType) == 4),
                  "cvGS (B200 build): warp takes CV_8U / CV_16U / CV_16S images with 3 or 4 channels");
    detail::WarpBatch b;
    b.n_planes = static_cast<int>(BATCH);
    b.src_type = InputType;
    b.used = usedPlanes;
    for (int i = 0; i < usedPlanes && i < static_cast<int>(BATCH); ++i) {
        if (InputType != inputs[i].type()) throw std::runtime_error("Input type does not match the input type of the operation.");
        if (dstSize[i].width != dstSize[0].width || dstSize[i].height != dstSize[0].height)
            throw std::runtime_error("cvGS (B200 build): the warps of a batch share one destination size");
        b.images.push_back(detail::crop_of(inputs[i]));
        b.warps.push_back(detail::inverse_of<WT>(transform_matrices[i]));
    }
    b.dst_w = dstSize[0].width;
    b.dst_h = dstSize[0].height;
    for (int c = 0; c < 4; ++c) b.bg[c] = static_cast<float>(defaultValue[c]);
    return b;
}
template <fk::WarpType WT, int InputType, size_t BATCH>
inline detail::WarpBatch warp(const std::array<cv::cuda::GpuMat, BATCH>& inputs, const std::array<cv::Mat, BATCH>& transform_matrices,
                              const std::array<cv::Size, BATCH>& dstSize) {
    return warp<WT, InputType, BATCH>(inputs, transform_matrices, dstSize, static_cast<int>(BATCH), cv::Scalar());
}
template <fk::WarpType WT, int InputType, size_t BATCH>
inline detail::WarpBatch warp(const std::array<cv::cuda::GpuMat, BATCH>& inputs, const std::array<cv::Mat, BATCH>& transform_matrices,
                              const cv::Size& dstSize) {
    std::array<cv::Size, BATCH> sizes;
    sizes.fill(dstSize);
    return warp<WT, InputType, BATCH>(inputs, transform_matrices, sizes, static_cast<int>(BATCH), cv::Scalar());
}
template <fk::WarpType WT, int InputType, size_t BATCH>
inline detail::WarpBatch warp(const std::array<cv::cuda::GpuMat, BATCH>& inputs, const std::array<cv::Mat, BATCH>& transform_matrices,
                              const cv::Size& dstSize, const int& usedPlanes, const cv::Scalar& defaultValue) {
    std::array<cv::Size, BATCH> sizes;
    sizes.fill(dstSize);
    return warp<WT, InputType, BATCH>(inputs, transform_matrices, sizes, usedPlanes, defaultValue);
}
template <typename... IOpTypes>
inline void executeOperations(const cv::cuda::Stream& stream, const detail::WarpBatch& read, const IOpTypes&... iops) {
    cvgs_pipeline_t p{};
    p.src_type = read.src_type;
    p.dst_width = read.dst_w;
    p.dst_height = read.dst_h;
    p.aspect_mode = CVGS_IGNORE_AR;
    p.fp_contract = fpContract();
    for (int c = 0; c < 4; ++c) p.background[c] = read.bg[c];
    (detail::append(p, iops), ...);
    detail::check(cvgs_b200_warp_launch(read.images.data(), read.warps.data(), read.n_planes, read.used, &p,
                                        cv::cuda::StreamAccessor::getStream(stream)),
                  "cvGS::executeOperations");
}

// ---- batch reads without a resize (reference :504-584; tests/batchread/test_batchread_x_write3D.cu) -----------
// fk::BatchRead<PerThreadRead> of N equally sized images.  The same kernel serves it: with destination size ==
// source size the scale factors are exactly 1, every tap weight is exactly 0 or 1 and the interpolated value is the
// source pixel bit for bit.
namespace detail {
template <size_t Batch>
inline ReadBatch read_batch(const std::array<cv::cuda::GpuMat, Batch>& input, size_t activeBatch, const cv::Scalar& def) {
    static_assert(Batch > 0, "empty batch");
    const int t = input[0].type();
    if (t != CV_8UC3 && t != CV_16UC3 && t != CV_16SC3 && t != CV_8UC4 && t != CV_16UC4 && t != CV_16SC4)
        throw std::runtime_error("cvGS (B200 build): batch reads take 8U / 16U / 16S images with 3 or 4 channels");
    ReadBatch r;
    r.n_planes = static_cast<int>(Batch);
    r.used = static_cast<int>(activeBatch < Batch ? activeBatch : Batch);
    r.dst_w = input[0].cols;
    r.dst_h = input[0].rows;
    r.aspect = CVGS_IGNORE_AR;
    r.src_type = t;
    for (int c = 0; c < 4; ++c) r.bg[c] = static_cast<float>(def[c]);
    r.crops.resize(Batch);
    r.parents.resize(Batch);
    for (int i = 0; i < r.used; ++i) {
        if (input[i].cols != r.dst_w || input[i].rows != r.dst_h || input[i].type() != t)
            throw std::runtime_error("cvGS::executeOperations: the images of a batch read must have one size and type");
        r.crops[i] = crop_of(input[i]);
        r.parents[i] = parent_of(input[i]);
    }
    return r;
}
}  // namespace detail
template <size_t Batch, typename... IOpTypes>
inline void executeOperations(const std::array<cv::cuda::GpuMat, Batch>& input, const size_t& activeBatch,
                              const cv::Scalar& defaultValue, const cv::cuda::Stream& stream, const IOpTypes&... iops) {
    executeOperations(stream, detail::read_batch(input, activeBatch, defaultValue), iops...);
}
template <size_t Batch, typename... IOpTypes>
inline void executeOperations(const std::array<cv::cuda::GpuMat, Batch>& input, const cv::cuda::Stream& stream,
                              const IOpTypes&... iops) {
    executeOperations(stream, detail::read_batch(input, Batch, cv::Scalar()), iops...);
}
// ... with the destination given as (GpuMat, plane size): the implicit final write is PerThreadWrite<_3D> (packed)
template <size_t Batch, typename... IOpTypes>
inline void executeOperations(const std::array<cv::cuda::GpuMat, Batch>& input, const size_t& activeBatch,
                              const cv::Scalar& defaultValue, const cv::cuda::GpuMat& output, const cv::Size& outputPlane,
                              const cv::cuda::Stream& stream, const IOpTypes&... iops) {
    executeOperations(stream, detail::read_batch(input, activeBatch, defaultValue), iops..., write<CV_32FC3>(output, outputPlane));
}
template <size_t Batch, typename... IOpTypes>
inline void executeOperations(const std::array<cv::cuda::GpuMat, Batch>& input, const cv::cuda::GpuMat& output,
                              const cv::Size& outputPlane, const cv::cuda::Stream& stream, const IOpTypes&... iops) {
    executeOperations(stream, detail::read_batch(input, Batch, cv::Scalar()), iops..., write<CV_32FC3>(output, outputPlane));
}
template <bool ENABLE_THREAD_FUSION, size_t Batch, typename... IOpTypes>
inline void executeOperations(const std::array<cv::cuda::GpuMat, Batch>& input, const cv::cuda::Stream& stream,
                              const IOpTypes&... iops) {
    executeOperations(input, stream, iops...);
}
// single image (reference :475-487)
template <typename... IOpTypes>
inline void executeOperations(const cv::cuda::GpuMat& input, const cv::cuda::Stream& stream, const IOpTypes&... iops) {
    executeOperations(std::array<cv::cuda::GpuMat, 1>{input}, stream, iops...);
}
// One image in, one image out (reference :489-503; tests/read/test_read_x_write.cu): PerThreadRead -> ops ->
// PerThreadWrite<_2D> with the output's own pitch.  The reference takes the pixel type of the write from the last
// operation; the operations are type-erased here, so it is the GpuMat's type (the library checks it against the chain).
template <typename... IOpTypes>
inline void executeOperations(const cv::cuda::GpuMat& input, cv::cuda::GpuMat& output, cv::cuda::Stream& stream,
                              const IOpTypes&... iops) {
    const int t = output.type();
    if (t != CV_32FC1 && t != CV_32FC3 && t != CV_32FC4 && t != CV_8UC3 && t != CV_8UC4)
        throw std::runtime_error("cvGS::executeOperations: output must be CV_32FC1/3/4 or CV_8UC3/4");
    if (output.cols != input.cols || output.rows != input.rows)
        throw std::runtime_error("cvGS::executeOperations: input and output sizes differ");
    detail::WriteOp w{output.data, CVGS_OUT_NHWC, 0, {}};
    w.dst_type = t;
    w.row_pitch = static_cast<long long>(output.step);
    executeOperations(stream, detail::read_batch(std::array<cv::cuda::GpuMat, 1>{input}, 1, cv::Scalar()), iops..., w);
}
template <bool ENABLE_THREAD_FUSION, typename... IOpTypes>
inline void executeOperations(const cv::cuda::GpuMat& input, cv::cuda::GpuMat& output, cv::cuda::Stream& stream,
                              const IOpTypes&... iops) {
    executeOperations(input, output, stream, iops...);
}

// ---- CircularTensor (reference :600-627 over fkl/.../core/data/circular_tensor.cuh:84-151) ----------------
template <int I, int O, int COLOR_PLANES, int BATCH, fk::CircularTensorOrder CT_ORDER,
          fk::ColorPlanes CP_MODE = fk::ColorPlanes::Standard>
class CircularTensor {
    // O names the tensor's element: a depth (CV_32F: planes of floats, COLOR_PLANES of them) or, with COLOR_PLANES == 1, a
    // packed pixel type (CV_32FC3 / CV_32FC4: TensorWrite, reference test_circularbatchread_x_write3D.cu:400-460)
    static constexpr int kElemChannels = COLOR_PLANES == 1 ? CV_MAT_CN(O) : 1;
    static_assert((I == CV_8UC3 || I == CV_16UC3 || I == CV_16SC3 || I == CV_8UC4 || I == CV_16UC4 || I == CV_16SC4) &&
                      CV_MAT_DEPTH(O) == CV_32F && (COLOR_PLANES == 1 || COLOR_PLANES == 3 || COLOR_PLANES == 4) &&
                      (kElemChannels == 1 || kElemChannels == 3 || kElemChannels == 4),
                  "cvGS (B200 build): CircularTensor of CV_8U / CV_16U / CV_16S frames with 3 or 4 channels into 1, 3 or 4 float planes, or "
                  "into one plane of CV_32FC3 / CV_32FC4 pixels");

public:
    CircularTensor() = default;
    CircularTensor(const uint& width_, const uint& height_, const int& deviceID_ = 0) { Alloc(width_, height_, deviceID_); }
    CircularTensor(const CircularTensor&) = delete;
    CircularTensor& operator=(const CircularTensor&) = delete;
    ~CircularTensor() {
        if (h_) cvgs_b200_ct_destroy(h_);
    }
    void Alloc(const uint& width_, const uint& height_, const int& deviceID_ = 0) {
        if (h_) detail::check(cvgs_b200_ct_destroy(h_), "CircularTensor::Alloc");
        h_ = nullptr;
        w_ = width_;
        hgt_ = height_;
        detail::check(cvgs_b200_ct_create_ex(&h_, static_cast<int>(width_), static_cast<int>(height_), COLOR_PLANES, kElemChannels, BATCH,
                                             CT_ORDER == fk::CircularTensorOrder::NewestFirst ? CVGS_CT_NEWEST_FIRST
                                                                                              : CVGS_CT_OLDEST_FIRST,
                                             CP_MODE == fk::ColorPlanes::Standard ? CVGS_CT_STANDARD : CVGS_CT_TRANSPOSED,
                                             deviceID_),
                      "CircularTensor::Alloc");
    }
    // update(stream, frame, ops..., write): the trailing write names this tensor in the reference
    // (fk::Write<TensorSplit/TensorTSplit>{myTensor.ptr()}); here the destination is implied, a WriteOp is accepted
    // and ignored.  The frame is resized to the tensor's plane size when its size differs.
    template <typename... IOpTypes>
    void update(const cv::cuda::Stream& stream, const cv::cuda::GpuMat& input, const IOpTypes&... iops) {
        cvgs_pipeline_t p{};
        p.src_type = I;
        p.dst_width = static_cast<int>(w_);
        p.dst_height = static_cast<int>(hgt_);
        p.aspect_mode = CVGS_IGNORE_AR;
        p.interp_mode = interpMode();
        p.fp_contract = fpContract();
        (detail::append(p, iops), ...);
        const cvgs_crop_t frame = detail::crop_of(input);
        detail::check(cvgs_b200_ct_update(h_, &frame, &p, cv::cuda::StreamAccessor::getStream(stream)),
                      "CircularTensor::update");
    }
    float* data() { return static_cast<float*>(cvgs_b200_ct_data(h_)); }
    using PtrT = fk::RawPtr<CP_MODE == fk::ColorPlanes::Standard ? fk::_3D : fk::T3D, float>;
    PtrT ptr() {
        PtrT r;
        r.data = data();
        r.dims = {w_, hgt_, BATCH, COLOR_PLANES};
        return r;
    }
    size_t sizeInBytes() const { return sizeof(float) * w_ * hgt_ * BATCH * COLOR_PLANES * kElemChannels; }

private:
    void* h_ = nullptr;
    uint w_ = 0, hgt_ = 0;
};

}  // namespace cvGS

namespace fk {
// fk::Cast<float3, uchar3>::build() in front of cvGS::write<CV_8UC3> (reference basic_ops/cast.cuh:22-29; chain of
// tests/warping/test_warping_opencv.cu:63): static_cast per channel, i.e. truncation.
template <typename I, typename O>
struct Cast {
    static_assert((std::is_same<I, float3>::value && std::is_same<O, uchar3>::value) ||
                      (std::is_same<I, float4>::value && std::is_same<O, uchar4>::value),
                  "cvGS (B200 build): fk::Cast<float3, uchar3> / <float4, uchar4> are the casts on this path");
    static cvGS::detail::ChainOp build() {
        cvGS::detail::ChainOp c{};
        c.to_u8 = std::is_same<O, uchar3>::value ? CVGS_8UC3 : CVGS_8UC4;
        c.cast = true;
        return c;
    }
};
}  // namespace fk
