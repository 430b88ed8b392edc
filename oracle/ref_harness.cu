// oracle/ref_harness.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" driver that CALLS the unmodified reference library
// (FusedKernelLibrary headers under /root/reference/fkl/include, found through
// -I at build time; nothing is copied into this repo) so that
//   * tests can compare the B200 kernels bit-for-bit with the reference's own
//     fused kernel on the same inputs (oracle/_ref/libfkref.so), and
//   * bench.py can time "the kernel to beat" on the same box.
//
// The call sequence mirrors benchmarks/benchmark_CPUandGPU_cvGS_vs_fk.cu:124-139
// (PerThreadRead::build_batch -> Resize::build_batch -> BatchRead::build ->
// fk::executeOperations(..., ColorConversion, Mul, Sub, Div, TensorSplit)) and
// tests/batchread/test_circularbatchread_x_write3D.cu:185-203 (CircularTensor).
//
// The reference fixes batch size, aspect-ratio mode and the op chain at compile
// time, so only the few instantiations listed at the bottom exist.
#include <array>
#include <cstdint>
#include <cstdio>
#include <string>
#include <thread>

#include <fused_kernel/fused_kernel.cuh>
#include <fused_kernel/core/data/circular_tensor.cuh>
#include <fused_kernel/algorithms/image_processing/resize.cuh>
#include <fused_kernel/algorithms/image_processing/color_conversion.cuh>
#include <fused_kernel/algorithms/image_processing/saturate.cuh>
#include <fused_kernel/algorithms/basic_ops/arithmetic.cuh>

#ifndef FKREF_BATCH
#define FKREF_BATCH 16
#endif

namespace {

thread_local std::string g_err;

struct CropIn { const void* data; int w, h, pitch; };

template <typename PixelT, int BATCH, fk::AspectRatio AR, bool SWAP>
int run_chain(const void* const* ptrs, const int* ws, const int* hs, const int* pitches,
              int used, int dst_w, int dst_h, const float* bg,
              const float* mul, const float* sub, const float* div,
              float* out, cudaStream_t stream) {
    using PixelReadOp = fk::PerThreadRead<fk::_2D, PixelT>;
    std::array<fk::RawPtr<fk::_2D, PixelT>, BATCH> crops{};
    for (int i = 0; i < BATCH; ++i) {
        const int j = i < used ? i : 0;  // unused slots still need a valid descriptor
        crops[i] = fk::RawPtr<fk::_2D, PixelT>{ (PixelT*)ptrs[j],
            { (uint)ws[j], (uint)hs[j], (uint)pitches[j] } };
    }
    constexpr int CN = fk::cn<PixelT>;
    using FloatT = fk::VectorType_t<float, CN>;
    const fk::Size dsize{ dst_w, dst_h };
    FloatT bgv, mulv, subv, divv;
    bgv.x = bg[0]; bgv.y = bg[1]; bgv.z = bg[2];
    mulv.x = mul[0]; mulv.y = mul[1]; mulv.z = mul[2];
    subv.x = sub[0]; subv.y = sub[1]; subv.z = sub[2];
    divv.x = div[0]; divv.y = div[1]; divv.z = div[2];
    if constexpr (CN == 4) { bgv.w = bg[3]; mulv.w = mul[3]; subv.w = sub[3]; divv.w = div[3]; }
    const auto readOP = PixelReadOp::build_batch(crops);
    const auto sizeArr = fk::make_set_std_array<BATCH>(dsize);
    using Resize = fk::Resize<fk::INTER_LINEAR, AR, fk::Read<PixelReadOp>>;
    const fk::Tensor<float> t_out(out, dst_w, dst_h, BATCH, CN);
    const auto mulOp = fk::Binary<fk::Mul<FloatT>>{ mulv };
    const auto subOp = fk::Binary<fk::Sub<FloatT>>{ subv };
    const auto divOp = fk::Binary<fk::Div<FloatT>>{ divv };
    const auto wrOp = fk::Write<fk::TensorSplit<FloatT>>{ t_out.ptr() };
    auto launch = [&](const auto& resizeOp) {
        if constexpr (SWAP) {
            if constexpr (CN == 4) {
                fk::executeOperations(stream, resizeOp,
                    fk::Unary<fk::ColorConversion<fk::COLOR_RGBA2BGRA, float4, float4>>{},
                    mulOp, subOp, divOp, wrOp);
            } else {
                fk::executeOperations(stream, resizeOp,
                    fk::Unary<fk::ColorConversion<fk::COLOR_RGB2BGR, float3, float3>>{},
                    mulOp, subOp, divOp, wrOp);
            }
        } else {
            fk::executeOperations(stream, resizeOp, mulOp, subOp, divOp, wrOp);
        }
    };
    if constexpr (AR == fk::IGNORE_AR) {
        const auto resizeDFs = Resize::build_batch(readOP, sizeArr);
        launch(fk::BatchRead<BATCH, fk::CONDITIONAL_WITH_DEFAULT>::build(resizeDFs, used, bgv));
    } else {
        const auto bgArr = fk::make_set_std_array<BATCH>(bgv);
        const auto resizeDFs = Resize::build_batch(readOP, sizeArr, bgArr);
        launch(fk::BatchRead<BATCH, fk::CONDITIONAL_WITH_DEFAULT>::build(resizeDFs, used, bgv));
    }
    return 0;
}

template <typename PixelT, int BATCH>
int dispatch(const void* const* ptrs, const int* ws, const int* hs, const int* pitches,
             int used, int dst_w, int dst_h, int ar, const float* bg, int swap,
             const float* mul, const float* sub, const float* div, float* out, cudaStream_t s) {
#define FKREF_CASE(ARV, SW) \
    if (ar == (int)ARV && swap == SW) \
        return run_chain<PixelT, BATCH, ARV, SW != 0>(ptrs, ws, hs, pitches, used, dst_w, dst_h, bg, mul, sub, div, out, s);
    FKREF_CASE(fk::IGNORE_AR, 0) FKREF_CASE(fk::IGNORE_AR, 1)
#ifdef FKREF_ALL_AR
    FKREF_CASE(fk::PRESERVE_AR, 0) FKREF_CASE(fk::PRESERVE_AR, 1)
    FKREF_CASE(fk::PRESERVE_AR_RN_EVEN, 0) FKREF_CASE(fk::PRESERVE_AR_RN_EVEN, 1)
    FKREF_CASE(fk::PRESERVE_AR_LEFT, 0) FKREF_CASE(fk::PRESERVE_AR_LEFT, 1)
#endif
#undef FKREF_CASE
    g_err = "fkref: unsupported (aspect mode, swap) combination in this build";
    return -1;
}

}  // namespace

#define FKREF_CAT_(a, b) a##b
#define FKREF_CAT(a, b) FKREF_CAT_(a, b)

extern "C" {

// Batch size this translation unit was compiled for (a template parameter in the reference).
int FKREF_CAT(fkref_batch_, FKREF_BATCH)(void) { return FKREF_BATCH; }

// resize(bilinear) -> [RGB2BGR] -> Mul -> Sub -> Div -> TensorSplit; `out` must hold FKREF_BATCH planes.
// Asynchronous on `stream`, like the reference. Returns 0, or -1 with fkref_last_error set.
int FKREF_CAT(fkref_preproc_, FKREF_BATCH)(const void* const* ptrs, const int* ws, const int* hs, const int* pitches,
                  int used, int dst_w, int dst_h, int aspect_mode, const float* bg, int swap_rb,
                  const float* mul, const float* sub, const float* div, float* out, void* stream) {
    try {
        return dispatch<uchar3, FKREF_BATCH>(ptrs, ws, hs, pitches, used, dst_w, dst_h, aspect_mode, bg, swap_rb,
                                             mul, sub, div, out, (cudaStream_t)stream);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// Frame loop in native code (what the reference's benchmarks time, tests/testsCommon.cuh:260-308): `steps`
// consecutive calls, call i using argument set i % n_sets -- so that a ctypes caller's per-call overhead is not
// charged to the reference.
int FKREF_CAT(fkref_preproc_sequence_, FKREF_BATCH)(const void* const* const* ptrs, const int* const* ws,
                  const int* const* hs, const int* const* pitches, int used, int dst_w, int dst_h, int aspect_mode,
                  const float* bg, int swap_rb, const float* mul, const float* sub, const float* div,
                  float* const* outs, int n_sets, int steps, void* stream) {
    try {
        for (int i = 0; i < steps; ++i) {
            const int s = i % n_sets;
            if (int rc = dispatch<uchar3, FKREF_BATCH>(ptrs[s], ws[s], hs[s], pitches[s], used, dst_w, dst_h, aspect_mode, bg,
                                               swap_rb, mul, sub, div, outs[s], (cudaStream_t)stream))
                return rc;
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// The same loop driven by `threads` host threads, one stream each (argument set s belongs to thread s % threads), forked
// from / joined into `stream` with events: the launch strategy of the product's multi-threaded frame loop, so that
// bench.py can compare kernels under the same strategy.
int FKREF_CAT(fkref_preproc_sequence_mt_, FKREF_BATCH)(const void* const* const* ptrs, const int* const* ws,
                  const int* const* hs, const int* const* pitches, int used, int dst_w, int dst_h, int aspect_mode,
                  const float* bg, int swap_rb, const float* mul, const float* sub, const float* div,
                  float* const* outs, int n_sets, int steps, void* stream, int threads) {
    constexpr int kMax = 8;
    static cudaStream_t st[kMax] = {};
    static cudaEvent_t done[kMax] = {}, start = nullptr;
    if (threads < 1 || threads > kMax) { g_err = "fkref: 1..8 threads"; return -1; }
    int device = 0;
    cudaGetDevice(&device);
    if (!start) {
        cudaEventCreateWithFlags(&start, cudaEventDisableTiming);
        for (int i = 0; i < kMax; ++i) {
            cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
        }
    }
    cudaEventRecord(start, (cudaStream_t)stream);
    int rcs[kMax] = {};
    std::string errs[kMax];
    std::thread pool[kMax];
    for (int w = 0; w < threads; ++w) {
        pool[w] = std::thread([&, w] {
            cudaSetDevice(device);
            cudaStreamWaitEvent(st[w], start, 0);
            try {
                for (int i = 0; i < steps && rcs[w] == 0; ++i) {
                    const int s = i % n_sets;
                    if (s % threads != w) continue;
                    rcs[w] = dispatch<uchar3, FKREF_BATCH>(ptrs[s], ws[s], hs[s], pitches[s], used, dst_w, dst_h, aspect_mode, bg,
                                                           swap_rb, mul, sub, div, outs[s], st[w]);
                }
            } catch (const std::exception& e) {
                errs[w] = e.what();
                rcs[w] = -1;
            }
            cudaEventRecord(done[w], st[w]);
        });
    }
    int rc = 0;
    for (int w = 0; w < threads; ++w) {
        pool[w].join();
        cudaStreamWaitEvent((cudaStream_t)stream, done[w], 0);
        if (rcs[w] != 0 && rc == 0) { rc = rcs[w]; g_err = errs[w]; }
    }
    return rc;
}

#ifdef FKREF_CT
// fk::CircularTensor<float, 3, BATCH, ORDER, Standard>::update (circular_tensor.cuh:111-146) with the pipeline of the
// cvGS wrapper (include/cvGPUSpeedup.cuh:600-627): Read<PerThreadRead<_2D, uchar3>> [-> Resize<INTER_LINEAR> when the
// frame is not plane-sized] -> SaturateCast<uchar3, float3> / the resize's float3 -> [RGB2BGR] -> Mul -> Sub -> Div ->
// TensorSplit<float3>.  BATCH is a template parameter: depths 4 and 16 are instantiated.
}  // extern "C"
namespace {
struct CtBase {
    virtual ~CtBase() {}
    virtual int update(const void* data, int w, int h, int pitch, int swap, const float* mul, const float* sub, const float* div,
                       cudaStream_t s) = 0;
    virtual float* data() = 0;
    virtual size_t bytes() = 0;
};
template <int BATCH, fk::CircularTensorOrder ORDER>
struct CtImpl : CtBase {
    fk::CircularTensor<float, 3, BATCH, ORDER, fk::ColorPlanes::Standard> t;
    int W, H;
    CtImpl(int w, int h) : t((uint)w, (uint)h), W(w), H(h) {}
    float* data() override { return t.ptr().data; }
    size_t bytes() override { return t.sizeInBytes(); }
    int update(const void* data, int w, int h, int pitch, int swap, const float* mul, const float* sub, const float* div,
               cudaStream_t s) override {
        const fk::RawPtr<fk::_2D, uchar3> img{ (uchar3*)data, { (uint)w, (uint)h, (uint)pitch } };
        const auto mulOp = fk::Binary<fk::Mul<float3>>{ float3{ mul[0], mul[1], mul[2] } };
        const auto subOp = fk::Binary<fk::Sub<float3>>{ float3{ sub[0], sub[1], sub[2] } };
        const auto divOp = fk::Binary<fk::Div<float3>>{ float3{ div[0], div[1], div[2] } };
        const auto wrOp = fk::Write<fk::TensorSplit<float3>>{ t.ptr() };
        const auto swapOp = fk::Unary<fk::ColorConversion<fk::COLOR_RGB2BGR, float3, float3>>{};
        if (w == W && h == H) {
            const auto rd = fk::Read<fk::PerThreadRead<fk::_2D, uchar3>>{ { img } };
            const auto cast = fk::Unary<fk::SaturateCast<uchar3, float3>>{};
            if (swap) t.update(s, rd, cast, swapOp, mulOp, subOp, divOp, wrOp);
            else t.update(s, rd, cast, mulOp, subOp, divOp, wrOp);
        } else {
            const auto rd = fk::Resize<fk::INTER_LINEAR>::build(fk::Read<fk::PerThreadRead<fk::_2D, uchar3>>{ { img } }, fk::Size(W, H));
            if (swap) t.update(s, rd, swapOp, mulOp, subOp, divOp, wrOp);
            else t.update(s, rd, mulOp, subOp, divOp, wrOp);
        }
        return 0;
    }
};
}  // namespace
extern "C" {
// order: 0 NewestFirst, 1 OldestFirst.  Returns a handle or NULL (unsupported depth).
void* fkref_ct_create(int batch, int order, int w, int h) {
    try {
        using O = fk::CircularTensorOrder;
        if (batch == 16 && order == 0) return new CtImpl<16, O::NewestFirst>(w, h);
        if (batch == 16 && order == 1) return new CtImpl<16, O::OldestFirst>(w, h);
        if (batch == 4 && order == 0) return new CtImpl<4, O::NewestFirst>(w, h);
        if (batch == 4 && order == 1) return new CtImpl<4, O::OldestFirst>(w, h);
        g_err = "fkref_ct: depth 4 or 16";
    } catch (const std::exception& e) {
        g_err = e.what();
    }
    return nullptr;
}
int fkref_ct_update(void* h, const void* data, int w, int hgt, int pitch, int swap_rb, const float* mul, const float* sub,
                    const float* div, void* stream) {
    try {
        return static_cast<CtBase*>(h)->update(data, w, hgt, pitch, swap_rb, mul, sub, div, (cudaStream_t)stream);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
float* fkref_ct_data(void* h) { return static_cast<CtBase*>(h)->data(); }
unsigned long long fkref_ct_bytes(void* h) { return static_cast<CtBase*>(h)->bytes(); }
void fkref_ct_destroy(void* h) { delete static_cast<CtBase*>(h); }
#endif

#ifdef FKREF_16BIT
// The same chain on CV_16UC3 (pixel_type 18, ushort3) and CV_16SC3 (19, short3) sources.
int FKREF_CAT(fkref_preproc16_, FKREF_BATCH)(int pixel_type, const void* const* ptrs, const int* ws, const int* hs,
                  const int* pitches, int used, int dst_w, int dst_h, int aspect_mode, const float* bg, int swap_rb,
                  const float* mul, const float* sub, const float* div, float* out, void* stream) {
    try {
        if (pixel_type == 18)
            return dispatch<ushort3, FKREF_BATCH>(ptrs, ws, hs, pitches, used, dst_w, dst_h, aspect_mode, bg, swap_rb,
                                                  mul, sub, div, out, (cudaStream_t)stream);
        if (pixel_type == 19)
            return dispatch<short3, FKREF_BATCH>(ptrs, ws, hs, pitches, used, dst_w, dst_h, aspect_mode, bg, swap_rb,
                                                 mul, sub, div, out, (cudaStream_t)stream);
        // 4-channel sources: bg / mul / sub / div hold four values, `out` holds FKREF_BATCH x 4 planes
        if (pixel_type == 24)
            return dispatch<uchar4, FKREF_BATCH>(ptrs, ws, hs, pitches, used, dst_w, dst_h, aspect_mode, bg, swap_rb,
                                                 mul, sub, div, out, (cudaStream_t)stream);
        if (pixel_type == 26)
            return dispatch<ushort4, FKREF_BATCH>(ptrs, ws, hs, pitches, used, dst_w, dst_h, aspect_mode, bg, swap_rb,
                                                  mul, sub, div, out, (cudaStream_t)stream);
        if (pixel_type == 27)
            return dispatch<short4, FKREF_BATCH>(ptrs, ws, hs, pitches, used, dst_w, dst_h, aspect_mode, bg, swap_rb,
                                                 mul, sub, div, out, (cudaStream_t)stream);
        g_err = "fkref: pixel type must be a 16-bit 3-channel or any 4-channel OpenCV type code";
        return -1;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
#endif

#ifdef FKREF_NV12
// One NV12 frame: Read<ReadYUV<NV12>> fused with ConvertYUVToRGB<NV12, range, primaries, false, float3> as the
// back-function of Resize<INTER_LINEAR> (tests/resize/test_fused_resize.cu:73-76,141-143), then Mul, Sub, Div and
// TensorSplit into out[3][dst_h][dst_w].  standard: 0 bt601 full, 1 bt709 full, 2 bt709 limited, 3 bt2020 full.
}  // extern "C"
namespace {
template <fk::ColorRange CR, fk::ColorPrimitives CP>
int run_nv12(const void* data, int w, int h, int pitch, int dst_w, int dst_h, const float* mul, const float* sub,
             const float* div, float* out, cudaStream_t stream) {
    const fk::RawPtr<fk::_2D, uchar> img{ (uchar*)data, { (uint)w, (uint)h, (uint)pitch } };
    const auto readBackOp = fk::fuse(fk::Read<fk::ReadYUV<fk::NV12>>{ img },
                                     fk::Unary<fk::ConvertYUVToRGB<fk::NV12, CR, CP, false, float3>>{});
    const auto readOp = fk::Resize<fk::INTER_LINEAR>::build(readBackOp, fk::Size(dst_w, dst_h));
    const fk::Tensor<float> t_out(out, dst_w, dst_h, 1, 3);
    fk::executeOperations(stream, readOp, fk::Binary<fk::Mul<float3>>{ float3{mul[0], mul[1], mul[2]} },
                          fk::Binary<fk::Sub<float3>>{ float3{sub[0], sub[1], sub[2]} },
                          fk::Binary<fk::Div<float3>>{ float3{div[0], div[1], div[2]} },
                          fk::Write<fk::TensorSplit<float3>>{ t_out.ptr() });
    return 0;
}
}  // namespace
extern "C" {
int FKREF_CAT(fkref_nv12_, FKREF_BATCH)(int standard, const void* data, int w, int h, int pitch, int dst_w, int dst_h,
                  const float* mul, const float* sub, const float* div, float* out, void* stream) {
    try {
        cudaStream_t s = (cudaStream_t)stream;
        switch (standard) {
            case 0: return run_nv12<fk::Full, fk::bt601>(data, w, h, pitch, dst_w, dst_h, mul, sub, div, out, s);
            case 1: return run_nv12<fk::Full, fk::bt709>(data, w, h, pitch, dst_w, dst_h, mul, sub, div, out, s);
            case 2: return run_nv12<fk::Limited, fk::bt709>(data, w, h, pitch, dst_w, dst_h, mul, sub, div, out, s);
            case 3: return run_nv12<fk::Full, fk::bt2020>(data, w, h, pitch, dst_w, dst_h, mul, sub, div, out, s);
            default: g_err = "fkref: bad yuv standard"; return -1;
        }
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
#endif

#ifdef FKREF_YUV
// The other ReadYUV formats (NV21 8-bit; P010, P210, Y210 10-bit in 16-bit words), same chain as fkref_nv12.
// format: 2 NV21, 3 P010, 4 P210, 5 Y210 (CVGS_NVxx & 0xf).  Instantiated for bt601 full (0) and bt2020 full (3).
}  // extern "C"
namespace {
template <fk::PixelFormat PF, fk::ColorRange CR, fk::ColorPrimitives CP>
int run_yuv(const void* data, int w, int h, int pitch, int dst_w, int dst_h, const float* mul, const float* sub,
            const float* div, float* out, cudaStream_t stream) {
    using Base = typename fk::ReadYUV<PF>::PixelBaseType;
    const fk::RawPtr<fk::_2D, Base> img{ (Base*)data, { (uint)w, (uint)h, (uint)pitch } };
    const auto readBackOp = fk::fuse(fk::Read<fk::ReadYUV<PF>>{ img },
                                     fk::Unary<fk::ConvertYUVToRGB<PF, CR, CP, false, float3>>{});
    const auto readOp = fk::Resize<fk::INTER_LINEAR>::build(readBackOp, fk::Size(dst_w, dst_h));
    const fk::Tensor<float> t_out(out, dst_w, dst_h, 1, 3);
    fk::executeOperations(stream, readOp, fk::Binary<fk::Mul<float3>>{ float3{mul[0], mul[1], mul[2]} },
                          fk::Binary<fk::Sub<float3>>{ float3{sub[0], sub[1], sub[2]} },
                          fk::Binary<fk::Div<float3>>{ float3{div[0], div[1], div[2]} },
                          fk::Write<fk::TensorSplit<float3>>{ t_out.ptr() });
    return 0;
}
template <fk::PixelFormat PF>
int run_yuv_std(int standard, const void* data, int w, int h, int pitch, int dst_w, int dst_h, const float* mul,
                const float* sub, const float* div, float* out, cudaStream_t s) {
    if (standard == 0) return run_yuv<PF, fk::Full, fk::bt601>(data, w, h, pitch, dst_w, dst_h, mul, sub, div, out, s);
    if (standard == 3) return run_yuv<PF, fk::Full, fk::bt2020>(data, w, h, pitch, dst_w, dst_h, mul, sub, div, out, s);
    g_err = "fkref: yuv standard not instantiated for this format";
    return -1;
}
}  // namespace
extern "C" {
int FKREF_CAT(fkref_yuv_, FKREF_BATCH)(int format, int standard, const void* data, int w, int h, int pitch, int dst_w,
                                       int dst_h, const float* mul, const float* sub, const float* div, float* out, void* stream) {
    try {
        cudaStream_t s = (cudaStream_t)stream;
        switch (format) {
            case 2: return run_yuv_std<fk::NV21>(standard, data, w, h, pitch, dst_w, dst_h, mul, sub, div, out, s);
            case 3: return run_yuv_std<fk::P010>(standard, data, w, h, pitch, dst_w, dst_h, mul, sub, div, out, s);
            case 4: return run_yuv_std<fk::P210>(standard, data, w, h, pitch, dst_w, dst_h, mul, sub, div, out, s);
            case 5: return run_yuv_std<fk::Y210>(standard, data, w, h, pitch, dst_w, dst_h, mul, sub, div, out, s);
            default: g_err = "fkref: bad yuv format"; return -1;
        }
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
#endif

#ifdef FKREF_WARP
// One image through fk::Warping<WT, PerThreadRead<_2D, uchar3>> (tests/warping/test_warping_opencv.cu:60-64):
//   mode 0:  warp -> Mul(mul) -> TensorSplit into float out[3][dst_h][dst_w]
//   mode 1:  warp -> Cast<float3, uchar3> -> PerThreadWrite into packed uchar3 out (row pitch out_pitch bytes)
// m is the inverse transform the wrapper hands to the kernel (cvGPUSpeedup.cuh:266-283), row-major.
}  // extern "C"
#include <fused_kernel/algorithms/image_processing/warping.cuh>
#include <fused_kernel/algorithms/basic_ops/cast.cuh>
namespace {
// Other pixel types of cvGS::warp<WT, InputType>: warp -> Mul -> TensorSplit into float out[C][dst_h][dst_w].
template <fk::WarpType WT, typename PixelT>
int run_warp_typed(const void* data, int w, int h, int pitch, const float* m, int dst_w, int dst_h, const float* mul,
                   float* out, cudaStream_t stream) {
    const auto read = fk::PerThreadRead<fk::_2D, PixelT>::build(
        fk::RawPtr<fk::_2D, PixelT>{ (PixelT*)data, { (uint)w, (uint)h, (uint)pitch } });
    fk::WarpingParameters<WT> params{};
    for (int r = 0; r < (WT == fk::Affine ? 2 : 3); ++r)
        for (int c = 0; c < 3; ++c) params.transformMatrix.data[r][c] = m[3 * r + c];
    params.dstSize = fk::Size(dst_w, dst_h);
    const auto warp = fk::Warping<WT, std::decay_t<decltype(read)>>::build({ params, read });
    constexpr int CN = fk::cn<PixelT>;
    using F = fk::VectorType_t<float, CN>;
    F mv;
    mv.x = mul[0]; mv.y = mul[1]; mv.z = mul[2];
    if constexpr (CN == 4) mv.w = mul[3];
    const fk::Tensor<float> t_out(out, dst_w, dst_h, 1, CN);
    fk::executeOperations(stream, warp, fk::Binary<fk::Mul<F>>{ mv }, fk::Write<fk::TensorSplit<F>>{ t_out.ptr() });
    return 0;
}
template <fk::WarpType WT>
int run_warp(int mode, const void* data, int w, int h, int pitch, const float* m, int dst_w, int dst_h, const float* mul,
             void* out, int out_pitch, cudaStream_t stream) {
    const auto read = fk::PerThreadRead<fk::_2D, uchar3>::build(
        fk::RawPtr<fk::_2D, uchar3>{ (uchar3*)data, { (uint)w, (uint)h, (uint)pitch } });
    fk::WarpingParameters<WT> params{};
    for (int r = 0; r < (WT == fk::Affine ? 2 : 3); ++r)
        for (int c = 0; c < 3; ++c) params.transformMatrix.data[r][c] = m[3 * r + c];
    params.dstSize = fk::Size(dst_w, dst_h);
    const auto warp = fk::Warping<WT, std::decay_t<decltype(read)>>::build({ params, read });
    if (mode == 0) {
        const fk::Tensor<float> t_out((float*)out, dst_w, dst_h, 1, 3);
        fk::executeOperations(stream, warp, fk::Binary<fk::Mul<float3>>{ float3{mul[0], mul[1], mul[2]} },
                              fk::Write<fk::TensorSplit<float3>>{ t_out.ptr() });
    } else {
        const fk::RawPtr<fk::_2D, uchar3> o{ (uchar3*)out, { (uint)dst_w, (uint)dst_h, (uint)out_pitch } };
        fk::executeOperations(stream, warp, fk::Unary<fk::Cast<float3, uchar3>>{},
                              fk::Write<fk::PerThreadWrite<fk::_2D, uchar3>>{ o });
    }
    return 0;
}
}  // namespace
extern "C" {
int FKREF_CAT(fkref_warp_, FKREF_BATCH)(int type, int mode, const void* data, int w, int h, int pitch, const float* m,
                                        int dst_w, int dst_h, const float* mul, void* out, int out_pitch, void* stream) {
    try {
        cudaStream_t s = (cudaStream_t)stream;
        if (type == 0) return run_warp<fk::Affine>(mode, data, w, h, pitch, m, dst_w, dst_h, mul, out, out_pitch, s);
        return run_warp<fk::Perspective>(mode, data, w, h, pitch, m, dst_w, dst_h, mul, out, out_pitch, s);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
// src_type: the CV type code (24 CV_8UC4, 18 CV_16UC3, 27 CV_16SC4); perspective (type 1) and affine (type 0)
int FKREF_CAT(fkref_warp_typed_, FKREF_BATCH)(int type, int src_type, const void* data, int w, int h, int pitch, const float* m,
                                              int dst_w, int dst_h, const float* mul, float* out, void* stream) {
    try {
        cudaStream_t s = (cudaStream_t)stream;
#define FKREF_WARP_CASE(CODE, PIX)                                                                                          \
    case CODE:                                                                                                              \
        return type == 0 ? run_warp_typed<fk::Affine, PIX>(data, w, h, pitch, m, dst_w, dst_h, mul, out, s)                 \
                         : run_warp_typed<fk::Perspective, PIX>(data, w, h, pitch, m, dst_w, dst_h, mul, out, s);
        switch (src_type) {
            FKREF_WARP_CASE(24, uchar4)
            FKREF_WARP_CASE(18, ushort3)
            FKREF_WARP_CASE(27, short4)
            default: g_err = "fkref: warp pixel type not instantiated"; return -1;
        }
#undef FKREF_WARP_CASE
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
#endif

#ifdef FKREF_CVT
// Colour conversions that change the channel count (color_conversion.cuh:364-461), fused behind the resize like
// cvGS::cvtColor<CODE, CV_32FCn, CV_32FCm> puts them (include/cvGPUSpeedup.cuh:151-161):
//   Resize<INTER_LINEAR>(PerThreadRead<uchar3 | uchar4>) -> ColorConversion<CODE, I, O> -> Mul<O> -> Sub<O> -> write
// O = float4 / float3: TensorSplit into out[C][dst_h][dst_w]; O = float (gray): PerThreadWrite into out[dst_h][dst_w].
// code: the cv:: / fk:: ColorConversionCodes value (0 BGR2BGRA, 1 BGRA2BGR, 2 BGR2RGBA, 3 RGBA2BGR, 6 BGR2GRAY,
// 7 RGB2GRAY, 10 BGRA2GRAY, 11 RGBA2GRAY).
}  // extern "C"
namespace {
template <fk::ColorConversionCodes CODE, typename SrcT, typename O>
int run_cvt(const void* data, int w, int h, int pitch, int dst_w, int dst_h, const float* mul, const float* sub, float* out,
            cudaStream_t stream) {
    using I = fk::VectorType_t<float, fk::cn<SrcT>>;
    const auto read = fk::PerThreadRead<fk::_2D, SrcT>::build(fk::RawPtr<fk::_2D, SrcT>{ (SrcT*)data, { (uint)w, (uint)h, (uint)pitch } });
    const auto rs = fk::Resize<fk::INTER_LINEAR>::build(read, fk::Size(dst_w, dst_h));
    const auto cvt = fk::Unary<fk::ColorConversion<CODE, I, O>>{};
    if constexpr (std::is_same_v<O, float>) {
        const fk::RawPtr<fk::_2D, float> o{ out, { (uint)dst_w, (uint)dst_h, (uint)(dst_w * 4) } };
        fk::executeOperations(stream, rs, cvt, fk::Binary<fk::Mul<float>>{ mul[0] }, fk::Binary<fk::Sub<float>>{ sub[0] },
                              fk::Write<fk::PerThreadWrite<fk::_2D, float>>{ o });
    } else {
        O m, b;
        m.x = mul[0]; m.y = mul[1]; m.z = mul[2];
        b.x = sub[0]; b.y = sub[1]; b.z = sub[2];
        if constexpr (fk::cn<O> == 4) { m.w = mul[3]; b.w = sub[3]; }
        const fk::Tensor<float> t_out(out, dst_w, dst_h, 1, fk::cn<O>);
        fk::executeOperations(stream, rs, cvt, fk::Binary<fk::Mul<O>>{ m }, fk::Binary<fk::Sub<O>>{ b },
                              fk::Write<fk::TensorSplit<O>>{ t_out.ptr() });
    }
    return 0;
}
}  // namespace
extern "C" {
int FKREF_CAT(fkref_cvt_, FKREF_BATCH)(int code, const void* data, int w, int h, int pitch, int dst_w, int dst_h,
                                       const float* mul, const float* sub, float* out, void* stream) {
    try {
        cudaStream_t s = (cudaStream_t)stream;
        switch (code) {
            case 0: return run_cvt<fk::COLOR_BGR2BGRA, uchar3, float4>(data, w, h, pitch, dst_w, dst_h, mul, sub, out, s);
            case 1: return run_cvt<fk::COLOR_BGRA2BGR, uchar4, float3>(data, w, h, pitch, dst_w, dst_h, mul, sub, out, s);
            case 2: return run_cvt<fk::COLOR_BGR2RGBA, uchar3, float4>(data, w, h, pitch, dst_w, dst_h, mul, sub, out, s);
            case 3: return run_cvt<fk::COLOR_RGBA2BGR, uchar4, float3>(data, w, h, pitch, dst_w, dst_h, mul, sub, out, s);
            case 6: return run_cvt<fk::COLOR_BGR2GRAY, uchar3, float>(data, w, h, pitch, dst_w, dst_h, mul, sub, out, s);
            case 7: return run_cvt<fk::COLOR_RGB2GRAY, uchar3, float>(data, w, h, pitch, dst_w, dst_h, mul, sub, out, s);
            case 10: return run_cvt<fk::COLOR_BGRA2GRAY, uchar4, float>(data, w, h, pitch, dst_w, dst_h, mul, sub, out, s);
            case 11: return run_cvt<fk::COLOR_RGBA2GRAY, uchar4, float>(data, w, h, pitch, dst_w, dst_h, mul, sub, out, s);
            default: g_err = "fkref: colour conversion code not instantiated"; return -1;
        }
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
#endif

const char* FKREF_CAT(fkref_last_error_, FKREF_BATCH)(void) { return g_err.c_str(); }

}  // extern "C"
