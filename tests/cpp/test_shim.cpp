// tests/cpp/test_shim.cpp -- the reference's own test programs for this path, restated against the header shim
// (cvgpuspeedup_b200/include/cvGPUSpeedup.cuh): same cvGS calls, same constants, same acceptance rules.
//   1. tests/batchresize/test_batchresize_x_split3D.cu:58-78,311-314  constant-image batch pipeline, |diff| <= 1e-4
//   2. tests/batchread/test_circularbatchread_x_write3D.cu:286-335,391,448  CircularTensor after 100 updates
//   3. tests/unit_tests/test_split.cu:47-62  (1,2,3) -> planes 1,2,3
//   4. random image vs the CPU oracle, bit for bit (the reference has no such test: SURVEY F2)
//   6. tests/warping/test_warping_opencv.cu:34-197  affine / perspective / batched perspective warp + fk::Cast + write,
//      against the oracle (cv::cuda::warpAffine is not in this image), plus exact checks of the translation case
//   7. cvGS::cvtColor codes that add / drop the alpha channel or reduce to gray, against the oracle
//   8. one image in / one pitched image out (tests/read/test_read_x_write.cu) and the CV_8UC4 round trip of
//      tests/resize/test_resize_CPUvsGPUresults.cu
//   5. error convention: std::runtime_error (gpuErrchk, fkl/.../core/utils/utils.h:42-60)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../cvgpuspeedup_b200/include/cvGPUSpeedup.cuh"

extern "C" int oracle_warp(const cvgs_crop_t* images, const cvgs_warp_t* warps, int n_planes, int used, const cvgs_pipeline_t* p,
                           int nthreads);
extern "C" int oracle_preproc(const cvgs_crop_t* crops, int n_planes, int used, const cvgs_pipeline_t* p, int nthreads);

#define REQUIRE(cond)                                                        \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
            return 1;                                                        \
        }                                                                    \
    } while (0)

static int test_batchresize_x_split3D() {
    constexpr int BATCH = 50;
    constexpr int CV_TYPE_I = CV_8UC3, CV_TYPE_O = CV_32FC3;
    const cv::Scalar val_init(5, 5, 5), val_alpha(0.3, 0.3, 0.3), val_sub(1, 4, 3.2), val_div(3.2, 0.6, 11.8);
    const cv::Size up(64, 128);
    cudaStream_t s;
    cudaStreamCreate(&s);
    cv::cuda::Stream cv_stream = cv::cuda::StreamAccessor::wrapStream(s);
    cv::cuda::GpuMat d_input(2160, 3840, CV_TYPE_I, val_init);
    std::array<cv::cuda::GpuMat, BATCH> crops;
    for (int i = 0; i < BATCH; ++i) crops[i] = d_input(cv::Rect(i, i, 60, 120));
    cv::cuda::GpuMat d_tensor_output(BATCH, up.width * up.height * 3, CV_32FC1);
    d_tensor_output.step = up.width * up.height * 3 * sizeof(float);  // as the reference test does (:86-87)

    cvGS::executeOperations(cv_stream, cvGS::resize<CV_TYPE_I, cv::INTER_LINEAR, BATCH>(crops, up, BATCH),
                            cvGS::cvtColor<cv::COLOR_RGB2BGR, CV_TYPE_O>(), cvGS::multiply<CV_TYPE_O>(val_alpha),
                            cvGS::subtract<CV_TYPE_O>(val_sub), cvGS::divide<CV_TYPE_O>(val_div),
                            cvGS::split<CV_TYPE_O>(d_tensor_output, up));
    cv_stream.waitForCompletion();
    std::vector<float> h(static_cast<size_t>(BATCH) * 3 * up.width * up.height);
    REQUIRE(cudaMemcpy(h.data(), d_tensor_output.data, h.size() * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess);
    const size_t plane = static_cast<size_t>(up.width) * up.height;
    for (int z = 0; z < BATCH; ++z)
        for (int c = 0; c < 3; ++c) {
            const float want = (5.0f * 0.3f - static_cast<float>(val_sub[c])) / static_cast<float>(val_div[c]);
            for (size_t i = 0; i < plane; ++i) REQUIRE(std::fabs(h[(z * 3 + c) * plane + i] - want) <= 1e-4f);
        }
    cudaStreamDestroy(s);
    return 0;
}

template <fk::CircularTensorOrder ORDER, fk::ColorPlanes MODE>
static int test_circular_tensor() {
    constexpr int BATCH = 15, WIDTH = 128, HEIGHT = 128, ITERS = 100;
    cvGS::CircularTensor<CV_8UC3, CV_32F, 3, BATCH, ORDER, MODE> myTensor(WIDTH, HEIGHT);
    cv::cuda::GpuMat input(HEIGHT, WIDTH, CV_8UC3);
    cudaStream_t s;
    cudaStreamCreate(&s);
    cv::cuda::Stream cv_stream = cv::cuda::StreamAccessor::wrapStream(s);
    for (int i = 0; i < ITERS; ++i) {
        input.setTo(cv::Scalar(i + 1, i + 1, i + 1));
        if constexpr (MODE == fk::ColorPlanes::Standard)
            myTensor.update(cv_stream, input, cvGS::convertTo<CV_8UC3, CV_32FC3>(), cvGS::split<CV_32FC3>(myTensor.ptr()));
        else
            myTensor.update(cv_stream, input, cvGS::convertTo<CV_8UC3, CV_32FC3>(), cvGS::splitT<CV_32FC3>(myTensor.ptr()));
        cv_stream.waitForCompletion();
    }
    std::vector<float> h(myTensor.sizeInBytes() / sizeof(float));
    REQUIRE(cudaMemcpy(h.data(), myTensor.data(), myTensor.sizeInBytes(), cudaMemcpyDeviceToHost) == cudaSuccess);
    const size_t px = static_cast<size_t>(WIDTH) * HEIGHT;
    for (int c = 0; c < 3; ++c)
        for (int z = 0; z < BATCH; ++z) {
            const float* plane = MODE == fk::ColorPlanes::Standard ? &h[(static_cast<size_t>(z) * 3 + c) * px]
                                                                   : &h[(static_cast<size_t>(c) * BATCH + z) * px];
            const float want = ORDER == fk::CircularTensorOrder::NewestFirst ? ITERS - z : ITERS - (BATCH - z - 1);
            for (size_t i = 0; i < px; ++i) REQUIRE(plane[i] == want);
        }
    cudaStreamDestroy(s);
    return 0;
}

// tests/batchread/test_circularbatchread_x_write3D.cu:400-460 (testOldestFirstCircularTensorcvGS_noSplit): CV_8UC4 frames
// into ONE plane of packed CV_32FC4 pixels, OldestFirst; after 100 updates plane z holds 100 - (BATCH - z - 1).
static int test_circular_tensor_no_split() {
    constexpr int BATCH = 15, WIDTH = 128, HEIGHT = 128, ITERS = 100;
    cvGS::CircularTensor<CV_8UC4, CV_32FC4, 1, BATCH, fk::CircularTensorOrder::OldestFirst> myTensor(WIDTH, HEIGHT);
    REQUIRE(myTensor.sizeInBytes() == sizeof(float) * 4 * WIDTH * HEIGHT * BATCH);
    cv::cuda::GpuMat input(HEIGHT, WIDTH, CV_8UC4);
    cv::cuda::Stream cv_stream;
    for (int i = 0; i < ITERS; ++i) {
        input.setTo(cv::Scalar::all(i + 1));
        myTensor.update(cv_stream, input, cvGS::convertTo<CV_8UC4, CV_32FC4>());
        cv_stream.waitForCompletion();
    }
    std::vector<float> h(myTensor.sizeInBytes() / sizeof(float));
    REQUIRE(cudaMemcpy(h.data(), myTensor.data(), myTensor.sizeInBytes(), cudaMemcpyDeviceToHost) == cudaSuccess);
    const size_t px = static_cast<size_t>(WIDTH) * HEIGHT * 4;
    for (int z = 0; z < BATCH; ++z)
        for (size_t i = 0; i < px; ++i) REQUIRE(h[z * px + i] == static_cast<float>(ITERS - (BATCH - z - 1)));
    return 0;
}

static int test_split() {
    cv::cuda::GpuMat d_input(16, 16, CV_8UC3, cv::Scalar(1, 2, 3));
    cv::cuda::GpuMat d_out(1, 16 * 16 * 3, CV_32FC1);
    cv::cuda::Stream st;
    cvGS::executeOperations(st, cvGS::resize<CV_8UC3, cv::INTER_LINEAR>(d_input, cv::Size(16, 16)),
                            cvGS::convertTo<CV_8UC3, CV_32FC3>(), cvGS::split<CV_32FC3>(d_out, cv::Size(16, 16)));
    st.waitForCompletion();
    std::vector<float> h(16 * 16 * 3);
    REQUIRE(cudaMemcpy(h.data(), d_out.data, h.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess);
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 256; ++i) REQUIRE(h[c * 256 + i] == static_cast<float>(c + 1));
    return 0;
}

// tests/resize/test_resize_x_split.cu:51,72-84: one crop -> resize -> alpha -> sub -> div -> split(vector<GpuMat>)
// (fk::SplitWrite), and the batch form split(array<vector<GpuMat>, N>); constant image so the answer is known.
static int test_resize_x_split_write() {
    cv::cuda::GpuMat d_input(480, 640, CV_8UC3, cv::Scalar(5, 5, 5));
    const cv::Size up(64, 128);
    const cv::Scalar sub(1, 4, 6), div(2, 8, 1);
    const double alpha = 0.5;
    cv::cuda::Stream st;
    std::vector<cv::cuda::GpuMat> planes(3);
    for (auto& m : planes) m = cv::cuda::GpuMat(up.height, up.width, CV_32FC1);
    cvGS::executeOperations(st, cvGS::resize<CV_8UC3, cv::INTER_LINEAR>(d_input(cv::Rect(200, 200, 60, 120)), up),
                            cvGS::multiply<CV_32FC3>(cv::Scalar(alpha, alpha, alpha)), cvGS::subtract<CV_32FC3>(sub),
                            cvGS::divide<CV_32FC3>(div), cvGS::split<CV_32FC3>(planes));
    constexpr int N = 3;
    std::array<cv::cuda::GpuMat, N> crops;
    std::array<std::vector<cv::cuda::GpuMat>, N> outs;
    for (int i = 0; i < N; ++i) {
        crops[i] = d_input(cv::Rect(i, i, 60, 120));
        outs[i].resize(3);
        for (auto& m : outs[i]) m = cv::cuda::GpuMat(up.height, up.width, CV_32FC1);
    }
    cvGS::executeOperations(st, cvGS::resize<CV_8UC3, cv::INTER_LINEAR, N, cvGS::IGNORE_AR>(crops, up, N),
                            cvGS::multiply<CV_32FC3>(cv::Scalar(alpha, alpha, alpha)), cvGS::subtract<CV_32FC3>(sub),
                            cvGS::divide<CV_32FC3>(div), cvGS::split<CV_32FC3, N>(outs));
    st.waitForCompletion();
    auto check = [&](const std::vector<cv::cuda::GpuMat>& v) {
        for (int c = 0; c < 3; ++c) {
            std::vector<float> h(static_cast<size_t>(up.width) * up.height);
            v[c].download(h.data(), up.width * sizeof(float));
            const float want = static_cast<float>((5.0 * alpha - sub[c]) / div[c]);
            for (float x : h)
                if (std::fabs(x - want) > 1e-4f) return 1;
        }
        return 0;
    };
    REQUIRE(check(planes) == 0);
    for (int i = 0; i < N; ++i) REQUIRE(check(outs[i]) == 0);
    return 0;
}

// tests/batchread/test_batchread_x_write3D.cu:83-97: batch read (no resize) -> convertTo(alpha) -> subtract -> divide
// -> write(tensor); images filled with a constant, so out = (v * alpha - sub) / div.  Also the (output, plane)
// overload with an active batch smaller than the batch and a default value.
static int test_batchread_x_write3D() {
    constexpr int BATCH = 6, W = 40, H = 24;
    const double alpha = 0.3;
    const cv::Scalar sub(1, 4, 3.2), div(3.2, 0.6, 11.8), def(7, 8, 9);
    cv::cuda::GpuMat d_big(H * 2, W * 3, CV_8UC3, cv::Scalar(5, 6, 7));
    std::array<cv::cuda::GpuMat, BATCH> crops;
    for (int i = 0; i < BATCH; ++i) crops[i] = d_big(cv::Rect(i, i, W, H));
    cv::cuda::GpuMat d_out(BATCH, W * H * 3, CV_32FC1);
    cv::cuda::Stream st;
    cvGS::executeOperations(crops, st, cvGS::convertTo<CV_8UC3, CV_32FC3>(alpha), cvGS::subtract<CV_32FC3>(sub),
                            cvGS::divide<CV_32FC3>(div), cvGS::write<CV_32FC3>(d_out, cv::Size(W, H)));
    st.waitForCompletion();
    std::vector<float> h(static_cast<size_t>(BATCH) * W * H * 3);
    REQUIRE(cudaMemcpy(h.data(), d_out.data, h.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess);
    const double src[3] = {5, 6, 7};
    for (size_t i = 0; i < h.size(); ++i) {
        const int c = static_cast<int>(i % 3);
        REQUIRE(std::fabs(h[i] - static_cast<float>((src[c] * alpha - sub[c]) / div[c])) <= 1e-4f);
    }
    cvGS::executeOperations(crops, static_cast<size_t>(BATCH - 2), def, d_out, cv::Size(W, H), st,
                            cvGS::convertTo<CV_8UC3, CV_32FC3>(alpha), cvGS::subtract<CV_32FC3>(sub), cvGS::divide<CV_32FC3>(div));
    st.waitForCompletion();
    REQUIRE(cudaMemcpy(h.data(), d_out.data, h.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess);
    for (size_t i = 0; i < h.size(); ++i) {
        const int c = static_cast<int>(i % 3), z = static_cast<int>(i / (static_cast<size_t>(W) * H * 3));
        const double v = z < BATCH - 2 ? src[c] : def[c];  // inactive planes: chain(default value), SURVEY F9
        REQUIRE(std::fabs(h[i] - static_cast<float>((v * alpha - sub[c]) / div[c])) <= 1e-4f);
    }
    return 0;
}

// tests/resize/test_resize_write.cu:55-56: resize -> convertTo<CV_32FC3, CV_8UC3>() -> write<CV_8UC3>(GpuMat), up and
// down, on a constant image (the 8-bit value survives the round trip) and into a pitched destination.
static int test_resize_write_8u() {
    cv::cuda::GpuMat d_input(120, 160, CV_8UC3, cv::Scalar(11, 129, 250));
    cv::cuda::Stream st;
    for (const cv::Size sz : {cv::Size(320, 200), cv::Size(50, 37)}) {
        cv::cuda::GpuMat d_out(sz.height, sz.width, CV_8UC3);
        cvGS::executeOperations(st, cvGS::resize<CV_8UC3, cv::INTER_LINEAR>(d_input, sz, 0., 0.),
                                cvGS::convertTo<CV_32FC3, CV_8UC3>(), cvGS::write<CV_8UC3>(d_out));
        st.waitForCompletion();
        std::vector<uchar> h(static_cast<size_t>(sz.width) * sz.height * 3);
        d_out.download(h.data(), sz.width * 3);
        for (size_t i = 0; i < h.size(); ++i) REQUIRE(h[i] == (i % 3 == 0 ? 11 : i % 3 == 1 ? 129 : 250));
    }
    bool threw = false;
    try {
        cv::cuda::GpuMat d_out(8, 8, CV_8UC3);
        cvGS::executeOperations(st, cvGS::resize<CV_8UC3, cv::INTER_LINEAR>(d_input, cv::Size(8, 8), 0., 0.), cvGS::write<CV_8UC3>(d_out));
    } catch (const std::runtime_error&) { threw = true; }
    REQUIRE(threw);  // an 8-bit write without the cast is a type error in the reference; here it throws
    return 0;
}

static int test_random_vs_oracle() {
    constexpr int BATCH = 8, W = 640, H = 480;
    std::mt19937 rng(7);
    cv::cuda::GpuMat d_img(H, W, CV_8UC3);
    std::vector<uchar> h_img(d_img.step * H);
    for (auto& b : h_img) b = static_cast<uchar>(rng());
    REQUIRE(cudaMemcpy(d_img.data, h_img.data(), h_img.size(), cudaMemcpyHostToDevice) == cudaSuccess);
    std::array<cv::cuda::GpuMat, BATCH> crops;
    std::vector<cvgs_crop_t> h_crops(BATCH);
    for (int i = 0; i < BATCH; ++i) {
        const cv::Rect r(3 + 37 * i, 5 + 11 * i, 24 + 40 * i, 48 + 50 * i);
        crops[i] = d_img(r);
        h_crops[i] = {h_img.data() + r.y * d_img.step + 3 * r.x, r.width, r.height, static_cast<int32_t>(d_img.step), 0};
    }
    const cv::Size up(64, 128);
    cv::cuda::GpuMat d_out(BATCH, up.width * up.height * 3, CV_32FC1);
    cv::cuda::Stream st;
    const cv::Scalar bg(128, 64, 32);
    cvGS::executeOperations(st, cvGS::resize<CV_8UC3, cv::INTER_LINEAR, BATCH, cvGS::PRESERVE_AR>(crops, up, BATCH - 2, bg),
                            cvGS::cvtColor<cv::COLOR_RGB2BGR, CV_32FC3>(), cvGS::multiply<CV_32FC3>(cv::Scalar(0.3, 0.3, 0.3)),
                            cvGS::subtract<CV_32FC3>(cv::Scalar(1, 4, 3.2)), cvGS::divide<CV_32FC3>(cv::Scalar(3.2, 0.6, 11.8)),
                            cvGS::split<CV_32FC3>(d_out, up));
    st.waitForCompletion();
    std::vector<float> got(static_cast<size_t>(BATCH) * 3 * up.width * up.height), want(got.size());
    REQUIRE(cudaMemcpy(got.data(), d_out.data, got.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess);
    cvgs_pipeline_t p{};
    p.src_type = CVGS_8UC3;
    p.dst_width = up.width;
    p.dst_height = up.height;
    p.aspect_mode = CVGS_PRESERVE_AR;
    p.background[0] = 128; p.background[1] = 64; p.background[2] = 32;
    p.n_ops = 4;
    p.ops[0].kind = CVGS_OP_REORDER; p.ops[0].perm[0] = 2; p.ops[0].perm[1] = 1; p.ops[0].perm[2] = 0;
    const float mul[3] = {0.3f, 0.3f, 0.3f}, sub[3] = {1.f, 4.f, 3.2f}, div[3] = {3.2f, 0.6f, 11.8f};
    p.ops[1].kind = CVGS_OP_MUL; p.ops[2].kind = CVGS_OP_SUB; p.ops[3].kind = CVGS_OP_DIV;
    for (int c = 0; c < 3; ++c) { p.ops[1].v[c] = mul[c]; p.ops[2].v[c] = sub[c]; p.ops[3].v[c] = div[c]; }
    p.out = want.data();
    REQUIRE(oracle_preproc(h_crops.data(), BATCH, BATCH - 2, &p, 0) == 0);
    REQUIRE(std::memcmp(got.data(), want.data(), got.size() * 4) == 0);
    return 0;
}

static int test_warping() {
    constexpr int W = 420, H = 410;
    std::mt19937 rng(11);
    cv::cuda::GpuMat d_img(H, W, CV_8UC3);
    std::vector<uchar> h_img(d_img.step * H);
    for (auto& b : h_img) b = static_cast<uchar>(rng());
    REQUIRE(cudaMemcpy(d_img.data, h_img.data(), h_img.size(), cudaMemcpyHostToDevice) == cudaSuccess);
    const cvgs_crop_t h_image{h_img.data(), W, H, static_cast<int32_t>(d_img.step), 0};
    cv::cuda::Stream stream;
    const cv::Size size(W, H);
    auto oracle_u8 = [&](const std::vector<cvgs_warp_t>& warps, std::vector<uchar>& out) {
        std::vector<cvgs_crop_t> images(warps.size(), h_image);
        cvgs_pipeline_t p{};
        p.src_type = CVGS_8UC3;
        p.dst_width = W;
        p.dst_height = H;
        p.aspect_mode = CVGS_IGNORE_AR;
        p.out_layout = CVGS_OUT_NHWC;
        p.dst_type = CVGS_8UC3;
        p.u8_cast = 1;
        out.assign(warps.size() * 3 * W * H, 0);
        p.out = out.data();
        return oracle_warp(images.data(), warps.data(), static_cast<int>(warps.size()), static_cast<int>(warps.size()), &p, 0);
    };
    auto download = [&](const cv::cuda::GpuMat& m, std::vector<uchar>& out) {
        out.resize(static_cast<size_t>(3) * m.cols * m.rows);
        return cudaMemcpy2D(out.data(), 3 * m.cols, m.data, m.step, 3 * m.cols, m.rows, cudaMemcpyDeviceToHost) == cudaSuccess;
    };

    // testAffine (:84-123): translation by (50, 100)
    {
        const double tx = 50, ty = 100;
        cv::Mat affine_matrix = (cv::Mat_<double>(2, 3) << 1, 0, tx, 0, 1, ty);
        cv::cuda::GpuMat d_result(H, W, CV_8UC3);
        const auto warpFunc = cvGS::warp<fk::WarpType::Affine, CV_8UC3>(d_img, affine_matrix, size);
        cvGS::executeOperations(stream, warpFunc, fk::Cast<float3, uchar3>::build(), cvGS::write<CV_8UC3>(d_result));
        stream.waitForCompletion();
        std::vector<uchar> got, want;
        REQUIRE(download(d_result, got));
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                for (int c = 0; c < 3; ++c) {
                    const uchar expect = (x >= 50 && y >= 100) ? h_img[(y - 100) * d_img.step + 3 * (x - 50) + c] : 0;
                    REQUIRE(got[(static_cast<size_t>(y) * W + x) * 3 + c] == expect);
                }
        cvgs_warp_t w{};
        w.type = CVGS_WARP_AFFINE;
        const float inv[6] = {1, 0, -50, 0, 1, -100};
        std::memcpy(w.m, inv, sizeof inv);
        REQUIRE(oracle_u8({w}, want) == 0);
        REQUIRE(got == want);
    }
    // testPerspective (:34-82) and testPerspectiveBatch (:125-197)
    {
        cv::Point2f src_points[4] = {cv::Point2f(56, 65), cv::Point2f(368, 52), cv::Point2f(28, 387), cv::Point2f(389, 390)};
        cv::Point2f dst1[4] = {cv::Point2f(0, 0), cv::Point2f(300, 0), cv::Point2f(0, 300), cv::Point2f(300, 300)};
        cv::Point2f dst2[4] = {cv::Point2f(0, 0), cv::Point2f(200, 0), cv::Point2f(0, 200), cv::Point2f(200, 200)};
        const cv::Mat m1 = cv::getPerspectiveTransform(src_points, dst1);
        const cv::Mat m2 = cv::getPerspectiveTransform(src_points, dst2);
        // the forward matrix maps the source points onto the destination points
        for (int i = 0; i < 4; ++i) {
            const double* h = m1.ptr<double>();
            const double d = h[6] * src_points[i].x + h[7] * src_points[i].y + h[8];
            REQUIRE(std::fabs((h[0] * src_points[i].x + h[1] * src_points[i].y + h[2]) / d - dst1[i].x) < 1e-6);
            REQUIRE(std::fabs((h[3] * src_points[i].x + h[4] * src_points[i].y + h[5]) / d - dst1[i].y) < 1e-6);
        }
        cv::cuda::GpuMat d_result(H, W, CV_8UC3);
        cvGS::executeOperations(stream, cvGS::warp<fk::WarpType::Perspective, CV_8UC3>(d_img, m1, size),
                                fk::Cast<float3, uchar3>::build(), cvGS::write<CV_8UC3>(d_result));
        stream.waitForCompletion();
        std::vector<uchar> got, want;
        REQUIRE(download(d_result, got));
        const cvgs_warp_t w1 = cvGS::detail::inverse_of<fk::WarpType::Perspective>(m1);
        const cvgs_warp_t w2 = cvGS::detail::inverse_of<fk::WarpType::Perspective>(m2);
        REQUIRE(oracle_u8({w1}, want) == 0);
        REQUIRE(got == want);
        size_t nonzero = 0;
        for (uchar b : got) nonzero += b != 0;
        REQUIRE(nonzero > got.size() / 4);

        constexpr size_t NUM_IMGS = 5;
        const std::array<cv::cuda::GpuMat, NUM_IMGS> d_imgs = {d_img, d_img, d_img, d_img, d_img};
        const std::array<cv::Mat, NUM_IMGS> mats = {m1, m2, m1, m2, m1};
        cv::cuda::GpuMat d_batch(static_cast<int>(NUM_IMGS) * H, W, CV_8UC3);
        d_batch.step = 3 * W;  // a tight tensor of five images, as write<CV_8UC3>(GpuMat, Size) defines it
        cvGS::executeOperations(stream, cvGS::warp<fk::WarpType::Perspective, CV_8UC3, NUM_IMGS>(d_imgs, mats, size),
                                fk::Cast<float3, uchar3>::build(), cvGS::write<CV_8UC3>(d_batch, size));
        stream.waitForCompletion();
        std::vector<uchar> gotb(NUM_IMGS * 3 * W * H), wantb;
        REQUIRE(cudaMemcpy(gotb.data(), d_batch.data, gotb.size(), cudaMemcpyDeviceToHost) == cudaSuccess);
        REQUIRE(oracle_u8({w1, w2, w1, w2, w1}, wantb) == 0);
        REQUIRE(gotb == wantb);
    }
    return 0;
}

// cvGS::cvtColor with the codes that change the channel count (reference include/cvGPUSpeedup.cuh:151-161,
// cv2cuda_types.cuh:77-86): against the oracle, plus exact checks of the alpha and gray values.
static int test_cvtcolor_channel_changes() {
    constexpr int W = 96, H = 64;
    std::mt19937 rng(21);
    cv::cuda::GpuMat d_img(H, W, CV_8UC3);
    std::vector<uchar> h_img(d_img.step * H);
    for (auto& b : h_img) b = static_cast<uchar>(rng());
    REQUIRE(cudaMemcpy(d_img.data, h_img.data(), h_img.size(), cudaMemcpyHostToDevice) == cudaSuccess);
    const cvgs_crop_t h_crop{h_img.data(), W, H, static_cast<int32_t>(d_img.step), 0};
    cv::cuda::Stream st;
    const cv::Size size(W, H);
    auto oracle = [&](const std::vector<cvgs_op_t>& ops, int nco, std::vector<float>& out) {
        cvgs_pipeline_t p{};
        p.src_type = CVGS_8UC3;
        p.dst_width = W;
        p.dst_height = H;
        p.aspect_mode = CVGS_IGNORE_AR;
        p.n_ops = static_cast<int>(ops.size());
        for (size_t i = 0; i < ops.size(); ++i) p.ops[i] = ops[i];
        out.assign(static_cast<size_t>(nco) * W * H, -1.f);
        p.out = out.data();
        return oracle_preproc(&h_crop, 1, 1, &p, 0);
    };
    auto op = [](int kind, float v0 = 0.f, float v1 = 0.f, float v2 = 0.f, float v3 = 0.f, int p0 = 0, int p1 = 1, int p2 = 2, int p3 = 3) {
        cvgs_op_t o{};
        o.kind = kind;
        o.v[0] = v0; o.v[1] = v1; o.v[2] = v2; o.v[3] = v3;
        o.perm[0] = p0; o.perm[1] = p1; o.perm[2] = p2; o.perm[3] = p3;
        return o;
    };
    {   // BGR -> RGBA, scaled per channel, planar CV_32FC4 tensor
        cv::cuda::GpuMat d_out(1, W * H * 4, CV_32FC1);
        cvGS::executeOperations(st, cvGS::resize<CV_8UC3, cv::INTER_LINEAR>(d_img, size),
                                cvGS::cvtColor<cv::COLOR_BGR2RGBA, CV_32FC3, CV_32FC4>(),
                                cvGS::multiply<CV_32FC4>(cv::Scalar(0.5, 0.25, 2.0, 1.0 / 255.0)), cvGS::split<CV_32FC4>(d_out, size));
        st.waitForCompletion();
        std::vector<float> got(static_cast<size_t>(4) * W * H), want;
        REQUIRE(cudaMemcpy(got.data(), d_out.data, got.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess);
        REQUIRE(oracle({op(CVGS_OP_REORDER, 0, 0, 0, 0, 2, 1, 0, 3), op(CVGS_OP_ADD_ALPHA, 255.f), op(CVGS_OP_MUL, 0.5f, 0.25f, 2.0f, 1.0f / 255.0f)}, 4, want) == 0);
        REQUIRE(std::memcmp(got.data(), want.data(), got.size() * 4) == 0);
        for (int i = 0; i < W * H; ++i) {
            REQUIRE(got[3 * W * H + i] == 255.f * (1.0f / 255.0f));
            REQUIRE(got[i] == 0.5f * h_img[(i / W) * d_img.step + 3 * (i % W) + 2]);  // R of a BGR pixel
        }
    }
    {   // RGB -> gray, one CV_32FC1 image
        cv::cuda::GpuMat d_out(H, W, CV_32FC1);
        d_out.step = W * sizeof(float);
        cvGS::executeOperations(st, cvGS::resize<CV_8UC3, cv::INTER_LINEAR>(d_img, size), cvGS::cvtColor<cv::COLOR_RGB2GRAY, CV_32FC3, CV_32FC1>(),
                                cvGS::write<CV_32FC1>(d_out));
        st.waitForCompletion();
        std::vector<float> got(static_cast<size_t>(W) * H), want;
        REQUIRE(cudaMemcpy(got.data(), d_out.data, got.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess);
        REQUIRE(oracle({op(CVGS_OP_GRAY, 0, 0, 0, 0, 1)}, 1, want) == 0);
        REQUIRE(std::memcmp(got.data(), want.data(), got.size() * 4) == 0);
        for (int i = 0; i < W * H; ++i) {
            const uchar* px = &h_img[(i / W) * d_img.step + 3 * (i % W)];
            const double lum = 0.299 * px[0] + 0.587 * px[1] + 0.114 * px[2];
            REQUIRE(got[i] == std::floor(got[i]) && std::fabs(got[i] - lum) <= 0.5 + 1e-3);
        }
    }
    {   // the write op names the pixel type the chain ends with
        cv::cuda::GpuMat d_out(1, W * H * 3, CV_32FC1);
        bool thrown = false;
        try {
            cvGS::executeOperations(st, cvGS::resize<CV_8UC3, cv::INTER_LINEAR>(d_img, size),
                                    cvGS::cvtColor<cv::COLOR_BGR2GRAY, CV_32FC3, CV_32FC1>(), cvGS::split<CV_32FC3>(d_out, size));
        } catch (const std::runtime_error& e) {
            thrown = std::strstr(e.what(), "dst_type") != nullptr;
        }
        REQUIRE(thrown);
    }
    return 0;
}

// tests/read/test_read_x_write.cu:60-75 (one image in, one pitched image out: convertTo, subtract, multiply, divide, add)
// and tests/resize/test_resize_CPUvsGPUresults.cu:45-47 (resize<CV_8UC4> -> convertTo<CV_32FC4, CV_8UC4> -> write<CV_8UC4>)
static int test_read_x_write_and_8uc4() {
    constexpr int W = 150, H = 70;
    std::mt19937 rng(31);
    cv::cuda::Stream st;
    {
        cv::cuda::GpuMat d_in(H, W, CV_8UC3), d_out(H, W, CV_32FC3);
        REQUIRE(d_out.step > static_cast<size_t>(W) * 12);  // the stand-in pads rows like cudaMallocPitch
        std::vector<uchar> h_in(d_in.step * H);
        for (auto& b : h_in) b = static_cast<uchar>(rng());
        REQUIRE(cudaMemcpy(d_in.data, h_in.data(), h_in.size(), cudaMemcpyHostToDevice) == cudaSuccess);
        REQUIRE(cudaMemset(d_out.data, 0xCD, d_out.step * H) == cudaSuccess);
        const cv::Scalar val_sub(1, 4, 3.2), val_mul(0.3, 0.5, 2.0), val_div(3.2, 0.6, 11.8), val_add(0.5, 1.5, 2.5);
        cvGS::executeOperations(d_in, d_out, st, cvGS::convertTo<CV_8UC3, CV_32FC3>(), cvGS::subtract<CV_32FC3>(val_sub),
                                cvGS::multiply<CV_32FC3>(val_mul), cvGS::divide<CV_32FC3>(val_div), cvGS::add<CV_32FC3>(val_add));
        st.waitForCompletion();
        std::vector<uchar> raw(d_out.step * H);
        REQUIRE(cudaMemcpy(raw.data(), d_out.data, raw.size(), cudaMemcpyDeviceToHost) == cudaSuccess);
        for (int y = 0; y < H; ++y) {
            const float* row = reinterpret_cast<const float*>(raw.data() + y * d_out.step);
            for (int x = 0; x < W; ++x)
                for (int c = 0; c < 3; ++c) {
                    const float px = h_in[y * d_in.step + 3 * x + c];
                    // separate roundings: sub, mul, div, add in the order of the chain (no mul directly before the add)
                    const float want = ((px - static_cast<float>(val_sub[c])) * static_cast<float>(val_mul[c])) / static_cast<float>(val_div[c]) +
                                       static_cast<float>(val_add[c]);
                    REQUIRE(row[3 * x + c] == want);
                }
            for (size_t b = static_cast<size_t>(W) * 12; b < d_out.step; ++b) REQUIRE(raw[y * d_out.step + b] == 0xCD);
        }
    }
    {
        cv::cuda::GpuMat d_in(H, W, CV_8UC4), d_out(40, 60, CV_8UC4);
        std::vector<uchar> h_in(d_in.step * H);
        for (auto& b : h_in) b = static_cast<uchar>(rng());
        REQUIRE(cudaMemcpy(d_in.data, h_in.data(), h_in.size(), cudaMemcpyHostToDevice) == cudaSuccess);
        cvGS::executeOperations(st, cvGS::resize<CV_8UC4, cv::INTER_LINEAR>(d_in, cv::Size(60, 40), 0., 0.),
                                cvGS::convertTo<CV_32FC4, CV_8UC4>(), cvGS::write<CV_8UC4>(d_out));
        st.waitForCompletion();
        std::vector<uchar> got(d_out.step * 40), want(d_out.step * 40, 0);
        REQUIRE(cudaMemcpy(got.data(), d_out.data, got.size(), cudaMemcpyDeviceToHost) == cudaSuccess);
        const cvgs_crop_t crop{h_in.data(), W, H, static_cast<int32_t>(d_in.step), 0};
        cvgs_pipeline_t p{};
        p.src_type = CVGS_8UC4;
        p.dst_width = 60;
        p.dst_height = 40;
        p.aspect_mode = CVGS_IGNORE_AR;
        p.out_layout = CVGS_OUT_NHWC;
        p.dst_type = CVGS_8UC4;
        p.out_row_pitch = static_cast<int64_t>(d_out.step);
        p.out = want.data();
        REQUIRE(oracle_preproc(&crop, 1, 1, &p, 0) == 0);
        for (int y = 0; y < 40; ++y) REQUIRE(std::memcmp(&got[y * d_out.step], &want[y * d_out.step], 60 * 4) == 0);
    }
    return 0;
}

// fkl/tests/algorithm/test_crop.cu:24-45 composed through the cvGS crop() overloads (include/cvGPUSpeedup.cuh:247-265,444):
// read.then(crop(rects)).then(resize(Size(100, 100))) has 100 x 100 x 2 active threads; crop(read, rect) shifts the
// thread by (rect.x, rect.y) and has rect.width x rect.height of them.  Values against the oracle.
static int test_crop() {
    constexpr int W = 128, H = 128;
    std::mt19937 rng(11);
    cv::cuda::GpuMat d_img(H, W, CV_8UC3);
    std::vector<uchar> h_img(d_img.step * H);
    for (auto& b : h_img) b = static_cast<uchar>(rng());
    REQUIRE(cudaMemcpy(d_img.data, h_img.data(), h_img.size(), cudaMemcpyHostToDevice) == cudaSuccess);
    cv::cuda::Stream st;
    const auto readIOp = cvGS::read<CV_8UC3>(d_img);
    const std::array<cv::Rect2d, 2> rects{cv::Rect2d(10, 12, 20, 30), cv::Rect2d(15, 15, 50, 20)};
    auto oracle_of = [&](const cv::Rect2d* rs, int n, cv::Size dst, std::vector<float>& want) {
        std::vector<cvgs_crop_t> hc(n);
        for (int i = 0; i < n; ++i)
            hc[i] = {h_img.data() + static_cast<int>(rs[i].y) * d_img.step + 3 * static_cast<int>(rs[i].x), static_cast<int>(rs[i].width),
                     static_cast<int>(rs[i].height), static_cast<int32_t>(d_img.step), 0};
        cvgs_pipeline_t p{};
        p.src_type = CVGS_8UC3;
        p.dst_width = dst.width;
        p.dst_height = dst.height;
        p.aspect_mode = CVGS_IGNORE_AR;
        p.n_ops = 1;
        p.ops[0].kind = CVGS_OP_MUL;
        for (int c = 0; c < 3; ++c) p.ops[0].v[c] = 0.5f;
        want.assign(static_cast<size_t>(n) * 3 * dst.width * dst.height, -1.f);
        p.out = want.data();
        return oracle_preproc(hc.data(), n, n, &p, 0);
    };
    // 1. batch crop + resize: both spellings of the reference
    for (int spelling = 0; spelling < 2; ++spelling) {
        const cv::Size dst(100, 100);
        cv::cuda::GpuMat d_out(2, dst.width * dst.height * 3, CV_32FC1);
        const cvGS::detail::CroppedRead cropped = spelling == 0 ? readIOp.then(cvGS::crop<2>(rects)) : cvGS::crop(readIOp, rects);
        REQUIRE(cropped.rects.size() == 2 && cropped.rects[1].x == 15 && cropped.rects[1].width == 50);
        cvGS::executeOperations(st, cropped.then(cvGS::resize<cv::INTER_LINEAR>(dst)), cvGS::multiply<CV_32FC3>(cv::Scalar(0.5, 0.5, 0.5)),
                                cvGS::split<CV_32FC3>(d_out, dst));
        st.waitForCompletion();
        std::vector<float> got(static_cast<size_t>(2) * 3 * dst.width * dst.height), want;
        REQUIRE(cudaMemcpy(got.data(), d_out.data, got.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess);
        REQUIRE(oracle_of(rects.data(), 2, dst, want) == 0);
        REQUIRE(std::memcmp(got.data(), want.data(), got.size() * 4) == 0);
    }
    // 2. one crop, no resize: the pixels of the rectangle, bit for bit (times 0.5)
    {
        const cv::Rect2d one(11, 9, 10, 10);
        cv::cuda::GpuMat d_out(1, 10 * 10 * 3, CV_32FC1);
        cvGS::executeOperations(st, cvGS::crop(readIOp, one), cvGS::multiply<CV_32FC3>(cv::Scalar(0.5, 0.5, 0.5)),
                                cvGS::split<CV_32FC3>(d_out, cv::Size(10, 10)));
        st.waitForCompletion();
        std::vector<float> got(3 * 10 * 10);
        REQUIRE(cudaMemcpy(got.data(), d_out.data, got.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess);
        for (int c = 0; c < 3; ++c)
            for (int y = 0; y < 10; ++y)
                for (int x = 0; x < 10; ++x)
                    REQUIRE(got[(c * 10 + y) * 10 + x] == 0.5f * h_img[(9 + y) * d_img.step + 3 * (11 + x) + c]);
    }
    // 3. a rectangle that leaves the image is an error (std::runtime_error, like gpuErrchk)
    bool threw = false;
    try {
        cv::cuda::GpuMat d_out(1, 8 * 8 * 3, CV_32FC1);
        cvGS::executeOperations(st, cvGS::crop(readIOp, cv::Rect2d(125, 0, 8, 8)), cvGS::split<CV_32FC3>(d_out, cv::Size(8, 8)));
    } catch (const std::runtime_error&) { threw = true; }
    REQUIRE(threw);
    return 0;
}

static int test_error_convention() {
    cv::cuda::GpuMat d_input(16, 16, CV_8UC3, cv::Scalar(1, 2, 3));
    cv::cuda::GpuMat d_null;  // data == nullptr
    cv::cuda::Stream st;
    bool thrown = false;
    try {
        cvGS::executeOperations(st, cvGS::resize<CV_8UC3, cv::INTER_LINEAR>(d_input, cv::Size(16, 16)),
                                cvGS::split<CV_32FC3>(d_null, cv::Size(16, 16)));
    } catch (const std::runtime_error& e) {
        thrown = std::strstr(e.what(), "output") != nullptr;
    }
    REQUIRE(thrown);
    return 0;
}

int main() {
    int failed = 0;
    failed += test_batchresize_x_split3D();
    failed += test_circular_tensor<fk::CircularTensorOrder::NewestFirst, fk::ColorPlanes::Standard>();
    failed += test_circular_tensor<fk::CircularTensorOrder::NewestFirst, fk::ColorPlanes::Transposed>();
    failed += test_circular_tensor<fk::CircularTensorOrder::OldestFirst, fk::ColorPlanes::Standard>();
    failed += test_circular_tensor<fk::CircularTensorOrder::OldestFirst, fk::ColorPlanes::Transposed>();
    failed += test_circular_tensor_no_split();
    failed += test_split();
    failed += test_resize_x_split_write();
    failed += test_batchread_x_write3D();
    failed += test_resize_write_8u();
    failed += test_random_vs_oracle();
    failed += test_warping();
    failed += test_cvtcolor_channel_changes();
    failed += test_read_x_write_and_8uc4();
    failed += test_crop();
    failed += test_error_convention();
    std::printf(failed ? "test_shim: %d FAILED\n" : "test_shim: all passed\n", failed);
    return failed ? 1 : 0;
}
