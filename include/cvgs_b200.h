/*
 * cvgs_b200.h -- C-ABI of the B200-native fused image-preprocessing path.
 *
 * This is the drop-in boundary for ONE hot path of cvGPUSpeedup / FusedKernelLibrary:
 *
 *   batched crop -> bilinear resize -> convertTo/scale -> per-channel mul/sub/div/add
 *                -> channel reorder (cvtColor) -> planar split           (one kernel launch)
 *   CircularTensor shift + process                                       (one kernel launch)
 * and the callers / data formats either side of it that SURVEY.md 8(f) ranks next: YUV frames as sources, batched
 * affine / perspective warps in front of the same chain, the other source depths, channel counts and output forms.
 *
 * The reference has no FFI: its public surface is header-only C++ templates
 * (reference include/cvGPUSpeedup.cuh:74-627) whose operation structs are PODs with public
 * `params` fields.  Every entry point below states which reference interface it replaces;
 * the header shim in cvgpuspeedup_b200/include/cvGPUSpeedup.cuh maps the reference's
 * operation-struct chains onto these calls.
 *
 * Conventions (same as the reference):
 *   - caller's thread, caller's stream; launches are asynchronous, nothing is synchronised
 *     or allocated on the hot call (descriptor staging buffers are created lazily once);
 *   - all image memory is caller-owned device memory (except CircularTensor, which owns
 *     its tensors like fk::CircularTensor does);
 *   - every function returns 0 on success or a non-zero cudaError_t-compatible code; the
 *     message is available from cvgs_b200_last_error() (thread-local).  The reference
 *     throws std::runtime_error from gpuErrchk (fkl/.../core/utils/utils.h:42-60); the
 *     header shim rethrows to keep that behaviour.
 *
 * No torch / OpenCV / CUDA types appear in any signature: streams are `void*`
 * (a cudaStream_t), device pointers are plain pointers.
 */
#ifndef CVGS_B200_H_
#define CVGS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVGS_B200_VERSION 105 /* 0.1.5 */

/* ---- error codes (subset of cudaError_t values so they can be passed through) ---- */
#define CVGS_OK 0
#define CVGS_ERR_INVALID_VALUE 1   /* cudaErrorInvalidValue: bad argument / unsupported chain */
#define CVGS_ERR_NO_DEVICE 100     /* cudaErrorNoDevice */
#define CVGS_ERR_NOT_SUPPORTED 801 /* cudaErrorNotSupported */

/* ---- pixel types: same numeric values as OpenCV's CV_MAKETYPE(depth, cn) ---- */
#define CVGS_8UC3 16  /* CV_8UC3  */
#define CVGS_16UC3 18 /* CV_16UC3: source only (SaturateCast saturate.cuh:267-298); TMA-staged kernel in the common geometry */
#define CVGS_16SC3 19 /* CV_16SC3: source only (saturate.cuh:358-378); TMA-staged kernel in the common geometry          */
#define CVGS_32FC1 5  /* CV_32FC1: output of a chain that ends in one channel (CVGS_OP_GRAY) */
#define CVGS_32FC3 21 /* CV_32FC3 */
#define CVGS_8UC4 24  /* CV_8UC4, CV_16UC4, CV_16SC4: 4-channel sources (the reference's test matrix,            */
#define CVGS_16UC4 26 /* tests/batchresize/test_batchresize_x_split3D.cu:427-432): four output channels, four      */
#define CVGS_16SC4 27 /* constants per operation, REORDER over four channels; all three in the common geometry
                         (IGNORE_AR, all planes used, planar float output) go through the TMA-staged kernel, the rest
                         through the direct-gather kernel                                                              */
#define CVGS_32FC4 29 /* CV_32FC4 */
/* Decoder output (not an OpenCV type code): 8-bit 4:2:0, a Y plane of `height` rows followed by an interleaved UV plane
 * of height/2 rows at data + pitch * height, read as fk::ReadYUV<fk::NV12> and converted per tap by
 * fk::ConvertYUVToRGB<NV12, range, primaries, false, float3> in front of the resize (reference
 * fkl/.../image_processing/color_conversion.cuh:235-362; tests/resize/test_fused_resize.cu:73-76,141-143).  A crop
 * of this type is a whole frame {Y plane, width, height, pitch}; the pipeline sees float RGB.  Batches of even-sized
 * Frames of every format below in the common geometry (IGNORE_AR, every plane used, planar float tensor, even sizes, pitch
 * a multiple of 16; P010 / P210: base a multiple of 4, Y210: of 8) take the TMA-staged kernel (csrc/preproc_yuv_tma.cuh),
 * everything else the direct-gather kernel. */
#define CVGS_NV12 0x1001
/* The other fk::PixelFormat readers the reference can instantiate (color_conversion.cuh:89-98,296-345), same contract:
 *   CVGS_NV21  as NV12 with the chroma bytes in V, U order
 *   CVGS_P010  16-bit words with the 10-bit sample in the high bits (>> 6): Y plane, then an interleaved UV plane of
 *              height/2 rows at data + pitch * height (4:2:0)
 *   CVGS_P210  as P010 with a chroma row per luma row (4:2:2)
 *   CVGS_Y210  packed 4:2:2: {Y0, U, Y1, V} 16-bit words per pixel pair, 10-bit samples in the high bits
 * The 10-bit formats are converted in their own range (chroma - 512, bt601 luma - 64) and the float RGB is multiplied by
 * 64 afterwards (NormalizeColorRangeDepth), i.e. it is 16-bit-scaled like the stored samples. */
#define CVGS_NV21 0x1002
#define CVGS_P010 0x1003
#define CVGS_P210 0x1004
#define CVGS_Y210 0x1005
#define CVGS_IS_YUV(t) ((t) >= CVGS_NV12 && (t) <= CVGS_Y210)

/* YCbCr -> RGB matrices of the reference (color_conversion.cuh:171-214): ccMatrix<range, primaries, YCbCr2RGB>. */
enum cvgs_yuv_standard {
    CVGS_YUV_BT601_FULL = 0,   /* luma offset 16, 1.164 / 1.596 / -0.392 / -0.813 / 2.017 */
    CVGS_YUV_BT709_FULL = 1,
    CVGS_YUV_BT709_LIMITED = 2,
    CVGS_YUV_BT2020_FULL = 3
};

/* Aspect-ratio policy of the resize; same numbering as cvGS::AspectRatio
 * (reference include/cvGPUSpeedup.cuh:32, fkl/.../image_processing/resize.cuh:41). */
enum cvgs_aspect_ratio {
    CVGS_PRESERVE_AR = 0,
    CVGS_IGNORE_AR = 1,
    CVGS_PRESERVE_AR_RN_EVEN = 2,
    CVGS_PRESERVE_AR_LEFT = 3
};

/* Post-resize per-pixel operations, applied in list order (reference
 * fkl/.../basic_ops/arithmetic.cuh:43-68, cuda_vector.cuh:45-54, cvGPUSpeedup.cuh:131-161). */
enum cvgs_op_kind {
    CVGS_OP_MUL = 1,     /* x * v[c]   cvGS::multiply / convertTo(alpha)          */
    CVGS_OP_SUB = 2,     /* x - v[c]   cvGS::subtract                            */
    CVGS_OP_DIV = 3,     /* x / v[c]   cvGS::divide   (IEEE-754 correctly rounded) */
    CVGS_OP_ADD = 4,     /* x + v[c]   cvGS::add / convertTo(alpha, beta)         */
    CVGS_OP_REORDER = 5, /* out[c] = in[perm[c]]   cvGS::cvtColor<RGB2BGR/BGR2RGB> = {2,1,0} */
    /* colour conversions that change the channel count (reference color_conversion.cuh:364-461; the codes that also
     * swap R and B are a REORDER followed by one of these).  Direct-gather kernel; the ops after them take the new
     * channel count, and the output has it (tensor layouts; dst_type 0 or the matching CVGS_32FCn). */
    CVGS_OP_ADD_ALPHA = 6,  /* 3 -> 4 channels, alpha = v[0]   cvtColor<BGR2BGRA / BGR2RGBA> (the wrapper passes 255:
                               AddOpaqueAlpha<float3, p8bit>, color_conversion.cuh:122-130)            */
    CVGS_OP_DROP_ALPHA = 7, /* 4 -> 3 channels                 cvtColor<BGRA2BGR / RGBA2BGR> (fk::Discard)            */
    CVGS_OP_GRAY = 8        /* 3 or 4 -> 1 channel, 0.299 R + 0.587 G + 0.114 B of channels (0, 1, 2) in that order
                               cvtColor<RGB2GRAY / RGBA2GRAY>; BGR2GRAY = REORDER {2,1,0} + GRAY (RGB2Gray<I, float>,
                               color_conversion.cuh:42-68; as compiled: FMUL, FFMA, FFMA; the reference then rounds
                               the luminance to the nearest integer even for float output -- std::is_signed_v<float> is
                               true, :55-60 -- and so does this op).  perm[0] = which product nvcc left as
                               the stand-alone FMUL in that instantiation: 1 (y * 0.587) for RGB2GRAY / RGBA2GRAY, 0
                               (x * 0.299) for BGR2GRAY / BGRA2GRAY, whose reorder is fused in front                  */
};

/* Floating-point contract:
 *   CVGS_FP_REFERENCE_FUSED  bit-identical to the reference's fused kernel as nvcc compiles
 *                            it: a MUL directly followed by ADD/SUB is one FMA.
 *   CVGS_FP_SEPARATE         every op rounded on its own, i.e. the result of running the
 *                            OpenCV-CUDA multiply/subtract/divide kernels one after another.
 * The bilinear interpolation itself is the same in both (1 FMUL + 3 FFMA, see oracle/). */
enum cvgs_fp_contract { CVGS_FP_REFERENCE_FUSED = 0, CVGS_FP_SEPARATE = 1 };

/* Resize output handed to the op chain:
 *   CVGS_INTERP_FLOAT     interpolated value stays float (what fk::Interpolate returns).
 *   CVGS_INTERP_ROUND_U8  value is rounded (RN-even) and saturated to the range of the SOURCE depth first
 *                         ([0,255], [0,65535] or [-32768,32767]), i.e. cv::cuda::resize on the integer image
 *                         followed by convertTo(CV_32F). */
enum cvgs_interp_mode { CVGS_INTERP_FLOAT = 0, CVGS_INTERP_ROUND_U8 = 1 };

/* Output layouts (reference fkl/.../memory_operations.cuh:168-220, ptr_nd.cuh:53-77). */
enum cvgs_out_layout {
    CVGS_OUT_NCHW = 0, /* fk::TensorSplit  : out[z][c][y][x]; cvGS::split(GpuMat, Size)   */
    CVGS_OUT_CNHW = 1, /* fk::TensorTSplit : out[c][z][y][x]; cvGS::splitT(RawPtr<T3D>)   */
    CVGS_OUT_NHWC = 2, /* fk::PerThreadWrite<_3D,float3>: packed; cvGS::write(GpuMat,Size) */
    CVGS_OUT_PLANES = 3 /* fk::SplitWrite: one 2-D float image per (crop, channel), each with its own pointer and
                           pitch; cvGS::split(vector<GpuMat>) / split(array<vector<GpuMat>,N>)
                           (reference memory_operations.cuh:331-360, cvGPUSpeedup.cuh:163-183).  `out` is then a
                           HOST pointer to n_planes * 3 cvgs_plane_t, crop-major: [z][c]. */
};

/* One destination image of CVGS_OUT_PLANES = fk::RawPtr<fk::_2D, float> built from a CV_32FC1 GpuMat. */
typedef struct cvgs_plane {
    void* data;          /* device pointer, float                                  */
    int64_t pitch_bytes; /* bytes between rows (GpuMat::step), a multiple of 4     */
} cvgs_plane_t;

/* One source crop = fk::RawPtr<fk::_2D, T> {data, {width, height, pitch}}
 * (reference fkl/.../core/data/ptr_nd.h:24-60; built by cvGS::gpuMat2RawPtr2D,
 * include/cvGPUSpeedup.cuh:40-44 from GpuMat::data/cols/rows/step). */
typedef struct cvgs_crop {
    const void* data; /* device pointer to the first pixel of the ROI            */
    int32_t width;    /* pixels                                                   */
    int32_t height;   /* pixels                                                   */
    int32_t pitch;    /* bytes between consecutive rows (GpuMat::step)            */
    int32_t reserved; /* must be 0                                                */
} cvgs_crop_t;

typedef struct cvgs_op {
    int32_t kind;    /* enum cvgs_op_kind                                         */
    int32_t perm[4]; /* CVGS_OP_REORDER only                                      */
    float v[4];      /* per-channel constant (cv::Scalar cast to float per channel,
                        reference include/cvGPUSpeedupHelpers.cuh:38-54)          */
} cvgs_op_t;

#define CVGS_MAX_OPS 8

/* Everything cvGS::executeOperations(stream, resize(...), ops..., split(...)) carries besides
 * the crops (reference include/cvGPUSpeedup.cuh:218-245 resize, :131-161 ops, :185-202 split). */
typedef struct cvgs_pipeline {
    int32_t src_type;    /* CVGS_8UC3, CVGS_16UC3, CVGS_16SC3 or their 4-channel forms */
    int32_t dst_width;   /* cv::Size dsize of cvGS::resize                         */
    int32_t dst_height;
    int32_t aspect_mode; /* enum cvgs_aspect_ratio                                 */
    int32_t interp_mode; /* enum cvgs_interp_mode                                  */
    int32_t fp_contract; /* enum cvgs_fp_contract                                  */
    float background[4]; /* backgroundValue of cvGS::resize: value of inactive planes
                            (z >= used) and of the bands outside the AR-preserved image;
                            it still flows through the op chain     */
    int32_t n_ops;
    cvgs_op_t ops[CVGS_MAX_OPS];
    int32_t out_layout; /* enum cvgs_out_layout                                    */
    int32_t dst_type;   /* 0 (float, channels as the chain leaves them) or the matching CVGS_32FC1 / CVGS_32FC3 /
                           CVGS_32FC4: float output.  CVGS_8UC3 / CVGS_8UC4 (3- / 4-channel sources): the chain ends with
                           convertTo<CV_32FCn, CV_8UCn> (SaturateCast<float, uchar>: round to nearest even, clamp to
                           [0, 255], reference saturate.cuh:127-147) and packed 8-bit pixels are written --
                           cvGS::write<CV_8UCn>(GpuMat), the form of the reference's tests/resize/test_resize_write.cu
                           and test_resize_CPUvsGPUresults.cu; needs CVGS_OUT_NHWC. */
    void* out;                 /* device pointer, float (CVGS_OUT_PLANES: host array of cvgs_plane_t) */
    int64_t out_plane_stride;  /* floats between consecutive batch planes z; 0 = tight
                                  (3*dst_width*dst_height for NCHW/NHWC, dst_width*dst_height
                                  for CNHW).  The reference ignores GpuMat::step.
                                  For 8-bit output the unit is bytes. */
    int64_t out_row_pitch;     /* CVGS_OUT_NHWC only: BYTES between rows of a packed destination image (GpuMat::step of
                                  cvGS::write<O>(GpuMat) / executeOperations(input, output, ...)); 0 = tight.  Float
                                  images: a multiple of 4; out_plane_stride then defaults to rows * pitch. */
    int32_t yuv_standard;      /* YUV sources (CVGS_NV12 ... CVGS_Y210) only: enum cvgs_yuv_standard */
    int32_t u8_cast;           /* 8-bit output only: 0 = SaturateCast (round to nearest even, clamp; convertTo),
                                  1 = fk::Cast<float3, uchar3> (C++ static_cast: truncation; values must lie in [0, 256),
                                  reference basic_ops/cast.cuh:22-29, as in tests/warping/test_warping_opencv.cu:63) */
} cvgs_pipeline_t;

/* ------------------------------------------------------------------------------------------
 * Library
 * ------------------------------------------------------------------------------------------ */
int cvgs_b200_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* cvgs_b200_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Fused batch pipeline.  Replaces
 *   cvGS::executeOperations(stream, cvGS::resize<CV_8UC3, INTER_LINEAR, N, AR>(crops, dsize,
 *       used, bg), [cvtColor], [multiply], [subtract], [divide], [add], cvGS::split<CV_32FC3>(...))
 * (reference include/cvGPUSpeedup.cuh:464-473 -> fkl/.../fused_kernel.cuh:22-38 ->
 *  executors.cuh:123-158 -> data_parallel_patterns.cuh:157-197,256-260).
 *
 *   crops       host array of n_planes descriptors (only the first `used` are read)
 *   n_planes    batch size N of the BatchRead = number of output planes written
 *   used        usedPlanes: planes z >= used are filled with chain(background)
 *   stream      cudaStream_t (NULL = legacy default stream)
 * ------------------------------------------------------------------------------------------ */
int cvgs_b200_preproc_launch(const cvgs_crop_t* crops, int32_t n_planes, int32_t used,
                             const cvgs_pipeline_t* pipeline, void* stream);

/* The same launch with the parent image of every crop named.  A cv::cuda::GpuMat ROI remembers the image it was cut
 * from (GpuMat::datastart and locateROI(wholeSize, ofs)); the reference has no use for that and passes only the ROI
 * (include/cvGPUSpeedup.cuh:54-65).  Here it lets the library stage every crop of an image through ONE cached
 * tensor map per image (the parent image is memory known to be readable, so the staging boxes may overhang a crop)
 * instead of encoding one tensor map per crop and launch: the host cost of a 50-crop launch drops from ~6.5 us to
 * ~3 us.  parents[i] describes crops[i]; a NULL `parents`, a NULL datastart or a crop that does not lie inside its
 * parent makes the call behave exactly like cvgs_b200_preproc_launch.  Results are identical either way. */
typedef struct cvgs_parent {
    const void* datastart; /* first byte of the parent image (GpuMat::datastart)            */
    int32_t whole_width;   /* parent size in pixels (locateROI wholeSize); its row pitch is  */
    int32_t whole_height;  /* the crop's pitch                                               */
} cvgs_parent_t;
int cvgs_b200_preproc_launch_ex(const cvgs_crop_t* crops, const cvgs_parent_t* parents, int32_t n_planes,
                                int32_t used, const cvgs_pipeline_t* pipeline, void* stream);

/* The same launch with the crops given as rectangles of ONE device frame: what
 *   cvGS::crop(readOfFrame, rects) / readOfFrame.then(cvGS::crop<BATCH>(rects))   [.then(Resize)]
 * builds (reference include/cvGPUSpeedup.cuh:247-265,444 -> fk::Crop<BackIOp>, fkl/.../image_processing/crop.cuh:23-55:
 * thread (x, y) of plane i reads pixel (x + rect.x, y + rect.y), rect.width x rect.height of them).  Rectangles must lie
 * inside the frame; the frame is named as every crop's parent, so the launch takes the cached per-image tensor maps.
 * Without a resize the caller passes dst size == rect size (all rectangles equal), which reads the pixels bit for bit. */
typedef struct cvgs_rect { int32_t x, y, width, height; } cvgs_rect_t;
int cvgs_b200_preproc_launch_rects(const void* frame, int32_t frame_width, int32_t frame_height, int32_t frame_pitch,
                                   const cvgs_rect_t* rects, int32_t n_planes, int32_t used,
                                   const cvgs_pipeline_t* pipeline, void* stream);

/* The same launch writing its tensor more than once: at pipeline->out and at each of replicas[0 .. n_replicas) (at most
 * 7), all with the layout the pipeline describes.  Meant for BASELINE config 5 -- crops sharded over the GPUs of one box,
 * every GPU ending with the whole [N][3][H][W] tensor: replicas are the other GPUs' tensors mapped into this process
 * (cvgs_b200_ipc_open), so the kernel's own stores travel over NVLink while it computes and no gather collective follows.
 * The reference has no multi-GPU path (fkl/.../execution_model/parallel_architectures.h:20-30 only names architectures);
 * what makes the split legal is that batch planes are independent (batch_operations.cuh:222-229).
 * Common geometry only (CV_8UC3, IGNORE_AR, used == n_planes, NCHW / CNHW float); other forms return
 * CVGS_ERR_NOT_SUPPORTED.  Ordering between GPUs is the caller's (a barrier / stream-ordered collective after the launch). */
int cvgs_b200_preproc_launch_replicated(const cvgs_crop_t* crops, const cvgs_parent_t* parents, int32_t n_planes,
                                        int32_t used, const cvgs_pipeline_t* pipeline, void* const* replicas,
                                        int32_t n_replicas, void* stream);
/* Device memory that can be mapped into the other processes of the box (one process per GPU): cudaMalloc / cudaFree and
 * cudaIpcGetMemHandle / cudaIpcOpenMemHandle (lazy peer access) / cudaIpcCloseMemHandle behind plain pointers.
 * handle64 = 64 bytes, to be exchanged between the processes by any means. */
int cvgs_b200_dev_alloc(void** ptr, uint64_t bytes);
int cvgs_b200_dev_free(void* ptr);
int cvgs_b200_ipc_export(void* ptr, void* handle64);
int cvgs_b200_ipc_open(const void* handle64, void** ptr);
int cvgs_b200_ipc_close(void* ptr);

/* ------------------------------------------------------------------------------------------
 * Batched affine / perspective warp in front of the same chain.  Replaces
 *   cvGS::executeOperations(stream, cvGS::warp<WT, InputType[, N]>(images, matrices, dstSize[, used, default]), ops..., write)
 * (reference include/cvGPUSpeedup.cuh:285-442 -> fk::Warping<WT, PerThreadRead> fkl/.../image_processing/warping.cuh:
 * 43-91, which shares fk::Interpolate<INTER_LINEAR> with the resize).  For plane z and destination pixel (x, y):
 *   affine       sx = (m00*x + m01*y) + m02,  sy = (m10*x + m11*y) + m12
 *   perspective  c = 1 / ((m20*x + m21*y) + m22),  sx = c * ((m00*x + m01*y) + m02),  sy = c * (...)
 * with m the INVERSE transform (destination -> source; the cvGS wrapper inverts the user's matrix on the host and
 * casts it to float); inside [0, w) x [0, h) the pixel is the bilinear interpolation at (sx, sy), outside it is 0;
 * the op chain and the output forms are those of the resize pipeline (pipeline->aspect_mode is ignored).
 * Sources: CV_8U / CV_16U / CV_16S with 3 or 4 channels (pipeline->src_type); gather kernel.
 * ------------------------------------------------------------------------------------------ */
enum cvgs_warp_type { CVGS_WARP_AFFINE = 0, CVGS_WARP_PERSPECTIVE = 1 };
typedef struct cvgs_warp {
    int32_t type; /* enum cvgs_warp_type                                              */
    float m[9];   /* inverse transform, row-major; affine uses m[0..5]                */
} cvgs_warp_t;
int cvgs_b200_warp_launch(const cvgs_crop_t* images, const cvgs_warp_t* warps, int32_t n_planes, int32_t used,
                          const cvgs_pipeline_t* pipeline, void* stream);

/* Same pipeline with HOST buffers, for callers that hold frames in (pinned) host memory:
 * brings the part of the source image its crops touch to the device, launches, copies the tensor back, all on
 * `stream` and without synchronising.  Crops are rectangles {x, y, w, h} of the one host image.
 * `pipeline->out` is ignored; the result goes to host_out (tight layout; channels as the chain leaves them).
 * Device staging buffers are owned by the library and reused across calls.  The frame is copied with
 * cudaMemcpy2DAsync, whole rows from the first to the last row a crop touches (see cvgs_b200_set_host_upload). */
int cvgs_b200_preproc_host(const void* host_image, int32_t image_width, int32_t image_height,
                           int32_t image_pitch, const cvgs_rect_t* rects, int32_t n_planes,
                           int32_t used, const cvgs_pipeline_t* pipeline, float* host_out,
                           void* stream);

/* Frame loops in native code: `steps` consecutive calls of cvgs_b200_preproc_launch /
 * cvgs_b200_preproc_host, call i using argument set (i % n_sets).  This is what a C++ caller's
 * per-frame loop does (the reference's benchmarks time exactly such loops,
 * tests/testsCommon.cuh:260-308); it exists so that language bindings with a slow call path
 * (ctypes, JNI) can drive and time many frames without their per-call overhead.
 * Stops at the first error and returns it.
 * cvgs_b200_preproc_launch_sequence_ex with cvgs_b200_set_overlap(1): when the argument sets are provably
 * independent (outputs pairwise disjoint and disjoint from every source) and steps >= 16, the loop is driven by
 * several host threads (CVGS_B200_SEQ_THREADS, default 3, at most 4), each launching the sets it owns into its own
 * stream; the caller's stream is forked into and joined from those streams, so for the caller the sequence is still
 * one ordered operation and every result is identical.  One host thread sustains ~3.9 us per 50-crop frame, three
 * ~2.1 us (GPU-bound). */
int cvgs_b200_preproc_launch_sequence(const cvgs_crop_t* const* crops, const int32_t* n_planes,
                                      const int32_t* used, const cvgs_pipeline_t* const* pipelines,
                                      int32_t n_sets, int32_t steps, void* stream);
int cvgs_b200_preproc_launch_sequence_ex(const cvgs_crop_t* const* crops, const cvgs_parent_t* const* parents,
                                         const int32_t* n_planes, const int32_t* used,
                                         const cvgs_pipeline_t* const* pipelines, int32_t n_sets,
                                         int32_t steps, void* stream);
int cvgs_b200_preproc_host_sequence(const void* const* host_images, int32_t image_width,
                                    int32_t image_height, int32_t image_pitch,
                                    const cvgs_rect_t* const* rects, const int32_t* n_planes,
                                    const int32_t* used, const cvgs_pipeline_t* const* pipelines,
                                    float* const* host_outs, int32_t n_sets, int32_t steps,
                                    void* stream);

/* Kernel-selection override, for tests and profiling: 0 = automatic, 1 = direct-gather kernel (warp: the general kernel),
 * 2 = TMA-staged kernel.  Returns the previous value. */
int cvgs_b200_set_kernel_variant(int variant);
/* Overlap of consecutive launches (default 0 = off; the environment variable CVGS_B200_OVERLAP=1 / 2 selects a mode at
 * load time).  The fused kernel is always launched with programmatic stream serialisation and, by default, waits for the
 * preceding kernel of the stream before its first global-memory access: plain stream order.
 *   mode 1  inside the library's own frame loops (cvgs_b200_preproc_launch_sequence_ex), where the library controls the
 *           whole stream segment: independent frames share launches (cvgs_b200_set_coalesce) or are driven from several
 *           host threads, and consecutive launches drop the early wait when the library proves them independent (their
 *           sources and outputs disjoint from the outputs -- and their outputs from the sources -- of the launches
 *           possibly still in flight).  Individual launches keep plain stream order.
 *   mode 2  additionally between individual launches of one stream (cvgs_b200_preproc_launch[_ex]), tracked per stream
 *           over the library's own last 8 launches.  The CALLER asserts that nothing it enqueues between two launches of
 *           the library produces a source image of the later one: the library cannot see foreign kernels, and a kernel
 *           that is ordered only by the implicit completion trigger of its predecessor is not guaranteed to see its writes.
 * Every kernel still waits for its predecessor before it completes, so later stream operations observe the usual order.
 * The reference has no equivalent (it launches plain kernels, executors.cuh:133-156).  Returns the previous mode. */
int cvgs_b200_set_overlap(int mode);
/* Frame loops (cvgs_b200_preproc_launch_sequence_ex) with overlap enabled: when the argument sets are provably
 * independent, carry the same pipeline apart from the output pointer, name their parent frames and have the common
 * geometry (CV_8UC3 sources, IGNORE_AR, every plane used, NCHW float output), consecutive steps SHARE kernel launches:
 * up to 928 crops of up to 32 argument sets per launch, each crop writing into its own set's tensor, one host thread,
 * the caller's stream.  A 50-crop frame is 7.5 MB of traffic (~1.5 us of HBM time), less than one kernel launch costs
 * the host and the GPU front end; ten frames per launch are bound by the GPU instead.  Results are identical.  Default
 * 1 (on; CVGS_B200_SEQ_COALESCE=0 turns it off at load time); 0 = one launch per step as before, driven by several
 * host threads.  The reference would express the same thing as one BatchRead over all the crops
 * (batch_operations.cuh:222-229), which its template batch size caps at 255 planes.  Returns the previous value. */
int cvgs_b200_set_coalesce(int enable);
/* Upload strategy of the host-buffer entry points: 0 (default) = cudaMemcpy2DAsync of the rows the crops span; 1 = for
 * pinned frames (base and pitch multiples of 16 bytes) a kernel reads the 128-byte x 16-row tiles some crop overlaps
 * from host memory in place (fewer bytes, but 128-byte PCIe requests: slower than the copy engine on PCIe hosts, see
 * DESIGN.md 6).  CVGS_B200_HOST_TILES=1 selects 1 at load time.  Returns the previous value. */
int cvgs_b200_set_host_upload(int mode);
/* Number of kernel launches issued by this library on the calling thread so far. */
int64_t cvgs_b200_launch_count(void);
/* Diagnostics (no device needed): how the fast warp kernel would form the source coordinate for this inverse matrix and
 * destination size -- 0 affine, 1 perspective with every denominator proven finite, of one sign and of magnitude in
 * [2^-100, 2^100] (reciprocal without the range check), 2 perspective in general; -1 on bad arguments. */
int cvgs_b200_debug_warp_mode(const float* m9, int32_t warp_type, int32_t dst_width, int32_t dst_height);
/* Diagnostics (no device needed): the chain kind the TMA-staged kernel would run this pipeline's chain with -- 0 interpreter,
 * 1 fma + two-operation division, 2 gray, 3 / 4 the former two plus a constant alpha plane (whose value goes to *alpha) --
 * or the negated error code. */
int cvgs_b200_debug_chain_kind(const cvgs_pipeline_t* pipeline, float* alpha);
/* Diagnostics (no device needed): n / d as the kernels' prologues compute it (multiply-shift by a launch constant,
 * FastDiv in csrc/preproc_tma.cuh); exact for n < 2^31. */
uint32_t cvgs_b200_debug_fast_div(uint32_t n, uint32_t d);
/* Diagnostics: bytes the host-buffer entry points moved host -> device and device -> host on the calling thread so far
 * (uploads count whole tiles / rows as issued). */
int cvgs_b200_debug_host_bytes(uint64_t* h2d, uint64_t* d2h, int reset);
/* Diagnostics: host-side cost of the small-batch TMA launch path on the calling thread, accumulated in
 * microseconds: out5 = {calls, descriptor fill, planning, tensor-map encoding, kernel launch}. */
int cvgs_b200_debug_host_profile(double* out5, int reset);
/* Diagnostics (no device needed): the launch plan the TMA-staged kernel would get for this batch geometry on a GPU
 * with sm_count SMs.  out12 = {plan exists, 32-column groups per band, row pairs per plane, bands per plane, work
 * items, bytes per staging slot, slots per warp, CTAs per SM, grid, staged row bytes needed, items covered by the
 * per-warp ranges, 1 if those ranges tile [0, items) in order}.  crops need valid sizes and pitches only. */
/* Diagnostics (no device needed): the normalised program the kernels run for pipeline->ops (a MUL
 * directly followed by ADD / SUB becomes one FMA, reorders become store offsets, AddOpaqueAlpha is hoisted to the front).
 * out[0..7] = {ops, channels of the output pixel, registers per pixel, 1 if the chain changes the channel count or
 * uses SET / GRAY, dst_chan[0..3]}; out[8 + 9 * i .. 8 + 9 * i + 8] = {kind, a[0..3], b[0..3]} of op i as floats
 * (kind: 1 MUL, 2 ADD, 3 DIV, 4 FMA, 5 SET, low byte 6 GRAY). */
int cvgs_b200_debug_program(const cvgs_pipeline_t* pipeline, float* out80);
int cvgs_b200_debug_plan(const cvgs_crop_t* crops, int32_t n_planes, int32_t used, const cvgs_pipeline_t* pipeline,
                         int32_t sm_count, int32_t image_mode, int32_t items_per_warp, int64_t* out12);
/* Diagnostics (no device needed): would a launch with these output / source byte ranges on the stream identified by
 * stream_key have to wait for its predecessor (1) or may it overlap (0)?  Updates the bookkeeping like a launch. */
int cvgs_b200_debug_overlap_query(void* stream_key, uint64_t out_lo, uint64_t out_hi, uint64_t src_lo, uint64_t src_hi);
/* Diagnostics: on the current device, compare the kernels' two-operation division by a launch constant with IEEE
 * division for every float x with 2^-73 <= |x| < 2^48 and the divisor d; *mismatches = differing results,
 * *reciprocal_used = RN(1/d), or 0 when the host-side proof rejects d (the kernels then use IEEE division). */
int cvgs_b200_debug_division_sweep(float d, unsigned long long* mismatches, unsigned* first_bad_bits,
                                   float* reciprocal_used);

/* ------------------------------------------------------------------------------------------
 * CircularTensor.  Replaces cvGS::CircularTensor<I, O, COLOR_PLANES, BATCH, ORDER, MODE>
 * (reference include/cvGPUSpeedup.cuh:600-627 over fkl/.../core/data/circular_tensor.cuh:84-151).
 * data() is a dense, time-ordered float tensor [BATCH][COLOR_PLANES][H][W] (Standard) or
 * [COLOR_PLANES][BATCH][H][W] (Transposed) that is valid after every update.
 * ------------------------------------------------------------------------------------------ */
enum cvgs_ct_order { CVGS_CT_NEWEST_FIRST = 0, CVGS_CT_OLDEST_FIRST = 1 }; /* fk::CircularTensorOrder */
enum cvgs_ct_planes { CVGS_CT_STANDARD = 0, CVGS_CT_TRANSPOSED = 1 };      /* fk::ColorPlanes        */

/* ctor / Alloc(width, height, deviceID) */
int cvgs_b200_ct_create(void** handle, int32_t width, int32_t height, int32_t color_planes,
                        int32_t batch, int32_t order, int32_t plane_mode, int32_t device);
/* The other instantiations of the reference (include/cvGPUSpeedup.cuh:600-627; tests/batchread/
 * test_circularbatchread_x_write3D.cu:400-460): color_planes 1, 3 or 4 planes of floats (elem_channels 1: TensorSplit /
 * TensorTSplit of a 1-, 3- or 4-channel pixel), or ONE plane of packed elem_channels = 3 / 4 float pixels (TensorWrite of
 * float3 / float4: cvGS::CircularTensor<CV_8UC4, CV_32FC4, 1, ...>).  The update's chain must end with color_planes *
 * elem_channels channels.  cvgs_b200_ct_create is the elem_channels = 1 form. */
int cvgs_b200_ct_create_ex(void** handle, int32_t width, int32_t height, int32_t color_planes, int32_t elem_channels,
                           int32_t batch, int32_t order, int32_t plane_mode, int32_t device);
/* update(stream, frame, ops..., write): runs `pipeline` (its out/out_layout/out_plane_stride
 * fields are ignored, dst size must equal the tensor plane size) on the new frame, stores it
 * as the newest plane and shifts the other BATCH-1 planes by one position. */
int cvgs_b200_ct_update(void* handle, const cvgs_crop_t* frame, const cvgs_pipeline_t* pipeline,
                        void* stream);
/* data(): device pointer of the dense tensor (stable for the lifetime of the handle). */
void* cvgs_b200_ct_data(void* handle);
int cvgs_b200_ct_destroy(void* handle);

#ifdef __cplusplus
}
#endif
#endif /* CVGS_B200_H_ */
