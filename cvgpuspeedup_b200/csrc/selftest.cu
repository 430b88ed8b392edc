// selftest.cu -- device-side self checks exported for the test-suite (diagnostics, not on the hot path).
//
// cvgs_b200_debug_division_sweep: compares the two-operation division by a launch constant of the fused kernels
// (div_by_const2, preproc_tma.cuh; constants and host-side proof in div_const.cpp) with the IEEE routine
// (__fdiv_rn) for EVERY float x with 2^-73 <= |x| < 2^48 (both signs, 2.03e9 values) and one divisor d, on the
// GPU.  Returns the number of mismatching bit patterns; *reciprocal_used = 0 when the host rejects d for the
// fast path (the kernels then divide with __fdiv_rn and there is nothing to compare).
#include <cuda_runtime.h>

#include "preproc_tma.cuh"

namespace cvgs {

__global__ void division_sweep_kernel(float d, float zh, float zl, unsigned long long* mismatches, unsigned* first_bad) {
    constexpr unsigned kExpLo = 127 - 73, kExps = 121;
    const unsigned long long total = 2ull * kExps * (1ull << 23);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const unsigned mant = (unsigned)(i & 0x7fffffu);
        const unsigned k = (unsigned)(i >> 23);
        const unsigned bits = ((k / kExps) << 31) | ((kExpLo + k % kExps) << 23) | mant;
        const float x = __uint_as_float(bits);
        const float fast = div_by_const2(make_float2(x, x), zh, zl).y;
        const float want = __fdiv_rn(x, d);
        if (__float_as_uint(fast) != __float_as_uint(want)) {
            if (!bad) atomicCAS(first_bad, 0u, bits);
            ++bad;
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace cvgs

extern "C" int cvgs_b200_debug_division_sweep(float d, unsigned long long* mismatches, unsigned* first_bad_bits,
                                              float* reciprocal_used) {
    using namespace cvgs;
    if (!mismatches) return fail(CVGS_ERR_INVALID_VALUE, "mismatches is NULL");
    unsigned long long* d_cnt = nullptr;
    unsigned* d_first = nullptr;
    CVGS_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_cnt), sizeof *d_cnt));
    CVGS_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_first), sizeof *d_first));
    CVGS_CUDA(cudaMemset(d_cnt, 0, sizeof *d_cnt));
    CVGS_CUDA(cudaMemset(d_first, 0, sizeof *d_first));
    const DivConst dc = div_const_prepare(d);
    if (reciprocal_used) *reciprocal_used = dc.exact ? dc.zh : 0.f;
    if (!dc.exact) {
        *mismatches = 0;
        if (first_bad_bits) *first_bad_bits = 0;
        cudaFree(d_cnt);
        cudaFree(d_first);
        return CVGS_OK;
    }
    division_sweep_kernel<<<148 * 16, 256>>>(d, dc.zh, dc.zl, d_cnt, d_first);
    CVGS_CUDA(cudaGetLastError());
    unsigned first = 0;
    CVGS_CUDA(cudaMemcpy(mismatches, d_cnt, sizeof *d_cnt, cudaMemcpyDeviceToHost));
    CVGS_CUDA(cudaMemcpy(&first, d_first, sizeof first, cudaMemcpyDeviceToHost));
    if (first_bad_bits) *first_bad_bits = first;
    cudaFree(d_cnt);
    cudaFree(d_first);
    return CVGS_OK;
}
