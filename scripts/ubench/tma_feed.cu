// Micro-benchmark: sustained TMA (cp.async.bulk.tensor.2d) feed rate from an L2-resident image into shared memory
// as a function of box shape and pipeline depth.  One producer thread per CTA, `stages` mbarriers, each stage is
// filled by `boxes` TMA ops of (rows x row_bytes); the stage is re-issued as soon as it completes.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void feed(const __grid_constant__ CUtensorMap map, int stages, int boxes, int rows, int row_bytes, int iters,
                     int img_rows, int img_bytes, unsigned long long* cycles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int stage_bytes = boxes * rows * row_bytes;
    unsigned rng = blockIdx.x * 2654435761u + 12345u;
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int it = 0; it < iters + stages; ++it) {
        const int s = it % stages;
        if (it >= stages) {  // wait for the previous fill of this stage
            uint32_t ok = 0;
            while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(phase) : "memory");
        }
        if (s == stages - 1 && it >= stages) phase ^= 1u;
        if (it < iters) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(stage_bytes) : "memory");
            for (int b = 0; b < boxes; ++b) {
                rng = rng * 1664525u + 1013904223u;
                const int y = (rng >> 8) % (img_rows - rows);
                const int x = ((rng >> 3) % ((img_bytes - row_bytes) / 16)) * 2;  // 16-byte aligned start, in 8-byte elements
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                                 smem_u32(smem) + s * stage_bytes + b * rows * row_bytes),
                             "l"(&map), "r"(x), "r"(y), "r"(smem_u32(&bar[s]))
                             : "memory");
            }
        }
    }
    if (blockIdx.x == 0) *cycles = clock64() - t0;
}
int main() {
    const int W = 3840 * 3, H = 2160, pitch = 11776;  // 4K BGR frame, 24.9 MB: L2 resident
    uint8_t* img; cudaMalloc(&img, (size_t)pitch * H); cudaMemset(img, 1, (size_t)pitch * H);
    unsigned long long* d_cyc; cudaMalloc(&d_cyc, 8);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto encode = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    cudaFuncSetAttribute(feed, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { int rows, row_bytes, boxes, stages, ctas_per_sm; };
    const Cfg cfgs[] = {{2, 1664, 8, 2, 4},  {2, 1664, 8, 1, 4}, {16, 1664, 1, 2, 4}, {4, 1664, 4, 2, 4}, {2, 832, 8, 4, 4}, {2, 832, 16, 2, 4},
                        {16, 832, 1, 4, 4}, {2, 1664, 8, 4, 2}, {2, 1664, 4, 4, 4}, {2, 1664, 2, 8, 4}, {2, 2048, 8, 1, 4}, {8, 2048, 2, 1, 4}, {2, 256, 32, 2, 4}, {2, 256, 32, 4, 4}};
    for (const Cfg& c : cfgs) {
        CUtensorMap map;
        cuuint64_t dim[2] = {(cuuint64_t)W / 8, (cuuint64_t)H}; cuuint64_t stride[1] = {(cuuint64_t)pitch};
        cuuint32_t box[2] = {(cuuint32_t)c.row_bytes / 8, (cuuint32_t)c.rows}; cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, img, dim, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int iters = 200, grid = 148 * c.ctas_per_sm;
        const size_t smem = (size_t)c.stages * c.boxes * c.rows * c.row_bytes;
        for (int rep = 0; rep < 2; ++rep) {
            feed<<<grid, 32, smem>>>(map, c.stages, c.boxes, c.rows, c.row_bytes, iters, H, W, d_cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        unsigned long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
        const double bytes_per_sm = (double)iters * c.boxes * c.rows * c.row_bytes * c.ctas_per_sm;
        printf("box %2d rows x %4d B, %2d boxes/stage, %d stages, %d CTAs/SM (smem %6zu B/CTA): %9llu cycles, %6.1f B/clk/SM, %5.0f cycles per stage fill\n",
               c.rows, c.row_bytes, c.boxes, c.stages, c.ctas_per_sm, smem, cyc, bytes_per_sm / cyc, (double)cyc / iters * c.stages);
    }
    return 0;
}
