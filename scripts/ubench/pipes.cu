// Micro-benchmark: issue rates of FFMA / FFMA2 / PRMT / mixed streams on sm_100a (one CTA per SM, 8..16 warps).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int MODE>
__global__ void k(float* out, float a, float b, unsigned sel) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    float2 y0 = make_float2(x0, x1), y1 = make_float2(x2, x3), y2 = make_float2(x4, x5), y3 = make_float2(x6, x7);
    float2 y4 = y0, y5 = y1, y6 = y2, y7 = y3;
    unsigned u0 = threadIdx.x, u1 = u0 * 3, u2 = u0 * 5, u3 = u0 * 7;
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        if (MODE == 0) {  // 8 independent FFMA
            x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
            x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
        } else if (MODE == 1) {  // 8 independent FFMA2 (scalar-broadcast operands)
            y0 = __ffma2_rn(y0, a2, b2); y1 = __ffma2_rn(y1, a2, b2); y2 = __ffma2_rn(y2, a2, b2); y3 = __ffma2_rn(y3, a2, b2);
            y4 = __ffma2_rn(y4, a2, b2); y5 = __ffma2_rn(y5, a2, b2); y6 = __ffma2_rn(y6, a2, b2); y7 = __ffma2_rn(y7, a2, b2);
        } else if (MODE == 2) {  // 8 FFMA2 with full vector operands
            y0 = __ffma2_rn(y0, y4, y5); y1 = __ffma2_rn(y1, y5, y6); y2 = __ffma2_rn(y2, y6, y7); y3 = __ffma2_rn(y3, y7, y4);
            y0 = __ffma2_rn(y0, y5, y6); y1 = __ffma2_rn(y1, y6, y7); y2 = __ffma2_rn(y2, y7, y4); y3 = __ffma2_rn(y3, y4, y5);
        } else if (MODE == 3) {  // 8 PRMT
            u0 = __byte_perm(u0, u1, sel); u1 = __byte_perm(u1, u2, sel); u2 = __byte_perm(u2, u3, sel); u3 = __byte_perm(u3, u0, sel);
            u0 = __byte_perm(u0, u2, sel); u1 = __byte_perm(u1, u3, sel); u2 = __byte_perm(u2, u0, sel); u3 = __byte_perm(u3, u1, sel);
        } else if (MODE == 4) {  // 4 PRMT + 4 FFMA2 interleaved
            u0 = __byte_perm(u0, u1, sel); y0 = __ffma2_rn(y0, a2, b2); u1 = __byte_perm(u1, u2, sel); y1 = __ffma2_rn(y1, a2, b2);
            u2 = __byte_perm(u2, u3, sel); y2 = __ffma2_rn(y2, a2, b2); u3 = __byte_perm(u3, u0, sel); y3 = __ffma2_rn(y3, a2, b2);
        } else if (MODE == 5) {  // 4 PRMT + 4 FFMA interleaved
            u0 = __byte_perm(u0, u1, sel); x0 = __fmaf_rn(x0, a, b); u1 = __byte_perm(u1, u2, sel); x1 = __fmaf_rn(x1, a, b);
            u2 = __byte_perm(u2, u3, sel); x2 = __fmaf_rn(x2, a, b); u3 = __byte_perm(u3, u0, sel); x3 = __fmaf_rn(x3, a, b);
        } else if (MODE == 6) {  // 8 FFMA, 3 distinct register operands
            x0 = __fmaf_rn(x0, x4, x5); x1 = __fmaf_rn(x1, x5, x6); x2 = __fmaf_rn(x2, x6, x7); x3 = __fmaf_rn(x3, x7, x4);
            x0 = __fmaf_rn(x0, x5, x6); x1 = __fmaf_rn(x1, x6, x7); x2 = __fmaf_rn(x2, x7, x4); x3 = __fmaf_rn(x3, x4, x5);
        } else if (MODE == 7) {  // 4 funnel shifts + 4 PRMT
            u0 = __funnelshift_r(u0, u1, sel); u1 = __byte_perm(u1, u2, sel); u2 = __funnelshift_r(u2, u3, sel); u3 = __byte_perm(u3, u0, sel);
            u0 = __funnelshift_r(u0, u2, sel); u1 = __byte_perm(u1, u3, sel); u2 = __funnelshift_r(u2, u0, sel); u3 = __byte_perm(u3, u1, sel);
        }
    }
    long long t1 = clock64();
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + y0.x + y0.y + y1.x + y1.y + y2.x + y2.y + y3.x + y3.y + y4.x + y5.x + y6.x + y7.x +
              __uint_as_float(u0 ^ u1 ^ u2 ^ u3);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int MODE>
void run(const char* name, int warps, float* d) {
    k<MODE><<<148, warps * 32>>>(d, 1.0001f, 0.5f, 0x5410);
    cudaDeviceSynchronize();
    k<MODE><<<148, warps * 32>>>(d, 1.0001f, 0.5f, 0x5410);
    cudaDeviceSynchronize();
    float cyc;
    cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
    // warp-instructions issued per SMSP per cycle: (warps/4) * 8 * N / cycles
    printf("%-44s warps/SM %2d  cycles %9.0f  inst/clk/SMSP %.3f\n", name, warps, cyc, (warps / 4.0) * 8 * N / cyc);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 1024 * 4);
    for (int w : {4, 8, 16}) {
        run<0>("FFMA (imm/const operands)", w, d);
        run<6>("FFMA (3 register operands)", w, d);
        run<1>("FFMA2 (scalar-broadcast operands)", w, d);
        run<2>("FFMA2 (3 vector register operands)", w, d);
        run<3>("PRMT", w, d);
        run<7>("SHF + PRMT", w, d);
        run<4>("PRMT + FFMA2 interleaved", w, d);
        run<5>("PRMT + FFMA interleaved", w, d);
    }
    return 0;
}
