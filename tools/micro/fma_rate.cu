// Microbenchmark: issue rate of FFMA (3-register), FFMA with an immediate, and packed FFMA2 per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) {            // 2 scalar FFMA (3-reg)
                    acc[i].x = fmaf(acc[i].x, a, b);
                    acc[i].y = fmaf(acc[i].y, a2.y, b2.y);
                } else if (MODE == 1) {     // 1 packed FFMA2
                    acc[i] = __ffma2_rn(acc[i], a2, b2);
                } else {                    // 2 scalar FFMA with immediate multiplier
                    acc[i].x = fmaf(acc[i].x, 1.0001f, b);
                    acc[i].y = fmaf(acc[i].y, 0.9999f, b2.y);
                }
            }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
    int sms = 148;
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 1024);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, 100, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)sms * warps_per_sm * 32 * iters * 4 * 8 * 2;     // scalar FMAs
    printf("%-28s warps/SM %2d: %7.3f ms  %6.2f TFLOP/s  %.1f FMA/clk/SM (at 1.965 GHz)\n", name, warps_per_sm, ms,
           2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / 1.965e9);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("FFMA 3-reg (x2)", w);
        run<1>("FFMA2 packed", w);
        run<2>("FFMA imm (x2)", w);
    }
    return 0;
}
