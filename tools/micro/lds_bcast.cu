// Microbenchmark: cost of shared-memory loads whose 32 lanes hit only five distinct, consecutive addresses
// (the access pattern of conv0_lean_bwd_kernel: lane <-> channel, address <-> the channel's winner position).
//   mode 0: three LDS.32 (words a, a+1, a+2),  mode 1: one LDS.128 of a pre-packed (x[a], x[a+1], x[a+2], 0)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bcast lds_bcast.cu && ./lds_bcast
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, int spread) {
    __shared__ __align__(16) float s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = (float)(i % 37) * 0.01f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int j = (lane * 7 + 3) % spread;                       // per-lane "winner" in [0, spread)
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) {
                const float* p = s + u * 48 + j;
                acc0 += p[0]; acc1 += p[1]; acc2 += p[2];
            } else {
                const float4 v = *reinterpret_cast<const float4*>(s + (u * 48 + j) * 4);
                acc0 += v.x; acc1 += v.y; acc2 += v.z;
            }
        }
        j = (j + 1) % spread;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2;
}

template <int MODE>
void run(const char* name, int spread) {
    float* out;
    cudaMalloc(&out, sizeof(float) * 148 * 256);
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 256>>>(out, 10, spread);
    cudaEventRecord(e0);
    k<MODE><<<148, 256>>>(out, iters, spread);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double groups = 8.0 * iters * 16;                 // (3 words) fetched per SM, in warp-groups
    printf("%-34s spread %d: %7.3f ms  %.2f cycles per 3-word group per SM (1.965 GHz)\n", name, spread, ms,
           ms * 1e-3 * 1.965e9 / groups);
    cudaFree(out);
}

int main() {
    for (int spread : {1, 5, 8}) {
        run<0>("3 x LDS.32", spread);
        run<1>("1 x LDS.128 (packed triples)", spread);
    }
    return 0;
}
