// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float *out, float a, float b, int iters) {
    float2 x[8];
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {
                x[i].x = fmaf(x[i].x, a, b);
                x[i].y = fmaf(x[i].y, a, b);
            } else {
                x[i] = __ffma2_rn(x[i], A, B);
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
            else k<1><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 8 * 256 * 16.0 * iters;
            if (rep) printf("%s: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s)\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
        }
    }
    return 0;
}
