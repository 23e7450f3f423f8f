// Microbenchmark: issue rate of FP32 FMA forms on sm_100a (warp instructions per cycle per SM sub-partition).
//   0  FFMA  R, R, R, R      three distinct register sources (FIR inner loop: tap reg x window reg + acc reg)
//   1  FFMA  R, R, c/UR, c/UR   two uniform sources
//   2  FFMA2 three register-pair sources
//   3  FFMA2 two uniform pair sources
//   4  FFMA  R, R, c/UR, R   one uniform source (tap from a kernel parameter)
//   5  as 0 but interleaved 1:1 with IADD3-class integer work (alu pipe)
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, const float *in, float A, float B, int iters) {
    float a[8], b[8], x[8], y[8];
    int n[8];
    for (int i = 0; i < 8; ++i) {
        a[i] = in[threadIdx.x + 32 * i];
        b[i] = in[threadIdx.x + 32 * i + 1000];
        x[i] = in[threadIdx.x + 32 * i + 2000];
        y[i] = in[threadIdx.x + 32 * i + 3000];
        n[i] = threadIdx.x + i;
    }
    const float2 A2 = make_float2(A, A), B2 = make_float2(B, B);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {
                x[i] = fmaf(a[i], b[i], x[i]);
                y[i] = fmaf(b[i], a[(i + 1) & 7], y[i]);
            } else if (MODE == 1) {
                x[i] = fmaf(x[i], A, B);
                y[i] = fmaf(y[i], A, B);
            } else if (MODE == 2) {
                float2 r = __ffma2_rn(make_float2(a[i], b[i]), make_float2(a[(i + 1) & 7], b[(i + 1) & 7]),
                                      make_float2(x[i], y[i]));
                x[i] = r.x;
                y[i] = r.y;
            } else if (MODE == 3) {
                float2 r = __ffma2_rn(make_float2(x[i], y[i]), A2, B2);
                x[i] = r.x;
                y[i] = r.y;
            } else if (MODE == 4) {
                x[i] = fmaf(a[i], A, x[i]);
                y[i] = fmaf(b[i], B, y[i]);
            } else {
                x[i] = fmaf(a[i], b[i], x[i]);
                n[i] = n[i] * 3 + (n[(i + 1) & 7] ^ it);
                y[i] = fmaf(b[i], a[(i + 1) & 7], y[i]);
                n[(i + 3) & 7] += n[i] >> 1;
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + (float)n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(float *out, const float *in, int iters, const char *name, double fp_per_iter, double inst_per_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int dev_clk = 0;
    cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<148 * 8, 256>>>(out, in, 0.999f, 0.001f, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double threads = 148.0 * 8 * 256;
        if (rep)
            printf("%-28s %8.3f ms  %6.2f TFMA/s   %.3f warp-inst/clk/SMSP (at %d MHz nominal)\n", name, ms,
                   threads * fp_per_iter * iters / ms / 1e9,
                   (threads / 32) * inst_per_iter * iters / (ms * 1e-3) / (dev_clk * 1e3) / (148.0 * 4), dev_clk / 1000);
    }
}

int main() {
    float *out, *in;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaMalloc(&in, 8192 * sizeof(float));
    cudaMemset(in, 0, 8192 * sizeof(float));
    const int iters = 20000;
    run<0>(out, in, iters, "FFMA R,R,R", 16, 16);
    run<1>(out, in, iters, "FFMA R,U,U", 16, 16);
    run<2>(out, in, iters, "FFMA2 RR,RR,RR", 16, 8);
    run<3>(out, in, iters, "FFMA2 RR,UU,UU", 16, 8);
    run<4>(out, in, iters, "FFMA R,U,R", 16, 16);
    run<5>(out, in, iters, "FFMA R,R,R + int mix", 16, 16);
    return 0;
}
