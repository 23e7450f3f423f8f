// Does FFMA2 (fma.rn.f32x2) free issue slots?  Per loop iteration:
//   A: 16 FFMA (R,U,R)  + 8 LOP3 (alu pipe)     24 issue slots if every instruction takes one
//   B:  8 FFMA2 (RR,UU,RR) + 8 LOP3              16 issue slots if FFMA2 is single-issue, 24 if it holds the port 2 cycles
//   C: 16 FFMA only      D: 8 FFMA2 only      E: 8 LOP3 only
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, const float *in, float A, float B, int iters, unsigned m) {
    float x[8], y[8];
    unsigned n[8];
    for (int i = 0; i < 8; ++i) {
        x[i] = in[threadIdx.x + 32 * i + 2000];
        y[i] = in[threadIdx.x + 32 * i + 3000];
        n[i] = threadIdx.x * 77u + i;
    }
    const float2 A2 = make_float2(A, A);
    float2 v[8];
    for (int i = 0; i < 8; ++i) v[i] = make_float2(x[i], y[i]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 2) {
                x[i] = fmaf(x[i], A, y[(i + 1) & 7]);
                y[i] = fmaf(y[i], A, x[(i + 3) & 7]);
            }
            if (MODE == 1 || MODE == 3) {
                v[i] = __ffma2_rn(v[i], A2, v[(i + 2) & 7]);
            }
            if (MODE == 0 || MODE == 1 || MODE == 4) n[i] = (n[i] & m) ^ n[(i + 1) & 7];
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + (float)n[i] + v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(float *out, const float *in, int iters, const char *name, double inst_per_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<148 * 8, 256>>>(out, in, 0.999f, 0.001f, iters, 0x7fffffffu);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double warps_per_smsp = 8.0 * 8 / 4;
        if (rep)
            printf("%-28s %8.3f ms   %.2f cycles/iteration/SMSP-warp-set (1.965 GHz)  %.3f inst/clk/SMSP\n", name, ms,
                   ms * 1e-3 * 1.965e9 / iters / warps_per_smsp, inst_per_iter * iters * warps_per_smsp / (ms * 1e-3 * 1.965e9));
    }
}

int main() {
    float *out, *in;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaMalloc(&in, 8192 * sizeof(float));
    cudaMemset(in, 0, 8192 * sizeof(float));
    const int iters = 20000;
    run<0>(out, in, iters, "A 16 FFMA + 8 LOP3", 24);
    run<1>(out, in, iters, "B 8 FFMA2 + 8 LOP3", 16);
    run<2>(out, in, iters, "C 16 FFMA", 16);
    run<3>(out, in, iters, "D 8 FFMA2", 8);
    run<4>(out, in, iters, "E 8 LOP3", 8);
    return 0;
}
