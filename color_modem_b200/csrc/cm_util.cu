// Stand-alone DSP kit entry points of the C ABI: the reference's utils.FilterFunction.__call__ (utils.py:28-36) as a
// kernel of its own, for callers of the reference's L0 kit (and for the luma notch of composed comb wrappers).
#include "cm_host.h"
#include "cm_iir.cuh"

void cm_build_filter_table(const cm_filter &f, FiltHdr &h, std::vector<double> &tab);      // cm_api.cu

// One warp per row: stage the row in shared memory (last sample replicated up to the padded length), run the exact
// chunk-parallel cascade (cm_iir.cuh), store with the group-delay shift.
template <typename T>
__global__ void __launch_bounds__(32)
k_filter_rows(const FiltHdr fh, const T *__restrict__ tab, const T *__restrict__ in, T *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *buf = reinterpret_cast<T *>(smem_raw);
    const T *src = in + (size_t)blockIdx.x * fh.n;
    T *dst = out + (size_t)blockIdx.x * fh.n;
    for (int j = threadIdx.x; j < fh.npad; j += 32) buf[j] = src[j < fh.n ? j : fh.n - 1];
    __syncwarp();
    warp_iir<T, 1>(tab, fh, [&](int q, int, int) { return buf[q]; }, [&](int j, T v) { dst[j] = v; });
}

template <typename T>
static int filter_rows(const cm_filter *f, const void *in, void *out, int32_t nrows, cudaStream_t st) {
    FiltHdr fh;
    memset(&fh, 0, sizeof(fh));
    std::vector<double> tab;
    cm_filter one = *f;
    one.rate = 1;
    cm_build_filter_table(one, fh, tab);
    std::vector<T> host(tab.size());
    for (size_t i = 0; i < tab.size(); ++i) host[i] = (T)tab[i];
    T *d_tab = nullptr;
    CUDA_TRY(cudaMalloc(&d_tab, host.size() * sizeof(T)));
    cudaError_t e = cudaMemcpyAsync(d_tab, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    const size_t smem = (size_t)fh.npad * sizeof(T);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_filter_rows<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        k_filter_rows<T><<<(unsigned)nrows, 32, smem, st>>>(fh, d_tab, (const T *)in, (T *)out);
        cm_count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);          // `host` and d_tab must outlive the launch
    cudaFree(d_tab);
    if (e != cudaSuccess) return cm_fail(CM_ERR_CUDA, "cm_filter_rows: %s", cudaGetErrorString(e));
    return CM_OK;
}

extern "C" int cm_filter_rows(const cm_filter *f, int precision, const void *in, void *out, int32_t nrows,
                              void *stream) {
    if (!f || !in || !out) return cm_fail(CM_ERR_INVALID, "null argument%s");
    if (precision != CM_FP32 && precision != CM_FP64) return cm_fail(CM_ERR_INVALID, "bad precision%s");
    if (f->nsec <= 0 || f->nsec > CM_MAX_SECTIONS || f->shift < 0 || f->n <= 0 || nrows < 0)
        return cm_fail(CM_ERR_INVALID, "bad cm_filter%s");
    if (nrows == 0) return CM_OK;
    return precision == CM_FP32 ? filter_rows<float>(f, in, out, nrows, (cudaStream_t)stream)
                                : filter_rows<double>(f, in, out, nrows, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------------
// Measured FP32 multiply-add peak of the device: the denominator of bench.py's `roofline_fma`.  Eight independent packed
// chains per thread, the multiplier and addend uniform (constant-bank) operands, 8 x 256-thread CTAs per SM — the form in
// which sm_100 issues FFMA2 at full rate (tools/ubench/fma_forms.cu).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fma_peak(float *out, float a, float b, int iters) {
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], A, B);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int cm_measure_fma_peak(double *tfma_per_s) {
    if (!tfma_per_s) return cm_fail(CM_ERR_INVALID, "null argument%s");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int ctas = sms * 8, iters = 20000;
    float *out = nullptr;
    CUDA_TRY(cudaMalloc(&out, (size_t)ctas * 256 * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 0.f;
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 4 && e == cudaSuccess; ++rep) {
        cudaEventRecord(e0);
        k_fma_peak<<<ctas, 256>>>(out, 0.999f, 0.001f, iters);
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms > 0.f && (best == 0.f || ms < best)) best = ms;          // rep 0 warms up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) return cm_fail(CM_ERR_CUDA, "cm_measure_fma_peak: %s", cudaGetErrorString(e));
    *tfma_per_s = (double)ctas * 256 * 16.0 * iters / (best * 1e-3) / 1e12;
    return CM_OK;
}
