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
