// Polyphase FIR resamplers on shared-memory line buffers.
//
// Reference: scipy.signal.resample_poly as called by qam.py:35,37,45,53-57, pal.py:72,77, secam.py:136,149,
// niir.py:106-145, protosecam.py:83-102, mac.py:51-111 — default filter firwin(2*half+1, 1/max(up,down),
// ('kaiser', 5.0)) * up with half = 10*max(up,down), zero-padded edges, centred:
//     out[j] = sum_i x[i] * h[half + j*down - i*up],   n_out = ceil(n*up/down)
// (restated and checked against scipy in oracle/dsp.py:resample_restated).  The taps themselves are computed
// by the host with the same scipy.signal.firwin call and passed in cm_desc.resamplers.
//
// The x2 / x3 filters are half-band / third-band: every tap at a non-zero multiple of `max` from the centre
// is zero to rounding (|h| < 1e-16, tests/test_oracle_dsp.py::test_halfband_structure) and is skipped.
#pragma once
#include "cm_common.cuh"

// x[0..n) -> out[0..2n)      h: 41 dense taps
template <typename T>
__device__ __forceinline__ void fir_up2(T *__restrict__ out, const T *__restrict__ x, int n, const T *__restrict__ h,
                                        int tid, int nthr) {
    T g[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) g[k] = h[39 - 2 * k];
    const T c0 = h[20];
    for (int m = tid; m < n; m += nthr) {
        T acc = (T)0;
        if (m >= 9 && m + 10 < n) {
#pragma unroll
            for (int k = 0; k < 20; ++k) acc = Real<T>::fma_(g[k], x[m - 9 + k], acc);
        } else {
#pragma unroll
            for (int k = 0; k < 20; ++k) {
                int i = m - 9 + k;
                if (i >= 0 && i < n) acc = Real<T>::fma_(g[k], x[i], acc);
            }
        }
        out[2 * m] = c0 * x[m];
        out[2 * m + 1] = acc;
    }
}

// x[0..n) -> out[0..ceil(n/2))    h: 41 dense taps.  `post(j, v)` consumes output j.
template <typename T, class Post>
__device__ __forceinline__ void fir_down2(const T *__restrict__ x, int n, const T *__restrict__ h, int tid, int nthr,
                                          Post post) {
    T g[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) g[k] = h[20 - (2 * k - 19)];     // tap for offset d = 2k-19
    const T c0 = h[20];
    const int n_out = (n + 1) >> 1;
    for (int j = tid; j < n_out; j += nthr) {
        const int ctr = 2 * j;
        T acc = c0 * x[ctr];
        if (ctr >= 19 && ctr + 19 < n) {
#pragma unroll
            for (int k = 0; k < 20; ++k) acc = Real<T>::fma_(g[k], x[ctr + 2 * k - 19], acc);
        } else {
#pragma unroll
            for (int k = 0; k < 20; ++k) {
                int i = ctr + 2 * k - 19;
                if (i >= 0 && i < n) acc = Real<T>::fma_(g[k], x[i], acc);
            }
        }
        post(j, acc);
    }
}

// x[0..n) -> out[0..3n)      h: 61 dense taps
template <typename T>
__device__ __forceinline__ void fir_up3(T *__restrict__ out, const T *__restrict__ x, int n, const T *__restrict__ h,
                                        int tid, int nthr) {
    const T c0 = h[30];
    for (int m = tid; m < n; m += nthr) {
        T a1 = (T)0, a2 = (T)0;
        if (m >= 9 && m + 10 < n) {
#pragma unroll
            for (int k = 0; k < 20; ++k) {
                T v = x[m - 9 + k];
                a1 = Real<T>::fma_(h[58 - 3 * k], v, a1);
                a2 = Real<T>::fma_(h[59 - 3 * k], v, a2);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 20; ++k) {
                int i = m - 9 + k;
                if (i >= 0 && i < n) {
                    T v = x[i];
                    a1 = Real<T>::fma_(h[58 - 3 * k], v, a1);
                    a2 = Real<T>::fma_(h[59 - 3 * k], v, a2);
                }
            }
        }
        out[3 * m] = c0 * x[m];
        out[3 * m + 1] = a1;
        out[3 * m + 2] = a2;
    }
}

// x[0..n) -> ceil(n/3) outputs   h: 61 dense taps
template <typename T, class Post>
__device__ __forceinline__ void fir_down3(const T *__restrict__ x, int n, const T *__restrict__ h, int tid, int nthr,
                                          Post post) {
    const int n_out = (n + 2) / 3;
    for (int j = tid; j < n_out; j += nthr) {
        const int ctr = 3 * j;
        T acc = h[30] * x[ctr];
        const bool inner = (ctr >= 29 && ctr + 29 < n);
#pragma unroll
        for (int d = -29; d <= 29; ++d) {
            if (d % 3 == 0) continue;
            int i = ctr + d;
            if (inner || (i >= 0 && i < n)) acc = Real<T>::fma_(h[30 - d], x[i], acc);
        }
        post(j, acc);
    }
}

// General rational resampler (MAC: 3/8, 3/16, 2/3, 3/2, ...).  x[0..n) -> n_out outputs.
template <typename T, class Post>
__device__ __forceinline__ void fir_general(const T *__restrict__ x, int n, int n_out, const ResHdr rh,
                                            const T *__restrict__ h, int tid, int nthr, Post post) {
    for (int j = tid; j < n_out; j += nthr) {
        const int c = rh.half + j * rh.down;
        int i_hi = c / rh.up;
        if (i_hi > n - 1) i_hi = n - 1;
        int lo_num = c - 2 * rh.half;
        int i_lo = lo_num <= 0 ? 0 : (lo_num + rh.up - 1) / rh.up;
        T acc = (T)0;
        for (int i = i_lo; i <= i_hi; ++i) acc = Real<T>::fma_(h[c - i * rh.up], x[i], acc);
        post(j, acc);
    }
}
