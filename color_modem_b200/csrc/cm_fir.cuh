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
//
// Layout: 1x signals are stored naturally; 2x signals polyphase [E | O] (E[m] = x2[2m], O[m] = x2[2m+1], `hb`
// elements each); 3x signals [P0 | P1 | P2].  In that layout both directions are unit-stride sliding windows:
//     up2:    E[m] = h[20] x[m]            O[m] = sum_k g[k] x[m-9+k]            g[k] = h[39-2k], k = 0..19
//     down2:  y[j] = h[20] E[j] + sum_k g[k] O[j-10+k]
// Each thread produces 4 consecutive outputs from 128-bit loads (7 loads for 80 FMAs).
#pragma once
#include "cm_common.cuh"

template <typename T> struct Vec4 { T v[4]; };

template <typename T>
__device__ __forceinline__ void ld4(const T *p, T *dst) {
    const Vec4<T> t = *reinterpret_cast<const Vec4<T> *>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = t.v[i];
}
template <>
__device__ __forceinline__ void ld4<float>(const float *p, float *dst) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
}
template <typename T>
__device__ __forceinline__ void st4(T *p, const T *src) {
    Vec4<T> t;
#pragma unroll
    for (int i = 0; i < 4; ++i) t.v[i] = src[i];
    *reinterpret_cast<Vec4<T> *>(p) = t;
}
template <>
__device__ __forceinline__ void st4<float>(float *p, const float *src) {
    *reinterpret_cast<float4 *>(p) = make_float4(src[0], src[1], src[2], src[3]);
}

// 28-sample window x[m0-12 .. m0+16) with zeros outside [0, n).  n and m0 are multiples of 4, so every aligned
// 4-vector is either entirely inside or entirely outside the line: 7 predicated 128-bit loads, no divergence.
template <typename T>
__device__ __forceinline__ void load_window28(const T *__restrict__ x, int n, int m0, T *w) {
#pragma unroll
    for (int v = 0; v < 7; ++v) {
        const int i = m0 - 12 + 4 * v;
        if (i >= 0 && i < n) {
            ld4(x + i, w + 4 * v);
        } else {
            w[4 * v] = w[4 * v + 1] = w[4 * v + 2] = w[4 * v + 3] = (T)0;
        }
    }
}

// x[0..n) natural -> E[0..n), O[0..n).  n % 4 == 0, buffers 16-byte aligned.      h: 41 dense taps
template <typename T>
__device__ __forceinline__ void fir_up2(T *__restrict__ E, T *__restrict__ O, const T *__restrict__ x, int n,
                                        const T *__restrict__ h, int tid, int nthr) {
    T g[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) g[k] = h[39 - 2 * k];
    const T c0 = h[20];
    for (int m0 = 4 * tid; m0 < n; m0 += 4 * nthr) {
        T w[28];
        load_window28(x, n, m0, w);
        T e[4], o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            e[r] = c0 * w[12 + r];
            T acc = (T)0;
#pragma unroll
            for (int k = 0; k < 20; ++k) acc = Real<T>::fma_(g[k], w[3 + r + k], acc);   // x[m0+r-9+k]
            o[r] = acc;
        }
        st4(E + m0, e);
        st4(O + m0, o);
    }
}

// E[0..n), O[0..n) -> n outputs; post(j0, y[4]) consumes outputs j0..j0+3.       h: 41 dense taps
template <typename T, class Post>
__device__ __forceinline__ void fir_down2(const T *__restrict__ E, const T *__restrict__ O, int n,
                                          const T *__restrict__ h, int tid, int nthr, Post post) {
    T g[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) g[k] = h[39 - 2 * k];
    const T c0 = h[20];
    for (int j0 = 4 * tid; j0 < n; j0 += 4 * nthr) {
        T w[28], e[4], y[4];
        load_window28(O, n, j0, w);
        ld4(E + j0, e);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            T acc = c0 * e[r];
#pragma unroll
            for (int k = 0; k < 20; ++k) acc = Real<T>::fma_(g[k], w[2 + r + k], acc);   // O[j0+r-10+k]
            y[r] = acc;
        }
        post(j0, y);
    }
}

// x[0..n) natural -> out[0..3n) natural (3x paths are not on the headline configs; kept simple)   h: 61 taps
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

// x[0..n) natural -> ceil(n/3) outputs   h: 61 dense taps
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
