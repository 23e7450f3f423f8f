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

// ------------------------------------------------------------------------------------------------------------
// fp32: packed f32x2 versions (FFMA2: one issue slot for two FMAs).  The window is held as the 14 naturally aligned
// pairs W2[i] = (w[2i], w[2i+1]) that the 128-bit loads deliver; a pair of neighbouring outputs (r, r+1) needs, for
// tap k, the pair (w[c+r+k], w[c+r+k+1]), which is aligned for every other tap.  Five accumulator pairs
//     (0,1), (2,3) over the taps of one parity  +  (-1,0), (1,2), (3,4) over the taps of the other parity
// give the four outputs with 50 FFMA2 instead of 80 FFMA (the outer halves of (-1,0) and (3,4) are discarded).
// The taps are (h, h) pairs in the kernel-parameter constant bank (DevParams::fircp), `hp` points at one of them and
// must be indexed with compile-time constants.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_window28_f2(const float *__restrict__ x, int n, int m0, float2 *w2) {
#pragma unroll
    for (int v = 0; v < 7; ++v) {
        const int i = m0 - 12 + 4 * v;
        if (i >= 0 && i < n) {
            const float4 t = *reinterpret_cast<const float4 *>(x + i);
            w2[2 * v] = make_float2(t.x, t.y);
            w2[2 * v + 1] = make_float2(t.z, t.w);
        } else {
            w2[2 * v] = w2[2 * v + 1] = make_float2(0.f, 0.f);
        }
    }
}

// out[r] (+)= sum_{k<20} h[T0 - TS k] w[C + r + k],  r = 0..3;  C = 3 (up) or 2 (down).  a01 / a23 carry the initial
// values of the outputs (0,1) / (2,3).
template <int C, int T0, int TS>
__device__ __forceinline__ void fir4_f2(const float (*hp)[2], const float2 *w2, float2 a01, float2 a23, float *out) {
    constexpr int PA = (C & 1) ? 1 : 0;      // parity of the taps whose (r, r+1) pairs are aligned for even r
    float2 bm = make_float2(0.f, 0.f), b12 = bm, b34 = bm;
#pragma unroll
    for (int kk = 0; kk < 10; ++kk) {
        const int k = 2 * kk + PA;
        const float2 g = make_float2(hp[T0 - TS * k][0], hp[T0 - TS * k][1]);
        a01 = __ffma2_rn(g, w2[(C + k) / 2], a01);
        a23 = __ffma2_rn(g, w2[(C + 2 + k) / 2], a23);
    }
#pragma unroll
    for (int kk = 0; kk < 10; ++kk) {
        const int k = 2 * kk + (1 - PA);
        const float2 g = make_float2(hp[T0 - TS * k][0], hp[T0 - TS * k][1]);
        bm = __ffma2_rn(g, w2[(C - 1 + k) / 2], bm);
        b12 = __ffma2_rn(g, w2[(C + 1 + k) / 2], b12);
        b34 = __ffma2_rn(g, w2[(C + 3 + k) / 2], b34);
    }
    out[0] = a01.x + bm.y;
    out[1] = a01.y + b12.x;
    out[2] = a23.x + b12.y;
    out[3] = a23.y + b34.x;
}

// Taps of one resampler: dense taps h (shared or constant memory) and, for the packed fp32 path, the duplicated
// pairs in the constant bank (nullptr: scalar path, e.g. the MAC kernels whose taps live in shared memory).
template <typename T>
struct FirTaps {
    const T *h;
    const float (*hp)[2];
#ifdef CM_NO_PACKED_FIR
    __device__ __forceinline__ FirTaps(const T *h_, const float (*)[2]) : h(h_), hp(nullptr) {}
#else
    __device__ __forceinline__ FirTaps(const T *h_, const float (*hp_)[2]) : h(h_), hp(hp_) {}
#endif
};

// x[0..n) natural -> E[0..n), O[0..n).  n % 4 == 0, buffers 16-byte aligned.      h: 41 dense taps
template <typename T>
__device__ __forceinline__ void fir_up2(T *__restrict__ E, T *__restrict__ O, const T *__restrict__ x, int n,
                                        const FirTaps<T> tp, int tid, int nthr) {
    const T *__restrict__ h = tp.h;
    if constexpr (IsF32<T>::value) {
        if (tp.hp) {
            for (int m0 = 4 * tid; m0 < n; m0 += 4 * nthr) {
                float2 w2[14];
                load_window28_f2(x, n, m0, w2);
                const float2 c0 = make_float2(tp.hp[20][0], tp.hp[20][1]), z = make_float2(0.f, 0.f);
                const float2 e01 = __fmul2_rn(c0, w2[6]), e23 = __fmul2_rn(c0, w2[7]);
                float o[4];
                fir4_f2<3, 39, 2>(tp.hp, w2, z, z, o);
                *reinterpret_cast<float4 *>(E + m0) = make_float4(e01.x, e01.y, e23.x, e23.y);
                *reinterpret_cast<float4 *>(O + m0) = make_float4(o[0], o[1], o[2], o[3]);
            }
            return;
        }
    }
    // taps are read with compile-time indices right in the FMA loops: for h in the kernel-parameter constant
    // bank (DevParams::firc) they become constant operands instead of 21 registers
    for (int m0 = 4 * tid; m0 < n; m0 += 4 * nthr) {
        T w[28];
        load_window28(x, n, m0, w);
        T e[4], o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            e[r] = h[20] * w[12 + r];
            T acc = (T)0;
#pragma unroll
            for (int k = 0; k < 20; ++k) acc = Real<T>::fma_(h[39 - 2 * k], w[3 + r + k], acc);   // x[m0+r-9+k]
            o[r] = acc;
        }
        st4(E + m0, e);
        st4(O + m0, o);
    }
}

// packed down2 of one quad: y[r] = h[20] E[j0+r] + sum_k h[39-2k] O[j0+r-10+k]
__device__ __forceinline__ void down2_quad_f2(const float (*hp)[2], const float *__restrict__ E,
                                              const float *__restrict__ O, int n, int j0, float *y) {
    float2 w2[14];
    load_window28_f2(O, n, j0, w2);
    const float4 e = *reinterpret_cast<const float4 *>(E + j0);
    const float2 c0 = make_float2(hp[20][0], hp[20][1]);
    fir4_f2<2, 39, 2>(hp, w2, __fmul2_rn(c0, make_float2(e.x, e.y)), __fmul2_rn(c0, make_float2(e.z, e.w)), y);
}

// One quad of a down2: y[r] = h[20] E[j0+r] + sum_k h[39-2k] O[j0+r-10+k], r = 0..3.        h: 41 dense taps
// (taps are read with compile-time indices right in the FMA loops: in the kernel-parameter constant bank,
// DevParams::firc, they become constant operands instead of 21 registers)
template <typename T>
__device__ __forceinline__ void down2_quad(const FirTaps<T> tp, const T *__restrict__ E, const T *__restrict__ O, int n,
                                           int j0, T *y) {
    if constexpr (IsF32<T>::value) {
        if (tp.hp) {
            down2_quad_f2(tp.hp, E, O, n, j0, y);
            return;
        }
    }
    const T *__restrict__ h = tp.h;
    T w[28], e[4];
    load_window28(O, n, j0, w);
    ld4(E + j0, e);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        T acc = h[20] * e[r];
#pragma unroll
        for (int k = 0; k < 20; ++k) acc = Real<T>::fma_(h[39 - 2 * k], w[2 + r + k], acc);   // O[j0+r-10+k]
        y[r] = acc;
    }
}

// E[0..n), O[0..n) -> n outputs; post(j0, y[4]) consumes outputs j0..j0+3.
template <typename T, class Post>
__device__ __forceinline__ void fir_down2(const T *__restrict__ E, const T *__restrict__ O, int n,
                                          const FirTaps<T> tp, int tid, int nthr, Post post) {
    for (int j0 = 4 * tid; j0 < n; j0 += 4 * nthr) {
        T y[4];
        down2_quad(tp, E, O, n, j0, y);
        post(j0, y);
    }
}

// Two down2's at once (same taps): post(j0, y1[4], y2[4]).  Used where two demodulated signals are combined
// sample by sample right after the decimation (PAL-D: (S, D) -> (u, v)).
template <typename T, class Post>
__device__ __forceinline__ void fir_down2_pair(const T *__restrict__ E1, const T *__restrict__ O1,
                                               const T *__restrict__ E2, const T *__restrict__ O2, int n,
                                               const FirTaps<T> tp, int tid, int nthr, Post post) {
    for (int j0 = 4 * tid; j0 < n; j0 += 4 * nthr) {
        T y1[4], y2[4];
        down2_quad(tp, E1, O1, n, j0, y1);
        down2_quad(tp, E2, O2, n, j0, y2);
        post(j0, y1, y2);
    }
}

// ---- x3 / :3, polyphase [P0 | P1 | P2] (P_r[m] = x3[3m + r]), 61 dense taps, every tap at a non-zero multiple of 3
// from the centre is zero:
//     up3:    P0[m] = h[30] x[m]     P1[m] = sum_k h[58-3k] x[m-9+k]     P2[m] = sum_k h[59-3k] x[m-9+k]   k = 0..19
//     down3:  y[j]  = h[30] P0[j] + sum_a ( h[29-3a] P1[j+a] + h[28-3a] P2[j+a] ),  a = -10..9
template <typename T>
__device__ __forceinline__ void fir_up3(T *__restrict__ P0, T *__restrict__ P1, T *__restrict__ P2,
                                        const T *__restrict__ x, int n, const FirTaps<T> tp, int tid, int nthr) {
    const T *__restrict__ h = tp.h;
    if constexpr (IsF32<T>::value) {
        if (tp.hp) {
            for (int m0 = 4 * tid; m0 < n; m0 += 4 * nthr) {
                float2 w2[14];
                load_window28_f2(x, n, m0, w2);
                const float2 c0 = make_float2(tp.hp[30][0], tp.hp[30][1]), z = make_float2(0.f, 0.f);
                const float2 e01 = __fmul2_rn(c0, w2[6]), e23 = __fmul2_rn(c0, w2[7]);
                float o1[4], o2[4];
                fir4_f2<3, 58, 3>(tp.hp, w2, z, z, o1);
                fir4_f2<3, 59, 3>(tp.hp, w2, z, z, o2);
                *reinterpret_cast<float4 *>(P0 + m0) = make_float4(e01.x, e01.y, e23.x, e23.y);
                *reinterpret_cast<float4 *>(P1 + m0) = make_float4(o1[0], o1[1], o1[2], o1[3]);
                *reinterpret_cast<float4 *>(P2 + m0) = make_float4(o2[0], o2[1], o2[2], o2[3]);
            }
            return;
        }
    }
    for (int m0 = 4 * tid; m0 < n; m0 += 4 * nthr) {
        T w[28], p0[4], p1[4], p2[4];
        load_window28(x, n, m0, w);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            p0[r] = h[30] * w[12 + r];
            T a1 = (T)0, a2 = (T)0;
#pragma unroll
            for (int k = 0; k < 20; ++k) {
                a1 = Real<T>::fma_(h[58 - 3 * k], w[3 + r + k], a1);
                a2 = Real<T>::fma_(h[59 - 3 * k], w[3 + r + k], a2);
            }
            p1[r] = a1;
            p2[r] = a2;
        }
        st4(P0 + m0, p0);
        st4(P1 + m0, p1);
        st4(P2 + m0, p2);
    }
}

// Taps of down3 in window order: h[59 - 3k] multiplies P1[j0 + r - 10 + k], h[58 - 3k] multiplies P2[j0 + r - 10 + k]
template <typename T>
struct Down3Taps {
    const T *h;       // 61 dense taps (constant bank): h[59 - 3k] multiplies P1, h[58 - 3k] multiplies P2   (a = k - 10)
    const float (*hp)[2];
    __device__ __forceinline__ explicit Down3Taps(const FirTaps<T> tp) : h(tp.h), hp(tp.hp) {}
    // w1 / w2: 28-sample windows of P1 / P2 starting at j0 - 12; e: P0[j0 .. j0+3]
    __device__ __forceinline__ void apply(const T *w1, const T *w2, const T *e, T *y) const {
        if constexpr (IsF32<T>::value) {
            if (hp) {
                float2 p1[14], p2[14];
#pragma unroll
                for (int i = 0; i < 14; ++i) {
                    p1[i] = make_float2(w1[2 * i], w1[2 * i + 1]);
                    p2[i] = make_float2(w2[2 * i], w2[2 * i + 1]);
                }
                const float2 c0 = make_float2(hp[30][0], hp[30][1]);
                float t[4];
                fir4_f2<2, 59, 3>(hp, p1, __fmul2_rn(c0, make_float2(e[0], e[1])), __fmul2_rn(c0, make_float2(e[2], e[3])), t);
                fir4_f2<2, 58, 3>(hp, p2, make_float2(t[0], t[1]), make_float2(t[2], t[3]), y);
                return;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            T acc = h[30] * e[r];
#pragma unroll
            for (int k = 0; k < 20; ++k) {
                acc = Real<T>::fma_(h[59 - 3 * k], w1[2 + r + k], acc);
                acc = Real<T>::fma_(h[58 - 3 * k], w2[2 + r + k], acc);
            }
            y[r] = acc;
        }
    }
};

// plain down3 of a polyphase buffer: 4 outputs y[j0..j0+3]
template <typename T>
__device__ __forceinline__ void down3_quad(const Down3Taps<T> &tp, const T *P0, const T *P1, const T *P2, int n, int j0,
                                           T *y) {
    T w1[28], w2[28], e[4];
    load_window28(P1, n, j0, w1);
    load_window28(P2, n, j0, w2);
    ld4(P0 + j0, e);
    tp.apply(w1, w2, e, y);
}

// down3 of the elementwise product of two polyphase buffers A * B
template <typename T>
__device__ __forceinline__ void down3_quad_prod(const Down3Taps<T> &tp, const T *A, const T *B, int hb, int n, int j0,
                                                T *y) {
    T w1[28], w2[28], e[4], t[28];
    load_window28(A + hb, n, j0, w1);
    load_window28(B + hb, n, j0, t);
#pragma unroll
    for (int i = 0; i < 28; ++i) w1[i] *= t[i];
    load_window28(A + 2 * hb, n, j0, w2);
    load_window28(B + 2 * hb, n, j0, t);
#pragma unroll
    for (int i = 0; i < 28; ++i) w2[i] *= t[i];
    T ea[4], eb[4];
    ld4(A + j0, ea);
    ld4(B + j0, eb);
#pragma unroll
    for (int i = 0; i < 4; ++i) e[i] = ea[i] * eb[i];
    tp.apply(w1, w2, e, y);
}

// General rational resampler (MAC: 3/8, 3/16, 2/3, 3/2, ...).  x[0..n) -> n_out outputs.
template <typename T, class Load, class Post>
__device__ __forceinline__ void fir_general(Load x, int n, int n_out, const ResHdr rh,
                                            const T *__restrict__ h, int tid, int nthr, Post post) {
    for (int j = tid; j < n_out; j += nthr) {
        const int c = rh.half + j * rh.down;
        int i_hi = c / rh.up;
        if (i_hi > n - 1) i_hi = n - 1;
        int lo_num = c - 2 * rh.half;
        int i_lo = lo_num <= 0 ? 0 : (lo_num + rh.up - 1) / rh.up;
        T acc = (T)0;
        for (int i = i_lo; i <= i_hi; ++i) acc = Real<T>::fma_(h[c - i * rh.up], x(i), acc);
        post(j, acc);
    }
}

// Rational resampler in aligned polyphase form (PolyHdr, cm_common.cuh): one thread per output GROUP (the UP outputs
// j = UP m + r), one 128-bit load of the line per four samples shared by the UP phases, the taps of each phase read with
// 128-bit loads that are broadcasts whenever `down` is a multiple of 4, UP independent accumulators.  Shared-memory
// bandwidth, not arithmetic, binds this loop: one output per thread (every sample loaded once per output) ran at 2.8
// TFMA/s.  `xpad` points at the element of the FRONT PAD that is FP samples ahead of the line (a multiple of 32 elements
// from the start of the skewed segment, so the skew is the same function for every resampler): xpad[skew(FP + i)] = x[i],
// zeros outside [0, n).
// Skewed line layout of the resampler inputs: element i of the padded line lives at i + 4 (i / 32) — one 16-byte chunk
// of slack per 128 bytes.  The lanes of a warp start their windows `down` samples apart (8 or 16 for the MAC ratios):
// unskewed, their 128-bit loads fall on 2 or 4 bank groups only (4-way conflicts measured as the bound of the loop).
// (mask: all ones with the skew, zero without — lines whose resamplers step by 2 or 3 samples need none)
__device__ __forceinline__ int poly_skew(int i, int mask) { return i + (((i >> 5) << 2) & mask); }

template <typename T, int UP, bool SKEW, class Post>
__device__ __forceinline__ void fir_poly_up(const T *__restrict__ xpad, int n_out, const PolyHdr &ph, const T *__restrict__ G,
                                            int tid, int nthr, Post post) {
    const int ngroups = (n_out + UP - 1) / UP;
    for (int m = tid; m < ngroups; m += nthr) {
        const int idx = m * ph.down + ph.lo0 + ph.FP;
        const int a = idx & 3, s0 = idx - a;
        const T *__restrict__ g = G + (size_t)(a * UP) * ph.stride;
        T acc[UP];
#pragma unroll
        for (int r = 0; r < UP; ++r) acc[r] = (T)0;
        for (int q = 0; q < ph.KU; q += 4) {
            T xv[4];
            ld4(xpad + (SKEW ? poly_skew(s0 + q, -1) : s0 + q), xv);
#pragma unroll
            for (int r = 0; r < UP; ++r) {
                T gv[4];
                ld4(g + r * ph.stride + q, gv);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[r] = Real<T>::fma_(gv[i], xv[i], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < UP; ++r)
            if (UP * m + r < n_out) post(UP * m + r, acc[r]);
    }
}

// The same with the taps in the constant bank (MacConst, cm_common.cuh), fully unrolled.  ALIGNED: `down` is a multiple of 4,
// every group has the alignment the host built the table ct[0] for.  Otherwise the groups are walked in four passes by
// alignment class j = m mod 4 (the alignment of group m depends on m mod 4 only), ct[j] the table of class j.
template <typename T, int UP, int KU, bool SKEW, bool ALIGNED, int NCLS, class Post>
__device__ __forceinline__ void fir_poly_ct(const T *__restrict__ xpad, int n_out, const PolyHdr &ph, const T (&ct)[NCLS][UP][KU],
                                            int tid, int nthr, Post post) {
    const int ngroups = (n_out + UP - 1) / UP;
#pragma unroll
    for (int j = 0; j < (ALIGNED ? 1 : 4); ++j) {
        for (int m = ALIGNED ? tid : 4 * tid + j; m < ngroups; m += ALIGNED ? nthr : 4 * nthr) {
            const int idx = m * ph.down + ph.lo0 + ph.FP;
            const int s0 = idx & ~3;
            T acc[UP];
            if constexpr (IsF32<T>::value) {
                // packed: the even and the odd taps of a phase accumulate side by side (FFMA2, taps as 64-bit uniform operands)
                float2 a2[UP];
#pragma unroll
                for (int r = 0; r < UP; ++r) a2[r] = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < KU; q += 4) {
                    const float4 xv = *reinterpret_cast<const float4 *>(xpad + (SKEW ? poly_skew(s0 + q, -1) : s0 + q));
                    const float2 x01 = make_float2(xv.x, xv.y), x23 = make_float2(xv.z, xv.w);
#pragma unroll
                    for (int r = 0; r < UP; ++r) {
                        const float (&c)[KU] = ct[ALIGNED ? 0 : j][r];
                        a2[r] = __ffma2_rn(make_float2(c[q], c[q + 1]), x01, a2[r]);
                        a2[r] = __ffma2_rn(make_float2(c[q + 2], c[q + 3]), x23, a2[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < UP; ++r) acc[r] = a2[r].x + a2[r].y;
            } else {
#pragma unroll
                for (int r = 0; r < UP; ++r) acc[r] = (T)0;
#pragma unroll
                for (int q = 0; q < KU; q += 4) {
                    T xv[4];
                    ld4(xpad + (SKEW ? poly_skew(s0 + q, -1) : s0 + q), xv);
#pragma unroll
                    for (int r = 0; r < UP; ++r)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[r] = Real<T>::fma_(ct[ALIGNED ? 0 : j][r][q + i], xv[i], acc[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < UP; ++r)
                if (UP * m + r < n_out) post(UP * m + r, acc[r]);
        }
    }
}

template <typename T, class Post>
__device__ __forceinline__ void fir_poly(const T *__restrict__ xpad, int n_out, const PolyHdr &ph, const T *__restrict__ G,
                                         int tid, int nthr, Post post) {
    if (ph.skew) {          // the skew only occurs with the 3/8 and 3/16 families (down a multiple of 8)
        if (ph.up == 3) fir_poly_up<T, 3, true>(xpad, n_out, ph, G, tid, nthr, post);
        else if (ph.up == 2) fir_poly_up<T, 2, true>(xpad, n_out, ph, G, tid, nthr, post);
        else if (ph.up == 1) fir_poly_up<T, 1, true>(xpad, n_out, ph, G, tid, nthr, post);
        else fir_poly_up<T, 4, true>(xpad, n_out, ph, G, tid, nthr, post);
        return;
    }
    if (ph.up == 1) fir_poly_up<T, 1, false>(xpad, n_out, ph, G, tid, nthr, post);
    else if (ph.up == 2) fir_poly_up<T, 2, false>(xpad, n_out, ph, G, tid, nthr, post);
    else if (ph.up == 3) fir_poly_up<T, 3, false>(xpad, n_out, ph, G, tid, nthr, post);
    else fir_poly_up<T, 4, false>(xpad, n_out, ph, G, tid, nthr, post);
}
