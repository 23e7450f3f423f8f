// Chunk-parallel exact IIR: one warp filters one line.
//
// Reference semantics (color_modem/utils.py:28-36, scipy.signal.lfilter): causal recursion with ZERO initial
// state at sample 0; with group-delay compensation `shift` the input is extended by `shift` copies of its last
// sample and the first `shift` outputs are dropped.
//
// Parallelisation: the line (n + shift samples) is cut into super-chunks of 32*L samples; lane t owns L
// consecutive samples in registers.  For every biquad section (DF-II transposed, state s = (s1, s2)):
//   1. each lane runs the recursion over its chunk from zero state           -> y0[i], end state e_t
//   2. true end states:  S_t = M S_{t-1} + e_t,  M = A^L,  A = [[-a1, 1], [-a2, 0]]   (Kogge-Stone scan over
//      lanes with the precomputed powers M, M^2, M^4, M^8, M^16; the carry of the previous super-chunk
//      enters through lane 0)
//   3. y[i] = y0[i] + (A^i S_{t-1})[0]  — the zero-input response of the incoming state, from a table h[i]
// This is exact (no impulse-response truncation): it is the same linear recursion, re-associated.
//
// Section table layout (elements of T), stride = FiltHdr::stride, built on the host in float64 (cm_api.cu):
//   [0..4]   b0 b1 b2 a1 a2
//   [5..24]  M^1, M^2, M^4, M^8, M^16   (row-major 2x2 each)
//   [25..]   h[i] = (A^i)[0][0], (A^i)[0][1]   for i = 0 .. L-1
#pragma once
#include "cm_common.cuh"

#define CM_SEC_COEF 0
#define CM_SEC_MPOW 5
#define CM_SEC_H 25

template <typename T>
__device__ __forceinline__ T shfl_up_t(T v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
template <typename T>
__device__ __forceinline__ T shfl_idx_t(T v, int l) { return __shfl_sync(0xffffffffu, v, l); }

// load(j)  -> input sample j (caller clamps nothing: j is already clamped to [0, n-1])
// store(j, v) is called for every output sample j in [0, n)
template <typename T, class Load, class Store>
__device__ __forceinline__ void warp_iir(const T *__restrict__ tab, const FiltHdr fh, Load load, Store store) {
    const int lane = threadIdx.x & 31;
    const int L = fh.L;
    const int n = fh.n;
    const int total = n + fh.shift;
    T carry1 = (T)0, carry2 = (T)0;           // lane s keeps the inter-super-chunk carry of section s

    for (int c = 0; c < fh.nsuper; ++c) {
        const int base = (c * 32 + lane) * L;
        T y[CM_LMAX];
#pragma unroll
        for (int i = 0; i < CM_LMAX; ++i) {
            if (i >= L) break;
            int j = base + i;
            y[i] = load(j < n ? j : n - 1);
        }
        for (int s = 0; s < fh.nsec; ++s) {
            const T *ts = tab + s * fh.stride;
            const T b0 = ts[0], b1 = ts[1], b2 = ts[2], na1 = -ts[3], na2 = -ts[4];
            T s1 = (T)0, s2 = (T)0;
#pragma unroll
            for (int i = 0; i < CM_LMAX; ++i) {
                if (i >= L) break;
                T x = y[i];
                T o = Real<T>::fma_(b0, x, s1);
                s1 = Real<T>::fma_(b1, x, s2);
                s1 = Real<T>::fma_(na1, o, s1);
                s2 = b2 * x;
                s2 = Real<T>::fma_(na2, o, s2);
                y[i] = o;
            }
            const T *mp = ts + CM_SEC_MPOW;
            const T c1 = shfl_idx_t(carry1, s), c2 = shfl_idx_t(carry2, s);
            if (lane == 0) {       // carry of the previous super-chunk (zero for the first)
                s1 += mp[0] * c1 + mp[1] * c2;
                s2 += mp[2] * c1 + mp[3] * c2;
            }
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int d = 1 << k;
                T v1 = shfl_up_t(s1, d), v2 = shfl_up_t(s2, d);
                if (lane >= d) {
                    s1 += mp[4 * k + 0] * v1 + mp[4 * k + 1] * v2;
                    s2 += mp[4 * k + 2] * v1 + mp[4 * k + 3] * v2;
                }
            }
            T in1 = shfl_up_t(s1, 1), in2 = shfl_up_t(s2, 1);
            if (lane == 0) { in1 = c1; in2 = c2; }
            {
                T e1 = shfl_idx_t(s1, 31), e2 = shfl_idx_t(s2, 31);
                if (lane == s) { carry1 = e1; carry2 = e2; }
            }
            const T *h = ts + CM_SEC_H;
#pragma unroll
            for (int i = 0; i < CM_LMAX; ++i) {
                if (i >= L) break;
                y[i] = Real<T>::fma_(h[2 * i], in1, Real<T>::fma_(h[2 * i + 1], in2, y[i]));
            }
        }
#pragma unroll
        for (int i = 0; i < CM_LMAX; ++i) {
            if (i >= L) break;
            int j = base + i;
            if (j >= fh.shift && j < total) store(j - fh.shift, y[i]);
        }
    }
}
