// Global-memory I/O helpers: 4 pixels per thread, 32/128-bit accesses, u8 <-> real conversion.
#pragma once
#include "cm_common.cuh"
#include "cm_fir.cuh"

template <typename T>
__device__ __forceinline__ void copy_taps(T *dst, const DevParams<T> &p, int nres) {
    int total = 0;
    for (int r = 0; r < nres; ++r) total = max(total, p.res[r].off + p.res[r].ntaps);
    for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = p.taps[i];
}

// 4 interleaved RGB pixels starting at pixel offset `px` (multiple of 4) of the input buffer
template <typename T>
__device__ __forceinline__ void load_rgb4(const IoArgs<T> &io, size_t px, T r[4], T g[4], T b[4]) {
    if (io.in_f) {
        T v[12];
#pragma unroll
        for (int q = 0; q < 3; ++q) ld4(io.in_f + px * 3 + 4 * q, v + 4 * q);
#pragma unroll
        for (int i = 0; i < 4; ++i) { r[i] = v[3 * i]; g[i] = v[3 * i + 1]; b[i] = v[3 * i + 2]; }
    } else {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(io.in_u8 + px * 3);
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
        unsigned char bytes[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bytes[i] = (w0 >> (8 * i)) & 0xff;
            bytes[4 + i] = (w1 >> (8 * i)) & 0xff;
            bytes[8 + i] = (w2 >> (8 * i)) & 0xff;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[i] = Real<T>::from_u8(bytes[3 * i]);
            g[i] = Real<T>::from_u8(bytes[3 * i + 1]);
            b[i] = Real<T>::from_u8(bytes[3 * i + 2]);
        }
    }
}

// composite row -> T (natural layout), from the u8 frame ((5*(v/255) - 1)/3, image.py:23-25,62) or the float buffer
template <typename T>
__device__ __forceinline__ void load_comp_row(T *dst, const IoArgs<T> &io, int fidx, int row, int Wc) {
    const size_t base = ((size_t)fidx * io.nrows + row) * Wc;
    for (int x = 4 * threadIdx.x; x < Wc; x += 4 * blockDim.x) {
        T v[4];
        if (io.in_f) {
            ld4(io.in_f + base + x, v);
        } else {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(io.in_u8 + base + x));
#pragma unroll
            for (int i = 0; i < 4; ++i)
                v[i] = ((T)5 * Real<T>::from_u8((w >> (8 * i)) & 0xff) - (T)1) * (T)(1.0 / 3.0);
        }
        st4(dst + x, v);
    }
}

// Several composite rows at once (at most 6): all global loads of a thread are issued back to back before any
// conversion, so the DRAM latency is paid once per CTA instead of once per row.  dst(k) -> shared row of index k in
// [0, nrows), src_row(k) -> buffer row number.
template <typename T, class Dst, class SrcRow>
__device__ __forceinline__ void load_comp_rows(const IoArgs<T> &io, int fidx, int nrows, int Wc, Dst dst, SrcRow src_row) {
    constexpr int kMaxRows = 6;
    const size_t frame_base = (size_t)fidx * io.nrows;
    for (int x = 4 * threadIdx.x; x < Wc; x += 4 * blockDim.x) {
        uint32_t w[kMaxRows];
        T vf[kMaxRows][4];
#pragma unroll
        for (int k = 0; k < kMaxRows; ++k) {
            if (k < nrows) {
                const size_t off = (frame_base + src_row(k)) * Wc + x;
                if (io.in_f) ld4(io.in_f + off, vf[k]);
                else w[k] = __ldg(reinterpret_cast<const uint32_t *>(io.in_u8 + off));
            }
        }
#pragma unroll
        for (int k = 0; k < kMaxRows; ++k) {
            if (k < nrows) {
                if (!io.in_f) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        vf[k][i] = ((T)5 * Real<T>::from_u8((w[k] >> (8 * i)) & 0xff) - (T)1) * (T)(1.0 / 3.0);
                }
                st4(dst(k) + x, vf[k]);
            }
        }
    }
}

// 4 composite samples out: float (before the level map) and/or u8 (0.6 v + 0.2, image.py:16-21)
template <typename T>
__device__ __forceinline__ void store_comp4(const IoArgs<T> &io, size_t off, const T v[4]) {
    if (io.out_f) st4(io.out_f + off, v);
    if (io.out_u8) {
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) w |= to_u8((T)0.6 * v[i] + (T)0.2) << (8 * i);
        *reinterpret_cast<uint32_t *>(io.out_u8 + off) = w;
    }
}

template <typename T>
__device__ __forceinline__ void store_yuv4(const DevParams<T> &p, const IoArgs<T> &io, int fidx, int row, int x0,
                                           const T y[4], const T c1[4], const T c2[4]);
template <typename T>
__device__ __forceinline__ void store_rgb4_direct(const DevParams<T> &p, const IoArgs<T> &io, int fidx, int row, int x0,
                                                  const T y[4], const T c1[4], const T c2[4]);

// 4 decoded pixels out (inverse colour matrix, e.g. ntsc.py:36-41), float before clipping and/or u8; or, when io.yuv is
// set, the planes to the finishing-pass scratch
template <typename T>
__device__ __forceinline__ void store_rgb4(const DevParams<T> &p, const IoArgs<T> &io, int fidx, int row, int x0,
                                           const T y[4], const T c1[4], const T c2[4]) {
    if (io.yuv) {
        store_yuv4(p, io, fidx, row, x0, y, c1, c2);
        return;
    }
    store_rgb4_direct(p, io, fidx, row, x0, y, c1, c2);
}

// the planes of 4 pixels to the finishing-pass scratch yuv[frame][row][3][Wo] (comb decoders with non-default knobs)
template <typename T>
__device__ __forceinline__ void store_yuv4(const DevParams<T> &p, const IoArgs<T> &io, int fidx, int row, int x0,
                                           const T y[4], const T c1[4], const T c2[4]) {
    T *dst = io.yuv + ((size_t)fidx * io.nrows + row) * 3 * p.Wo + x0;
    st4(dst, y);
    st4(dst + p.Wo, c1);
    st4(dst + 2 * p.Wo, c2);
}

template <typename T>
__device__ __forceinline__ void store_rgb4_direct(const DevParams<T> &p, const IoArgs<T> &io, int fidx, int row, int x0,
                                                  const T y[4], const T c1[4], const T c2[4]) {
    T v[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[3 * i] = p.dec[0] * y[i] + p.dec[1] * c1[i] + p.dec[2] * c2[i];
        v[3 * i + 1] = p.dec[3] * y[i] + p.dec[4] * c1[i] + p.dec[5] * c2[i];
        v[3 * i + 2] = p.dec[6] * y[i] + p.dec[7] * c1[i] + p.dec[8] * c2[i];
    }
    if (p.ident_dec) {          // demodulate_components: the planes themselves (a product with 0 would eat the sign of -0.0)
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[3 * i] = y[i]; v[3 * i + 1] = c1[i]; v[3 * i + 2] = c2[i]; }
    }
    const size_t o = (((size_t)fidx * io.nrows + row) * p.Wo + x0) * 3;
    if (io.out_f) {
#pragma unroll
        for (int q = 0; q < 3; ++q) st4(io.out_f + o + 4 * q, v + 4 * q);
    }
    if (io.out_u8) {
        uint32_t w[3] = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < 12; ++i) w[i >> 2] |= to_u8(v[i]) << (8 * (i & 3));
        uint32_t *dst = reinterpret_cast<uint32_t *>(io.out_u8 + o);
        dst[0] = w[0];
        dst[1] = w[1];
        dst[2] = w[2];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Line-sequential colour systems (SECAM, proto-SECAM, MAC): every row carries ONE colour-difference signal X and
// the decoder pairs it with the X of the previous row of the field (zeros at the field top), e.g. secam.py:297-300.
// The heavy per-row kernel writes (luma, X) of each row once to a float scratch in HBM (aux: [frame][row][2][Wo]);
// this light kernel pairs neighbouring rows, applies the inverse matrix and stores RGB.  Compared with recomputing
// the previous row as a halo inside the heavy kernel this trades 20 B/pixel of HBM traffic (the kernels use < 3 %
// of the HBM bandwidth) for 25-100 % less arithmetic.
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_pair_rows_store(const __grid_constant__ DevParams<T> p,
                                                         const __grid_constant__ IoArgs<T> io) {
    const int Wo = p.Wo, q_per_row = Wo >> 2;
    const int f = blockIdx.z, row = io.out_begin + blockIdx.y;
    const long long frame = io.first_frame + f;
    const bool alt = is_alternate(p, frame, io.y0 + row);
    const bool hp = row >= 2;
    const T *cur = io.aux + ((size_t)f * io.nrows + row) * 2 * Wo;
    const T *prev = io.aux + ((size_t)f * io.nrows + (hp ? row - 2 : row)) * 2 * Wo;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < q_per_row; q += gridDim.x * blockDim.x) {
        T y[4], a[4], b[4] = {(T)0, (T)0, (T)0, (T)0};
        ld4(cur + 4 * q, y);
        ld4(cur + Wo + 4 * q, a);
        if (hp) ld4(prev + Wo + 4 * q, b);
        // non-alternate rows carry D'R (dr = current, db = previous), alternate rows D'B
        if (alt) store_rgb4(p, io, f, row, 4 * q, y, b, a);
        else store_rgb4(p, io, f, row, 4 * q, y, a, b);
    }
}

// u8 composite row held in registers between its (early) global load and its staging into shared memory: the DRAM
// latency of the row a CTA works on next is hidden behind the row it is working on now.
struct RowPrefetch {
    static constexpr int kMaxQuads = 4;          // 4 pixels per quad: rows of up to 16 * blockDim samples
    uint32_t w[kMaxQuads];
    template <typename T>
    __device__ __forceinline__ void fetch(const IoArgs<T> &io, int f, int row, int Wc) {
        const uint8_t *src = io.in_u8 + ((size_t)f * io.nrows + row) * Wc;
#pragma unroll
        for (int q = 0; q < kMaxQuads; ++q) {
            const int x = 4 * (threadIdx.x + q * blockDim.x);
            if (x < Wc) w[q] = __ldg(reinterpret_cast<const uint32_t *>(src + x));
        }
    }
    template <typename T>
    __device__ __forceinline__ void stage(T *dst, int Wc) const {
#pragma unroll
        for (int q = 0; q < kMaxQuads; ++q) {
            const int x = 4 * (threadIdx.x + q * blockDim.x);
            if (x < Wc) {
                T v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = ((T)5 * Real<T>::from_u8((w[q] >> (8 * i)) & 0xff) - (T)1) * (T)(1.0 / 3.0);
                st4(dst + x, v);
            }
        }
    }
};

// sin/cos of the subcarrier at 4 consecutive 1x samples starting at x0 (exact seed + 3 rotations)
template <typename T>
__device__ __forceinline__ void carrier4(unsigned long long ph_x0, T rs, T rc, T s[4], T c[4]) {
    Real<T>::sincos_turns(ph_x0, s[0], c[0]);
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        s[i] = Real<T>::fma_(s[i - 1], rc, c[i - 1] * rs);
        c[i] = Real<T>::fma_(c[i - 1], rc, -(s[i - 1] * rs));
    }
}

// the same with the seed from the hardware sin / cos approximation (Real<T>::sincos_turns_fast): the u8 row encoders
template <typename T>
__device__ __forceinline__ void carrier4_fast(unsigned long long ph_x0, T rs, T rc, T s[4], T c[4]) {
    Real<T>::sincos_turns_fast(ph_x0, s[0], c[0]);
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        s[i] = Real<T>::fma_(s[i - 1], rc, c[i - 1] * rs);
        c[i] = Real<T>::fma_(c[i - 1], rc, -(s[i - 1] * rs));
    }
}
