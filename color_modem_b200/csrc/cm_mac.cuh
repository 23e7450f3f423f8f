// D2-MAC kernels (time-compressed chroma + luma multiplex).  Reference: color_modem/color/mac.py.
// Pure resampling and assembly: no IIR, no carrier.  Lines are resampled with the general polyphase FIR
// (ratios 720/W, 360/W, width/1080, 1080/width; e.g. 3/8 -> 161 taps, 3/16 -> 321 taps, 2/3, 3/2).
#pragma once
#include "cm_common.cuh"
#include "cm_fir.cuh"
#include "cm_io.cuh"
#include "cm_slots.h"

// Shared-memory line that a resampler reads: [FP zeros | n samples (rounded up to 4) | BP zeros]  (cm_fir.cuh: fir_poly)
// in the skewed layout of cm_fir.cuh (poly_skew): sample i of the line lives at seg[poly_skew(FP + i, MSK)]
// (MSK: the skew mask as a compile-time constant — the kernels are instantiated with and without the skew)
template <typename T, int MSK>
__device__ __forceinline__ int mac_seg(const DevParams<T> &p, int n) {
    return poly_skew(p.mac_fp + ((n + 3) & ~3) + p.mac_bp, MSK) + 4;
}

template <typename T, int MSK>
__device__ __forceinline__ void mac_zero_pads(const DevParams<T> &p, T *seg, int n) {
    const int n4 = (n + 3) & ~3;
    for (int i = threadIdx.x; i < p.mac_fp; i += blockDim.x) seg[poly_skew(i, MSK)] = (T)0;
    for (int i = n + threadIdx.x; i < n4 + p.mac_bp; i += blockDim.x) seg[poly_skew(p.mac_fp + i, MSK)] = (T)0;
}

// dst[0..n_out) = resample(src line of n_in samples, `seg` = its padded segment) with resampler slot r (identity when the
// ratio is 1).  tabs: the polyphase tables staged in shared memory.
template <typename T, int MSK>
__device__ __forceinline__ void mac_fit(const DevParams<T> &p, const T *tabs, int r, const T *seg, int n_in, T *dst,
                                        int n_out, T add) {
    const ResHdr rh = p.res[r];
    const int fp = p.mac_fp;
    if (rh.ntaps == 0) {
        for (int j = threadIdx.x; j < n_out; j += blockDim.x) dst[j] = seg[poly_skew(fp + j, MSK)] + add;
    } else if (p.poly[r].up) {
        fir_poly(seg, n_out, p.poly[r], tabs + p.poly[r].off, threadIdx.x, blockDim.x, [&](int j, T v) { dst[j] = v + add; });
    } else {        // ratios with up > 4 (odd composite widths): one output per thread straight from the dense taps
        fir_general<T>([&](int i) { return seg[poly_skew(fp + i, MSK)]; }, n_in, n_out, rh, p.taps + rh.off, threadIdx.x,
                       blockDim.x, [&](int j, T v) { dst[j] = v + add; });
    }
}

// Encode.  smem: tables + R * ( luma seg(W) | chroma seg(W) | luma720 | ch360 | line seg(1080) | out[1080] )
template <typename T, int MSK>
__global__ void __launch_bounds__(CM_NTHREADS)
k_mac_encode(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io, int taps_len,
             const __grid_constant__ MacConst<T> mc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, W4 = W >> 2, Wc = p.Wc;
    const bool avg = (p.flags & 2) != 0;
    T *taps = sm;
    T *rows = sm + taps_len;
    const int segW = mac_seg<T, MSK>(p, W), seg1080 = mac_seg<T, MSK>(p, 1080);
    const size_t per_row = 2 * (size_t)segW + 720 + 360 + seg1080 + 1080;
    for (int i = threadIdx.x; i < taps_len; i += blockDim.x) taps[i] = p.ptab[i];
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const int nrow = (row + 2 < io.nrows) ? row + 2 : row;
        const int ci = is_alternate(p, g.frame, io.y0 + row) ? 6 : 3;      // D'B on alternate lines, else D'R
        T *yseg = rows + k * per_row, *cseg = yseg + segW;
        mac_zero_pads<T, MSK>(p, yseg, W);
        mac_zero_pads<T, MSK>(p, cseg, W);
        mac_zero_pads<T, MSK>(p, cseg + segW + 720 + 360, 1080);
        for (int q = threadIdx.x; q < W4; q += blockDim.x) {
            const int x = 4 * q;
            T r[4], gg[4], b[4], y[4], c[4];
            load_rgb4(io, ((size_t)g.fidx * io.nrows + row) * W + x, r, gg, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                y[i] = p.enc[0] * r[i] + p.enc[1] * gg[i] + p.enc[2] * b[i];
                c[i] = p.enc[ci] * r[i] + p.enc[ci + 1] * gg[i] + p.enc[ci + 2] * b[i];
            }
            if (avg) {
                load_rgb4(io, ((size_t)g.fidx * io.nrows + nrow) * W + x, r, gg, b);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    c[i] = (T)0.5 * ((p.enc[ci] * r[i] + p.enc[ci + 1] * gg[i] + p.enc[ci + 2] * b[i]) + c[i]);
            }
            st4(yseg + poly_skew(p.mac_fp + x, MSK), y);
            st4(cseg + poly_skew(p.mac_fp + x, MSK), c);
        }
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {
        T *yseg = rows + k * per_row, *cseg = yseg + segW, *l720 = cseg + segW, *c360 = l720 + 720;
        if (mc.ok_luma)
            fir_poly_ct<T, 3, MacShape::KL, MSK != 0, true>(yseg, 720, p.poly[MR_LUMA_IN], reinterpret_cast<const T (&)[1][3][MacShape::KL]>(mc.luma),
                                                            threadIdx.x, blockDim.x, [&](int j, T v) { l720[j] = v; });
        else mac_fit<T, MSK>(p, taps, MR_LUMA_IN, yseg, W, l720, 720, (T)0);
        if (mc.ok_chroma)                                                       // mac.py:57 chroma += 0.5
            fir_poly_ct<T, 3, MacShape::KC, MSK != 0, true>(cseg, 360, p.poly[MR_CHROMA_IN], reinterpret_cast<const T (&)[1][3][MacShape::KC]>(mc.chroma),
                                                            threadIdx.x, blockDim.x, [&](int j, T v) { c360[j] = v + (T)0.5; });
        else mac_fit<T, MSK>(p, taps, MR_CHROMA_IN, cseg, W, c360, 360, (T)0.5);
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {                                    // mac.py:58-69
        const T *l = rows + k * per_row + 2 * segW, *c = l + 720;
        T *oseg = rows + k * per_row + 2 * segW + 1080;
        for (int i = threadIdx.x; i < 1080; i += blockDim.x) {
            T v = (T)0.5;
            if (i == 15) v = (T)0.4375 + (T)0.125 * c[2];
            else if (i == 16) v = (T)0.25 + (T)0.5 * c[3];
            else if (i == 17) v = (T)0.0625 + (T)0.875 * c[4];
            else if (i >= 18 && i < 369) v = c[i - 13];
            else if (i == 369) v = (T)0.875 * c[356] + (T)0.125 * l[8];
            else if (i == 370) v = (T)0.5 * c[357] + (T)0.5 * l[9];
            else if (i == 371) v = (T)0.125 * c[358] + (T)0.875 * l[10];
            else if (i >= 372 && i < 1071) v = l[i - 361];
            else if (i == 1071) v = (T)0.0625 + (T)0.875 * l[710];
            else if (i == 1072) v = (T)0.25 + (T)0.5 * l[711];
            else if (i == 1073) v = (T)0.4375 + (T)0.125 * l[712];
            oseg[poly_skew(p.mac_fp + i, MSK)] = v;
        }
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {
        const T *oseg = rows + k * per_row + 2 * segW + 1080;
        T *outrow = rows + k * per_row + 2 * segW + 1080 + seg1080;
        if (mc.ok_out)
            fir_poly_ct<T, 2, MacShape::KO, MSK != 0, false>(oseg, Wc, p.poly[MR_OUT], mc.out, threadIdx.x, blockDim.x,
                                                             [&](int j, T v) { outrow[j] = v; });
        else mac_fit<T, MSK>(p, taps, MR_OUT, oseg, 1080, outrow, Wc, (T)0);
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const T *outrow = rows + k * per_row + 2 * segW + 1080 + seg1080;
        for (int q = threadIdx.x; q < (Wc >> 2); q += blockDim.x) {
            T o[4];
            ld4(outrow + 4 * q, o);
            store_comp4(io, ((size_t)g.fidx * io.nrows + row) * Wc + 4 * q, o);
        }
    }
}

// Decode.  smem: tables + (R+1) rows x ( comp seg(Wc) | c1080 | luma720 | ch360 | XE[360] | XO[360] )
template <typename T, int MSK>
__global__ void __launch_bounds__(CM_NTHREADS)
k_mac_decode(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io, int taps_len,
             const __grid_constant__ MacConst<T> mc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int Wc = p.Wc;
    T *taps = sm;
    T *rows = sm + taps_len;
    const int segC = mac_seg<T, MSK>(p, Wc);
    const size_t per_row = (size_t)segC + 1080 + 720 + 360 + 720;
    const bool has_prev0 = g.r0 >= 2;
    const int k_lo = has_prev0 ? -1 : 0;
    auto rowp = [&](int k) { return rows + (size_t)(k - k_lo) * per_row; };
    for (int i = threadIdx.x; i < taps_len; i += blockDim.x) taps[i] = p.ptab[i];
    for (int k = k_lo; k < g.count; ++k) {
        mac_zero_pads<T, MSK>(p, rowp(k), Wc);
        T *cseg = rowp(k);
        const size_t base = ((size_t)g.fidx * io.nrows + g.r0 + 2 * k) * Wc;
        for (int x = 4 * threadIdx.x; x < Wc; x += 4 * blockDim.x) {       // load_comp_row into the skewed segment
            T v[4];
            if (io.in_f) {
                ld4(io.in_f + base + x, v);
            } else {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(io.in_u8 + base + x));
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = ((T)5 * Real<T>::from_u8((w >> (8 * i)) & 0xff) - (T)1) * (T)(1.0 / 3.0);
            }
            st4(cseg + poly_skew(p.mac_fp + x, MSK), v);
        }
    }
    __syncthreads();
    for (int k = k_lo; k < g.count; ++k) {
        T *c1080 = rowp(k) + segC;
        if (mc.ok_comp)
            fir_poly_ct<T, 3, MacShape::KI, MSK != 0, false>(rowp(k), 1080, p.poly[MR_COMP_IN], mc.comp, threadIdx.x, blockDim.x,
                                                             [&](int j, T v) { c1080[j] = v; });
        else mac_fit<T, MSK>(p, taps, MR_COMP_IN, rowp(k), Wc, c1080, 1080, (T)0);
    }
    __syncthreads();
    for (int k = k_lo; k < g.count; ++k) {                                 // mac.py:86-109
        const T *c = rowp(k) + segC;
        T *l = rowp(k) + segC + 1080, *ch = l + 720;
        for (int i = threadIdx.x; i < 720 + 360; i += blockDim.x) {
            if (i < 720) {
                const T ch355 = c[368];                                     // chroma[355] = composite[368]
                const T l8 = (T)8 * c[369] - (T)7 * ch355;
                const T l712 = (T)8 * c[1073] - (T)3.5;
                T v;
                if (i < 8) v = l8;
                else if (i == 8) v = l8;
                else if (i == 9) v = (T)2 * c[370] - ch355;
                else if (i == 10) v = (c[371] - (T)0.125 * ch355) / (T)0.875;
                else if (i < 710) v = c[i + 361];
                else if (i == 710) v = (c[1071] - (T)0.0625) / (T)0.875;
                else if (i == 711) v = (T)2 * c[1072] - (T)0.5;
                else v = l712;
                l[i] = v;
            } else {
                const int m = i - 720;
                const T l11 = c[372];                                       // luma[11] = composite[372]
                const T c2 = (T)8 * c[15] - (T)3.5;
                const T c358 = (T)8 * c[371] - (T)7 * l11;
                T v;
                if (m == 0) v = c2;
                else if (m == 1) v = (T)0.5;                                // mac.py:105 writes element 0 only
                else if (m == 2) v = c2;
                else if (m == 3) v = (T)2 * c[16] - (T)0.5;
                else if (m == 4) v = (c[17] - (T)0.0625) / (T)0.875;
                else if (m < 356) v = c[m + 13];
                else if (m == 356) v = (c[369] - (T)0.125 * l11) / (T)0.875;
                else if (m == 357) v = (T)2 * c[370] - l11;
                else v = c358;                                              // 358 and 359
                ch[m] = v;
            }
        }
    }
    __syncthreads();
    const FirTaps<T> hup{p.taps + p.res[MR_UP2].off, nullptr};
    for (int k = k_lo; k < g.count; ++k) {
        T *ch = rowp(k) + segC + 1080 + 720, *xe = ch + 360, *xo = xe + 360;
        fir_up2(xe, xo, ch, 360, hup, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const bool alt = is_alternate(p, g.frame, io.y0 + row);
        const bool hp = (k > 0) || has_prev0;
        const T *l = rowp(k) + segC + 1080, *xe = l + 720 + 360, *xo = xe + 360;
        const T *pe = hp ? rowp(k - 1) + segC + 1080 + 720 + 360 : nullptr, *po = hp ? pe + 360 : nullptr;
        for (int q = threadIdx.x; q < 180; q += blockDim.x) {
            T y[4], a[4], b[4];
            ld4(l + 4 * q, y);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int x = 4 * q + i, m = x >> 1;
                a[i] = ((x & 1) ? xo[m] : xe[m]) - (T)0.5;                   // mac.py:111
                b[i] = hp ? ((x & 1) ? po[m] : pe[m]) - (T)0.5 : (T)0;
            }
            // mac.py:113-118: non-alternate rows carry D'R (dr = current, db = previous)
            if (alt) store_rgb4(p, io, g.fidx, row, 4 * q, y, b, a);
            else store_rgb4(p, io, g.fidx, row, 4 * q, y, a, b);
        }
    }
}
