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

// post(j, v) for the n_out samples of resample(src line of n_in samples, `seg` = its padded segment) with resampler slot r
// (identity when the ratio is 1).  tabs: the polyphase tables staged in shared memory.
template <typename T, int MSK, class Post>
__device__ __forceinline__ void mac_fit(const DevParams<T> &p, const T *tabs, int r, const T *seg, int n_in, int n_out, Post post) {
    const ResHdr rh = p.res[r];
    const int fp = p.mac_fp;
    if (rh.ntaps == 0) {
        for (int j = threadIdx.x; j < n_out; j += blockDim.x) post(j, seg[poly_skew(fp + j, MSK)]);
    } else if (p.poly[r].up) {
        fir_poly(seg, n_out, p.poly[r], tabs + p.poly[r].off, threadIdx.x, blockDim.x, post);
    } else {        // ratios with up > 4 (odd composite widths): one output per thread straight from the dense taps
        fir_general<T>([&](int i) { return seg[poly_skew(fp + i, MSK)]; }, n_in, n_out, rh, p.taps + rh.off, threadIdx.x,
                       blockDim.x, post);
    }
}

// Shared-memory floats of one row of the kernels below (the launchers size the dynamic shared memory with the same rule)
__host__ __device__ __forceinline__ size_t mac_encode_row_elems(int segW, int seg1080) {
    return 2 * (size_t)segW + 16 + seg1080 + (segW >= 1080 ? 0 : 1080);        // the output row reuses the luma segment when it fits
}
__host__ __device__ __forceinline__ size_t mac_decode_row_elems(int segC) { return (size_t)segC + 16 + 720 + 360 + 720; }

// Encode.  smem: tables + R * ( luma seg(W) | chroma seg(W) | side[16] | line seg(1080) | out[1080] unless it fits the luma seg )
// The two input resamplers write the 1080-sample multiplex (mac.py:58-69) directly: luma[11..709] -> line[372..1070],
// chroma[5..355] -> line[18..368]; the twelve cross-faded samples around them come from side[] in a fix-up step.
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
    const size_t per_row = mac_encode_row_elems(segW, seg1080);
    const int fp = p.mac_fp;
    for (int i = threadIdx.x; i < taps_len; i += blockDim.x) taps[i] = p.ptab[i];
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const int nrow = (row + 2 < io.nrows) ? row + 2 : row;
        const int ci = is_alternate(p, g.frame, io.y0 + row) ? 6 : 3;      // D'B on alternate lines, else D'R
        T *yseg = rows + k * per_row, *cseg = yseg + segW;
        mac_zero_pads<T, MSK>(p, yseg, W);
        mac_zero_pads<T, MSK>(p, cseg, W);
        mac_zero_pads<T, MSK>(p, cseg + segW + 16, 1080);
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
            st4(yseg + poly_skew(fp + x, MSK), y);
            st4(cseg + poly_skew(fp + x, MSK), c);
        }
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {
        T *yseg = rows + k * per_row, *cseg = yseg + segW, *side = cseg + segW, *oseg = side + 16;
        auto put_l = [&](int j, T v) {                                       // luma[8..10] -> side[0..2], luma[710..712] -> side[3..5]
            if (j >= 11 && j < 710) oseg[poly_skew(fp + j + 361, MSK)] = v;
            else if (j >= 8 && j < 11) side[j - 8] = v;
            else if (j >= 710 && j < 713) side[j - 707] = v;
        };
        auto put_c = [&](int j, T v) {                                       // mac.py:57 chroma += 0.5;  [2..4] -> side[6..8], [356..358] -> side[9..11]
            v += (T)0.5;
            if (j >= 5 && j < 356) oseg[poly_skew(fp + j + 13, MSK)] = v;
            else if (j >= 2 && j < 5) side[j + 4] = v;
            else if (j >= 356 && j < 359) side[j - 347] = v;
        };
        if (mc.ok_luma)
            fir_poly_ct<T, 3, MacShape::KL, MSK != 0, true>(yseg, 713, p.poly[MR_LUMA_IN], reinterpret_cast<const T (&)[1][3][MacShape::KL]>(mc.luma),
                                                            threadIdx.x, blockDim.x, put_l);
        else mac_fit<T, MSK>(p, taps, MR_LUMA_IN, yseg, W, 713, put_l);
        if (mc.ok_chroma)
            fir_poly_ct<T, 3, MacShape::KC, MSK != 0, true>(cseg, 359, p.poly[MR_CHROMA_IN], reinterpret_cast<const T (&)[1][3][MacShape::KC]>(mc.chroma),
                                                            threadIdx.x, blockDim.x, put_c);
        else mac_fit<T, MSK>(p, taps, MR_CHROMA_IN, cseg, W, 359, put_c);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 30 * g.count; t += blockDim.x) {          // mac.py:58-69: the constant and cross-faded samples
        const int k = t / 30, e = t - 30 * k;
        const T *side = rows + k * per_row + 2 * segW;
        T *oseg = rows + k * per_row + 2 * segW + 16;
        const T *l = side, *c = side + 6, *c2 = side + 9;                  // l[8..10 | 710..712], c[2..4], c[356..358]
        int i;
        T v = (T)0.5;
        if (e < 15) i = e;
        else if (e < 21) i = 1074 + (e - 15);
        else {
            i = e < 24 ? e - 6 : e < 27 ? e + 345 : e + 1044;               // 15..17, 369..371, 1071..1073
            switch (e) {
            case 21: v = (T)0.4375 + (T)0.125 * c[0]; break;
            case 22: v = (T)0.25 + (T)0.5 * c[1]; break;
            case 23: v = (T)0.0625 + (T)0.875 * c[2]; break;
            case 24: v = (T)0.875 * c2[0] + (T)0.125 * l[0]; break;
            case 25: v = (T)0.5 * c2[1] + (T)0.5 * l[1]; break;
            case 26: v = (T)0.125 * c2[2] + (T)0.875 * l[2]; break;
            case 27: v = (T)0.0625 + (T)0.875 * l[3]; break;
            case 28: v = (T)0.25 + (T)0.5 * l[4]; break;
            default: v = (T)0.4375 + (T)0.125 * l[5]; break;
            }
        }
        oseg[poly_skew(fp + i, MSK)] = v;
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {
        const T *oseg = rows + k * per_row + 2 * segW + 16;
        T *outrow = segW >= 1080 ? rows + k * per_row : rows + k * per_row + 2 * segW + 16 + seg1080;
        auto put = [&](int j, T v) { outrow[j] = v; };
        if (mc.ok_out)
            fir_poly_ct<T, 2, MacShape::KO, MSK != 0, false>(oseg, Wc, p.poly[MR_OUT], mc.out, threadIdx.x, blockDim.x, put);
        else mac_fit<T, MSK>(p, taps, MR_OUT, oseg, 1080, Wc, put);
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const T *outrow = segW >= 1080 ? rows + k * per_row : rows + k * per_row + 2 * segW + 16 + seg1080;
        for (int q = threadIdx.x; q < (Wc >> 2); q += blockDim.x) {
            T o[4];
            ld4(outrow + 4 * q, o);
            store_comp4(io, ((size_t)g.fidx * io.nrows + row) * Wc + 4 * q, o);
        }
    }
}

// Decode.  smem: tables + (R+1) rows x ( comp seg(Wc) | side[16] | luma720 | ch360 | XE[360] | XO[360] )
// The composite resampler writes the demultiplexed line directly (mac.py:86-109): line[372..1070] -> luma[11..709],
// line[18..368] -> chroma[5..355]; the thirty extrapolated / cross-faded samples come from side[] in a fix-up step.  Of the
// row ahead of the group (the other chroma component) only the chroma part of the line is resampled.
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
    const size_t per_row = mac_decode_row_elems(segC);
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
        T *side = rowp(k) + segC, *l = side + 16, *ch = l + 720;
        auto put = [&](int j, T v) {          // side: line[15..17] -> [0..2], line[368..372] -> [3..7], line[1071..1073] -> [8..10]
            if (j >= 372 && j < 1071) l[j - 361] = v;
            else if (j >= 18 && j < 369) ch[j - 13] = v;
            if (j >= 368 && j < 373) side[j - 365] = v;
            else if (j >= 15 && j < 18) side[j - 15] = v;
            else if (j >= 1071 && j < 1074) side[j - 1063] = v;
        };
        const int n = k < 0 ? 373 : 1074;
        if (mc.ok_comp) fir_poly_ct<T, 3, MacShape::KI, MSK != 0, false>(rowp(k), n, p.poly[MR_COMP_IN], mc.comp, threadIdx.x, blockDim.x, put);
        else mac_fit<T, MSK>(p, taps, MR_COMP_IN, rowp(k), Wc, n, put);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 30 * (g.count - k_lo); t += blockDim.x) {       // mac.py:86-109: the samples around the cross-fades
        const int k = k_lo + t / 30, e = t % 30;
        const T *c = rowp(k) + segC;                                        // c[0..2] = line[15..17], c[3..7] = line[368..372], c[8..10] = line[1071..1073]
        T *l = rowp(k) + segC + 16, *ch = l + 720;
        const T ch355 = c[3], l11 = c[7];                                   // chroma[355] = line[368], luma[11] = line[372]
        if (e < 21) {
            const int i = e < 11 ? e : 699 + e;                             // luma 0..10, 710..719
            T v;
            if (i < 9) v = (T)8 * c[4] - (T)7 * ch355;
            else if (i == 9) v = (T)2 * c[5] - ch355;
            else if (i == 10) v = (c[6] - (T)0.125 * ch355) / (T)0.875;
            else if (i == 710) v = (c[8] - (T)0.0625) / (T)0.875;
            else if (i == 711) v = (T)2 * c[9] - (T)0.5;
            else v = (T)8 * c[10] - (T)3.5;
            l[i] = v;
        } else {
            const int m = e < 26 ? e - 21 : 330 + e;                        // chroma 0..4, 356..359
            T v;
            if (m == 0 || m == 2) v = (T)8 * c[0] - (T)3.5;
            else if (m == 1) v = (T)0.5;                                    // mac.py:105 writes element 0 only
            else if (m == 3) v = (T)2 * c[1] - (T)0.5;
            else if (m == 4) v = (c[2] - (T)0.0625) / (T)0.875;
            else if (m == 356) v = (c[4] - (T)0.125 * l11) / (T)0.875;
            else if (m == 357) v = (T)2 * c[5] - l11;
            else v = (T)8 * c[6] - (T)7 * l11;                              // 358 and 359
            ch[m] = v;
        }
    }
    __syncthreads();
    const FirTaps<T> hup{p.taps + p.res[MR_UP2].off, nullptr};
    for (int k = k_lo; k < g.count; ++k) {
        T *ch = rowp(k) + segC + 16 + 720, *xe = ch + 360, *xo = xe + 360;
        fir_up2(xe, xo, ch, 360, hup, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const bool alt = is_alternate(p, g.frame, io.y0 + row);
        const bool hp = (k > 0) || has_prev0;
        const T *l = rowp(k) + segC + 16, *xe = l + 720 + 360, *xo = xe + 360;
        const T *pe = hp ? rowp(k - 1) + segC + 16 + 720 + 360 : nullptr, *po = hp ? pe + 360 : nullptr;
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
