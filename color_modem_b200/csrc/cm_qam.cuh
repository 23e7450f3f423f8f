// QAM family kernels: NTSC / PAL encode, band-split decode, PAL-D delay-line decode, NTSC 2-line / 3-line comb,
// PAL 3-line comb.  Reference: color_modem/qam.py, color/ntsc.py, color/pal.py, comb.py.
#pragma once
#include "cm_common.cuh"
#include "cm_fir.cuh"
#include "cm_iir.cuh"
#include "cm_slots.h"

// ------------------------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void copy_taps(T *dst, const DevParams<T> &p, int nres) {
    int total = 0;
    for (int r = 0; r < nres; ++r) total = max(total, p.res[r].off + p.res[r].ntaps);
    for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = p.taps[i];
}

// composite row -> T, either from the u8 frame ((5*(v/255) - 1)/3, image.py:23-25,62) or from the float buffer
template <typename T>
__device__ __forceinline__ void load_comp_row(T *dst, const IoArgs<T> &io, int fidx, int row, int Wc) {
    const size_t base = ((size_t)fidx * io.nrows + row) * Wc;
    if (io.in_f) {
        for (int x = threadIdx.x; x < Wc; x += blockDim.x) dst[x] = io.in_f[base + x];
    } else {
        for (int x = threadIdx.x; x < Wc; x += blockDim.x)
            dst[x] = ((T)5 * Real<T>::from_u8(io.in_u8[base + x]) - (T)1) / (T)3;
    }
}

template <typename T>
__device__ __forceinline__ void store_rgb(const DevParams<T> &p, const IoArgs<T> &io, int fidx, int row, int x,
                                          T y, T c1, T c2) {
    T r = p.dec[0] * y + p.dec[1] * c1 + p.dec[2] * c2;
    T g = p.dec[3] * y + p.dec[4] * c1 + p.dec[5] * c2;
    T b = p.dec[6] * y + p.dec[7] * c1 + p.dec[8] * c2;
    const size_t o = (((size_t)fidx * io.nrows + row) * p.Wo + x) * 3;
    if (io.out_f) { io.out_f[o] = r; io.out_f[o + 1] = g; io.out_f[o + 2] = b; }
    if (io.out_u8) {
        io.out_u8[o] = (uint8_t)to_u8(r);
        io.out_u8[o + 1] = (uint8_t)to_u8(g);
        io.out_u8[o + 2] = (uint8_t)to_u8(b);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Encode: RGB -> Y + sin(phi) LP(U) + cos(phi) LP(+-V)          qam.py:20-32, ntsc.py:27-45, pal.py:32-52
// optional ColorAveragingModem front end (comb.py:141-152)
// smem: R * 3 * W
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(CM_NTHREADS)
k_qam_encode(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool avg = (p.flags & 2) != 0;

    for (int idx = threadIdx.x; idx < g.count * W; idx += blockDim.x) {
        const int k = idx / W, x = idx - k * W;
        const int row = g.r0 + 2 * k;
        T rgb[3], nrgb[3];
        const size_t o = (((size_t)g.fidx * io.nrows + row) * W + x) * 3;
        const int nrow = (row + 2 < io.nrows) ? row + 2 : row;
        const size_t on = (((size_t)g.fidx * io.nrows + nrow) * W + x) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            rgb[c] = io.in_f ? io.in_f[o + c] : Real<T>::from_u8(io.in_u8[o + c]);
            if (avg) nrgb[c] = io.in_f ? io.in_f[on + c] : Real<T>::from_u8(io.in_u8[on + c]);
        }
        T *row_sm = sm + (size_t)k * 3 * W;
        row_sm[x] = p.enc[0] * rgb[0] + p.enc[1] * rgb[1] + p.enc[2] * rgb[2];
        T u = p.enc[3] * rgb[0] + p.enc[4] * rgb[1] + p.enc[5] * rgb[2];
        T v = p.enc[6] * rgb[0] + p.enc[7] * rgb[1] + p.enc[8] * rgb[2];
        if (avg) {
            T un = p.enc[3] * nrgb[0] + p.enc[4] * nrgb[1] + p.enc[5] * nrgb[2];
            T vn = p.enc[6] * nrgb[0] + p.enc[7] * nrgb[1] + p.enc[8] * nrgb[2];
            u = (T)0.5 * (un + u);
            v = (T)0.5 * (vn + v);
        }
        row_sm[W + x] = u;
        row_sm[2 * W + x] = v;
    }
    __syncthreads();
    for (int t = warp; t < 2 * g.count; t += nwarps) {
        T *buf = sm + (size_t)(t >> 1) * 3 * W + (1 + (t & 1)) * W;
        warp_iir<T>(p.tab + p.filt[QF_PRE_LP].off, p.filt[QF_PRE_LP],
                    [&](int j) { return buf[j]; }, [&](int j, T v) { buf[j] = v; });
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < g.count * W; idx += blockDim.x) {
        const int k = idx / W, x = idx - k * W;
        const int row = g.r0 + 2 * k;
        const int line = io.y0 + row;
        const T *row_sm = sm + (size_t)k * 3 * W;
        unsigned long long ph = start_phase(p, g.frame, line) + (unsigned long long)x * p.phases[QP_STEP1X];
        T s, c;
        Real<T>::sincos_turns(ph, s, c);
        T v = row_sm[2 * W + x];
        if ((p.flags & 1) && is_alternate(p, g.frame, line)) v = -v;
        T comp = row_sm[x] + (s * row_sm[W + x] + c * v);
        const size_t o = ((size_t)g.fidx * io.nrows + row) * p.Wc + x;
        if (io.out_f) io.out_f[o] = comp;
        if (io.out_u8) io.out_u8[o] = (uint8_t)to_u8((T)0.6 * comp + (T)0.2);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Band-split decode of independent rows                       qam.py:43-58, ntsc.py:47-49, pal.py:54-59
//   luma_mode 0: luma = down2(BS(up2 c))        (strip_chroma=True; NtscModem / PalSModem; field-top rows of
//                                                NtscCombModem / PalDModem, comb.py:48-49)
//   luma_mode 1: luma = c - remod(u, v)         (field-top rows of Pal3DModem, pal.py:191-202,225-226)
//   luma_mode 2: luma = c                       (strip_chroma=False: the reset-branch return value of the comb
//                                                decoders, CM_MODE_BANDSPLIT_NOSTRIP)
// 2 warps per row.  smem per row: c[W] + 4 x [2W]
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(CM_NTHREADS)
k_qam_bandsplit(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io, int luma_mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, W2 = 2 * W;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    T *taps = sm;                       // 128 elements reserved
    T *rows = sm + 128;
    const int per_row = 9 * W;
    copy_taps(taps, p, 2);
    for (int k = 0; k < g.count; ++k) load_comp_row(rows + (size_t)k * per_row, io, g.fidx, g.r0 + 2 * k, W);
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {
        T *c = rows + (size_t)k * per_row;
        fir_up2(c + W, c, W, taps + p.res[QR_UP2].off, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    // IIR phase 1: band-pass -> b2x, band-stop -> l2x
    for (int t = warp; t < 2 * g.count; t += nwarps) {
        T *c = rows + (size_t)(t >> 1) * per_row;
        const T *a2x = c + W;
        if ((t & 1) == 0) {
            T *b2x = c + W + W2;
            warp_iir<T>(p.tab + p.filt[QF_BP2X].off, p.filt[QF_BP2X],
                        [&](int j) { return a2x[j]; }, [&](int j, T v) { b2x[j] = v; });
        } else if (luma_mode == 0) {
            T *l2x = c + W + 2 * W2;
            warp_iir<T>(p.tab + p.filt[QF_BS2X].off, p.filt[QF_BS2X],
                        [&](int j) { return a2x[j]; }, [&](int j, T v) { l2x[j] = v; });
        }
    }
    __syncthreads();
    // IIR phase 2: product demodulation + low-pass.  u2x -> a2x (c2x is dead), v2x -> 4th buffer
    for (int t = warp; t < 2 * g.count; t += nwarps) {
        const int k = t >> 1;
        T *c = rows + (size_t)k * per_row;
        const T *b2x = c + W + W2;
        T *dst = (t & 1) ? (c + W + 3 * W2) : (c + W);
        const int line = io.y0 + g.r0 + 2 * k;
        // sin for u, cos for v: cos(x) = sin(x + 1/4 turn)
        const unsigned long long ph0 = start_phase(p, g.frame, line) + p.phases[QP_BP_SHIFT] +
                                       ((t & 1) ? 0x4000000000000000ull : 0ull);
        const unsigned long long step = p.phases[QP_STEP2X];
        warp_iir<T>(p.tab + p.filt[QF_DEMOD_LP].off, p.filt[QF_DEMOD_LP],
                    [&](int j) {
                        T s, cc;
                        Real<T>::sincos_turns(ph0 + (unsigned long long)j * step, s, cc);
                        return (T)2 * s * b2x[j];
                    },
                    [&](int j, T v) { dst[j] = v; });
    }
    __syncthreads();
    // down2 of u2x, v2x (and luma) -> u, v, y at 1x into the b2x region (dead now)
    for (int k = 0; k < g.count; ++k) {
        T *c = rows + (size_t)k * per_row;
        T *uo = c + W + W2, *vo = uo + W;
        const T *h = taps + p.res[QR_DOWN2].off;
        const bool alt = (p.flags & 1) && is_alternate(p, g.frame, io.y0 + g.r0 + 2 * k);
        fir_down2(c + W, W2, h, threadIdx.x, blockDim.x, [&](int j, T v) { uo[j] = v; });
        fir_down2(c + W + 3 * W2, W2, h, threadIdx.x, blockDim.x, [&](int j, T v) { vo[j] = alt ? -v : v; });
    }
    __syncthreads();
    if (luma_mode == 0) {
        for (int k = 0; k < g.count; ++k) {
            T *c = rows + (size_t)k * per_row;
            const T *uo = c + W + W2, *vo = uo + W;
            const int row = g.r0 + 2 * k;
            fir_down2(c + W + 2 * W2, W2, taps + p.res[QR_DOWN2].off, threadIdx.x, blockDim.x,
                      [&](int j, T y) { store_rgb(p, io, g.fidx, row, j, y, uo[j], vo[j]); });
        }
        return;
    }
    if (luma_mode == 2) {
        for (int idx = threadIdx.x; idx < g.count * W; idx += blockDim.x) {
            const int k = idx / W, x = idx - k * W;
            const T *c = rows + (size_t)k * per_row;
            store_rgb(p, io, g.fidx, g.r0 + 2 * k, x, c[x], c[W + W2 + x], c[W + W2 + W + x]);
        }
        return;
    }
    // luma_mode 1: re-modulate (u, v) through the encoder's pre-lowpass and subtract from the composite
    for (int t = warp; t < 2 * g.count; t += nwarps) {
        T *c = rows + (size_t)(t >> 1) * per_row;
        const T *src = c + W + W2 + (t & 1) * W;
        T *dst = c + W + (t & 1) * W;           // a2x region (u2x is dead after down2)
        warp_iir<T>(p.tab + p.filt[QF_PRE_LP].off, p.filt[QF_PRE_LP],
                    [&](int j) { return src[j]; }, [&](int j, T v) { dst[j] = v; });
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < g.count * W; idx += blockDim.x) {
        const int k = idx / W, x = idx - k * W;
        const T *c = rows + (size_t)k * per_row;
        const int row = g.r0 + 2 * k, line = io.y0 + row;
        T s, cc;
        Real<T>::sincos_turns(start_phase(p, g.frame, line) + (unsigned long long)x * p.phases[QP_STEP1X], s, cc);
        T vl = c[W + W + x];
        if ((p.flags & 1) && is_alternate(p, g.frame, line)) vl = -vl;
        T y = c[x] - (s * c[W + x] + cc * vl);
        store_rgb(p, io, g.fidx, row, x, y, c[W + W2 + x], c[W + W2 + W + x]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// PAL-D delay-line decode of rows that have a predecessor        pal.py:79-127 + comb.py:50-53
// Per row k (k = -1 is the halo row y0-2):  G[k] = up2(down2(BP(up2 c[k])))       (extract_chroma, then the
// up2 of _demodulate_am; all linear, so G of the sum/difference is the sum/difference of the G's)
//   S = down2(LP(sin(ph)       * (G[k] + G[k-1])))      D = down2(LP(sin(ph + pi/2) * (G[k] - G[k-1])))
//   u = D sin(LS/2) + S cos(LS/2);  v = D cos(LS/2) - S sin(LS/2);  v = -v on alternate lines
//   y = c - (sin(phi) LP_pre(u) + cos(phi) LP_pre(+-v))
// smem: taps[128] + (R+1) * (c[W] + G[2W]) + R * 2 * [2W] work
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(CM_NTHREADS)
k_pald_combed(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, W2 = 2 * W;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int R = io.rows_per_cta;
    T *taps = sm;
    T *cbuf = sm + 128;                         // (R+1) x W     composite rows, index k+1
    T *gbuf = cbuf + (size_t)(R + 1) * W;       // (R+1) x 2W    G rows, index k+1
    T *work = gbuf + (size_t)(R + 1) * W2;      // 2R x 2W
    const int nin = g.count + 1;
    copy_taps(taps, p, 2);
    for (int k = 0; k < nin; ++k) load_comp_row(cbuf + (size_t)k * W, io, g.fidx, g.r0 + 2 * (k - 1), W);
    __syncthreads();
    for (int k = 0; k < nin; ++k)
        fir_up2(gbuf + (size_t)k * W2, cbuf + (size_t)k * W, W, taps + p.res[QR_UP2].off, threadIdx.x, blockDim.x);
    __syncthreads();
    for (int t = warp; t < nin; t += nwarps) {          // band-pass in place
        T *b = gbuf + (size_t)t * W2;
        warp_iir<T>(p.tab + p.filt[QF_BP2X].off, p.filt[QF_BP2X],
                    [&](int j) { return b[j]; }, [&](int j, T v) { b[j] = v; });
    }
    __syncthreads();
    for (int k = 0; k < nin; ++k) {                     // E = down2(b2x) -> work[k][0..W)
        T *e = work + (size_t)k * W;
        fir_down2(gbuf + (size_t)k * W2, W2, taps + p.res[QR_DOWN2].off, threadIdx.x, blockDim.x,
                  [&](int j, T v) { e[j] = v; });
    }
    __syncthreads();
    for (int k = 0; k < nin; ++k)                       // G = up2(E)
        fir_up2(gbuf + (size_t)k * W2, work + (size_t)k * W, W, taps + p.res[QR_UP2].off, threadIdx.x, blockDim.x);
    __syncthreads();
    for (int t = warp; t < 2 * g.count; t += nwarps) {  // AM demodulation low-pass of sum / difference
        const int k = t >> 1;
        const T *gc = gbuf + (size_t)(k + 1) * W2, *gl = gbuf + (size_t)k * W2;
        T *dst = work + (size_t)t * W2;
        const int line = io.y0 + g.r0 + 2 * k;
        const T sgn = (t & 1) ? (T)-1 : (T)1;
        const unsigned long long ph0 = start_phase(p, g.frame, line) + p.phases[QP_BP_SHIFT] - p.phases[QP_HALF_LS] +
                                       ((t & 1) ? 0x4000000000000000ull : 0ull);
        const unsigned long long step = p.phases[QP_STEP2X];
        warp_iir<T>(p.tab + p.filt[QF_PALD_LP].off, p.filt[QF_PALD_LP],
                    [&](int j) {
                        T s, cc;
                        Real<T>::sincos_turns(ph0 + (unsigned long long)j * step, s, cc);
                        return (gc[j] + sgn * gl[j]) * s;
                    },
                    [&](int j, T v) { dst[j] = v; });
    }
    __syncthreads();
    // S, D at 1x -> u, v (kept in gbuf rows: u at [k][0..W), v at [k][W..2W))
    const T sf = p.scalars[QS_PALD_SIN], cf = p.scalars[QS_PALD_COS];
    for (int k = 0; k < g.count; ++k) {
        T *so = gbuf + (size_t)k * W2, *dout = so + W;
        const T *h = taps + p.res[QR_DOWN2].off;
        fir_down2(work + (size_t)(2 * k) * W2, W2, h, threadIdx.x, blockDim.x, [&](int j, T v) { so[j] = v; });
        fir_down2(work + (size_t)(2 * k + 1) * W2, W2, h, threadIdx.x, blockDim.x, [&](int j, T v) { dout[j] = v; });
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < g.count * W; idx += blockDim.x) {
        const int k = idx / W, x = idx - k * W;
        T *so = gbuf + (size_t)k * W2;
        const T s = so[x], d = so[W + x];
        T u = d * sf + s * cf;
        T v = d * cf - s * sf;
        if (is_alternate(p, g.frame, io.y0 + g.r0 + 2 * k)) v = -v;
        so[x] = u;
        so[W + x] = v;
    }
    __syncthreads();
    for (int t = warp; t < 2 * g.count; t += nwarps) {  // encoder pre-lowpass for the re-modulation
        const T *src = gbuf + (size_t)(t >> 1) * W2 + (t & 1) * W;
        T *dst = work + (size_t)t * W;
        warp_iir<T>(p.tab + p.filt[QF_PRE_LP].off, p.filt[QF_PRE_LP],
                    [&](int j) { return src[j]; }, [&](int j, T v) { dst[j] = v; });
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < g.count * W; idx += blockDim.x) {
        const int k = idx / W, x = idx - k * W;
        const int row = g.r0 + 2 * k, line = io.y0 + row;
        T s, cc;
        Real<T>::sincos_turns(start_phase(p, g.frame, line) + (unsigned long long)x * p.phases[QP_STEP1X], s, cc);
        T vl = work[(size_t)(2 * k + 1) * W + x];
        if (is_alternate(p, g.frame, line)) vl = -vl;
        const T y = cbuf[(size_t)(k + 1) * W + x] - (s * work[(size_t)(2 * k) * W + x] + cc * vl);
        store_rgb(p, io, g.fidx, row, x, y, gbuf[(size_t)k * W2 + x], gbuf[(size_t)k * W2 + W + x]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Line-comb decoders that work on the band-passed 2x signal B[k] = BP(up2 c[k]) of neighbouring rows.
// qam.demodulate (qam.py:43-58) is linear in its composite argument, so demodulating a line difference equals
// combining the B's first; likewise the wrappers' 0.5*(a+b) averages commute with the low-pass and down2.
//
//   COMB_NTSC2  NtscCombModem rows with a predecessor (ntsc.py:61-82 + comb.py:50-53):
//         u =  f down2 LP(2 cos(phi) (B[k]-B[k-1])),  v = -f down2 LP(2 sin(phi) (B[k]-B[k-1])),
//         phi = start_phase(line) - LS/2 + bp_shift
//   COMB_NTSC3  Simple3DCombModem(NtscCombModem) (comb.py:96-113): average of the NtscComb chroma of this row
//         (band-split chroma if the row has no predecessor) and of the next row of the field (zero at the bottom,
//         where the driver re-feeds the row itself, image.py:51-53)
//   COMB_PAL3   Pal3DModem rows with a predecessor (pal.py:198-226):
//         S = demod(c[k+1]-c[k-1]), D = demod(c[k+1]-2c[k]+c[k-1]) at the phase of row k,
//         u = a_ss S.v + a_cu D.u,  v = a_ss S.u + a_cv D.v, V switch; c[k+1] := c[k] at the bottom
//   then  y = c[k] - remod(u, v)  for all three.
// smem: taps[128] + (R+2) * (c[W] + B[2W]) + 2R * [2W]
// ------------------------------------------------------------------------------------------------------------
enum { COMB_NTSC2 = 0, COMB_NTSC3 = 1, COMB_PAL3 = 2 };

template <typename T, int MODE>
__global__ void __launch_bounds__(CM_NTHREADS)
k_qam_comb(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, W2 = 2 * W;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int R = io.rows_per_cta;
    T *taps = sm;
    T *cbuf = sm + 128;                         // (R+2) x W    index k+1, k = -1 .. R
    T *bbuf = cbuf + (size_t)(R + 2) * W;       // (R+2) x 2W
    T *work = bbuf + (size_t)(R + 2) * W2;      // 2R x 2W
    const bool has_prev0 = g.r0 >= 2;                               // row k = -1 exists
    const bool has_next_last = g.r0 + 2 * g.count < io.nrows;       // row k = count exists
    const int k_lo = has_prev0 ? -1 : 0;
    const int k_hi = (MODE != COMB_NTSC2 && has_next_last) ? g.count : g.count - 1;
    copy_taps(taps, p, 2);
    for (int k = k_lo; k <= k_hi; ++k) load_comp_row(cbuf + (size_t)(k + 1) * W, io, g.fidx, g.r0 + 2 * k, W);
    __syncthreads();
    for (int k = k_lo; k <= k_hi; ++k)
        fir_up2(bbuf + (size_t)(k + 1) * W2, cbuf + (size_t)(k + 1) * W, W, taps + p.res[QR_UP2].off, threadIdx.x,
                blockDim.x);
    __syncthreads();
    for (int t = warp; t < k_hi - k_lo + 1; t += nwarps) {
        T *b = bbuf + (size_t)(k_lo + t + 1) * W2;
        warp_iir<T>(p.tab + p.filt[QF_BP2X].off, p.filt[QF_BP2X],
                    [&](int j) { return b[j]; }, [&](int j, T v) { b[j] = v; });
    }
    __syncthreads();
    const unsigned long long step = p.phases[QP_STEP2X];
    for (int t = warp; t < 2 * g.count; t += nwarps) {
        const int k = t >> 1;
        const bool is_v = (t & 1) != 0;
        const int line = io.y0 + g.r0 + 2 * k;
        const bool hp = (k > 0) || has_prev0;
        const bool hn = (k + 1 < g.count) || has_next_last;
        const T *bc = bbuf + (size_t)(k + 1) * W2;
        const T *bp = bbuf + (size_t)(hp ? k : k + 1) * W2;
        const T *bn = bbuf + (size_t)(hn ? k + 2 : k + 1) * W2;
        T *dst = work + (size_t)t * W2;
        const unsigned long long psi = start_phase(p, g.frame, line) + p.phases[QP_BP_SHIFT];
        if (MODE == COMB_NTSC2) {
            const T f2 = (T)2 * p.scalars[QS_NTSC_FACTOR];
            const unsigned long long phi = psi - p.phases[QP_HALF_LS];
            warp_iir<T>(p.tab + p.filt[QF_DEMOD_LP].off, p.filt[QF_DEMOD_LP],
                        [&](int j) {
                            T s, c;
                            Real<T>::sincos_turns(phi + (unsigned long long)j * step, s, c);
                            const T d = bc[j] - bp[j];
                            return is_v ? -f2 * s * d : f2 * c * d;
                        },
                        [&](int j, T v) { dst[j] = v; });
        } else if (MODE == COMB_NTSC3) {
            const T f = p.scalars[QS_NTSC_FACTOR];      // 0.5 * (2 f ...) = f ...
            const unsigned long long phi = psi - p.phases[QP_HALF_LS];
            const unsigned long long phin = start_phase(p, g.frame, line + 2) + p.phases[QP_BP_SHIFT] -
                                            p.phases[QP_HALF_LS];
            warp_iir<T>(p.tab + p.filt[QF_DEMOD_LP].off, p.filt[QF_DEMOD_LP],
                        [&](int j) {
                            T s, c, acc;
                            const unsigned long long dj = (unsigned long long)j * step;
                            if (hp) {
                                Real<T>::sincos_turns(phi + dj, s, c);
                                const T d = bc[j] - bp[j];
                                acc = is_v ? -f * s * d : f * c * d;
                            } else {
                                Real<T>::sincos_turns(psi + dj, s, c);
                                acc = is_v ? c * bc[j] : s * bc[j];
                            }
                            if (hn) {
                                Real<T>::sincos_turns(phin + dj, s, c);
                                const T d = bn[j] - bc[j];
                                acc += is_v ? -f * s * d : f * c * d;
                            }
                            return acc;
                        },
                        [&](int j, T v) { dst[j] = v; });
        } else {
            const T a_ss = (T)2 * p.scalars[QS_P3D_SINSUM];
            const T a_c = (T)2 * (is_v ? p.scalars[QS_P3D_COSV] : p.scalars[QS_P3D_COSU]);
            warp_iir<T>(p.tab + p.filt[QF_DEMOD_LP].off, p.filt[QF_DEMOD_LP],
                        [&](int j) {
                            T s, c;
                            Real<T>::sincos_turns(psi + (unsigned long long)j * step, s, c);
                            const T curr_diff = bn[j] - bc[j], last_diff = bc[j] - bp[j];
                            const T ssig = curr_diff + last_diff, dsig = curr_diff - last_diff;
                            return is_v ? (a_ss * s * ssig + a_c * c * dsig) : (a_ss * c * ssig + a_c * s * dsig);
                        },
                        [&](int j, T v) { dst[j] = v; });
        }
    }
    __syncthreads();
    // u, v at 1x into bbuf rows: u at [k+1][0..W), v at [k+1][W..2W)   (the B's are dead now)
    for (int t = 0; t < 2 * g.count; ++t) {
        const int k = t >> 1;
        T *o = bbuf + (size_t)(k + 1) * W2 + (t & 1) * W;
        const bool neg = (MODE == COMB_PAL3) && (t & 1) && is_alternate(p, g.frame, io.y0 + g.r0 + 2 * k);
        fir_down2(work + (size_t)t * W2, W2, taps + p.res[QR_DOWN2].off, threadIdx.x, blockDim.x,
                  [&](int j, T v) { o[j] = neg ? -v : v; });
    }
    __syncthreads();
    for (int t = warp; t < 2 * g.count; t += nwarps) {
        const T *src = bbuf + (size_t)((t >> 1) + 1) * W2 + (t & 1) * W;
        T *dst = work + (size_t)t * W;
        warp_iir<T>(p.tab + p.filt[QF_PRE_LP].off, p.filt[QF_PRE_LP],
                    [&](int j) { return src[j]; }, [&](int j, T v) { dst[j] = v; });
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < g.count * W; idx += blockDim.x) {
        const int k = idx / W, x = idx - k * W;
        const int row = g.r0 + 2 * k, line = io.y0 + row;
        T s, cc;
        Real<T>::sincos_turns(start_phase(p, g.frame, line) + (unsigned long long)x * p.phases[QP_STEP1X], s, cc);
        T vl = work[(size_t)(2 * k + 1) * W + x];
        if ((p.flags & 1) && is_alternate(p, g.frame, line)) vl = -vl;
        const T y = cbuf[(size_t)(k + 1) * W + x] - (s * work[(size_t)(2 * k) * W + x] + cc * vl);
        store_rgb(p, io, g.fidx, row, x, y, bbuf[(size_t)(k + 1) * W2 + x], bbuf[(size_t)(k + 1) * W2 + W + x]);
    }
}
