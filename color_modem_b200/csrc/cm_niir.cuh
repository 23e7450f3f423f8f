// NIIR / SECAM-IV kernels.  Reference: color_modem/color/niir.py.
//
// Encode (niir.py:40-48, 67-88, 166-202): chroma vector gets a +0.1 saturation offset in polar form (or, for the
// hue-correcting encoder, the saturation-weighted hue of this row and the next row of the field with this row's
// saturation), both components are low-passed, then normal rows carry db sin + dr cos (QAM) and alternate rows
// the reference carrier -|c| sin.
//
// Decode (niir.py:102-163) at 3x: band-pass, envelope (pi/2 |x| -> low-pass) = saturation, x / saturation =
// phase-modulated carrier; each row is demodulated against the carrier of its neighbour row of the field (the
// previous row; a synthetic carrier at the field top): products with the carrier and with its derivative give
// sin/cos of the hue after down3, normalisation and rotation by +-line_shift; luma = composite - re-synthesised
// chroma; the saturation offset is removed last.  3x buffers are stored naturally (odd chunk lengths keep the
// IIR loads conflict-free, and stride-3 FIR accesses are conflict-free by themselves).
#pragma once
#include "cm_common.cuh"
#include "cm_fir.cuh"
#include "cm_iir.cuh"
#include "cm_io.cuh"
#include "cm_slots.h"

// sample j of a polyphase 3x buffer
template <typename T>
__device__ __forceinline__ T &poly3(T *buf, int hb, int j) { return buf[(j % 3) * hb + j / 3]; }

// The NIIR encoders add their +0.1 saturation offset along the *direction* of the chroma vector (niir.py:43-48,
// 186-197).  For grey pixels that vector is zero in exact arithmetic, and the reference's direction is whatever the
// float64 rounding residue (~1e-17) of its matrix product points at — an O(0.1) effect on the composite decided by the
// last bit.  To stay a drop-in on such pixels (every black, white or grey area of a real picture) the vector is
// computed here in float64 with the reference's own operation order and no FMA contraction, from u8 / 255.0 exactly as
// image.py:35-37 does; only the result is rounded to T.
template <typename T>
__device__ __forceinline__ void niir_load_rgb4_f64(const IoArgs<T> &io, size_t px, double r[4], double g[4], double b[4]) {
    if (io.in_f) {
        T v[12];
#pragma unroll
        for (int q = 0; q < 3; ++q) ld4(io.in_f + px * 3 + 4 * q, v + 4 * q);
#pragma unroll
        for (int i = 0; i < 4; ++i) { r[i] = (double)v[3 * i]; g[i] = (double)v[3 * i + 1]; b[i] = (double)v[3 * i + 2]; }
    } else {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(io.in_u8 + px * 3);
        const uint32_t ww[3] = {__ldg(w), __ldg(w + 1), __ldg(w + 2)};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = 3 * i;
            r[i] = __ddiv_rn((double)((ww[k >> 2] >> (8 * (k & 3))) & 0xff), 255.0);
            g[i] = __ddiv_rn((double)((ww[(k + 1) >> 2] >> (8 * ((k + 1) & 3))) & 0xff), 255.0);
            b[i] = __ddiv_rn((double)((ww[(k + 2) >> 2] >> (8 * ((k + 2) & 3))) & 0xff), 255.0);
        }
    }
}

// niir.py:35-38: c0 * r + c1 * g + c2 * b evaluated left to right in float64 (c2 carries the sign of the reference's "-")
__device__ __forceinline__ void niir_chroma_f64(const double *e, double r, double g, double b, double &db, double &dr) {
    db = __dadd_rn(__dadd_rn(__dmul_rn(e[3], r), __dmul_rn(e[4], g)), __dmul_rn(e[5], b));
    dr = __dadd_rn(__dadd_rn(__dmul_rn(e[6], r), __dmul_rn(e[7], g)), __dmul_rn(e[8], b));
}

// One pixel of the NIIR encoder front end in float64 with the reference's operation order (niir.py:40-48, 186-197;
// comb.py:149-150): (r, g, b) of the row and (r2, g2, b2) of its field neighbour -> (db, dr) with the saturation offset.
// (not inlined: grey pixels are the exception, and twelve to sixteen inlined copies of the float64 square roots and divisions
// made the row encoder 138 KB of code)
template <typename T>
__device__ __noinline__ void niir_exact_pixel(const DevParams<T> &p, bool hue, bool avg, double r, double g, double b,
                                                 double r2, double g2, double b2, T &db, T &dr) {
    double vb, vr, nb = 0.0, nr = 0.0, mag_out;
    niir_chroma_f64(p.encd, r, g, b, vb, vr);
    if (avg || hue) niir_chroma_f64(p.encd, r2, g2, b2, nb, nr);
    if (p.ident_enc) {
        vb = g;
        vr = b;
        if (avg || hue) { nb = g2; nr = b2; }
    }
    if (hue) {
        const double ls = sqrt(__dadd_rn(__dmul_rn(vb, vb), __dmul_rn(vr, vr)));
        const double sn = sqrt(__dadd_rn(__dmul_rn(nb, nb), __dmul_rn(nr, nr)));
        double div = __dadd_rn(ls, sn);
        if (div == 0.0) div = 1.0;
        const double ab = __ddiv_rn(__dadd_rn(__dmul_rn(vb, ls), __dmul_rn(nb, sn)), div);
        const double ar = __ddiv_rn(__dadd_rn(__dmul_rn(vr, ls), __dmul_rn(nr, sn)), div);
        mag_out = ls + 0.1;
        vb = ab;
        vr = ar;
    } else {
        if (avg) {                                                    // comb.py:149-150
            vb = __dmul_rn(0.5, __dadd_rn(nb, vb));
            vr = __dmul_rn(0.5, __dadd_rn(nr, vr));
        }
        mag_out = sqrt(__dadd_rn(__dmul_rn(vb, vb), __dmul_rn(vr, vr))) + 0.1;
    }
    const double m2 = vb * vb + vr * vr;
    if (m2 > 0.0) {
        const double sc = mag_out / sqrt(m2);
        db = (T)(vb * sc);
        dr = (T)(vr * sc);
    } else if (!signbit(vr)) {        // atan2(+-0, +0) = +-0 -> (sin, cos) = (+-0, 1)
        db = (T)(mag_out * copysign(0.0, vb));
        dr = (T)mag_out;
    } else {                          // atan2(+-0, -0) = +-pi -> (sin, cos) = (+-1.2e-16, -1), as numpy has it
        db = (T)(mag_out * copysign(1.2246467991473532e-16, vb));
        dr = (T)(-mag_out);
    }
}

// The pixels of a quad whose chroma vector is too short for fp32 (mask bit i), redone in float64 from the frame itself.
template <typename T>
__device__ __noinline__ void niir_exact_quad(const DevParams<T> &p, const IoArgs<T> &io, size_t px, size_t pxn, bool hue, bool avg,
                                             unsigned mask, T *db, T *dr) {
    double rd[4], gd[4], bd[4], r2d[4] = {0, 0, 0, 0}, g2d[4] = {0, 0, 0, 0}, b2d[4] = {0, 0, 0, 0};
    niir_load_rgb4_f64(io, px, rd, gd, bd);
    if (avg || hue) niir_load_rgb4_f64(io, pxn, r2d, g2d, b2d);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (mask & (1u << i)) niir_exact_pixel<T>(p, hue, avg, rd[i], gd[i], bd[i], r2d[i], g2d[i], b2d[i], db[i], dr[i]);
}

// ------------------------------------------------------------------------------------------------------------
// Encode.  2 warps per row.  smem: R * 3 * N1   (luma | db | dr)
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(CM_NTHREADS)
k_niir_encode(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, N1 = p.n1p, W4 = W >> 2;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool avg = (p.flags & 2) != 0, hue = (p.flags & 4) != 0;
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const int nrow = (row + 2 < io.nrows) ? row + 2 : row;
        T *ys = sm + (size_t)k * 3 * N1, *bs = ys + N1, *rs_ = bs + N1;
        for (int q = threadIdx.x; q < W4; q += blockDim.x) {
            const int x = 4 * q;
            const size_t px = ((size_t)g.fidx * io.nrows + row) * W + x, pxn = ((size_t)g.fidx * io.nrows + nrow) * W + x;
            T r[4], gg[4], b[4], r2[4], g2[4], b2[4], y[4], db[4], dr[4];
            load_rgb4(io, px, r, gg, b);
            if (avg || hue) load_rgb4(io, pxn, r2, g2, b2);
            bool exact[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                y[i] = p.enc[0] * r[i] + p.enc[1] * gg[i] + p.enc[2] * b[i];
                T vb = p.enc[3] * r[i] + p.enc[4] * gg[i] + p.enc[5] * b[i];
                T vr = p.enc[6] * r[i] + p.enc[7] * gg[i] + p.enc[8] * b[i];
                if (p.ident_enc) { y[i] = r[i]; vb = gg[i]; vr = b[i]; }       // modulate_components: planes as given
                T mag_out;
                if (hue) {
                    // niir.py:186-197: hue of the saturation-weighted mean of this row and the next, this row's saturation
                    T nb = p.enc[3] * r2[i] + p.enc[4] * g2[i] + p.enc[5] * b2[i];
                    T nr = p.enc[6] * r2[i] + p.enc[7] * g2[i] + p.enc[8] * b2[i];
                    if (p.ident_enc) { nb = g2[i]; nr = b2[i]; }
                    const T ls = Real<T>::sqrt_(vb * vb + vr * vr), sn = Real<T>::sqrt_(nb * nb + nr * nr);
                    T div = ls + sn;
                    if (div == (T)0) div = (T)1;
                    const T ab = (vb * ls + nb * sn) / div, ar = (vr * ls + nr * sn) / div;
                    mag_out = ls + (T)0.1;
                    vb = ab;
                    vr = ar;
                } else {
                    if (avg) {
                        T nb = p.enc[3] * r2[i] + p.enc[4] * g2[i] + p.enc[5] * b2[i];
                        T nr = p.enc[6] * r2[i] + p.enc[7] * g2[i] + p.enc[8] * b2[i];
                        if (p.ident_enc) { nb = g2[i]; nr = b2[i]; }
                        vb = (T)0.5 * (nb + vb);
                        vr = (T)0.5 * (nr + vr);
                    }
                    mag_out = Real<T>::sqrt_(vb * vb + vr * vr) + (T)0.1;      // niir.py:43
                }
                // mag_out * (sin, cos)(atan2(vb, vr))
                const T m2 = vb * vb + vr * vr;
                // The direction of a vector shorter than 1e-4 is not reliable in fp32 (and is rounding noise of the
                // reference itself when it is ~1e-17): redo those pixels in float64 the reference's way.  The fp64 build
                // always takes the exact path.
                exact[i] = sizeof(T) == 8 || !(m2 >= (T)1e-8);
                const T sc = mag_out * Real<T>::rsqrt_(m2);
                db[i] = vb * sc;
                dr[i] = vr * sc;
            }
            if (exact[0] || exact[1] || exact[2] || exact[3]) {
                double rd[4], gd[4], bd[4], r2d[4], g2d[4], b2d[4];
                niir_load_rgb4_f64(io, px, rd, gd, bd);
                if (avg || hue) niir_load_rgb4_f64(io, pxn, r2d, g2d, b2d);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!exact[i]) continue;
                    niir_exact_pixel<T>(p, hue, avg, rd[i], gd[i], bd[i], r2d[i], g2d[i], b2d[i], db[i], dr[i]);
                }
            }
            st4(ys + x, y);
            st4(bs + x, db);
            st4(rs_ + x, dr);
        }
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) cta_fill_tail<T, 1>(sm + (size_t)k * 3 * N1 + N1, (size_t)N1, 2, N1, W, N1);
    __syncthreads();
    const FiltHdr &fpre = p.filt[NF_PRE_LP];
    for (int t = warp; t < 2 * g.count; t += nwarps) {
        T *buf = sm + (size_t)(t >> 1) * 3 * N1 + (1 + (t & 1)) * N1;
        warp_iir<T, 1>(p.tab + fpre.off, fpre, [&](int q, int, int) { return buf[q]; },
                       [&](int j, T v) { buf[j] = v; });
    }
    __syncthreads();
    T crs, crc;
    Real<T>::sincos_turns(p.phases[NP_STEP1X], crs, crc);
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k, line = io.y0 + row;
        const T *ys = sm + (size_t)k * 3 * N1, *bs = ys + N1, *rs_ = bs + N1;
        const unsigned long long ph0 = start_phase(p, g.frame, line);
        const bool alt = is_alternate(p, g.frame, line);
        for (int q = threadIdx.x; q < W4; q += blockDim.x) {
            const int x = 4 * q;
            T y[4], db[4], dr[4], o[4], s[4], c[4];
            ld4(ys + x, y);
            ld4(bs + x, db);
            ld4(rs_ + x, dr);
            // u8 frames: hardware sin / cos seed + three rotations; float frames (parity path): exact per sample
            if (io.in_u8) {
                carrier4_fast(ph0 + (unsigned long long)x * p.phases[NP_STEP1X], crs, crc, s, c);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) Real<T>::sincos_turns(ph0 + (unsigned long long)(x + i) * p.phases[NP_STEP1X], s[i], c[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                o[i] = y[i] + (alt ? -Real<T>::sqrt_(db[i] * db[i] + dr[i] * dr[i]) * s[i] : db[i] * s[i] + dr[i] * c[i]);
            store_comp4(io, ((size_t)g.fidx * io.nrows + row) * p.Wc + x, o);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Encoder for u8 frames, second generation (k_niir_encode2): a CTA of 2 or 4 warps (EncGeo, cm_qam.cuh) walks a STRIP of
// consecutive rows of one field, one row at a time.  The hue-correcting and the averaging front ends read the row and its field
// neighbour (niir.py:186-197, comb.py:149-150): the neighbour's planes before the polar step are kept in shared memory (each thread
// its own pixels) and are the row's own planes one iteration later, so every row is fetched, unpacked and matrixed once per
// strip; the words of the row after are in flight during the filtering.  The two chroma
// low-passes are packed DF-I team recursions (team_iir_pk).  fp32: |v| = m rsqrt(m) and one hardware reciprocal per pixel
// instead of two square roots and two divisions (the parity bound is 1e-4; the float64 build keeps the exact forms).
// smem: scratch[128] | y[N1] | db[N1] | dr[N1] | stash: y, db, dr, |c| of the row [4 N1]
// ------------------------------------------------------------------------------------------------------------
#define NF_ENC_PRE 9     // DevParams::filt slot of this kernel's low-pass site (cm_api.cu: plan_encode_kernel)
template <typename T> struct NiirFast {
    static __device__ __forceinline__ T norm(T m2) { return Real<T>::sqrt_(m2); }
};
template <> struct NiirFast<float> {
    static __device__ __forceinline__ float norm(float m2) { return m2 > 0.f ? m2 * rsqrtf(m2) : 0.f; }
};

template <typename T, int GEO>
__global__ void __launch_bounds__(32 * EncGeo<GEO>::NW, sizeof(T) == 8 ? 1 : 16 / EncGeo<GEO>::NW)
k_niir_encode2(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;
    typedef EncGeo<GEO> EG;
    constexpr int kQ = EG::KQ, NT = 32 * EG::NW, TH = EG::NW / 2;
    const int W = p.W, N1 = p.n1p, W4 = W >> 2;
    const int warp = threadIdx.x >> 5, task = warp / TH, wr = warp - task * TH;
    const bool avg = (p.flags & 2) != 0, hue = (p.flags & 4) != 0, two = avg || hue;
    const int field = blockIdx.y, f = blockIdx.z;
    const long long frame = io.first_frame + f;
    const int first = io.out_begin + field, nout = (io.out_count - field + 1) >> 1;
    const int R = io.rows_per_cta;
    const int k0 = blockIdx.x * R, k1 = min(k0 + R, nout);
    if (k0 >= nout) return;
    const FiltHdr &fpre = p.filt[NF_ENC_PRE];
    T *ys = sm, *bs = ys + N1, *rs_ = bs + N1;
    T *st_y = rs_ + N1, *st_b = st_y + N1, *st_r = st_b + N1, *st_s = st_r + N1;      // the row's own planes, kept from the iteration before
    uint32_t wn[kQ][3], wt[kQ][3];
    auto next_of = [&](int row) { return (row + 2 < io.nrows) ? row + 2 : row; };
    auto fetch = [&](int row, uint32_t (*w)[3]) {
        const uint32_t *a = reinterpret_cast<const uint32_t *>(io.in_u8 + ((size_t)f * io.nrows + row) * W * 3);
#pragma unroll
        for (int j = 0; j < kQ; ++j) {
            const int q = threadIdx.x + j * NT;
            if (q < W4) {
#pragma unroll
                for (int i = 0; i < 3; ++i) w[j][i] = __ldg(a + 3 * q + i);
            }
        }
    };
    // (y, db, dr) of four pixels before the polar step, and |(db, dr)| for the hue-correcting front end
    auto planes4 = [&](const uint32_t *w, T *y, T *vb, T *vr, T *mag) {
        unsigned char bytes[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bytes[i] = (w[0] >> (8 * i)) & 0xff;
            bytes[4 + i] = (w[1] >> (8 * i)) & 0xff;
            bytes[8 + i] = (w[2] >> (8 * i)) & 0xff;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const T r = Real<T>::from_u8(bytes[3 * i]), g = Real<T>::from_u8(bytes[3 * i + 1]), b = Real<T>::from_u8(bytes[3 * i + 2]);
            y[i] = p.enc[0] * r + p.enc[1] * g + p.enc[2] * b;
            vb[i] = p.enc[3] * r + p.enc[4] * g + p.enc[5] * b;
            vr[i] = p.enc[6] * r + p.enc[7] * g + p.enc[8] * b;
            if (p.ident_enc) { y[i] = r; vb[i] = g; vr[i] = b; }                    // modulate_components: planes as given
            mag[i] = hue ? NiirFast<T>::norm(vb[i] * vb[i] + vr[i] * vr[i]) : (T)0;
        }
    };
    T crs, crc;
    Real<T>::sincos_turns(p.phases[NP_STEP1X], crs, crc);
    fetch(first + 2 * k0, wn);
    if (two) {                  // the strip's first row: its planes into the stash, then the words of its field neighbour
#pragma unroll
        for (int j = 0; j < kQ; ++j) {
            const int q = threadIdx.x + j * NT;
            if (q < W4) {
                T y[4], vb[4], vr[4], mg[4];
                planes4(wn[j], y, vb, vr, mg);
                st4(st_y + 4 * q, y);
                st4(st_b + 4 * q, vb);
                st4(st_r + 4 * q, vr);
                st4(st_s + 4 * q, mg);
            }
        }
        fetch(next_of(first + 2 * k0), wn);
    }
    for (int k = k0; k < k1; ++k) {
        const int row = first + 2 * k, nrow = next_of(row), line = io.y0 + row;
#pragma unroll
        for (int j = 0; j < kQ; ++j) {
            const int q = threadIdx.x + j * NT;
            if (q < W4) {
                const int x = 4 * q;
                T y[4], vb4[4], vr4[4], ls4[4], yn[4], nb4[4], nr4[4], sn4[4], db[4], dr[4];
                planes4(wn[j], yn, nb4, nr4, sn4);                                  // wn: the neighbour row (or, without one, the row)
                if (two) {
                    ld4(st_y + x, y);
                    ld4(st_b + x, vb4);
                    ld4(st_r + x, vr4);
                    ld4(st_s + x, ls4);
                    st4(st_y + x, yn);                                              // the neighbour is the next row of the strip
                    st4(st_b + x, nb4);
                    st4(st_r + x, nr4);
                    st4(st_s + x, sn4);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { y[i] = yn[i]; vb4[i] = nb4[i]; vr4[i] = nr4[i]; ls4[i] = sn4[i]; }
                }
                bool exact[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    T vb = vb4[i], vr = vr4[i];
                    const T nb = nb4[i], nr = nr4[i];
                    T mag_out, m2;
                    if (hue) {                                                      // niir.py:186-197
                        const T ls = ls4[i], sn = sn4[i];
                        T div = ls + sn;
                        if (div == (T)0) div = (T)1;
                        const T inv = FastRcp<T>::rcp(div);
                        const T ab = (vb * ls + nb * sn) * inv, ar = (vr * ls + nr * sn) * inv;
                        mag_out = ls + (T)0.1;
                        vb = ab;
                        vr = ar;
                        m2 = vb * vb + vr * vr;
                    } else {
                        if (avg) {                                                  // comb.py:149-150
                            vb = (T)0.5 * (nb + vb);
                            vr = (T)0.5 * (nr + vr);
                        }
                        m2 = vb * vb + vr * vr;
                        mag_out = NiirFast<T>::norm(m2) + (T)0.1;                   // niir.py:43
                    }
                    // direction of a vector shorter than 1e-4: float64, the reference's way (see k_niir_encode)
                    exact[i] = sizeof(T) == 8 || !(m2 >= (T)1e-8);
                    const T sc = mag_out * Real<T>::rsqrt_(m2);
                    db[i] = vb * sc;
                    dr[i] = vr * sc;
                }
                if (exact[0] || exact[1] || exact[2] || exact[3])
                    niir_exact_quad<T>(p, io, ((size_t)f * io.nrows + row) * W + x, ((size_t)f * io.nrows + nrow) * W + x, hue, avg,
                                       (exact[0] ? 1u : 0u) | (exact[1] ? 2u : 0u) | (exact[2] ? 4u : 0u) | (exact[3] ? 8u : 0u), db, dr);
                st4(ys + x, y);
                st4(bs + x, db);
                st4(rs_ + x, dr);
            }
        }
        __syncthreads();
        if (k + 1 < k1) fetch(two ? next_of(row + 2) : row + 2, wt);               // in flight during the filtering
        {
            T *buf = task ? rs_ : bs;
            warp_fill_tail<T, 1>(buf, N1, W, iir_tail_end(fpre));                  // every warp of the team writes the same values
            team_iir_pk<T, 1, EG::PRE, TH>(p.tab + fpre.off, fpre, LoadLinear<T, EG::PRE>{buf}, [&](int j, T x) { buf[j] = x; },
                                           wr, 2 + task, scratch + 32 * task);
        }
        __syncthreads();
        const unsigned long long ph0 = start_phase(p, frame, line);
        const bool alt = is_alternate(p, frame, line);
        for (int q = threadIdx.x; q < W4; q += NT) {
            const int x = 4 * q;
            T y[4], db[4], dr[4], o[4], s[4], c[4];
            ld4(ys + x, y);
            ld4(bs + x, db);
            ld4(rs_ + x, dr);
            carrier4_fast(ph0 + (unsigned long long)x * p.phases[NP_STEP1X], crs, crc, s, c);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                o[i] = y[i] + (alt ? -NiirFast<T>::norm(db[i] * db[i] + dr[i] * dr[i]) * s[i] : db[i] * s[i] + dr[i] * c[i]);
            store_comp4(io, ((size_t)f * io.nrows + row) * p.Wc + x, o);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kQ; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) wn[j][i] = wt[j][i];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Decode.  smem: scratch[128] (IIR team scratch) + (R+1) rows x ( c[N1] | up[N3] | mod->pm[N3] | sat[N3] ), 3x buffers polyphase
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(192, 2)
k_niir_decode(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, N1 = p.n1p, hb = p.hb3, N3 = 3 * hb, n3 = 3 * W;
    T *taps = sm;
    T *rows = sm + 128;
    const size_t per_row = (size_t)N1 + 3 * (size_t)N3;
    const bool has_prev0 = g.r0 >= 2;
    const int nin = g.count + 1;                       // row slot 0 = previous row (real or synthetic), slot k+1 = row k
    const FirTaps<T> hup{p.firc[NR_UP3], p.fircp[NR_UP3]}, hdn{p.firc[NR_DOWN3], p.fircp[NR_DOWN3]};   // constant bank (kernel parameter)
    auto rowp = [&](int k) { return rows + (size_t)(k + 1) * per_row; };
    load_comp_rows(io, g.fidx, has_prev0 ? nin : g.count, W,
                   [&](int k) { return rowp(has_prev0 ? k - 1 : k); },
                   [&](int k) { return g.r0 + 2 * (has_prev0 ? k - 1 : k); });
    if (!has_prev0) {
        // niir.py:103-106: synthetic reference carrier of line-2: sin(phi) on normal lines, -sin(phi) on alternate ones
        const int line = io.y0 + g.r0 - 2;
        const unsigned long long ph0 = start_phase(p, g.frame, line);
        const T sgn = is_alternate(p, g.frame, line) ? (T)-1 : (T)1;
        T *c = rowp(-1);
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            T s, co;
            Real<T>::sincos_turns(ph0 + (unsigned long long)x * p.phases[NP_STEP1X], s, co);
            c[x] = sgn * s;
        }
    }
    __syncthreads();
    for (int k = -1; k < g.count; ++k) {
        T *u = rowp(k) + N1;
        fir_up3(u, u + hb, u + 2 * hb, rowp(k), W, hup, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    cta_fill_tail<T, 3>(rows + N1, per_row, nin, hb, n3, N3);
    __syncthreads();
    const FiltHdr &fbp = p.filt[NF_UP_BP], &flp = p.filt[NF_BASE_LP];
    // band-pass: up -> mod.  The 3x lines are 2-4 super-chunks long: each task is run by a team of warps (cm_iir.cuh)
    for_each_iir_task<T, true>(fbp, nin, taps, [&](int t, const IirTeam<T> &tm) {
        const T *u = rows + (size_t)t * per_row + N1;
        T *m = rows + (size_t)t * per_row + N1 + N3;
        team_iir<T, 3, true>(p.tab + fbp.off, fbp, [&](int q, int ph, int) { return u[ph * hb + q]; },
                             Poly3Out<T>{m, hb}, tm);
    });
    __syncthreads();
    // envelope low-pass: sat = LP(pi/2 |mod|).  The synthetic top-of-field carrier is used un-normalised
    // (niir.py:105-106 stores the band-passed reference itself), so its slot skips this step.
    for_each_iir_task<T, true>(flp, nin, taps, [&](int t, const IirTeam<T> &tm) {
        if (t == 0 && !has_prev0) return;
        T *m = rows + (size_t)t * per_row + N1 + N3;
        T *s = rows + (size_t)t * per_row + N1 + 2 * (size_t)N3;
        warp_fill_tail<T, 3>(m, hb, n3, N3);           // every warp of the team writes the same values
        team_iir<T, 3, true>(p.tab + flp.off, flp,
                             [&](int q, int ph, int) { return (T)1.57079632679489661923 * Real<T>::abs_(m[ph * hb + q]); },
                             Poly3Out<T>{s, hb}, tm);
    });
    __syncthreads();
    for (int k = -1; k < g.count; ++k) {                // pm = mod / sat in place
        if (k == -1 && !has_prev0) continue;
        T *m = rowp(k) + N1 + N3;
        const T *s = rowp(k) + N1 + 2 * (size_t)N3;
        for (int ph = 0; ph < 3; ++ph) {
            T *mp = m + ph * hb;
            const T *sp = s + ph * hb;
            for (int q = 4 * threadIdx.x; q < W; q += 4 * blockDim.x) {
                T a[4], b[4];
                ld4(mp + q, a);
                ld4(sp + q, b);
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = a[i] / b[i];
                st4(mp + q, a);
            }
        }
    }
    __syncthreads();
    // derivative of the carrier (niir.py:123-125) into the (dead) `up` buffer of each output row
    const T inv_step3 = p.scalars[NS_INV_STEP3] * (T)0.5;
    for (int k = 0; k < g.count; ++k) {
        const bool alt = is_alternate(p, g.frame, io.y0 + g.r0 + 2 * k);
        const T *car = (alt ? rowp(k) : rowp(k - 1)) + N1 + N3;
        const T *c0 = car, *c1 = car + hb, *c2 = car + 2 * hb;
        T *a = rowp(k) + N1;
        for (int m = threadIdx.x; m < W; m += blockDim.x) {
            a[m] = m > 0 ? (c1[m] - c2[m - 1]) * inv_step3 : (T)0;                 // j = 3m
            a[hb + m] = (c2[m] - c0[m]) * inv_step3;                               // j = 3m + 1
            a[2 * hb + m] = m + 1 < W ? (c0[m + 1] - c1[m]) * inv_step3 : (T)0;    // j = 3m + 2
        }
    }
    __syncthreads();
    const Down3Taps<T> tp(hdn);
    T ls_s, ls_c;
    Real<T>::sincos_turns(p.phases[NP_LINE_SHIFT], ls_s, ls_c);
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k, line = io.y0 + row;
        const bool alt = is_alternate(p, g.frame, line);
        const T *c = rowp(k);
        const T *pm = rowp(k) + N1 + N3, *last = rowp(k - 1) + N1 + N3, *satu = rowp(k) + N1 + 2 * (size_t)N3;
        const T *carrier = alt ? pm : last, *huemod = alt ? last : pm, *altc = rowp(k) + N1;
        const T sh_s = alt ? -ls_s : ls_s, sh_c = ls_c;                           // niir.py:114-121,136-137
        T rot_s, rot_c;                                                           // niir.py:146-156
        Real<T>::sincos_turns(p.phases[NP_LUMA_ROT] + (alt ? 0ull : p.phases[NP_LINE_SHIFT]), rot_s, rot_c);
        for (int q = threadIdx.x; q < (W >> 2); q += blockDim.x) {
            const int j0 = 4 * q;
            T sinphi[4], cosphi[4], sat[4], sincar[4], coscar[4], y[4], ob[4], orr[4], cc[4];
            down3_quad_prod(tp, huemod, carrier, hb, W, j0, sinphi);
            down3_quad_prod(tp, huemod, altc, hb, W, j0, cosphi);
            down3_quad(tp, satu, satu + hb, satu + 2 * hb, W, j0, sat);
            down3_quad(tp, carrier, carrier + hb, carrier + 2 * hb, W, j0, sincar);
            down3_quad(tp, altc, altc + hb, altc + 2 * hb, W, j0, coscar);
            ld4(c + j0, cc);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const T norm = Real<T>::sqrt_(cosphi[i] * cosphi[i] + sinphi[i] * sinphi[i]);
                const T cp0 = cosphi[i] / norm, sp0 = sinphi[i] / norm;
                const T sp = -cp0 * sh_s - sp0 * sh_c, cp = sp0 * sh_s - cp0 * sh_c;
                T db = sat[i] * sp, dr = sat[i] * cp;
                const T us0 = alt ? -Real<T>::sqrt_(db * db + dr * dr) : db, vs0 = alt ? (T)0 : dr;
                const T us = us0 * rot_c - vs0 * rot_s, vs = us0 * rot_s + vs0 * rot_c;
                y[i] = cc[i] - (us * sincar[i] + vs * coscar[i]);
                const T m2 = db * db + dr * dr;                                   // niir.py:61-65
                if (m2 > (T)0) {
                    const T mag = Real<T>::sqrt_(m2);
                    T ns = mag - (T)0.1;
                    ns = ns > (T)0 ? ns : (T)0;
                    const T sc = ns / mag;
                    db *= sc;
                    dr *= sc;
                } else {
                    db = (T)0;
                    dr = (T)0;
                }
                ob[i] = db;
                orr[i] = dr;
            }
            store_rgb4(p, io, g.fidx, row, j0, y, ob, orr);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Decode, second generation (k_niir_decode2).  A CTA of 2 (lines up to ~800 samples) or 4 warps (up to ~2100) walks a STRIP
// of consecutive rows of one field, one row at a time, keeping the normalised carrier of the row before in shared memory:
// the previous row is recomputed once per strip instead of once per two rows, every recursion is a packed DF-I team of all
// the warps (team_iir_pk), and the 3x buffers that live only within a row are shared:
//   c[N1] | sat[N1] | A[N3] | pm[2][N3]         A: up3(c) -> envelope at 3x -> derivative of the reference carrier
// (2 N1 + 3 N3 elements: 36 KB at 720 samples; the float64 build of 1920-sample lines fits one CTA per SM).
// Per row: up3, band-pass (3 sections), pi/2 |.| -> low-pass = envelope, carrier = band-passed / envelope, sat = down3(envelope),
// then the products with the neighbour row's carrier and its derivative, 4 more down3, hue / luma arithmetic, store.
// ------------------------------------------------------------------------------------------------------------
#define NF_ROW_BP 4          // DevParams::filt slots of this kernel's use-sites (cm_api.cu: plan_niir_kernel)
#define NF_ROW_LP 5
template <int GEO> struct NiirGeo;
template <> struct NiirGeo<1> { static constexpr int NW = 2, L3 = 39; };
template <> struct NiirGeo<3> { static constexpr int NW = 4, L3 = 51; };

template <typename T, int GEO>
__global__ void __launch_bounds__(32 * NiirGeo<GEO>::NW, sizeof(T) == 8 ? 1 : 12 / NiirGeo<GEO>::NW)
k_niir_decode2(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;
    typedef NiirGeo<GEO> NG;
    constexpr int NW = NG::NW, NT = 32 * NW, L3 = NG::L3;
    const int W = p.W, N1 = p.n1p, hb = p.hb3, N3 = 3 * hb, n3 = 3 * W;
    const int warp = threadIdx.x >> 5;
    const int field = blockIdx.y, f = blockIdx.z;
    const long long frame = io.first_frame + f;
    const int first = io.out_begin + field, nout = (io.out_count - field + 1) >> 1;      // rows first, first + 2, ...
    const int R = io.rows_per_cta;
    const int k0 = blockIdx.x * R, k1 = min(k0 + R, nout);
    if (k0 >= nout) return;
    T *c = sm, *sat = c + N1, *A = sat + N1, *pm0 = A + N3;
    const FirTaps<T> hup{p.firc[NR_UP3], p.fircp[NR_UP3]}, hdn{p.firc[NR_DOWN3], p.fircp[NR_DOWN3]};
    const Down3Taps<T> tp(hdn);
    const FiltHdr &fbp = p.filt[NF_ROW_BP], &flp = p.filt[NF_ROW_LP];
    const T inv_step3 = p.scalars[NS_INV_STEP3] * (T)0.5;
    T ls_s, ls_c;
    Real<T>::sincos_turns(p.phases[NP_LINE_SHIFT], ls_s, ls_c);

    // k = k0 - 1 is the row before the strip: a real row, or the synthetic reference carrier at the top of the field
    // (niir.py:103-106: sin(phi) on normal lines, -sin(phi) on alternate ones, band-passed but NOT normalised).  It only
    // leaves its carrier behind.  One copy of the per-row code serves both cases (the unrolled recursions and the five
    // decimators are ~60 KB of instructions; three inlined copies cost 18 % of the issue slots in instruction fetch).
    for (int k = k0 - 1; k < k1; ++k) {
        const int row = first + 2 * k, line = io.y0 + row;
        T *pm = pm0 + (size_t)(k & 1) * N3;
        const T *last = pm0 + (size_t)((k + 1) & 1) * N3;
        const bool synthetic = row < 0;
        if (!synthetic) {
            load_comp_row(c, io, f, row, W);
        } else {
            const unsigned long long ph0 = start_phase(p, frame, line);
            const T sgn = is_alternate(p, frame, line) ? (T)-1 : (T)1;
            for (int x = threadIdx.x; x < W; x += NT) {
                T s, co;
                Real<T>::sincos_turns(ph0 + (unsigned long long)x * p.phases[NP_STEP1X], s, co);
                c[x] = sgn * s;
            }
        }
        __syncthreads();
        fir_up3(A, A + hb, A + 2 * hb, c, W, hup, threadIdx.x, NT);
        __syncthreads();
        warp_fill_tail<T, 3>(A, hb, n3, iir_tail_end(fbp));         // every warp writes the same values
        team_iir_pk<T, 3, L3, NW>(p.tab + fbp.off, fbp, LoadPoly3<T, L3, false>{A, hb}, Poly3Out<T>{pm, hb}, warp, 1, scratch);
        __syncthreads();
        if (synthetic) continue;
        warp_fill_tail<T, 3>(pm, hb, n3, iir_tail_end(flp));
        team_iir_pk<T, 3, L3, NW>(p.tab + flp.off, flp, LoadPoly3<T, L3, true>{pm, hb}, Poly3Out<T>{A, hb}, warp, 1, scratch);
        __syncthreads();
        for (int q = 4 * threadIdx.x; q < W; q += 4 * NT) {          // saturation at 1x; carrier = band-passed / envelope
            T y[4];
            down3_quad(tp, A, A + hb, A + 2 * hb, W, q, y);
            st4(sat + q, y);
#pragma unroll
            for (int ph = 0; ph < 3; ++ph) {
                T a[4], b[4];
                ld4(pm + ph * hb + q, a);
                ld4(A + ph * hb + q, b);
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = a[i] * FastRcp<T>::rcp(b[i]);
                st4(pm + ph * hb + q, a);
            }
        }
        __syncthreads();
        if (k < k0) continue;
        const bool alt = is_alternate(p, frame, line);
        const T *carrier = alt ? pm : last, *huemod = alt ? last : pm;
        {   // derivative of the reference carrier (niir.py:123-125) into A (the envelope is dead)
            const T *c0 = carrier, *c1 = carrier + hb, *c2 = carrier + 2 * hb;
            for (int m = threadIdx.x; m < W; m += NT) {
                A[m] = m > 0 ? (c1[m] - c2[m - 1]) * inv_step3 : (T)0;                 // j = 3m
                A[hb + m] = (c2[m] - c0[m]) * inv_step3;                               // j = 3m + 1
                A[2 * hb + m] = m + 1 < W ? (c0[m + 1] - c1[m]) * inv_step3 : (T)0;    // j = 3m + 2
            }
        }
        __syncthreads();
        const T sh_s = alt ? -ls_s : ls_s, sh_c = ls_c;                           // niir.py:114-121,136-137
        T rot_s, rot_c;                                                           // niir.py:146-156
        Real<T>::sincos_turns(p.phases[NP_LUMA_ROT] + (alt ? 0ull : p.phases[NP_LINE_SHIFT]), rot_s, rot_c);
        for (int q = threadIdx.x; q < (W >> 2); q += NT) {
            const int j0 = 4 * q;
            T sinphi[4], cosphi[4], sv[4], sincar[4], coscar[4], y[4], ob[4], orr[4], cc[4];
            down3_quad_prod(tp, huemod, carrier, hb, W, j0, sinphi);
            down3_quad_prod(tp, huemod, (const T *)A, hb, W, j0, cosphi);
            down3_quad(tp, carrier, carrier + hb, carrier + 2 * hb, W, j0, sincar);
            down3_quad(tp, (const T *)A, (const T *)A + hb, (const T *)A + 2 * hb, W, j0, coscar);
            ld4(sat + j0, sv);
            ld4(c + j0, cc);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const T inv_norm = Real<T>::rsqrt_(cosphi[i] * cosphi[i] + sinphi[i] * sinphi[i]);
                const T cp0 = cosphi[i] * inv_norm, sp0 = sinphi[i] * inv_norm;
                const T sp = -cp0 * sh_s - sp0 * sh_c, cp = sp0 * sh_s - cp0 * sh_c;
                T db = sv[i] * sp, dr = sv[i] * cp;
                const T m2 = db * db + dr * dr;
                const T inv_mag = Real<T>::rsqrt_(m2), mag = m2 * inv_mag;      // (only used where m2 > 0)
                const T us0 = alt ? (m2 > (T)0 ? -mag : (T)0) : db, vs0 = alt ? (T)0 : dr;
                const T us = us0 * rot_c - vs0 * rot_s, vs = us0 * rot_s + vs0 * rot_c;
                y[i] = cc[i] - (us * sincar[i] + vs * coscar[i]);
                if (m2 > (T)0) {                                                  // niir.py:61-65
                    T ns = mag - (T)0.1;
                    ns = ns > (T)0 ? ns : (T)0;
                    const T sc = ns * inv_mag;
                    db *= sc;
                    dr *= sc;
                } else {
                    db = (T)0;
                    dr = (T)0;
                }
                ob[i] = db;
                orr[i] = dr;
            }
            store_rgb4(p, io, f, row, j0, y, ob, orr);
        }
        __syncthreads();
    }
}
