// QAM family kernels: NTSC / PAL encode, band-split decode, PAL-D delay-line decode, NTSC 2-line / 3-line comb,
// PAL 3-line comb.  Reference: color_modem/qam.py, color/ntsc.py, color/pal.py, comb.py.
//
// Buffer conventions (cm_iir.cuh / cm_fir.cuh): 1x signals natural with pitch N1 = p.n1p; 2x signals polyphase
// [E | O] with p.hb2 elements per phase, pitch N2 = 2 * p.hb2 (>= 2 * N1, so one 2x buffer can hold two 1x rows).
#pragma once
#include "cm_common.cuh"
#include "cm_fir.cuh"
#include "cm_iir.cuh"
#include "cm_io.cuh"
#include "cm_slots.h"

// one-row kernels: a CTA of two warps; 8 CTAs per SM at 128 registers per thread (more CTAs at fewer registers spill
// and lose)
#define CM_ROW_THREADS 64
#ifndef CM_ROWS_MINB
#define CM_ROWS_MINB 8
#endif


// ------------------------------------------------------------------------------------------------------------
// Encode: RGB -> Y + sin(phi) LP(U) + cos(phi) LP(+-V)          qam.py:20-32, ntsc.py:27-45, pal.py:32-52
// optional ColorAveragingModem front end (comb.py:141-152).     smem: R * 3 * N1
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool TEAMS>
__global__ void __launch_bounds__(CM_NTHREADS, 3)
k_qam_encode(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw) + 128;         // [0, 128): IIR team scratch
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, N1 = p.n1p, W4 = W >> 2;
    const bool avg = (p.flags & 2) != 0;

    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const int nrow = (row + 2 < io.nrows) ? row + 2 : row;
        T *ys = sm + (size_t)k * 3 * N1, *us = ys + N1, *vs = us + N1;
        for (int q = threadIdx.x; q < W4; q += blockDim.x) {
            const int x = 4 * q;
            T r[4], gg[4], b[4], y[4], u[4], v[4];
            load_rgb4(io, ((size_t)g.fidx * io.nrows + row) * W + x, r, gg, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                y[i] = p.enc[0] * r[i] + p.enc[1] * gg[i] + p.enc[2] * b[i];
                u[i] = p.enc[3] * r[i] + p.enc[4] * gg[i] + p.enc[5] * b[i];
                v[i] = p.enc[6] * r[i] + p.enc[7] * gg[i] + p.enc[8] * b[i];
            }
            if (avg) {
                load_rgb4(io, ((size_t)g.fidx * io.nrows + nrow) * W + x, r, gg, b);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const T un = p.enc[3] * r[i] + p.enc[4] * gg[i] + p.enc[5] * b[i];
                    const T vn = p.enc[6] * r[i] + p.enc[7] * gg[i] + p.enc[8] * b[i];
                    u[i] = (T)0.5 * (un + u[i]);
                    v[i] = (T)0.5 * (vn + v[i]);
                }
            }
            st4(ys + x, y);
            st4(us + x, u);
            st4(vs + x, v);
        }
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) cta_fill_tail<T, 1>(sm + (size_t)k * 3 * N1 + N1, (size_t)N1, 2, N1, W, N1);
    __syncthreads();
    const FiltHdr &fpre = p.filt[QF_PRE_LP];
    for_each_iir_task<T, TEAMS>(fpre, 2 * g.count, sm - 128, [&](int t, const IirTeam<T> &tm) {
        T *buf = sm + (size_t)(t >> 1) * 3 * N1 + (1 + (t & 1)) * N1;
        team_iir<T, 1, TEAMS>(p.tab + fpre.off, fpre, [&](int q, int, int) { return buf[q]; },
                       [&](int j, T v) { buf[j] = v; }, tm);
    });
    __syncthreads();
    T rs, rc;
    Real<T>::sincos_turns(p.phases[QP_STEP1X], rs, rc);
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k, line = io.y0 + row;
        const T *ys = sm + (size_t)k * 3 * N1, *us = ys + N1, *vs = us + N1;
        const unsigned long long ph0 = start_phase(p, g.frame, line);
        const bool neg = (p.flags & 1) && is_alternate(p, g.frame, line);
        for (int q = threadIdx.x; q < W4; q += blockDim.x) {
            const int x = 4 * q;
            T y[4], u[4], v[4], s[4], c[4], o[4];
            ld4(ys + x, y);
            ld4(us + x, u);
            ld4(vs + x, v);
            carrier4(ph0 + (unsigned long long)x * p.phases[QP_STEP1X], rs, rc, s, c);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = y[i] + (s[i] * u[i] + c[i] * (neg ? -v[i] : v[i]));
            store_comp4(io, ((size_t)g.fidx * io.nrows + row) * p.Wc + x, o);
        }
    }
}

// Encoder for u8 frames of narrow lines: the same chain, one row at a time per CTA of two warps, the RGB words of the
// row the CTA works on next (and of its field neighbour for the ColorAveraging front end) prefetched into registers
// while the current row is filtered — the global-load latency was 24 % of this kernel's warp-stall samples.
// Grid: x = rows of a field handled round-robin, y = field, z = frame.
template <typename T>
__global__ void __launch_bounds__(CM_ROW_THREADS, 8)
k_qam_encode_row(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;
    constexpr int kQ = 3;                                   // quads per thread: W <= 4 * 3 * 64 = 768
    const int W = p.W, N1 = p.n1p, W4 = W >> 2;
    const int warp = threadIdx.x >> 5;
    const bool avg = (p.flags & 2) != 0;
    const int field = blockIdx.y, f = blockIdx.z;
    const long long frame = io.first_frame + f;
    const int first = io.out_begin + field, nout = (io.out_count - field + 1) >> 1;
    T *ys = sm, *us = ys + N1, *vs = us + N1;
    uint32_t wc[kQ][3], wn[kQ][3];
    auto fetch = [&](int row) {
        const int nrow = (row + 2 < io.nrows) ? row + 2 : row;
        const uint32_t *a = reinterpret_cast<const uint32_t *>(io.in_u8 + ((size_t)f * io.nrows + row) * W * 3);
        const uint32_t *b = reinterpret_cast<const uint32_t *>(io.in_u8 + ((size_t)f * io.nrows + nrow) * W * 3);
#pragma unroll
        for (int j = 0; j < kQ; ++j) {
            const int q = threadIdx.x + j * CM_ROW_THREADS;
            if (q < W4) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    wc[j][i] = __ldg(a + 3 * q + i);
                    if (avg) wn[j][i] = __ldg(b + 3 * q + i);
                }
            }
        }
    };
    auto unpack = [&](const uint32_t *w, T *r, T *g, T *b) {
        unsigned char bytes[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bytes[i] = (w[0] >> (8 * i)) & 0xff;
            bytes[4 + i] = (w[1] >> (8 * i)) & 0xff;
            bytes[8 + i] = (w[2] >> (8 * i)) & 0xff;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[i] = Real<T>::from_u8(bytes[3 * i]);
            g[i] = Real<T>::from_u8(bytes[3 * i + 1]);
            b[i] = Real<T>::from_u8(bytes[3 * i + 2]);
        }
    };
    T rs, rc;
    Real<T>::sincos_turns(p.phases[QP_STEP1X], rs, rc);
    int k = blockIdx.x;
    if (k < nout) fetch(first + 2 * k);
    for (; k < nout; k += gridDim.x) {
        const int row = first + 2 * k, line = io.y0 + row;
#pragma unroll
        for (int j = 0; j < kQ; ++j) {
            const int q = threadIdx.x + j * CM_ROW_THREADS;
            if (q < W4) {
                T r[4], gg[4], b[4], y[4], u[4], v[4];
                unpack(wc[j], r, gg, b);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    y[i] = p.enc[0] * r[i] + p.enc[1] * gg[i] + p.enc[2] * b[i];
                    u[i] = p.enc[3] * r[i] + p.enc[4] * gg[i] + p.enc[5] * b[i];
                    v[i] = p.enc[6] * r[i] + p.enc[7] * gg[i] + p.enc[8] * b[i];
                }
                if (avg) {
                    unpack(wn[j], r, gg, b);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const T un = p.enc[3] * r[i] + p.enc[4] * gg[i] + p.enc[5] * b[i];
                        const T vn = p.enc[6] * r[i] + p.enc[7] * gg[i] + p.enc[8] * b[i];
                        u[i] = (T)0.5 * (un + u[i]);
                        v[i] = (T)0.5 * (vn + v[i]);
                    }
                }
                st4(ys + 4 * q, y);
                st4(us + 4 * q, u);
                st4(vs + 4 * q, v);
            }
        }
        __syncthreads();
        if (k + (int)gridDim.x < nout) fetch(first + 2 * (k + gridDim.x));       // in flight during the filtering
        {
            const FiltHdr &fpre = p.filt[QF_PRE_LP];
            T *buf = warp ? vs : us;
            warp_fill_tail<T, 1>(buf, N1, W, N1);
            warp_iir<T, 1>(p.tab + fpre.off, fpre, [&](int q, int, int) { return buf[q]; }, [&](int j, T x) { buf[j] = x; });
        }
        __syncthreads();
        const unsigned long long ph0 = start_phase(p, frame, line);
        const bool neg = (p.flags & 1) && is_alternate(p, frame, line);
        for (int q = threadIdx.x; q < W4; q += blockDim.x) {
            const int x = 4 * q;
            T y[4], u[4], v[4], s[4], c[4], o[4];
            ld4(ys + x, y);
            ld4(us + x, u);
            ld4(vs + x, v);
            carrier4(ph0 + (unsigned long long)x * p.phases[QP_STEP1X], rs, rc, s, c);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = y[i] + (s[i] * u[i] + c[i] * (neg ? -v[i] : v[i]));
            store_comp4(io, ((size_t)f * io.nrows + row) * p.Wc + x, o);
        }
        __syncthreads();
    }
}

// Encoder for u8 frames, second generation: the same chain as k_qam_encode_row for lines of up to 2048 samples.  The two
// chroma low-passes are run by two teams of NW / 2 warps as packed DF-I recursions (team_iir_pk); KQ pixel quads per
// thread (RGB words of the next row, and of its field neighbour for the ColorAveraging front end, prefetched in registers).
//   GEO 1: 2 warps, 3 quads (<= 768 samples);  2: 4 warps, 3 quads (<= 1536);  3: 4 warps, 4 quads, longer chunks (<= 2048)
#define QF_ENC_PRE 9     // DevParams::filt slot of this kernel's low-pass site (cm_api.cu: plan_encode_kernel)
// (EncGeo: cm_common.cuh — the NIIR row encoder shares the geometries)

template <typename T, int GEO>
__global__ void __launch_bounds__(32 * EncGeo<GEO>::NW, sizeof(T) == 8 ? 1 : 16 / EncGeo<GEO>::NW)
k_qam_encode_row2(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;
    typedef EncGeo<GEO> EG;
    constexpr int kQ = EG::KQ, NT = 32 * EG::NW, TH = EG::NW / 2;
    const int W = p.W, N1 = p.n1p, W4 = W >> 2;
    const int warp = threadIdx.x >> 5, task = warp / TH, wr = warp - task * TH;
    const bool avg = (p.flags & 2) != 0;
    const int field = blockIdx.y, f = blockIdx.z;
    const long long frame = io.first_frame + f;
    const int first = io.out_begin + field, nout = (io.out_count - field + 1) >> 1;
    const FiltHdr &fpre = p.filt[QF_ENC_PRE];
    T *ys = sm, *us = ys + N1, *vs = us + N1;
    uint32_t wc[kQ][3], wn[kQ][3];
    auto fetch = [&](int row) {
        const int nrow = (row + 2 < io.nrows) ? row + 2 : row;
        const uint32_t *a = reinterpret_cast<const uint32_t *>(io.in_u8 + ((size_t)f * io.nrows + row) * W * 3);
        const uint32_t *b = reinterpret_cast<const uint32_t *>(io.in_u8 + ((size_t)f * io.nrows + nrow) * W * 3);
#pragma unroll
        for (int j = 0; j < kQ; ++j) {
            const int q = threadIdx.x + j * NT;
            if (q < W4) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    wc[j][i] = __ldg(a + 3 * q + i);
                    if (avg) wn[j][i] = __ldg(b + 3 * q + i);
                }
            }
        }
    };
    auto unpack = [&](const uint32_t *w, T *r, T *g, T *b) {
        unsigned char bytes[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bytes[i] = (w[0] >> (8 * i)) & 0xff;
            bytes[4 + i] = (w[1] >> (8 * i)) & 0xff;
            bytes[8 + i] = (w[2] >> (8 * i)) & 0xff;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[i] = Real<T>::from_u8(bytes[3 * i]);
            g[i] = Real<T>::from_u8(bytes[3 * i + 1]);
            b[i] = Real<T>::from_u8(bytes[3 * i + 2]);
        }
    };
    T rs, rc;
    Real<T>::sincos_turns(p.phases[QP_STEP1X], rs, rc);
    int k = blockIdx.x;
    if (k < nout) fetch(first + 2 * k);
    for (; k < nout; k += gridDim.x) {
        const int row = first + 2 * k, line = io.y0 + row;
#pragma unroll
        for (int j = 0; j < kQ; ++j) {
            const int q = threadIdx.x + j * NT;
            if (q < W4) {
                T r[4], gg[4], b[4], y[4], u[4], v[4];
                unpack(wc[j], r, gg, b);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    y[i] = p.enc[0] * r[i] + p.enc[1] * gg[i] + p.enc[2] * b[i];
                    u[i] = p.enc[3] * r[i] + p.enc[4] * gg[i] + p.enc[5] * b[i];
                    v[i] = p.enc[6] * r[i] + p.enc[7] * gg[i] + p.enc[8] * b[i];
                }
                if (avg) {
                    unpack(wn[j], r, gg, b);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const T un = p.enc[3] * r[i] + p.enc[4] * gg[i] + p.enc[5] * b[i];
                        const T vn = p.enc[6] * r[i] + p.enc[7] * gg[i] + p.enc[8] * b[i];
                        u[i] = (T)0.5 * (un + u[i]);
                        v[i] = (T)0.5 * (vn + v[i]);
                    }
                }
                st4(ys + 4 * q, y);
                st4(us + 4 * q, u);
                st4(vs + 4 * q, v);
            }
        }
        __syncthreads();
        if (k + (int)gridDim.x < nout) fetch(first + 2 * (k + gridDim.x));       // in flight during the filtering
        {
            T *buf = task ? vs : us;
            warp_fill_tail<T, 1>(buf, N1, W, iir_tail_end(fpre));  // every warp of the team writes the same values
            team_iir_pk<T, 1, EG::PRE, TH>(p.tab + fpre.off, fpre, LoadLinear<T, EG::PRE>{buf}, [&](int j, T x) { buf[j] = x; },
                                           wr, 2 + task, scratch + 32 * task);
        }
        __syncthreads();
        const unsigned long long ph0 = start_phase(p, frame, line);
        const bool neg = (p.flags & 1) && is_alternate(p, frame, line);
        for (int q = threadIdx.x; q < W4; q += NT) {
            const int x = 4 * q;
            T y[4], u[4], v[4], s[4], c[4], o[4];
            ld4(ys + x, y);
            ld4(us + x, u);
            ld4(vs + x, v);
            carrier4_fast(ph0 + (unsigned long long)x * p.phases[QP_STEP1X], rs, rc, s, c);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = y[i] + (s[i] * u[i] + c[i] * (neg ? -v[i] : v[i]));
            store_comp4(io, ((size_t)f * io.nrows + row) * p.Wc + x, o);
        }
        __syncthreads();
    }
}

// Re-modulation of (u, v) through the encoder and subtraction from the composite, then colour matrix + store:
//   y = c - (sin(phi) u_lp + cos(phi) (+-v_lp))            comb.py:52-53, pal.py:225-226
template <typename T>
__device__ __forceinline__ void remod_store_row(const DevParams<T> &p, const IoArgs<T> &io, const RowGroup &g, int k,
                                                const T *c, const T *ulp, const T *vlp, const T *u, const T *v,
                                                T rs, T rc) {
    const int row = g.r0 + 2 * k, line = io.y0 + row;
    const unsigned long long ph0 = start_phase(p, g.frame, line);
    const bool neg = (p.flags & 1) && is_alternate(p, g.frame, line);
    for (int q = threadIdx.x; q < (p.W >> 2); q += blockDim.x) {
        const int x = 4 * q;
        T cc[4], a[4], b[4], uu[4], vv[4], s[4], co[4], y[4];
        ld4(c + x, cc);
        ld4(ulp + x, a);
        ld4(vlp + x, b);
        ld4(u + x, uu);
        ld4(v + x, vv);
        carrier4(ph0 + (unsigned long long)x * p.phases[QP_STEP1X], rs, rc, s, co);
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = cc[i] - (s[i] * a[i] + co[i] * (neg ? -b[i] : b[i]));
        store_rgb4(p, io, g.fidx, row, x, y, uu, vv);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Band-split decode of independent rows                       qam.py:43-58, ntsc.py:47-49, pal.py:54-59
//   luma_mode 0: luma = down2(BS(up2 c))        (strip_chroma=True; NtscModem / PalSModem; field-top rows of
//                                                NtscCombModem / PalDModem, comb.py:48-49)
//   luma_mode 1: luma = c - remod(u, v)         (field-top rows of Pal3DModem, pal.py:191-202,225-226)
//   luma_mode 2: luma = c                       (strip_chroma=False: the reset-branch return value of the comb
//                                                decoders, CM_MODE_BANDSPLIT_NOSTRIP)
// 2 warps per row.  smem: scratch[256] (IIR team scratch) + R * (c[N1] + 4 x [N2])
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool TEAMS>
__global__ void __launch_bounds__(CM_NTHREADS, 2)
k_qam_bandsplit(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io, int luma_mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, W2 = 2 * W, N1 = p.n1p, hb = p.hb2, N2 = 2 * hb;
    T *rows = sm + CM_TAPS_ELEMS, *scratch = sm + 128;
    const size_t per_row = (size_t)N1 + 4 * (size_t)N2;     // c | a2 | b2 | l2 | v2
    const FirTaps<T> hup{p.firc[QR_UP2], p.fircp[QR_UP2]}, hdn{p.firc[QR_DOWN2], p.fircp[QR_DOWN2]};   // constant bank (kernel parameter)
    load_comp_rows(io, g.fidx, g.count, W, [&](int k) { return rows + k * per_row; },
                   [&](int k) { return g.r0 + 2 * k; });
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {
        T *c = rows + k * per_row, *a2 = c + N1;
        fir_up2(a2, a2 + hb, c, W, hup, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    cta_fill_tail<T, 2>(rows + N1, per_row, g.count, hb, W2, N2);
    __syncthreads();
    // IIR phase 1: band-pass a2 -> b2, band-stop a2 -> l2
    {
        const FiltHdr &fbp = p.filt[QF_BP2X], &fbs = p.filt[QF_BS2X];
        for_each_iir_task<T, TEAMS>(fbp, g.count, scratch, [&](int k, const IirTeam<T> &tm) {
            T *c = rows + k * per_row;
            const T *ae = c + N1, *ao = ae + hb;
            T *be = c + N1 + N2, *bo = be + hb;
            team_iir<T, 2, TEAMS>(p.tab + fbp.off, fbp, [&](int q, int ph, int) { return (ph ? ao : ae)[q]; },
                           Poly2Out<T>{be, bo}, tm);
        });
        if (luma_mode == 0)
            for_each_iir_task<T, TEAMS>(fbs, g.count, scratch, [&](int k, const IirTeam<T> &tm) {
                T *c = rows + k * per_row;
                const T *ae = c + N1, *ao = ae + hb;
                T *le = c + N1 + 2 * N2, *lo = le + hb;
                team_iir<T, 2, TEAMS>(p.tab + fbs.off, fbs, [&](int q, int ph, int) { return (ph ? ao : ae)[q]; },
                               Poly2Out<T>{le, lo}, tm);
            }, g.count);
    }
    __syncthreads();
    cta_fill_tail<T, 2>(rows + N1 + N2, per_row, g.count, hb, W2, N2);      // b2 feeds the low-pass stage
    __syncthreads();
    // IIR phase 2: product demodulation + low-pass.  u2 -> a2 (dead), v2 -> 4th buffer
    for_each_iir_task<T, TEAMS>(p.filt[QF_DEMOD_LP], 2 * g.count, scratch, [&](int t, const IirTeam<T> &tm) {
        const int k = t >> 1;
        T *c = rows + k * per_row;
        const T *be = c + N1 + N2, *bo = be + hb;
        T *de = (t & 1) ? (c + N1 + 3 * N2) : (c + N1), *dod = de + hb;
        const int line = io.y0 + g.r0 + 2 * k;
        // sin for u, cos for v: cos(x) = sin(x + 1/4 turn)
        Carrier<T> car(start_phase(p, g.frame, line) + p.phases[QP_BP_SHIFT] + ((t & 1) ? CM_QUARTER_TURN : 0ull),
                       p.phases[QP_STEP2X], W2);
        const FiltHdr &f = p.filt[QF_DEMOD_LP];
        team_iir<T, 2, TEAMS>(p.tab + f.off, f,
                       [&](int q, int ph, int i) {
                           car.at(2 * q + ph, i);
                           return (T)2 * car.s * (ph ? bo : be)[q];
                       },
                       Poly2Out<T>{de, dod}, tm);
    });
    __syncthreads();
    // down2 of u2, v2 -> u, v at 1x into the b2 region (dead now): u at [0, N1), v at [N1, 2 N1)
    for (int k = 0; k < g.count; ++k) {
        T *c = rows + k * per_row;
        T *uo = c + N1 + N2, *vo = uo + N1;
        const bool alt = (p.flags & 1) && is_alternate(p, g.frame, io.y0 + g.r0 + 2 * k);
        fir_down2(c + N1, c + N1 + hb, W, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *y) { st4(uo + j0, y); });
        fir_down2(c + N1 + 3 * N2, c + N1 + 3 * N2 + hb, W, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *y) {
            T v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = alt ? -y[i] : y[i];
            st4(vo + j0, v);
        });
    }
    __syncthreads();
    if (luma_mode == 0) {
        for (int k = 0; k < g.count; ++k) {
            T *c = rows + k * per_row;
            const T *uo = c + N1 + N2, *vo = uo + N1;
            const int row = g.r0 + 2 * k;
            fir_down2(c + N1 + 2 * N2, c + N1 + 2 * N2 + hb, W, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *y) {
                T u[4], v[4];
                ld4(uo + j0, u);
                ld4(vo + j0, v);
                store_rgb4(p, io, g.fidx, row, j0, y, u, v);
            });
        }
        return;
    }
    if (luma_mode == 2) {
        for (int k = 0; k < g.count; ++k) {
            const T *c = rows + k * per_row;
            for (int q = threadIdx.x; q < (W >> 2); q += blockDim.x) {
                T y[4], u[4], v[4];
                ld4(c + 4 * q, y);
                ld4(c + N1 + N2 + 4 * q, u);
                ld4(c + N1 + N2 + N1 + 4 * q, v);
                store_rgb4(p, io, g.fidx, g.r0 + 2 * k, 4 * q, y, u, v);
            }
        }
        return;
    }
    // luma_mode 1: encoder pre-lowpass of (u, v) into the a2 region, then re-modulate and subtract
    for (int k = 0; k < g.count; ++k) cta_fill_tail<T, 1>(rows + k * per_row + N1 + N2, (size_t)N1, 2, N1, W, N1);
    __syncthreads();
    const FiltHdr &fpre = p.filt[QF_PRE_LP];
    for_each_iir_task<T, TEAMS>(fpre, 2 * g.count, scratch, [&](int t, const IirTeam<T> &tm) {
        T *c = rows + (t >> 1) * per_row;
        const T *src = c + N1 + N2 + (t & 1) * N1;
        T *dst = c + N1 + (t & 1) * N1;
        team_iir<T, 1, TEAMS>(p.tab + fpre.off, fpre, [&](int q, int, int) { return src[q]; },
                       [&](int j, T v) { dst[j] = v; }, tm);
    });
    __syncthreads();
    T rs, rc;
    Real<T>::sincos_turns(p.phases[QP_STEP1X], rs, rc);
    for (int k = 0; k < g.count; ++k) {
        const T *c = rows + k * per_row;
        remod_store_row(p, io, g, k, c, c + N1, c + N1 + N1, c + N1 + N2, c + N1 + N2 + N1, rs, rc);
    }
}

// ------------------------------------------------------------------------------------------------------------
// PAL-D delay-line decode of rows that have a predecessor        pal.py:79-127 + comb.py:50-53
// Per row k (k = -1 is the halo row):  G[k] = up2(down2(BP(up2 c[k])))     (extract_chroma, then the up2 of
// _demodulate_am; all linear, so G of the sum/difference of two rows is the sum/difference of their G's)
//   S = down2(LP(sin(ph)       * (G[k] + G[k-1])))      D = down2(LP(sin(ph + pi/2) * (G[k] - G[k-1])))
//   u = D sin(LS/2) + S cos(LS/2);  v = D cos(LS/2) - S sin(LS/2);  v = -v on alternate lines
//   y = c - (sin(phi) LP_pre(u) + cos(phi) LP_pre(+-v))
// smem: scratch[256] (IIR team scratch) + (R+1) * (c[N1] + G[N2]) + 2R * [N2] work
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool TEAMS>
__global__ void __launch_bounds__(CM_NTHREADS, 2)
k_pald_combed(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    PhaseClock pc(io.prof);
    const int W = p.W, W2 = 2 * W, N1 = p.n1p, hb = p.hb2, N2 = 2 * hb;
    const int R = io.rows_per_cta;
    T *scratch = sm + 128;
    T *cbuf = sm + CM_TAPS_ELEMS;                          // (R+1) x N1    composite rows, index k+1
    T *gbuf = cbuf + (size_t)(R + 1) * N1;       // (R+1) x N2    G rows, index k+1
    T *work = gbuf + (size_t)(R + 1) * N2;       // 2R x N2
    const int nin = g.count + 1;
    const FirTaps<T> hup{p.firc[QR_UP2], p.fircp[QR_UP2]}, hdn{p.firc[QR_DOWN2], p.fircp[QR_DOWN2]};   // constant bank (kernel parameter)
    load_comp_rows(io, g.fidx, nin, W, [&](int k) { return cbuf + (size_t)k * N1; },
                   [&](int k) { return g.r0 + 2 * (k - 1); });
    __syncthreads();
    pc.mark();
    for (int k = 0; k < nin; ++k) {
        T *b = gbuf + (size_t)k * N2;
        fir_up2(b, b + hb, cbuf + (size_t)k * N1, W, hup, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    pc.mark();
    cta_fill_tail<T, 2>(gbuf, (size_t)N2, nin, hb, W2, N2);
    __syncthreads();
    pc.mark();
    for_each_iir_task<T, TEAMS>(p.filt[QF_BP2X], nin, scratch, [&](int t, const IirTeam<T> &tm) {   // band-pass in place
        T *be = gbuf + (size_t)t * N2, *bo = be + hb;
        const FiltHdr &f = p.filt[QF_BP2X];
        team_iir<T, 2, TEAMS>(p.tab + f.off, f, [&](int q, int ph, int) { return (ph ? bo : be)[q]; },
                       Poly2Out<T>{be, bo}, tm);
    });
    __syncthreads();
    pc.mark();
    for (int k = 0; k < nin; ++k) {                     // E = down2(b2) -> work[k][0..W)
        T *e = work + (size_t)k * N1;
        const T *b = gbuf + (size_t)k * N2;
        fir_down2(b, b + hb, W, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *y) { st4(e + j0, y); });
    }
    __syncthreads();
    pc.mark();
    for (int k = 0; k < nin; ++k) {                     // G = up2(E)
        T *b = gbuf + (size_t)k * N2;
        fir_up2(b, b + hb, work + (size_t)k * N1, W, hup, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    pc.mark();
    cta_fill_tail<T, 2>(gbuf, (size_t)N2, nin, hb, W2, N2);
    __syncthreads();
    pc.mark();
    for_each_iir_task<T, TEAMS>(p.filt[QF_PALD_LP], 2 * g.count, scratch, [&](int t, const IirTeam<T> &tm) {
        // AM demodulation low-pass of sum / difference
        const int k = t >> 1;
        const T *gce = gbuf + (size_t)(k + 1) * N2, *gco = gce + hb;
        const T *gle = gbuf + (size_t)k * N2, *glo = gle + hb;
        T *de = work + (size_t)t * N2, *dod = de + hb;
        const int line = io.y0 + g.r0 + 2 * k;
        const T sgn = (t & 1) ? (T)-1 : (T)1;
        Carrier<T> car(start_phase(p, g.frame, line) + p.phases[QP_BP_SHIFT] - p.phases[QP_HALF_LS] +
                           ((t & 1) ? CM_QUARTER_TURN : 0ull),
                       p.phases[QP_STEP2X], W2);
        const FiltHdr &f = p.filt[QF_PALD_LP];
        team_iir<T, 2, TEAMS>(p.tab + f.off, f,
                       [&](int q, int ph, int i) {
                           car.at(2 * q + ph, i);
                           return Real<T>::fma_(sgn, (ph ? glo : gle)[q], (ph ? gco : gce)[q]) * car.s;
                       },
                       Poly2Out<T>{de, dod}, tm);
    });
    __syncthreads();
    pc.mark();
    // S, D at 1x, rotated to (u, v) with the V switch (pal.py:121-125), into the gbuf rows (G is dead):
    // u at [k][0..N1), v at [k][N1..2 N1)
    const T sf = p.scalars[QS_PALD_SIN], cf = p.scalars[QS_PALD_COS];
    for (int k = 0; k < g.count; ++k) {
        T *uo = gbuf + (size_t)k * N2, *vo = uo + N1;
        const T *ws = work + (size_t)(2 * k) * N2, *wd = work + (size_t)(2 * k + 1) * N2;
        const T vsgn = is_alternate(p, g.frame, io.y0 + g.r0 + 2 * k) ? (T)-1 : (T)1;
        fir_down2_pair(ws, ws + hb, wd, wd + hb, W, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *s, const T *d) {
            T u[4], v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                u[i] = d[i] * sf + s[i] * cf;
                v[i] = vsgn * (d[i] * cf - s[i] * sf);
            }
            st4(uo + j0, u);
            st4(vo + j0, v);
        });
    }
    __syncthreads();
    pc.mark();
    for (int k = 0; k < g.count; ++k) cta_fill_tail<T, 1>(gbuf + (size_t)k * N2, (size_t)N1, 2, N1, W, N1);
    __syncthreads();
    pc.mark();
    const FiltHdr &fpre = p.filt[QF_PRE_LP];
    for_each_iir_task<T, TEAMS>(fpre, 2 * g.count, scratch, [&](int t, const IirTeam<T> &tm) {
        // encoder pre-lowpass for the re-modulation
        const T *src = gbuf + (size_t)(t >> 1) * N2 + (t & 1) * N1;
        T *dst = work + (size_t)t * N1;
        team_iir<T, 1, TEAMS>(p.tab + fpre.off, fpre, [&](int q, int, int) { return src[q]; },
                       [&](int j, T v) { dst[j] = v; }, tm);
    });
    __syncthreads();
    pc.mark();
    T rs, rc;
    Real<T>::sincos_turns(p.phases[QP_STEP1X], rs, rc);
    for (int k = 0; k < g.count; ++k)
        remod_store_row(p, io, g, k, cbuf + (size_t)(k + 1) * N1, work + (size_t)(2 * k) * N1,
                        work + (size_t)(2 * k + 1) * N1, gbuf + (size_t)k * N2, gbuf + (size_t)k * N2 + N1, rs, rc);
    pc.mark();
}

// ------------------------------------------------------------------------------------------------------------
// PAL-D in two passes over independent rows (the path taken when every IIR use-site fits one super-chunk, i.e.
// 720-sample lines).  The AM demodulation of the line sum / difference (pal.py:116-119) is linear in the lines, and
// the carrier of row k is the carrier of row k-1 advanced by the line shift LS, so with the per-row quadrature pair
//     a_k = down2(LP(sin(psi_k) G_k)),   b_k = down2(LP(cos(psi_k) G_k)),   psi_k = start_phase(k) + bp_shift - LS/2
// (G_k = up2(down2(BP(up2 c_k))) as above) the sum / difference channels are
//     S_k = a_k + cos(LS) a_{k-1} + sin(LS) b_{k-1},     D_k = b_k - cos(LS) b_{k-1} + sin(LS) a_{k-1}.
// Pass 1 (k_qam_rows<PALD>, all the arithmetic: 1 row at a time per CTA of two warps, ~21 KB of shared memory, 8 CTAs per
// SM, no halo row) writes (a_k, b_k) and their encoder-low-passed copies to an fp32 scratch; pass 2 (k_qam_combine) combines
// neighbouring rows, rotates to (u, v), re-modulates through the encoder low-pass and stores RGB.
// ------------------------------------------------------------------------------------------------------------

// Common tail of the pass-1 row kernels.  wa / wb hold LP(sin X), LP(cos X) at 2x (polyphase).  Writes four planes of
// the row to the scratch  aux[frame][row][4][W]:
//     a = down2(wa),  b = down2(wb),  alpha = LPpre(a),  beta = LPpre(b)
// LPpre is the encoder's chroma low-pass (qam.py:16) that the comb decoders apply to (u, v) before re-modulating
// (comb.py:53).  (u, v) are fixed linear combinations of the (a, b) of neighbouring rows and the filter is linear, so
// filtering a and b here, where both warps of the CTA are free, leaves pass 2 purely elementwise.
template <typename T, bool TEAMS>
__device__ __forceinline__ void rows_epilogue(const DevParams<T> &p, T *__restrict__ dst, T *scratch, T *cb, T *g,
                                              T *wa, T *wb) {
    const int W = p.W, N1 = p.n1p, hb = p.hb2;
    const FirTaps<T> hdn{p.firc[QR_DOWN2], p.fircp[QR_DOWN2]};
    T *sa = cb, *sb = g;                                   // both dead by now; g holds two N1 rows
    fir_down2_pair(wa, wa + hb, wb, wb + hb, W, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *a, const T *b) {
        st4(dst + j0, a);
        st4(dst + W + j0, b);
        st4(sa + j0, a);
        st4(sb + j0, b);
    });
    __syncthreads();
    {
        const FiltHdr &fpre = p.filt[QF_PRE_LP];
        for_row_tasks<T, TEAMS>(fpre, 2, scratch, [&](int t, const IirTeam<T> &tm) {
            T *src = t ? sb : sa, *out = t ? wb : wa;
            warp_fill_tail<T, 1>(src, N1, W, N1);
            team_iir<T, 1, TEAMS>(p.tab + fpre.off, fpre, [&](int q, int, int) { return src[q]; },
                                  [&](int j, T x) { out[j] = x; }, tm);
        });
    }
    __syncthreads();
    for (int x = 4 * threadIdx.x; x < W; x += 4 * blockDim.x) {
        T v[4];
        ld4(wa + x, v);
        st4(dst + 2 * W + x, v);
        ld4(wb + x, v);
        st4(dst + 3 * W + x, v);
    }
}

// The planes of one row.  PALD: G-path of pal.py:79-127 (a, b from G = up2(down2(BP(up2 c))) through PalDModem._filter
// at phase psi - LS/2); otherwise the qam.py:43-58 path (a, b from B = BP(up2 c) through _demod_lowpass at phase psi,
// without the factor 2).  The composite row must already be staged in cb.  On return the four planes are in `dst`
// (global, plane pitch W) and in shared memory at cb (a), g (b), wa (alpha), wb (beta); the CTA is synchronised.
template <typename T, bool PALD, bool TEAMS>
__device__ __forceinline__ void row_planes(const DevParams<T> &p, const IoArgs<T> &io, int f, long long frame, int row,
                                           T *__restrict__ dst, T *scratch, T *cb, T *g, T *wa, T *wb) {
    const int W = p.W, W2 = 2 * W, hb = p.hb2, N2 = 2 * hb;
    const FirTaps<T> hup{p.firc[QR_UP2], p.fircp[QR_UP2]}, hdn{p.firc[QR_DOWN2], p.fircp[QR_DOWN2]};
    // precondition: the composite row is staged in cb and the CTA is synchronised
    fir_up2(g, g + hb, cb, W, hup, threadIdx.x, blockDim.x);
    __syncthreads();
    {
        const FiltHdr &fb = p.filt[QF_BP2X];
        for_row_tasks<T, TEAMS>(fb, 1, scratch, [&](int, const IirTeam<T> &tm) {
            warp_fill_tail<T, 2>(g, hb, W2, N2);
            T *ge = g, *go = g + hb;
            team_iir<T, 2, TEAMS>(p.tab + fb.off, fb, [&](int q, int ph, int) { return (ph ? go : ge)[q]; },
                                  Poly2Out<T>{ge, go}, tm);
        });
    }
    __syncthreads();
    if (PALD) {
        fir_down2(g, g + hb, W, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *y) { st4(cb + j0, y); });
        __syncthreads();
        fir_up2(g, g + hb, cb, W, hup, threadIdx.x, blockDim.x);
        __syncthreads();
    }
    {
        const FiltHdr &fl = p.filt[PALD ? QF_PALD_LP : QF_DEMOD_LP];
        const unsigned long long psi = start_phase(p, frame, io.y0 + row) + p.phases[QP_BP_SHIFT] -
                                       (PALD ? p.phases[QP_HALF_LS] : 0ull);
        for_row_tasks<T, TEAMS>(fl, 2, scratch, [&](int t, const IirTeam<T> &tm) {
            warp_fill_tail<T, 2>(g, hb, W2, N2);          // every warp writes the same values
            const T *ge = g, *go = g + hb;
            T *de = t ? wb : wa, *dod = de + hb;
            Carrier<T> car(psi + (t ? CM_QUARTER_TURN : 0ull), p.phases[QP_STEP2X], W2);
            team_iir<T, 2, TEAMS>(p.tab + fl.off, fl,
                                  [&](int q, int ph, int i) {
                                      car.at(2 * q + ph, i);
                                      return (ph ? go : ge)[q] * car.s;
                                  },
                                  Poly2Out<T>{de, dod}, tm);
        });
    }
    __syncthreads();
    rows_epilogue<T, TEAMS>(p, dst, scratch, cb, g, wa, wb);
    __syncthreads();
}

// Pass 1 of the two-pass decoders: a CTA works through rows blockIdx.x, blockIdx.x + gridDim.x, ... of one frame (one at
// a time; the next one is prefetched), planes to aux[frame][row][4][W].
template <typename T, bool PALD, bool TEAMS>
__global__ void __launch_bounds__(TEAMS ? CM_NTHREADS : CM_ROW_THREADS, TEAMS ? 2 : CM_ROWS_MINB)
k_qam_rows(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;      // [0, 128): IIR team scratch
    const int W = p.W, N1 = p.n1p, N2 = 2 * p.hb2;
    const int f = blockIdx.z, end = io.out_begin + io.out_count;
    T *cb = sm, *g = cb + N1, *wa = g + N2, *wb = wa + N2;
    const bool pre = io.in_u8 != nullptr && W <= 4 * RowPrefetch::kMaxQuads * (int)blockDim.x;
    RowPrefetch pf;
    int row = io.out_begin + blockIdx.x;
    if (pre && row < end) pf.fetch(io, f, row, W);
    for (; row < end; row += gridDim.x) {
        if (pre) pf.stage(cb, W);
        else load_comp_row(cb, io, f, row, W);
        __syncthreads();
        if (pre && row + (int)gridDim.x < end) pf.fetch(io, f, row + gridDim.x, W);
        row_planes<T, PALD, TEAMS>(p, io, f, io.first_frame + f, row, io.aux + ((size_t)f * io.nrows + row) * 4 * W,
                                   scratch, cb, g, wa, wb);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Pass 1, second generation (k_qam_rows2): the same planes as k_qam_rows, with
//   * every IIR stage run by ALL warps of the CTA: the single band-pass task by a team of NW warps, the task pairs
//     (sin || cos low-pass, alpha || beta pre-low-pass) by two teams of NW / 2 warps — no warp idles at a barrier while
//     another one filters (15 % of the warp-stall samples of k_qam_rows), and 1920-sample lines get the packed f32x2
//     recursion that the older multi-warp path (team_iir_L, scalar DF-II-T) lacked;
//   * the demodulating carrier taken from a row-independent table: sin / cos(2 pi j step) for the 2x sample index j, laid
//     out for coalesced loads, tail replicated (DevParams::ctab).  The row's own phase psi is applied afterwards, as a
//     rotation of the decimated pair:  a = cos(psi) a' + sin(psi) b',  b = cos(psi) b' - sin(psi) a'  (low-pass and
//     decimation are linear) — 4 operations per output sample instead of a rotating carrier with an end-of-line
//     predicate per 2x sample and channel (13 % of the instructions of k_qam_rows).
// Geometry GEO (warps per CTA, chunk lengths per lane; chosen by the host, cm_api.cu: plan_row_kernel):
//   1: 2 warps, lines up to ~760 samples;  2: 4 warps, up to ~1520;  3: 4 warps with long chunks, up to ~1970 (1920-wide
//   rasters).  Short chunks on more warps lose: the per-chunk cost of a section (warp scan, team fold, barrier) is fixed,
//   8 warps x 16 samples measured 98 us per 1080p frame against 65 of the first-generation kernel.
// ------------------------------------------------------------------------------------------------------------
template <int GEO> struct RowL;
// (every L / RATE is odd: lane t starts at element t * L / RATE of its polyphase plane, an odd stride keeps the 32 lanes of
// a load or store on 32 different banks; 24 and 32 measured 17 M and 74 M bank conflicts per 64 / 16 frames)
template <> struct RowL<1> { static constexpr int NW = 2, BP = 26, LPA = 46, LPB = 50, PRE = 23; };
template <> struct RowL<2> { static constexpr int NW = 4, BP = 26, LPA = 46, LPB = 50, PRE = 23; };
template <> struct RowL<3> { static constexpr int NW = 4, BP = 34, LPA = 62, LPB = 66, PRE = 31; };

#define QF_ROW_BP 6      // DevParams::filt slots of the row kernel's use-sites (same filters as QF_BP2X, QF_DEMOD_LP or
#define QF_ROW_LP 7      // QF_PALD_LP, QF_PRE_LP; chunk lengths and tables for its team geometry, built in cm_api.cu)
#define QF_ROW_PRE 8

#ifndef CM_ROWS2_MINB
#define CM_ROWS2_MINB 8        // CTAs per SM of the 2-warp geometry (128 registers per thread); 9 / 10 are A/B variants
#endif
template <typename T, bool PALD, int GEO>
__global__ void __launch_bounds__(32 * RowL<GEO>::NW, sizeof(T) == 8 ? 1 : (RowL<GEO>::NW == 2 ? CM_ROWS2_MINB : 4))
k_qam_rows2(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;      // [0, 128): team exchange
    typedef RowL<GEO> RL;
    constexpr int NW = RL::NW;
    constexpr int TH = NW / 2;                                              // warps per team of a task pair
    const int W = p.W, W2 = 2 * W, N1 = p.n1p, hb = p.hb2, N2 = 2 * hb;
    const int f = blockIdx.z, end = io.out_begin + io.out_count;
    const int warp = threadIdx.x >> 5;
    const int task = warp / TH, wr = warp - task * TH;
    // cb | sb: 1x rows; g | wb: 2x rows.  The sin-channel low-pass writes over its own input (wa == g: both teams load
    // their chunks, the CTA synchronises, then they filter and store), which keeps 1920-sample rows at 4 CTAs per SM
    T *cb = sm, *sb = cb + N1, *g = sb + N1, *wa = g, *wb = g + N2;
    const FirTaps<T> hup{p.firc[QR_UP2], p.fircp[QR_UP2]}, hdn{p.firc[QR_DOWN2], p.fircp[QR_DOWN2]};
    const FiltHdr &fb = p.filt[QF_ROW_BP], &fl = p.filt[QF_ROW_LP], &fpre = p.filt[QF_ROW_PRE];
    const long long frame = io.first_frame + f;
    const bool pre = io.in_u8 != nullptr && W <= 4 * RowPrefetch::kMaxQuads * (int)blockDim.x;
    RowPrefetch pf;
    int row = io.out_begin + blockIdx.x;
    if (pre && row < end) pf.fetch(io, f, row, W);
    for (; row < end; row += gridDim.x) {
        if (pre) pf.stage(cb, W);
        else load_comp_row(cb, io, f, row, W);
        __syncthreads();
        if (pre && row + (int)gridDim.x < end) pf.fetch(io, f, row + gridDim.x, W);
        T *__restrict__ dst = io.aux + ((size_t)f * io.nrows + row) * 4 * W;
        fir_up2(g, g + hb, cb, W, hup, threadIdx.x, blockDim.x);
        __syncthreads();
        // band-pass in place, all warps
        warp_fill_tail<T, 2>(g, hb, W2, iir_tail_end(fb));          // every warp writes the same values
        team_iir_pk<T, 2, RL::BP, NW>(p.tab + fb.off, fb, LoadPoly2<T, RL::BP>{g, g + hb}, Poly2Out<T>{g, g + hb}, warp, 1,
                                      scratch);
        __syncthreads();
        if (PALD) {
            fir_down2(g, g + hb, W, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *y) { st4(cb + j0, y); });
            __syncthreads();
            fir_up2(g, g + hb, cb, W, hup, threadIdx.x, blockDim.x);
            __syncthreads();
        }
        {   // a' = LP(sin(theta) X) by team 0, b' = LP(cos(theta) X) by team 1
            warp_fill_tail<T, 2>(g, hb, W2, iir_tail_end(fl));
            T *de = task ? wb : wa;
            const T *ct = p.ctab + (size_t)task * fl.npad;
            if (fl.L == RL::LPA)
                team_iir_pk<T, 2, RL::LPA, TH, true>(p.tab + fl.off, fl, LoadPoly2Carrier<T, RL::LPA>{g, g + hb, ct, 32 * TH},
                                                     Poly2Out<T>{de, de + hb}, wr, 2 + task, scratch + 32 + 32 * task);
            else
                team_iir_pk<T, 2, RL::LPB, TH, true>(p.tab + fl.off, fl, LoadPoly2Carrier<T, RL::LPB>{g, g + hb, ct, 32 * TH},
                                                     Poly2Out<T>{de, de + hb}, wr, 2 + task, scratch + 32 + 32 * task);
        }
        __syncthreads();
        {   // decimate, rotate to the row's own phase, keep (a, b) for the pre-low-pass
            T sphi, cphi;
            Real<T>::sincos_turns(start_phase(p, frame, io.y0 + row) + p.phases[QP_BP_SHIFT] -
                                      (PALD ? p.phases[QP_HALF_LS] : 0ull), sphi, cphi);
            T *sa = cb;
            fir_down2_pair(wa, wa + hb, wb, wb + hb, W, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *a0, const T *b0) {
                T a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    a[i] = Real<T>::fma_(cphi, a0[i], sphi * b0[i]);
                    b[i] = Real<T>::fma_(cphi, b0[i], -(sphi * a0[i]));
                }
                st4(dst + j0, a);
                st4(dst + W + j0, b);
                st4(sa + j0, a);
                st4(sb + j0, b);
            });
        }
        __syncthreads();
        {   // alpha = LPpre(a) by team 0, beta = LPpre(b) by team 1
            T *src = task ? sb : cb, *out = task ? wb : wa;
            warp_fill_tail<T, 1>(src, N1, W, iir_tail_end(fpre));
            team_iir_pk<T, 1, RL::PRE, TH>(p.tab + fpre.off, fpre, LoadLinear<T, RL::PRE>{src}, [&](int j, T x) { out[j] = x; },
                                           wr, 2 + task, scratch + 32 + 32 * task);
        }
        __syncthreads();
        for (int x = 4 * threadIdx.x; x < W; x += 4 * blockDim.x) {
            T v[4];
            ld4(wa + x, v);
            st4(dst + 2 * W + x, v);
            ld4(wb + x, v);
            st4(dst + 3 * W + x, v);
        }
        __syncthreads();
    }
}

// Band-split decode (qam.py:43-58 with strip_chroma=True: NtscModem, PalSModem) of one row per CTA of two warps:
// the two IIR stages are pairs of independent tasks (band-pass || band-stop, then the u || v low-pass), so both warps
// are busy in every phase; 26.5 KB of shared memory, 8 CTAs per SM.
template <typename T, bool TEAMS>
__global__ void __launch_bounds__(TEAMS ? CM_NTHREADS : CM_ROW_THREADS, TEAMS ? 2 : CM_ROWS_MINB)
k_qam_bs_row(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;      // [0, 128): IIR team scratch
    const int W = p.W, W2 = 2 * W, N1 = p.n1p, hb = p.hb2, N2 = 2 * hb;
    const int row = io.out_begin + blockIdx.x, f = blockIdx.z;
    const long long frame = io.first_frame + f;
    T *cb = sm;                                       // N1: composite row
    T *a2 = cb + N1, *b2 = a2 + N2, *l2 = b2 + N2, *v2 = l2 + N2;
    T *u2 = a2;                                       // up2(c) is dead once both filters have read it
    const FirTaps<T> hup{p.firc[QR_UP2], p.fircp[QR_UP2]}, hdn{p.firc[QR_DOWN2], p.fircp[QR_DOWN2]};
    load_comp_row(cb, io, f, row, W);
    __syncthreads();
    fir_up2(a2, a2 + hb, cb, W, hup, threadIdx.x, blockDim.x);
    __syncthreads();
    {
        const FiltHdr &fbp = p.filt[QF_BP2X], &fbs = p.filt[QF_BS2X];
        const T *ae = a2, *ao = a2 + hb;
        if constexpr (!TEAMS) {
            // one call site for both filters (warp 0: band-pass -> b2, warp 1: band-stop -> l2) keeps the code small
            const int t = threadIdx.x >> 5;
            const FiltHdr &ff = t ? fbs : fbp;
            T *de = t ? l2 : b2;
            warp_fill_tail<T, 2>(a2, hb, W2, N2);          // both warps write the same values
            warp_iir<T, 2>(p.tab + ff.off, ff, [&](int q, int ph, int) { return (ph ? ao : ae)[q]; },
                           Poly2Out<T>{de, de + hb});
        } else {
            for_row_tasks<T, TEAMS>(fbp, 1, scratch, [&](int, const IirTeam<T> &tm) {
                warp_fill_tail<T, 2>(a2, hb, W2, N2);
                team_iir<T, 2, TEAMS>(p.tab + fbp.off, fbp, [&](int q, int ph, int) { return (ph ? ao : ae)[q]; },
                                      Poly2Out<T>{b2, b2 + hb}, tm);
            });
            for_row_tasks<T, TEAMS>(fbs, 1, scratch, [&](int, const IirTeam<T> &tm) {
                warp_fill_tail<T, 2>(a2, hb, W2, N2);
                team_iir<T, 2, TEAMS>(p.tab + fbs.off, fbs, [&](int q, int ph, int) { return (ph ? ao : ae)[q]; },
                                      Poly2Out<T>{l2, l2 + hb}, tm);
            }, 1);
        }
    }
    __syncthreads();
    {
        const FiltHdr &fl = p.filt[QF_DEMOD_LP];
        for_row_tasks<T, TEAMS>(fl, 2, scratch, [&](int t, const IirTeam<T> &tm) {
            warp_fill_tail<T, 2>(b2, hb, W2, N2);
            const T *be = b2, *bo = b2 + hb;
            T *de = t ? v2 : u2, *dod = de + hb;
            Carrier<T> car(start_phase(p, frame, io.y0 + row) + p.phases[QP_BP_SHIFT] + (t ? CM_QUARTER_TURN : 0ull),
                           p.phases[QP_STEP2X], W2);
            team_iir<T, 2, TEAMS>(p.tab + fl.off, fl,
                                  [&](int q, int ph, int i) {
                                      car.at(2 * q + ph, i);
                                      return (T)2 * car.s * (ph ? bo : be)[q];
                                  },
                                  Poly2Out<T>{de, dod}, tm);
        });
    }
    __syncthreads();
    const bool alt = (p.flags & 1) && is_alternate(p, frame, io.y0 + row);
    for (int j0 = 4 * threadIdx.x; j0 < W; j0 += 4 * blockDim.x) {
        T y[4], u[4], v[4];
        down2_quad(hdn, u2, u2 + hb, W, j0, u);
        down2_quad(hdn, v2, v2 + hb, W, j0, v);
        down2_quad(hdn, l2, l2 + hb, W, j0, y);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = alt ? -v[i] : v[i];
        store_rgb4(p, io, f, row, j0, y, u, v);
    }
}

// Band-split decode, second generation (k_qam_bs_row2; NtscModem / PalSModem, BASELINE configs[0]): the chain of
// k_qam_bs_row with the techniques of k_qam_rows2 — packed DF-I teams, band-pass || band-stop and u || v low-pass as task
// pairs with one of each pair in place, the demodulating carrier from the row-independent table with the row's phase (and
// the factor 2 of qam.py:50-51) applied as a rotation after the decimation.  One chunk length serves the three use-sites
// (host: plan_row_kernel).  smem: scratch[128] | cb[N1] | a2[N2] (-> luma) | b2[N2] (-> u) | v2[N2]
#define QF_BSROW_BP 6        // DevParams::filt slots (kind QAM_BANDSPLIT only; the comb kinds use them for k_qam_rows2)
#define QF_BSROW_LP 7
#define QF_BSROW_BS 8
template <typename T, int GEO, int L>
__device__ __forceinline__ void bs_row2_body(const DevParams<T> &p, const IoArgs<T> &io, T *scratch, T *sm) {
    typedef RowL<GEO> RL;
    constexpr int NW = RL::NW, TH = NW / 2, NT = 32 * NW;
    const int W = p.W, W2 = 2 * W, N1 = p.n1p, hb = p.hb2, N2 = 2 * hb;
    const int f = blockIdx.z, end = io.out_begin + io.out_count;
    const int warp = threadIdx.x >> 5, task = warp / TH, wr = warp - task * TH;
    T *cb = sm, *a2 = cb + N1, *b2 = a2 + N2, *v2 = b2 + N2;
    const FirTaps<T> hup{p.firc[QR_UP2], p.fircp[QR_UP2]}, hdn{p.firc[QR_DOWN2], p.fircp[QR_DOWN2]};
    const FiltHdr &fbp = p.filt[QF_BSROW_BP], &fbs = p.filt[QF_BSROW_BS], &fl = p.filt[QF_BSROW_LP];
    const long long frame = io.first_frame + f;
    const bool pre = io.in_u8 != nullptr && W <= 4 * RowPrefetch::kMaxQuads * NT;
    RowPrefetch pf;
    int row = io.out_begin + blockIdx.x;
    if (pre && row < end) pf.fetch(io, f, row, W);
    for (; row < end; row += gridDim.x) {
        if (pre) pf.stage(cb, W);
        else load_comp_row(cb, io, f, row, W);
        __syncthreads();
        if (pre && row + (int)gridDim.x < end) pf.fetch(io, f, row + gridDim.x, W);
        fir_up2(a2, a2 + hb, cb, W, hup, threadIdx.x, NT);
        __syncthreads();
        warp_fill_tail<T, 2>(a2, hb, W2, max(iir_tail_end(fbp), iir_tail_end(fbs)));      // every warp writes the same values
        if (task == 0)                                           // band-pass a2 -> b2 || band-stop a2 -> a2 (= luma at 2x)
            team_iir_pk<T, 2, L, TH, true>(p.tab + fbp.off, fbp, LoadPoly2<T, L>{a2, a2 + hb}, Poly2Out<T>{b2, b2 + hb}, wr, 2, scratch);
        else
            team_iir_pk<T, 2, L, TH, true>(p.tab + fbs.off, fbs, LoadPoly2<T, L>{a2, a2 + hb}, Poly2Out<T>{a2, a2 + hb}, wr, 3,
                                           scratch + 32);
        __syncthreads();
        {   // u' = LP(sin(theta) B) (team 0, over b2) || v' = LP(cos(theta) B) (team 1 -> v2)
            warp_fill_tail<T, 2>(b2, hb, W2, iir_tail_end(fl));
            T *de = task ? v2 : b2;
            const T *ct = p.ctab + (size_t)task * fl.npad;
            team_iir_pk<T, 2, L, TH, true>(p.tab + fl.off, fl, LoadPoly2Carrier<T, L>{b2, b2 + hb, ct, 32 * TH},
                                           Poly2Out<T>{de, de + hb}, wr, 2 + task, scratch + 32 * task);
        }
        __syncthreads();
        T sphi, cphi;
        Real<T>::sincos_turns(start_phase(p, frame, io.y0 + row) + p.phases[QP_BP_SHIFT], sphi, cphi);
        sphi *= (T)2;                                            // qam.py:50-51: 2 sin, 2 cos
        cphi *= (T)2;
        const bool alt = (p.flags & 1) && is_alternate(p, frame, io.y0 + row);
        for (int j0 = 4 * threadIdx.x; j0 < W; j0 += 4 * NT) {
            T y[4], a0[4], b0[4], u[4], v[4];
            down2_quad(hdn, b2, b2 + hb, W, j0, a0);
            down2_quad(hdn, v2, v2 + hb, W, j0, b0);
            down2_quad(hdn, a2, a2 + hb, W, j0, y);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                u[i] = Real<T>::fma_(cphi, a0[i], sphi * b0[i]);
                const T vv = Real<T>::fma_(cphi, b0[i], -(sphi * a0[i]));
                v[i] = alt ? -vv : vv;
            }
            store_rgb4(p, io, f, row, j0, y, u, v);
        }
        __syncthreads();
    }
}

template <typename T, int GEO>
__global__ void __launch_bounds__(32 * RowL<GEO>::NW, sizeof(T) == 8 ? 1 : 16 / RowL<GEO>::NW)
k_qam_bs_row2(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;
    if (p.filt[QF_BSROW_LP].L == RowL<GEO>::LPA) bs_row2_body<T, GEO, RowL<GEO>::LPA>(p, io, scratch, sm);
    else bs_row2_body<T, GEO, RowL<GEO>::LPB>(p, io, scratch, sm);
}

// QamColorModem.extract_chroma (qam.py:34-37) of one row per CTA: E = down2(BP(up2 c)), stored as the first output plane
// (float output only; the L0 kit call of the drop-in boundary, not on the frame path).
template <typename T>
__global__ void __launch_bounds__(CM_NTHREADS)
k_qam_extract(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;
    const int W = p.W, W2 = 2 * W, N1 = p.n1p, hb = p.hb2, N2 = 2 * hb;
    const int row = io.out_begin + blockIdx.x, f = blockIdx.z;
    T *cb = sm, *g = cb + N1;
    const FirTaps<T> hup{p.firc[QR_UP2], p.fircp[QR_UP2]}, hdn{p.firc[QR_DOWN2], p.fircp[QR_DOWN2]};
    load_comp_row(cb, io, f, row, W);
    __syncthreads();
    fir_up2(g, g + hb, cb, W, hup, threadIdx.x, blockDim.x);
    __syncthreads();
    cta_fill_tail<T, 2>(g, (size_t)N2, 1, hb, W2, N2);
    __syncthreads();
    if (threadIdx.x < 32) {
        const FiltHdr &fb = p.filt[QF_BP2X];
        T *ge = g, *go = g + hb;
        warp_iir<T, 2>(p.tab + fb.off, fb, [&](int q, int ph, int) { return (ph ? go : ge)[q]; }, Poly2Out<T>{ge, go});
    }
    __syncthreads();
    T *dst = io.out_f + ((size_t)f * io.nrows + row) * W * 3;
    fir_down2(g, g + hb, W, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *y) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dst[3 * (j0 + i)] = y[i];
            dst[3 * (j0 + i) + 1] = (T)0;
            dst[3 * (j0 + i) + 2] = (T)0;
        }
    });
}

// Pass 2 (elementwise): combine the planes of neighbouring rows of a field into (u, v) and the low-passed (u, v) the
// re-modulation needs, y = c - remod, inverse matrix, store.  A thread owns 4 consecutive samples of CM_SEG consecutive
// rows of one field and walks down the rows keeping the previous rows' planes in registers, so every plane is read
// ~once; the carrier of the next row is the carrier of this row rotated by LS.
// With ch/sh = cos/sin(LS/2), cl/sl = cos/sin(LS) and D[.] = down2(LP(.)):
//     D[sin(psi_k + t) B_k] = cos(t) a_k + sin(t) b_k,     D[cos(psi_k + t) B_k] = cos(t) b_k - sin(t) a_k,
// and psi_{k+1} = psi_k + LS.
//   PAIR_PALD   pal.py:113-125   S = a_k + D[sin(psi_k) G_{k-1}],  D = b_k - D[cos(psi_k) G_{k-1}]
//   PAIR_NTSC2  ntsc.py:74-81    phase psi_k - LS/2 on the line difference, u from cos, v from -sin
//   PAIR_NTSC3  comb.py:96-113   average of the 2-line chroma of this row (band-split chroma at a field top) and of
//                                the next row (nothing at the bottom, where the driver re-feeds the row, image.py:51-53)
//   PAIR_PAL3   pal.py:198-226   sums / differences of three rows at the phase of the middle one
enum { PAIR_PALD = 0, PAIR_NTSC2 = 1, PAIR_NTSC3 = 2, PAIR_PAL3 = 3 };
#ifndef CM_SEG
#define CM_SEG 8       // rows of a field per thread of the combine pass
#endif

template <typename T>
struct PairCoef {
    T sh, ch, sl, cl, sf, cf, f2, a_ss, a_cu, a_cv;
};

// (u, v) of one sample from the (a, b) of the previous / current / next row of the field
// comb.py:13-15: the estimate of smaller magnitude where the two agree in sign (signbit, so -0 counts as negative),
// zero where they do not
template <typename T>
__device__ __forceinline__ T minavg_(T a, T b) {
    const T sign = ((T)1 - (signbit(a) ? (T)1 : (T)0)) - (signbit(b) ? (T)1 : (T)0);
    const T aa = Real<T>::abs_(a), ab = Real<T>::abs_(b);
    return sign * (aa < ab ? aa : ab);
}

// mn: the two estimates of the 3-line decoders are combined by comb.minavg instead of their mean (CM_FLAG_MINAVG; the
// host then passes the PAL factors without the 0.5 of comb.avg)
template <typename T, int MODE, bool mn>
__device__ __forceinline__ void pair_uv(const PairCoef<T> &k, bool hp, bool hn, bool alt, T ap, T bp, T ac, T bc, T an,
                                        T bn, T &u, T &v) {
    if (MODE == PAIR_PALD) {
        const T s = ac + (k.cl * ap + k.sl * bp);
        const T d = bc - (k.cl * bp - k.sl * ap);
        u = d * k.sf + s * k.cf;
        v = d * k.cf - s * k.sf;
        v = alt ? -v : v;
    } else if (MODE == PAIR_NTSC2) {
        u = k.f2 * ((k.ch * bc + k.sh * ac) - (k.ch * bp - k.sh * ap));
        v = -k.f2 * ((k.ch * ac - k.sh * bc) - (k.ch * ap + k.sh * bp));
    } else if (MODE == PAIR_NTSC3) {
        const T u0 = hp ? k.f2 * ((k.ch * bc + k.sh * ac) - (k.ch * bp - k.sh * ap)) : (T)2 * ac;
        const T v0 = hp ? -k.f2 * ((k.ch * ac - k.sh * bc) - (k.ch * ap + k.sh * bp)) : (T)2 * bc;
        // at the field bottom the re-fed row gives a zero line difference: (+0) * f, (+0) * (-f)   (ntsc.py:78-81)
        const T u1 = hn ? k.f2 * ((k.ch * bn + k.sh * an) - (k.ch * bc - k.sh * ac)) : (T)0;
        const T v1 = hn ? -k.f2 * ((k.ch * an - k.sh * bn) - (k.ch * ac + k.sh * bc)) : -(T)0;
        if (mn) {
            u = minavg_(u0, u1);
            v = minavg_(v0, v1);
        } else {
            u = (T)0.5 * (u0 + (hn ? u1 : (T)0));
            v = (T)0.5 * (v0 + (hn ? v1 : (T)0));
        }
    } else {
        const T sin_n = hn ? k.cl * an - k.sl * bn : ac, cos_n = hn ? k.cl * bn + k.sl * an : bc;
        const T sin_p = k.cl * ap + k.sl * bp, cos_p = k.cl * bp - k.sl * ap;
        const T us = k.a_ss * (cos_n - cos_p), ud = k.a_cu * (sin_n - (T)2 * ac + sin_p);
        const T vs = k.a_ss * (sin_n - sin_p), vd = k.a_cv * (cos_n - (T)2 * bc + cos_p);
        if (mn) {
            u = minavg_(us, ud);
            v = minavg_(vs, vd);
        } else {
            u = us + ud;
            v = vs + vd;
        }
        v = alt ? -v : v;
    }
}
// OUT: 0 = RGB; 1 = (y, u, v) to io.yuv (luma notch follows); 2 = (composite, u, v) to io.yuv with comb.minavg
// (no minimum-blocks bound: with one the compiler spends registers freely and the 3-line modes lose an occupancy step —
// 2.22 vs 1.90 us/frame; CM_COMBINE_MINB is the A/B hook of tools/variants.sh)
template <typename T, int MODE, int OUT>
#ifdef CM_COMBINE_MINB
__global__ void __launch_bounds__(128, sizeof(T) == 4 ? CM_COMBINE_MINB : 1)
#else
__global__ void __launch_bounds__(128)
#endif
k_qam_combine(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    const int W = p.W, W4 = W >> 2;
    const int field = blockIdx.y, f = blockIdx.z;
    const int first = io.out_begin + field;
    const int rows_in_field = (io.out_count - field + 1) >> 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = idx / W4, q = idx - seg * W4;
    const int k0 = seg * CM_SEG;
    if (k0 >= rows_in_field) return;
    const int k1 = min(k0 + CM_SEG, rows_in_field);
    const int x = 4 * q;
    const long long frame = io.first_frame + f;
    PairCoef<T> kc;
    Real<T>::sincos_turns(p.phases[QP_HALF_LS], kc.sh, kc.ch);
    Real<T>::sincos_turns(p.line_shift, kc.sl, kc.cl);
    kc.sf = p.scalars[QS_PALD_SIN];
    kc.cf = p.scalars[QS_PALD_COS];
    kc.f2 = (T)2 * p.scalars[QS_NTSC_FACTOR];
    kc.a_ss = (T)2 * p.scalars[QS_P3D_SINSUM];
    kc.a_cu = (T)2 * p.scalars[QS_P3D_COSU];
    kc.a_cv = (T)2 * p.scalars[QS_P3D_COSV];
    const T *aux = io.aux + (size_t)f * io.nrows * 4 * W + x;
    // planes of a row: [0] a, [1] b, [2] alpha, [3] beta
    T P[4][4], C[4][4], Nx[4][4];
    auto ldrow = [&](int row, T (*dst)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ld4(aux + ((size_t)row * 4 + j) * W, dst[j]);
    };
    int row = first + 2 * k0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) P[j][i] = Nx[j][i] = (T)0;
    if (row >= 2) ldrow(row - 2, P);
    ldrow(row, C);
    T rs, rc, s[4], co[4];
    Real<T>::sincos_turns(p.phases[QP_STEP1X], rs, rc);
    carrier4(start_phase(p, frame, io.y0 + row) + (unsigned long long)x * p.phases[QP_STEP1X], rs, rc, s, co);
    for (int k = k0; k < k1; ++k, row += 2) {
        const bool hp = row >= 2, hn = row + 2 < io.nrows;
        const bool need_next = (MODE >= PAIR_NTSC3) ? hn : (k + 1 < k1);
        if (need_next) ldrow(row + 2, Nx);
        const int line = io.y0 + row;
        const bool alt = is_alternate(p, frame, line);
        const bool neg = (p.flags & 1) && alt;
        constexpr bool mn = OUT == 2 && MODE >= PAIR_NTSC3;
        T cc[4], y[4], u[4], v[4];
        const size_t cbase = ((size_t)f * io.nrows + row) * W + x;
        if (io.in_f) {
            ld4(io.in_f + cbase, cc);
        } else {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(io.in_u8 + cbase));
#pragma unroll
            for (int i = 0; i < 4; ++i) cc[i] = ((T)5 * Real<T>::from_u8((w >> (8 * i)) & 0xff) - (T)1) * (T)(1.0 / 3.0);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            T ul, vl;
            pair_uv<T, MODE, mn>(kc, hp, hn, alt, P[0][i], P[1][i], C[0][i], C[1][i], Nx[0][i], Nx[1][i], u[i], v[i]);
            pair_uv<T, MODE, false>(kc, hp, hn, alt, P[2][i], P[3][i], C[2][i], C[3][i], Nx[2][i], Nx[3][i], ul, vl);
            y[i] = cc[i] - (s[i] * ul + co[i] * (neg ? -vl : vl));
        }
        // comb.minavg is not linear: the encoder low-pass of (u, v) cannot be taken from the pre-filtered planes, so the
        // row goes to k_finish_rows as (composite, u, v) through io.yuv
        if (OUT == 0) store_rgb4_direct(p, io, f, row, x, y, u, v);
        else store_yuv4(p, io, f, row, x, mn ? cc : y, u, v);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) { P[j][i] = C[j][i]; C[j][i] = Nx[j][i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {          // carrier of the next row of the field: advance by LS
            const T ns = Real<T>::fma_(s[i], kc.cl, co[i] * kc.sl);
            co[i] = Real<T>::fma_(co[i], kc.cl, -(s[i] * kc.sl));
            s[i] = ns;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Finishing pass of the comb decoders for the non-default knobs: rows whose (y, u, v) the decoders above diverted to
// io.yuv (store_rgb4) are completed here, one row per CTA of two warps.
//   REMOD (avg=comb.minavg, comb.py:13-15): the y plane holds the composite; (u, v) go through the encoder low-pass
//         (one warp each) and y = c - remod(u, v)                                   comb.py:104-107, pal.py:225-226
//   notch (notch=Q, CM_FLAG_NOTCH): y is filtered along the row by one biquad at fsc, zero initial state
//                                                                 comb.py:18-20, 54-55, 108-109, pal.py:227-228
// then the inverse matrix and the store.
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool REMOD>
__global__ void __launch_bounds__(CM_ROW_THREADS, 8)
k_finish_rows(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *y = reinterpret_cast<T *>(smem_raw);              // REMOD: + ulp[N1] | vlp[N1]
    const int W = p.Wo, N1 = p.n1p;
    const int row = io.out_begin + blockIdx.x, f = blockIdx.z;
    const long long frame = io.first_frame + f;
    const T *src = io.yuv + ((size_t)f * io.nrows + row) * 3 * W;
    T *ulp = y + N1, *vlp = ulp + N1;
    for (int x = 4 * threadIdx.x; x < W; x += 4 * blockDim.x) {
        T v[4];
        ld4(src + x, v);
        st4(y + x, v);
        if (REMOD) {
            ld4(src + W + x, v);
            st4(ulp + x, v);
            ld4(src + 2 * W + x, v);
            st4(vlp + x, v);
        }
    }
    __syncthreads();
    if (REMOD) {
        const FiltHdr &fpre = p.filt[QF_PRE_LP];
        T *buf = (threadIdx.x >> 5) ? vlp : ulp;
        warp_fill_tail<T, 1>(buf, N1, W, N1);
        warp_iir<T, 1>(p.tab + fpre.off, fpre, [&](int q, int, int) { return buf[q]; }, [&](int j, T x) { buf[j] = x; });
        __syncthreads();
        T rs, rc;
        Real<T>::sincos_turns(p.phases[QP_STEP1X], rs, rc);
        const int line = io.y0 + row;
        const unsigned long long ph0 = start_phase(p, frame, line);
        const bool neg = (p.flags & 1) && is_alternate(p, frame, line);
        for (int x = 4 * threadIdx.x; x < W; x += 4 * blockDim.x) {
            T cc[4], a[4], b[4], s[4], co[4];
            ld4(y + x, cc);
            ld4(ulp + x, a);
            ld4(vlp + x, b);
            carrier4(ph0 + (unsigned long long)x * p.phases[QP_STEP1X], rs, rc, s, co);
#pragma unroll
            for (int i = 0; i < 4; ++i) cc[i] = cc[i] - (s[i] * a[i] + co[i] * (neg ? -b[i] : b[i]));
            st4(y + x, cc);
        }
        __syncthreads();
    }
    if ((p.flags & CM_FLAG_NOTCH) && threadIdx.x < 32) {
        const FiltHdr &fn = p.filt[QF_NOTCH];
        warp_fill_tail<T, 1>(y, N1, W, N1);
        warp_iir<T, 1>(p.tab + fn.off, fn, [&](int q, int, int) { return y[q]; }, [&](int j, T v) { y[j] = v; });
    }
    __syncthreads();
    IoArgs<T> out = io;
    out.yuv = nullptr;
    for (int x = 4 * threadIdx.x; x < W; x += 4 * blockDim.x) {
        T yy[4], u[4], v[4];
        ld4(y + x, yy);
        ld4(src + W + x, u);
        ld4(src + 2 * W + x, v);
        store_rgb4(p, out, f, row, x, yy, u, v);
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
// smem: scratch[256] (IIR team scratch) + (R+2) * (c[N1] + B[N2]) + 2R * [N2]
// ------------------------------------------------------------------------------------------------------------
enum { COMB_NTSC2 = 0, COMB_NTSC3 = 1, COMB_PAL3 = 2 };

template <typename T, int MODE, bool TEAMS>
__global__ void __launch_bounds__(CM_NTHREADS, 2)
k_qam_comb(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, W2 = 2 * W, N1 = p.n1p, hb = p.hb2, N2 = 2 * hb;
    const int R = io.rows_per_cta;
    T *scratch = sm + 128;
    T *cbuf = sm + CM_TAPS_ELEMS;                          // (R+2) x N1    index k+1, k = -1 .. R
    T *bbuf = cbuf + (size_t)(R + 2) * N1;       // (R+2) x N2
    T *work = bbuf + (size_t)(R + 2) * N2;       // 2R x N2
    const bool has_prev0 = g.r0 >= 2;                               // row k = -1 exists
    const bool has_next_last = g.r0 + 2 * g.count < io.nrows;       // row k = count exists
    const int k_lo = has_prev0 ? -1 : 0;
    const int k_hi = (MODE != COMB_NTSC2 && has_next_last) ? g.count : g.count - 1;
    const int nin = k_hi - k_lo + 1;
    const FirTaps<T> hup{p.firc[QR_UP2], p.fircp[QR_UP2]}, hdn{p.firc[QR_DOWN2], p.fircp[QR_DOWN2]};   // constant bank (kernel parameter)
    load_comp_rows(io, g.fidx, nin, W, [&](int k) { return cbuf + (size_t)(k_lo + k + 1) * N1; },
                   [&](int k) { return g.r0 + 2 * (k_lo + k); });
    __syncthreads();
    for (int k = k_lo; k <= k_hi; ++k) {
        T *b = bbuf + (size_t)(k + 1) * N2;
        fir_up2(b, b + hb, cbuf + (size_t)(k + 1) * N1, W, hup, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    cta_fill_tail<T, 2>(bbuf + (size_t)(k_lo + 1) * N2, (size_t)N2, nin, hb, W2, N2);
    __syncthreads();
    for_each_iir_task<T, TEAMS>(p.filt[QF_BP2X], nin, scratch, [&](int t, const IirTeam<T> &tm) {
        T *be = bbuf + (size_t)(k_lo + t + 1) * N2, *bo = be + hb;
        const FiltHdr &f = p.filt[QF_BP2X];
        team_iir<T, 2, TEAMS>(p.tab + f.off, f, [&](int q, int ph, int) { return (ph ? bo : be)[q]; },
                       Poly2Out<T>{be, bo}, tm);
    });
    __syncthreads();
    cta_fill_tail<T, 2>(bbuf + (size_t)(k_lo + 1) * N2, (size_t)N2, nin, hb, W2, N2);
    __syncthreads();
    const unsigned long long step = p.phases[QP_STEP2X];
    const FiltHdr &flp = p.filt[QF_DEMOD_LP];
    for_each_iir_task<T, TEAMS>(flp, 2 * g.count, scratch, [&](int t, const IirTeam<T> &tm) {
        const int k = t >> 1;
        const bool is_v = (t & 1) != 0;
        const int line = io.y0 + g.r0 + 2 * k;
        const bool hp = (k > 0) || has_prev0;
        const bool hn = (k + 1 < g.count) || has_next_last;
        const T *bce = bbuf + (size_t)(k + 1) * N2, *bco = bce + hb;
        const T *bpe = bbuf + (size_t)(hp ? k : k + 1) * N2, *bpo = bpe + hb;
        const T *bne = bbuf + (size_t)(hn ? k + 2 : k + 1) * N2, *bno = bne + hb;
        T *de = work + (size_t)t * N2, *dod = de + hb;
        auto st = Poly2Out<T>{de, dod};
        const unsigned long long psi = start_phase(p, g.frame, line) + p.phases[QP_BP_SHIFT];
        if (MODE == COMB_NTSC2) {
            const T f2 = (T)2 * p.scalars[QS_NTSC_FACTOR];
            // u: f 2 cos(phi) d = f2 sin(phi + 1/4) d;   v: -f 2 sin(phi) d
            Carrier<T> car(psi - p.phases[QP_HALF_LS] + (is_v ? 0ull : CM_QUARTER_TURN), step, W2);
            const T amp = is_v ? -f2 : f2;
            team_iir<T, 2, TEAMS>(p.tab + flp.off, flp,
                           [&](int q, int ph, int i) {
                               car.at(2 * q + ph, i);
                               return amp * car.s * ((ph ? bco : bce)[q] - (ph ? bpo : bpe)[q]);
                           },
                           st, tm);
        } else if (MODE == COMB_NTSC3) {
            const T f = p.scalars[QS_NTSC_FACTOR];      // 0.5 * (2 f ...) = f ...
            const unsigned long long quarter = is_v ? 0ull : CM_QUARTER_TURN;
            Carrier<T> car(hp ? (psi - p.phases[QP_HALF_LS] + quarter) : (psi + (is_v ? CM_QUARTER_TURN : 0ull)),
                           step, W2);
            Carrier<T> carn(start_phase(p, g.frame, line + 2) + p.phases[QP_BP_SHIFT] - p.phases[QP_HALF_LS] + quarter,
                            step, W2);
            const T amp = is_v ? -f : f;
            team_iir<T, 2, TEAMS>(p.tab + flp.off, flp,
                           [&](int q, int ph, int i) {
                               const int j = 2 * q + ph;
                               car.at(j, i);
                               const T bc = (ph ? bco : bce)[q];
                               T acc = hp ? amp * car.s * (bc - (ph ? bpo : bpe)[q]) : car.s * bc;
                               if (hn) {
                                   carn.at(j, i);
                                   acc = Real<T>::fma_(amp * carn.s, (ph ? bno : bne)[q] - bc, acc);
                               }
                               return acc;
                           },
                           st, tm);
        } else {
            const T a_ss = (T)2 * p.scalars[QS_P3D_SINSUM];
            const T a_c = (T)2 * (is_v ? p.scalars[QS_P3D_COSV] : p.scalars[QS_P3D_COSU]);
            Carrier<T> car(psi, step, W2);
            team_iir<T, 2, TEAMS>(p.tab + flp.off, flp,
                           [&](int q, int ph, int i) {
                               car.at(2 * q + ph, i);
                               const T bc = (ph ? bco : bce)[q];
                               const T curr_diff = (ph ? bno : bne)[q] - bc, last_diff = bc - (ph ? bpo : bpe)[q];
                               const T ssig = curr_diff + last_diff, dsig = curr_diff - last_diff;
                               return is_v ? (a_ss * car.s * ssig + a_c * car.c * dsig)
                                           : (a_ss * car.c * ssig + a_c * car.s * dsig);
                           },
                           st, tm);
        }
    });
    __syncthreads();
    // u, v at 1x into bbuf rows (the B's are dead): u at [k+1][0..N1), v at [k+1][N1..2 N1)
    for (int t = 0; t < 2 * g.count; ++t) {
        const int k = t >> 1;
        T *o = bbuf + (size_t)(k + 1) * N2 + (t & 1) * N1;
        const T *w = work + (size_t)t * N2;
        const bool neg = (MODE == COMB_PAL3) && (t & 1) && is_alternate(p, g.frame, io.y0 + g.r0 + 2 * k);
        fir_down2(w, w + hb, W, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *y) {
            T v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = neg ? -y[i] : y[i];
            st4(o + j0, v);
        });
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) cta_fill_tail<T, 1>(bbuf + (size_t)(k + 1) * N2, (size_t)N1, 2, N1, W, N1);
    __syncthreads();
    const FiltHdr &fpre = p.filt[QF_PRE_LP];
    for_each_iir_task<T, TEAMS>(fpre, 2 * g.count, scratch, [&](int t, const IirTeam<T> &tm) {
        const T *src = bbuf + (size_t)((t >> 1) + 1) * N2 + (t & 1) * N1;
        T *dst = work + (size_t)t * N1;
        team_iir<T, 1, TEAMS>(p.tab + fpre.off, fpre, [&](int q, int, int) { return src[q]; },
                       [&](int j, T v) { dst[j] = v; }, tm);
    });
    __syncthreads();
    T rs, rc;
    Real<T>::sincos_turns(p.phases[QP_STEP1X], rs, rc);
    for (int k = 0; k < g.count; ++k)
        remod_store_row(p, io, g, k, cbuf + (size_t)(k + 1) * N1, work + (size_t)(2 * k) * N1,
                        work + (size_t)(2 * k + 1) * N1, bbuf + (size_t)(k + 1) * N2, bbuf + (size_t)(k + 1) * N2 + N1,
                        rs, rc);
}
