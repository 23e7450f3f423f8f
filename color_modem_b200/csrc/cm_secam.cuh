// SECAM kernels.  Reference: color_modem/color/secam.py.
//
// Encode (secam.py:261-276, 240-246): line-sequential D'R / D'B -> low-pass -> optional LF pre-emphasis ->
// instantaneous frequency (clipped) -> FM synthesis with the complex "bell" gain G(f) evaluated per sample and a
// running phase  start - angle(G[0]) + sum_{i=1..j} pi f[i].  The running phase is accumulated in 0.64
// fixed-point turns with an integer prefix sum over the line (per-lane chunk sums + warp-shuffle scan), so it
// cannot drift: the centre frequency enters as an exact integer constant and only the deviation goes through
// floating point.
//
// Decode (secam.py:278-304, 127-149): Bessel band-stop -> luma; flipped warm-up prefix + Chebyshev band-pass
// (+ anti-bell) -> quadrature FM discriminator at 2x (mix to baseband, low-pass I and Q, wrapped phase
// difference of consecutive analytic samples == diff(unwrap(angle)), down2) -> clip -> scale -> optional
// de-emphasis; each output row pairs its own colour-difference signal with the one of the previous row of the
// field (zeros at the top).  The previous row's signal is recomputed by the CTA as a halo row.
#pragma once
#include "cm_common.cuh"
#include "cm_fir.cuh"
#include "cm_iir.cuh"
#include "cm_io.cuh"
#include "cm_slots.h"

template <typename T> struct Fix64;
template <> struct Fix64<float> {
    static __device__ __forceinline__ long long half_turns(float f) { return __float2ll_rn(f * 9.2233720368547758e18f); }
};
template <> struct Fix64<double> {
    static __device__ __forceinline__ long long half_turns(double f) { return __double2ll_rn(f * 9.2233720368547758e18); }
};

// secam.py:248-256 — 625-line numbers are hard-coded in the reference regardless of the line standard
template <typename T>
__device__ __forceinline__ bool secam_phase_inverted(const DevParams<T> &p, long long frame, int line) {
    const int fr = (int)(frame % 6);
    const int half = line >> 1;                                   // floor, also for negative lines
    const int lif = ((line & 1) == 0 ? 23 : 336) + half;
    int seq = (fr * 625 + lif) % 6;
    if (seq < 0) seq += 6;
    const bool inv = (p.phases[SP_INVERSIONS] >> seq) & 1ull;
    return inv ^ ((fr & 1) == 1);
}

// ------------------------------------------------------------------------------------------------------------
// Encode.  One warp per row.   smem: R * 2 * N1   (luma | selected chroma -> composite)
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void secam_bell(T f, T f0, T inv_f0, T m0, T kn, T kd, T &re, T &im) {
    // G = m0 (1 + j kn F) / (1 + j kd F),  F = f/f0 - f0/f      (secam.py:241-243)
    const T F = f * inv_f0 - f0 / f;
    const T g = m0 / ((T)1 + kd * kd * F * F);
    re = g * ((T)1 + kn * kd * F * F);
    im = g * (kn - kd) * F;
}

template <typename T, int L>
__device__ __forceinline__ void secam_fm_row(const DevParams<T> &p, const T *__restrict__ chroma, T *__restrict__ luma,
                                             int W, bool alt, bool inverted) {
    const int lane = threadIdx.x & 31;
    const T fsc = alt ? p.scalars[SS_FSC_DB] : p.scalars[SS_FSC_DR];
    const T fdev = alt ? p.scalars[SS_FDEV_DB] : p.scalars[SS_FDEV_DR];
    const T dlo = p.scalars[SS_F_LO] - fsc, dhi = p.scalars[SS_F_HI] - fsc;
    const T f0 = p.scalars[SS_BELL_F0], inv_f0 = (T)1 / f0;
    const T m0 = p.scalars[SS_M0], kn = p.scalars[SS_KN], kd = p.scalars[SS_KD];
    const unsigned long long centre = alt ? p.phases[SP_FSC_DB_HALF] : p.phases[SP_FSC_DR_HALF];
    // phase[j] = start - angle(G[0]) + sum_{i=1..j} pi f[i]   (secam.py:244-245), in 0.64 fixed-point turns
    unsigned long long run = inverted ? 0x8000000000000000ull : 0ull;     // start phase: pi or 0 (secam.py:273)
    {
        T d0 = fdev * chroma[0];
        d0 = d0 < dlo ? dlo : (d0 > dhi ? dhi : d0);
        T re, im;
        secam_bell(fsc + d0, f0, inv_f0, m0, kn, kd, re, im);
        run -= (unsigned long long)Fix64<T>::half_turns(Real<T>::atan2_(im, re) * (T)0.31830988618379067154);
    }
    const int nsuper = (W + 32 * L - 1) / (32 * L);
    for (int c = 0; c < nsuper; ++c) {
        const int base = (c * 32 + lane) * L;
        T d[L];
        unsigned long long acc[L];
        unsigned long long sum = 0;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const int j = base + i;
            T dev = fdev * chroma[j < W ? j : W - 1];
            dev = dev < dlo ? dlo : (dev > dhi ? dhi : dev);
            d[i] = dev;
            // advance of sample j: pi*f rad = f half-turns; sample 0 contributes nothing
            if (j > 0 && j < W) sum += centre + (unsigned long long)Fix64<T>::half_turns(dev);
            acc[i] = sum;
        }
        unsigned long long incl = sum;                                    // inclusive scan of the lane totals
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, 1 << k);
            if (lane >= (1 << k)) incl += v;
        }
        const unsigned long long before = run + incl - sum;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const int j = base + i;
            if (j < W) {
                T re, im, s, co;
                secam_bell(fsc + d[i], f0, inv_f0, m0, kn, kd, re, im);
                Real<T>::sincos_turns(before + acc[i], s, co);
                luma[j] += re * co - im * s;
            }
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

template <typename T>
__global__ void __launch_bounds__(CM_NTHREADS)
k_secam_encode(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, N1 = p.n1p, W4 = W >> 2;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool avg = (p.flags & 2) != 0;
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const int nrow = (row + 2 < io.nrows) ? row + 2 : row;
        const bool alt = is_alternate(p, g.frame, io.y0 + row);
        const int ci = alt ? 6 : 3;                         // D'B (row 2 of the matrix) on alternate lines, else D'R
        T *ys = sm + (size_t)k * 2 * N1, *cs = ys + N1;
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
            st4(ys + x, y);
            st4(cs + x, c);
        }
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) cta_fill_tail<T, 1>(sm + (size_t)k * 2 * N1 + N1, (size_t)N1, 1, N1, W, N1);
    __syncthreads();
    for (int k = warp; k < g.count; k += nwarps) {
        T *ys = sm + (size_t)k * 2 * N1, *cs = ys + N1;
        const int line = io.y0 + g.r0 + 2 * k;
        const FiltHdr &f = p.filt[SF_PRE_LP];
        warp_iir<T, 1>(p.tab + f.off, f, [&](int q, int, int) { return cs[q]; }, [&](int j, T v) { cs[j] = v; });
        if (p.flags & 128) {
            warp_fill_tail<T, 1>(cs, N1, W, N1);
            const FiltHdr &fe = p.filt[SF_PRE_EMPH];
            warp_iir<T, 1>(p.tab + fe.off, fe, [&](int q, int, int) { return cs[q]; }, [&](int j, T v) { cs[j] = v; });
        }
        __syncwarp();
        secam_fm_row<T, 23>(p, cs, ys, W, is_alternate(p, g.frame, line), secam_phase_inverted(p, g.frame, line));
    }
    __syncthreads();
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const T *ys = sm + (size_t)k * 2 * N1;
        for (int q = threadIdx.x; q < W4; q += blockDim.x) {
            T o[4];
            ld4(ys + 4 * q, o);
            store_comp4(io, ((size_t)g.fidx * io.nrows + row) * p.Wc + 4 * q, o);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Decode, per row (no coupling between rows).  smem: taps/scratch[256] + R rows x ( c[N1] | cc[N1] | U2[N2] | I2[N2] | Q2[N2] )
//   c  : composite, then luma (band-stop in place)
//   cc : warm-up prefix + composite -> band-passed chroma -> later the colour-difference signal X
//   U2 : up2(chroma), later the instantaneous frequency at 2x
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool TEAMS>
__global__ void __launch_bounds__(CM_NTHREADS, 2)
k_secam_decode(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, N1 = p.n1p, hb = p.hb2, N2 = 2 * hb;
    const int pre = W / 40 - 1;                 // samples of flipped warm-up (secam.py:283)
    const int ncc = W + pre, ncc4 = (ncc + 3) & ~3, n2 = 2 * ncc;
    T *scratch = sm + 128;
    T *rows = sm + CM_TAPS_ELEMS;
    const size_t per_row = 2 * (size_t)N1 + 3 * (size_t)N2;
    const int k_lo = 0;                          // rows are independent here: pairing with the previous row happens
    const int nin = g.count;                     // in k_pair_rows_store
    const FirTaps<T> hup{p.firc[SR_UP2], p.fircp[SR_UP2]}, hdn{p.firc[SR_DOWN2], p.fircp[SR_DOWN2]};   // constant bank (kernel parameter)
    auto rowp = [&](int k) { return rows + (size_t)k * per_row; };
    load_comp_rows(io, g.fidx, nin, W, [&](int k) { return rowp(k); }, [&](int k) { return g.r0 + 2 * k; });
    __syncthreads();
    for (int k = k_lo; k < g.count; ++k) {
        const T *c = rowp(k);
        T *cc = rowp(k) + N1;
        for (int i = threadIdx.x; i < N1; i += blockDim.x) {
            T v;
            if (i < pre) v = c[pre - i];                     // flip(composite[1 : W/40])
            else if (i < ncc) v = c[i - pre];
            else v = c[W - 1];                               // replicated tail for the IIR
            cc[i] = v;
        }
        if (k >= 0) {
            T *cm = rowp(k);
            for (int i = W + threadIdx.x; i < N1; i += blockDim.x) cm[i] = c[W - 1];
        }
    }
    __syncthreads();
    // IIR phase 1: luma band-stop (rows k >= 0, in place) and chroma band-pass (+ anti-bell) on cc (all rows)
    {
        const FiltHdr &fl = p.filt[SF_LUMA_BS], &fc = p.filt[SF_CHROMA_BP];
        for_each_iir_task<T, TEAMS>(fl, g.count, scratch, [&](int t, const IirTeam<T> &tm) {
            T *c = rowp(t);
            team_iir<T, 1, TEAMS>(p.tab + fl.off, fl, [&](int q, int, int) { return c[q]; },
                                  [&](int j, T v) { c[j] = v; }, tm);
        });
        for_each_iir_task<T, TEAMS>(fc, nin, scratch, [&](int t, const IirTeam<T> &tm) {
            T *cc = rowp(k_lo + t) + N1;
            team_iir<T, 1, TEAMS>(p.tab + fc.off, fc, [&](int q, int, int) { return cc[q]; },
                                  [&](int j, T v) { cc[j] = v; }, tm);
        }, g.count);
    }
    __syncthreads();
    if (p.flags & 64) {
        cta_fill_tail<T, 1>(rows + N1, per_row, nin, N1, ncc, N1);
        __syncthreads();
        const FiltHdr &fb = p.filt[SF_ANTI_BELL];
        for_each_iir_task<T, TEAMS>(fb, nin, scratch, [&](int t, const IirTeam<T> &tm) {
            T *cc = rowp(k_lo + t) + N1;
            team_iir<T, 1, TEAMS>(p.tab + fb.off, fb, [&](int q, int, int) { return cc[q]; },
                                  [&](int j, T v) { cc[j] = v; }, tm);
        });
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < nin * (ncc4 - ncc); idx += blockDim.x) {   // zero pad for the resampler
        const int r = idx / (ncc4 - ncc), i = ncc + idx - r * (ncc4 - ncc);
        rows[(size_t)r * per_row + N1 + i] = (T)0;
    }
    __syncthreads();
    for (int k = k_lo; k < g.count; ++k) {
        T *u2 = rowp(k) + 2 * N1;
        fir_up2(u2, u2 + hb, rowp(k) + N1, ncc4, hup, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    cta_fill_tail<T, 2>(rows + 2 * N1, per_row, nin, hb, n2, N2);
    __syncthreads();
    // IIR phase 2: mix to baseband and low-pass: I = LP(x cos), Q = LP(x sin)        secam.py:137-142
    const FiltHdr &flp = p.filt[SF_FM_LP];
    for_each_iir_task<T, TEAMS>(flp, 2 * nin, scratch, [&](int t, const IirTeam<T> &tm) {
        T *r = rows + (size_t)(t >> 1) * per_row;
        const T *ue = r + 2 * N1, *uo = ue + hb;
        T *de = r + 2 * N1 + ((t & 1) ? 2 : 1) * (size_t)N2, *dod = de + hb;
        Carrier<T> car((t & 1) ? 0ull : CM_QUARTER_TURN, p.phases[SP_FM_STEP2X], n2);   // I: cos, Q: sin
        team_iir<T, 2, TEAMS>(p.tab + flp.off, flp,
                              [&](int q, int ph, int i) {
                                  car.at(2 * q + ph, i);
                                  return car.s * (ph ? uo : ue)[q];
                              },
                              Poly2Out<T>{de, dod}, tm);
    });
    __syncthreads();
    // discriminator: wrapped phase step of z = I - jQ between consecutive 2x samples, first step 0 (secam.py:143-148)
    const T fc = p.scalars[SS_FM_FC];
    for (int k = k_lo; k < g.count; ++k) {
        T *r = rowp(k);
        const T *ie = r + 2 * N1 + N2, *io_ = ie + hb, *qe = r + 2 * N1 + 2 * (size_t)N2, *qo = qe + hb;
        T *fe = r + 2 * N1, *fo = fe + hb;
        for (int m = threadIdx.x; m < ncc4; m += blockDim.x) {
            T f_even = (T)0, f_odd = (T)0;
            if (m < ncc) {
                const T i0 = ie[m], q0 = qe[m], i1 = io_[m], q1 = qo[m];
                // sample 2m+1 vs 2m
                f_odd = fc + (T)0.63661977236758134308 *
                                 Real<T>::atan2_(i1 * q0 - q1 * i0, i1 * i0 + q1 * q0);
                if (m > 0) {
                    const T ip = io_[m - 1], qp = qo[m - 1];
                    f_even = fc + (T)0.63661977236758134308 *
                                      Real<T>::atan2_(i0 * qp - q0 * ip, i0 * ip + q0 * qp);
                } else {
                    f_even = fc;
                }
            }
            fe[m] = f_even;
            fo[m] = f_odd;
        }
    }
    __syncthreads();
    // down2, keep the last W samples, clip, scale to the colour-difference signal -> X in the cc buffer
    for (int k = k_lo; k < g.count; ++k) {
        T *r = rowp(k);
        T *x = r + N1;
        const bool alt = is_alternate(p, g.frame, io.y0 + g.r0 + 2 * k);
        const T fsc = alt ? p.scalars[SS_FSC_DB] : p.scalars[SS_FSC_DR];
        const T inv_dev = (T)1 / (alt ? p.scalars[SS_FDEV_DB] : p.scalars[SS_FDEV_DR]);
        const T lo = p.scalars[SS_F_LO], hi = p.scalars[SS_F_HI];
        fir_down2(r + 2 * N1, r + 2 * N1 + hb, ncc4, hdn, threadIdx.x, blockDim.x, [&](int j0, const T *y) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = j0 + i - pre;
                if (j >= 0 && j < W) {
                    T f = y[i];
                    f = f < lo ? lo : (f > hi ? hi : f);
                    x[j] = (f - fsc) * inv_dev;                // cc buffer is dead after the up2
                }
            }
        });
    }
    __syncthreads();
    cta_fill_tail<T, 1>(rows + N1, per_row, nin, N1, W, N1);
    __syncthreads();
    if (p.flags & 128) {
        const FiltHdr &fd = p.filt[SF_DE_EMPH];
        for_each_iir_task<T, TEAMS>(fd, nin, scratch, [&](int t, const IirTeam<T> &tm) {
            T *x = rows + (size_t)t * per_row + N1;
            team_iir<T, 1, TEAMS>(p.tab + fd.off, fd, [&](int q, int, int) { return x[q]; },
                                  [&](int j, T v) { x[j] = v; }, tm);
        });
        __syncthreads();
    }
    // (luma, X) of every row to the pairing scratch; k_pair_rows_store combines rows y and y-2 (secam.py:297-300)
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const T *luma = rowp(k), *xc = rowp(k) + N1;
        T *dst = io.aux + ((size_t)g.fidx * io.nrows + row) * 2 * W;
        for (int q = threadIdx.x; q < (W >> 2); q += blockDim.x) {
            T y[4], a[4];
            ld4(luma + 4 * q, y);
            ld4(xc + 4 * q, a);
            st4(dst + 4 * q, y);
            st4(dst + W + 4 * q, a);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Decode, second generation (k_secam_decode2): the same chain, one row at a time per CTA of 2 (lines up to ~750 samples)
// or 4 warps (up to ~2000), every recursion a packed DF-I team (cm_iir.cuh: team_iir_pk):
//   A  luma band-stop (team 0) || chroma band-pass (team 1), each in place
//   B  anti-bell (team 1)
//   C  up2 of the chroma
//   D  I = LP(cos x) (team 0, over its own input) || Q = LP(sin x) (team 1); the mixing carrier starts at phase 0 on every
//      line (secam.py:137), so it comes from a row-independent table (DevParams::ctab) instead of a rotating carrier
//   E  discriminator (results held in registers across one barrier, then stored over I), down2, clip, scale -> X
//   F  de-emphasis (team 1)
// (luma, X) of the row go to the pairing scratch; k_pair_rows_store finishes.  smem: scratch[128] | c[N1] | cc[N1] | u2[N2] | q2[N2]
// ------------------------------------------------------------------------------------------------------------
#define SF_ROW_BS 7          // DevParams::filt slots of this kernel's use-sites (cm_api.cu: plan_secam_kernel)
#define SF_ROW_BP 8
#define SF_ROW_BELL 9
#define SF_ROW_FMLP 10
#define SF_ROW_DEEMPH 11
template <int GEO> struct SecGeo;
template <> struct SecGeo<1> { static constexpr int NW = 2, L1 = 25, L2 = 50, ITER = 12; };
template <> struct SecGeo<3> { static constexpr int NW = 4, L1 = 33, L2 = 66, ITER = 16; };

template <typename T, int GEO>
__global__ void __launch_bounds__(32 * SecGeo<GEO>::NW, sizeof(T) == 8 ? 1 : 16 / SecGeo<GEO>::NW)
k_secam_decode2(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;
    typedef SecGeo<GEO> SG;
    constexpr int NW = SG::NW, TH = NW / 2, NT = 32 * NW;
    const int W = p.W, N1 = p.n1p, hb = p.hb2, N2 = 2 * hb;
    const int pre = W / 40 - 1;                 // samples of flipped warm-up (secam.py:283)
    const int ncc = W + pre, ncc4 = (ncc + 3) & ~3, n2 = 2 * ncc;
    const int f = blockIdx.z, end = io.out_begin + io.out_count;
    const int warp = threadIdx.x >> 5, task = warp / TH, wr = warp - task * TH;
    T *c = sm, *cc = c + N1, *u2 = cc + N1, *q2 = u2 + N2;
    const FirTaps<T> hup{p.firc[SR_UP2], p.fircp[SR_UP2]}, hdn{p.firc[SR_DOWN2], p.fircp[SR_DOWN2]};
    const FiltHdr &fbs = p.filt[SF_ROW_BS], &fbp = p.filt[SF_ROW_BP], &fbell = p.filt[SF_ROW_BELL],
                  &flp = p.filt[SF_ROW_FMLP], &fde = p.filt[SF_ROW_DEEMPH];
    const long long frame = io.first_frame + f;
    const bool pref = io.in_u8 != nullptr && W <= 4 * RowPrefetch::kMaxQuads * NT;
    RowPrefetch pf;
    int row = io.out_begin + blockIdx.x;
    if (pref && row < end) pf.fetch(io, f, row, W);
    for (; row < end; row += gridDim.x) {
        if (pref) pf.stage(c, W);
        else load_comp_row(c, io, f, row, W);
        __syncthreads();
        if (pref && row + (int)gridDim.x < end) pf.fetch(io, f, row + gridDim.x, W);
        for (int i = threadIdx.x; i < iir_tail_end(fbp); i += NT) {   // warm-up prefix + composite + replicated tail
            T v;
            if (i < pre) v = c[pre - i];                     // flip(composite[1 : W/40])
            else if (i < ncc) v = c[i - pre];
            else v = c[W - 1];
            cc[i] = v;
        }
        __syncthreads();
        if (task == 0) {                                     // A: luma band-stop, in place
            warp_fill_tail<T, 1>(c, N1, W, iir_tail_end(fbs));
            team_iir_pk<T, 1, SG::L1, TH>(p.tab + fbs.off, fbs, LoadLinear<T, SG::L1>{c}, [&](int j, T v) { c[j] = v; }, wr, 2,
                                          scratch);
        } else {                                             //    chroma band-pass, in place
            team_iir_pk<T, 1, SG::L1, TH>(p.tab + fbp.off, fbp, LoadLinear<T, SG::L1>{cc}, [&](int j, T v) { cc[j] = v; }, wr,
                                          3, scratch + 32);
        }
        __syncthreads();
        if (p.flags & 64) {                                  // B: anti-bell
            if (task == 1) {
                warp_fill_tail<T, 1>(cc, N1, ncc, iir_tail_end(fbell));
                team_iir_pk<T, 1, SG::L1, TH>(p.tab + fbell.off, fbell, LoadLinear<T, SG::L1>{cc}, [&](int j, T v) { cc[j] = v; },
                                              wr, 3, scratch + 32);
            }
            __syncthreads();
        }
        for (int i = ncc + threadIdx.x; i < ncc4; i += NT) cc[i] = (T)0;       // zero pad for the resampler
        __syncthreads();
        fir_up2(u2, u2 + hb, cc, ncc4, hup, threadIdx.x, NT);                 // C
        __syncthreads();
        {                                                    // D: I (cos, team 0) || Q (sin, team 1)
            warp_fill_tail<T, 2>(u2, hb, n2, iir_tail_end(flp)); // every warp writes the same values
            T *de = task ? q2 : u2;
            const T *ct = p.ctab + (task ? 0 : (size_t)flp.npad);
            team_iir_pk<T, 2, SG::L2, TH, true>(p.tab + flp.off, flp, LoadPoly2Carrier<T, SG::L2>{u2, u2 + hb, ct, 32 * TH},
                                                Poly2Out<T>{de, de + hb}, wr, 2 + task, scratch + 32 * task);
        }
        __syncthreads();
        {   // E: wrapped phase step of z = I - jQ between consecutive 2x samples, first step 0 (secam.py:143-148)
            const T fc = p.scalars[SS_FM_FC];
            const T *ie = u2, *io_ = ie + hb, *qe = q2, *qo = qe + hb;
            T fev[SG::ITER], fod[SG::ITER];
#pragma unroll
            for (int k = 0; k < SG::ITER; ++k) {
                const int m = threadIdx.x + k * NT;
                T f_even = (T)0, f_odd = (T)0;
                if (m < ncc) {
                    const T i0 = ie[m], q0 = qe[m], i1 = io_[m], q1 = qo[m];
                    f_odd = fc + (T)0.63661977236758134308 * Real<T>::atan2_(i1 * q0 - q1 * i0, i1 * i0 + q1 * q0);
                    if (m > 0) {
                        const T ip = io_[m - 1], qp = qo[m - 1];
                        f_even = fc + (T)0.63661977236758134308 * Real<T>::atan2_(i0 * qp - q0 * ip, i0 * ip + q0 * qp);
                    } else {
                        f_even = fc;
                    }
                }
                fev[k] = f_even;
                fod[k] = f_odd;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < SG::ITER; ++k) {
                const int m = threadIdx.x + k * NT;
                if (m < ncc4) {
                    u2[m] = fev[k];
                    u2[hb + m] = fod[k];
                }
            }
        }
        __syncthreads();
        {   // down2, keep the last W samples, clip, scale to the colour-difference signal -> X in the cc buffer
            const bool alt = is_alternate(p, frame, io.y0 + row);
            const T fsc = alt ? p.scalars[SS_FSC_DB] : p.scalars[SS_FSC_DR];
            const T inv_dev = (T)1 / (alt ? p.scalars[SS_FDEV_DB] : p.scalars[SS_FDEV_DR]);
            const T lo = p.scalars[SS_F_LO], hi = p.scalars[SS_F_HI];
            fir_down2(u2, u2 + hb, ncc4, hdn, threadIdx.x, NT, [&](int j0, const T *y) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = j0 + i - pre;
                    if (j >= 0 && j < W) {
                        T fr = y[i];
                        fr = fr < lo ? lo : (fr > hi ? hi : fr);
                        cc[j] = (fr - fsc) * inv_dev;
                    }
                }
            });
        }
        __syncthreads();
        if (p.flags & 128) {                                 // F: de-emphasis
            if (task == 1) {
                warp_fill_tail<T, 1>(cc, N1, W, iir_tail_end(fde));
                team_iir_pk<T, 1, SG::L1, TH>(p.tab + fde.off, fde, LoadLinear<T, SG::L1>{cc}, [&](int j, T v) { cc[j] = v; }, wr,
                                              3, scratch + 32);
            }
            __syncthreads();
        }
        T *dst = io.aux + ((size_t)f * io.nrows + row) * 2 * W;
        for (int q = threadIdx.x; q < (W >> 2); q += NT) {
            T y[4], a[4];
            ld4(c + 4 * q, y);
            ld4(cc + 4 * q, a);
            st4(dst + 4 * q, y);
            st4(dst + W + 4 * q, a);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------
// Encode, second generation (k_secam_encode_row2): one row at a time per CTA of 2 (lines up to 768 samples) or 4 warps (up to
// 2048), the RGB words of the next row (and of its field neighbour for the ColorAveraging front end) prefetched in
// registers.  The chroma low-pass (+ LF pre-emphasis) is a short recursion run by the first team; the FM synthesis — the
// expensive part: bell gain, running phase, sin / cos per sample — is spread over ALL lanes of the CTA, LF samples per
// lane, with the integer phase prefix sum carried across the warps through shared memory.  The bell gain
//   G = m0 (1 + j kn F) / (1 + j kd F),  F = f/f0 - f0/f     becomes, with a = f/f0 and b = a^2 - 1 (F = b / a),
//   G = m0 (a + j kn b)(a - j kd b) / (a^2 + kd^2 b^2):      one reciprocal per sample instead of two divisions.
// ------------------------------------------------------------------------------------------------------------
#define SF_ENC_PRE 12        // DevParams::filt slots (cm_api.cu: plan_secam_kernel)
#define SF_ENC_EMPH 13
template <int GEO> struct SecEncGeo;
// (LF odd: lane t of the FM synthesis starts at sample t * LF — an odd stride keeps the 32 lanes on 32 banks; 12 and 16
// measured 4- and 16-way conflicts, 47 M per 16 frames of 1920x1080)
template <> struct SecEncGeo<1> { static constexpr int NW = 2, KQ = 3, L1 = 13, LF = 13; };
template <> struct SecEncGeo<3> { static constexpr int NW = 4, KQ = 4, L1 = 17, LF = 15; };


template <typename T>
__device__ __forceinline__ void secam_bell2(T f, T inv_f0, T m0, T kn, T kd, T &re, T &im) {
    const T a = f * inv_f0, b = a * a - (T)1;
    const T kb = kd * b;
    const T g = m0 * FastRcp<T>::rcp(a * a + kb * kb);
    re = g * (a * a + kn * b * kb);
    im = g * (kn - kd) * a * b;
}

template <typename T, int GEO>
__global__ void __launch_bounds__(32 * SecEncGeo<GEO>::NW, sizeof(T) == 8 ? 1 : 16 / SecEncGeo<GEO>::NW)
k_secam_encode_row2(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;
    typedef SecEncGeo<GEO> EG;
    constexpr int kQ = EG::KQ, NW = EG::NW, NT = 32 * NW, LF = EG::LF;
    unsigned long long *wsum = reinterpret_cast<unsigned long long *>(scratch + 64);        // [NW] warp totals of the phase scan
    const int W = p.W, N1 = p.n1p, W4 = W >> 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool avg = (p.flags & 2) != 0;
    const int field = blockIdx.y, f = blockIdx.z;
    const long long frame = io.first_frame + f;
    const int first = io.out_begin + field, nout = (io.out_count - field + 1) >> 1;
    const FiltHdr &fpre = p.filt[SF_ENC_PRE], &femph = p.filt[SF_ENC_EMPH];
    T *ys = sm, *cs = ys + N1;
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
    const T f0 = p.scalars[SS_BELL_F0], inv_f0 = (T)1 / f0;
    const T m0 = p.scalars[SS_M0], kn = p.scalars[SS_KN], kd = p.scalars[SS_KD];
    int k = blockIdx.x;
    if (k < nout) fetch(first + 2 * k);
    for (; k < nout; k += gridDim.x) {
        const int row = first + 2 * k, line = io.y0 + row;
        const bool alt = is_alternate(p, frame, line);
        const int ci = alt ? 6 : 3;                         // D'B (row 2 of the matrix) on alternate lines, else D'R
#pragma unroll
        for (int j = 0; j < kQ; ++j) {
            const int q = threadIdx.x + j * NT;
            if (q < W4) {
                T r[4], gg[4], b[4], y[4], c[4];
                unpack(wc[j], r, gg, b);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    y[i] = p.enc[0] * r[i] + p.enc[1] * gg[i] + p.enc[2] * b[i];
                    c[i] = p.enc[ci] * r[i] + p.enc[ci + 1] * gg[i] + p.enc[ci + 2] * b[i];
                }
                if (avg) {
                    unpack(wn[j], r, gg, b);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        c[i] = (T)0.5 * ((p.enc[ci] * r[i] + p.enc[ci + 1] * gg[i] + p.enc[ci + 2] * b[i]) + c[i]);
                }
                st4(ys + 4 * q, y);
                st4(cs + 4 * q, c);
            }
        }
        __syncthreads();
        if (k + (int)gridDim.x < nout) fetch(first + 2 * (k + gridDim.x));       // in flight during the rest of the row
        // chroma low-pass (+ LF pre-emphasis): one short task, run by all the warps as one team
        warp_fill_tail<T, 1>(cs, N1, W, iir_tail_end(fpre));
        team_iir_pk<T, 1, EG::L1, NW>(p.tab + fpre.off, fpre, LoadLinear<T, EG::L1>{cs}, [&](int j, T v) { cs[j] = v; }, warp, 2,
                                      scratch);
        __syncthreads();
        if (p.flags & 128) {
            warp_fill_tail<T, 1>(cs, N1, W, iir_tail_end(femph));
            team_iir_pk<T, 1, EG::L1, NW>(p.tab + femph.off, femph, LoadLinear<T, EG::L1>{cs}, [&](int j, T v) { cs[j] = v; }, warp,
                                          2, scratch);
            __syncthreads();
        }
        {   // FM synthesis (secam.py:240-246): phase[j] = start - angle(G[0]) + sum_{i=1..j} pi f[i], 0.64 fixed-point turns
            const T fsc = alt ? p.scalars[SS_FSC_DB] : p.scalars[SS_FSC_DR];
            const T fdev = alt ? p.scalars[SS_FDEV_DB] : p.scalars[SS_FDEV_DR];
            const T dlo = p.scalars[SS_F_LO] - fsc, dhi = p.scalars[SS_F_HI] - fsc;
            const unsigned long long centre = alt ? p.phases[SP_FSC_DB_HALF] : p.phases[SP_FSC_DR_HALF];
            unsigned long long run = secam_phase_inverted(p, frame, line) ? 0x8000000000000000ull : 0ull;
            {
                T d0 = fdev * cs[0];
                d0 = d0 < dlo ? dlo : (d0 > dhi ? dhi : d0);
                T re, im;
                secam_bell2(fsc + d0, inv_f0, m0, kn, kd, re, im);
                run -= (unsigned long long)Fix64<T>::half_turns(Real<T>::atan2_(im, re) * (T)0.31830988618379067154);
            }
            const int base = threadIdx.x * LF;
            T d[LF];
            unsigned long long acc[LF];
            unsigned long long sum = 0;
#pragma unroll
            for (int i = 0; i < LF; ++i) {
                const int j = base + i;
                T dev = fdev * cs[j < W ? j : W - 1];
                dev = dev < dlo ? dlo : (dev > dhi ? dhi : dev);
                d[i] = dev;
                if (j > 0 && j < W) sum += centre + (unsigned long long)Fix64<T>::half_turns(dev);
                acc[i] = sum;
            }
            unsigned long long incl = sum;                                    // inclusive scan of the lane totals
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, 1 << s);
                if (lane >= (1 << s)) incl += v;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            unsigned long long before = run + incl - sum;
            for (int w = 0; w < warp; ++w) before += wsum[w];
#pragma unroll
            for (int i = 0; i < LF; ++i) {
                const int j = base + i;
                if (j < W) {
                    T re, im, s, co;
                    secam_bell2(fsc + d[i], inv_f0, m0, kn, kd, re, im);
                    Real<T>::sincos_turns_fast(before + acc[i], s, co);
                    ys[j] += re * co - im * s;
                }
            }
        }
        __syncthreads();
        for (int q = threadIdx.x; q < W4; q += NT) {
            T o[4];
            ld4(ys + 4 * q, o);
            store_comp4(io, ((size_t)f * io.nrows + row) * p.Wc + 4 * q, o);
        }
        __syncthreads();
    }
}
