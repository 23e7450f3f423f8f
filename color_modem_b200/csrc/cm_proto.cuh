// 819-line AM proto-SECAM kernels.  Reference: color_modem/color/protosecam.py.
//
// Encode (protosecam.py:74-90): line-sequential D'R / D'B -> low-pass -> 0.125 (1 + c) cos(phi) on a luma that
// has had the chroma band removed at 3x (up3 -> band-stop -> down3, optional).
// Decode (protosecam.py:92-112) at 3x: band-pass -> full-wave rectifier (pi/2 |x|) -> low-pass -> down3 ->
// 8 c - 1; luma = down3(band-stop(up3)); rows pair their colour-difference signal with the previous row's.
#pragma once
#include "cm_niir.cuh"

// Encode.  2 warps per row.  smem: scratch[128] (IIR team scratch) + R * (luma[N1] | chroma[N1] | up[N3])
template <typename T>
__global__ void __launch_bounds__(CM_NTHREADS)
k_proto_encode(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, N1 = p.n1p, hb = p.hb3, N3 = 3 * hb, n3 = 3 * W, W4 = W >> 2;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool avg = (p.flags & 2) != 0, luma_filter = (p.flags & 256) != 0;
    T *rows = sm + 128;
    const size_t per_row = 2 * (size_t)N1 + N3;
    const FirTaps<T> hup{p.firc[PR_UP3], p.fircp[PR_UP3]}, hdn{p.firc[PR_DOWN3], p.fircp[PR_DOWN3]};   // constant bank (kernel parameter)
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const int nrow = (row + 2 < io.nrows) ? row + 2 : row;
        const int ci = is_alternate(p, g.frame, io.y0 + row) ? 6 : 3;      // D'B on alternate lines, else D'R
        T *ys = rows + k * per_row, *cs = ys + N1;
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
    for (int k = 0; k < g.count; ++k) {
        cta_fill_tail<T, 1>(rows + k * per_row + N1, (size_t)N1, 1, N1, W, N1);
        if (luma_filter) {
            T *u = rows + k * per_row + 2 * N1;
            fir_up3(u, u + hb, u + 2 * hb, rows + k * per_row, W, hup, threadIdx.x, blockDim.x);
        }
    }
    __syncthreads();
    if (luma_filter) cta_fill_tail<T, 3>(rows + 2 * N1, per_row, g.count, hb, n3, N3);
    __syncthreads();
    for (int t = warp; t < 2 * g.count; t += nwarps) {
        T *r = rows + (t >> 1) * per_row;
        if ((t & 1) == 0) {
            T *cs = r + N1;
            const FiltHdr &f = p.filt[PF_PRE_LP];
            warp_iir<T, 1>(p.tab + f.off, f, [&](int q, int, int) { return cs[q]; }, [&](int j, T v) { cs[j] = v; });
        } else if (luma_filter) {
            T *u = r + 2 * N1;
            const FiltHdr &f = p.filt[PF_BS_UP];
            warp_iir<T, 3>(p.tab + f.off, f, [&](int q, int ph, int) { return u[ph * hb + q]; },
                           Poly3Out<T>{u, hb});
        }
    }
    __syncthreads();
    const Down3Taps<T> tp(hdn);
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k, line = io.y0 + row;
        const T *ys = rows + k * per_row, *cs = ys + N1, *u = ys + 2 * N1;
        const unsigned long long ph0 = start_phase(p, g.frame, line);
        for (int q = threadIdx.x; q < W4; q += blockDim.x) {
            T o[4], luma[4];
            if (luma_filter) down3_quad(tp, u, u + hb, u + 2 * hb, W, 4 * q, luma);
            else ld4(ys + 4 * q, luma);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int x = 4 * q + i;
                T s, c;
                Real<T>::sincos_turns(ph0 + (unsigned long long)x * p.phases[PP_STEP1X], s, c);
                o[i] = luma[i] + c * ((T)0.125 * ((T)1 + cs[x]));
            }
            store_comp4(io, ((size_t)g.fidx * io.nrows + row) * p.Wc + 4 * q, o);
        }
    }
}

// Decode.  smem: scratch[128] (IIR team scratch) + (R+1) rows x ( c -> X [N1] | up[N3] | chroma[N3] | luma[N3] )
template <typename T>
__global__ void __launch_bounds__(CM_NTHREADS, 2)
k_proto_decode(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    RowGroup g;
    if (!decode_group(io, g)) return;
    const int W = p.W, N1 = p.n1p, hb = p.hb3, N3 = 3 * hb, n3 = 3 * W;
    T *taps = sm;
    T *rows = sm + 128;
    const size_t per_row = (size_t)N1 + 3 * (size_t)N3;
    const int k_lo = 0;                          // rows are independent here: pairing with the previous row happens
    const int nin = g.count;                     // in k_pair_rows_store (cm_io.cuh)
    const FirTaps<T> hup{p.firc[PR_UP3], p.fircp[PR_UP3]}, hdn{p.firc[PR_DOWN3], p.fircp[PR_DOWN3]};   // constant bank (kernel parameter)
    auto rowp = [&](int k) { return rows + (size_t)(k - k_lo) * per_row; };
    load_comp_rows(io, g.fidx, nin, W, [&](int k) { return rowp(k_lo + k); }, [&](int k) { return g.r0 + 2 * (k_lo + k); });
    __syncthreads();
    for (int k = k_lo; k < g.count; ++k) {
        T *u = rowp(k) + N1;
        fir_up3(u, u + hb, u + 2 * hb, rowp(k), W, hup, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    cta_fill_tail<T, 3>(rows + N1, per_row, nin, hb, n3, N3);
    __syncthreads();
    const FiltHdr &fbp = p.filt[PF_BP_UP], &fbs = p.filt[PF_BS_UP], &fpost = p.filt[PF_POST_LP];
    // chroma band-pass (u -> b) and luma band-stop (u -> l) side by side, each task run by a team of warps
    for_each_iir_task<T, true>(fbp, nin, taps, [&](int t, const IirTeam<T> &tm) {
        T *r = rows + (size_t)t * per_row;
        const T *u = r + N1;
        T *b = r + N1 + N3;
        team_iir<T, 3, true>(p.tab + fbp.off, fbp, [&](int q, int ph, int) { return u[ph * hb + q]; },
                             Poly3Out<T>{b, hb}, tm);
    });
    for_each_iir_task<T, true>(fbs, nin, taps, [&](int t, const IirTeam<T> &tm) {
        T *r = rows + (size_t)t * per_row;
        const T *u = r + N1;
        T *l = r + N1 + 2 * (size_t)N3;
        team_iir<T, 3, true>(p.tab + fbs.off, fbs, [&](int q, int ph, int) { return u[ph * hb + q]; },
                             Poly3Out<T>{l, hb}, tm);
    }, nin);
    __syncthreads();
    for_each_iir_task<T, true>(fpost, nin, taps, [&](int t, const IirTeam<T> &tm) {     // rectifier + low-pass, in place
        T *b = rows + (size_t)t * per_row + N1 + N3;
        warp_fill_tail<T, 3>(b, hb, n3, N3);           // every warp of the team writes the same values
        team_iir<T, 3, true>(p.tab + fpost.off, fpost,
                             [&](int q, int ph, int) { return (T)1.57079632679489661923 * Real<T>::abs_(b[ph * hb + q]); },
                             Poly3Out<T>{b, hb}, tm);
    });
    __syncthreads();
    const Down3Taps<T> tp(hdn);
    for (int k = k_lo; k < g.count; ++k) {              // X = 8 down3(chroma_up) - 1 into the (dead) composite buffer
        T *x = rowp(k);
        const T *b = rowp(k) + N1 + N3;
        for (int q = threadIdx.x; q < (W >> 2); q += blockDim.x) {
            T y[4];
            down3_quad(tp, b, b + hb, b + 2 * hb, W, 4 * q, y);
#pragma unroll
            for (int i = 0; i < 4; ++i) y[i] = (T)8 * y[i] - (T)1;
            st4(x + 4 * q, y);
        }
    }
    __syncthreads();
    // (luma, X) of every row to the pairing scratch; k_pair_rows_store combines rows y and y-2 (protosecam.py:105-108)
    for (int k = 0; k < g.count; ++k) {
        const int row = g.r0 + 2 * k;
        const T *xc = rowp(k), *l = rowp(k) + N1 + 2 * (size_t)N3;
        T *dst = io.aux + ((size_t)g.fidx * io.nrows + row) * 2 * W;
        for (int q = threadIdx.x; q < (W >> 2); q += blockDim.x) {
            T y[4], a[4];
            down3_quad(tp, l, l + hb, l + 2 * hb, W, 4 * q, y);
            ld4(xc + 4 * q, a);
            st4(dst + 4 * q, y);
            st4(dst + W + 4 * q, a);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Second generation: one row at a time per CTA of 4 (lines up to ~800 samples) or 8 warps (up to ~2100), two teams of
// NW / 2 warps, packed DF-I recursions at the 3x rate (team_iir_pk), the next composite / RGB row prefetched.
//   decode  up3 -> chroma band-pass (team 0, u -> b) || luma band-stop (team 1, in place) -> rectifier + low-pass (team 0, in
//           place) -> down3 of both -> (luma, X) to the pairing scratch (k_pair_rows_store finishes).  smem: c[N1] | u[N3] | b[N3]
//   encode  luma band-stop at 3x (team 0) || chroma low-pass at 1x (team 1) -> down3, carrier, level map
// ------------------------------------------------------------------------------------------------------------
#define PF_ROW_BP 4          // DevParams::filt slots of these kernels' use-sites (cm_api.cu: plan_proto_kernel)
#define PF_ROW_BS 5
#define PF_ROW_POST 6
#define PF_ENC_PRE 7
template <int GEO> struct ProtoGeo;
template <> struct ProtoGeo<1> { static constexpr int NW = 4, L3 = 39, L1 = 13, KQ = 2; };
template <> struct ProtoGeo<3> { static constexpr int NW = 8, L3 = 51, L1 = 17, KQ = 2; };

template <typename T, int GEO>
__global__ void __launch_bounds__(32 * ProtoGeo<GEO>::NW, sizeof(T) == 8 ? 1 : 16 / ProtoGeo<GEO>::NW)
k_proto_decode2(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;
    typedef ProtoGeo<GEO> PG;
    constexpr int NW = PG::NW, TH = NW / 2, NT = 32 * NW, L3 = PG::L3;
    const int W = p.W, N1 = p.n1p, hb = p.hb3, N3 = 3 * hb, n3 = 3 * W;
    const int f = blockIdx.z, end = io.out_begin + io.out_count;
    const int warp = threadIdx.x >> 5, task = warp / TH, wr = warp - task * TH;
    T *c = sm, *u = c + N1, *b = u + N3;
    const FirTaps<T> hup{p.firc[PR_UP3], p.fircp[PR_UP3]}, hdn{p.firc[PR_DOWN3], p.fircp[PR_DOWN3]};
    const Down3Taps<T> tp(hdn);
    const FiltHdr &fbp = p.filt[PF_ROW_BP], &fbs = p.filt[PF_ROW_BS], &fpost = p.filt[PF_ROW_POST];
    const bool pref = io.in_u8 != nullptr && W <= 4 * RowPrefetch::kMaxQuads * NT;
    RowPrefetch pf;
    int row = io.out_begin + blockIdx.x;
    if (pref && row < end) pf.fetch(io, f, row, W);
    for (; row < end; row += gridDim.x) {
        if (pref) pf.stage(c, W);
        else load_comp_row(c, io, f, row, W);
        __syncthreads();
        if (pref && row + (int)gridDim.x < end) pf.fetch(io, f, row + gridDim.x, W);
        fir_up3(u, u + hb, u + 2 * hb, c, W, hup, threadIdx.x, NT);
        __syncthreads();
        warp_fill_tail<T, 3>(u, hb, n3, max(iir_tail_end(fbp), iir_tail_end(fbs)));       // every warp writes the same values
        if (task == 0)
            team_iir_pk<T, 3, L3, TH, true>(p.tab + fbp.off, fbp, LoadPoly3<T, L3, false>{u, hb}, Poly3Out<T>{b, hb}, wr, 2, scratch);
        else
            team_iir_pk<T, 3, L3, TH, true>(p.tab + fbs.off, fbs, LoadPoly3<T, L3, false>{u, hb}, Poly3Out<T>{u, hb}, wr, 3,
                                            scratch + 32);
        __syncthreads();
        if (task == 0) {                                         // rectifier + low-pass, in place (the team's own barrier
            warp_fill_tail<T, 3>(b, hb, n3, iir_tail_end(fpost));  // separates its loads from its stores)
            team_iir_pk<T, 3, L3, TH>(p.tab + fpost.off, fpost, LoadPoly3<T, L3, true>{b, hb}, Poly3Out<T>{b, hb}, wr, 2, scratch);
        }
        __syncthreads();
        T *dst = io.aux + ((size_t)f * io.nrows + row) * 2 * W;
        for (int q = threadIdx.x; q < (W >> 2); q += NT) {       // luma = down3(band-stopped), X = 8 down3(envelope) - 1
            T y[4], x[4];
            down3_quad(tp, u, u + hb, u + 2 * hb, W, 4 * q, y);
            down3_quad(tp, b, b + hb, b + 2 * hb, W, 4 * q, x);
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = (T)8 * x[i] - (T)1;
            st4(dst + 4 * q, y);
            st4(dst + W + 4 * q, x);
        }
        __syncthreads();
    }
}

template <typename T, int GEO>
__global__ void __launch_bounds__(32 * ProtoGeo<GEO>::NW, sizeof(T) == 8 ? 1 : 16 / ProtoGeo<GEO>::NW)
k_proto_encode_row2(const __grid_constant__ DevParams<T> p, const __grid_constant__ IoArgs<T> io) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *scratch = reinterpret_cast<T *>(smem_raw), *sm = scratch + 128;
    typedef ProtoGeo<GEO> PG;
    constexpr int NW = PG::NW, TH = NW / 2, NT = 32 * NW, L3 = PG::L3, L1 = PG::L1, kQ = PG::KQ;
    const int W = p.W, N1 = p.n1p, hb = p.hb3, N3 = 3 * hb, n3 = 3 * W, W4 = W >> 2;
    const int warp = threadIdx.x >> 5, task = warp / TH, wr = warp - task * TH;
    const bool avg = (p.flags & 2) != 0, luma_filter = (p.flags & 256) != 0;
    const int field = blockIdx.y, f = blockIdx.z;
    const long long frame = io.first_frame + f;
    const int first = io.out_begin + field, nout = (io.out_count - field + 1) >> 1;
    const FiltHdr &fpre = p.filt[PF_ENC_PRE], &fbs = p.filt[PF_ROW_BS];
    const FirTaps<T> hup{p.firc[PR_UP3], p.fircp[PR_UP3]}, hdn{p.firc[PR_DOWN3], p.fircp[PR_DOWN3]};
    const Down3Taps<T> tp(hdn);
    T *ys = sm, *cs = ys + N1, *u = cs + N1;
    (void)N3;
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
    Real<T>::sincos_turns(p.phases[PP_STEP1X], rs, rc);
    int k = blockIdx.x;
    if (k < nout) fetch(first + 2 * k);
    for (; k < nout; k += gridDim.x) {
        const int row = first + 2 * k, line = io.y0 + row;
        const int ci = is_alternate(p, frame, line) ? 6 : 3;      // D'B on alternate lines, else D'R
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
        if (k + (int)gridDim.x < nout) fetch(first + 2 * (k + gridDim.x));
        if (luma_filter) {
            fir_up3(u, u + hb, u + 2 * hb, ys, W, hup, threadIdx.x, NT);
            __syncthreads();
        }
        if (task == 0) {
            if (luma_filter) {                                   // luma band-stop at 3x, in place
                warp_fill_tail<T, 3>(u, hb, n3, iir_tail_end(fbs));
                team_iir_pk<T, 3, L3, TH>(p.tab + fbs.off, fbs, LoadPoly3<T, L3, false>{u, hb}, Poly3Out<T>{u, hb}, wr, 2, scratch);
            }
        } else {                                                 // chroma low-pass at 1x, in place
            warp_fill_tail<T, 1>(cs, N1, W, iir_tail_end(fpre));
            team_iir_pk<T, 1, L1, TH>(p.tab + fpre.off, fpre, LoadLinear<T, L1>{cs}, [&](int j, T v) { cs[j] = v; }, wr, 3,
                                      scratch + 32);
        }
        __syncthreads();
        const unsigned long long ph0 = start_phase(p, frame, line);
        for (int q = threadIdx.x; q < W4; q += NT) {
            T o[4], luma[4], ch[4], s[4], c[4];
            if (luma_filter) down3_quad(tp, u, u + hb, u + 2 * hb, W, 4 * q, luma);
            else ld4(ys + 4 * q, luma);
            ld4(cs + 4 * q, ch);
            carrier4_fast(ph0 + (unsigned long long)(4 * q) * p.phases[PP_STEP1X], rs, rc, s, c);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = luma[i] + c[i] * ((T)0.125 * ((T)1 + ch[i]));
            store_comp4(io, ((size_t)f * io.nrows + row) * p.Wc + 4 * q, o);
        }
        __syncthreads();
    }
}
