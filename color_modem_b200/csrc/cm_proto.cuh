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
