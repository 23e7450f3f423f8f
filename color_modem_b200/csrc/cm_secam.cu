// Launchers of the SECAM family: kernels in cm_secam.cuh.
#include "cm_host.h"
#include "cm_secam.cuh"

template <typename T>
int secam_encode(cm_modem *m, IoArgs<T> io, cudaStream_t st) {
    const DevParams<T> &p = params_of<T>(m);
    if (io.out_count <= 0) return CM_OK;
    auto bytes = [&](int r) { return (size_t)r * 2 * p.n1p * sizeof(T); };
    if (io.in_u8 && p.row_geo && p.filt[SF_ENC_PRE].nsec && !m->tune.rows_v1 && !m->tune.onepass) {
        void (*kern)(const DevParams<T>, const IoArgs<T>) = p.row_geo == 1 ? k_secam_encode_row2<T, 1> : k_secam_encode_row2<T, 3>;
        const size_t b2 = (128 + 2 * (size_t)p.n1p) * sizeof(T);
        int rc2 = set_smem(kern, b2);
        if (rc2) return rc2;
        const int rpc = cm_rows_per_cta(m, (long long)io.out_count * io.nframes), nf = (io.out_count + 1) >> 1;
        {
            LaunchTimer lt(m, CM_K_ENCODE, st);
            kern<<<dim3((unsigned)((nf + rpc - 1) / rpc), 2u, (unsigned)io.nframes), p.row_geo == 1 ? 64 : 128, b2, st>>>(p, io);
        }
        cm_count_launch();
        CUDA_TRY(cudaGetLastError());
        return CM_OK;
    }
    int R = pick_rows(m, 2, (size_t)m->smem_optin / 2, bytes);
    if (!R) return cm_fail(CM_ERR_UNSUPPORTED, "line too wide for the SECAM encode kernel%s");
    set_groups(io, R);
    int rc = set_smem(k_secam_encode<T>, bytes(R));
    if (rc) return rc;
    dim3 grid = cm_grid(io);
    {
        LaunchTimer lt(m, CM_K_ENCODE, st);
        k_secam_encode<T><<<grid, cta_threads(m, R), bytes(R), st>>>(p, io);
    }
    cm_count_launch();
    CUDA_TRY(cudaGetLastError());
    return CM_OK;
}

template <typename T>
int secam_decode(cm_modem *m, IoArgs<T> io, cudaStream_t st) {
    const DevParams<T> &p = params_of<T>(m);
    if (io.out_count <= 0) return CM_OK;
    // pass 1 (heavy): (luma, X) per row for the output rows and the two rows above them, into a pairing scratch of
    // 2 * Wo elements per row; the batch is cut so that the scratch stays within 2 GiB
    const size_t frame_elems = (size_t)io.nrows * 2 * p.Wo;
    int chunk = (int)(((size_t)2 << 30) / (frame_elems * sizeof(T)));
    if (chunk < 1) chunk = 1;
    if (m->tune.chunk > 0) chunk = m->tune.chunk;
    if (chunk > io.nframes) chunk = io.nframes;
    T *aux = (T *)cm_ensure_aux(m, (size_t)chunk * frame_elems * sizeof(T));
    if (!aux) return CM_ERR_NOMEM;
    const bool rows2 = p.row_geo && !m->tune.rows_v1 && !m->tune.onepass;
    auto bytes = [&](int r) { return (CM_TAPS_ELEMS + (size_t)r * (2 * (size_t)p.n1p + 6 * (size_t)p.hb2)) * sizeof(T); };
    const size_t b2 = (128 + 2 * (size_t)p.n1p + 4 * (size_t)p.hb2) * sizeof(T);
    void (*row_kernel)(const DevParams<T>, const IoArgs<T>) = p.row_geo == 1 ? k_secam_decode2<T, 1> : k_secam_decode2<T, 3>;
    int R = 0, rc;
    bool teams = false;
    if (rows2) {
        if (b2 > (size_t)m->smem_optin) return cm_fail(CM_ERR_UNSUPPORTED, "line too wide for the SECAM decode kernel%s");
        rc = set_smem(row_kernel, b2);
        if (rc) return rc;
    } else {
        R = pick_rows(m, 2, (size_t)m->smem_optin / 2, bytes);
        if (!R) R = pick_rows(m, 1, (size_t)m->smem_optin, bytes);
        if (!R) return cm_fail(CM_ERR_UNSUPPORTED, "line too wide for the SECAM decode kernel%s");
        for (int i = 0; i < 7; ++i) teams = teams || (p.filt[i].nsec && p.filt[i].nsuper > 1);
        rc = teams ? set_smem(k_secam_decode<T, true>, bytes(R)) : set_smem(k_secam_decode<T, false>, bytes(R));
        if (rc) return rc;
    }
    const size_t in_frame = (size_t)io.nrows * p.Wc, out_frame = (size_t)io.nrows * p.Wo * 3;
    for (int f0 = 0; f0 < io.nframes; f0 += chunk) {
        IoArgs<T> c = io;
        c.nframes = io.nframes - f0 < chunk ? io.nframes - f0 : chunk;
        c.first_frame = io.first_frame + f0;
        c.aux = aux;
        if (c.in_u8) c.in_u8 += (size_t)f0 * in_frame;
        if (c.in_f) c.in_f += (size_t)f0 * in_frame;
        if (c.out_u8) c.out_u8 += (size_t)f0 * out_frame;
        if (c.out_f) c.out_f += (size_t)f0 * out_frame;
        IoArgs<T> a = c;
        a.out_begin = c.out_begin >= 2 ? c.out_begin - 2 : 0;
        a.out_count = c.out_begin + c.out_count - a.out_begin;
        {
            LaunchTimer lt(m, CM_K_DECODE_OTHER, st);
            if (rows2) {
                const int rpc = cm_rows_per_cta(m, (long long)a.out_count * c.nframes);
                row_kernel<<<dim3((unsigned)((a.out_count + rpc - 1) / rpc), 1u, (unsigned)c.nframes), p.row_geo == 1 ? 64 : 128,
                             b2, st>>>(p, a);
            } else {
                set_groups(a, R);
                if (teams) k_secam_decode<T, true><<<cm_grid(a), CM_NTHREADS, bytes(R), st>>>(p, a);
                else k_secam_decode<T, false><<<cm_grid(a), cta_threads(m, 2 * R), bytes(R), st>>>(p, a);
            }
        }
        cm_count_launch();
        CUDA_TRY(cudaGetLastError());
        // pass 2 (light): pair rows y / y-2, inverse matrix, store
        {
            LaunchTimer lt(m, CM_K_DECODE_OTHER, st);
            dim3 grid(1u, (unsigned)c.out_count, (unsigned)c.nframes);
            k_pair_rows_store<T><<<grid, 192, 0, st>>>(p, c);
        }
        cm_count_launch();
        CUDA_TRY(cudaGetLastError());
    }
    return CM_OK;
}

CM_INSTANTIATE(template int secam_encode<float>(cm_modem *, IoArgs<float>, cudaStream_t);,
               template int secam_encode<double>(cm_modem *, IoArgs<double>, cudaStream_t);)
CM_INSTANTIATE(template int secam_decode<float>(cm_modem *, IoArgs<float>, cudaStream_t);,
               template int secam_decode<double>(cm_modem *, IoArgs<double>, cudaStream_t);)
