// Launchers of the SECAM family: kernels in cm_secam.cuh.
#include "cm_host.h"
#include "cm_secam.cuh"

template <typename T>
int secam_encode(cm_modem *m, IoArgs<T> io, cudaStream_t st) {
    const DevParams<T> &p = params_of<T>(m);
    if (io.out_count <= 0) return CM_OK;
    auto bytes = [&](int r) { return (size_t)r * 2 * p.n1p * sizeof(T); };
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
    // pass 1 (heavy): (luma, X) per row for the output rows and the two rows above them
    const size_t aux_bytes = (size_t)io.nframes * io.nrows * 2 * p.Wo * sizeof(T);
    io.aux = (T *)cm_ensure_aux(m, aux_bytes);
    if (!io.aux) return CM_ERR_NOMEM;
    int rc;
    IoArgs<T> a = io;
    a.out_begin = io.out_begin >= 2 ? io.out_begin - 2 : 0;
    a.out_count = io.out_begin + io.out_count - a.out_begin;
    auto bytes = [&](int r) { return (CM_TAPS_ELEMS + (size_t)r * (2 * (size_t)p.n1p + 6 * (size_t)p.hb2)) * sizeof(T); };
    int R = pick_rows(m, 2, (size_t)m->smem_optin / 2, bytes);
    if (!R) R = pick_rows(m, 1, (size_t)m->smem_optin, bytes);
    if (!R) return cm_fail(CM_ERR_UNSUPPORTED, "line too wide for the SECAM decode kernel%s");
    set_groups(a, R);
    bool teams = false;
    for (int i = 0; i < CM_NFILT; ++i) teams = teams || (p.filt[i].nsec && p.filt[i].nsuper > 1);
    rc = teams ? set_smem(k_secam_decode<T, true>, bytes(R)) : set_smem(k_secam_decode<T, false>, bytes(R));
    if (rc) return rc;
    {
        LaunchTimer lt(m, CM_K_DECODE_OTHER, st);
        if (teams) k_secam_decode<T, true><<<cm_grid(a), CM_NTHREADS, bytes(R), st>>>(p, a);
        else k_secam_decode<T, false><<<cm_grid(a), cta_threads(m, 2 * R), bytes(R), st>>>(p, a);
    }
    cm_count_launch();
    CUDA_TRY(cudaGetLastError());
    // pass 2 (light): pair rows y / y-2, inverse matrix, store
    {
        LaunchTimer lt(m, CM_K_DECODE_OTHER, st);
        dim3 grid(1u, (unsigned)io.out_count, (unsigned)io.nframes);
        k_pair_rows_store<T><<<grid, 192, 0, st>>>(p, io);
    }
    cm_count_launch();
    CUDA_TRY(cudaGetLastError());
    return CM_OK;
}

CM_INSTANTIATE(template int secam_encode<float>(cm_modem *, IoArgs<float>, cudaStream_t);,
               template int secam_encode<double>(cm_modem *, IoArgs<double>, cudaStream_t);)
CM_INSTANTIATE(template int secam_decode<float>(cm_modem *, IoArgs<float>, cudaStream_t);,
               template int secam_decode<double>(cm_modem *, IoArgs<double>, cudaStream_t);)
