// Launchers of the D2-MAC family: kernels in cm_mac.cuh.
#include "cm_host.h"
#include "cm_mac.cuh"

// elements of the polyphase tables (DevParams::ptab) that the kernels stage in shared memory
template <typename T>
static int mac_taps_len(const DevParams<T> &p) {
    int total = 0;
    for (int r = 0; r < CM_NRES; ++r)
        if (p.poly[r].up) total = p.poly[r].off + 4 * p.poly[r].up * p.poly[r].stride;
    return (total + 3) & ~3;
}

template <typename T>
static const MacConst<T> &mac_const(const cm_modem *m) {
    if constexpr (IsF32<T>::value) return m->mcf;
    else return m->mcd;
}

template <typename T>
int mac_encode(cm_modem *m, IoArgs<T> io, cudaStream_t st) {
    const DevParams<T> &p = params_of<T>(m);
    if (io.out_count <= 0) return CM_OK;
    if (p.Wc > 1080) return cm_fail(CM_ERR_UNSUPPORTED, "MAC widths above 1080 are not built%s");
    const MacConst<T> &mc = mac_const<T>(m);
    const int tl = (mc.ok_luma && mc.ok_chroma && mc.ok_out) ? 0 : mac_taps_len(p);     // nothing to stage: every ratio runs from the constant bank
    auto seg = [&](int n) { const int e = p.mac_fp + ((n + 3) & ~3) + p.mac_bp; return (size_t)(e + (((e >> 5) << 2) & p.mac_skew) + 4); };
    auto bytes = [&](int r) { return ((size_t)tl + (size_t)r * mac_encode_row_elems((int)seg(p.W), (int)seg(1080))) * sizeof(T); };
    // measured (threads x rows per CTA, 720 and 1920 wide): four warps walking one row; eight CTAs per SM in different phases
    int R = pick_rows(m, 1, (size_t)m->smem_optin / 2, bytes);
    if (!R) return cm_fail(CM_ERR_UNSUPPORTED, "line too wide for the MAC encode kernel%s");
    set_groups(io, R);
    void (*kern)(const DevParams<T>, const IoArgs<T>, int, const MacConst<T>) = p.mac_skew ? k_mac_encode<T, -1> : k_mac_encode<T, 0>;
    int rc = set_smem(kern, bytes(R));
    if (rc) return rc;
    dim3 grid = cm_grid(io);
    {
        LaunchTimer lt(m, CM_K_ENCODE, st);
        kern<<<grid, m->tune.mac_threads, bytes(R), st>>>(p, io, tl, mac_const<T>(m));
    }
    cm_count_launch();
    CUDA_TRY(cudaGetLastError());
    return CM_OK;
}

template <typename T>
int mac_decode(cm_modem *m, IoArgs<T> io, cudaStream_t st) {
    const DevParams<T> &p = params_of<T>(m);
    if (io.out_count <= 0) return CM_OK;
    const int tl = mac_const<T>(m).ok_comp ? 0 : mac_taps_len(p);
    auto bytes = [&](int r) {
        const int e = p.mac_fp + ((p.Wc + 3) & ~3) + p.mac_bp;
        return ((size_t)tl + (size_t)(r + 1) * mac_decode_row_elems(e + (((e >> 5) << 2) & p.mac_skew) + 4)) * sizeof(T);
    };
    int R = pick_rows(m, 2, (size_t)m->smem_optin / 2, bytes);           // measured: two rows + the chroma of the row ahead, four warps
    if (!R) return cm_fail(CM_ERR_UNSUPPORTED, "line too wide for the MAC decode kernel%s");
    set_groups(io, R);
    void (*kern)(const DevParams<T>, const IoArgs<T>, int, const MacConst<T>) = p.mac_skew ? k_mac_decode<T, -1> : k_mac_decode<T, 0>;
    int rc = set_smem(kern, bytes(R));
    if (rc) return rc;
    dim3 grid = cm_grid(io);
    {
        LaunchTimer lt(m, CM_K_DECODE_OTHER, st);
        kern<<<grid, m->tune.mac_threads, bytes(R), st>>>(p, io, tl, mac_const<T>(m));
    }
    cm_count_launch();
    CUDA_TRY(cudaGetLastError());
    return CM_OK;
}

CM_INSTANTIATE(template int mac_encode<float>(cm_modem *, IoArgs<float>, cudaStream_t);,
               template int mac_encode<double>(cm_modem *, IoArgs<double>, cudaStream_t);)
CM_INSTANTIATE(template int mac_decode<float>(cm_modem *, IoArgs<float>, cudaStream_t);,
               template int mac_decode<double>(cm_modem *, IoArgs<double>, cudaStream_t);)
