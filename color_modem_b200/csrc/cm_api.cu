// C ABI of color_modem_b200 (include/color_modem_b200.h): handle management, filter-table construction,
// kernel launches.  Host code only; the arithmetic lives in cm_*.cuh.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <new>
#include <vector>

#include "../../include/color_modem_b200.h"
#include "cm_common.cuh"
#include "cm_qam.cuh"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return fail(CM_ERR_CUDA, #expr ": %s", cudaGetErrorString(_e)); \
    } while (0)

struct cm_modem {
    cm_desc desc;
    int precision;
    int device;
    int sm_count;
    int smem_optin;
    DevParams<float> pf;
    DevParams<double> pd;
    void *d_tab = nullptr;
    void *d_taps = nullptr;
    bool timing = false;
    struct Ev { cudaEvent_t a, b; int id; };
    std::vector<Ev> events;
    // scratch for the *_host entry points
    void *d_in = nullptr, *d_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
};

// ------------------------------------------------------------------------------------------------------------
// filter tables (float64 on the host, cast to the handle's precision)
// ------------------------------------------------------------------------------------------------------------
static void mat_mul(const double a[4], const double b[4], double out[4]) {
    double r[4] = {a[0] * b[0] + a[1] * b[2], a[0] * b[1] + a[1] * b[3], a[2] * b[0] + a[3] * b[2],
                   a[2] * b[1] + a[3] * b[3]};
    memcpy(out, r, sizeof(r));
}

// Chunk length per lane: the kernels instantiate warp_iir for a few compile-time values per rate
// (cm_iir.cuh: warp_iir); pick the one that wastes the least padding, preferring longer chunks on ties.
static void pick_chunk(int rate, int total, int &L, int &nsuper) {
    static const int c1[] = {47, 31, 23}, c2[] = {46, 30}, c3[] = {45, 39};
    const int *cand = rate == 1 ? c1 : (rate == 2 ? c2 : c3);
    const int ncand = rate == 1 ? 3 : 2;
    long best = -1;
    for (int i = 0; i < ncand; ++i) {
        int ns = (total + 32 * cand[i] - 1) / (32 * cand[i]);
        if (ns < 1) ns = 1;
        long padded = (long)ns * 32 * cand[i];
        if (best < 0 || padded < best) { best = padded; L = cand[i]; nsuper = ns; }
    }
}

static void build_filter(const cm_filter &f, FiltHdr &h, std::vector<double> &tab) {
    int L = 0, nsuper = 0;
    pick_chunk(f.rate, f.n + f.shift, L, nsuper);
    h.nsec = f.nsec;
    h.shift = f.shift;
    h.n = f.n;
    h.L = L;
    h.nsuper = nsuper;
    h.rate = f.rate;
    h.npad = 32 * L * nsuper;
    h.off = (int)tab.size();
    for (int s = 0; s < f.nsec; ++s) {
        const double *c = f.sos[s];
        std::vector<double> sec(CM_SEC_STRIDE, 0.0);
        for (int i = 0; i < 5; ++i) sec[i] = c[i];
        const double A[4] = {-c[3], 1.0, -c[4], 0.0};
        double M[4] = {1.0, 0.0, 0.0, 1.0};
        for (int i = 0; i < L; ++i) {                     // M = A^L
            double Q[4];
            mat_mul(A, M, Q);
            memcpy(M, Q, sizeof(M));
        }
        for (int k = 0; k < 5; ++k) {                     // M, M^2, M^4, M^8, M^16
            for (int i = 0; i < 4; ++i) sec[CM_SEC_MPOW + 4 * k + i] = M[i];
            double Q[4];
            mat_mul(M, M, Q);
            memcpy(M, Q, sizeof(M));
        }
        tab.insert(tab.end(), sec.begin(), sec.end());
    }
}

template <typename T>
static void fill_params(const cm_desc &d, DevParams<T> &p, const FiltHdr *fh, const ResHdr *rh, const void *tab,
                        const void *taps) {
    memset(&p, 0, sizeof(p));
    p.kind = d.kind;
    p.flags = d.flags;
    p.W = d.width;
    p.H = d.height;
    p.Wc = d.comp_width;
    p.Wo = d.out_width;
    p.digital_shift = d.digital_shift;
    p.odd_first = d.odd_first;
    p.even_first = d.even_first;
    p.ref_line = d.ref_line;
    p.frame_cycle = d.frame_cycle;
    p.frame_shift = d.frame_shift_turns;
    p.line_shift = d.line_shift_turns;
    for (int i = 0; i < CM_NPHASE; ++i) p.phases[i] = d.phases[i];
    for (int i = 0; i < CM_NSCAL; ++i) p.scalars[i] = (T)d.scalars[i];
    for (int i = 0; i < 9; ++i) { p.enc[i] = (T)d.enc_matrix[i]; p.dec[i] = (T)d.dec_matrix[i]; }
    for (int i = 0; i < CM_NFILT; ++i) p.filt[i] = fh[i];
    // line-buffer geometry: every buffer that feeds an IIR is padded to the site's 32*L*nsuper
    auto up4 = [](int v) { return (v + 3) & ~3; };
    int n1p = up4(d.width > d.comp_width ? d.width : d.comp_width), n2p = 2 * up4(d.comp_width),
        n3p = 3 * up4(d.comp_width);
    for (int i = 0; i < CM_NFILT; ++i) {
        if (!fh[i].nsec) continue;
        if (fh[i].rate == 1 && fh[i].npad > n1p) n1p = fh[i].npad;
        if (fh[i].rate == 2 && fh[i].npad > n2p) n2p = fh[i].npad;
        if (fh[i].rate == 3 && fh[i].npad > n3p) n3p = fh[i].npad;
    }
    p.n1p = up4(n1p);
    p.hb2 = up4((n2p + 1) / 2);
    if (p.hb2 < p.n1p) p.hb2 = p.n1p;
    p.hb3 = up4((n3p + 2) / 3);
    for (int i = 0; i < CM_NRES; ++i) p.res[i] = rh[i];
    p.tab = (const T *)tab;
    p.taps = (const T *)taps;
}

template <typename T>
static int upload(const std::vector<double> &src, void **dst) {
    std::vector<T> tmp(src.size() ? src.size() : 1);
    for (size_t i = 0; i < src.size(); ++i) tmp[i] = (T)src[i];
    CUDA_TRY(cudaMalloc(dst, tmp.size() * sizeof(T)));
    CUDA_TRY(cudaMemcpy(*dst, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
    return CM_OK;
}

extern "C" int cm_abi_version(void) { return CM_ABI_VERSION; }
extern "C" const char *cm_last_error(void) { return g_err; }
extern "C" int64_t cm_launch_count(void) { return g_launches.load(); }

extern "C" int cm_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return CM_OK;
}

extern "C" int cm_create(const cm_desc *desc, int precision, cm_modem **out) {
    if (!desc || !out) return fail(CM_ERR_INVALID, "null argument%s");
    if (desc->abi_version != CM_ABI_VERSION) return fail(CM_ERR_INVALID, "cm_desc.abi_version mismatch%s");
    if (precision != CM_FP32 && precision != CM_FP64) return fail(CM_ERR_INVALID, "bad precision%s");
    if (desc->width <= 0 || desc->height <= 0 || desc->comp_width <= 0 || desc->out_width <= 0)
        return fail(CM_ERR_INVALID, "bad raster size%s");
    if (desc->nfilters < 0 || desc->nfilters > CM_MAX_FILTERS || desc->nresamplers < 0 ||
        desc->nresamplers > CM_MAX_RESAMPLERS)
        return fail(CM_ERR_INVALID, "bad filter / resampler count%s");
    if (desc->frame_cycle <= 0) return fail(CM_ERR_INVALID, "frame_cycle must be positive%s");
    if ((desc->width & 3) || (desc->comp_width & 3) || (desc->out_width & 3))
        return fail(CM_ERR_UNSUPPORTED, "line widths must be multiples of 4 samples%s");
    switch (desc->kind) {
        case CM_KIND_QAM_BANDSPLIT:
        case CM_KIND_NTSC_COMB:
        case CM_KIND_NTSC_3D:
        case CM_KIND_PAL_D:
        case CM_KIND_PAL_3D:
            break;
        default:
            return fail(CM_ERR_UNSUPPORTED, "modem kind not built%s");
    }
    cm_modem *m = new (std::nothrow) cm_modem();
    if (!m) return fail(CM_ERR_NOMEM, "out of host memory%s");
    m->desc = *desc;
    m->precision = precision;
    cudaError_t e = cudaGetDevice(&m->device);
    if (e != cudaSuccess) { delete m; return fail(CM_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e)); }
    cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device);
    cudaDeviceGetAttribute(&m->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device);

    FiltHdr fh[CM_NFILT];
    ResHdr rh[CM_NRES];
    memset(fh, 0, sizeof(fh));
    memset(rh, 0, sizeof(rh));
    std::vector<double> tab, taps;
    for (int i = 0; i < desc->nfilters; ++i) {
        const cm_filter &f = desc->filters[i];
        if (f.nsec == 0) continue;
        if (f.nsec < 0 || f.nsec > CM_MAX_SECTIONS || f.shift < 0 || f.n <= 0 || f.rate < 1 || f.rate > 3) {
            delete m;
            return fail(CM_ERR_INVALID, "bad cm_filter%s");
        }
        build_filter(f, fh[i], tab);
    }
    for (int i = 0; i < desc->nresamplers; ++i) {
        const cm_resampler &r = desc->resamplers[i];
        if (r.ntaps == 0) continue;
        if (r.ntaps < 0 || r.ntaps > CM_MAX_TAPS || r.up <= 0 || r.down <= 0 || r.ntaps != 2 * r.half + 1) {
            delete m;
            return fail(CM_ERR_INVALID, "bad cm_resampler%s");
        }
        rh[i].up = r.up;
        rh[i].down = r.down;
        rh[i].half = r.half;
        rh[i].ntaps = r.ntaps;
        rh[i].off = (int)taps.size();
        taps.insert(taps.end(), r.taps, r.taps + r.ntaps);
    }
    int rc;
    if (precision == CM_FP32) {
        rc = upload<float>(tab, &m->d_tab);
        if (rc == CM_OK) rc = upload<float>(taps, &m->d_taps);
        fill_params<float>(*desc, m->pf, fh, rh, m->d_tab, m->d_taps);
    } else {
        rc = upload<double>(tab, &m->d_tab);
        if (rc == CM_OK) rc = upload<double>(taps, &m->d_taps);
        fill_params<double>(*desc, m->pd, fh, rh, m->d_tab, m->d_taps);
    }
    if (rc != CM_OK) { cm_destroy(m); return rc; }
    *out = m;
    return CM_OK;
}

extern "C" void cm_destroy(cm_modem *m) {
    if (!m) return;
    cudaFree(m->d_tab);
    cudaFree(m->d_taps);
    cudaFree(m->d_in);
    cudaFree(m->d_out);
    cm_timing_reset(m);
    delete m;
}

// ------------------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------------------
struct LaunchTimer {
    cm_modem *m;
    cudaStream_t st;
    cm_modem::Ev ev;
    bool on;
    LaunchTimer(cm_modem *m_, int id, cudaStream_t st_) : m(m_), st(st_), on(m_->timing) {
        if (on) {
            ev.id = id;
            cudaEventCreate(&ev.a);
            cudaEventCreate(&ev.b);
            cudaEventRecord(ev.a, st);
        }
    }
    ~LaunchTimer() {
        if (on) {
            cudaEventRecord(ev.b, st);
            m->events.push_back(ev);
        }
    }
};

extern "C" int cm_timing_enable(cm_modem *m, int on) {
    if (!m) return fail(CM_ERR_INVALID, "null handle%s");
    m->timing = on != 0;
    return CM_OK;
}

extern "C" int cm_timing_reset(cm_modem *m) {
    if (!m) return fail(CM_ERR_INVALID, "null handle%s");
    for (auto &e : m->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    m->events.clear();
    return CM_OK;
}

extern "C" int cm_timing_read(cm_modem *m, int id, double *total_ms, int64_t *launches) {
    if (!m) return fail(CM_ERR_INVALID, "null handle%s");
    double tot = 0.0;
    int64_t n = 0;
    for (auto &e : m->events) {
        if (e.id != id) continue;
        CUDA_TRY(cudaEventSynchronize(e.b));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e.a, e.b));
        tot += ms;
        ++n;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return CM_OK;
}

template <typename T> static const DevParams<T> &params_of(const cm_modem *m);
template <> const DevParams<float> &params_of<float>(const cm_modem *m) { return m->pf; }
template <> const DevParams<double> &params_of<double>(const cm_modem *m) { return m->pd; }

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return CM_OK;
}

// Largest R in [1, rmax] whose shared-memory footprint fits `budget`; 0 if even R = 1 does not fit.
template <class F>
static int pick_rows(int rmax, size_t budget, F bytes_for) {
    for (int r = rmax; r >= 1; --r)
        if (bytes_for(r) <= budget) return r;
    return 0;
}

template <typename T>
static void set_groups(IoArgs<T> &io, int R) {
    io.rows_per_cta = R;
    int rows_in_field = (io.out_count + 1) >> 1;
    io.groups_per_field = (rows_in_field + R - 1) / R;
}

template <typename T>
static int launch_encode(cm_modem *m, IoArgs<T> io, cudaStream_t st) {
    const DevParams<T> &p = params_of<T>(m);
    if (io.out_count <= 0) return CM_OK;
    switch (p.kind) {
        case CM_KIND_QAM_BANDSPLIT:
        case CM_KIND_NTSC_COMB:
        case CM_KIND_NTSC_3D:
        case CM_KIND_PAL_D:
        case CM_KIND_PAL_3D: {
            auto bytes = [&](int r) { return (size_t)r * 3 * p.n1p * sizeof(T); };
            int R = pick_rows(4, (size_t)m->smem_optin / 2, bytes);
            if (!R) return fail(CM_ERR_UNSUPPORTED, "line too wide for the encode kernel%s");
            set_groups(io, R);
            int rc = set_smem(k_qam_encode<T>, bytes(R));
            if (rc) return rc;
            dim3 grid((unsigned)(io.nframes * 2 * io.groups_per_field));
            LaunchTimer lt(m, CM_K_ENCODE, st);
            k_qam_encode<T><<<grid, 64 * R, bytes(R), st>>>(p, io);
            g_launches++;
            break;
        }
        default:
            return fail(CM_ERR_UNSUPPORTED, "encode: modem kind not built%s");
    }
    CUDA_TRY(cudaGetLastError());
    return CM_OK;
}

template <typename T>
static int launch_bandsplit(cm_modem *m, IoArgs<T> io, int luma_mode, cudaStream_t st) {
    const DevParams<T> &p = params_of<T>(m);
    if (io.out_count <= 0) return CM_OK;
    auto bytes = [&](int r) { return (128 + (size_t)r * (p.n1p + 8 * (size_t)p.hb2)) * sizeof(T); };
    int R = pick_rows(4, (size_t)m->smem_optin / 2, bytes);
    if (!R) R = pick_rows(1, (size_t)m->smem_optin, bytes);
    if (!R) return fail(CM_ERR_UNSUPPORTED, "line too wide for the band-split kernel%s");
    set_groups(io, R);
    int rc = set_smem(k_qam_bandsplit<T>, bytes(R));
    if (rc) return rc;
    dim3 grid((unsigned)(io.nframes * 2 * io.groups_per_field));
    LaunchTimer lt(m, CM_K_BANDSPLIT, st);
    k_qam_bandsplit<T><<<grid, 64 * R, bytes(R), st>>>(p, io, luma_mode);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CM_OK;
}

template <typename T>
static int launch_pald(cm_modem *m, IoArgs<T> io, cudaStream_t st) {
    const DevParams<T> &p = params_of<T>(m);
    if (io.out_count <= 0) return CM_OK;
    auto bytes = [&](int r) {
        return (128 + (size_t)(r + 1) * (p.n1p + 2 * (size_t)p.hb2) + (size_t)r * 4 * p.hb2) * sizeof(T);
    };
    int R = pick_rows(4, (size_t)m->smem_optin / 2, bytes);
    if (!R) R = pick_rows(2, (size_t)m->smem_optin, bytes);
    if (!R) return fail(CM_ERR_UNSUPPORTED, "line too wide for the PAL-D kernel%s");
    set_groups(io, R);
    int rc = set_smem(k_pald_combed<T>, bytes(R));
    if (rc) return rc;
    dim3 grid((unsigned)(io.nframes * 2 * io.groups_per_field));
    LaunchTimer lt(m, CM_K_PALD, st);
    k_pald_combed<T><<<grid, 64 * R, bytes(R), st>>>(p, io);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CM_OK;
}

template <typename T, int MODE>
static int launch_comb(cm_modem *m, IoArgs<T> io, cudaStream_t st) {
    const DevParams<T> &p = params_of<T>(m);
    if (io.out_count <= 0) return CM_OK;
    auto bytes = [&](int r) {
        return (128 + (size_t)(r + 2) * (p.n1p + 2 * (size_t)p.hb2) + (size_t)r * 4 * p.hb2) * sizeof(T);
    };
    int R = pick_rows(4, (size_t)m->smem_optin / 2, bytes);
    if (!R) R = pick_rows(2, (size_t)m->smem_optin, bytes);
    if (!R) return fail(CM_ERR_UNSUPPORTED, "line too wide for the comb kernel%s");
    set_groups(io, R);
    int rc = set_smem(k_qam_comb<T, MODE>, bytes(R));
    if (rc) return rc;
    dim3 grid((unsigned)(io.nframes * 2 * io.groups_per_field));
    LaunchTimer lt(m, CM_K_COMB, st);
    k_qam_comb<T, MODE><<<grid, 64 * R, bytes(R), st>>>(p, io);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return CM_OK;
}

// rows of [begin, begin+count) that have no predecessor in the window (r < 2) / that have one (r >= 2)
template <typename T>
static void split_top(const IoArgs<T> &io, IoArgs<T> &top, IoArgs<T> &rest) {
    top = io;
    rest = io;
    int end = io.out_begin + io.out_count;
    int top_end = end < 2 ? end : 2;
    top.out_count = top_end > io.out_begin ? top_end - io.out_begin : 0;
    int rest_begin = io.out_begin > 2 ? io.out_begin : 2;
    rest.out_begin = rest_begin;
    rest.out_count = end > rest_begin ? end - rest_begin : 0;
}

template <typename T>
static int launch_decode(cm_modem *m, IoArgs<T> io, int mode, cudaStream_t st) {
    const DevParams<T> &p = params_of<T>(m);
    if (mode == CM_MODE_BANDSPLIT_NOSTRIP) {
        if (p.kind < CM_KIND_QAM_BANDSPLIT || p.kind > CM_KIND_PAL_3D)
            return fail(CM_ERR_INVALID, "CM_MODE_BANDSPLIT_NOSTRIP is only defined for the QAM family%s");
        return launch_bandsplit<T>(m, io, 2, st);
    }
    switch (p.kind) {
        case CM_KIND_QAM_BANDSPLIT:
            return launch_bandsplit<T>(m, io, 0, st);
        case CM_KIND_PAL_D: {
            IoArgs<T> top, rest;
            split_top(io, top, rest);
            int rc = launch_bandsplit<T>(m, top, 0, st);
            if (rc) return rc;
            return launch_pald<T>(m, rest, st);
        }
        case CM_KIND_NTSC_COMB: {
            IoArgs<T> top, rest;
            split_top(io, top, rest);
            int rc = launch_bandsplit<T>(m, top, 0, st);
            if (rc) return rc;
            if (p.flags & CM_FLAG_NTSC_NO_COMB)      // ntsc.py:71-72: chroma of the row itself, luma = c - remod
                return launch_bandsplit<T>(m, rest, 1, st);
            return launch_comb<T, COMB_NTSC2>(m, rest, st);
        }
        case CM_KIND_NTSC_3D:
            if (p.flags & CM_FLAG_NTSC_NO_COMB) return launch_bandsplit<T>(m, io, 1, st);
            return launch_comb<T, COMB_NTSC3>(m, io, st);
        case CM_KIND_PAL_3D: {
            if (!(p.flags & (CM_FLAG_PAL3D_SIN | CM_FLAG_PAL3D_COS))) {   // pal.py:181-182: plain PAL-D
                IoArgs<T> top, rest;
                split_top(io, top, rest);
                int rc = launch_bandsplit<T>(m, top, 0, st);
                if (rc) return rc;
                return launch_pald<T>(m, rest, st);
            }
            IoArgs<T> top, rest;
            split_top(io, top, rest);
            int rc = launch_bandsplit<T>(m, top, 1, st);
            if (rc) return rc;
            return launch_comb<T, COMB_PAL3>(m, rest, st);
        }
        default:
            return fail(CM_ERR_UNSUPPORTED, "decode: modem kind not built%s");
    }
}

template <typename T>
static int run_ex(cm_modem *m, bool encode, const cm_window *win, const uint8_t *in_u8, const void *in_f,
                  uint8_t *out_u8, void *out_f, int64_t first_frame, int32_t nframes, void *stream) {
    if (!m) return fail(CM_ERR_INVALID, "null handle%s");
    if ((in_u8 == nullptr) == (in_f == nullptr)) return fail(CM_ERR_INVALID, "exactly one input buffer required%s");
    if (!out_u8 && !out_f) return fail(CM_ERR_INVALID, "no output buffer%s");
    if (nframes < 0 || first_frame < 0) return fail(CM_ERR_INVALID, "bad frame range%s");
    if (nframes == 0) return CM_OK;
    IoArgs<T> io;
    memset(&io, 0, sizeof(io));
    io.in_u8 = in_u8;
    io.in_f = (const T *)in_f;
    io.out_u8 = out_u8;
    io.out_f = (T *)out_f;
    io.first_frame = first_frame;
    io.nframes = nframes;
    if (win) {
        if (win->nrows <= 0 || win->out_begin < 0 || win->out_count < 0 || win->out_begin + win->out_count > win->nrows)
            return fail(CM_ERR_INVALID, "bad cm_window%s");
        io.nrows = win->nrows;
        io.y0 = win->y0;
        io.out_begin = win->out_begin;
        io.out_count = win->out_count;
    } else {
        io.nrows = m->desc.height;
        io.y0 = 0;
        io.out_begin = 0;
        io.out_count = m->desc.height;
    }
    CUDA_TRY(cudaSetDevice(m->device));
    const int mode = win ? win->mode : CM_MODE_DEFAULT;
    return encode ? launch_encode<T>(m, io, (cudaStream_t)stream)
                  : launch_decode<T>(m, io, mode, (cudaStream_t)stream);
}

extern "C" int cm_encode_ex(cm_modem *m, const cm_window *win, const uint8_t *rgb_u8, const void *rgb_float,
                            uint8_t *comp_u8, void *comp_float, int64_t first_frame, int32_t nframes,
                            void *stream) {
    if (!m) return fail(CM_ERR_INVALID, "null handle%s");
    return m->precision == CM_FP32
               ? run_ex<float>(m, true, win, rgb_u8, rgb_float, comp_u8, comp_float, first_frame, nframes, stream)
               : run_ex<double>(m, true, win, rgb_u8, rgb_float, comp_u8, comp_float, first_frame, nframes, stream);
}

extern "C" int cm_decode_ex(cm_modem *m, const cm_window *win, const uint8_t *comp_u8, const void *comp_float,
                            uint8_t *rgb_u8, void *rgb_float, int64_t first_frame, int32_t nframes, void *stream) {
    if (!m) return fail(CM_ERR_INVALID, "null handle%s");
    return m->precision == CM_FP32
               ? run_ex<float>(m, false, win, comp_u8, comp_float, rgb_u8, rgb_float, first_frame, nframes, stream)
               : run_ex<double>(m, false, win, comp_u8, comp_float, rgb_u8, rgb_float, first_frame, nframes, stream);
}

extern "C" int cm_encode_frames(cm_modem *m, const uint8_t *rgb, uint8_t *comp, int64_t first_frame,
                                int32_t nframes, void *stream) {
    return cm_encode_ex(m, nullptr, rgb, nullptr, comp, nullptr, first_frame, nframes, stream);
}

extern "C" int cm_decode_frames(cm_modem *m, const uint8_t *comp, uint8_t *rgb, int64_t first_frame,
                                int32_t nframes, void *stream) {
    return cm_decode_ex(m, nullptr, comp, nullptr, rgb, nullptr, first_frame, nframes, stream);
}

static int ensure(void **buf, size_t *cap, size_t need) {
    if (*cap >= need) return CM_OK;
    cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    CUDA_TRY(cudaMalloc(buf, need));
    *cap = need;
    return CM_OK;
}

extern "C" int cm_encode_frames_host(cm_modem *m, const uint8_t *rgb, uint8_t *comp, int64_t first_frame,
                                     int32_t nframes) {
    if (!m || !rgb || !comp) return fail(CM_ERR_INVALID, "null argument%s");
    CUDA_TRY(cudaSetDevice(m->device));
    size_t in_b = (size_t)nframes * m->desc.height * m->desc.width * 3;
    size_t out_b = (size_t)nframes * m->desc.height * m->desc.comp_width;
    int rc = ensure(&m->d_in, &m->in_cap, in_b);
    if (!rc) rc = ensure(&m->d_out, &m->out_cap, out_b);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(m->d_in, rgb, in_b, cudaMemcpyHostToDevice, 0));
    rc = cm_encode_frames(m, (const uint8_t *)m->d_in, (uint8_t *)m->d_out, first_frame, nframes, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(comp, m->d_out, out_b, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return CM_OK;
}

extern "C" int cm_decode_frames_host(cm_modem *m, const uint8_t *comp, uint8_t *rgb, int64_t first_frame,
                                     int32_t nframes) {
    if (!m || !rgb || !comp) return fail(CM_ERR_INVALID, "null argument%s");
    CUDA_TRY(cudaSetDevice(m->device));
    size_t in_b = (size_t)nframes * m->desc.height * m->desc.comp_width;
    size_t out_b = (size_t)nframes * m->desc.height * m->desc.out_width * 3;
    int rc = ensure(&m->d_in, &m->in_cap, in_b);
    if (!rc) rc = ensure(&m->d_out, &m->out_cap, out_b);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(m->d_in, comp, in_b, cudaMemcpyHostToDevice, 0));
    rc = cm_decode_frames(m, (const uint8_t *)m->d_in, (uint8_t *)m->d_out, first_frame, nframes, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(rgb, m->d_out, out_b, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return CM_OK;
}
