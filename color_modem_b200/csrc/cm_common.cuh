// Common device-side types of the colour-modem kernels (sm_100a).
//
// Execution model:
//   * a CTA owns a small group of consecutive rows of ONE field of one frame; every intermediate signal of
//     those rows lives in shared memory, HBM is touched once per input and once per output byte;
//   * FIR / elementwise stages are parallel over samples (all threads of the CTA);
//   * IIR stages (the reference's scipy.signal.lfilter calls, utils.py:28-36) are parallel over *chunks of a
//     line*: one warp per (row, signal), lane t owns samples [t*L, (t+1)*L) in registers, runs the biquad
//     cascade from zero state, and the true chunk-boundary states are recovered exactly with a warp-shuffle
//     scan over the 2x2 state-transition powers (cm_iir.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CM_NWARPS 8
#define CM_NTHREADS (CM_NWARPS * 32)
#define CM_MAXSEC 6
#define CM_NFILT 14
#define CM_NRES 6
#define CM_NSCAL 48
#define CM_NPHASE 16
#define CM_TAPS_ELEMS 256   // head of every decode kernel's shared memory: resampler taps [0,128) + IIR team scratch [128,256)
#define CM_QUARTER_TURN 0x4000000000000000ull

struct FiltHdr {
    int nsec, shift, n, L, nsuper, rate;     // n = input length of this use-site; L*32*nsuper >= n+shift
    int off;                                 // offset (in elements) of the section tables in DevParams::tab
    int npad;                                // 32 * L * nsuper
};

struct ResHdr {
    int up, down, half, ntaps;
    int off;                                 // offset (in elements) into DevParams::taps
    int _pad[3];
};

// Rational resampler (MAC) in aligned polyphase form, built by the host (cm_api.cu: build_poly).  The `up` outputs
// j = up m + r (r = 0 .. up-1) of one group m read overlapping stretches of the line, so one thread computes the whole
// group from ONE window of KU samples x[s .. s + KU), s = (m down + lo0 + FP) & ~3, of a zero-padded line (FP zeros in
// front, KU behind), against the taps G[((a * up) + r) * stride + q], a = (m down + lo0 + FP) & 3.  up <= 4.
struct PolyHdr {
    int up, down, KU, stride, FP, off;       // off: offset (elements) of G in DevParams::ptab; up == 0: not built
    int lo0, skew;                            // skew: mask of poly_skew (cm_fir.cuh), the same for every resampler of a handle
};

// Taps of the MAC resamplers whose shape matches one of the compiled ones, as a kernel parameter: read with compile-time
// indices they are constant-bank operands of the FMAs (cm_fir.cuh: fir_poly_ct_*), and the loop no longer competes with
// itself for shared-memory bandwidth (taps and line both came from shared memory: 55 multiply-adds per clock and SM).
//   luma / chroma: up = 3, `down` a multiple of 4 (every output group has the same window alignment): 3/8 and 3/16
//   out / comp:    up = 2 / 3, any `down`: one table per alignment class m mod 4 of the output group
struct MacShape { static constexpr int KL = 64, KC = 128, KO = 40, KI = 32; };
template <typename T>
struct alignas(16) MacConst {      // 16-byte aligned in the parameter bank: tap pairs / quads are single uniform loads
    T luma[3][MacShape::KL];
    T chroma[3][MacShape::KC];
    T out[4][2][MacShape::KO];
    T comp[4][3][MacShape::KI];
    int ok_luma, ok_chroma, ok_out, ok_comp;
};

template <typename T>
struct DevParams {
    int kind, flags;
    int ident_enc, ident_dec;   // the colour matrix is the identity (component-level handles): planes pass through untouched,
                                // so that signed zeros survive (numpy.arctan2 of niir.py:45,64 tells -0.0 from +0.0)
    int W, H, Wc, Wo;
    int digital_shift, odd_first, even_first, ref_line, frame_cycle;
    int n1p;          // padded length of 1x line buffers (multiple of 4, >= every 1x IIR site's npad)
    int hb2;          // elements per phase of 2x polyphase buffers (multiple of 4, >= n1p)
    int hb3;          // elements per phase of 3x polyphase buffers
    unsigned long long frame_shift, line_shift;
    unsigned long long phase0;     // cm_window.phase_offset of the current call (explicit start phase, qam.py:28,43)
    unsigned long long phases[CM_NPHASE];
    T scalars[CM_NSCAL];
    T enc[9];
    T dec[9];
    double encd[9];   // encoder matrix in float64 (NIIR: the *direction* of the chroma vector is decided in float64 with the
                      // reference's own operation order, see cm_niir.cuh)
    FiltHdr filt[CM_NFILT];
    ResHdr res[CM_NRES];
    const T *tab;
    const T *taps;
    const T *ptab;    // MAC: polyphase tap tables of the resamplers (PolyHdr::off)
    PolyHdr poly[CM_NRES];
    int mac_skew;         // mask of poly_skew: -1 when some resampler steps by a multiple of 8 samples, else 0
    int mac_fp, mac_bp;   // zero padding (elements, multiples of 4) in front of / behind every line a resampler reads
    const T *ctab;    // k_qam_rows2: row-independent carrier table [sin | cos][npad of the QF_ROW_LP site] (cm_api.cu)
    int enc_geo;      // k_qam_encode_row2 / k_niir_encode2: geometry 1..3 (cm_api.cu: plan_encode_kernel); 0 = not served
    int row_geo;      // k_qam_rows2: geometry 1..3 (cm_qam.cuh: RowL) / k_secam_decode2: 1 or 3 (cm_secam.cuh: SecGeo);
                      // 0 = the row kernel does not serve this line length
    // dense taps of resampler slots 0 and 1 (the x2 / x3 half-band pair of every family except MAC), zero-filled:
    // read with compile-time indices they become constant-bank operands of the FIR FMAs (FFMA R, R, c[0][..], R
    // issues at full rate; with the tap in a third register the sm_100 register file caps FFMA at ~0.7 / clk,
    // tools/ubench/fma_forms.cu)
    T firc[2][64];
    float fircp[2][64][2];     // the same taps duplicated (h, h): 64-bit uniform operands of the packed f32x2 FIRs
};

template <typename T> struct IsF32 { static const bool value = false; };
template <> struct IsF32<float> { static const bool value = true; };

// Launch geometry / buffers of one call.
template <typename T>
struct IoArgs {
    const uint8_t *in_u8;
    const T *in_f;
    uint8_t *out_u8;
    T *out_f;
    long long first_frame;
    int nframes;
    int nrows;        // rows per frame in the buffers (window)
    int y0;           // line number of buffer row 0
    int out_begin, out_count;
    int rows_per_cta; // R
    int groups_per_field;
    T *aux;                 // kernel-specific scratch in global memory (line-sequential decoders: per-row luma | chroma)
    T *yuv;                 // non-null: store_rgb4 writes the planes (y, c1, c2) of the row to yuv[frame][row][3][Wo] instead
                            // of RGB (comb decoders with a luma notch: k_notch_rows finishes the row)
    unsigned long long *prof;   // optional per-phase cycle counters (tuning aid, cm_phase_profile); nullptr normally
};

// Phase timer: thread 0 of every CTA accumulates the cycles between consecutive marks.
struct PhaseClock {
    unsigned long long *prof;
    long long t;
    int idx;
    __device__ __forceinline__ PhaseClock(unsigned long long *p) : prof(p), t(0), idx(0) {
        if (prof && threadIdx.x == 0) t = clock64();
    }
    __device__ __forceinline__ void mark() {
        if (prof && threadIdx.x == 0) {
            const long long now = clock64();
            atomicAdd(prof + idx, (unsigned long long)(now - t));
            t = now;
        }
        ++idx;
    }
};

// ------------------------------------------------------------------------------------------------------------
// Real-number traits
// ------------------------------------------------------------------------------------------------------------
template <typename T> struct Real;

template <> struct Real<float> {
    static __device__ __forceinline__ void sincos_turns(unsigned long long ph, float &s, float &c) {
        // ph: phase in turns, 0.64 fixed point.  Top 32 bits as signed => [-0.5, 0.5) turns => [-1, 1) half-turns.
        float x = (float)(int)(unsigned)(ph >> 32) * 4.656612873077393e-10f;   // 2^-31
        sincospif(x, &s, &c);
    }
    // the same from the hardware approximations (MUFU.SIN / MUFU.COS, absolute error < 5e-7 on [-pi, pi]): for carriers
    // that are evaluated per sample or per pixel quad and multiply a chroma amplitude well below 1 (encoders; the
    // parity bound is 1e-4); sincospif costs ~28 instructions, this 5
    static __device__ __forceinline__ void sincos_turns_fast(unsigned long long ph, float &s, float &c) {
        const float x = (float)(int)(unsigned)(ph >> 32) * 1.4629180792671596e-09f;   // 2 pi * 2^-32: radians in [-pi, pi)
        s = __sinf(x);
        c = __cosf(x);
    }
    static __device__ __forceinline__ float from_u8(unsigned v) { return (float)v * (1.0f / 255.0f); }
    static __device__ __forceinline__ float rsqrt_(float x) { return rsqrtf(x); }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    // atan2 in ~20 instructions (atan2f: ~45): octant reduction to a = min/max in [0, 1], a * P(a^2) with a degree-8
    // near-minimax P (max error 1.0e-7 rad in fp32 evaluation, fitted with numpy; the FM discriminator amplifies
    // phase errors by ~15, the parity bound is 1e-4), then the reflections.  atan2(+-0, x<0) = +-pi as in numpy.
    static __device__ __forceinline__ float atan2_(float y, float x) {
        const float ax = fabsf(x), ay = fabsf(y);
        const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
        const float a = mx > 0.f ? __fdividef(mn, mx) : 0.f;
        const float s = a * a;
        float p = 2.456721384e-03f;
        p = fmaf(p, s, -1.440134458e-02f);
        p = fmaf(p, s, 3.978120163e-02f);
        p = fmaf(p, s, -7.234855741e-02f);
        p = fmaf(p, s, 1.049894542e-01f);
        p = fmaf(p, s, -1.416122913e-01f);
        p = fmaf(p, s, 1.998590678e-01f);
        p = fmaf(p, s, -3.333259821e-01f);
        p = fmaf(p, s, 9.999998808e-01f);
        float r = a * p;
        r = ay > ax ? 1.57079632679489661923f - r : r;
        r = x < 0.f ? 3.14159265358979323846f - r : r;
        return copysignf(r, y);
    }
    static __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
};

template <> struct Real<double> {
    static __device__ __forceinline__ void sincos_turns(unsigned long long ph, double &s, double &c) {
        double x = (double)(long long)ph * 1.084202172485504434e-19;           // 2^-63 => half-turns in [-1, 1)
        sincospi(x, &s, &c);
    }
    static __device__ __forceinline__ void sincos_turns_fast(unsigned long long ph, double &s, double &c) { sincos_turns(ph, s, c); }
    static __device__ __forceinline__ double from_u8(unsigned v) { return (double)v / 255.0; }
    static __device__ __forceinline__ double rsqrt_(double x) { return 1.0 / sqrt(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
    static __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
};

// 1 / x: float32 takes the hardware approximation (MUFU.RCP, ~1 ulp) instead of the IEEE division sequence; the float64
// verification build divides.
// Geometries of the u8 row encoders (k_qam_encode_row2, k_niir_encode2): warps per CTA, pixel quads per thread, chunk length
// of the chroma low-pass site (cm_api.cu: plan_encode_kernel)
template <int GEO> struct EncGeo;
template <> struct EncGeo<1> { static constexpr int NW = 2, KQ = 3, PRE = 23; };
template <> struct EncGeo<2> { static constexpr int NW = 4, KQ = 3, PRE = 23; };
template <> struct EncGeo<3> { static constexpr int NW = 4, KQ = 4, PRE = 31; };

template <typename T> struct FastRcp;
template <> struct FastRcp<float> { static __device__ __forceinline__ float rcp(float x) { return __fdividef(1.0f, x); } };
template <> struct FastRcp<double> { static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; } };

// uint8(rint(255 * clip(v, 0, 1)))  — reference image.py:7-8 (numpy.rint = round-half-even = cvt.rni)
template <typename T>
__device__ __forceinline__ unsigned to_u8(T v) {
    v = v < (T)0 ? (T)0 : (v > (T)1 ? (T)1 : v);
    return (unsigned)__double2int_rn((double)(v * (T)255));
}
template <>
__device__ __forceinline__ unsigned to_u8<float>(float v) {
    return (unsigned)__float2int_rn(__saturatef(v) * 255.0f);
}

// ------------------------------------------------------------------------------------------------------------
// Raster / carrier phase — reference line.py:57-65, utils.py:82-88
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int analog_line(const DevParams<T> &p, int line) {
    int adj = line + p.digital_shift;
    int half = adj >> 1;                       // floor division, also for negative adj (priming / field-top rows)
    return ((adj & 1) == 0 ? p.even_first : p.odd_first) + half;
}

template <typename T>
__device__ __forceinline__ bool is_alternate(const DevParams<T> &p, long long frame, int line) {
    return ((analog_line(p, line) & 1) == (int)(frame & 1));
}

// Subcarrier phase at the first sample of `line`, in turns (0.64 fixed point; integer wrap == mod 2*pi).
template <typename T>
__device__ __forceinline__ unsigned long long start_phase(const DevParams<T> &p, long long frame, int line) {
    unsigned long long fr = (unsigned long long)(frame % (long long)p.frame_cycle);
    long long dl = (long long)(analog_line(p, line) - p.ref_line);
    return fr * p.frame_shift + (unsigned long long)dl * p.line_shift + p.phase0;
}

// Row bookkeeping of a CTA: which buffer rows it outputs.
struct RowGroup {
    long long frame;      // absolute frame number
    int fidx;             // frame index inside the batch
    int r0;               // first output buffer row of the group
    int count;            // rows in the group (same field: r0, r0+2, ...)
};

// Output rows [out_begin, out_begin+out_count) split by parity into two "fields"; each field into groups of R.
// Grid: x = group inside the field, y = field, z = frame of the batch (no integer divisions in the prologue).
template <typename T>
__device__ __forceinline__ bool decode_group(const IoArgs<T> &io, RowGroup &g) {
    const int field = blockIdx.y, gi = blockIdx.x, f = blockIdx.z;
    const int first = io.out_begin + field;                         // first row of this parity class
    const int rows_in_field = (io.out_count - field + 1) >> 1;      // rows out_begin+field, +2, ... < out_begin+out_count
    const int start = gi * io.rows_per_cta;
    g.fidx = f;
    g.frame = io.first_frame + f;
    g.r0 = first + 2 * start;
    g.count = min(io.rows_per_cta, rows_in_field - start);
    return g.count > 0;
}
