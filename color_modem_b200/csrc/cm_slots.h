/* Slot assignment of cm_desc.filters[] / resamplers[] / scalars[] / phases[] per modem family.
 * Mirrored by color_modem_b200/_slots.py (kept in sync by tests/test_abi.py). */
#ifndef CM_SLOTS_H
#define CM_SLOTS_H

/* ---- QAM family: NtscModem, PalSModem and the comb decoders built on them (qam.py, ntsc.py, pal.py, comb.py) */
#define QF_PRE_LP 0      /* QamColorModem._chroma_precorrect_lowpass  qam.py:16   n = W   */
#define QF_BP2X 1        /* _extract_chroma2x                         qam.py:17   n = 2W  */
#define QF_BS2X 2        /* _remove_chroma2x                          qam.py:17   n = 2W  */
#define QF_DEMOD_LP 3    /* _demod_lowpass                            qam.py:18   n = 2W  */
#define QF_PALD_LP 4     /* PalDModem._filter                         pal.py:67   n = 2W  */
#define QR_UP2 0         /* resample_poly(up=2, down=1) */
#define QR_DOWN2 1       /* resample_poly(up=1, down=2) */
#define QP_STEP1X 0      /* carrier advance per sample at 1x rate, turns (2*carrier_phase_step/2pi, qam.py:24) */
#define QP_STEP2X 1      /* carrier advance per sample at 2x rate (carrier_phase_step/2pi, qam.py:47)          */
#define QP_BP_SHIFT 2    /* extract_chroma_phase_shift / 2pi  (qam.py:39-41)                                   */
#define QP_HALF_LS 3     /* 0.5 * line_shift / 2pi            (ntsc.py:74, pal.py:113-114)                      */
#define QS_NTSC_FACTOR 0 /* NtscCombModem._factor = 0.5 / sin(LS/2)          ntsc.py:55-59 */
#define QS_PALD_SIN 1    /* sin(LS/2)                                         pal.py:65     */
#define QS_PALD_COS 2    /* cos(LS/2)                                         pal.py:66     */
#define QS_P3D_SINSUM 3  /* 0.5 / sin(LS)                                     pal.py:168-169 */
#define QS_P3D_COSU 4    /* -0.5 / (1 - cos(LS))                              pal.py:171-173 */
#define QS_P3D_COSV 5    /* -0.5 / (1 + cos(LS))                                             */

#endif
