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
#define QF_NOTCH 5       /* comb._notch(backend, q): luma notch at fsc comb.py:18-20 n = W  (optional) */
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


/* ---- SECAM (secam.py) */
#define SF_PRE_LP 0      /* _chroma_precorrect_lowpass   secam.py:171   rate 1, n = W            */
#define SF_PRE_EMPH 1    /* _chroma_precorrect           secam.py:176   rate 1, n = W  (optional) */
#define SF_LUMA_BS 2     /* _chroma_demod_luma_filter    secam.py:185   rate 1, n = W            */
#define SF_CHROMA_BP 3   /* _chroma_demod_chroma_filter  secam.py:183   rate 1, n = W + W/40 - 1 */
#define SF_ANTI_BELL 4   /* _chroma_demod_bell           secam.py:169   rate 1, same n (optional) */
#define SF_FM_LP 5       /* FmDecoder._lowpass           secam.py:131   rate 2, n = 2 (W + W/40 - 1) */
#define SF_DE_EMPH 6     /* _reverse_chroma_precorrect   secam.py:176   rate 1, n = W  (optional) */
#define SR_UP2 0
#define SR_DOWN2 1
#define SP_FM_STEP2X 0   /* FmDecoder mixing phase per 2x sample: fc/4 turns (secam.py:137) */
#define SP_FSC_DR_HALF 1 /* pi*fsc_dr per sample = fsc_dr/2 turns (secam.py:244) */
#define SP_FSC_DB_HALF 2
#define SP_INVERSIONS 3  /* bit i = _start_phase_inversions[i] (secam.py:163-166) -- raw integer */
#define SS_FSC_DR 0      /* normalised (Nyquist = 1) frequencies, secam.py:156-162 */
#define SS_FSC_DB 1
#define SS_FDEV_DR 2
#define SS_FDEV_DB 3
#define SS_F_LO 4
#define SS_F_HI 5
#define SS_BELL_F0 6
#define SS_M0 7
#define SS_KN 8
#define SS_KD 9
#define SS_FM_FC 10      /* FmDecoder centre frequency (secam.py:179,187) */


/* ---- NIIR / SECAM-IV (niir.py) */
#define NF_PRE_LP 0      /* _chroma_precorrect_lowpass              niir.py:17   rate 1, n = W  */
#define NF_BASE_LP 1     /* _demodulate_upsampled_baseband_filter   niir.py:21   rate 3, n = 3W */
#define NF_UP_BP 2       /* _demodulate_upsampled_filter            niir.py:21   rate 3, n = 3W */
#define NR_UP3 0
#define NR_DOWN3 1
#define NP_STEP1X 0      /* _carrier_phase_step / 2pi = fsc/fs turns per sample (niir.py:12) */
#define NP_LINE_SHIFT 1  /* line_shift / 2pi                                                  */
#define NP_LUMA_ROT 2    /* (pi - up_bp.phase_shift) / 2pi   (niir.py:154)                    */
#define NS_INV_STEP3 0   /* resample_factor / carrier_phase_step = 3 / (2 pi fsc/fs)  (niir.py:124-125) */

/* ---- 819-line AM proto-SECAM (protosecam.py) */
#define PF_PRE_LP 0      /* _chroma_precorrect_lowpass     protosecam.py:34   rate 1, n = W  */
#define PF_BP_UP 1       /* _extract_chroma_up             protosecam.py:37   rate 3, n = 3W */
#define PF_BS_UP 2       /* _remove_chroma_up              protosecam.py:37   rate 3, n = 3W */
#define PF_POST_LP 3     /* _chroma_up_post_demod_filter   protosecam.py:46   rate 3, n = 3W */
#define PR_UP3 0
#define PR_DOWN3 1
#define PP_STEP1X 0      /* 2 * _carrier_phase_step / 2pi = fsc/fs turns per sample (protosecam.py:33,88) */

/* ---- D2-MAC (mac.py) */
#define MR_LUMA_IN 0     /* W -> 720          mac.py:47-51  */
#define MR_CHROMA_IN 1   /* W -> 360          mac.py:48-54  */
#define MR_OUT 2         /* 1080 -> width     mac.py:70-73  */
#define MR_COMP_IN 3     /* width -> 1080     mac.py:81-83  */
#define MR_UP2 4         /* 360 -> 720        mac.py:111    */

#endif
