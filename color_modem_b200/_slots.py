"""Slot numbers of cm_desc arrays — mirror of csrc/cm_slots.h (tests/test_abi.py keeps them in sync)."""
# QAM family
QF_PRE_LP, QF_BP2X, QF_BS2X, QF_DEMOD_LP, QF_PALD_LP = 0, 1, 2, 3, 4
QR_UP2, QR_DOWN2 = 0, 1
QP_STEP1X, QP_STEP2X, QP_BP_SHIFT, QP_HALF_LS = 0, 1, 2, 3
QS_NTSC_FACTOR, QS_PALD_SIN, QS_PALD_COS, QS_P3D_SINSUM, QS_P3D_COSU, QS_P3D_COSV = 0, 1, 2, 3, 4, 5
