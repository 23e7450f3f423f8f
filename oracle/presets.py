"""Standard presets (oracle; test infrastructure only).

Numbers restated from reference color_modem/color/ntsc.py:8-20, pal.py:9-25, secam.py:10-124,
protosecam.py:25, mac.py:9-12.
"""
import collections

Qam = collections.namedtuple('Qam', 'fsc bw3 bw20')

# NB: evaluation order matters at the ulp level (the reference writes  k * 15750.0 * 1000.0 / 1001.0).

NTSC = {
    'NTSC': Qam(227.5 * 15750.0 * 1000.0 / 1001.0, 1.3e6, 3.6e6),
    'NTSC_A': Qam(2657812.5, 1.0e6, 2.5e6),
    'NTSC_I': Qam(4429687.5, 1.3e6, 3.6e6),
    'NTSC443': Qam(4433618.75, 1.3e6, 3.6e6),
    'NTSC_N': Qam(3585937.5, 1.3e6, 3.6e6),
    'NTSC361': Qam(229.5 * 15750.0 * 1000.0 / 1001.0, 1.3e6, 3.6e6),
    # not a preset of the reference: a subcarrier at a whole multiple of the line rate, for which NtscCombModem gives up
    # combing (|sin(LS/2)| <= 0.05, ntsc.py:55-59,71-72); built from the reference's own NtscVariant(fsc=...) constructor
    'NTSC_NOCOMB': Qam(227.0 * 15750.0 * 1000.0 / 1001.0, 1.3e6, 3.6e6),
}

PAL = {
    'PAL': Qam(4433618.75, 1.3e6, 4.0e6),
    'PAL_M': Qam(227.25 * 15750.0 * 1000.0 / 1001.0, 1.3e6, 3.6e6),
    'PAL_N': Qam(3582056.25, 1.3e6, 3.6e6),
}

PROTOSECAM = {'SECAM_1957': Qam(8384512.5, 0.8e6, 2.0e6)}

Secam = collections.namedtuple(
    'Secam', 'fsc_dr fsc_db fdev_dr fdev_db flim_lo flim_hi m0 bell_f0 bell_kn bell_kd lf_f1 lf_k')

SECAM = {
    'SECAM_I': Secam(4437500.0, 4437500.0, 250e3, 250e3, -250e3, 250e3, 0.2, 4437500.0, 1.0, 1.0, 0.0, 1.0),
    'SECAM_II': Secam(4437500.0, 4437500.0, 250e3, 250e3, -250e3, 250e3, 0.1, 4437500.0, 16.0, 1.26, 0.0, 1.0),
    'SECAM_III': Secam(4437500.0, 4437500.0, 230e3, 230e3, -450e3, 350e3, 0.1, 4437500.0, 16.0, 1.26, 70e3, 5.6),
    'SECAM': Secam(4406250.0, 4250000.0, 280e3, 230e3, -386e3, 470250.0, 0.115, 4286000.0, 16.0, 1.26, 85e3, 3.0),
    'SECAM_A': Secam(2660000.0, 2660000.0, 250e3, 250e3, -250e3, 250e3, 0.2, 2660000.0, 1.0, 1.0, 0.0, 1.0),
    'SECAM_E': Secam(8370000.0, 8370000.0, 250e3, 250e3, -250e3, 250e3, 0.2, 8370000.0, 1.0, 1.0, 0.0, 1.0),
    'SECAM_M': Secam(227.5 * 15750.0 * 1000.0 / 1001.0, 227.5 * 15750.0 * 1000.0 / 1001.0, 230e3, 230e3, -500e3, 500e3, 0.1,
                     227.5 * 15750.0 * 1000.0 / 1001.0, 16.0, 1.26, 70e3, 5.6),
    'SECAM_N': Secam(3578125.0, 3578125.0, 230e3, 230e3, -500e3, 500e3, 0.1, 3578125.0, 16.0, 1.26, 70e3, 5.6),
}

MAC = {'D2MAC_12MHZ': 1080, 'D2MAC_7MHZ': 720}
