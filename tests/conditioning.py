"""Sample-wise conditioning of a decoder, measured on the float64 oracle (test infrastructure).

The non-linear decoders divide by a demodulated amplitude: the FM discriminator by |I - jQ|^2 (secam.py:143-148), NIIR by
the envelope and by |(sin, cos)| (niir.py:112,132-134).  Where that amplitude passes close to zero (the first samples of
a line, isolated nulls) the reference's own output is ill-conditioned: any float32 implementation carries a noise floor
of a few ulp of full scale through its filter recursions, and at those samples that floor is amplified by orders of
magnitude more than anywhere else.  Instead of a hard-coded list of exceptions the parity tests measure it: the float64
oracle is re-run on its input plus white noise of 16 ulp(1.0f) (~1e-6 absolute, the floor of a cascade of ~10 float32
second-order sections), and a sample may miss the 1e-4 bound only by what the reference itself moves under that noise.
"""
import numpy as np

FP32_FLOOR = 16.0 * 2.0 ** -24


def sensitivity(decode, comp_in, out_ref, trials=3, seed=1234):
    """max over `trials` of |decode(comp_in + noise) - out_ref|, noise uniform in +-FP32_FLOOR."""
    rng = np.random.default_rng(seed)
    s = np.zeros_like(out_ref)
    for _ in range(trials):
        d = decode(comp_in + rng.uniform(-FP32_FLOOR, FP32_FLOOR, comp_in.shape))
        s = np.maximum(s, np.abs(d - out_ref))
    return s


def fp32_bound(decode, comp_in, out_ref, tol, factor=4.0, max_relaxed=1e-3):
    """Per-sample bound of the float32 build: `tol`, except where the float64 reference itself moves by more than
    tol / factor under the float32 noise floor — there factor x that movement.  At most `max_relaxed` of the samples
    may be relaxed (well-conditioned decoders: none)."""
    s = sensitivity(decode, comp_in, out_ref)
    bound = np.maximum(tol, factor * s)
    relaxed = float((bound > tol).mean())
    assert relaxed <= max_relaxed, 'decoder ill-conditioned on %.3f %% of the samples' % (100 * relaxed)
    return bound
