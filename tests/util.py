"""Shared helpers for the test-suite: the parity metric and fixtures."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# BASELINE.json north_star: "max abs error <= 1e-3 and max rel error <= 1e-4 on
# cepstra, fp32".  Relative error is taken against max(|ref|, 1) because the
# higher cepstra cross zero (SURVEY.md §7.3 "Relative-error metric near zero").
ABS_TOL = 1e-3
REL_TOL = 1e-4


def parity_errors(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if got.size == 0:
        return 0.0, 0.0
    d = np.abs(got - ref)
    return float(d.max()), float((d / np.maximum(np.abs(ref), 1.0)).max())


def assert_parity(got, ref, abs_tol=ABS_TOL, rel_tol=REL_TOL, what="", truth=None):
    """``got`` must match the f32 oracle ``ref`` within the tolerance.  With ``truth`` (the f64 oracle)
    an element may instead be no further from the truth than twice the f32 oracle itself: a band
    that sits at the fp32 FFT noise floor (e.g. the DC-only filter of an 80-band bank after
    pre-emphasis, 65 dB under the spectrum) is not determined to 1e-4 by ANY f32 evaluation order —
    the f32 oracle is off the truth by 1.4e-3 there — so "equal to the f32 oracle" is not a
    meaningful bar for it, "as accurate as the f32 oracle" is."""
    assert np.isfinite(np.asarray(got)).all(), f"{what}: non-finite output"
    a, r = parity_errors(got, ref)
    if truth is None:
        assert a <= abs_tol and r <= rel_tol, f"{what}: max abs {a:.3e} (tol {abs_tol}), max rel {r:.3e} (tol {rel_tol})"
        return a, r
    g, f32, f64 = (np.asarray(x, np.float64) for x in (got, ref, truth))
    d = np.abs(g - f32)
    ok = (d <= abs_tol) & (d / np.maximum(np.abs(f32), 1.0) <= rel_tol)
    ok |= np.abs(g - f64) <= 2.0 * np.abs(f32 - f64) + 1e-6
    assert ok.all(), (f"{what}: {int((~ok).sum())} elements off both the f32 oracle (max abs {a:.3e}, max rel {r:.3e}) "
                      f"and the f64 truth")
    return a, r


def golden():
    return np.load(os.path.join(GOLDEN, "mfcc_golden.npz"))
