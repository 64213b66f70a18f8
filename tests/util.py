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


def assert_parity(got, ref, abs_tol=ABS_TOL, rel_tol=REL_TOL, what=""):
    assert np.isfinite(np.asarray(got)).all(), f"{what}: non-finite output"
    a, r = parity_errors(got, ref)
    assert a <= abs_tol and r <= rel_tol, f"{what}: max abs {a:.3e} (tol {abs_tol}), max rel {r:.3e} (tol {rel_tol})"
    return a, r


def golden():
    return np.load(os.path.join(GOLDEN, "mfcc_golden.npz"))
