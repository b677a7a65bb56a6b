"""Shared helpers of the test-suite."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerance for qfrc_inverse: 1e-9 relative / 1e-12 absolute
RTOL, ATOL = 1e-9, 1e-12


def golden(name):
    """(path of the gzip'd MJB, dict of reference outputs) of a committed fixture."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return os.path.join(GOLDEN, name + ".mjb.gz"), {k: z[k] for k in z.files}


def qfrc_violations(got, ref, rtol=RTOL, atol=ATOL):
    """Entries outside |got - ref| <= atol + rtol*|ref|, and the worst ratio to that bound."""
    d = np.abs(got - ref)
    tol = atol + rtol * np.abs(ref)
    return int((d > tol).sum()), float((d / tol).max()) if d.size else 0.0


def qfrc_violations_scaled(got, ref, rtol=RTOL, atol=ATOL):
    """Same bound with the relative part taken against the largest force of the STATE: a
    generalized force that is the sum of contact terms of magnitude F carries rounding of order
    eps*F in every component, also in those that cancel to ~0 (the CPU engine's own summation
    order has the same property)."""
    d = np.abs(got - ref)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    tol = atol + rtol * np.maximum(np.abs(ref), 1e-3 * scale)
    return int((d > tol).sum()), float((d / tol).max()) if d.size else 0.0


def ref_available():
    from oracle import reflib
    return reflib.available()
