"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PARAM_ATTR = {"use_bfecc": "UseBFECC", "confinement": "Confinement"}
FIELDS = ("U", "V", "M", "p")


def apply_preset(f, preset):
    f.edit(preset.init)
    for k, v in preset.params.items():
        setattr(f, PARAM_ATTR[k], v)


def diff_report(name, got, want):
    """max-abs and relative-L2 difference plus the count of bit mismatches."""
    got = np.asarray(got, dtype=np.float32)
    want = np.asarray(want, dtype=np.float32)
    bits = int(np.count_nonzero(got.view(np.uint32) != want.view(np.uint32)))
    # +0 / -0 compare equal numerically; count only numeric mismatches as errors
    neq = int(np.count_nonzero(~((got == want) | (np.isnan(got) & np.isnan(want)))))
    d = got.astype(np.float64) - want.astype(np.float64)
    max_abs = float(np.max(np.abs(d))) if d.size else 0.0
    nrm = float(np.sqrt(np.sum(want.astype(np.float64) ** 2)))
    rel_l2 = float(np.sqrt(np.sum(d ** 2)) / nrm) if nrm > 0 else float(np.sqrt(np.sum(d ** 2)))
    return {"field": name, "bit_mismatch": bits, "numeric_mismatch": neq, "max_abs": max_abs, "rel_l2": rel_l2}


def assert_bit_exact(name, got, want):
    r = diff_report(name, got, want)
    assert r["numeric_mismatch"] == 0, r


def copy_state(dst, src, fields=("U", "V", "newU", "newV", "p", "S", "M", "newM")):
    """Clone the full solver state of `src` (any impl) into `dst` (any impl)."""
    for name in fields:
        dst.set(name, src.get(name))
    for attr in ("Confinement", "ViscosityDiffusion", "PressureDamping", "TurbulenceStrength", "SmokeAdvection",
                 "UseBFECC"):
        setattr(dst, attr, getattr(src, attr))
