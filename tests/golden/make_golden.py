"""Generates the golden fixtures in this directory from the CPU oracle.

The reference (pure Go) cannot run in this environment and ships no golden
vectors, so these are outputs of oracle/fluid_oracle.c -- the restatement that
passes the reference's own test-suite (tests/test_reference_suite.py) -- on the
three presets of main/main.go at small sizes.  Re-run with
    python tests/golden/make_golden.py
Files: <case>.npz with U,V,M,p at the listed steps (keys like U_10).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from fluid_b200 import presets  # noqa: E402

CASES = {
    # name: (preset factory, width, height, snapshot steps, solver)
    "jet_130x66": (lambda: presets.jet(130, 66), (1, 2, 10, 100), oracle.SOLVER_EXACT),
    "cavity_96x96_bfecc": (lambda: presets.cavity(96, 96), (1, 10, 100), oracle.SOLVER_EXACT),
    "karman_160x80_bfecc_conf": (lambda: presets.karman(160, 80), (1, 10, 100), oracle.SOLVER_EXACT),
    "jet_300x251_default": (lambda: presets.jet(), (100,), oracle.SOLVER_EXACT),
    "karman_160x80_redblack": (lambda: presets.karman(160, 80), (1, 10, 100), oracle.SOLVER_REDBLACK),
}


def run_case(name):
    make, steps, solver = CASES[name]
    p = make()
    f = oracle.New(p.density, p.width, p.height, p.h, solver=solver)
    f.edit(p.init)
    for k, v in p.params.items():
        setattr(f, {"use_bfecc": "UseBFECC", "confinement": "Confinement"}[k], v)
    out, done = {}, 0
    for s in steps:
        f.step(p.dt, s - done, p.per_step)
        done = s
        for fld in ("U", "V", "M", "p"):
            out[f"{fld}_{s}"] = f.get(fld)
    return out


def main():
    for name in CASES:
        out = run_case(name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
