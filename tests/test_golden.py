"""Golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py
from the oracle): the oracle must keep reproducing them bit for bit on this host
(guards against libm / compiler drift), and the CUDA path must reproduce them
bit for bit on the GPU box (integer-exact float32 parity, no tolerance)."""
import os

import numpy as np
import pytest

from common import FIELDS, GOLDEN, apply_preset, assert_bit_exact
from golden.make_golden import CASES


def _run(f, name):
    make, steps, _solver = CASES[name]
    p = make()
    apply_preset(f, p)
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    done = 0
    for s in steps:
        f.step(p.dt, s - done, p.per_step)
        done = s
        for fld in FIELDS:
            assert_bit_exact(f"{name}:{fld}@{s}", f.get(fld), gold[f"{fld}_{s}"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    import oracle
    make, _steps, solver = CASES[name]
    p = make()
    _run(oracle.New(p.density, p.width, p.height, p.h, solver=solver), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_golden(name):
    import fluid_b200
    make, _steps, solver = CASES[name]
    p = make()
    with fluid_b200.New(p.density, p.width, p.height, p.h, solver=solver) as f:
        _run(f, name)


def test_goldens_against_the_go_reference():
    """Where a Go toolchain and the reference checkout exist, the committed goldens must equal the output of the
    REAL reference (go/cmd/dump driven by tests/golden/verify_with_go.sh), bit for bit.  Skipped -- and parity stays
    "unpinned" -- where they do not (this image, the GPU boxes)."""
    import subprocess
    from conftest import go_toolchain
    ref = os.environ.get("FLUID_REFERENCE", "/root/reference")
    if not go_toolchain():
        pytest.skip("no Go toolchain: goldens pinned by the oracle only")
    if not os.path.isdir(os.path.join(ref, "pkg", "fluid")):
        pytest.skip("no reference checkout at " + ref)
    r = subprocess.run([os.path.join(GOLDEN, "verify_with_go.sh"), ref], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
