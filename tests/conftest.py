import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def go_toolchain():
    """`go version` if a Go toolchain is on PATH, else None.  The reference is pure Go: where Go exists the goldens
    are re-checked against the reference itself (tests/golden/verify_with_go.sh); here it does not."""
    import shutil
    import subprocess
    exe = shutil.which("go")
    if not exe:
        return None
    try:
        return subprocess.run([exe, "version"], capture_output=True, text=True, timeout=30).stdout.strip() or None
    except Exception:
        return None


def pytest_report_header(config):
    return f"go toolchain: {go_toolchain() or 'absent (goldens pinned by the oracle only: parity unpinned by the reference)'}"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def _make_oracle(density, width, height, h, **kw):
    import oracle
    return oracle.New(density, width, height, h, **kw)


def _make_gpu(density, width, height, h, **kw):
    import fluid_b200
    kw.setdefault("compat", True)
    return fluid_b200.New(density, width, height, h, **kw)


@pytest.fixture(params=["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def impl(request):
    """Factory with the signature of fluid.New.  'oracle' = CPU restatement,
    'gpu' = fluid_b200 in white-box (compat) mode so that f.U[i, j] = x works
    like the reference's in-package tests."""
    f = _make_oracle if request.param == "oracle" else _make_gpu
    f.kind = request.param
    return f
