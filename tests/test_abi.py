"""The C-ABI library loads without a GPU and exports every symbol that
include/fluidb200.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fluidb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("fb_create", "fb_destroy", "fb_step", "fb_phase", "fb_edit", "fb_view", "fb_reduce",
                 "fb_upload", "fb_download", "fb_sample_velocity", "fb_halo_region", "fb_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from fluid_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in fluidb200.h but not exported"
    # and the binding table covers the header exactly
    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_struct_layouts_match_header():
    from fluid_b200 import _lib, edits
    assert ctypes.sizeof(_lib.Config) == 9 * 4
    assert ctypes.sizeof(_lib.Params) == 11 * 4
    assert ctypes.sizeof(_lib.SolveStats) == 2 * 4 + 32 * 4
    assert edits.EDIT_DTYPE.itemsize == 7 * 4


def test_no_gpu_calls_fail_cleanly():
    """Without a device fb_create must return an error code, not crash; with one it
    must succeed.  Either way the ABI is callable."""
    from fluid_b200 import _lib
    p = _lib.Params()
    assert _lib.lib.fb_default_params(ctypes.byref(p)) == 0
    assert abs(p.relaxation - 1.9) < 1e-6 and p.iters == 8 and abs(p.turbulence_strength - 0.02) < 1e-7
    assert _lib.lib.fb_version() >= 100
    h = ctypes.c_void_p()
    cfg = _lib.Config(8, 8, 1.0, 1.0, 0, 0, 1, 0, 0)
    st = _lib.lib.fb_create(ctypes.byref(cfg), ctypes.byref(h))
    if st == 0:
        assert _lib.lib.fb_destroy(h) == 0
    else:
        assert st in (-2, -3) and not h.value


def test_product_does_not_import_oracle():
    """The product path must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "fluid_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text, fn
                assert "fluid_oracle" not in text, fn
