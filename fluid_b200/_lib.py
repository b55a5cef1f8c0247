"""ctypes binding of libfluidb200.so -- the same entry points a cgo binding uses
(include/fluidb200.h).  The library is opened on the FIRST use of ``lib`` (so that the
pure-Python helpers -- presets, edit lists, slab planning -- import on a checkout
without it); there is no CPU fallback: if it is missing or does not load, that first
use raises ImportError."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FLUIDB200_LIB: load another build of the SAME library (kernel experiments, tools/variants.sh)
LIB_PATH = os.environ.get("FLUIDB200_LIB") or os.path.join(_HERE, "libfluidb200.so")

FB_OK = 0
STATUS = {0: "FB_OK", -1: "FB_ERR_INVALID", -2: "FB_ERR_CUDA", -3: "FB_ERR_NOMEM",
          -4: "FB_ERR_UNSUPPORTED", -5: "FB_ERR_HALO"}

# fb_field
U, V, NEWU, NEWV, P, S, M, NEWM = range(8)
FIELD_NAMES = {"U": U, "V": V, "newU": NEWU, "newV": NEWV, "p": P, "S": S, "M": M, "newM": NEWM}
# fb_solver
SOLVER_EXACT, SOLVER_REDBLACK, SOLVER_REDBLACK_PRESSURE = 0, 1, 2
OPT_SOLVE_STATS = 0
OPT_HALO_OVERLAP = 1
# fb_flags
FLAG_LITERAL, FLAG_EXACT_SHADOW = 1, 2
# fb_phase_id
(PHASE_MAKE_INCOMPRESSIBLE, PHASE_ADVECT_VELOCITY, PHASE_ADVECT_SMOKE, PHASE_HANDLE_BORDERS, PHASE_CONFINEMENT,
 PHASE_TURBULENCE, PHASE_ADVECT_VELOCITY_BFECC, PHASE_ADVECT_SMOKE_BFECC, PHASE_VISCOSITY,
 PHASE_CLEAR_PRESSURE, PHASE_PROJECT) = range(11)
# fb_view_kind / fb_reduce_kind
VIEW_SMOKE, VIEW_PRESSURE, VIEW_VELOCITY_MAGNITUDE, VIEW_VORTICITY = range(4)
REDUCE_MAX_DIVERGENCE, REDUCE_MAX_ABS_VELOCITY = range(2)
PROF_PHASES = ("edits", "clear_pressure", "viscosity", "project", "confinement", "turbulence", "borders",
               "advect_velocity", "advect_smoke",
               # single kernels, one event pair per launch (inside the phase pairs above)
               "k_pressure_solve", "k_advect_velocity_full", "k_bfecc_velocity_correct", "k_advect_smoke_full",
               "k_bfecc_smoke_correct", "k_confine_turbulence",
               # slabs: the halo exchange in front of a step, or (overlapped) the wait for it at the end of the step
               "halo")
PROF_KERNELS = PROF_PHASES[9:15]


class Config(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("density", C.c_float), ("h", C.c_float),
                ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("ghost", C.c_int32),
                ("flags", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("relaxation", C.c_float), ("confinement", C.c_float), ("viscosity_diffusion", C.c_float),
                ("pressure_damping", C.c_float), ("turbulence_strength", C.c_float),
                ("smoke_advection", C.c_float), ("use_multigrid", C.c_int32), ("multigrid_levels", C.c_int32),
                ("use_bfecc", C.c_int32), ("solver", C.c_int32), ("iters", C.c_int32)]


class SolveStats(C.Structure):
    _fields_ = [("sweeps_run", C.c_int32), ("rolled_back", C.c_int32), ("max_div", C.c_float * 32)]


# every symbol include/fluidb200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "fb_create": (C.c_int, [C.POINTER(Config), C.POINTER(_H)]),
    "fb_destroy": (C.c_int, [_H]),
    "fb_last_error": (C.c_char_p, [_H]),
    "fb_default_params": (C.c_int, [C.POINTER(Params)]),
    "fb_dims": (C.c_int, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fb_step": (C.c_int, [_H, C.POINTER(Params), C.c_float, C.c_int32, C.c_void_p, C.c_size_t]),
    "fb_step_local": (C.c_int, [_H, C.POINTER(Params), C.c_float, C.c_int32, C.c_void_p, C.c_size_t]),
    "fb_phase": (C.c_int, [_H, C.c_int32, C.POINTER(Params), C.c_float, C.c_uint32]),
    "fb_get_solve_stats": (C.c_int, [_H, C.POINTER(SolveStats)]),
    "fb_edit": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "fb_apply_force_radius": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32]),
    "fb_upload": (C.c_int, [_H, C.c_int32, C.c_void_p]),
    "fb_download": (C.c_int, [_H, C.c_int32, C.c_void_p]),
    "fb_host_mirror": (C.c_int, [_H, C.c_int32, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_size_t)]),
    "fb_view": (C.c_int, [_H, C.c_int32, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "fb_halo_export": (C.c_int, [_H, C.c_int32, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)]),
    "fb_halo_connect": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_uint64]),
    "fb_halo_connect_local": (C.c_int, [_H, C.c_int32, _H]),
    "fb_halo_post": (C.c_int, [_H]),
    "fb_halo_pull": (C.c_int, [_H]),
    "fb_halo_exchange": (C.c_int, [_H]),
    "fb_view_begin": (C.c_int, [_H, C.c_int32, C.c_void_p]),
    "fb_view_end": (C.c_int, [_H, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "fb_view_u8_begin": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "fb_render_begin": (C.c_int, [_H, C.c_int32, C.c_void_p, C.POINTER(C.c_float)]),
    "fb_render_end": (C.c_int, [_H, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "fb_render": (C.c_int, [_H, C.c_int32, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "fb_advect_particles": (C.c_int, [_H, C.c_void_p, C.c_size_t, C.c_float, C.POINTER(C.c_size_t)]),
    "fb_reduce": (C.c_int, [_H, C.c_int32, C.POINTER(C.c_float)]),
    "fb_sample_velocity": (C.c_int, [_H, C.c_size_t, C.c_void_p, C.c_void_p]),
    "fb_halo_region": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_size_t)]),
    "fb_check_halo": (C.c_int, [_H]),
    "fb_ghost_lines": (C.c_int, [_H, C.POINTER(C.c_int32)]),
    "fb_stream": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "fb_synchronize": (C.c_int, [_H]),
    "fb_timer_start": (C.c_int, [_H]),
    "fb_timer_stop": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "fb_profile_enable": (C.c_int, [_H, C.c_int32]),
    "fb_profile_read": (C.c_int, [_H, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "fb_set_option": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "fb_launch_count": (C.c_int, [_H, C.POINTER(C.c_uint64)]),
    "fb_selftest_fastmath": (C.c_int, [C.c_int32, C.c_uint64, C.c_uint32, C.c_int32, C.POINTER(C.c_uint64)]),
    "fb_version": (C.c_int, []),
}


def load(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). fluid_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


class _LazyLib:
    """``lib.fb_xxx`` opens the library on first attribute access and then gets out of the way."""
    _cdll = None

    def __getattr__(self, name):
        if _LazyLib._cdll is None:
            _LazyLib._cdll = load()
        fn = getattr(_LazyLib._cdll, name)
        setattr(self, name, fn)
        return fn


lib = _LazyLib()


class FluidError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


def check(handle, status: int):
    if status != FB_OK:
        msg = lib.fb_last_error(handle) if handle else b"(no handle)"
        raise FluidError(status, (msg or b"").decode("utf-8", "replace"))
