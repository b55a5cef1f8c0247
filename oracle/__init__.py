"""CPU oracle: ctypes wrapper of oracle/fluid_oracle.c (a restatement of the
reference's pkg/fluid in C with Go/amd64 float32 semantics).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from fluid_b200/.
PARITY UNPINNED by reference goldens (the reference has none and cannot be run
here); pinned by the reference tests' own assertions, see
tests/test_oracle_reference_suite.py.

``OracleFluid`` exposes the same names as ``fluid_b200.Fluid`` (i.e. the Go API
of pkg/fluid) so parity tests drive both with one script.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfluid_oracle.so")

EDIT_DTYPE = np.dtype(
    [("op", "<i4"), ("i0", "<i4"), ("j0", "<i4"), ("i1", "<i4"), ("j1", "<i4"), ("a", "<f4"), ("b", "<f4")]
)


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, n) for n in ("fluid_oracle.c", "fluid_oracle.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return LIB_PATH


class _FoFluid(C.Structure):
    _fields_ = [("density", C.c_float), ("h", C.c_float),
                ("NumX", C.c_int64), ("NumY", C.c_int64), ("numCells", C.c_int64),
                ("U", C.POINTER(C.c_float)), ("V", C.POINTER(C.c_float)),
                ("newU", C.POINTER(C.c_float)), ("newV", C.POINTER(C.c_float)),
                ("p", C.POINTER(C.c_float)), ("S", C.POINTER(C.c_float)),
                ("M", C.POINTER(C.c_float)), ("newM", C.POINTER(C.c_float)),
                ("Confinement", C.c_float), ("ViscosityDiffusion", C.c_float), ("PressureDamping", C.c_float),
                ("TurbulenceStrength", C.c_float), ("SmokeAdvection", C.c_float),
                ("UseMultigrid", C.c_int), ("MultigridLevels", C.c_int), ("UseBFECC", C.c_int),
                ("Relaxation", C.c_float),
                ("threads", C.c_int), ("last_iters", C.c_int), ("last_maxdiv", C.c_float)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB_PATH)
        P = C.POINTER(_FoFluid)
        f32p = C.POINTER(C.c_float)
        sig = {
            "fo_new": (P, [C.c_float, C.c_int64, C.c_int64, C.c_float]),
            "fo_free": (None, [P]),
            "fo_set_threads": (None, [P, C.c_int]),
            "fo_simulate": (None, [P, C.c_float]),
            "fo_apply_viscosity": (None, [P, C.c_float]),
            "fo_make_incompressible": (None, [P, C.c_uint, C.c_float]),
            "fo_pressure_iteration": (C.c_float, [P, C.c_float, C.c_float]),
            "fo_handle_borders": (None, [P]),
            "fo_advect_velocity": (None, [P, C.c_float]),
            "fo_advect_smoke": (None, [P, C.c_float]),
            "fo_copy_border": (None, [P, C.c_void_p, C.c_void_p]),
            "fo_apply_vorticity_confinement": (None, [P, C.c_float]),
            "fo_add_turbulence": (None, [P, C.c_float]),
            "fo_get_adaptive_time_step": (C.c_float, [P, C.c_float]),
            "fo_advect_velocity_bfecc": (None, [P, C.c_float]),
            "fo_advect_smoke_bfecc": (None, [P, C.c_float]),
            "fo_sample_field": (C.c_float, [P, C.c_float, C.c_float, C.c_int]),
            "fo_set_solid": (C.c_int, [P, C.c_int64, C.c_int64, C.c_int]),
            "fo_is_solid": (C.c_int, [P, C.c_int64, C.c_int64]),
            "fo_set_velocity": (C.c_int, [P, C.c_int64, C.c_int64, C.c_float, C.c_float]),
            "fo_add_smoke": (C.c_int, [P, C.c_int64, C.c_int64, C.c_float]),
            "fo_reset": (None, [P]),
            "fo_apply_force": (None, [P, C.c_int64, C.c_int64, C.c_float, C.c_float]),
            "fo_apply_force_radius": (None, [P, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_int64]),
            "fo_set_circular_obstacle": (None, [P, C.c_int64, C.c_int64, C.c_int64]),
            "fo_minmax": (None, [C.c_void_p, C.c_int64, f32p, f32p]),
            "fo_vorticity": (None, [P, C.c_void_p, f32p, f32p]),
            "fo_velocity_magnitude": (None, [P, C.c_void_p, f32p, f32p]),
            "fo_max_divergence": (C.c_float, [P]),
            "fo_sample_velocity": (None, [P, C.c_float, C.c_float, f32p, f32p]),
            "fo_render": (None, [P, C.c_int, C.c_void_p]),
            "fo_advect_particles": (C.c_int64, [P, C.c_void_p, C.c_int64, C.c_float]),
            "fo_apply_edits": (C.c_int, [P, C.c_void_p, C.c_int64]),
            "fo_run": (C.c_int, [P, C.c_float, C.c_int64, C.c_void_p, C.c_int64]),
            "fo_project_redblack": (C.c_float, [P, C.c_uint, C.c_float]),
            "fo_project_redblack_q": (C.c_float, [P, C.c_uint, C.c_float]),
            "fo_project_multigrid_redblack": (None, [P, C.c_uint, C.c_float]),
            "fo_project_multigrid_redblack_q": (None, [P, C.c_uint, C.c_float]),
            "fo_project_redblack_sched": (C.c_float, [P, C.c_void_p, C.c_uint, C.c_float]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


SOLVER_EXACT, SOLVER_REDBLACK, SOLVER_REDBLACK_PRESSURE = 0, 1, 2


class ScalarField:
    def __init__(self, values, mn, mx):
        self.NumX, self.NumY = values.shape
        self.values, self.MinValue, self.MaxValue = values, mn, mx

    def Value(self, i, j):
        if i < 0 or i >= self.NumX:
            raise IndexError(f"x index ({i}) out of range, must be between 0 and {self.NumX - 1}")
        if j < 0 or j >= self.NumY:
            raise IndexError(f"y index ({j}) out of range, must be between 0 and {self.NumY - 1}")
        return float(self.values[i, j])


class VectorField:
    def __init__(self, u, v):
        self.NumX, self.NumY = u.shape
        self.valuesU, self.valuesV = u, v

    def Value(self, i, j):
        if i < 0 or i >= self.NumX or j < 0 or j >= self.NumY:
            raise IndexError("index out of range")
        return float(self.valuesU[i, j]), float(self.valuesV[i, j])


_KNOBS = ("Confinement", "ViscosityDiffusion", "PressureDamping", "TurbulenceStrength", "SmokeAdvection",
          "UseMultigrid", "MultigridLevels", "UseBFECC", "Relaxation")


class OracleFluid:
    """pkg/fluid's ``Fluid`` on the CPU.  Arrays U, V, S, M, p, newU, newV, newM are
    live numpy views of the C arrays (like the Go slices), shape [NumX, NumY]."""

    def __init__(self, density, width, height, h, threads: int | None = None, solver: int = SOLVER_EXACT):
        self._l = lib()
        self._f = self._l.fo_new(density, width, height, h)
        if not self._f:
            raise MemoryError("fo_new")
        c = self._f.contents
        self.NumX, self.NumY, self.numCells = c.NumX, c.NumY, c.numCells
        self.h, self.density = c.h, c.density
        self.Solver = solver      # SOLVER_REDBLACK: use the fast-mode restatement for the projection
        self.NumIters = 8
        for name in ("U", "V", "newU", "newV", "p", "S", "M", "newM"):
            arr = np.ctypeslib.as_array(getattr(c, name), shape=(self.numCells,)).reshape(self.NumX, self.NumY)
            object.__setattr__(self, name, arr)
        if threads is not None:
            self._l.fo_set_threads(self._f, threads)

    # knobs live in the C struct
    def __getattr__(self, name):
        if name in _KNOBS:
            v = getattr(self._f.contents, name)
            return bool(v) if name in ("UseMultigrid", "UseBFECC") else v
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in _KNOBS:
            setattr(self._f.contents, name, int(value) if name in ("UseMultigrid", "UseBFECC", "MultigridLevels") else value)
        else:
            object.__setattr__(self, name, value)

    def close(self):
        if getattr(self, "_f", None):
            for name in ("U", "V", "newU", "newV", "p", "S", "M", "newM"):
                self.__dict__.pop(name, None)
            self._l.fo_free(self._f)
            object.__setattr__(self, "_f", None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def threads(self):
        return self._f.contents.threads

    def H(self):
        return self.h

    # ---- edits
    def _chk(self, rc, i, j):
        if rc < 0:
            raise IndexError(f"invalid index: ({i},{j})")

    def SetSolid(self, i, j, value):
        self._chk(self._l.fo_set_solid(self._f, i, j, int(bool(value))), i, j)

    def IsSolid(self, i, j):
        rc = self._l.fo_is_solid(self._f, i, j)
        self._chk(rc, i, j)
        return bool(rc)

    def SetVelocity(self, i, j, u, v):
        self._chk(self._l.fo_set_velocity(self._f, i, j, u, v), i, j)

    def AddSmoke(self, i, j, s):
        self._chk(self._l.fo_add_smoke(self._f, i, j, s), i, j)

    def Reset(self):
        self._l.fo_reset(self._f)

    def ApplyForce(self, i, j, fx, fy):
        self._l.fo_apply_force(self._f, i, j, fx, fy)

    def ApplyForceRadius(self, cx, cy, fx, fy, radius):
        self._l.fo_apply_force_radius(self._f, cx, cy, fx, fy, radius)

    def SetCircularObstacle(self, cx, cy, radius):
        self._l.fo_set_circular_obstacle(self._f, cx, cy, radius)

    def edit(self, cmds):
        arr = np.ascontiguousarray(np.array(list(cmds), dtype=EDIT_DTYPE) if not isinstance(cmds, np.ndarray) else cmds)
        if len(arr) and self._l.fo_apply_edits(self._f, arr.ctypes.data, len(arr)) != 0:
            raise IndexError("edit out of range")

    def flush(self):
        pass

    # ---- hot path
    def _project_rb(self, iters, dt):
        if self.UseMultigrid and self.MultigridLevels > 1:
            if self.Solver == SOLVER_REDBLACK_PRESSURE:
                return self._l.fo_project_multigrid_redblack_q(self._f, iters, dt)
            return self._l.fo_project_multigrid_redblack(self._f, iters, dt)
        fn = self._l.fo_project_redblack_q if self.Solver == SOLVER_REDBLACK_PRESSURE else self._l.fo_project_redblack
        return fn(self._f, iters, dt)

    def Simulate(self, dt):
        if self.Solver != SOLVER_EXACT:
            self._simulate_redblack(dt)
        else:
            self._l.fo_simulate(self._f, dt)

    def _simulate_redblack(self, dt):
        """Simulate (fluid.go:79-109) with the projection replaced by the red-black
        restatement -- mirrors fb_step with FB_SOLVER_REDBLACK."""
        l, f = self._l, self._f
        self.p[...] = 0
        if self.ViscosityDiffusion > 0:
            l.fo_apply_viscosity(f, dt)
        self._project_rb(self.NumIters, dt)
        if self.Confinement != 0:
            l.fo_apply_vorticity_confinement(f, dt)
        if self.TurbulenceStrength > 0:
            l.fo_add_turbulence(f, dt)
        l.fo_handle_borders(f)
        if self.UseBFECC:
            l.fo_advect_velocity_bfecc(f, dt)
            l.fo_advect_smoke_bfecc(f, dt)
        else:
            l.fo_advect_velocity(f, dt)
            l.fo_advect_smoke(f, dt)

    def step(self, dt, nsteps=1, per_step=None):
        arr = None
        if per_step is not None and len(per_step):
            arr = np.ascontiguousarray(per_step)
        if self.Solver != SOLVER_EXACT:
            for _ in range(nsteps):
                if arr is not None:
                    self.edit(arr)
                self._simulate_redblack(dt)
            return
        rc = self._l.fo_run(self._f, dt, nsteps, arr.ctypes.data if arr is not None else None,
                            len(arr) if arr is not None else 0)
        if rc != 0:
            raise IndexError("per-step edit out of range")

    def makeIncompressible(self, numIters, dt):
        if self.Solver != SOLVER_EXACT:
            self._project_rb(numIters, dt)
        else:
            self._l.fo_make_incompressible(self._f, numIters, dt)

    def redblackIteration(self, relaxation, dt):
        """NOT in the reference: one red + one black half sweep at `relaxation` (cp = density*h/dt)."""
        om = np.array([relaxation, relaxation], dtype=np.float32)
        return self._l.fo_project_redblack_sched(self._f, om.ctypes.data, 1, dt)

    def pressureIteration(self, relaxation, cp):
        return self._l.fo_pressure_iteration(self._f, relaxation, cp)

    def advectVelocity(self, dt):
        self._l.fo_advect_velocity(self._f, dt)

    def advectSmoke(self, dt):
        self._l.fo_advect_smoke(self._f, dt)

    def handleBorders(self):
        self._l.fo_handle_borders(self._f)

    def copyBorder(self, dst, src):
        self._l.fo_copy_border(self._f, dst.ctypes.data, src.ctypes.data)

    def applyVorticityConfinement(self, dt):
        self._l.fo_apply_vorticity_confinement(self._f, dt)

    def addTurbulence(self, dt):
        self._l.fo_add_turbulence(self._f, dt)

    def applyViscosity(self, dt):
        self._l.fo_apply_viscosity(self._f, dt)

    def advectVelocityBFECC(self, dt):
        self._l.fo_advect_velocity_bfecc(self._f, dt)

    def advectSmokeBFECC(self, dt):
        self._l.fo_advect_smoke_bfecc(self._f, dt)

    def clearPressure(self):
        self.p[...] = 0

    def project(self, numIters, dt):
        """fill(p, 0) + makeIncompressible as Simulate runs them (fluid.go:83, 90)."""
        self.clearPressure()
        self.makeIncompressible(numIters, dt)

    def solve_stats(self):
        c = self._f.contents
        return {"sweeps_run": c.last_iters, "last_max_div": c.last_maxdiv}

    def GetAdaptiveTimeStep(self, basedt):
        return self._l.fo_get_adaptive_time_step(self._f, basedt)

    # ---- field access (same names as fluid_b200.Fluid)
    def get(self, name):
        return np.array(getattr(self, name), copy=True)

    def set(self, name, values):
        getattr(self, name)[...] = np.asarray(values, dtype=np.float32).reshape(self.NumX, self.NumY)

    # ---- views
    def _minmax(self, a):
        mn, mx = C.c_float(), C.c_float()
        self._l.fo_minmax(a.ctypes.data, a.size, C.byref(mn), C.byref(mx))
        return mn.value, mx.value

    def Smoke(self):
        return ScalarField(self.M, *self._minmax(self.M))

    def Pressure(self):
        return ScalarField(self.p, *self._minmax(self.p))

    def Velocity(self):
        return VectorField(self.U, self.V)

    def Vorticity(self):
        out = np.zeros((self.NumX, self.NumY), dtype=np.float32)
        mn, mx = C.c_float(), C.c_float()
        self._l.fo_vorticity(self._f, out.ctypes.data, C.byref(mn), C.byref(mx))
        return ScalarField(out, mn.value, mx.value)

    def VelocityMagnitude(self):
        out = np.zeros((self.NumX, self.NumY), dtype=np.float32)
        mn, mx = C.c_float(), C.c_float()
        self._l.fo_velocity_magnitude(self._f, out.ctypes.data, C.byref(mn), C.byref(mx))
        return ScalarField(out, mn.value, mx.value)

    def MaxDivergence(self):
        return self._l.fo_max_divergence(self._f)

    def SampleVelocity(self, x, y):
        u, v = C.c_float(), C.c_float()
        self._l.fo_sample_velocity(self._f, x, y, C.byref(u), C.byref(v))
        return u.value, v.value

    def Render(self, kind):
        """Draw's pixel pass (main/main.go:550-574, 620-652): RGBA image [NumY][NumX][4] of view `kind`."""
        out = np.zeros((self.NumY, self.NumX, 4), dtype=np.uint8)
        self._l.fo_render(self._f, int(kind), out.ctypes.data)
        return out

    def AdvectParticles(self, particles, dt):
        """advectParticles (main/main.go:512-546) on a structured array (fluid_b200.PARTICLE_DTYPE)."""
        ps = np.ascontiguousarray(particles).copy()
        n = self._l.fo_advect_particles(self._f, ps.ctypes.data, len(ps), dt)
        return ps[:n]

    def SampleVelocities(self, xy):
        xy = np.asarray(xy, dtype=np.float32).reshape(-1, 2)
        return np.array([self.SampleVelocity(float(x), float(y)) for x, y in xy], dtype=np.float32)


def New(density, width, height, h, **kw):
    return OracleFluid(density, width, height, h, **kw)
