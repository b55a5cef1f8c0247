"""Host-side mirror of the reference's Go API (package ``fluid``).

Every exported name of pkg/fluid keeps its name, argument meaning and error
behaviour: ``New`` (fluid.go:42), ``Simulate`` (fluid.go:79), the edits of
walls.go, the views of pressure.go / smoke.go / velocity.go / fluid.go:799-891.
Where Go panics (walls.go:6-11) this raises ``IndexError``; where Go returns an
``error`` (scalar_field.go:14-19) this raises ``IndexError`` too.

All arithmetic happens in libfluidb200.so through the C ABI; this file only
batches edits, owns the parameter struct and moves arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from . import edits as E

# package var `Relaxation` (fluid.go:7-9)
Relaxation = 1.9

_MAXF = float(np.finfo(np.float32).max)


# main/main.go:139-144 (`type Particle`), the layout of fb_particle
PARTICLE_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("pad", "u1"),
                           ("age", "<f4"), ("max_age", "<f4")])


class ScalarField:
    """scalar_field.go:7-22."""

    def __init__(self, values: np.ndarray, min_value: float, max_value: float):
        self.NumX, self.NumY = values.shape
        self.values = values
        self.MinValue = min_value
        self.MaxValue = max_value

    def Value(self, i: int, j: int) -> float:
        if i < 0 or i >= self.NumX:
            raise IndexError(f"x index ({i}) out of range, must be between 0 and {self.NumX - 1}")
        if j < 0 or j >= self.NumY:
            raise IndexError(f"y index ({j}) out of range, must be between 0 and {self.NumY - 1}")
        return float(self.values[i, j])


class VectorField:
    """vector_field.go:5-19."""

    def __init__(self, u: np.ndarray, v: np.ndarray):
        self.NumX, self.NumY = u.shape
        self.valuesU, self.valuesV = u, v

    def Value(self, i: int, j: int):
        if i < 0 or i >= self.NumX:
            raise IndexError(f"x index out of range, must be between 0 and {self.NumX - 1}")
        if j < 0 or j >= self.NumY:
            raise IndexError(f"y index out of range, must be between 0 and {self.NumY - 1}")
        return float(self.valuesU[i, j]), float(self.valuesV[i, j])


class Fluid:
    """``type Fluid`` (fluid.go:11-40) on one B200 (or one slab of a multi-GPU grid)."""

    def __init__(self, density: float, width: int, height: int, h: float, *, device: int = 0,
                 solver: int = L.SOLVER_EXACT, rank: int = 0, nranks: int = 1, ghost: int = 0,
                 compat: bool = False, literal: bool = False, exact_shadow: bool | None = None):
        self._h = C.c_void_p()
        if exact_shadow is None:
            exact_shadow = compat     # white-box callers may rewrite S directly
        flags = (L.FLAG_LITERAL if literal else 0) | (L.FLAG_EXACT_SHADOW if exact_shadow else 0)
        cfg = L.Config(width, height, density, h, device, rank, nranks, ghost, flags)
        st = L.lib.fb_create(C.byref(cfg), C.byref(self._h))
        if st != L.FB_OK:
            self._h = C.c_void_p()
            raise L.FluidError(st, "fb_create failed (is a CUDA device visible?)")
        self.density = float(np.float32(density))
        self.h = float(np.float32(h))
        self.NumX, self.NumY = width + 2, height + 2
        self.numCells = self.NumX * self.NumY
        p = L.Params()
        L.check(self._h, L.lib.fb_default_params(C.byref(p)))
        # exported knobs, defaults of fluid.go:59-66
        self.Confinement = p.confinement
        self.ViscosityDiffusion = p.viscosity_diffusion
        self.PressureDamping = p.pressure_damping
        self.TurbulenceStrength = p.turbulence_strength
        self.SmokeAdvection = p.smoke_advection
        self.UseMultigrid = bool(p.use_multigrid)
        self.MultigridLevels = p.multigrid_levels
        self.UseBFECC = bool(p.use_bfecc)
        # not in the reference: which ordering the projection uses, and numIters
        self.Solver = solver
        self.NumIters = 8          # fluid.go:81
        self.Relaxation = None     # None -> module var `Relaxation` at call time
        self.compat = compat
        self._pending: list = []   # edits queued since the last flush
        self._mirrors: dict = {}
        self._s_valid = False
        lo, hi = C.c_int64(), C.c_int64()
        L.check(self._h, L.lib.fb_dims(self._h, None, None, C.byref(lo), C.byref(hi)))
        self.i_lo, self.i_hi = lo.value, hi.value
        self.rank, self.nranks = rank, nranks

    # ---- lifecycle -------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            L.lib.fb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- parameters ----------------------------------------------------------------
    def params(self) -> L.Params:
        relax = Relaxation if self.Relaxation is None else self.Relaxation
        return L.Params(relax, self.Confinement, self.ViscosityDiffusion, self.PressureDamping,
                        self.TurbulenceStrength, self.SmokeAdvection, int(self.UseMultigrid), self.MultigridLevels,
                        int(self.UseBFECC), self.Solver, self.NumIters)

    def H(self) -> float:   # fluid.go:71
        return self.h

    # ---- edits (walls.go) ------------------------------------------------------------
    def _check_index(self, i: int, j: int):
        if i < 0 or i >= self.NumX:
            raise IndexError(f"invalid x-index: {i}")      # walls.go:7
        if j < 0 or j >= self.NumY:
            raise IndexError(f"invalid y-index: {j}")      # walls.go:10

    def SetSolid(self, i: int, j: int, value: bool):       # walls.go:5
        self._check_index(i, j)
        self._pending.append(E.set_solid(i, j, value))
        self._s_valid = False

    def IsSolid(self, i: int, j: int) -> bool:             # walls.go:51
        self._check_index(i, j)
        if self.compat:
            self.flush()
            return bool(self._mirror(L.S)[i, j] == 0.0)
        return bool(self._solid_mirror()[i, j] == 0.0)

    def SetVelocity(self, i: int, j: int, u: float, v: float):   # walls.go:62
        self._check_index(i, j)
        self._pending.append(E.set_velocity(i, j, u, v))

    def AddSmoke(self, i: int, j: int, smoke: float):      # walls.go:74
        self._check_index(i, j)
        self._pending.append(E.add_smoke(i, j, smoke))

    def Reset(self):                                       # walls.go:85
        self._pending.append(E.reset())

    def ApplyForce(self, i: int, j: int, fx: float, fy: float):   # fluid.go:761
        self._pending.append(E.apply_force(i, j, fx, fy))

    def ApplyForceRadius(self, cx: int, cy: int, fx: float, fy: float, radius: int):   # fluid.go:774
        self.flush()
        if self.compat:
            self._upload_compat()
        L.check(self._h, L.lib.fb_apply_force_radius(self._h, cx, cy, fx, fy, radius))
        if self.compat:
            self._download_compat()

    def SetCircularObstacle(self, cx: int, cy: int, radius: int):   # fluid.go:894
        self._pending.append(E.circle_obstacle(cx, cy, radius))
        self._s_valid = False

    def edit(self, cmds):
        """Apply a command list (fluid_b200.edits) after anything already queued."""
        self.flush()
        arr = E.pack(cmds)
        if len(arr):
            if self.compat:
                self._upload_compat()
            L.check(self._h, L.lib.fb_edit(self._h, arr.ctypes.data, len(arr)))
            if self.compat:
                self._download_compat()
        self._s_valid = False

    def flush(self):
        if self._pending:
            arr = E.pack(self._pending)
            self._pending = []
            if self.compat:
                # white-box mode: the host mirrors are authoritative between calls
                self._upload_compat()
            L.check(self._h, L.lib.fb_edit(self._h, arr.ctypes.data, len(arr)))
            if self.compat:
                self._download_compat()

    # ---- the hot path ----------------------------------------------------------------------
    def Simulate(self, dt: float):                         # fluid.go:79
        self.step(dt, 1)

    def step(self, dt: float, nsteps: int = 1, per_step=None):
        """``Simulate`` x nsteps; ``per_step`` edits are replayed before every step
        (what main/main.go:233-241 and 474-486 do each frame)."""
        self.flush()
        if self.compat:
            self._upload_compat()
        p = self.params()
        if per_step is not None and len(per_step):
            arr = E.pack(per_step)
            L.check(self._h, L.lib.fb_step(self._h, C.byref(p), dt, nsteps, arr.ctypes.data, len(arr)))
        else:
            L.check(self._h, L.lib.fb_step(self._h, C.byref(p), dt, nsteps, None, 0))
        if self.compat:
            self._download_compat()

    def _phase(self, phase: int, dt: float = 0.0, iters: int = 0):
        self.flush()
        if self.compat:
            self._upload_compat()
        p = self.params()
        L.check(self._h, L.lib.fb_phase(self._h, phase, C.byref(p), dt, iters))
        if self.compat:
            self._download_compat()

    # white-box phases the reference's in-package tests call directly
    def makeIncompressible(self, numIters: int, dt: float):   # fluid.go:144
        self._phase(L.PHASE_MAKE_INCOMPRESSIBLE, dt, numIters)

    def advectVelocity(self, dt: float):                      # fluid.go:291
        self._phase(L.PHASE_ADVECT_VELOCITY, dt)

    def advectSmoke(self, dt: float):                         # fluid.go:400
        self._phase(L.PHASE_ADVECT_SMOKE, dt)

    def handleBorders(self):                                  # fluid.go:236
        self._phase(L.PHASE_HANDLE_BORDERS)

    def applyVorticityConfinement(self, dt: float):           # fluid.go:449
        self._phase(L.PHASE_CONFINEMENT, dt)

    def addTurbulence(self, dt: float):                       # fluid.go:496
        self._phase(L.PHASE_TURBULENCE, dt)

    def applyViscosity(self, dt: float):                      # fluid.go:112
        self._phase(L.PHASE_VISCOSITY, dt)

    def advectVelocityBFECC(self, dt: float):                 # fluid.go:911
        self._phase(L.PHASE_ADVECT_VELOCITY_BFECC, dt)

    def advectSmokeBFECC(self, dt: float):                    # fluid.go:997
        self._phase(L.PHASE_ADVECT_SMOKE_BFECC, dt)

    def clearPressure(self):                                  # fluid.go:83
        self._phase(L.PHASE_CLEAR_PRESSURE)

    def project(self, numIters: int, dt: float):              # fluid.go:83 + 90, as Simulate runs them
        self._phase(L.PHASE_PROJECT, dt, numIters)

    def solve_stats(self):
        st = L.SolveStats()
        L.check(self._h, L.lib.fb_get_solve_stats(self._h, C.byref(st)))
        return {"sweeps_run": st.sweeps_run, "rolled_back": bool(st.rolled_back),
                "max_div": [float(x) for x in st.max_div[:max(st.sweeps_run, 0)]]}

    # ---- field transfer ---------------------------------------------------------------------
    def _mirror(self, field: int) -> np.ndarray:
        if field not in self._mirrors:
            ptr, n = C.POINTER(C.c_float)(), C.c_size_t()
            L.check(self._h, L.lib.fb_host_mirror(self._h, field, C.byref(ptr), C.byref(n)))
            self._mirrors[field] = np.ctypeslib.as_array(ptr, shape=(n.value,)).reshape(self.NumX, self.NumY)
        return self._mirrors[field]

    def get(self, name: str) -> np.ndarray:
        """Download a field into a fresh dense [NumX, NumY] array (owned lines only
        are filled when this handle is one slab of several)."""
        self.flush()
        out = np.zeros((self.NumX, self.NumY), dtype=np.float32)
        L.check(self._h, L.lib.fb_download(self._h, L.FIELD_NAMES[name], out.ctypes.data))
        return out

    def set(self, name: str, values):
        """Upload a dense [NumX, NumY] array (the white-box `f.U[...] = x` of the
        reference's tests and benchmarks, fluid_bench_test.go:9-13)."""
        self.flush()
        arr = np.ascontiguousarray(np.asarray(values, dtype=np.float32).reshape(self.NumX, self.NumY))
        L.check(self._h, L.lib.fb_upload(self._h, L.FIELD_NAMES[name], arr.ctypes.data))
        if name == "S":
            self._s_valid = False

    def _solid_mirror(self) -> np.ndarray:
        m = self._mirror(L.S)
        if not self._s_valid or self._pending:
            self.flush()
            L.check(self._h, L.lib.fb_download(self._h, L.S, m.ctypes.data))
            self._s_valid = True
        return m

    def _download_into_mirror(self, field: int) -> np.ndarray:
        self.flush()
        m = self._mirror(field)
        L.check(self._h, L.lib.fb_download(self._h, field, m.ctypes.data))
        return m

    # exported slices U, V, S, M (fluid.go:17-22): pinned host mirrors, refreshed on access
    @property
    def U(self) -> np.ndarray:
        return self._compat_mirror(L.U) if self.compat else self._download_into_mirror(L.U)

    @property
    def V(self) -> np.ndarray:
        return self._compat_mirror(L.V) if self.compat else self._download_into_mirror(L.V)

    @property
    def M(self) -> np.ndarray:
        return self._compat_mirror(L.M) if self.compat else self._download_into_mirror(L.M)

    @property
    def S(self) -> np.ndarray:
        return self._compat_mirror(L.S) if self.compat else self._solid_mirror()

    @property
    def p(self) -> np.ndarray:
        return self._compat_mirror(L.P) if self.compat else self._download_into_mirror(L.P)

    def _compat_mirror(self, field: int) -> np.ndarray:
        self.flush()
        return self._mirror(field)

    def _upload_compat(self):
        for fld in (L.U, L.V, L.M, L.S, L.P):
            L.check(self._h, L.lib.fb_upload(self._h, fld, self._mirror(fld).ctypes.data))

    def _download_compat(self):
        for fld in (L.U, L.V, L.M, L.S, L.P):
            L.check(self._h, L.lib.fb_download(self._h, fld, self._mirror(fld).ctypes.data))
        self._s_valid = True

    # ---- views (Q-14) -----------------------------------------------------------------------
    def _view(self, kind: int) -> ScalarField:
        self.flush()
        if self.compat:
            self._upload_compat()
        out = np.zeros((self.NumX, self.NumY), dtype=np.float32)
        mn, mx = C.c_float(), C.c_float()
        L.check(self._h, L.lib.fb_view(self._h, kind, out.ctypes.data, C.byref(mn), C.byref(mx)))
        return ScalarField(out, mn.value, mx.value)

    def view_begin(self, kind: int, out) -> None:
        """Pipelined view (fb_view_begin): queue the view behind the submitted work and return.
        ``out`` is a pinned float32 host array of the dense global shape (or its address); the
        transfer overlaps whatever is queued next.  Pair with view_end()."""
        self.flush()
        addr = out if isinstance(out, int) else out.ctypes.data
        L.check(self._h, L.lib.fb_view_begin(self._h, kind, addr))

    def view_u8_begin(self, kind: int, stride: int, out):
        """Pipelined decimated 8-bit view (fb_view_u8_begin): every ``stride``-th cell of every ``stride``-th line,
        quantised on the device against the full field's min / max; ``out`` is a pinned uint8 host array (or its
        address) of at least ceil(lines/stride) * ceil(NumY/stride) bytes.  Returns (lines, cols); pair with view_end()."""
        self.flush()
        addr = out if isinstance(out, int) else out.ctypes.data
        nl, nc = C.c_int32(), C.c_int32()
        L.check(self._h, L.lib.fb_view_u8_begin(self._h, kind, stride, addr, C.byref(nl), C.byref(nc)))
        return nl.value, nc.value

    def view_end(self):
        """Wait for the view started by view_begin() / view_u8_begin(); returns (min, max)."""
        mn, mx = C.c_float(), C.c_float()
        L.check(self._h, L.lib.fb_view_end(self._h, C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    # ---- the frame loop either side of Simulate (main/main.go Draw / advectParticles) -----------
    def Render(self, kind: int, color_range=None) -> np.ndarray:
        """Draw's pixel pass on the device (main/main.go:550-574, 620-652; main/colors.go): the view
        through the UI's colormap, solid cells black, as the RGBA image [NumY][NumX][4] of
        fluidToImageIndex.  Returns the image; .last_range holds the (min, max) it was coloured with."""
        self.flush()
        out = np.zeros((self.NumY, self.NumX, 4), dtype=np.uint8)
        rng = (C.c_float * 2)(*color_range) if color_range is not None else None
        mn, mx = C.c_float(), C.c_float()
        L.check(self._h, L.lib.fb_render(self._h, kind, out.ctypes.data, rng, C.byref(mn), C.byref(mx)))
        self.last_range = (mn.value, mx.value)
        return out

    def render_begin(self, kind: int, out, color_range=None) -> None:
        """Pipelined Render (fb_render_begin): `out` is a pinned uint8 array [NumY][NumX][4] or its address."""
        self.flush()
        addr = out if isinstance(out, int) else out.ctypes.data
        rng = (C.c_float * 2)(*color_range) if color_range is not None else None
        L.check(self._h, L.lib.fb_render_begin(self._h, kind, addr, rng))

    def render_end(self):
        mn, mx = C.c_float(), C.c_float()
        L.check(self._h, L.lib.fb_render_end(self._h, C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    def AdvectParticles(self, particles: np.ndarray, dt: float) -> np.ndarray:
        """advectParticles (main/main.go:512-546) for a structured array of PARTICLE_DTYPE; returns the
        survivors in their original order."""
        self.flush()
        ps = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE).copy()
        n = C.c_size_t()
        L.check(self._h, L.lib.fb_advect_particles(self._h, ps.ctypes.data, len(ps), dt, C.byref(n)))
        return ps[:n.value]

    def Smoke(self) -> ScalarField:               # smoke.go:5
        return self._view(L.VIEW_SMOKE)

    def Pressure(self) -> ScalarField:            # pressure.go:5
        return self._view(L.VIEW_PRESSURE)

    def VelocityMagnitude(self) -> ScalarField:   # fluid.go:841
        return self._view(L.VIEW_VELOCITY_MAGNITUDE)

    def Vorticity(self) -> ScalarField:           # fluid.go:806
        return self._view(L.VIEW_VORTICITY)

    def Velocity(self) -> VectorField:            # velocity.go:3
        return VectorField(self.get("U"), self.get("V"))

    def _reduce(self, kind: int) -> float:
        self.flush()
        if self.compat:
            self._upload_compat()
        out = C.c_float()
        L.check(self._h, L.lib.fb_reduce(self._h, kind, C.byref(out)))
        return out.value

    def MaxDivergence(self) -> float:             # fluid.go:876
        return self._reduce(L.REDUCE_MAX_DIVERGENCE)

    def GetAdaptiveTimeStep(self, basedt: float) -> float:   # fluid.go:529-557
        f32 = np.float32
        max_vel = f32(self._reduce(L.REDUCE_MAX_ABS_VELOCITY))
        if max_vel == 0:
            return float(f32(basedt))
        adaptive = (f32(0.8) * f32(self.h)) / max_vel
        adaptive = max(min(adaptive, f32(basedt) * f32(2.0)), f32(basedt) * f32(0.1))
        return float(adaptive)

    def SampleVelocity(self, x: float, y: float):            # fluid.go:799
        uv = self.SampleVelocities(np.array([[x, y]], dtype=np.float32))
        return float(uv[0, 0]), float(uv[0, 1])

    def SampleVelocities(self, xy: np.ndarray) -> np.ndarray:
        """Batched SampleVelocity: the particle loop of main/main.go:512-546 makes
        two samples per particle per frame."""
        self.flush()
        xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
        uv = np.zeros_like(xy)
        L.check(self._h, L.lib.fb_sample_velocity(self._h, len(xy), xy.ctypes.data, uv.ctypes.data))
        return uv

    # ---- plumbing --------------------------------------------------------------------------
    def synchronize(self):
        L.check(self._h, L.lib.fb_synchronize(self._h))

    def timer_start(self):
        L.check(self._h, L.lib.fb_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        L.check(self._h, L.lib.fb_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def set_option(self, option: int, value: int):
        L.check(self._h, L.lib.fb_set_option(self._h, option, value))

    def profile(self, on: bool = True):
        L.check(self._h, L.lib.fb_profile_enable(self._h, int(on)))

    def profile_read(self) -> dict:
        """{phase: (ms_total, calls)} since the last read (CUDA events per phase)."""
        n = len(L.PROF_PHASES)
        ms, calls = (C.c_float * n)(), (C.c_int32 * n)()
        L.check(self._h, L.lib.fb_profile_read(self._h, ms, calls))
        return {name: (float(ms[k]), int(calls[k])) for k, name in enumerate(L.PROF_PHASES)}

    def launch_count(self) -> int:
        n = C.c_uint64()
        L.check(self._h, L.lib.fb_launch_count(self._h, C.byref(n)))
        return n.value

    def cuda_stream(self) -> int:
        s = C.c_void_p()
        L.check(self._h, L.lib.fb_stream(self._h, C.byref(s)))
        return s.value or 0


def New(density: float, width: int, height: int, h: float, **kw) -> Fluid:
    """fluid.New (fluid.go:42)."""
    return Fluid(density, width, height, h, **kw)
