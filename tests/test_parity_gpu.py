"""Parity of the CUDA path against the CPU oracle on the same inputs, through the
C ABI.  The bar for every kernel with a counterpart in the reference is
BIT-EXACT float32 (the kernels perform the reference's operations in the
reference's order, without FMA contraction); nothing here uses a tolerance
except where stated.
"""
import numpy as np
import pytest

from common import FIELDS, apply_preset, assert_bit_exact, copy_state, diff_report

pytestmark = pytest.mark.gpu

STATE = ("U", "V", "newU", "newV", "p", "S", "M", "newM")
# The fused path keeps the scratch arrays newU/newV/newM current only where the
# reference can observe them (faces / cells that advection skips); they are compared
# in full when the handle is created with exact_shadow (white-box mode) or literal.
OBSERVABLE = ("U", "V", "p", "S", "M")


def new_pair(preset, solver=0):
    import fluid_b200
    import oracle
    o = oracle.New(preset.density, preset.width, preset.height, preset.h, solver=solver)
    g = fluid_b200.New(preset.density, preset.width, preset.height, preset.h, solver=solver)
    apply_preset(o, preset)
    apply_preset(g, preset)
    return o, g


def developed_state(preset, steps=25):
    """An oracle a few steps into a preset (non-trivial fields, stale scratch buffers)."""
    import oracle
    o = oracle.New(preset.density, preset.width, preset.height, preset.h)
    apply_preset(o, preset)
    o.step(preset.dt, steps, preset.per_step)
    o.edit(preset.per_step)   # jet re-imposed: inflow faces differ from the stale scratch (Q-6)
    return o


def gpu_clone(o, preset, **kw):
    import fluid_b200
    g = fluid_b200.New(preset.density, preset.width, preset.height, preset.h, **kw)
    copy_state(g, o)
    return g


def assert_state_equal(g, o, tag, fields=OBSERVABLE):
    for name in fields:
        assert_bit_exact(f"{tag}:{name}", g.get(name), o.get(name))


def test_edits_and_preset_init_match():
    from fluid_b200 import presets
    for p in (presets.jet(64, 48), presets.cavity(40, 40), presets.karman(120, 60)):
        o, g = new_pair(p)
        o.edit(p.per_step)
        g.edit(p.per_step)
        assert_state_equal(g, o, p.name, STATE)
        g.close()


@pytest.mark.parametrize("mode", ["literal", "exact_shadow"])
@pytest.mark.parametrize("preset_name", ["jet", "cavity", "karman"])
def test_full_state_including_scratch_arrays(mode, preset_name):
    """With the reference-shaped kernels (literal) or with exact_shadow the unexported
    scratch arrays newU/newV/newM match the reference everywhere too."""
    import fluid_b200
    import oracle
    from fluid_b200 import presets
    p = {"jet": presets.jet(90, 70), "cavity": presets.cavity(64, 64), "karman": presets.karman(120, 64)}[preset_name]
    o = oracle.New(p.density, p.width, p.height, p.h)
    g = fluid_b200.New(p.density, p.width, p.height, p.h, **{mode: True})
    apply_preset(o, p)
    apply_preset(g, p)
    o.step(p.dt, 25, p.per_step)
    g.step(p.dt, 25, p.per_step)
    assert_state_equal(g, o, f"{mode}/{preset_name}", STATE)
    g.close()


@pytest.mark.parametrize("solver", [0, 1])
def test_fused_path_equals_literal_path(solver):
    """A/B: the fused kernels (pointer swaps, mask, fused BFECC / confinement / red-black
    passes) against the first, reference-shaped CUDA implementation, bit for bit."""
    import fluid_b200
    from fluid_b200 import presets
    for p in (presets.karman(300, 200), presets.jet(257, 130, bfecc=False), presets.cavity(130, 190)):
        a = fluid_b200.New(p.density, p.width, p.height, p.h, solver=solver)
        b = fluid_b200.New(p.density, p.width, p.height, p.h, solver=solver, literal=True)
        for f in (a, b):
            apply_preset(f, p)
            f.step(p.dt, 30, p.per_step)
        for name in OBSERVABLE:
            assert_bit_exact(f"{p.name}:{name}", a.get(name), b.get(name))
        sa, sb = a.solve_stats(), b.solve_stats()
        assert sa["sweeps_run"] == sb["sweeps_run"]
        assert np.array_equal(np.float32(sa["max_div"]), np.float32(sb["max_div"])), (sa, sb)
        a.close()
        b.close()


def test_smoke_in_solid_cells_keeps_stale_value():
    """newM is never written at solid cells (fluid.go:411): smoke added to a solid cell
    reverts at the next advectSmoke, smoke in a cell that then becomes solid freezes."""
    from fluid_b200 import presets
    p = presets.jet(60, 40)
    o, g = new_pair(p)
    for f in (o, g):
        f.step(p.dt, 5, p.per_step)
        f.AddSmoke(20, 20, 0.75)
        f.step(p.dt, 2, p.per_step)
        f.SetSolid(20, 20, True)
        f.SetSolid(21, 20, True)
        f.step(p.dt, 2, p.per_step)
        f.AddSmoke(21, 20, 2.0)        # into a solid cell
        f.step(p.dt, 3, p.per_step)
        f.SetSolid(20, 20, False)
        f.step(p.dt, 3, p.per_step)
    assert_state_equal(g, o, "solid-smoke")
    g.close()


PHASES = [
    ("handleBorders", lambda f, dt: f.handleBorders()),
    ("advectVelocity", lambda f, dt: f.advectVelocity(dt)),
    ("advectSmoke", lambda f, dt: f.advectSmoke(dt)),
    ("addTurbulence", lambda f, dt: f.addTurbulence(dt)),
    ("applyVorticityConfinement", lambda f, dt: f.applyVorticityConfinement(dt)),
    ("advectVelocityBFECC", lambda f, dt: f.advectVelocityBFECC(dt)),
    ("advectSmokeBFECC", lambda f, dt: f.advectSmokeBFECC(dt)),
    ("applyViscosity", lambda f, dt: f.applyViscosity(dt)),
    ("makeIncompressible8", lambda f, dt: f.makeIncompressible(8, dt)),
    ("makeIncompressible20", lambda f, dt: f.makeIncompressible(20, dt)),
    ("makeIncompressible3", lambda f, dt: f.makeIncompressible(3, dt)),
]


@pytest.mark.parametrize("phase", [p[0] for p in PHASES])
@pytest.mark.parametrize("preset_name", ["jet", "karman"])
def test_single_phase_bit_exact(phase, preset_name):
    """One kernel group at a time on a developed flow (single-kernel parity)."""
    from fluid_b200 import presets
    p = presets.jet(150, 97) if preset_name == "jet" else presets.karman(141, 90)
    o = developed_state(p)
    o.Confinement = 0.1
    o.ViscosityDiffusion = 0.05 if phase in ("applyViscosity", "advectSmoke") else 0.0
    g = gpu_clone(o, p)
    fn = dict(PHASES)[phase]
    fn(o, p.dt)
    fn(g, p.dt)
    assert_state_equal(g, o, f"{preset_name}/{phase}")
    g.close()


@pytest.mark.parametrize("make,steps", [
    ("jet", 60), ("cavity", 40), ("karman", 60),
])
def test_presets_exact_solver_bit_exact(make, steps):
    """All three presets, exact (lexicographic wavefront) solver: every field
    bit-identical to the oracle after N steps."""
    from fluid_b200 import presets
    p = {"jet": presets.jet(200, 121), "cavity": presets.cavity(128, 128), "karman": presets.karman(220, 110)}[make]
    o, g = new_pair(p)
    o.step(p.dt, steps, p.per_step)
    g.step(p.dt, steps, p.per_step)
    assert_state_equal(g, o, make)
    st = g.solve_stats()
    assert st["sweeps_run"] == 8
    assert np.float32(st["max_div"][-1]) == np.float32(o.solve_stats()["last_max_div"])
    g.close()


def test_default_grid_jet_600_steps_bit_exact():
    """BASELINE config 1: jet preset at the reference's default 300x251 grid."""
    from fluid_b200 import presets
    p = presets.jet()
    o, g = new_pair(p)
    for n in (1, 1, 8, 90, 500):
        o.step(p.dt, n, p.per_step)
        g.step(p.dt, n, p.per_step)
        assert_state_equal(g, o, "jet300x251", FIELDS)
    g.close()


def test_exact_solver_early_exit_matches():
    """fluid.go:175: a quiescent field converges after one sweep; the fused wavefront
    must roll back and report exactly the sweeps the reference runs."""
    from fluid_b200 import presets
    p = presets.jet(64, 40)
    o, g = new_pair(p)
    o.Simulate(p.dt)          # no jet imposed: all-zero velocities
    g.Simulate(p.dt)
    st = g.solve_stats()
    assert st["sweeps_run"] == o.solve_stats()["sweeps_run"] == 1
    assert st["rolled_back"]
    assert_state_equal(g, o, "quiescent")
    # tiny divergence that dies out after a few sweeps
    for f in (o, g):
        f.SetVelocity(20, 20, 3e-5, 0.0)
        f.Simulate(p.dt)
    assert g.solve_stats()["sweeps_run"] == o.solve_stats()["sweeps_run"]
    assert_state_equal(g, o, "tiny")
    g.close()


@pytest.mark.parametrize("iters", [1, 4, 8, 11])
def test_redblack_solver_bit_exact_and_reaches_reference_residual(iters):
    """Fast mode: red-black ordering of the reference's per-cell update.  (1) the CUDA
    solve is bit-identical to the CPU restatement of the same ordering; (2) with the
    reference's 8 iterations and omega schedule its residual is not worse than the
    reference's lexicographic solve on the same input (north_star: 'reach the same
    residual, iteration count stated')."""
    import oracle
    from fluid_b200 import presets
    p = presets.karman(200, 120)
    base = developed_state(p, steps=30)
    rb_cpu = oracle.New(p.density, p.width, p.height, p.h, solver=oracle.SOLVER_REDBLACK)
    copy_state(rb_cpu, base)
    g = gpu_clone(base, p, solver=1)
    rb_cpu.makeIncompressible(iters, p.dt)
    g.makeIncompressible(iters, p.dt)
    assert_state_equal(g, rb_cpu, f"redblack{iters}")
    if iters == 8:
        before = base.MaxDivergence()
        base.makeIncompressible(8, p.dt)
        lex, rb = base.MaxDivergence(), g.MaxDivergence()
        assert rb <= lex and rb < before, (before, lex, rb)
    g.close()


@pytest.mark.parametrize("iters", [1, 3, 8, 11, 16])
@pytest.mark.parametrize("size", [(200, 120), (520, 75), (61, 700)])
def test_pressure_form_solver_bit_exact_vs_its_restatement(iters, size):
    """FB_SOLVER_REDBLACK_PRESSURE (rbq_fused.cuh): bit-identical to the CPU restatement of
    the same arithmetic (oracle fo_project_redblack_q), for pass splits 8+3 and 8+8 too, on
    grids that exercise several strips / chunks; and equal to the face-form red-black solve
    within float32 rounding (stated tolerance: 2e-5 max-abs on velocities of O(1..10))."""
    import oracle
    from fluid_b200 import presets
    p = presets.karman(*size)
    base = developed_state(p, steps=12)
    cpu = oracle.New(p.density, p.width, p.height, p.h, solver=oracle.SOLVER_REDBLACK_PRESSURE)
    copy_state(cpu, base)
    g = gpu_clone(base, p, solver=2)
    cpu.makeIncompressible(iters, p.dt)
    g.makeIncompressible(iters, p.dt)
    assert_state_equal(g, cpu, f"rbq{iters}")
    st = g.solve_stats()
    assert st["sweeps_run"] == iters
    assert np.float32(st["max_div"][-1]) == np.float32(cpu.solve_stats()["last_max_div"])
    face = oracle.New(p.density, p.width, p.height, p.h, solver=oracle.SOLVER_REDBLACK)
    copy_state(face, base)
    face.makeIncompressible(iters, p.dt)
    for name in ("U", "V"):
        r = diff_report(name, g.get(name), face.get(name))
        assert r["max_abs"] <= 2e-5, r
    g.close()


@pytest.mark.parametrize("preset_name", ["jet", "cavity", "karman"])
def test_pressure_form_multi_step_and_residual(preset_name):
    """Whole steps with the throughput solver: bit-identical to its CPU restatement after 30
    steps (fused turbulence in the solve's write-out included), and every step's residual
    max|div| is not worse than the reference's lexicographic solve on the same input."""
    import fluid_b200
    import oracle
    from fluid_b200 import presets
    p = {"jet": presets.jet(210, 133), "cavity": presets.cavity(140, 140), "karman": presets.karman(260, 140)}[preset_name]
    o, g = new_pair(p, solver=2)
    o.step(p.dt, 30, p.per_step)
    g.step(p.dt, 30, p.per_step)
    assert_state_equal(g, o, f"rbq/{preset_name}")
    # residual criterion on the developed state
    o.edit(p.per_step)
    lex = oracle.New(p.density, p.width, p.height, p.h)
    copy_state(lex, o)
    g2 = gpu_clone(o, p, solver=2)
    lex.makeIncompressible(8, p.dt)
    g2.makeIncompressible(8, p.dt)
    assert g2.MaxDivergence() <= lex.MaxDivergence(), (g2.MaxDivergence(), lex.MaxDivergence())
    g.close()
    g2.close()


def test_redblack_multi_step_bit_exact():
    import oracle
    from fluid_b200 import presets
    p = presets.karman(180, 100)
    o, g = new_pair(p, solver=1)
    o.step(p.dt, 40, p.per_step)
    g.step(p.dt, 40, p.per_step)
    assert_state_equal(g, o, "karman-rb")
    g.close()


def test_views_and_reductions_match():
    from fluid_b200 import presets
    p = presets.karman(150, 90)
    o = developed_state(p)
    g = gpu_clone(o, p)
    for name in ("Smoke", "Pressure", "VelocityMagnitude", "Vorticity"):
        a, b = getattr(g, name)(), getattr(o, name)()
        assert_bit_exact(name, a.values, b.values)
        assert np.float32(a.MinValue) == np.float32(b.MinValue), name
        assert np.float32(a.MaxValue) == np.float32(b.MaxValue), name
        assert a.Value(10, 10) == b.Value(10, 10)
    assert np.float32(g.MaxDivergence()) == np.float32(o.MaxDivergence())
    assert np.float32(g.GetAdaptiveTimeStep(0.08)) == np.float32(o.GetAdaptiveTimeStep(0.08))
    rng = np.random.default_rng(7)
    xy = rng.uniform(-0.2, 1.8, size=(500, 2)).astype(np.float32)
    assert_bit_exact("SampleVelocity", g.SampleVelocities(xy), o.SampleVelocities(xy))
    vf, of = g.Velocity(), o.Velocity()
    assert vf.Value(7, 9) == of.Value(7, 9)
    # sentinels when no fluid cell exists (fluid.go:810-811)
    import fluid_b200
    e = fluid_b200.New(1.0, 4, 4, 1.0)
    v = e.Vorticity()
    assert v.MinValue == np.finfo(np.float32).max and v.MaxValue == -np.finfo(np.float32).max
    e.close()
    g.close()


def test_pipelined_view_equals_blocking_view_and_overlaps_next_step():
    """fb_view_begin / fb_view_end: the frame that lands in host memory is the state at the time of
    `begin`, even though a Simulate is queued behind it before `end`; min / max as fb_view."""
    import fluid_b200
    from fluid_b200 import presets, _lib as L
    p = presets.karman(150, 90)
    o = developed_state(p)
    g = gpu_clone(o, p)
    kinds = {"Smoke": L.VIEW_SMOKE, "Pressure": L.VIEW_PRESSURE, "VelocityMagnitude": L.VIEW_VELOCITY_MAGNITUDE,
             "Vorticity": L.VIEW_VORTICITY}
    for name, kind in kinds.items():
        want = getattr(o, name)()
        out = np.full((g.NumX, g.NumY), np.nan, dtype=np.float32)
        g.view_begin(kind, out)
        with pytest.raises(fluid_b200.FluidError):
            g.view_begin(kind, out)                      # one view in flight per handle
        g.step(p.dt, 1, p.per_step)                      # overwrites M / U / V / p behind the snapshot
        mn, mx = g.view_end()
        assert_bit_exact(name, out, want.values)
        assert np.float32(mn) == np.float32(want.MinValue) and np.float32(mx) == np.float32(want.MaxValue), name
        o.edit(p.per_step); o.Simulate(p.dt)
    with pytest.raises(fluid_b200.FluidError):
        g.view_end()                                     # nothing in flight
    assert_bit_exact("Smoke after", g.Smoke().values, o.Smoke().values)
    g.close()


@pytest.mark.parametrize("stride", [1, 3, 4])
def test_decimated_u8_view_is_the_quantised_full_view(stride):
    """fb_view_u8_begin / fb_view_end: min / max equal the full view's, and every byte is the stated quantisation
    (float32, no FMA) of the cell it samples -- the frame-loop view for fields larger than a display."""
    from fluid_b200 import _lib as L
    from fluid_b200 import presets
    p = presets.karman(150, 90)
    o = developed_state(p)
    g = gpu_clone(o, p)
    for kind in (L.VIEW_SMOKE, L.VIEW_PRESSURE, L.VIEW_VELOCITY_MAGNITUDE, L.VIEW_VORTICITY):
        full = np.zeros((g.NumX, g.NumY), dtype=np.float32)
        g.view_begin(kind, full)
        mn, mx = g.view_end()
        ni, nj = -(-g.NumX // stride), -(-g.NumY // stride)
        out = np.zeros(ni * nj + 64, dtype=np.uint8)
        got = g.view_u8_begin(kind, stride, out)
        mn8, mx8 = g.view_end()
        assert got == (ni, nj) and (mn8, mx8) == (mn, mx)
        lo, hi = np.float32(mn), np.float32(mx)
        scale = np.float32(255.0) / (hi - lo) if hi > lo else np.float32(0.0)
        v = full[::stride, ::stride]
        want = (np.minimum(np.maximum((v - lo) * scale, np.float32(0.0)), np.float32(255.0)) + np.float32(0.5)).astype(np.uint32)
        assert np.array_equal(out[:ni * nj].reshape(ni, nj), want.astype(np.uint8)), kind
    g.close()


@pytest.mark.parametrize("preset_name", ["karman", "cavity"])
def test_render_matches_the_ui_pixel_pass(preset_name):
    """fb_render == Draw's pixel pass (main/main.go:550-574, 620-652; main/colors.go) restated in the oracle:
    all four views, both colormaps, solid overlay, image layout -- byte for byte."""
    from fluid_b200 import presets, _lib as L
    p = {"karman": presets.karman(150, 90), "cavity": presets.cavity(97, 61)}[preset_name]
    o = developed_state(p)
    g = gpu_clone(o, p)
    for kind in (L.VIEW_SMOKE, L.VIEW_PRESSURE, L.VIEW_VELOCITY_MAGNITUDE, L.VIEW_VORTICITY):
        got, want = g.Render(kind), o.Render(kind)
        assert got.shape == (g.NumY, g.NumX, 4)
        bad = np.argwhere((got != want).any(axis=2))
        assert len(bad) == 0, f"kind {kind}: {len(bad)} pixels differ, first {bad[:3].tolist()}"
    # explicit colour range (what a multi-GPU host passes after its all-reduce) and the pipelined form
    mn, mx = g.last_range
    out = np.zeros((g.NumY, g.NumX, 4), dtype=np.uint8)
    g.render_begin(L.VIEW_VORTICITY, out, (mn, mx))
    g.step(p.dt, 1, p.per_step)
    assert g.render_end() == (mn, mx)
    assert (out == o.Render(L.VIEW_VORTICITY)).all()
    # degenerate ranges: a field that is constant (d <= 0 -> 0.5) and all-zero vorticity (white)
    import fluid_b200, oracle
    e, eo = fluid_b200.New(1.0, 6, 5, 1.0), oracle.New(1.0, 6, 5, 1.0)
    for f in (e, eo):
        for i in range(1, 7):
            for j in range(1, 6):
                f.SetSolid(i, j, False)
    for kind in (L.VIEW_SMOKE, L.VIEW_VORTICITY):
        assert (e.Render(kind) == eo.Render(kind)).all()
    e.close(); g.close()


def test_advect_particles_matches_the_ui_tracer():
    """fb_advect_particles == advectParticles (main/main.go:512-546): ageing, RK2 midpoint through
    SampleVelocity, bounds / solid culling, survivors in order -- bit for bit."""
    import fluid_b200
    from fluid_b200 import presets
    p = presets.karman(150, 90)
    o = developed_state(p)
    g = gpu_clone(o, p)
    rng = np.random.default_rng(11)
    n = 5000
    ps = np.zeros(n, dtype=fluid_b200.PARTICLE_DTYPE)
    ps["x"] = rng.uniform(-0.05, (p.width + 2) * p.h + 0.05, n).astype(np.float32)     # some start outside
    ps["y"] = rng.uniform(-0.05, (p.height + 2) * p.h + 0.05, n).astype(np.float32)
    ps["r"], ps["g"], ps["b"] = rng.integers(0, 256, (3, n), dtype=np.uint8)
    ps["age"] = rng.uniform(0, 2, n).astype(np.float32)
    ps["max_age"] = rng.uniform(0.5, 3, n).astype(np.float32)                          # some expire
    ps["x"][:3] = [np.nan, np.inf, -np.inf]
    a, b = ps.copy(), ps.copy()
    for _ in range(5):
        a, b = g.AdvectParticles(a, 0.05), o.AdvectParticles(b, 0.05)
        assert len(a) == len(b) and 0 < len(a) < n
        assert a.tobytes() == b.tobytes()
    assert len(g.AdvectParticles(ps[:0], 0.05)) == 0
    g.close()


def test_apply_force_radius_and_misc_edits_match():
    from fluid_b200 import presets
    p = presets.jet(60, 40)
    o, g = new_pair(p)
    for f in (o, g):
        f.ApplyForceRadius(20, 20, 7.5, -2.5, 6)
        f.ApplyForceRadius(1, 1, 1.0, 1.0, 3)       # clipped at the ring
        f.ApplyForceRadius(30, 30, 1.0, 2.0, 0)     # radius 0 -> ApplyForce
        f.SetCircularObstacle(40, 20, 5)
        f.SetCircularObstacle(0, 0, 4)              # clipped to the domain
        f.ApplyForce(40, 20, 9.0, 9.0)              # solid: ignored
        f.AddSmoke(3, 3, 0.25)
        f.SetSolid(10, 10, True)
        f.SetSolid(10, 10, False)
        f.flush()
    assert_state_equal(g, o, "edits")
    g.close()


def test_quirk_q6_stale_scratch_at_inlet():
    """Q-6: the inlet face set by the caller reverts to the stale newU before
    advectSmoke reads it; a ping-pong implementation would leave 4.0 there."""
    from fluid_b200 import presets
    p = presets.jet(80, 60)
    o, g = new_pair(p)
    g.step(p.dt, 3, p.per_step)
    o.step(p.dt, 3, p.per_step)
    U = g.get("U")
    assert np.all(U[1, 1:-1] == 0.0)
    assert_state_equal(g, o, "q6")
    g.close()


def test_large_grid_properties():
    """At a BASELINE-scale size the oracle is too slow to run inside the suite, so check
    size-independent properties: the fused/unfused solvers agree bit for bit, the
    projection lowers max|div|, nothing goes non-finite, smoke stays non-negative."""
    import fluid_b200
    from fluid_b200 import presets
    p = presets.karman(2048, 1024)
    a = fluid_b200.New(p.density, p.width, p.height, p.h, solver=0)
    apply_preset(a, p)
    a.step(p.dt, 5, p.per_step)
    a.edit(p.per_step)
    before = a.MaxDivergence()
    a.makeIncompressible(8, p.dt)
    after = a.MaxDivergence()
    assert after < before
    a.step(p.dt, 3, p.per_step)
    for name in FIELDS:
        assert np.all(np.isfinite(a.get(name))), name
    assert a.get("M").min() >= 0.0
    a.close()


def test_parity_report_three_presets():
    """Field-by-field max-abs and relative-L2 after N steps for the three presets in both
    solver modes (the numbers DESIGN.md quotes).  Exact mode must be 0.  Red-black mode is
    a different ordering of an unconverged 8-sweep solve, so its fields differ from the
    reference's at the size of the solver residual (tens of percent in V after one step,
    see DESIGN.md); what it must match is the CPU restatement of the same ordering (bit
    for bit) and the reference's residual: max|div| after each step's solve."""
    import fluid_b200
    import oracle
    from fluid_b200 import presets
    rows = []
    for p in (presets.jet(200, 121), presets.cavity(128, 128), presets.karman(220, 110)):
        o = oracle.New(p.density, p.width, p.height, p.h)
        apply_preset(o, p)
        o.step(p.dt, 20, p.per_step)
        orb = oracle.New(p.density, p.width, p.height, p.h, solver=1)
        apply_preset(orb, p)
        orb.step(p.dt, 20, p.per_step)
        for solver in (0, 1):
            g = fluid_b200.New(p.density, p.width, p.height, p.h, solver=solver)
            apply_preset(g, p)
            g.step(p.dt, 20, p.per_step)
            for name in FIELDS:
                r = diff_report(name, g.get(name), o.get(name))
                r.update(preset=p.name, solver="exact" if solver == 0 else "redblack", steps=20)
                rows.append(r)
                if solver == 0:
                    assert r["numeric_mismatch"] == 0, r
                else:
                    assert_bit_exact(f"{p.name}:rb:{name}", g.get(name), orb.get(name))
                    assert np.all(np.isfinite(g.get(name)))
            g.close()
    print("\nPARITY_REPORT " + repr(rows))


@pytest.mark.parametrize("preset_name,nslabs,transport", [("jet", 2, "peer"), ("karman", 2, "peer"), ("cavity", 3, "peer"),
                                                          ("karman_plain", 4, "peer"), ("karman", 3, "nccl")])
def test_slab_decomposition_is_bit_identical(preset_name, nslabs, transport):
    """Row slabs with ghost lines, one halo exchange per step and redundant ghost
    computation (fb_step_local) give exactly the single-domain result.  The slabs live in
    one process here: transport "peer" runs the library's own pack / publish / pull kernels
    (fb_halo_post / fb_halo_pull, neighbours attached by handle), "nccl" stands in for the
    send/recv path with plain device copies of the fb_halo_region lines.  The multi-process
    versions of both run in tests/multi_gpu_check.py."""
    import fluid_b200
    from fluid_b200 import presets
    from fluid_b200.parallel import LocalSlabGroup, required_ghost
    p = {"jet": presets.jet(330, 140), "karman": presets.karman(420, 150),
         "cavity": presets.cavity(390, 130), "karman_plain": presets.karman(500, 90, bfecc=False, confinement=0.0)}[preset_name]
    bfecc = bool(p.params.get("use_bfecc", False))
    conf = float(p.params.get("confinement", 0.0)) != 0.0
    reach = 6
    ghost = required_ghost(reach, bfecc, conf)
    single = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2)
    group = LocalSlabGroup(p.density, p.width, p.height, p.h, nslabs, solver=2, ghost=ghost, reach=reach,
                           transport=transport)
    for f in (single, group):
        f.edit(p.init)
        f.UseBFECC = bfecc
        f.Confinement = float(p.params.get("confinement", 0.0))
        f.step(p.dt, 25, p.per_step)
    for name in OBSERVABLE:
        assert_bit_exact(f"slabs/{preset_name}:{name}", group.get(name), single.get(name))
    group.close()
    single.close()


def projection_case(size):
    """BASELINE config 5 at a small size: SplitMix64 velocity field, obstacle lattice, walls, sources."""
    from fluid_b200 import presets
    p = presets.projection_stress(*size)
    u, v = presets.projection_fields(size[0] + 2, size[1] + 2, 0, size[0] + 2)

    def prepare(f):
        f.set("U", u); f.set("V", v)
        f.edit(p.init); f.edit(p.per_step)
    return p, prepare


@pytest.mark.parametrize("size", [(256, 256), (600, 200)])
def test_projection_stress_config5_parity_and_residual(size):
    """configs[4] input at a size the oracle finishes at once: fb_phase(PROJECT) of the
    pressure-form solver == its CPU restatement bit for bit (U, V, p), and the residual criterion of
    the north star: max|div| after the 8 red-black iterations <= the reference's after its 8
    lexicographic sweeps on the same input."""
    import fluid_b200
    import oracle
    p, prepare = projection_case(size)
    lex = oracle.New(p.density, p.width, p.height, p.h, solver=oracle.SOLVER_EXACT)
    cpu = oracle.New(p.density, p.width, p.height, p.h, solver=oracle.SOLVER_REDBLACK_PRESSURE)
    g = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2)
    for f in (lex, cpu, g):
        prepare(f)
    for name in ("U", "V", "S"):
        assert_bit_exact(f"config5 input:{name}", g.get(name), cpu.get(name))
    before = lex.MaxDivergence()
    assert np.float32(g.MaxDivergence()) == np.float32(before)
    for f in (lex, cpu, g):
        f.project(8, p.dt)
    assert_state_equal(g, cpu, "config5 project", fields=("U", "V", "p"))
    after_lex, after_gpu = lex.MaxDivergence(), g.MaxDivergence()
    assert after_gpu <= after_lex < before, (before, after_lex, after_gpu)
    # a second solve starts from p = 0 again (Simulate's fill) and keeps converging
    cpu.project(8, p.dt); g.project(8, p.dt)
    assert_state_equal(g, cpu, "config5 project x2", fields=("U", "V", "p"))
    assert g.MaxDivergence() < after_gpu
    g.close()


def test_pressure_form_many_chunks_bit_exact_and_deterministic():
    """A grid tall enough for 16 chunks of 257 lines and several strips (4098 x 1282 cells): the fused solve equals
    its CPU restatement bit for bit and repeats bit for bit.  (A hand-off race at the first line of a chunk --
    iteration 0 read line 0 of its window before the loader warp that owns it had written it -- only showed at
    this scale: ~0.1 % of the cells of the first owned line of every chunk but the first, differently on every run.)"""
    import fluid_b200
    import oracle
    p, prepare = projection_case((4096, 1280))
    cpu = oracle.New(p.density, p.width, p.height, p.h, solver=oracle.SOLVER_REDBLACK_PRESSURE)
    prepare(cpu)
    cpu.project(8, p.dt)
    runs = []
    for _ in range(3):
        g = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2)
        prepare(g)
        g.project(8, p.dt)
        for name in ("U", "V", "p"):       # a failure says WHERE (a hand-off race shows in particular lines of a chunk)
            got, want = g.get(name), cpu.get(name)
            d = np.argwhere(~((got == want) | (np.isnan(got) & np.isnan(want))))      # +0 / -0 compare equal, as in assert_bit_exact
            assert len(d) == 0, (f"many chunks, solve {len(runs)}, {name}: {len(d)} cells differ; lines {d[:, 0].min()}..{d[:, 0].max()} "
                                 f"(first distinct: {np.unique(d[:, 0])[:16].tolist()}), columns {d[:, 1].min()}..{d[:, 1].max()} "
                                 f"(first distinct: {np.unique(d[:, 1])[:16].tolist()}); tools/rbq_race_hunt.py repeats this solve")
        runs.append(g.get("p"))
        g.close()
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])


@pytest.mark.parametrize("nslabs,size", [(2, (512, 160)), (4, (512, 160)), (2, (4096, 1280))])
def test_projection_on_slabs_is_bit_identical(nslabs, size):
    """The strong-scaling path of bench.py --workload project*: ghost lines of U, V refreshed, then
    fb_phase(PROJECT) on every slab -- same bits as the single-domain solve, solve after solve."""
    import fluid_b200
    from fluid_b200.parallel import LocalSlabGroup
    p, prepare = projection_case(size)
    single = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2)
    prepare(single)
    group = LocalSlabGroup(p.density, p.width, p.height, p.h, nslabs, solver=2, ghost=32, reach=1)
    u, v = single.get("U"), single.get("V")
    for s in group.slabs:
        s.f.set("U", u); s.f.set("V", v)
        s.f.edit(p.init)
    for k in range(3):
        single.project(8, p.dt)
        group.project(8, p.dt)
        for name in ("U", "V", "p"):
            assert_bit_exact(f"config5 slabs solve {k}:{name}", group.get(name), single.get(name))
    group.close()
    single.close()


def test_step_local_single_rank_equals_step():
    import fluid_b200
    from fluid_b200 import presets
    from fluid_b200.parallel import SlabFluid
    p = presets.karman(200, 100)
    a = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2)
    b = SlabFluid(p.density, p.width, p.height, p.h, solver=2)
    for f in (a, b):
        f.edit(p.init)
        f.UseBFECC = True
        f.Confinement = 0.1
        f.step(p.dt, 10, p.per_step)
    for name in OBSERVABLE:
        assert_bit_exact(name, b.get(name), a.get(name))
    a.close()
    b.close()


def test_ghost_zone_too_narrow_is_an_error():
    import fluid_b200
    from fluid_b200 import presets
    from fluid_b200.parallel import LocalSlabGroup
    p = presets.karman(300, 80)
    group = LocalSlabGroup(p.density, p.width, p.height, p.h, 2, solver=2, ghost=20, reach=6)
    group.edit(p.init)
    group.UseBFECC = True
    with pytest.raises(fluid_b200.FluidError):
        group.step(p.dt, 1, p.per_step)
    group.close()


def test_fastmath_sequences_equal_the_ieee_instructions():
    """k_confine_fast's straight-line division / square root against div.rn.f32 / sqrt.rn.f32 (fb_selftest_fastmath):
    whatever operand the range test accepts must give the IEEE result bit for bit -- on random bit patterns (mode 0:
    denormals, infinities, NaN included; most are rejected) and on operands drawn inside the range (mode 1: all accepted)."""
    import ctypes as C
    from fluid_b200 import _lib as L
    for mode, n in ((0, 1 << 27), (1, 1 << 28), (2, 1 << 27), (3, 1 << 28)):
        cnt = (C.c_uint64 * 6)()
        rc = L.lib.fb_selftest_fastmath(0, n, 20261017 + mode, mode, cnt)
        assert rc == 0
        div_n, div_ok, div_diff, sq_n, sq_ok, sq_diff = [int(x) for x in cnt]
        print(f"\nFASTMATH mode={mode} div: {div_n} drawn, {div_ok} accepted, {div_diff} differ; sqrt: {sq_n} drawn, {sq_ok} accepted, {sq_diff} differ")
        assert div_n == n and sq_n == n
        assert div_diff == 0 and sq_diff == 0
        if mode >= 1:
            assert div_ok == n and sq_ok == n
        else:
            assert div_ok > 0 and sq_ok > 0


@pytest.mark.parametrize("phase", ["applyVorticityConfinement", "addTurbulence"])
def test_confinement_operand_extremes_bit_exact(phase):
    """The confinement pass on velocity fields whose magnitudes span the whole float32 range in patches -- exact zeros,
    -0, denormals, 1e-38 .. 1e-20 (below the fast sequences' range: the IEEE fall-back must take over, per thread and per
    CTA), ordinary values, and 1e15 .. 1e30 (above it; overflowing sums) -- against the oracle, bit for bit."""
    from fluid_b200 import presets
    p = presets.karman(300, 150)
    o = developed_state(p, steps=5)
    o.Confinement = 0.1
    rng = np.random.default_rng(77)
    nx, ny = o.NumX, o.NumY
    scales = np.array([0.0, 1e-45, 1e-41, 1e-38, 1e-33, 1e-30, 1e-26, 1e-20, 1e-12, 1e-5, 1.0, 30.0, 1e8, 1e15, 1e19, 1e25, 1e30],
                      dtype=np.float64)
    for name in ("U", "V"):
        # patches of 7 x 9 cells, each with its own scale (so neighbouring patches mix magnitudes at their seams)
        pick = rng.integers(0, len(scales), size=(nx // 7 + 1, ny // 9 + 1))
        sc = np.repeat(np.repeat(scales[pick], 7, axis=0), 9, axis=1)[:nx, :ny]
        a = (rng.standard_normal((nx, ny)) * sc).astype(np.float32)
        a[rng.random((nx, ny)) < 0.05] = np.float32(-0.0)
        a[rng.random((nx, ny)) < 0.05] = np.float32(0.0)
        o.set(name, a)
    g = gpu_clone(o, p)
    fn = dict(PHASES)[phase]
    with np.errstate(all="ignore"):
        fn(o, p.dt)
    fn(g, p.dt)
    for name in ("U", "V"):
        got, want = g.get(name), o.get(name)
        assert np.array_equal(np.isnan(got), np.isnan(want)), name
        assert_bit_exact(f"extremes/{phase}:{name}", got, want)
    g.close()
