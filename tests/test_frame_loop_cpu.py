"""The oracle's restatement of the UI code either side of Simulate (main/main.go Draw /
advectParticles, main/colors.go) against answers worked out by hand from the Go source.
The reference has no tests for main/; these known answers are what pins the restatement."""
import numpy as np
import pytest

import oracle


def small(nx=6, ny=5, h=1.0):
    o = oracle.New(1.0, nx, ny, h)
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            o.SetSolid(i, j, False)
    return o


def test_sci_colormap_known_answers():
    o = small()
    M = o.get("M")
    M[1, 1], M[2, 1], M[3, 1], M[4, 1], M[5, 1] = 0.0, 0.25, 0.5, 0.75, 1.0
    o.set("M", M)
    img = o.Render(0)                      # smoke: getSciValue(val, 0, 1), colors.go:48-84
    NY = o.NumY
    px = lambda i, j: tuple(int(v) for v in img[NY - 1 - j, i])
    assert px(1, 1) == (0, 0, 255, 255)    # num 0, s 0
    assert px(2, 1) == (0, 255, 255, 255)  # num 1, s 0
    assert px(3, 1) == (0, 255, 0, 255)    # num 2, s 0
    assert px(4, 1) == (255, 255, 0, 255)  # num 3, s 0
    # val = max - 0.0001 -> 0.9999: num 3, s = 0.9996, g = 1 - s = 0.0004 -> uint8(0.102) = 0
    assert px(5, 1) == (255, 0, 0, 255)
    assert px(0, 0) == (0, 0, 0, 255)      # ring cells are solid: black (main.go:564-574)
    assert img.shape == (o.NumY, o.NumX, 4) and (img[..., 3] == 255).all()


def test_constant_field_and_diverging_colormap():
    o = small()
    assert tuple(o.Render(1)[2, 2]) == (0, 255, 0, 255)       # pressure all zero: d <= 0 -> val 0.5 -> num 2, s 0
    assert tuple(o.Render(3)[2, 2]) == (255, 255, 255, 255)   # vorticity all zero: absMax < 1e-8 -> white
    V = o.get("V")
    V[3, 2] = 1.0                                              # curl = dV/dx - dU/dy: +0.5 at i=2, -0.5 at i=4
    o.set("V", V)
    img = o.Render(3)
    NY = o.NumY
    assert tuple(img[NY - 1 - 2, 2]) == (255, 0, 0, 255)       # t = +1: white -> red
    assert tuple(img[NY - 1 - 2, 4]) == (0, 0, 255, 255)       # t = -1: white -> blue


def test_particles_uniform_flow_known_answers():
    import fluid_b200                                          # dtype only; no GPU call
    o = small(8, 8, 0.5)
    U = o.get("U"); U[:] = 2.0; o.set("U", U)                  # u = 2 everywhere, v = 0
    ps = np.zeros(4, dtype=fluid_b200.PARTICLE_DTYPE)
    ps["x"], ps["y"] = [1.0, 2.0, 4.4, 1.0], [2.0, 2.0, 2.0, 0.2]
    ps["max_age"] = [1.0, 0.05, 1.0, 1.0]
    out = o.AdvectParticles(ps, 0.125)
    # particle 1 expires (age 0.125 > 0.05); particle 2 ends at x = 4.65 -> cell 9 = NumX-1, the solid ring;
    # particle 3 sits in the solid ring row j = 0
    assert len(out) == 1
    assert out["x"][0] == np.float32(1.25) and out["y"][0] == np.float32(2.0) and out["age"][0] == np.float32(0.125)
