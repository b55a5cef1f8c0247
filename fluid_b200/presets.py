"""Synthetic preset inputs, as edit command lists.

The reference keeps its presets in the UI, not in pkg/fluid: ``NewGame``
(main/main.go:200-229), ``resetToPreset`` (main/main.go:747-793),
``applyWallSettings`` (main/main.go:840-849), the per-frame jet
(main/main.go:236-241) and ``applySources`` (main/main.go:474-486).  A preset
here is ``(init_cmds, per_step_cmds, params_overrides)``: ``init_cmds`` is what
the UI does once, ``per_step_cmds`` what it re-imposes before every Simulate.
Sizes other than the reference's 300x251 scale the jet span and the obstacle
radius with the height.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import edits as E

DENSITY = 1000.0          # main/main.go:201
SPACING = 1.0 / 100.0     # main/main.go:201
DT = float(np.float32(1.0 / 120.0))   # main/main.go:243 at speed 1


@dataclass
class Preset:
    name: str
    width: int
    height: int
    init: np.ndarray
    per_step: np.ndarray
    params: dict = field(default_factory=dict)
    density: float = DENSITY
    h: float = SPACING
    dt: float = DT


def _walls(num_x: int, num_y: int, top: bool, bottom: bool, left: bool, right: bool):
    """applyWallSettings (main/main.go:840-849), in the reference's call order:
    the j-loop runs second, so the left/right setting wins at the four corners."""
    return [
        E.set_solid_rect(0, 0, num_x, 1, bottom),
        E.set_solid_rect(0, num_y - 1, num_x, num_y, top),
        E.set_solid_rect(0, 0, 1, num_y, left),
        E.set_solid_rect(num_x - 1, 0, num_x, num_y, right),
    ]


def _clear(num_x: int, num_y: int):
    """SetSolid(i, j, false) for every cell (main/main.go:222-226, 754-758)."""
    return [E.set_solid_rect(0, 0, num_x, num_y, False)]


def _jet(height: int, span: int):
    """Per-frame jet (main/main.go:236-241): SetVelocity(1,j,4,0); AddSmoke(1,j,1)."""
    j0, j1 = height // 2 - span, height // 2 + span
    return [E.cmd(E.SET_VELOCITY, 1, j0, 2, j1, 4.0, 0.0), E.cmd(E.ADD_SMOKE, 1, j0, 2, j1, 1.0)]


def _scaled(value_at_251: int, height: int) -> int:
    return int(round(value_at_251 * height / 251.0))


def jet(width: int = 300, height: int = 251, bfecc: bool = False) -> Preset:
    """PresetJet (main/main.go:760-765) with the jet switched on."""
    nx, ny = width + 2, height + 2
    span = 100 if height == 251 else _scaled(100, height)
    init = _clear(nx, ny) + _walls(nx, ny, True, True, True, False)
    return Preset("jet", width, height, E.pack(init), E.pack(_jet(height, span)), {"use_bfecc": bfecc})


def cavity(width: int = 300, height: int = 251, bfecc: bool = True) -> Preset:
    """PresetCavity (main/main.go:767-779): four walls, a source row under the lid.
    applySources guards every source with !IsSolid (main/main.go:475-477)."""
    nx, ny = width + 2, height + 2
    init = _clear(nx, ny) + _walls(nx, ny, True, True, True, True)
    per = [
        E.cmd(E.SET_VELOCITY_IF_FLUID, 2, ny - 2, nx - 2, ny - 1, 3.0, 0.0),
        E.cmd(E.ADD_SMOKE_IF_FLUID, 2, ny - 2, nx - 2, ny - 1, 0.5),
    ]
    return Preset("cavity", width, height, E.pack(init), E.pack(per), {"use_bfecc": bfecc})


def karman(width: int = 300, height: int = 251, bfecc: bool = True, confinement: float = 0.1) -> Preset:
    """PresetKarman (main/main.go:781-790) + jet; obstacle radius 8 at H=251."""
    nx, ny = width + 2, height + 2
    radius = 8 if height == 251 else max(1, _scaled(8, height))
    span = 100 if height == 251 else _scaled(100, height)
    init = _clear(nx, ny) + [E.circle_obstacle(nx // 4, ny // 2, radius)] + _walls(nx, ny, True, True, True, False)
    return Preset("karman", width, height, E.pack(init), E.pack(_jet(height, span)),
                  {"use_bfecc": bfecc, "confinement": confinement})


def projection_stress(width: int, height: int, seed: int = 0x5EED):
    """Config 5 walls and sources: four walls, an 8x8 lattice of
    circular obstacles of radius H/64, +-5 sources on three rows.  The random
    pre-projection velocity field is produced by ``splitmix_uniform``."""
    nx, ny = width + 2, height + 2
    init = _clear(nx, ny)
    r = max(1, height // 64)
    for a in range(8):
        for b in range(8):
            init.append(E.circle_obstacle((2 * a + 1) * nx // 16, (2 * b + 1) * ny // 16, r))
    init += _walls(nx, ny, True, True, True, True)
    src = []
    for row, j in enumerate((ny // 4, ny // 2, 3 * ny // 4)):
        for k, i in enumerate(range(64, nx - 1, 64)):
            sign = 1.0 if (k + row) % 2 == 0 else -1.0
            src.append(E.cmd(E.SET_VELOCITY_IF_FLUID, i, j, i + 1, j + 1, 5.0 * sign, 0.0))
    return Preset("projection", width, height, E.pack(init), E.pack(src), {})


def splitmix_uniform(n: int, seed: int, start: int = 0) -> np.ndarray:
    """float32 uniform(-1,1) from SplitMix64, consumed in index order: values ``start .. start+n-1``
    of the stream (value k comes from state ``seed + (k+1)*gamma``, so any window can be generated
    on its own -- a rank fills only the lines it holds)."""
    idx = np.arange(start + 1, start + n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    return (u * 2.0 - 1.0).astype(np.float32)


def projection_fields(num_x: int, num_y: int, line0: int, nlines: int, seed: int = 0x5EED, chunk_lines: int = 256):
    """Config 5's pre-projection velocity field for lines ``line0 .. line0+nlines-1``:
    U takes the first NumX*NumY values of the stream in linear-index order, V the next NumX*NumY."""
    n = num_x * num_y
    out = []
    for base in (0, n):
        a = np.empty((nlines, num_y), dtype=np.float32)
        for l0 in range(0, nlines, chunk_lines):
            l1 = min(l0 + chunk_lines, nlines)
            a[l0:l1] = splitmix_uniform((l1 - l0) * num_y, seed, base + (line0 + l0) * num_y).reshape(l1 - l0, num_y)
        out.append(a)
    return out[0], out[1]


BY_NAME = {"jet": jet, "cavity": cavity, "karman": karman}
