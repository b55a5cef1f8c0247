"""Edit command lists: the wire format of ``fb_edit_cmd`` (include/fluidb200.h).

The reference edits one cell per call (pkg/fluid/walls.go:5-93, fluid.go:761-771,
894-907).  A command here is the same edit applied to a half-open rectangle
``[i0,i1) x [j0,j1)``; a point edit is the 1x1 rectangle.  Lists apply in order.
"""
from __future__ import annotations

import numpy as np

EDIT_DTYPE = np.dtype(
    [("op", "<i4"), ("i0", "<i4"), ("j0", "<i4"), ("i1", "<i4"), ("j1", "<i4"), ("a", "<f4"), ("b", "<f4")]
)
assert EDIT_DTYPE.itemsize == 28

SET_SOLID = 0
SET_VELOCITY = 1
ADD_SMOKE = 2
APPLY_FORCE = 3
CIRCLE_OBSTACLE = 4
RESET = 5
SET_VELOCITY_IF_FLUID = 6
ADD_SMOKE_IF_FLUID = 7
SET_SMOKE = 8


def cmd(op: int, i0: int = 0, j0: int = 0, i1: int = 0, j1: int = 0, a: float = 0.0, b: float = 0.0):
    return (op, i0, j0, i1, j1, a, b)


def pack(cmds) -> np.ndarray:
    """List of ``cmd(...)`` tuples (or an already packed array) -> contiguous array."""
    if isinstance(cmds, np.ndarray) and cmds.dtype == EDIT_DTYPE:
        return np.ascontiguousarray(cmds)
    return np.array(list(cmds), dtype=EDIT_DTYPE)


def set_solid(i: int, j: int, value: bool):
    return cmd(SET_SOLID, i, j, i + 1, j + 1, 1.0 if value else 0.0)


def set_solid_rect(i0: int, j0: int, i1: int, j1: int, value: bool):
    return cmd(SET_SOLID, i0, j0, i1, j1, 1.0 if value else 0.0)


def set_velocity(i: int, j: int, u: float, v: float):
    return cmd(SET_VELOCITY, i, j, i + 1, j + 1, u, v)


def add_smoke(i: int, j: int, amount: float):
    return cmd(ADD_SMOKE, i, j, i + 1, j + 1, amount)


def apply_force(i: int, j: int, fx: float, fy: float):
    return cmd(APPLY_FORCE, i, j, i + 1, j + 1, fx, fy)


def circle_obstacle(cx: int, cy: int, radius: int):
    return cmd(CIRCLE_OBSTACLE, cx, cy, radius, 0)


def reset():
    return cmd(RESET)
