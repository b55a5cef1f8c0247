import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluid_b200
from fluid_b200 import _lib as L
from fluid_b200 import presets
size = (int(sys.argv[1]), int(sys.argv[2]))
p = presets.projection_stress(*size)
u, v = presets.projection_fields(size[0] + 2, size[1] + 2, 0, size[0] + 2)
for mode in ("sync", "async", "async-nostats"):
    f = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2)
    f.set("U", u); f.set("V", v); f.edit(p.init); f.edit(p.per_step)
    if mode == "async-nostats":
        f.set_option(L.OPT_SOLVE_STATS, 0)
    out = []
    for k in range(4):
        f.project(8, p.dt)
        if mode == "sync":
            out.append(f.MaxDivergence())
    out.append(f.MaxDivergence())
    U = f.get("U")
    print(size, mode, out, "finite" if np.isfinite(U).all() else "NOT FINITE", float(np.abs(U).max()))
    f.close()
