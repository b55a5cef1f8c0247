"""Run-to-run determinism of the fused pressure solve on the config-5 input (debugging aid): the same solve three
times on fresh handles, compared bit for bit; then against the face-form kernel within its stated tolerance.
usage: [FLUIDB200_LIB=...] python tools/dbg_proj_determinism.py W H"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluid_b200
from fluid_b200 import presets
size = (int(sys.argv[1]), int(sys.argv[2]))
p = presets.projection_stress(*size)
u, v = presets.projection_fields(size[0] + 2, size[1] + 2, 0, size[0] + 2)
res = []
for solver in (2, 2, 2, 1):
    f = fluid_b200.New(p.density, p.width, p.height, p.h, solver=solver)
    f.set("U", u); f.set("V", v); f.edit(p.init); f.edit(p.per_step)
    f.clearPressure(); f.makeIncompressible(8, p.dt)
    res.append((f.get("U"), f.get("V"), f.get("p")))
    f.close()
tag = os.path.basename(os.environ.get("FLUIDB200_LIB", "default"))
for k in (1, 2):
    out = []
    for name, a, b in zip("UVp", res[0], res[k]):
        bad = np.argwhere(a != b)
        out.append(f"{name}: {len(bad)}" + (" (lines %d..%d)" % (bad[:, 0].min(), bad[:, 0].max()) if len(bad) else ""))
    print(f"[{tag} {size}] run0 vs run{k} mismatches  " + "  ".join(out))
d = [float(np.abs(a - b).max()) for a, b in zip(res[0][:2], res[3][:2])]
print(f"[{tag} {size}] pressure form vs face form: max abs U {d[0]:.3e} V {d[1]:.3e}")
