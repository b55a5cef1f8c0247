import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluid_b200
from fluid_b200 import presets
from fluid_b200.parallel import LocalSlabGroup
size = (int(sys.argv[1]), int(sys.argv[2])); nslabs = int(sys.argv[3])
p = presets.projection_stress(*size)
u, v = presets.projection_fields(size[0] + 2, size[1] + 2, 0, size[0] + 2)
single = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2)
single.set("U", u); single.set("V", v); single.edit(p.init); single.edit(p.per_step)
group = LocalSlabGroup(p.density, p.width, p.height, p.h, nslabs, solver=2, ghost=32, reach=1)
u0, v0 = single.get("U"), single.get("V")
for s in group.slabs:
    s.f.set("U", u0); s.f.set("V", v0); s.f.edit(p.init)
for k in range(3):
    single.project(8, p.dt); group.project(8, p.dt)
    for name in ("U", "V", "p"):
        got, want = group.get(name), single.get(name)
        bad = np.argwhere(got != want)
        print(f"solve {k} {name}: mismatches={len(bad)}", "lines %d..%d cols %d..%d" % (bad[:,0].min(), bad[:,0].max(), bad[:,1].min(), bad[:,1].max()) if len(bad) else "", flush=True)
