"""Time of one fused pressure solve against the number of iterations (= active pipeline stages) on the
config-5 input at 4096^2; with FLUIDB200_RBQ_X set, parts of the kernel are switched off (timing only).
usage: [FLUIDB200_RBQ_X=31] python tools/rbq_iters.py [iterations ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluid_b200
from fluid_b200 import _lib as L
from fluid_b200 import presets
size = 4096
p = presets.projection_stress(size, size)
f = fluid_b200.New(p.density, size, size, p.h, solver=2)
u, v = presets.projection_fields(size + 2, size + 2, 0, size + 2)
f.set("U", u); f.set("V", v); f.edit(p.init); f.edit(p.per_step)
f.set_option(L.OPT_SOLVE_STATS, 0)
# fingerprint of one 8-iteration solve of the prepared field: equal across builds <=> bit-identical results
import hashlib
f.project(8, p.dt)
fp = hashlib.sha1(b"".join(f.get(n).tobytes() for n in ("U", "V", "p"))).hexdigest()[:12]
out = []
for k in [int(a) for a in sys.argv[1:]] or (1, 2, 4, 8):
    for _ in range(3):
        f.project(k, p.dt)
    f.timer_start()
    for _ in range(10):
        f.project(k, p.dt)
    out.append((k, round(f.timer_stop() / 10, 4)))
print("X=" + os.environ.get("FLUIDB200_RBQ_X", "0"), "ms per solve by iterations:", out, "sha1(U,V,p after the first solve)", fp)
