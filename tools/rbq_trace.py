"""Timeline of CTA (0, 0) of k_rbq_fused from a -DRQ_TRACE build (FLUIDB200_LIB=fluid_b200/variants/lib_trace.so): per role and line
the clock64 stamps 0 = step entered, 1 = waits satisfied, 2 = (stages) past the named barrier / (loader, writer) arrived, 3 = step left.
usage: FLUIDB200_RBQ_TRACE=/tmp/t.bin python tools/rbq_trace.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluid_b200
from fluid_b200 import _lib as L
from fluid_b200 import presets

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
path = os.environ["FLUIDB200_RBQ_TRACE"]
size = 4096
p = presets.projection_stress(size, size)
f = fluid_b200.New(p.density, size, size, p.h, solver=2)
u, v = presets.projection_fields(size + 2, size + 2, 0, size + 2)
f.set("U", u); f.set("V", v); f.edit(p.init); f.edit(p.per_step)
f.set_option(L.OPT_SOLVE_STATS, 0)
for _ in range(3):
    f.project(iters, p.dt)
t = np.fromfile(path, dtype=np.int64).reshape(10, 512, 4)
names = ["loader"] + [f"stage{k}" for k in range(8)] + ["writer"]
t0 = t[t > 0].min()
print(f"X={os.environ.get('FLUIDB200_RBQ_X', '0')} iterations={iters}")
for r in range(10):
    a = t[r]
    lines = np.nonzero(a[:, 0] > 0)[0]
    if len(lines) < 20:
        continue
    mid = lines[len(lines) // 4: 3 * len(lines) // 4]                      # steady state
    enter = a[mid, 0]
    order = np.argsort(enter)
    period = np.diff(np.sort(enter)).mean()
    wait = (a[mid, 1] - a[mid, 0]).mean()
    seg2 = (a[mid, 2] - a[mid, 1]).mean() if (a[mid, 2] > 0).all() else float("nan")
    last = a[mid, 3] if (a[mid, 3] > 0).all() else a[mid, 2]
    prev = a[mid, 2] if (a[mid, 3] > 0).all() else a[mid, 1]
    seg3 = (last - prev).mean()
    print(f"{names[r]:8s} lines {lines.min()}..{lines.max()} ({len(lines)}): period {period:7.1f} cycles per line; in the step: waiting {wait:7.1f}, "
          f"then {seg2:7.1f}, then {seg3:7.1f}; first entered at {a[lines.min(), 0] - t0}, last left at {a[lines.max()].max() - t0}")
