"""Repeat one fused pressure solve of the many-chunks case (4098 x 1282 cells, 16 chunks x 3 strips) from the same input and
compare every result with the first: a hand-off race shows as a run that differs.  Prints where.
usage: python tools/rbq_race_hunt.py [repetitions] [width height]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluid_b200
from fluid_b200 import _lib as L
from fluid_b200 import presets

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
size = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4096, 1280)
p = presets.projection_stress(*size)
u, v = presets.projection_fields(size[0] + 2, size[1] + 2, 0, size[0] + 2)
stats = int(os.environ.get("HUNT_STATS", "1"))
ref = None
bad_runs = 0
t0 = time.time()
for r in range(reps):
    # a new handle every few runs (the test makes one per solve), otherwise re-upload the input
    if r % 8 == 0:
        if r:
            g.close()
        g = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2)
        g.set_option(L.OPT_SOLVE_STATS, stats)
    g.set("U", u); g.set("V", v); g.set("p", np.zeros_like(u))
    g.edit(p.init); g.edit(p.per_step)
    g.project(8, p.dt)
    out = {n: g.get(n) for n in ("U", "V", "p")}
    if ref is None:
        ref = out
        continue
    for n in ("U", "V", "p"):
        d = np.argwhere(out[n].view(np.uint32) != ref[n].view(np.uint32))
        if len(d):
            bad_runs += 1
            ii, jj = d[:, 0], d[:, 1]
            print(f"run {r}: {n}: {len(d)} cells differ; lines {ii.min()}..{ii.max()} (distinct {len(np.unique(ii))}: {np.unique(ii)[:12]}...), "
                  f"columns {jj.min()}..{jj.max()} (distinct {len(np.unique(jj))}); max|diff| {np.abs(out[n] - ref[n]).max()}", flush=True)
            break
print(f"{reps} runs, {bad_runs} differ from the first ({time.time() - t0:.0f} s); library {L.LIB_PATH}")
