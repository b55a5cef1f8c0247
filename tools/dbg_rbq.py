import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, fluid_b200, oracle
from fluid_b200 import presets
from common import apply_preset, copy_state, diff_report
for size, iters in (((200,120),1), ((200,120),8), ((520,75),8), ((2048,2048),8)):
    p = presets.karman(*size)
    g = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2)
    apply_preset(g, p); g.edit(p.per_step)
    t=time.time()
    try:
        g.makeIncompressible(iters, p.dt); g.synchronize(); print(size, iters, "ok", round(time.time()-t,3), g.solve_stats()["max_div"][-1])
    except Exception as e:
        print(size, iters, "ERR", round(time.time()-t,3), e)
    g.close()
