"""Ablation timing of the fused pressure solve as Simulate runs it (pressure cleared before every solve).
FLUIDB200_RBQ_X bits: 1 skip sweeps, 2 skip writer body, 4 skip TMA."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluid_b200
from fluid_b200 import presets, _lib as L

N = int(os.environ.get("DBG_N", "4096"))
p = presets.jet(N, N)
g = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2)
g.edit(p.init); g.step(p.dt, 5, p.per_step); g.edit(p.per_step)
g.set_option(L.OPT_SOLVE_STATS, 0)
REP = 10
def timed(fn):
    for _ in range(3):
        fn()
    g.synchronize()
    g.timer_start()
    for _ in range(REP):
        fn()
    return g.timer_stop() / REP
t_clear = timed(g.clearPressure)
out = []
for iters in [int(x) for x in os.environ.get("DBG_ITERS", "1,8").split(",")]:
    def both():
        g.clearPressure(); g.makeIncompressible(iters, p.dt)
    ms = timed(both) - t_clear
    out.append(f"iters={iters}: {ms*1000:.1f} us")
print(f"x={os.environ.get('FLUIDB200_RBQ_X','0')} clear={t_clear*1000:.1f} us  " + "  ".join(out))
g.close()
