"""Ablation timing of the fused pressure solve (FLUIDB200_RBQ_X bits: 1 skip sweeps, 2 skip writer I/O, 4 skip TMA)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluid_b200
from fluid_b200 import presets, _lib as L

p = presets.jet(4096, 4096)
g = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2)
g.edit(p.init); g.step(p.dt, 5, p.per_step); g.edit(p.per_step)
g.set_option(L.OPT_SOLVE_STATS, 0)
for iters in (1, 2, 4, 8):
    for _ in range(3):
        g.makeIncompressible(iters, p.dt)
    g.synchronize()
    g.timer_start()
    for _ in range(10):
        g.makeIncompressible(iters, p.dt)
    ms = g.timer_stop() / 10
    print(f"x={os.environ.get('FLUIDB200_RBQ_X','0')} iters={iters} nstages={2*iters}: {ms*1000:.1f} us")
g.close()
