"""debug: projection on slabs, back-to-back solves without host synchronisation between them"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import fluid_b200
from fluid_b200 import presets
from fluid_b200.parallel import SlabFluid
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
size = (int(sys.argv[1]), int(sys.argv[2]))
nsolves = int(sys.argv[3])
mode = sys.argv[4]      # sync | async
p = presets.projection_stress(*size)
u, v = presets.projection_fields(size[0] + 2, size[1] + 2, 0, size[0] + 2)
slab = SlabFluid(p.density, p.width, p.height, p.h, solver=2, device=local, rank=rank, nranks=world, ghost=32, reach=1)
slab.f.set("U", u); slab.f.set("V", v)
slab.edit(p.init); slab.edit(p.per_step)
for k in range(nsolves):
    slab.project(8, p.dt)
    if mode == "sync":
        slab.MaxDivergence()
md = slab.MaxDivergence()
fields = {name: slab.get(name) for name in ("U", "V", "p")}
slab.check_halo()
if rank == 0:
    single = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2, device=local)
    single.set("U", u); single.set("V", v); single.edit(p.init); single.edit(p.per_step)
    for k in range(nsolves):
        single.project(8, p.dt)
    for name, got in fields.items():
        want = single.get(name)
        bad = np.argwhere(got != want)
        print(f"[{size} {mode} x{nsolves}] {name}: mismatches={len(bad)}", "lines %d..%d cols %d..%d" % (bad[:,0].min(), bad[:,0].max(), bad[:,1].min(), bad[:,1].max()) if len(bad) else "")
    print("maxdiv", md, single.MaxDivergence())
dist.destroy_process_group()
