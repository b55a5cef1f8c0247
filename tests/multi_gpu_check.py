"""Multi-process slab check, run under torchrun on N GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank owns a row slab (fluid_b200.parallel.SlabFluid, NCCL halo exchange); rank 0 also
runs the whole grid on its own GPU and compares the gathered fields bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import fluid_b200
    from fluid_b200 import presets
    from fluid_b200.parallel import SlabFluid, required_ghost

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = 0
    cases = [presets.jet(256 * world, 192), presets.karman(200 * world, 160),
             presets.karman(240 * world, 128, bfecc=False, confinement=0.0)]
    # halo exchange: CUDA-IPC peer memory (explicit, then overlapped with the step: FB_OPT_HALO_OVERLAP) and NCCL send/recv
    runs = [(cases[0], "peer", False), (cases[1], "peer", False), (cases[2], "nccl", False),
            (cases[0], "peer", True), (cases[1], "peer", True), (cases[2], "peer", True)]
    for p, transport, overlap in runs:
        bfecc = bool(p.params.get("use_bfecc", False))
        conf = float(p.params.get("confinement", 0.0))
        reach = 6
        slab = SlabFluid(p.density, p.width, p.height, p.h, solver=2, device=local, rank=rank, nranks=world,
                         ghost=required_ghost(reach, bfecc, conf != 0.0), reach=reach, transport=transport)
        slab.edit(p.init)
        slab.UseBFECC = bfecc
        slab.Confinement = conf
        slab.set_overlap(overlap)
        slab.step(p.dt, 13, p.per_step)
        if overlap:                       # a host-side edit between overlapped steps reaches the ghost lines too
            slab.edit(p.per_step)
        slab.step(p.dt, 17, p.per_step)
        slab.check_halo()
        fields = {name: slab.get(name) for name in ("U", "V", "M", "p")}
        md = slab.MaxDivergence()
        if rank == 0:
            single = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2, device=local)
            single.edit(p.init)
            single.UseBFECC = bfecc
            single.Confinement = conf
            single.step(p.dt, 13, p.per_step)
            if overlap:
                single.edit(p.per_step)
            single.step(p.dt, 17, p.per_step)
            for name, got in fields.items():
                want = single.get(name)
                bad = int(np.count_nonzero(~((got == want) | (np.isnan(got) & np.isnan(want)))))
                print(f"[{p.name} {p.width}x{p.height} bfecc={bfecc} conf={conf} {transport}{' overlapped' if overlap else ''}] "
                      f"{name}: mismatches={bad}")
                failures += bad != 0
            assert np.float32(md) == np.float32(single.MaxDivergence())
            single.close()
        slab.close()
        dist.barrier()
    # BASELINE config 5: projection only on the slabs (ghost lines of U, V refreshed before every solve)
    for transport in ("peer", "nccl"):
        size = (192 * world, 224)
        p = presets.projection_stress(*size)
        u, v = presets.projection_fields(size[0] + 2, size[1] + 2, 0, size[0] + 2)
        slab = SlabFluid(p.density, p.width, p.height, p.h, solver=2, device=local, rank=rank, nranks=world,
                         ghost=32, reach=1, transport=transport)
        slab.f.set("U", u); slab.f.set("V", v)
        slab.edit(p.init); slab.edit(p.per_step)
        divs = []
        for _ in range(3):
            slab.project(8, p.dt)
            divs.append(slab.MaxDivergence())
        fields = {name: slab.get(name) for name in ("U", "V", "p")}
        if rank == 0:
            single = fluid_b200.New(p.density, p.width, p.height, p.h, solver=2, device=local)
            single.set("U", u); single.set("V", v)
            single.edit(p.init); single.edit(p.per_step)
            want_divs = []
            for _ in range(3):
                single.project(8, p.dt)
                want_divs.append(single.MaxDivergence())
            for name, got in fields.items():
                bad = int(np.count_nonzero(got != single.get(name)))
                print(f"[projection {size[0]}x{size[1]} {transport}] {name}: mismatches={bad}")
                failures += bad != 0
            print(f"[projection {transport}] max|div| after 1..3 solves: {divs} (single GPU: {want_divs})")
            failures += [np.float32(x) for x in divs] != [np.float32(x) for x in want_divs]
            single.close()
        slab.close()
        dist.barrier()
    t = torch.tensor([failures], device="cuda")
    dist.broadcast(t, 0)
    dist.destroy_process_group()
    if int(t.item()):
        sys.exit(1)
    if rank == 0:
        print(f"multi_gpu_check OK on {world} ranks")


if __name__ == "__main__":
    main()
