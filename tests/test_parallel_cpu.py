"""Host-side logic of the slab decomposition, on CPU: partition arithmetic, ghost-zone
requirements, and the halo exchange itself over torch.distributed with the gloo backend
at world_size 2 (the same exchange_halos() the NCCL path calls)."""
import os
import socket

import numpy as np
import pytest


def test_partition_covers_grid_contiguously():
    from fluid_b200.parallel import partition
    for width in (7, 300, 4096, 16384 * 8):
        for n in (1, 2, 3, 4, 8):
            if width < n:
                continue
            parts = partition(width, n)
            assert parts[0][0] == 0 and parts[-1][1] == width + 2
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c and a < b
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 2 + 1      # ring lines go to the end ranks


def test_required_ghost_matches_step_extents():
    """required_ghost() mirrors the extents fb_step_local derives (csrc/fluidb200.cu)."""
    from fluid_b200.parallel import PROJECTION_HALO, reach_for, required_ghost
    for w in (1, 3, 6, 9):
        for bfecc in (False, True):
            for conf in (False, True):
                if bfecc:
                    e_final = 2 * w + 1
                    e_ct = e_final + 3 * w
                else:
                    e_ct = 1 + w
                e_proj = e_ct + (2 if conf else 0)
                assert required_ghost(w, bfecc, conf) == max(e_proj + PROJECTION_HALO, 3 * w)
    # the reference's jet: 4.0 * (1/120) / 0.01 = 3.33 cells a step
    assert reach_for(1.0 / 120.0, 0.01, 4.0) == 6


def test_halo_plan():
    from fluid_b200.parallel import halo_plan
    assert halo_plan(0, 1) == []
    assert halo_plan(0, 2) == [(1, 1)]
    assert halo_plan(1, 2) == [(0, 0)]
    assert halo_plan(2, 4) == [(0, 1), (1, 3)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, width, pitch, ghost, q):
    import torch
    import torch.distributed as dist
    from fluid_b200.parallel import exchange_halos, halo_plan, partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = partition(width, world)
        lo, hi = parts[rank]
        alloc0 = max(lo - ghost, 0)
        alloc1 = min(hi + ghost, width + 2)
        # plane value at (i, j) = 1000*field + i + j/4096: the global pattern a correct exchange reproduces
        fields = {}
        for f in range(3):
            t = torch.full((alloc1 - alloc0, pitch), -1.0)
            ii = torch.arange(lo, hi, dtype=torch.float32)[:, None]
            jj = torch.arange(pitch, dtype=torch.float32)[None, :]
            t[lo - alloc0:hi - alloc0] = 1000.0 * f + ii + jj / 4096.0
            fields[f] = t
        plan = halo_plan(rank, world)
        regions = {}
        for f, t in fields.items():
            for side, _peer in plan:
                if side == 0:
                    send, recv = t[lo - alloc0:lo - alloc0 + ghost], t[lo - ghost - alloc0:lo - alloc0]
                else:
                    send, recv = t[hi - ghost - alloc0:hi - alloc0], t[hi - alloc0:hi + ghost - alloc0]
                regions[(f, side)] = (send.reshape(-1), recv.reshape(-1))
        exchange_halos(dist, regions, plan)
        ok = True
        for f, t in fields.items():
            for i in range(alloc0, alloc1):
                want = 1000.0 * f + i + torch.arange(pitch, dtype=torch.float32) / 4096.0
                if not torch.equal(t[i - alloc0], want):
                    ok = False
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 64 * world, 40, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]


def test_slab_step_exchanges_only_when_the_ghost_lines_are_stale():
    """Host logic of SlabFluid.step (no GPU): with the exchange in front of the step every step exchanges; with the
    overlapped exchange (FB_OPT_HALO_OVERLAP) only the first step does, and again after project(), which leaves the
    ghost lines one solve behind; the halo check runs every `check_every` steps."""
    from fluid_b200 import parallel

    class Stub(parallel.SlabFluid):
        def __init__(self):                      # none of the device state
            object.__setattr__(self, "log", [])
            self.nranks, self.transport, self.overlap = 2, "peer", False
            self._ghost_fresh, self._ghost_stale, self._steps_since_check = False, False, 0
            self.check_every, self.adaptive_reach, self.reach = 4, False, 6

            class F:                             # what project() / set_overlap() touch
                def project(_, n, dt): self.log.append("solve")
                def set_option(_, o, v): self.log.append(("option", o, v))
            self.f = F()

        def exchange(self):
            self.log.append("exchange")
            self._ghost_stale = self._ghost_fresh = False

        def step_no_exchange(self, dt, per_step=None):
            self.log.append("step")
            self._steps_since_check += 1

        def check_halo(self):
            self.log.append("check")
            self._steps_since_check = 0

    s = Stub()
    s.step(0.01, 3)
    assert s.log == ["exchange", "step"] * 3
    s.log.clear()
    s.set_overlap(True)
    s.step(0.01, 5)
    assert s.log == [("option", parallel.L.OPT_HALO_OVERLAP, 1), "exchange", "step", "check", "step", "step", "step", "step", "check"]
    s.log.clear()
    s.project(8, 0.01)                           # exchanges for itself, then the ghost lines lag
    s.step(0.01, 2)
    assert s.log == ["exchange", "solve", "exchange", "step", "step"]
    s.log.clear()
    s.set_overlap(False)
    s.step(0.01, 2)
    assert s.log == [("option", parallel.L.OPT_HALO_OVERLAP, 0), "exchange", "step", "exchange", "step", "check"]
    with pytest.raises(ValueError):
        s.transport = "nccl"
        s.set_overlap(True)
