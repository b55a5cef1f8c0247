"""Row-slab decomposition of one grid over the GPUs of a node: one process per GPU.

The reference's only parallelism is a fork-join over lines of constant i
(pkg/fluid/parallel.go:10-39).  Here a line of constant i is a contiguous row of
NumY floats, so a slab of consecutive lines is one contiguous block per field.
Rank r owns interior lines ``1 + W*r/R .. 1 + W*(r+1)/R`` (the ring lines 0 and
NumX-1 go to the first / last rank) plus ``ghost`` lines on each side.

Per step there is ONE exchange: the ``ghost`` lines of U, V and M on each side
are refreshed from the neighbours.  Default transport ``"peer"``: every rank packs
its boundary lines into a send buffer exported with CUDA IPC and pulls its ghost
lines straight out of the neighbours' buffers over NVLink, the pull kernel waiting
on an epoch flag in peer memory (fb_halo_exchange) -- no collective, no host
hand-shake, three kernels on the library's own stream.  Transport ``"nccl"``:
send/recv of the same lines through torch.distributed (kept for comparison).  Inside the step every
phase is recomputed redundantly on as many ghost lines as later phases read
(``fb_step_local``), so results are bit-identical to the single-GPU run.  Views and
reductions are local passes followed by a max/min all-reduce of two floats.

The exact (lexicographic) solver cannot be decomposed this way (it is a wavefront
through the whole grid); slabs use the red-black solvers.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib as L
from . import edits as E
from .fluid import Fluid

PROJECTION_HALO = 16   # dependence radius of 8 fused red-black iterations (rbq_fused.cuh)


def partition(width: int, nranks: int):
    """[(i_lo, i_hi)] per rank -- the same arithmetic as fb_create (csrc/fluidb200.cu)."""
    out = []
    for r in range(nranks):
        lo = 1 + width * r // nranks
        hi = 1 + width * (r + 1) // nranks
        if r == 0:
            lo = 0
        if r == nranks - 1:
            hi = width + 2
        out.append((lo, hi))
    return out


def reach_for(dt: float, h: float, max_speed: float) -> int:
    """Lines a semi-Lagrangian trace can travel in one step, bilinear tap and the
    staggering half cell included."""
    return int(math.ceil(dt * max_speed / h)) + 2


def required_ghost(reach: int, bfecc: bool, confinement: bool) -> int:
    """Ghost lines needed so that a whole step needs no communication (mirrors the
    extents of fb_step_local)."""
    w = max(reach, 1)
    e_ct = (5 * w + 1) if bfecc else (1 + w)
    return max(e_ct + (2 if confinement else 0) + PROJECTION_HALO, 3 * w)


def halo_plan(rank: int, nranks: int):
    """[(side, peer)] of the exchanges this rank takes part in: side 0 = lower i."""
    plan = []
    if rank > 0:
        plan.append((0, rank - 1))
    if rank < nranks - 1:
        plan.append((1, rank + 1))
    return plan


def exchange_halos(dist, regions, plan, group=None):
    """regions[(field, side)] = (send_tensor, recv_tensor).  Posts every send/recv as one
    batch (ncclGroupStart/End under NCCL; plain isend/irecv under gloo) and waits on the
    current stream.  Transport-agnostic: the CPU tests drive it with gloo tensors."""
    ops = []
    for (field, side), (send_t, recv_t) in regions.items():
        peer = dict(plan)[side]
        ops.append(dist.P2POp(dist.isend, send_t, peer, group))
        ops.append(dist.P2POp(dist.irecv, recv_t, peer, group))
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def exchange_local(slabs):
    """Halo exchange between slabs that live in ONE process (several handles on one or
    more devices).  Peer transport: the same pack / publish / pull kernels as between
    processes, connected by device address; every slab posts before any slab pulls.
    Otherwise plain device-to-device copies."""
    import torch
    if slabs and slabs[0].transport == "peer":
        for s in slabs:
            L.check(s.f._h, L.lib.fb_halo_post(s.f._h))
        for s in slabs:
            L.check(s.f._h, L.lib.fb_halo_pull(s.f._h))
        return
    for s in slabs:
        s.f.synchronize()
    regs = [s._regions() for s in slabs]
    for r, s in enumerate(slabs):
        for field in SlabFluid.EXCHANGED:
            if r + 1 < len(slabs):
                send_hi, recv_hi = regs[r][(field, 1)]
                send_lo, recv_lo = regs[r + 1][(field, 0)]
                recv_lo.copy_(send_hi)
                recv_hi.copy_(send_lo)
    torch.cuda.synchronize()


class _DeviceSpan:
    """A float32 span of device memory, visible to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False),
                                         "version": 2, "strides": None}


class SlabFluid:
    """One rank's slab of a grid split over ``nranks`` GPUs, with the stepping API of
    ``Fluid`` (step / edit / get / MaxDivergence / timers)."""

    EXCHANGED = (L.U, L.V, L.M)
    # every exported knob of Fluid is forwarded to the slab's handle; anything else that looks like one is rejected
    KNOBS = ("UseBFECC", "Confinement", "TurbulenceStrength", "PressureDamping", "SmokeAdvection", "NumIters", "Solver",
             "ViscosityDiffusion", "UseMultigrid", "MultigridLevels", "Relaxation")

    def __init__(self, density, width, height, h, *, solver=L.SOLVER_REDBLACK_PRESSURE, device=0, rank=0,
                 nranks=1, ghost=None, reach=6, transport="peer", connect=True):
        import torch
        import torch.distributed as dist
        if solver == L.SOLVER_EXACT:
            raise ValueError("the lexicographic solver does not decompose into slabs; use a red-black solver")
        if ghost is None:            # enough for any step fb_step_local accepts at this reach (BFECC + confinement)
            ghost = required_ghost(reach, True, True)
        self.torch, self.dist = torch, dist
        self.rank, self.nranks, self.device = rank, nranks, device
        self.reach = self.reach_cap = reach
        self.adaptive_reach = False        # bench.py / callers opt in; the tests pin the reach they were written for
        self.check_every = 16              # steps between halo checks (+ reach updates): each one drains the stream
        self.f = Fluid(density, width, height, h, device=device, solver=solver, rank=rank, nranks=nranks,
                       ghost=ghost if nranks > 1 else 0)
        self.NumX, self.NumY = self.f.NumX, self.f.NumY
        self.i_lo, self.i_hi = self.f.i_lo, self.f.i_hi
        self.ghost = ghost if nranks > 1 else 0
        self.global_cells = self.NumX * self.NumY
        self.plan = halo_plan(rank, nranks)
        self.stream = torch.cuda.ExternalStream(self.f.cuda_stream(), device=torch.device("cuda", device)) \
            if nranks > 1 else None
        self._steps_since_check = 0
        if transport not in ("peer", "nccl"):
            raise ValueError("transport must be 'peer' or 'nccl'")
        self.transport = transport if nranks > 1 else "none"
        self.overlap = False               # set_overlap(True): the exchange for step k+1 rides inside step k
        self._ghost_fresh = False          # the ghost lines are what the neighbours hold now (only kept under overlap)
        if self.transport == "peer" and connect:
            self.connect_peers()

    # ---- peer-memory transport ---------------------------------------------------------------
    def export_halo(self):
        """(ipc_handle_bytes, device_address) of this rank's send buffer."""
        handle = (C.c_ubyte * 64)()
        ptr, nbytes = C.c_uint64(), C.c_size_t()
        L.check(self.f._h, L.lib.fb_halo_export(self.f._h, self.ghost, handle, C.byref(ptr), C.byref(nbytes)))
        return bytes(handle), int(ptr.value)

    def connect_peers(self):
        """One process per GPU: all-gather the IPC handles and attach both neighbours."""
        handle, _ptr = self.export_halo()
        gathered = [None] * self.nranks
        self.dist.all_gather_object(gathered, handle)
        for side, peer in self.plan:
            buf = (C.c_ubyte * 64).from_buffer_copy(gathered[peer])
            L.check(self.f._h, L.lib.fb_halo_connect(self.f._h, side, buf, 0))
        self.dist.barrier()

    # knobs are forwarded to the slab's Fluid
    def __getattr__(self, name):
        if name in SlabFluid.KNOBS:
            return getattr(self.f, name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in SlabFluid.KNOBS:
            setattr(self.f, name, value)       # unsupported combinations (viscosity, multigrid on slabs) fail in fb_step_local
        elif name[:1].isupper() and name not in ("NumX", "NumY"):
            raise AttributeError(f"SlabFluid has no knob {name!r}")
        else:
            object.__setattr__(self, name, value)

    def close(self):
        self.f.close()

    def set_overlap(self, on=True):
        """FB_OPT_HALO_OVERLAP: fb_step_local refreshes the ghost lines for the next step while it finishes the current one
        (U, V during the smoke passes, M during the interior of the last smoke pass, on a second stream), so that step()
        exchanges explicitly only before the first step and after host-side changes.  Peer-memory transport between
        processes only; bit-identical to the explicit exchange (tests/multi_gpu_check.py, bench.py's parity_check)."""
        if on and self.nranks > 1 and self.transport != "peer":
            raise ValueError("the overlapped exchange needs the peer-memory transport")
        self.overlap = bool(on) and self.nranks > 1
        self._ghost_fresh = False
        if self.nranks > 1:
            self.f.set_option(L.OPT_HALO_OVERLAP, 1 if self.overlap else 0)

    def edit(self, cmds):
        """Every rank applies the same (global-coordinate) commands; the kernels clip them
        to the lines the rank holds, ghosts included, so ghosts stay consistent."""
        self.f.edit(cmds)

    def _regions(self):
        f, regions = self.f, {}
        for field in self.EXCHANGED:
            for side, _peer in self.plan:
                sp, rp, nb = C.c_void_p(), C.c_void_p(), C.c_size_t()
                L.check(f._h, L.lib.fb_halo_region(f._h, field, side, self.ghost, C.byref(sp), C.byref(rp), C.byref(nb)))
                dev = self.torch.device("cuda", self.device)
                regions[(field, side)] = (self.torch.as_tensor(_DeviceSpan(sp.value, nb.value), device=dev),
                                          self.torch.as_tensor(_DeviceSpan(rp.value, nb.value), device=dev))
        return regions

    def exchange(self):
        self._ghost_stale = False
        self._ghost_fresh = False
        if self.nranks == 1:
            return
        if self.transport == "peer":
            L.check(self.f._h, L.lib.fb_halo_exchange(self.f._h))
            return
        with self.torch.cuda.stream(self.stream):
            exchange_halos(self.dist, self._regions(), self.plan)

    def step_no_exchange(self, dt, per_step=None):
        """One fb_step_local; the caller has refreshed the ghost lines."""
        f = self.f
        f.flush()
        p = f.params()
        arr = E.pack(per_step) if per_step is not None and len(per_step) else None
        L.check(f._h, L.lib.fb_step_local(f._h, C.byref(p), dt, self.reach,
                                          arr.ctypes.data if arr is not None else None,
                                          len(arr) if arr is not None else 0))
        self._steps_since_check += 1

    def step(self, dt, nsteps=1, per_step=None):
        arr = E.pack(per_step) if per_step is not None and len(per_step) else None
        for _ in range(nsteps):
            if not (self.overlap and self._ghost_fresh):
                self.exchange()
            self.step_no_exchange(dt, arr)
            self._ghost_fresh = self.overlap      # the step left the next step's ghost lines behind
            if self._steps_since_check >= self.check_every:
                self.check_halo()
                if self.adaptive_reach:
                    self.adapt_reach(dt)

    def adapt_reach(self, dt):
        """size the semi-Lagrangian reach from the all-reduced max |u| instead of a constant.  Every phase
        of fb_step_local is recomputed on as many ghost lines as the reach asks for, so a reach that follows the flow
        (with a 1.5x margin, re-measured every `check_every` steps together with the halo check) trims the redundant work; a trace
        that outruns it still raises FB_ERR_HALO at the next check, never a wrong result.  Never above the reach the
        ghost zone was allocated for."""
        speed = self.max_speed()
        self.reach = max(2, min(self.reach_cap, reach_for(dt, self.f.h, 1.5 * speed)))

    def project(self, numIters, dt):
        """fill(p,0) + makeIncompressible(numIters) on the slab, after refreshing the ghost lines of U, V
        (one pass of <= 8 iterations reads 16 of them; BASELINE config 5)."""
        if numIters > 8 and self.nranks > 1:
            raise ValueError("a slab solve is one pass of at most 8 iterations between halo exchanges")
        self.exchange()
        self.f.project(numIters, dt)
        self._ghost_stale = True      # only owned lines are solved: the ghost lines now lag one solve behind
        self._ghost_fresh = False

    def check_halo(self):
        self._steps_since_check = 0
        L.check(self.f._h, L.lib.fb_check_halo(self.f._h))

    # ---- reductions / gathers --------------------------------------------------------------
    def _allreduce(self, value: float, op):
        if self.nranks == 1:
            return value
        t = self.torch.tensor([value], dtype=self.torch.float32, device=self.torch.device("cuda", self.device))
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def MaxDivergence(self) -> float:
        if getattr(self, "_ghost_stale", False):
            self.exchange()           # the divergence of the last owned line reads the first ghost line
        return self._allreduce(self.f.MaxDivergence(), self.dist.ReduceOp.MAX)

    def max_speed(self) -> float:
        return self._allreduce(self.f._reduce(L.REDUCE_MAX_ABS_VELOCITY), self.dist.ReduceOp.MAX)

    def get(self, name: str):
        """Gather a field on every rank (tests / checkpoints; not a hot path)."""
        local = self.f.get(name)
        if self.nranks == 1:
            return local
        t = self.torch.from_numpy(local).to(self.torch.device("cuda", self.device))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)   # owned line ranges are disjoint, the rest is zero
        return t.cpu().numpy()

    def solve_stats(self):
        return self.f.solve_stats()

    # plumbing used by bench.py
    def timer_start(self):
        self.f.timer_start()

    def timer_stop(self):
        return self.f.timer_stop()

    def launch_count(self):
        return self.f.launch_count()

    def profile(self, on=True):
        self.f.profile(on)

    def profile_read(self):
        return self.f.profile_read()

    def set_option(self, option, value):
        self.f.set_option(option, value)

    def synchronize(self):
        self.f.synchronize()


class LocalSlabGroup:
    """All slabs of a grid inside ONE process (handles on one device, or one per visible
    device), halos exchanged with device copies.  Same results as the multi-process path;
    used to check slab-vs-single-domain bit identity on a single-GPU box."""

    def __init__(self, density, width, height, h, nslabs, *, solver=L.SOLVER_REDBLACK_PRESSURE, devices=None,
                 ghost=None, reach=6, transport="peer"):
        devices = devices or [0] * nslabs
        self.slabs = [SlabFluid(density, width, height, h, solver=solver, device=devices[r], rank=r, nranks=nslabs,
                                ghost=ghost, reach=reach, transport=transport, connect=False) for r in range(nslabs)]
        if transport == "peer" and nslabs > 1:
            for s in self.slabs:
                s.export_halo()
            for s in self.slabs:                                    # same process: attach by handle
                for side, peer in s.plan:
                    L.check(s.f._h, L.lib.fb_halo_connect_local(s.f._h, side, self.slabs[peer].f._h))
        self.NumX, self.NumY = self.slabs[0].NumX, self.slabs[0].NumY

    def __setattr__(self, name, value):
        if name in SlabFluid.KNOBS:
            for s in self.slabs:
                setattr(s, name, value)
        elif name in ("NumX", "NumY") or not name[:1].isupper():
            object.__setattr__(self, name, value)
        else:
            raise AttributeError(f"LocalSlabGroup has no knob {name!r}")

    def edit(self, cmds):
        for s in self.slabs:
            s.edit(cmds)

    def step(self, dt, nsteps=1, per_step=None):
        arr = E.pack(per_step) if per_step is not None and len(per_step) else None
        for _ in range(nsteps):
            exchange_local(self.slabs)
            for s in self.slabs:
                s.step_no_exchange(dt, arr)
        for s in self.slabs:
            s.check_halo()

    def project(self, numIters, dt):
        exchange_local(self.slabs)
        for s in self.slabs:
            s.f.project(numIters, dt)

    def get(self, name):
        out = np.zeros((self.NumX, self.NumY), dtype=np.float32)
        for s in self.slabs:
            local = s.f.get(name)
            out[s.i_lo:s.i_hi] = local[s.i_lo:s.i_hi]
        return out

    def close(self):
        for s in self.slabs:
            s.close()
