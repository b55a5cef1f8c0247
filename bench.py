#!/usr/bin/env python
"""bench.py -- throughput of the pkg/fluid per-step hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--solver redblack|exact]
    python bench.py --impl reference ...      # the CPU restatement of the Go reference on host cores

One JSON line on stdout (rank 0).  A "step" is one (*Fluid).Simulate over the whole
grid with the preset's per-frame edits (jet / sources) replayed first, exactly what
main/main.go:233-245 does per frame.  metric = cell-steps/s, a cell-step being one grid
cell (ring included: NumX*NumY per step) advanced by one Simulate.

value     device-resident: inputs live in HBM, K steps timed with CUDA events on the
          library's stream between barriers + synchronizes, max over ranks.
e2e       the frame loop a user of the Go API runs: per step the edit commands go
          host->device, Simulate runs, and the Smoke() view (field + min/max) comes
          back device->host into pinned memory (main/main.go Update + Draw), pipelined
          with fb_view_begin/end so that frame k's transfer overlaps Simulate k+1; the
          blocking variant of the same loop is reported beside it.
roofline  dominant phase of the step (largest share of device time, measured with CUDA
          events per phase inside the timed region) against the measured HBM peak.
cpu_baseline  oracle/ (C restatement of the Go reference, all host threads) on a bounded
          sample of the same preset.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# algorithmic HBM bytes per cell-step by phase
BYTES = {"project": 24, "turbulence": 20, "confinement": 20, "advect_velocity": 20, "advect_smoke": 20,
         "bfecc_velocity": 68, "bfecc_smoke": 64}

WORKLOADS = {
    # name: (preset, width, height, bfecc, confinement, BASELINE.json config it restates)
    "karman4096": ("karman", 4096, 4096, True, 0.1, "configs[2]: Karman 4096^2, BFECC + vorticity confinement"),
    "jet16384": ("jet", 16384, 16384, False, 0.0, "configs[3]: jet 16384^2 per GPU, plain semi-Lagrangian"),
    "jet4096": ("jet", 4096, 4096, False, 0.0, "jet 4096^2, plain semi-Lagrangian"),
    "cavity1024": ("cavity", 1024, 1024, True, 0.0, "configs[1]: lid-driven cavity 1024^2, BFECC (fits L2)"),
    "jet300": ("jet", 300, 251, False, 0.0, "configs[0]: jet at the reference's default 300x251 grid"),
    "karman1024": ("karman", 1024, 1024, True, 0.1, "Karman 1024^2 (CPU sample size)"),
    # projection only (fill(p,0) + makeIncompressible, fluid.go:83 + 144-234) on a FIXED grid split over the ranks
    "project32768": ("projection", 32768, 32768, False, 0.0, "configs[4]: pressure projection on a fixed 32768^2 grid with walls and sources (strong scaling)"),
    "project8192": ("projection", 8192, 8192, False, 0.0, "configs[4] at 8192^2"),
    "project4096": ("projection", 4096, 4096, False, 0.0, "configs[4] at 4096^2 (the size the CPU oracle's residual target is computed at)"),
}


def step_bytes(bfecc: bool, confinement: float, turbulence: bool = True) -> int:
    b = BYTES["project"]
    if confinement != 0.0 or turbulence:
        b += 20                      # a6+a7 share one pass
    b += (BYTES["bfecc_velocity"] + BYTES["bfecc_smoke"]) if bfecc else (BYTES["advect_velocity"] + BYTES["advect_smoke"])
    return b


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 5 ms from a thread (a 160 ms region gets
    ~30 samples), `nvidia-smi -lms` when the NVML binding is missing (its fastest loop gives a handful)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* / nvmlClocksThrottleReason* bits
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []          # (sm MHz, reasons bit mask)
        self.sm_max = None
        self.running = False
        self.source = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.device < len(ids) and ids[self.device].isdigit():
                return int(ids[self.device])
        return self.device

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.running = True
            self.source = "nvml, 5 ms"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                 str(self._physical_index())], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 20"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.running:
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)), int(reasons(self.handle))))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.running = False
            self.thread.join(timeout=1.0)
            sm = [x[0] for x in self.samples]
            mask = 0
            for x in self.samples:
                mask |= x[1]
            try:
                self.nvml.nvmlShutdown()
            except Exception:
                pass
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "samples": len(sm),
                    "reasons": sorted(k for k, b in self.BITS.items() if mask & b), "source": self.source}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": self.source}


def make_fluid(mod, preset, solver, **kw):
    f = mod.New(preset.density, preset.width, preset.height, preset.h, solver=solver, **kw)
    f.edit(preset.init)
    f.UseBFECC = bool(preset.params.get("use_bfecc", False))
    f.Confinement = float(preset.params.get("confinement", 0.0))
    return f


def build_preset(name, width, height, bfecc, confinement):
    from fluid_b200 import presets
    if name == "karman":
        return presets.karman(width, height, bfecc=bfecc, confinement=confinement)
    if name == "cavity":
        return presets.cavity(width, height, bfecc=bfecc)
    return presets.jet(width, height, bfecc=bfecc)


def go_probe():
    """`go version` when a Go toolchain is on PATH (then the reference itself could be timed), else "absent"."""
    import shutil
    exe = shutil.which("go")
    if not exe:
        return "absent"
    try:
        return subprocess.run([exe, "version"], capture_output=True, text=True, timeout=30).stdout.strip() or "absent"
    except Exception:
        return "absent"


def cpu_reference_run(workload, steps, warmup, sample_size=None, budget_s=420.0):
    """The C restatement of the Go reference (oracle/) on the host cores, all threads, on the SAME preset and grid as
    the GPU arm (`sample_size` shrinks the grid; the caller then says so).  The first warm-up step is timed: if
    (warmup + steps) of them would not fit `budget_s`, warm-up and steps are cut (never below 1 + 2) and the line
    reports what was actually run."""
    import oracle
    pname, w, h, bfecc, conf, _ = WORKLOADS[workload]
    sw, sh = sample_size if sample_size is not None else (w, h)
    p = build_preset(pname, sw, sh, bfecc, conf)
    f = make_fluid(oracle, p, oracle.SOLVER_EXACT)
    t0 = time.perf_counter()
    f.step(p.dt, 1, p.per_step)
    first = time.perf_counter() - t0
    asked = (steps, warmup)
    if first * (warmup + steps) > budget_s:
        warmup = 1
        steps = max(2, min(steps, int(budget_s / max(first, 1e-9)) - 1))
    if warmup > 1:
        f.step(p.dt, warmup - 1, p.per_step)
    t0 = time.perf_counter()
    f.step(p.dt, steps, p.per_step)
    dt = time.perf_counter() - t0
    cells = f.NumX * f.NumY
    return {"value": cells * steps / dt, "unit": "cell-steps/s", "cores": int(f.threads), "kind": "port",
            "sample": f"{pname} preset {sw}x{sh} (+ring), bfecc={bfecc}, confinement={conf}, {steps} steps after "
                      f"{warmup} warm-up, oracle/ C restatement of the Go reference (lexicographic solver), "
                      f"{int(f.threads)} threads (GOMAXPROCS-equivalent), nproc={os.cpu_count()}",
            "ms_per_step": dt / steps * 1e3, "steps_run": steps, "warmup_run": warmup, "asked": asked,
            "grid": [sw + 2, sh + 2], "same_grid_as_gpu_arm": (sw, sh) == (w, h)}


def cpu_anchor_config1():
    """BASELINE config[0] / the reference's own BenchmarkSimulateWithJet (pkg/fluid/fluid_bench_test.go:29-43): the jet
    preset at the default 302x253 grid, BFECC off, 600 steps; the author quotes ~4.6 ms/step on an Apple M4 Max."""
    import oracle
    p = build_preset("jet", 300, 251, False, 0.0)
    f = make_fluid(oracle, p, oracle.SOLVER_EXACT)
    f.step(p.dt, 20, p.per_step)
    t0 = time.perf_counter()
    f.step(p.dt, 600, p.per_step)
    dt = time.perf_counter() - t0
    return {"ms_per_step": dt / 600 * 1e3, "cell_steps_per_s": f.NumX * f.NumY * 600 / dt, "steps": 600, "grid": [f.NumX, f.NumY],
            "threads": int(f.threads), "reference_author_ms_per_step": 4.6,
            "note": "oracle/ C restatement, jet preset as main/ runs it (jet re-imposed every step); the author's figure is "
                    "Go on an Apple M4 Max (fluid_bench_test.go:29)"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def splitmix_uniform_torch(torch, n, seed, start, device):
    """presets.splitmix_uniform on a torch device (int64 arithmetic wraps like uint64; logical shifts masked)."""
    def lsr(z, k):
        return (z >> k) & ((1 << (64 - k)) - 1)

    def s64(v):
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v
    idx = torch.arange(start + 1, start + n + 1, dtype=torch.int64, device=device)
    z = idx * s64(0x9E3779B97F4A7C15) + s64(seed)
    z = (z ^ lsr(z, 30)) * s64(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * s64(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    u = lsr(z, 40).to(torch.float64) / float(1 << 24)
    return (u * 2.0 - 1.0).to(torch.float32)


def projection_cpu_run(size, solves, warmup=1):
    """The reference's projection (lexicographic GS/SOR, sequential: one core) on the config-5 input at `size`^2:
    time per solve and the residual it reaches -- the target the GPU solver is held to."""
    import oracle
    from fluid_b200 import presets
    p = presets.projection_stress(size, size)
    f = oracle.New(p.density, size, size, p.h, solver=oracle.SOLVER_EXACT)
    u, v = presets.projection_fields(size + 2, size + 2, 0, size + 2)
    f.set("U", u); f.set("V", v)
    f.edit(p.init); f.edit(p.per_step)
    u0, v0 = f.get("U").copy(), f.get("V").copy()
    before = float(f.MaxDivergence())
    f.project(8, p.dt)
    after = float(f.MaxDivergence())
    times = []
    for k in range(warmup + solves):
        f.set("U", u0); f.set("V", v0)
        t0 = time.perf_counter()
        f.project(8, p.dt)
        if k >= warmup:
            times.append(time.perf_counter() - t0)
    cells = (size + 2) * (size + 2)
    sec = sum(times) / len(times)
    return {"value": cells / sec, "unit": "cell-steps/s", "cores": 1, "kind": "port", "ms_per_step": sec * 1e3,
            "sample": f"config-5 input at {size}^2 (+ring): fill(p,0) + makeIncompressible(8) of the oracle/ C restatement, "
                      f"lexicographic in-place GS/SOR -- sequential in the reference too (fluid.go:188-234), 1 thread; "
                      f"{solves} solves after {warmup} warm-up",
            "max_div_before": before, "max_div_after_8_sweeps": after}


def run_projection(args, rank, world, local):
    """configs[4]: projection only, FIXED global grid (strong scaling), row slabs over the ranks.  One "step" is
    fill(p,0) + makeIncompressible(8) (fluid.go:83, 144-234) over the whole grid, preceded for N > 1 by the
    exchange of 16 ghost lines of U, V with both neighbours."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    import fluid_b200
    from fluid_b200 import _lib as L
    from fluid_b200 import parallel, presets

    pname, width, height, _b, _c, cfg_desc = WORKLOADS[args.workload]
    if args.impl == "reference":
        if rank != 0:
            return 0
        size = min(width, 4096)
        r = projection_cpu_run(size, max(min(args.steps, 5), 1), 1)
        line = {"impl": "reference", "metric": "cell-steps/s", "value": r["value"], "unit": "cell-steps/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "restates": cfg_desc, "step": "one projection (8 sweeps)",
                           "solver": "lexicographic (reference)", "sample_grid": [size + 2, size + 2]},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    solver = {"pressure": fluid_b200.SOLVER_REDBLACK_PRESSURE, "redblack": fluid_b200.SOLVER_REDBLACK,
              "exact": fluid_b200.SOLVER_EXACT}[args.solver]
    preset = presets.projection_stress(width, height)
    ghost = 32 if world > 1 else 0
    if world > 1:
        sim = parallel.SlabFluid(preset.density, width, height, preset.h, solver=solver, device=local, rank=rank,
                                 nranks=world, ghost=ghost, reach=1, transport=args.transport)
        f = sim.f
    else:
        sim = f = fluid_b200.New(preset.density, width, height, preset.h, solver=solver, device=local)
    NX, NY = f.NumX, f.NumY
    a0 = max(0, f.i_lo - ghost)
    lines = min(NX, f.i_hi + ghost) - a0
    # the rank's lines of the SplitMix64 field (U = values 0 .. NX*NY-1 of the stream, V the next NX*NY), generated on
    # the device and parked in pinned host memory: the end-to-end leg uploads them every step
    host = {}
    for name, base in (("U", 0), ("V", NX * NY)):
        buf = torch.empty((lines, NY), dtype=torch.float32, pin_memory=True)
        for l0 in range(0, lines, 2048):
            l1 = min(l0 + 2048, lines)
            buf[l0:l1].copy_(splitmix_uniform_torch(torch, (l1 - l0) * NY, 0x5EED, base + (a0 + l0) * NY, dev).view(l1 - l0, NY))
        host[name] = buf
    torch.cuda.synchronize()

    def upload():
        # fb_upload addresses a GLOBAL [NumX][NumY] array and copies the lines this rank holds: pass the window's origin
        for name in ("U", "V"):
            L.check(f._h, L.lib.fb_upload(f._h, L.FIELD_NAMES[name], C.c_void_p(host[name].data_ptr() - a0 * NY * 4)))

    upload()
    sim.edit(preset.init)        # all fluid, 8x8 obstacle lattice, four walls: SetSolid zeroes the faces of solid cells (Q-16)
    sim.edit(preset.per_step)    # the +-5 sources on three rows
    # keep the PREPARED field in the pinned buffers (faces zeroed, sources set)
    for name in ("U", "V"):
        L.check(f._h, L.lib.fb_download(f._h, L.FIELD_NAMES[name], C.c_void_p(host[name].data_ptr() - a0 * NY * 4)))
    if world > 1:
        upload()                 # fb_download fills owned lines only; ghosts come from the exchange below

    def solve():
        sim.project(8, preset.dt)      # N > 1: SlabFluid.project refreshes the ghost lines first

    div_before = sim.MaxDivergence()
    residuals = {}
    done = 0
    for target in (1, 2, 4, 8):
        while done < target:
            solve(); done += 1
        residuals[str(8 * target)] = sim.MaxDivergence()
    for _ in range(max(args.warmup - 8, 0)):
        solve()
    f.set_option(L.OPT_SOLVE_STATS, 0)
    f.profile(True); f.profile_read()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = f.launch_count()
    barrier()
    f.timer_start()
    for _ in range(args.steps):
        solve()
    ms = f.timer_stop()
    barrier()
    ms = allmax(ms)
    launches = f.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else {}
    phases = f.profile_read()
    f.profile(False)
    cells_total = NX * NY
    value = cells_total * args.steps / (ms * 1e-3)
    peak, peak_src = measured_peak()
    proj_ms = allmax(phases["project"][0] / max(phases["project"][1], 1))
    cells_rank = (f.i_hi - f.i_lo) * NY
    achieved = 24 * cells_rank / (proj_ms * 1e-3) / 1e9
    kname = ("k_rbq_stream" if os.environ.get("FLUIDB200_RBQ_STREAM") else "k_rbq_fused") if args.solver == "pressure" else "k_rb_fused"
    traffic, traffic_src = traffic_for(args.workload, kname)
    roofline = {"bound": "hbm", "kernel": kname, "phase": "project",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "alg_bytes_per_cell": 24.0, "ms_per_launch": proj_ms,
                "share_of_step": proj_ms / (ms / args.steps),
                "pressure_solve": {"alg_bytes_per_cell": 24, "ms": proj_ms, "achieved": achieved, "frac": achieved / peak,
                                   "aggregate_GBps": 24 * cells_total / (ms / args.steps * 1e-3) / 1e9},
                "phases_ms_per_step": {k: v[0] / args.steps for k, v in phases.items() if v[1] > 0}}

    # ---- e2e: the same projection through the C ABI with HOST buffers: per step the rank's lines of U and V go
    # host -> device from pinned memory, the halo exchange and the solve run, and the step's result (max |div|) comes back
    e2e = None
    if not args.no_secondary:
        n_e2e = max(1, min(args.steps, 3))
        def e2e_step():
            upload()
            solve()
            return sim.MaxDivergence()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        e2e_s = allmax(time.perf_counter() - t0)
        e2e = {"value": cells_total * n_e2e / e2e_s, "unit": "cell-steps/s", "steps": n_e2e,
               "h2d_bytes_per_step": int(2 * lines * NY * 4) * world, "d2h_bytes_per_step": 4 * world,
               "ms_per_step": e2e_s / n_e2e * 1e3,
               "what": "per step and rank: fb_upload of U and V (the lines the rank holds) from pinned host memory, halo "
                       "exchange, fill(p,0) + makeIncompressible(8), MaxDivergence() read back"}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        r = projection_cpu_run(min(width, 4096), 3, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "max_div_before", "max_div_after_8_sweeps")}
    if world > 1:
        sim.check_halo()
    if rank == 0:
        line = {
            "metric": "cell-steps/s", "value": value, "unit": "cell-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "restates": cfg_desc, "step": "one projection: fill(p,0) + makeIncompressible(8 iterations)",
                       "grid_total": [NX, NY], "cells_total": cells_total, "grid_per_gpu": [f.i_hi - f.i_lo, NY],
                       "input": "U, V = float32 uniform(-1,1) from SplitMix64(0x5EED) in linear-index order; four walls, 8x8 lattice of "
                                "circular obstacles of radius H/64, +-5 sources every 64th cell on three rows (SURVEY.md 8d)",
                       "parallelism": "single GPU" if world == 1 else
                                      f"row slabs over i, {world} ranks, {ghost} ghost lines of U, V refreshed before every solve "
                                      f"({'peer memory over NVLink' if args.transport == 'peer' else 'NCCL send/recv'})",
                       "solver": {"pressure": "red-black SOR in pressure form, 8 iterations fused in one pass",
                                  "redblack": "red-black SOR on the face velocities, 8 iterations fused",
                                  "exact": "lexicographic GS/SOR (wavefront), 8 sweeps"}[args.solver],
                       "l2": "working set >> 126 MB L2 (inputs larger than L2)" if cells_rank > 8e6 else "working set fits L2",
                       "residual": {"max_div_before": div_before, "max_div_after_iterations": residuals,
                                    "note": "target = the reference's (lexicographic) max|div| after its 8 sweeps on the same "
                                            "input, cpu_baseline.max_div_after_8_sweeps, computed at min(size, 4096)^2"}},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def source_stamp():
    """sha1 of the kernel sources: profiles/*_traffic.json carries the stamp of the build its ncu captures were taken
    from, and `roofline.traffic` is only reported when it matches THIS build."""
    import hashlib
    hh = hashlib.sha1()
    csrc = os.path.join(ROOT, "fluid_b200", "csrc")
    for fn in sorted(os.listdir(csrc)):
        if fn.endswith((".cu", ".cuh")):
            with open(os.path.join(csrc, fn), "rb") as fh:
                hh.update(fh.read())
    return hh.hexdigest()[:16]


def traffic_for(workload, kernel):
    """DRAM bytes per launch of `kernel` from the newest profiles/rNN_traffic.json whose stamp matches this build."""
    prof = os.path.join(ROOT, "profiles")
    try:
        names = sorted(fn for fn in os.listdir(prof) if fn.endswith("_traffic.json"))
    except OSError:
        return None, "no profiles/"
    stamp = source_stamp()
    for fn in reversed(names):
        try:
            with open(os.path.join(prof, fn)) as fh:
                d = json.load(fh)
        except Exception:
            continue
        if d.get("source_stamp") != stamp:
            continue
        v = d.get(workload, {}).get(kernel)
        if v is not None:
            return v, fn
    return None, f"no ncu capture of this build (source stamp {stamp}); see tools/collect_profiles.sh"


def residual_on_state(sim, preset, fluid_b200):
    """The benchmarked solver's parity claim ON THE BENCHMARKED STATE: the projection's input of the next step (the
    per-frame edits applied to the final state of the timed region) goes through the reference's 8 lexicographic sweeps
    (oracle, CPU) and through the GPU solver's 8 iterations; max|div| over the cells the projection updates before and
    after both.  north_star: an order-dependent reference solver is matched by residual, with the iteration count stated."""
    import oracle
    sim.edit(preset.per_step)
    state = {k: sim.get(k) for k in ("U", "V", "M", "p")}
    o = oracle.New(preset.density, preset.width, preset.height, preset.h, solver=oracle.SOLVER_EXACT)
    o.edit(preset.init)
    o.set("U", state["U"]); o.set("V", state["V"])
    before = float(o.MaxDivergence())
    t0 = time.perf_counter()
    o.project(8, preset.dt)
    lex_s = time.perf_counter() - t0
    lex = float(o.MaxDivergence())
    o.close()
    gpu_before = float(sim.MaxDivergence())
    sim.project(8, preset.dt)
    gpu = float(sim.MaxDivergence())
    for k, v in state.items():        # put the state back: later legs continue from it
        sim.set(k, v)
    return {"max_div_before": before, "gpu_before": gpu_before, "gpu_after_8": gpu, "reference_lex_after_8": lex,
            "iterations": 8, "ok": bool(gpu <= lex), "reference_solve_s": lex_s,
            "what": "input = final state of the timed region + the per-frame edits; reference = oracle/ lexicographic GS/SOR, "
                    "8 sweeps (fluid.go:157-234); gpu = the benchmarked solver, 8 iterations; MaxDivergence() of fluid.go:876"}


def slab_parity_check(fluid_b200, parallel, torch, dist, rank, world, local, transport, overlap):
    """N ranks vs one GPU on the same small preset (Karman, BFECC + confinement, pressure-form solver): every field must
    be bit-identical.  Runs inside the driver's own SCALE invocation so that N-process correctness is on record."""
    from fluid_b200 import presets
    width, height, steps = 160 * world, 96, 10
    p = presets.karman(width, height)
    solver = fluid_b200.SOLVER_REDBLACK_PRESSURE
    reach = parallel.reach_for(p.dt, p.h, 8.0)
    sim = parallel.SlabFluid(p.density, width, height, p.h, solver=solver, device=local, rank=rank, nranks=world,
                             ghost=parallel.required_ghost(reach, True, True), reach=reach, transport=transport)
    sim.edit(p.init)
    sim.UseBFECC = True
    sim.Confinement = 0.1
    sim.set_overlap(overlap)
    sim.step(p.dt, steps, p.per_step)
    sim.check_halo()
    got = {k: sim.get(k) for k in ("U", "V", "M", "p")}
    sim.close()
    out = None
    if rank == 0:
        one = fluid_b200.New(p.density, width, height, p.h, solver=solver, device=local)
        one.edit(p.init)
        one.UseBFECC = True
        one.Confinement = 0.1
        one.step(p.dt, steps, p.per_step)
        diffs = {k: float(np.max(np.abs(got[k].astype(np.float64) - one.get(k).astype(np.float64)))) for k in got}
        one.close()
        out = {"ok": all(v == 0.0 for v in diffs.values()), "max_abs_diff": diffs, "ranks": world, "grid": [width + 2, height + 2],
               "steps": steps, "overlapped_exchange": bool(overlap),
               "what": "Karman preset, BFECC + confinement, pressure-form solver: N row slabs (peer-memory halo exchange, as "
                       "benchmarked) vs the single-GPU run of the same grid, U V M p compared bit for bit"}
    if world > 1:
        dist.barrier()
    return out


def run_preset(args, workload, rank, world, local, primary):
    """One preset workload on `world` GPUs (weak scaling over i); returns the JSON line's dict on rank 0."""
    import torch
    import torch.distributed as dist

    import fluid_b200
    from fluid_b200 import _lib as L
    from fluid_b200 import parallel

    pname, width, height, bfecc, conf, cfg_desc = WORKLOADS[workload]
    bpc = step_bytes(bfecc, conf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    solver = {"pressure": fluid_b200.SOLVER_REDBLACK_PRESSURE, "redblack": fluid_b200.SOLVER_REDBLACK,
              "exact": fluid_b200.SOLVER_EXACT}[args.solver]
    preset = build_preset(pname, width, height, bfecc, conf)

    # ---- N > 1: weak scaling -- the grid grows with N along i (row slabs, one per GPU), every rank steps its slab,
    # halos travel through peer memory (fluid_b200.parallel); N == 1 is the plain handle.
    if world > 1:
        if solver == fluid_b200.SOLVER_EXACT:
            raise SystemExit("the lexicographic solver does not decompose into slabs; use --solver pressure")
        preset = build_preset(pname, width * world, height, bfecc, conf)
        reach = parallel.reach_for(preset.dt, preset.h, 8.0)          # jets run at 4; allow 2x
        ghost = parallel.required_ghost(reach, bfecc, conf != 0.0)
        sim = parallel.SlabFluid(preset.density, preset.width, preset.height, preset.h, solver=solver, device=local,
                                 rank=rank, nranks=world, ghost=ghost, reach=reach, transport=args.transport)
        sim.edit(preset.init)
        sim.UseBFECC = bfecc
        sim.Confinement = conf
        sim.adaptive_reach = True       # reach (hence the ghost lines every phase recomputes) follows the all-reduced max |u|
        sim.check_every = 64            # halo check + reach update: a stream drain and an all-reduce each time
        overlap = args.transport == "peer" and not args.no_overlap
        sim.set_overlap(overlap)        # the exchange for step k+1 rides inside step k (FB_OPT_HALO_OVERLAP)
        cells_total = sim.global_cells
        how = ("pulled from the neighbours' CUDA-IPC send buffers over NVLink (flag in peer memory, no collective)"
               if args.transport == "peer" else "NCCL send/recv")
        when = ("overlapped with the step: U, V travel on a second stream during the smoke passes, M during the interior of the last "
                "smoke pass, whose boundary strips are computed first" if overlap else "in front of the step")
        parallelism = (f"row slabs over i, {world} ranks, {ghost} ghost lines allocated (reach for |u| <= 8), the reach used per step "
                       f"re-measured every 64 steps from the all-reduced max |u| x 1.5; 1 halo exchange per step {how}, {when}")
    else:
        sim = make_fluid(fluid_b200, preset, solver, device=local)
        cells_total = sim.NumX * sim.NumY
        parallelism = "single GPU"

    def run_steps(n):
        sim.step(preset.dt, n, preset.per_step)

    enqueue = []       # host time to queue a block (the host must stay ahead of the device: N > 1 steps from Python)

    def timed(n):
        barrier()
        sim.timer_start()
        t0 = time.perf_counter()
        run_steps(n)
        enqueue.append((time.perf_counter() - t0) * 1e3 / n)
        ms = sim.timer_stop()
        barrier()
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    K = args.steps
    run_steps(args.warmup)
    if hasattr(sim, "set_option"):
        sim.set_option(L.OPT_SOLVE_STATS, 0)      # per-iteration residual tracking off in the timed region
    # ---- (1) the quiescent state round 1 reported: the jet has moved W + K of NumX lines, most cells are exact zeros
    quiescent_ms = timed(K)
    # ---- (2) develop the flow: the jet front (CFL 3.33 lines per step at u = 4) crosses one GPU's share of the domain
    preroll = args.preroll
    if preroll < 0:
        preroll = min(int(np.ceil((width + 2) * preset.h / (4.0 * preset.dt))), args.preroll_cap)
    done = 0
    while done < preroll:
        n = min(256, preroll - done)
        run_steps(n)
        done += n
        if world > 1:
            sim.check_halo()
    # ---- (3) the timed region: blocks of EXACTLY K steps, each between barrier + synchronize, CUDA events on the
    # library's stream, max over ranks; enough blocks for >= args.min_timed_steps steps so that the clock sampler sees the
    # run.  The reported step time is the MEDIAN block; every block is listed.
    nblocks = max(1, -(-args.min_timed_steps // K))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = sim.launch_count()
    blocks = [timed(K) for _ in range(nblocks)]
    launches = (sim.launch_count() - launches0) // nblocks
    clk = clocks.stop() if rank == 0 else {}
    ms = float(np.median(blocks))
    value = cells_total * K / (ms * 1e-3)
    # ---- (4) one more block with the per-phase / per-kernel CUDA events on (not part of the headline)
    # Three blocks, per slot the MEDIAN block: one disturbed block (a kernel 60 % slower for 20 steps was seen once) must not
    # pick the "dominant" kernel.
    sim.profile(True)
    sim.profile_read()
    prof_blocks = []
    for _ in range(3):
        t_block = timed(K)
        prof_blocks.append((t_block, sim.profile_read()))
    sim.profile(False)
    prof_ms = float(np.median([b[0] for b in prof_blocks]))
    phases = {}
    for k in prof_blocks[0][1]:
        rows = [b[1].get(k, (0.0, 0)) for b in prof_blocks]
        calls = rows[0][1]
        phases[k] = (float(np.median([r[0] for r in rows])), calls) if all(r[1] == calls for r in rows) else rows[0]

    # ---- roofline of the dominant kernel (largest share of the step among single kernels) and of the whole step
    peak, peak_src = measured_peak()
    cells_rank = cells_total // world
    kernel_alg_bytes = {"k_pressure_solve": 24, "k_advect_velocity_full": 20, "k_bfecc_velocity_correct": 28,
                        "k_advect_smoke_full": 20, "k_bfecc_smoke_correct": 24, "k_confine_turbulence": 20}
    kernel_real = {"k_pressure_solve": {"pressure": "k_rbq_stream" if os.environ.get("FLUIDB200_RBQ_STREAM") else "k_rbq_fused",
                                        "redblack": "k_rb_fused", "exact": "k_gs_wavefront"}[args.solver],
                   "k_confine_turbulence": ("k_confine_tile" if os.environ.get("FLUIDB200_CONFINE_TILE") else
                                            "k_confine_turbulence" if os.environ.get("FLUIDB200_CONFINE_IEEE") else "k_confine_fast")}
    if not os.environ.get("FLUIDB200_ADV_FULL"):     # the slots are named after the round-1 kernels; what runs is the tile form
        kernel_real.update({"k_advect_velocity_full": "k_advect_velocity_tile", "k_bfecc_velocity_correct": "k_bfecc_velocity_tile",
                            "k_advect_smoke_full": "k_advect_smoke_tile", "k_bfecc_smoke_correct": "k_bfecc_smoke_tile"})
    kern = {}
    for k in L.PROF_KERNELS:
        tot, calls = phases.get(k, (0.0, 0))
        if calls > 0 and tot > 0:
            per = tot / calls
            ach = kernel_alg_bytes[k] * cells_rank / (per * 1e-3) / 1e9
            kern[kernel_real.get(k, k)] = {"ms_per_launch": per, "launches_per_step": calls / K, "ms_per_step": tot / K,
                                           "alg_bytes_per_cell": kernel_alg_bytes[k], "achieved": ach, "frac": ach / peak,
                                           "share_of_step": tot / max(prof_ms, 1e-9)}
    dom = max(kern, key=lambda k: kern[k]["ms_per_step"]) if kern else None
    proj = kern.get(kernel_real["k_pressure_solve"], {})
    phase_alg_bytes = {"project": 24, "confinement": 20, "turbulence": 0 if conf != 0.0 else 20,
                       "advect_velocity": 68 if bfecc else 20, "advect_smoke": 64 if bfecc else 20}
    step_gbs = bpc * cells_rank / (ms / K * 1e-3) / 1e9
    roofline = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src}
    if dom:
        d = kern[dom]
        traffic, traffic_src = traffic_for(workload, dom)
        roofline.update({"kernel": dom, "achieved": d["achieved"], "frac": d["frac"], "traffic": traffic, "traffic_source": traffic_src,
                         "alg_bytes_per_cell": d["alg_bytes_per_cell"], "alg_bytes_per_launch": d["alg_bytes_per_cell"] * cells_rank,
                         "ms_per_launch": d["ms_per_launch"], "launches_per_step": d["launches_per_step"],
                         "share_of_step": d["share_of_step"],
                         "dominant_by": "largest total time per step among single kernels (CUDA events around every launch; median of 3 profiled blocks)"})
    roofline.update({
        "step": {"alg_bytes_per_cell_step": bpc, "achieved": step_gbs, "frac": step_gbs / peak},
        "pressure_solve": {"kernel": kernel_real["k_pressure_solve"], "alg_bytes_per_cell": 24, "ms": proj.get("ms_per_launch"),
                           "achieved": proj.get("achieved"), "frac": proj.get("frac")},
        "kernels": kern,
        "phases_ms_per_step": {k: v[0] / K for k, v in phases.items() if v[1] > 0 and k not in L.PROF_KERNELS},
        "phases_frac_of_peak": {k: (phase_alg_bytes[k] * cells_rank / (v[0] / K * 1e-3) / 1e9 / peak)
                                for k, v in phases.items() if v[1] > 0 and v[0] > 0 and phase_alg_bytes.get(k, 0)},
        "profiled_block_ms_per_step": prof_ms / K})

    line = {
        "metric": "cell-steps/s", "value": value, "unit": "cell-steps/s", "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "restates": cfg_desc, "grid_per_gpu": [width + 2, height + 2],
                   "grid_total": [width * world + 2, height + 2], "cells_total": cells_total, "preset": pname, "bfecc": bfecc,
                   "confinement": conf, "turbulence": 0.02, "dt": preset.dt, "parallelism": parallelism,
                   "solver": {"pressure": "red-black SOR in pressure form, 8 iterations fused in one pass (k_rbq_fused), "
                                          "reference omega schedule with damped close (1.0, 0.5)",
                              "redblack": "red-black SOR on the face velocities, 8 iterations fused, "
                                          "reference omega schedule with damped close (1.0, 0.5)",
                              "exact": "lexicographic GS/SOR (bit-exact wavefront), 8 sweeps"}[args.solver],
                   "l2": "working set >> 126 MB L2 (inputs larger than L2)" if cells_total // world > 8e6
                         else "working set fits L2: HBM fraction not meaningful",
                   "state": f"developed flow: {preroll} untimed pre-roll steps after the {args.warmup} warm-up steps (the jet front "
                            f"crosses one GPU's {width + 2} lines at 3.33 lines per step); the quiescent start-up state round 1 "
                            "timed is reported beside it (`quiescent`)",
                   "timing": f"{nblocks} blocks of exactly {K} steps, CUDA events on the library's stream between barrier + "
                             "synchronize, max over ranks; value = median block; profiling events off"},
        "timed_blocks_ms": blocks, "quiescent": {"ms_per_step": quiescent_ms / K, "value": cells_total * K / (quiescent_ms * 1e-3),
                                                "what": f"the same {K} steps timed right after the warm-up, before the pre-roll"},
        "roofline": roofline, "gpu_launches": int(launches), "clocks": clk,
        "host_enqueue_ms_per_step": float(np.median(enqueue)) if enqueue else None,
    }
    if not primary:
        if world > 1:
            sim.check_halo()
        sim.close()
        return line if rank == 0 else None

    # ---- e2e: the frame loop through the public API with host buffers
    e2e = None
    secondary = {}
    if (not args.no_secondary) and world == 1:
        import ctypes as C
        per = fluid_b200.edits.pack(preset.per_step)
        h2d = int(per.nbytes)
        mirror = sim._mirror(L.M)            # pinned host memory owned by the library
        mn, mx = C.c_float(), C.c_float()

        def frame():
            sim.step(preset.dt, 1, per)      # edit commands host->device + Simulate
            L.check(sim._h, L.lib.fb_view(sim._h, L.VIEW_SMOKE, mirror.ctypes.data, C.byref(mn), C.byref(mx)))

        for _ in range(3):
            frame()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            frame()
        barrier()
        sync_s = time.perf_counter() - t0

        # the same loop pipelined: frame k's view travels to the host while Simulate k+1 runs
        # (fb_view_begin / fb_view_end, two pinned frame buffers -- the renderer reads one while
        # the other fills).  Every step's view still reaches host memory inside the timed region.
        frames = [torch.empty((sim.NumX * sim.NumY,), dtype=torch.float32, pin_memory=True) for _ in range(2)]

        def pipelined(n):
            sim.step(preset.dt, 1, per)
            sim.view_begin(L.VIEW_SMOKE, frames[0].data_ptr())
            for k in range(1, n):
                sim.step(preset.dt, 1, per)
                sim.view_end()
                sim.view_begin(L.VIEW_SMOKE, frames[k & 1].data_ptr())
            return sim.view_end()

        pipelined(3)
        barrier()
        t0 = time.perf_counter()
        pipelined(K)
        barrier()
        full_s = time.perf_counter() - t0

        # the display loop: the same pipelined frame loop delivering what a window can show -- the Smoke() view decimated
        # to <= 1024 cells a side and quantised to one byte per cell on the device (fb_view_u8_begin), min / max included
        stride = max(1, -(-max(sim.NumX, sim.NumY) // 1024))
        dn_i, dn_j = -(-sim.NumX // stride), -(-sim.NumY // stride)
        shots = [torch.empty((dn_i * dn_j,), dtype=torch.uint8, pin_memory=True) for _ in range(2)]

        def display(n):
            sim.step(preset.dt, 1, per)
            sim.view_u8_begin(L.VIEW_SMOKE, stride, shots[0].data_ptr())
            for k in range(1, n):
                sim.step(preset.dt, 1, per)
                sim.view_end()
                sim.view_u8_begin(L.VIEW_SMOKE, stride, shots[k & 1].data_ptr())
            return sim.view_end()

        display(3)
        barrier()
        t0 = time.perf_counter()
        display(K)
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e = {"value": cells_total * K / e2e_s, "unit": "cell-steps/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(dn_i * dn_j + 8), "ms_per_step": e2e_s / K * 1e3,
               "what": f"per step: edit commands H2D, Simulate, the Smoke() view as a display frame D2H into pinned memory -- every "
                       f"{stride}th cell of every {stride}th line as one byte (quantised on the device against the full field's min / max, "
                       f"which travel too): {dn_i}x{dn_j} bytes; pipelined frame loop (fb_view_u8_begin / fb_view_end): the frame of step k "
                       "travels while step k+1 computes.  The FULL float field per step (round 1's definition) is `full_field_loop`",
               "full_field_loop": {"value": cells_total * K / full_s, "ms_per_step": full_s / K * 1e3,
                                   "d2h_bytes_per_step": int(sim.NumX * sim.NumY * 4 + 8),
                                   "what": "same loop with fb_view_begin / fb_view_end: the whole float32 Smoke() field every step (PCIe-bound)"},
               "blocking_loop": {"value": cells_total * K / sync_s, "ms_per_step": sync_s / K * 1e3,
                                 "what": "full float field with the blocking fb_view after every step"}}

        # ---- the frame loop's neighbours of the hot path, reported beside e2e:
        # the same pipelined loop delivering PIXELS (Draw's colormap + solid overlay on the device, fb_render_begin/end)
        # instead of the float field, and advectParticles for 1 M tracers (4 bilinear samples each).
        images = [torch.empty((sim.NumX * sim.NumY * 4,), dtype=torch.uint8, pin_memory=True) for _ in range(2)]

        def render_loop(n):
            sim.step(preset.dt, 1, per)
            sim.render_begin(L.VIEW_SMOKE, images[0].data_ptr())
            for k in range(1, n):
                sim.step(preset.dt, 1, per)
                sim.render_end()
                sim.render_begin(L.VIEW_SMOKE, images[k & 1].data_ptr())
            return sim.render_end()

        render_loop(3)
        barrier()
        t0 = time.perf_counter()
        render_loop(K)
        barrier()
        render_s = time.perf_counter() - t0
        nparts = 1 << 20
        rng = np.random.default_rng(5)
        parts = np.zeros(nparts, dtype=fluid_b200.PARTICLE_DTYPE)
        parts["x"] = rng.uniform(preset.h, sim.NumX * preset.h, nparts).astype(np.float32)
        parts["y"] = rng.uniform(preset.h, sim.NumY * preset.h, nparts).astype(np.float32)
        parts["max_age"] = 1e9
        sim.AdvectParticles(parts[:1024], preset.dt)
        t0 = time.perf_counter()
        alive = sim.AdvectParticles(parts, preset.dt)
        parts_s = time.perf_counter() - t0
        e2e["frame_loop_neighbours"] = {
            "render_loop": {"value": cells_total * K / render_s, "ms_per_step": render_s / K * 1e3,
                            "d2h_bytes_per_step": int(sim.NumX * sim.NumY * 4 + 8),
                            "what": "the pipelined frame loop with fb_render_begin/end: the RGBA image of Draw "
                                    "(colormap + solid overlay computed on the device) lands in pinned memory every step"},
            "advect_particles": {"particles": nparts, "alive": int(len(alive)), "ms": parts_s * 1e3,
                                 "particles_per_s": nparts / parts_s,
                                 "what": "fb_advect_particles, host array in / out (H2D + RK2 midpoint on the device + D2H + "
                                         "stable filter on the host)"}}

        # ---- the other solver on the same workload (reported, not the headline)
        other = fluid_b200.SOLVER_EXACT if solver != fluid_b200.SOLVER_EXACT else fluid_b200.SOLVER_REDBLACK_PRESSURE
        sim.Solver = other
        run_steps(2)
        n2 = max(K // 3, 3)
        ms2 = timed(n2)
        secondary = {"solver": "exact" if other == fluid_b200.SOLVER_EXACT else "pressure",
                     "value": cells_total * n2 / (ms2 * 1e-3), "ms_per_step": ms2 / n2,
                     "note": "exact = lexicographic wavefront, bit-identical to the reference restatement"}
        sim.Solver = solver

    if (not args.no_secondary) and world > 1:
        # frame loop per rank: edit commands H2D, slab step (with its halo exchange), the rank's
        # share of the Smoke() view D2H into pinned memory, min/max all-reduced
        per = fluid_b200.edits.pack(preset.per_step)
        # pinned buffers for this rank's slab only; fb_view addresses them as a window of the global array
        slab = torch.empty(((sim.i_hi - sim.i_lo) * sim.NumY,), dtype=torch.float32, pin_memory=True)
        slabs = [slab, torch.empty_like(slab).pin_memory()]
        bases = [b.data_ptr() - sim.i_lo * sim.NumY * 4 for b in slabs]

        def reduce_minmax(mm):
            t = torch.tensor([-mm[0], mm[1]], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)

        def pipelined(n):
            sim.step(preset.dt, 1, per)
            sim.f.view_begin(L.VIEW_SMOKE, bases[0])
            for k in range(1, n):
                sim.step(preset.dt, 1, per)
                reduce_minmax(sim.f.view_end())
                sim.f.view_begin(L.VIEW_SMOKE, bases[k & 1])
            reduce_minmax(sim.f.view_end())

        def loop_time(fn):
            fn(3)
            barrier()
            t0 = time.perf_counter()
            fn(K)
            barrier()
            t = torch.tensor([time.perf_counter() - t0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        full_s = loop_time(pipelined)
        # the display loop (see the single-GPU leg): each rank's lines decimated / quantised on its device
        stride = max(1, -(-max(width + 2, height + 2) // 1024))
        nlines = -(-sim.i_hi // stride) - (-(-sim.i_lo // stride))
        ncols = -(-sim.NumY // stride)
        shots = [torch.empty((max(nlines, 1) * ncols,), dtype=torch.uint8, pin_memory=True) for _ in range(2)]

        def display(n):
            sim.step(preset.dt, 1, per)
            sim.f.view_u8_begin(L.VIEW_SMOKE, stride, shots[0].data_ptr())
            for k in range(1, n):
                sim.step(preset.dt, 1, per)
                reduce_minmax(sim.f.view_end())
                sim.f.view_u8_begin(L.VIEW_SMOKE, stride, shots[k & 1].data_ptr())
            reduce_minmax(sim.f.view_end())

        e2e_s = loop_time(display)
        e2e = {"value": cells_total * K / e2e_s, "unit": "cell-steps/s", "h2d_bytes_per_step": int(per.nbytes) * world,
               "d2h_bytes_per_step": int(nlines * ncols + 8) * world, "ms_per_step": e2e_s / K * 1e3,
               "what": f"per step and rank: edit commands H2D, slab Simulate + halo exchange, the rank's lines of the Smoke() view as a "
                       f"display frame (every {stride}th cell of every {stride}th line, one byte, quantised on the device) D2H into pinned "
                       "memory, min / max all-reduce; pipelined (fb_view_u8_begin / fb_view_end).  Full float field: `full_field_loop`",
               "full_field_loop": {"value": cells_total * K / full_s, "ms_per_step": full_s / K * 1e3,
                                   "d2h_bytes_per_step": int(cells_total * 4 + 8 * world),
                                   "what": "same loop with fb_view_begin / fb_view_end: every rank's float32 slab every step (one host's PCIe)"}}

    # ---- the benchmarked solver against the reference's on the benchmarked state, the CPU baseline at the SAME grid
    residual, cpu = None, None
    if world > 1:
        sim.check_halo()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        big = cells_total > 2.0e7          # the oracle needs ~0.4 us per cell-step: 16386^2 would take minutes per step
        if not big:
            residual = residual_on_state(sim, preset, fluid_b200)
        r = cpu_reference_run(workload, 2, 1, sample_size=(4096, 4096) if big else None, budget_s=60.0)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "ms_per_step", "grid", "same_grid_as_gpu_arm")}
        cpu["go_toolchain"] = go_probe()
        cpu["config1_anchor"] = cpu_anchor_config1()
    sim.close()
    line.update({"cpu_baseline": cpu, "e2e": e2e, "other_solver": secondary})
    line["config"]["residual"] = residual
    return line if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="karman4096", choices=sorted(WORKLOADS))
    ap.add_argument("--solver", default="pressure", choices=["pressure", "redblack", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"], help="halo exchange between slabs (N > 1)")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: exchange halos in front of every step instead of inside it")
    ap.add_argument("--no-secondary", action="store_true", help="skip the e2e, other-solver, parity-check and config[3] legs")
    ap.add_argument("--preroll", type=int, default=-1, help="untimed steps that develop the flow (-1: the jet front crosses one GPU's lines)")
    ap.add_argument("--preroll-cap", type=int, default=1500)
    ap.add_argument("--min-timed-steps", type=int, default=200, help="timed blocks of --steps steps are repeated up to this many steps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank, world, local = dist_env()
    pname, width, height, bfecc, conf, cfg_desc = WORKLOADS[args.workload]
    if pname == "projection":
        return run_projection(args, rank, world, local)

    if args.impl == "reference":
        # rank 0 alone runs the CPU implementation of the SAME preset at the SAME grid; other ranks exit without work.
        # Nothing of the product is loaded here: presets / edit lists are pure Python, the library opens on first use.
        if rank != 0:
            return 0
        r = cpu_reference_run(args.workload, args.steps, args.warmup)
        line = {
            "impl": "reference", "metric": "cell-steps/s", "value": r["value"], "unit": "cell-steps/s",
            "n_gpus": args.gpus, "steps": r["steps_run"], "warmup": r["warmup_run"], "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "restates": cfg_desc, "grid_per_gpu": r["grid"], "preset": pname, "bfecc": bfecc,
                       "confinement": conf, "solver": "lexicographic GS/SOR, 8 sweeps (the reference's)",
                       "asked_steps_warmup": list(r["asked"]), "go_toolchain": go_probe(),
                       "what": "oracle/ (C restatement of the Go reference, parallelRange chunking on all host threads) on the same "
                               "preset and grid as the GPU arm; one rank's grid whatever --gpus is"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    line = run_preset(args, args.workload, rank, world, local, primary=True)
    if not args.no_secondary:
        if world > 1:
            import fluid_b200
            from fluid_b200 import parallel
            pc = slab_parity_check(fluid_b200, parallel, torch, dist, rank, world, local, args.transport,
                                   args.transport == "peer" and not args.no_overlap)
            if rank == 0:
                line["parity_check"] = pc
        if args.workload == "karman4096":
            # BASELINE configs[3] (jet 16384^2 per GPU, weak scaling) inside the driver's default invocation
            import copy
            a2 = copy.copy(args)
            a2.min_timed_steps = args.steps
            a2.preroll, a2.preroll_cap = -1, 600
            extra = run_preset(a2, "jet16384", rank, world, local, primary=False)
            if rank == 0:
                line["config3_jet16384"] = {k: extra[k] for k in ("value", "unit", "n_gpus", "steps", "ms_per_step", "config", "quiescent",
                                                                   "timed_blocks_ms", "roofline")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
