"""Time ONE multigrid V-cycle (fluid.go:560-599) on the config-5 scene at SIZE^2, beside the
single-grid 8-sweep projection of the same solver, with CUDA events through fb_timer_*.  The prepared
field is uploaded again before every repetition (untimed): the reference's cycle amplifies the field, so
repeated cycles on the same handle would overflow.  Prints one JSON line per solver.

    python tools/mg_measure.py [--size 4096] [--reps 5] [--solvers 2,0]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--solvers", default="2,0")
    ap.add_argument("--no-single-grid", action="store_true")
    args = ap.parse_args()
    import fluid_b200
    from fluid_b200 import presets
    preset = presets.projection_stress(args.size, args.size)
    NX = NY = args.size + 2
    U0, V0 = presets.projection_fields(NX, NY, 0, NX)
    for solver in [int(s) for s in args.solvers.split(",")]:
        f = fluid_b200.New(preset.density, args.size, args.size, preset.h, solver=solver)
        f.set("U", U0); f.set("V", V0)
        f.edit(preset.init); f.edit(preset.per_step)
        U, V = f.get("U"), f.get("V")
        out = {"size": args.size, "solver": {0: "exact", 1: "redblack", 2: "pressure"}[solver], "cells": NX * NY,
               "max_div_before": f.MaxDivergence()}
        for label, mg, iters in (("vcycle", True, 1), ("single_grid_8", False, 8)):
            if not mg and args.no_single_grid:
                continue
            f.UseMultigrid = mg
            ms = []
            for rep in range(args.reps + 1):
                f.set("U", U); f.set("V", V)
                n0 = f.launch_count()
                f.timer_start()
                f.project(iters, preset.dt)
                t = f.timer_stop()
                if rep:
                    ms.append(t)
                launches = f.launch_count() - n0
            out[label] = {"ms": float(np.median(ms)), "ms_all": [round(x, 4) for x in ms], "launches": int(launches),
                          "max_div_after": f.MaxDivergence()}
        print(json.dumps(out), flush=True)
        f.close()


if __name__ == "__main__":
    main()
