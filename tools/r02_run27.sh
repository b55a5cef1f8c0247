#!/bin/bash
# round 2, GPU run 27: the next step's hand-off probed during the arithmetic (RQ_EARLYPROBE): timing, determinism, parity, timeline
export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_probe.so
echo -n "probe: "; timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-130
timeout 120 python tools/rbq_race_hunt.py 100 2>&1 | tail -2 | sed "s/^/probe: /"
timeout 900 python -m pytest tests -m gpu -x -q -k "many_chunks or pressure_form or slab or fused_path or step_local or projection or karman_4098" 2>&1 | tail -1 | sed "s/^/probe pytest: /"
export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_probet.so
export FLUIDB200_RBQ_TRACE=/tmp/rbq_trace.bin
timeout 120 python tools/rbq_trace.py 8 2>&1 | tail -11
