#!/bin/bash
# Round-2 experiment, prepared at the end of round 1 (no GPU minutes were left to run it).
#
# Finding (profiles/README.md section 2, "closed loop"): in the ncu source view of k_rbq_fused every role polls
# ~18 times per line -- the sweeps on their predecessor, the WRITER on the last sweep and the LOADER on a free ring
# slot -- while the TMA "full" barriers are hit on the first poll.  A line stays in the shared-memory ring for
# ~26-28 line periods (4 loader + 8 iterations x 2 lines of lag + 4-5 writer) and the ring has 28 slots: the loop
# has no slack, so any hand-off latency stalls everybody.  24 slots cost +17 % (0.181 -> 0.212 ms).  More slots
# need shared memory; the writer's staging ring can give 16 KB back (one line in flight per writer warp).
#
# Second idea on the same finding: 256-column slots (RQ_WL=256, one warp per stage).  Half the bytes per slot, so either
# TWO CTAs per SM with 28 slots each (64 registers; when one pipeline stalls the other issues: wl256x2) or one CTA
# with a 56-slot ring (wl256deep).  Halo overhead 1.30 instead of 1.26.  The default build (RQ_WL=512) is unchanged:
# its SASS was compared instruction for instruction when these knobs were added.
#
# Third idea: the loader's and the writer's four warps are line-interleaved, so a line spends four line periods in each
# role.  With RQ_LSPLIT / RQ_WSPLIT = 2 or 4 warps of a role on ONE line (half / a quarter of the column groups each,
# hand-off count 32 x split like the two warps of a sweep stage) that is two / one period each: up to ~6 line periods
# less residency, the same as 6 more slots, for no shared memory (but 2-4x the hand-off polls in those roles).
# 32 and more slots switch the hand-off rings to 128 mbarriers per role (RQ_RING > 2 * RQ_NL).
#
#   here:    tools/r02_rbq_ring.sh build
#   gpurun:  tools/r02_rbq_ring.sh run        (prints ms for 1 and 8 iterations and a hash of U, V, p per variant:
#                                              equal hashes <=> bit-identical results; then run pytest -m gpu with
#                                              FLUIDB200_LIB pointing at the winner before adopting it)
set -e
cd "$(dirname "$0")/.."
case "$1" in
build)
  tools/variants.sh base "" \
                    nl32w4 "-DRQ_NL=32 -DRQ_WSTG=4" \
                    nl34w4 "-DRQ_NL=34 -DRQ_WSTG=4" \
                    nl36w4 "-DRQ_NL=36 -DRQ_WSTG=4" \
                    nl28w4 "-DRQ_WSTG=4" \
                    wl256x2 "-DRQ_WL=256 -DRQ_SPLIT=1 -DRQ_MINB=2" \
                    wl256deep "-DRQ_WL=256 -DRQ_SPLIT=1 -DRQ_NL=56" \
                    ls2ws2 "-DRQ_LSPLIT=2 -DRQ_WSPLIT=2" \
                    ls4ws4 "-DRQ_LSPLIT=4 -DRQ_WSPLIT=4" \
                    nl32w4s2 "-DRQ_NL=32 -DRQ_WSTG=4 -DRQ_LSPLIT=2 -DRQ_WSPLIT=2" ;;
run)
  for v in base nl28w4 nl32w4 nl34w4 nl36w4 wl256x2 wl256deep ls2ws2 ls4ws4 nl32w4s2; do
    echo -n "$v: "; FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so timeout 40 python tools/rbq_iters.py 1 8 2>&1 | tail -1
  done ;;
*) echo "usage: $0 build|run"; exit 2 ;;
esac
