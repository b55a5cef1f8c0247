#!/bin/bash
# round 2, GPU run 14: sanity after the per-CTA debug pointer (static shared memory) + solve-kernel role switches
set -x
O=gpurun_out/r02_run14; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1; tail -3 $O/pytest.txt
bash tools/r02_rbq_x.sh 2>&1 | tee $O/rbq_x.txt
