#!/bin/bash
# round 2, GPU run 20: the full GPU suite five more times (rate of the one-off many-chunks mismatch of run 14) + smoke
O=gpurun_out/r02_run20; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for k in 1 2 3; do
  timeout 900 python -m pytest tests -m gpu -q > $O/pytest_$k.txt 2>&1; tail -1 $O/pytest_$k.txt | sed "s/^/full suite $k: /"; grep -A6 "AssertionError" $O/pytest_$k.txt | head -12
done
