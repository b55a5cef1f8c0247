#!/bin/bash
# round 2, GPU run 16: does the many-chunks mismatch of run 14 come back in the full suite?  + initcheck of that test
O=gpurun_out/r02_run16; mkdir -p $O
for k in 1 2 3; do
  timeout 900 python -m pytest tests -m gpu -q > $O/pytest_$k.txt 2>&1; tail -1 $O/pytest_$k.txt | sed "s/^/full suite $k: /"; grep FAILED $O/pytest_$k.txt | head -5
done
timeout 900 compute-sanitizer --tool initcheck --print-limit 20 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "many_chunks" > $O/initcheck.txt 2>&1
grep -c "Uninitialized" $O/initcheck.txt; grep -A12 "Uninitialized" $O/initcheck.txt | head -60; tail -3 $O/initcheck.txt
