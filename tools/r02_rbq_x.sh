#!/bin/bash
# round 2: the solve kernel with parts switched off (FLUIDB200_RBQ_X: 1 sweeps, 2 writer I/O, 4 TMA, 8 / 16 loader reads / stores): timing only
for x in 0 1 2 3 4 5 7 8 24 25 27 31; do
  FLUIDB200_RBQ_X=$x timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-120
done
