#!/bin/bash
# traversal-only microbench for every tuning build in gpurun_in/ — development tool
for so in default gpurun_in/libkdpt_*.so; do
  if [ "$so" = default ]; then unset KDPT_LIB; else [ -f "$so" ] || continue; export KDPT_LIB=$PWD/$so; fi
  echo "== $so $KDPT_BLOCK_TICKETS"
  python tools/microbench.py --modes fullstack,shortstack 2>&1 | tail -2
  python tools/microbench.py --modes fullstack --soup 1000000 --rays 4194304 --reps 3 2>&1 | tail -1
  python tools/microbench.py --modes fullstack --soup 10000000 --rays 16777216 --reps 3 2>&1 | tail -1
done
