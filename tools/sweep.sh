#!/bin/bash
# run bench.py (value only) for every tuning build in gpurun_in/ — development tool
for w in "$@"; do
  for so in default gpurun_in/libkdpt_*.so; do
    if [ "$so" = default ]; then unset KDPT_LIB; else export KDPT_LIB=$PWD/$so; fi
    python bench.py --workload $w --no-cpu-baseline --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', '$so', 'Mrays/s %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.1f'%d['e2e']['value'])"
  done
done
