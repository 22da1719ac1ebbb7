#!/bin/bash
# run bench.py (value only) for the lane-queue policies of the bounce kernel (0 direct, 1 queued, 2 kernel decides) — development tool
for w in "$@"; do
  for pol in 0 1 2; do
    KDPT_QUEUE=$pol python bench.py --workload $w --no-cpu-baseline --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', 'KDPT_QUEUE=$pol', 'Mrays/s %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.1f'%d['e2e']['value'])"
  done
done
