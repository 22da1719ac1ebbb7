#!/bin/bash
# round 2, call 6 (ON the GPU box): lean v1 variants (occupancy, refill threshold, leaf unroll, staged top levels, SAH constants)
O=gpurun_out/r2; mkdir -p $O
run() {  # label
  echo "== $1"
  python tools/microbench.py --modes fullstack 2>&1 | grep stack
  python tools/microbench.py --modes fullstack --soup 1000000 2>&1 | grep stack
  for w in dragon1m dragon5; do
    timeout 400 python bench.py --workload $w --no-cpu-baseline --no-extras --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w %.1f Mrays/s, %.3f ms/step, check %s'%(d['value'],d['ms_per_step'],d['check'].get('ok')))"
  done
}
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/t6.log 2>&1; tail -4 $O/t6.log
run default
for n in b6 r4 r8 u2; do export KDPT_LIB=$PWD/gpurun_in/libkdpt_$n.so; run $n; done
for L in 8 10 11; do export KDPT_LIB=$PWD/gpurun_in/libkdpt_top$L.so; export KDPT_TOP_LEVELS=$L; run top$L; done
unset KDPT_LIB; unset KDPT_TOP_LEVELS
for ct in 2 4; do for lf in 4 8; do export KDPT_REFINE_CT=$ct; export KDPT_REFINE_LEAF=$lf; run "ct$ct leaf$lf"; done; done
