#!/bin/bash
# round 2, call 5 (ON the GPU box): lean v2 (one unit of work per round) — parity, then variants
O=gpurun_out/r2; mkdir -p $O
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/t5.log 2>&1; tail -4 $O/t5.log
for so in default gpurun_in/libkdpt_*.so; do
  if [ "$so" = default ]; then unset KDPT_LIB; else export KDPT_LIB=$PWD/$so; fi
  echo "== $so"
  python tools/microbench.py --modes fullstack,shortstack 2>&1 | grep stack
  python tools/microbench.py --modes fullstack --cls outside 2>&1 | grep stack
  python tools/microbench.py --modes fullstack --soup 1000000 2>&1 | grep stack
  for w in dragon1m dragon5 bunny; do
    timeout 400 python bench.py --workload $w --no-cpu-baseline --no-extras --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w %.1f Mrays/s, %.3f ms/step, e2e %.1f, check %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['check'].get('ok')))"
  done
done
unset KDPT_LIB
python bench.py --workload soup10m --steps 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('soup10m %.1f Mrays/s e2e %.1f check %s'%(d['value'],d['e2e']['value'],d['check'].get('ok')))"
M="python tools/microbench.py --modes fullstack --reps 2"
$M > $O/mb_lean2_d5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:traceLean -s 1 -c 1 -o $O/prof_lean2_d5 -f $M > $O/ncu_lean2_d5.log 2>&1
