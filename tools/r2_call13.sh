#!/bin/bash
# round 2, call 13 (ON the GPU box): refined leaves as indices into the shared hot records (KDPT_SHARE_HOT=1, new default) vs own copies
O=gpurun_out/r2; mkdir -p $O
(time timeout 1800 python -m pytest tests/test_trace_gpu.py tests/test_large_gpu.py tests/test_render_gpu.py -x -q) > $O/t13.log 2>&1; grep -E "passed|failed" $O/t13.log
run() {
  echo "== $1"
  python tools/microbench.py --modes fullstack 2>&1 | grep stack
  python tools/microbench.py --modes fullstack --soup 1000000 2>&1 | grep stack
  for w in dragon1m dragon5 bunny; do
    timeout 400 python bench.py --workload $w --no-cpu-baseline --no-extras --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w %.1f Mrays/s, %.3f ms/step, check %s, tree bytes %d'%(d['value'],d['ms_per_step'],d['check'].get('ok'),d['config']['tree']['device_bytes']))"
  done
  python bench.py --workload soup10m --steps 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('soup10m %.1f Mrays/s e2e %.1f check %s create %.2f'%(d['value'],d['e2e']['value'],d['check'].get('ok'),d['config']['create_s']))"
}
KDPT_SHARE_HOT=1 run share1
KDPT_SHARE_HOT=0 run share0
export KDPT_SHARE_HOT=1
KDPT_REFINE_LEAF=4 KDPT_REFINE_CT=1 run "share1 leaf4 ct1"
KDPT_REFINE_LEAF=4 KDPT_REFINE_CT=2 run "share1 leaf4 ct2"
KDPT_REFINE_LEAF=8 KDPT_REFINE_CT=1 run "share1 leaf8 ct1"
