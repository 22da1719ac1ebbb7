#!/bin/bash
# round 2, call 3 (ON the GPU box): lean kernel — parity tests, then microbenchmarks against the packet kernels
O=gpurun_out/r2; mkdir -p $O
(time timeout 900 python -m pytest tests/test_trace_gpu.py tests/test_large_gpu.py -x -q) > $O/t3.log 2>&1; tail -5 $O/t3.log
for cfg in "1 4" "0 4" "0 0" "1 8" "1 2"; do set -- $cfg
  echo "== KDPT_LEAN=$1 KDPT_REFINE_LEAF=$2"
  KDPT_LEAN=$1 KDPT_REFINE_LEAF=$2 python tools/microbench.py --modes fullstack,shortstack 2>&1 | grep stack
  KDPT_LEAN=$1 KDPT_REFINE_LEAF=$2 python tools/microbench.py --modes fullstack,shortstack --cls outside 2>&1 | grep stack
  KDPT_LEAN=$1 KDPT_REFINE_LEAF=$2 python tools/microbench.py --modes fullstack --soup 1000000 2>&1 | grep stack
done
for cfg in "1 4" "0 0" "1 8"; do set -- $cfg
  echo "== soup10m KDPT_LEAN=$1 KDPT_REFINE_LEAF=$2"
  KDPT_LEAN=$1 KDPT_REFINE_LEAF=$2 python bench.py --workload soup10m --steps 5 --no-cpu-baseline 2>$O/s10_$1_$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f Mrays/s e2e %.1f check %s create %.2fs'%(d['value'],d['e2e']['value'],d['check'],d['config']['create_s']))"
done
for b in 6 10 12; do echo "== lean blocks $b"; KDPT_LEAN_BLOCKS=$b python tools/microbench.py --modes fullstack 2>&1 | grep stack; KDPT_LEAN_BLOCKS=$b python tools/microbench.py --modes fullstack --soup 1000000 2>&1 | grep stack; done
