#!/bin/bash
# round 2, call 8 (ON the GPU box): capped while-while rounds (lean v3) vs v1; new modes' tests
O=gpurun_out/r2; mkdir -p $O
(time timeout 1800 python -m pytest tests/test_trace_gpu.py tests/test_render_gpu.py -x -q) > $O/t8.log 2>&1; tail -4 $O/t8.log
(time KDPT_LIB=$PWD/gpurun_in/libkdpt_c8_8.so timeout 1800 python -m pytest tests/test_trace_gpu.py tests/test_large_gpu.py tests/test_render_gpu.py -x -q -k "not 10m") > $O/t8c.log 2>&1; tail -4 $O/t8c.log
run() {
  echo "== $1"
  python tools/microbench.py --modes fullstack 2>&1 | grep stack
  python tools/microbench.py --modes fullstack --soup 1000000 2>&1 | grep stack
  for w in dragon1m dragon5; do
    timeout 400 python bench.py --workload $w --no-cpu-baseline --no-extras --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w %.1f Mrays/s, %.3f ms/step, check %s'%(d['value'],d['ms_per_step'],d['check'].get('ok')))"
  done
}
run default
for n in c8_8 c4_4 c6_4 c12_8 c16_8 c8_16; do export KDPT_LIB=$PWD/gpurun_in/libkdpt_$n.so; run $n; done
export KDPT_LIB=$PWD/gpurun_in/libkdpt_c8_8.so
M="python tools/microbench.py --modes fullstack --reps 2"
$M > $O/mb_lean3_d5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:traceLean -s 1 -c 1 -o $O/prof_lean3_d5 -f $M > $O/ncu_lean3_d5.log 2>&1
