#!/bin/bash
# round 2, call 16 (ON the GPU box): leaf tests of a round dealt evenly over the lanes (KDPT_LEAN_DEALT=1) vs the while-while loop
O=gpurun_out/r2; mkdir -p $O
(KDPT_LIB=$PWD/gpurun_in/libkdpt_d.so timeout 1800 python -m pytest tests/test_trace_gpu.py tests/test_large_gpu.py tests/test_render_gpu.py -x -q) > $O/t16d.log 2>&1; grep -E "passed|failed|Error" $O/t16d.log | head -5
run() {
  echo "== $1"
  python tools/microbench.py --modes fullstack,shortstack 2>&1 | grep stack
  python tools/microbench.py --modes fullstack --soup 1000000 2>&1 | grep stack
  for w in dragon1m dragon5 bunny; do
    timeout 400 python bench.py --workload $w --no-cpu-baseline --no-extras --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w %.1f Mrays/s, %.3f ms/step, check %s'%(d['value'],d['ms_per_step'],d['check'].get('ok')))"
  done
  python bench.py --workload soup10m --steps 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('soup10m %.1f Mrays/s check %s'%(d['value'],d['check'].get('ok')))"
}
KDPT_LIB=$PWD/gpurun_in/libkdpt_d.so run dealt
KDPT_LIB=$PWD/gpurun_in/libkdpt_d6.so run dealt6
KDPT_LIB=$PWD/gpurun_in/libkdpt_d.so python tools/microbench.py --modes fullstack --reps 2 > /dev/null 2>&1 && KDPT_LIB=$PWD/gpurun_in/libkdpt_d.so ncu --set full --clock-control none -k regex:traceLean -s 1 -c 1 -o /tmp/prof_leand_d5 -f python tools/microbench.py --modes fullstack --reps 2 > $O/ncu_leand_d5.log 2>&1
python tools/ncu_traffic.py /tmp/prof_leand_d5.ncu-rep > $O/leand_d5.txt 2>&1; cat $O/leand_d5.txt | cut -c1-300
