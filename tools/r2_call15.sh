#!/bin/bash
# round 2, call 15 (ON the GPU box): postponed 32-wide triangle tests (KDPT_LEAN_QUEUED=1) vs the while-while loop; ncu traffic of one 16-iteration batch
O=gpurun_out/r2; mkdir -p $O
(KDPT_LIB=$PWD/gpurun_in/libkdpt_q.so timeout 1800 python -m pytest tests/test_trace_gpu.py tests/test_large_gpu.py tests/test_render_gpu.py -x -q) > $O/t15q.log 2>&1; grep -E "passed|failed|Error" $O/t15q.log | head -5
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
run default
KDPT_LIB=$PWD/gpurun_in/libkdpt_q.so run queued
KDPT_LIB=$PWD/gpurun_in/libkdpt_q6.so run queued6
B="python bench.py --steps 2 --warmup 3 --spp-per-step 16 --no-cpu-baseline --no-extras"
$B > $O/plain_dragon1m4.log 2>&1 && ncu --set full --clock-control none -k regex:"frontKernel|traceQueueKernel|finishQueueKernel|queueBin" -s 135 -c 45 -o /tmp/prof_dragon1m_batch -f $B > $O/ncu_dragon1m_batch.log 2>&1
python tools/ncu_traffic.py /tmp/prof_dragon1m_batch.ncu-rep --group "front + queue sort + traceQueue + finishQueue (one bounce of a 16-iteration 1080p batch)" --per 8 > $O/traffic_dragon1m_batch16.txt 2>&1; tail -4 $O/traffic_dragon1m_batch16.txt | cut -c1-500
KDPT_LIB=$PWD/gpurun_in/libkdpt_q.so python tools/microbench.py --modes fullstack --reps 2 > /dev/null 2>&1 && KDPT_LIB=$PWD/gpurun_in/libkdpt_q.so ncu --set full --clock-control none -k regex:traceLean -s 1 -c 1 -o /tmp/prof_leanq_d5 -f python tools/microbench.py --modes fullstack --reps 2 > $O/ncu_leanq_d5.log 2>&1
python tools/ncu_traffic.py /tmp/prof_leanq_d5.ncu-rep > $O/leanq_d5.txt 2>&1; cat $O/leanq_d5.txt | cut -c1-300
du -sh gpurun_out
