#!/bin/bash
# round 2, call 12 (ON the GPU box): final-state suite, queue / batch binning A/B, every bench workload, ncu traffic of one whole batch
# (ncu reports stay in /tmp on the box: only their summaries are small enough to travel back)
O=gpurun_out/r2; mkdir -p $O
(time timeout 1800 python -m pytest tests -m gpu -x -q) > $O/t12.log 2>&1; grep -E "passed|failed" $O/t12.log
(KDPT_QUEUE_BIN=1 timeout 900 python -m pytest tests/test_render_gpu.py tests/test_converged_gpu.py -x -q) > $O/t12q.log 2>&1; grep -E "passed|failed" $O/t12q.log
for q in 0 1; do for w in dragon1m dragon5 bunny; do
  KDPT_QUEUE_BIN=$q timeout 400 python bench.py --workload $w --no-cpu-baseline --no-extras --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('queue_bin=$q $w %.1f Mrays/s, %.3f ms/step, check %s'%(d['value'],d['ms_per_step'],d['check'].get('ok')))"
done; done
for b in 1 0; do KDPT_BIN_RAYS=$b python bench.py --workload soup10m --steps 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bin_rays=$b soup10m %.1f Mrays/s e2e %.1f check %s'%(d['value'],d['e2e']['value'],d['check'].get('ok')))"
KDPT_BIN_RAYS=$b python tools/microbench.py --modes fullstack,shortstack --cls outside 2>&1 | grep stack; done
for w in dragon1m bunny dragon5 sphere soup10m soup1m; do timeout 600 python bench.py --workload $w > $O/final_$w.log 2> $O/final_$w.err; tail -c 200 $O/final_$w.log; echo; done
timeout 600 python bench.py --impl reference > $O/final_ref_dragon1m.log 2>&1
B="python bench.py --steps 2 --warmup 3 --spp-per-step 16 --no-cpu-baseline --no-extras"
$B > $O/plain_dragon1m3.log 2>&1 && ncu --set full --clock-control none -k regex:"frontKernel|traceQueueKernel|finishQueueKernel" -s 24 -c 24 -o /tmp/prof_dragon1m_batch -f $B > $O/ncu_dragon1m_batch.log 2>&1
python tools/ncu_traffic.py /tmp/prof_dragon1m_batch.ncu-rep --group "front + traceQueue + finishQueue (one bounce of a 16-iteration 1080p batch)" --per 8 > $O/traffic_dragon1m_batch.txt 2>&1; tail -3 $O/traffic_dragon1m_batch.txt | cut -c1-600
B="python bench.py --workload bunny --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$B > $O/plain_bunny3.log 2>&1 && ncu --set full --clock-control none -k regex:"frontKernel|traceQueueKernel|finishQueueKernel" -s 12 -c 12 -o /tmp/prof_bunny_batch -f $B > $O/ncu_bunny_batch.log 2>&1
python tools/ncu_traffic.py /tmp/prof_bunny_batch.ncu-rep --group "front + traceQueue + finishQueue (one bounce of a 64-iteration 800x800 batch)" --per 4 > $O/traffic_bunny_batch.txt 2>&1; tail -1 $O/traffic_bunny_batch.txt | cut -c1-600
du -sh gpurun_out
