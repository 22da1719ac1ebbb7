#!/bin/bash
# round 2, call 14 (ON the GPU box): final state — suite, smoke, every bench line, ncu traffic of one 16-iteration batch (summaries only travel back)
O=gpurun_out/r2; mkdir -p $O
(time timeout 1800 python -m pytest tests -m gpu -x -q) > $O/t14.log 2>&1; grep -E "passed|failed" $O/t14.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/microbench.py 2>&1 | tail -4; python tools/microbench.py --cls outside 2>&1 | tail -4
for w in dragon1m bunny dragon5 sphere soup10m soup1m; do timeout 600 python bench.py --workload $w > $O/final_$w.log 2> $O/final_$w.err; python -c "
import json; d=json.loads(open('$O/final_$w.log').read().strip().splitlines()[-1]); print('$w %.1f Mrays/s e2e %.1f check %s'%(d['value'],d['e2e']['value'],d['check'].get('ok')))"; done
timeout 600 python bench.py --impl reference > $O/final_ref_dragon1m.log 2>&1; tail -c 300 $O/final_ref_dragon1m.log; echo
B="python bench.py --steps 2 --warmup 3 --spp-per-step 16 --no-cpu-baseline --no-extras"
$B > $O/plain_dragon1m4.log 2>&1 && ncu --set full --clock-control none -k regex:"frontKernel|traceQueueKernel|finishQueueKernel|queueBin" -s 135 -c 45 -o /tmp/prof_dragon1m_batch -f $B > $O/ncu_dragon1m_batch.log 2>&1
python tools/ncu_traffic.py /tmp/prof_dragon1m_batch.ncu-rep --group "front + queue sort + traceQueue + finishQueue (one bounce of a 16-iteration 1080p batch)" --per 8 > $O/traffic_dragon1m_batch16.txt 2>&1; tail -4 $O/traffic_dragon1m_batch16.txt | cut -c1-500
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_dragon1m_final.csv $B > $O/ncu_dragon1m_list2.log 2>&1
du -sh gpurun_out
