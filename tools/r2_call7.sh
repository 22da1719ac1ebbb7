#!/bin/bash
# round 2, call 7 (ON the GPU box): full GPU suite, cooperative ref modes, new defaults, ncu captures of the new pipeline
O=gpurun_out/r2; mkdir -p $O
(time timeout 1800 python -m pytest tests -m gpu -x -q) > $O/t7.log 2>&1; tail -4 $O/t7.log
python tools/microbench.py 2>&1 | tail -5
python tools/microbench.py --cls outside 2>&1 | tail -4
for n in r12 r16 b10; do export KDPT_LIB=$PWD/gpurun_in/libkdpt_$n.so; echo "== $n"; python tools/microbench.py --modes fullstack 2>&1 | grep stack; python tools/microbench.py --modes fullstack --soup 1000000 2>&1 | grep stack
  timeout 400 python bench.py --no-cpu-baseline --no-extras --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dragon1m %.1f Mrays/s, %.3f ms/step, check %s'%(d['value'],d['ms_per_step'],d['check'].get('ok')))"
done
unset KDPT_LIB
timeout 900 python bench.py > $O/bench_dragon1m.log 2> $O/bench_dragon1m.err; tail -c 3000 $O/bench_dragon1m.log
B="python bench.py --steps 2 --warmup 3 --spp-per-step 16 --no-cpu-baseline --no-extras"
$B > $O/plain_dragon1m.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_dragon1m.csv $B > $O/ncu_dragon1m_list.log 2>&1
$B > $O/plain_dragon1m2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:traceQueueKernel -s 12 -c 3 -o $O/prof_dragon1m_tq -f $B > $O/ncu_dragon1m_tq.log 2>&1
T="python bench.py --workload soup10m --steps 1 --warmup 3 --no-cpu-baseline"
$T > $O/plain_soup10m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:traceLean -s 3 -c 1 -o $O/prof_soup10m_lean -f $T > $O/ncu_soup10m_lean.log 2>&1
tail -c 600 $O/plain_soup10m.log
