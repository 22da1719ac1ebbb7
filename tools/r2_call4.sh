#!/bin/bash
# round 2, call 4 (ON the GPU box): wavefront render form — parity tests (default policy and forced on), A/B, ncu of the lean kernel
O=gpurun_out/r2; mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/t4.log 2>&1; tail -4 $O/t4.log
(time KDPT_WAVEFRONT=1 timeout 900 python -m pytest tests/test_render_gpu.py tests/test_dropin_gpu.py -x -q) > $O/t4w.log 2>&1; tail -4 $O/t4w.log
for w in dragon1m dragon5 bunny sphere; do
  for wf in 0 1; do
    KDPT_WAVEFRONT=$wf timeout 400 python bench.py --workload $w --no-cpu-baseline --no-extras --steps 8 > $O/wf_${w}_$wf.log 2> $O/wf_${w}_$wf.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/wf_${w}_$wf.log").read().strip().splitlines()[-1])
    print("$w wavefront $wf: %.1f Mrays/s, %.3f ms/step, e2e %.1f, check %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['check']))
except Exception as e:
    print("$w wavefront $wf: FAILED", e)
PY
  done
done
M="python tools/microbench.py --modes fullstack --reps 2"
$M > $O/mb_lean_d5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:traceLean -s 1 -c 1 -o $O/prof_lean_d5 -f $M > $O/ncu_lean_d5.log 2>&1
M="python tools/microbench.py --modes fullstack --reps 2 --soup 1000000"
$M > $O/mb_lean_s1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:traceLean -s 1 -c 1 -o $O/prof_lean_s1 -f $M > $O/ncu_lean_s1.log 2>&1
KDPT_DEBUG_STATS=1 python bench.py --workload dragon1m --no-cpu-baseline --no-extras --steps 1 2>&1 | grep "kdpt stats" | tail -8
