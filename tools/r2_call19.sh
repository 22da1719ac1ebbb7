#!/bin/bash
# round 2, call 19 (ON the GPU box): empty leaves popped inside the descent (new libkdpt.so) vs the previous kernel (libkdpt_base.so,
# built from the parent commit), 16-byte nodes vs two-level records (KDPT_WIDE), refill thresholds, SAH parameters under the new loop
O=gpurun_out/r2; mkdir -p $O; L=$O/call19.txt; : > $L
P=kdtreepathtraceroptimization_b200
mb() { python tools/microbench.py --modes fullstack,shortstack "$@" 2>&1 | grep stack | awk '{printf "%s %s ; ", $1, $4}'; }
bench() { timeout 400 python bench.py --workload $1 --no-cpu-baseline --no-extras --steps ${2:-6} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 %.1f Mrays/s %.3f ms/step check %s ;'%(d['value'],d['ms_per_step'],d['check'].get('ok')), end=' ')"; }
run() {   # name, what to run
  echo -n "== $1: " | tee -a $L
  shift
  for w in "$@"; do
    case $w in
      d5) echo -n "d5 " | tee -a $L; mb | tee -a $L;;
      s1) echo -n "soup1m " | tee -a $L; mb --soup 1000000 | tee -a $L;;
      *) bench $w | tee -a $L;;
    esac
  done
  echo | tee -a $L
}
KDPT_WIDE=0 run "new, 16-byte nodes" d5 s1 dragon1m dragon5 bunny
KDPT_WIDE=1 run "new, two-level records" d5 s1 dragon1m dragon5 bunny
KDPT_LIB=$PWD/$P/libkdpt_base.so run "base (parent commit)" d5 s1 dragon1m
KDPT_WIDE=1 run "new, two-level, soup10m" soup10m
KDPT_WIDE=0 run "new, 16-byte, soup10m" soup10m
# correctness of the two-level form on the GPU (the default suite covers it through nodeForm = 2 as well)
(KDPT_WIDE=1 timeout 600 python -m pytest tests/test_trace_gpu.py tests/test_large_gpu.py tests/test_render_gpu.py -x -q -k "not 10m and not headless and not peer") > $O/t19.log 2>&1; grep -E "passed|failed|error" $O/t19.log | tee -a $L
for r in 4 16; do
  KDPT_LIB=$PWD/$P/libkdpt_r$r.so KDPT_WIDE=1 run "refill at $r idle lanes, two-level" d5 dragon1m
done
KDPT_WIDE=1 KDPT_REFINE_LEAF=4 run "two-level, leaf 4" d5 dragon1m
KDPT_WIDE=1 KDPT_REFINE_LEAF=16 run "two-level, leaf 16" d5 dragon1m
KDPT_WIDE=1 KDPT_REFINE_CT=1 run "two-level, SAH cost 1" d5 dragon1m
KDPT_WIDE=1 KDPT_REFINE_CT=4 run "two-level, SAH cost 4" d5 dragon1m
KDPT_WIDE=0 KDPT_REFINE_LEAF=4 run "16-byte, leaf 4" d5 dragon1m
