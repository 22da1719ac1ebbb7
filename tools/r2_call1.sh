#!/bin/bash
# round 2, call 1 (run ON the GPU box): GPU tests with the refined tree, then A/B of leaf refinement on every workload
mkdir -p gpurun_out/r2
O=gpurun_out/r2
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/t1.log 2>&1; tail -4 $O/t1.log
for w in dragon1m soup10m soup1m dragon5 bunny sphere; do
  for r in 0 4 8; do
    KDPT_REFINE_LEAF=$r timeout 400 python bench.py --workload $w --no-cpu-baseline --steps 8 > $O/ab_${w}_r$r.log 2> $O/ab_${w}_r$r.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/ab_${w}_r$r.log").read().strip().splitlines()[-1])
    print("$w refine $r: %.1f Mrays/s, %.3f ms/step, e2e %.1f, refined %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('tree_refined')))
except Exception as e:
    print("$w refine $r: FAILED", e)
PY
  done
done
for r in 0 4; do echo "microbench dragon_5 refine $r"; KDPT_REFINE_LEAF=$r python tools/microbench.py 2>&1 | tail -5; done
