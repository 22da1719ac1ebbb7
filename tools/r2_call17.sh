#!/bin/bash
# round 2, call 17 (ON the GPU box): the committed final state once more — whole GPU suite, smoke, the driver's two bench commands
O=gpurun_out/r2; mkdir -p $O
(time timeout 1800 python -m pytest tests -m gpu -x -q) > $O/t17.log 2>&1; grep -E "passed|failed" $O/t17.log; grep -E "^real" $O/t17.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 3) > $O/drv_ref.log 2>&1; tail -4 $O/drv_ref.log | cut -c1-300
(time python bench.py --gpus 1) > $O/drv_kdpt.log 2> $O/drv_kdpt.err; grep -E "^real" $O/drv_kdpt.err; python -c "
import json; d=json.loads(open('$O/drv_kdpt.log').read().strip().splitlines()[-1]); r=d['roofline']
print('dragon1m %.1f Mrays/s e2e %.1f frac %.3f traffic %.3e ach bytes %.3e check %s ref_t2 %.0f cpu %.3f'%(d['value'],d['e2e']['value'],r['frac'],r['traffic'],r['algorithmic_bytes_per_launch'],d['check'].get('ok'),d['ref_t2']['value'],d['cpu_baseline']['value']))"
