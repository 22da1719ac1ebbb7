#!/bin/bash
# round 2, call 22 (ON the GPU box): the final state — whole GPU suite, smoke, the driver's two bench commands, the 10 M soup with its new
# default (two-level records) and a fresh ncu capture of that kernel
O=gpurun_out/r2; mkdir -p $O
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/t22.log 2>&1; grep -E "passed|failed|rror" $O/t22.log | tail -3; grep -E "^real" $O/t22.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 3) > $O/drv22_ref.log 2>&1; tail -4 $O/drv22_ref.log | cut -c1-300
(time python bench.py --gpus 1) > $O/drv22_kdpt.log 2> $O/drv22_kdpt.err; grep -E "^real" $O/drv22_kdpt.err; python -c "
import json; d=json.loads(open('$O/drv22_kdpt.log').read().strip().splitlines()[-1]); r=d['roofline']
print('dragon1m %.1f Mrays/s e2e %.1f frac %.3f traffic %.3e ach bytes %.3e check %s ref_t2 %.0f cpu %.3f launches %d kernel %s'%(d['value'],d['e2e']['value'],r['frac'],r['traffic'],r['algorithmic_bytes_per_launch'],d['check'].get('ok'),d['ref_t2']['value'],d['cpu_baseline']['value'],d['gpu_launches'],r['kernel']))"
T="python bench.py --workload soup10m --steps 5 --warmup 3 --no-cpu-baseline"
$T > $O/final22_soup10m.log 2> $O/final22_soup10m.err; python -c "
import json; d=json.loads(open('$O/final22_soup10m.log').read().strip().splitlines()[-1]); print('soup10m %.1f Mrays/s e2e %.1f frac %s check %s'%(d['value'],d['e2e']['value'],d['roofline'].get('frac'),d['check'].get('ok')))"
T="python bench.py --workload soup10m --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:traceLean -s 3 -c 1 -o $O/prof_soup10m_wide -f $T > $O/ncu_soup10m_wide.log 2>&1; ls -la $O/prof_soup10m_wide.ncu-rep
