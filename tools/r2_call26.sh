#!/bin/bash
# round 2, call 26 (ON the GPU box): the driver's bench command on the final default (1.6 M dragon on two-level records)
O=gpurun_out/r2; mkdir -p $O
timeout 80 python bench.py --gpus 1 > $O/drv26_kdpt.log 2> $O/drv26_kdpt.err; python -c "
import json; d=json.loads(open('$O/drv26_kdpt.log').read().strip().splitlines()[-1]); r=d['roofline']
print('dragon1m %.1f Mrays/s e2e %.1f frac %.3f check %s ref_t2 %.0f launches %d'%(d['value'],d['e2e']['value'],r['frac'],d['check'].get('ok'),d['ref_t2']['value'],d['gpu_launches']))"
