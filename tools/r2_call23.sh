#!/bin/bash
# round 2, call 23 (ON the GPU box): branch-free axis select in the descent (compSel, libkdpt.so) vs the state validated in call 22
# (libkdpt_prev.so), both node forms; then the whole GPU suite, smoke and the driver's bench command on the new library
O=gpurun_out/r2; mkdir -p $O; L=$O/call23.txt; : > $L
P=kdtreepathtraceroptimization_b200
mb() { python tools/microbench.py --modes fullstack,shortstack "$@" 2>&1 | grep stack | awk '{printf "%s %s ; ", $1, $4}'; }
bench() { timeout 400 python bench.py --workload $1 --no-cpu-baseline --no-extras --steps ${2:-6} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 %.1f Mrays/s %.3f ms/step check %s ;'%(d['value'],d['ms_per_step'],d['check'].get('ok')), end=' ')"; }
run() {
  echo -n "== $1: " | tee -a $L
  shift
  for w in "$@"; do
    case $w in
      d5) echo -n "d5 " | tee -a $L; mb | tee -a $L;;
      *) bench $w | tee -a $L;;
    esac
  done
  echo | tee -a $L
}
KDPT_LIB=$PWD/$P/libkdpt_prev.so KDPT_WIDE=0 run "previous, 16-byte nodes" d5 dragon1m
KDPT_WIDE=0 run "select, 16-byte nodes" d5 dragon1m dragon5
KDPT_WIDE=1 run "select, two-level records" d5 dragon1m
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $O/t23.log 2>&1; grep -E "passed|failed|rror" $O/t23.log | tail -3 | tee -a $L; grep -E "^real" $O/t23.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $L
(time python bench.py --gpus 1) > $O/drv23_kdpt.log 2> $O/drv23_kdpt.err; python -c "
import json; d=json.loads(open('$O/drv23_kdpt.log').read().strip().splitlines()[-1]); r=d['roofline']
print('dragon1m %.1f Mrays/s e2e %.1f frac %.3f check %s ref_t2 %.0f launches %d'%(d['value'],d['e2e']['value'],r['frac'],d['check'].get('ok'),d['ref_t2']['value'],d['gpu_launches']))" | tee -a $L
