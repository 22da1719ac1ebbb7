#!/bin/bash
# round 2, call 20 (ON the GPU box): the empty-leaf policies of the lean loop (kdpt_lean.cuh KDPT_LEAN_SKIP_EMPTY: 0 = libkdpt.so, whose
# 16-byte-node kernel is instruction for instruction the previous one; 2 = libkdpt_s2.so, one loop of node steps and pops; s2c8 = the same
# with rounds capped at 8 trips) x node form (KDPT_WIDE: 0 = 16-byte nodes, 1 = two-level records)
O=gpurun_out/r2; mkdir -p $O; L=$O/call20.txt; : > $L
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
      s1) echo -n "soup1m " | tee -a $L; mb --soup 1000000 | tee -a $L;;
      *) bench $w | tee -a $L;;
    esac
  done
  echo | tee -a $L
}
KDPT_WIDE=1 run "policy 0, two-level records" d5 dragon1m s1 soup10m
KDPT_WIDE=0 run "policy 0, 16-byte nodes (= previous kernel)" d5 dragon1m
KDPT_LIB=$PWD/$P/libkdpt_s2.so KDPT_WIDE=1 run "policy 2, two-level records" d5 dragon1m s1 soup10m
KDPT_LIB=$PWD/$P/libkdpt_s2.so KDPT_WIDE=0 run "policy 2, 16-byte nodes" d5 dragon1m
KDPT_LIB=$PWD/$P/libkdpt_s2c8.so KDPT_WIDE=1 run "policy 2 capped at 8 trips, two-level records" d5 dragon1m
KDPT_LIB=$PWD/$P/libkdpt_s2c8.so KDPT_WIDE=0 run "policy 2 capped at 8 trips, 16-byte nodes" d5 dragon1m
KDPT_WIDE=1 run "policy 0, two-level records, other workloads" dragon5 bunny
