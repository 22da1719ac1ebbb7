#!/bin/bash
# round 2, call 24 (ON the GPU box): ncu --set full of the wavefront traversal kernel on the 1.6 M dragon with two-level records
# (bounces 0-3 of the second 16-iteration batch); the same bench command without ncu exited 0 in call 23
O=gpurun_out/r2; mkdir -p $O
D="python bench.py --workload dragon1m --steps 1 --warmup 1 --spp-per-step 16 --no-cpu-baseline --no-extras"
KDPT_WIDE=1 timeout 170 ncu --set full --clock-control none --import-source on -k regex:traceQueue -s 8 -c 4 -o $O/prof_dragon1m_wide -f $D > $O/ncu_dragon1m_wide.log 2>&1
ls -la $O/prof_dragon1m_wide.ncu-rep; tail -2 $O/ncu_dragon1m_wide.log | cut -c1-200
