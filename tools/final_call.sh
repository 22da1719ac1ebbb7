#!/bin/bash
# development helper, run ON the GPU box: full GPU test suite, every bench workload, smoke, refreshed ncu capture of the trace kernel
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/t_final.log 2>&1; tail -4 gpurun_out/t_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
for w in bunny dragon5 sphere dragon1m soup10m soup1m; do timeout 400 python bench.py --workload $w > gpurun_out/final_$w.log 2> gpurun_out/final_$w.err; tail -c 400 gpurun_out/final_$w.log; echo; done
timeout 300 python bench.py --impl reference > gpurun_out/final_ref_bunny.log 2>&1; tail -c 600 gpurun_out/final_ref_bunny.log




B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_bunny_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bunny_final.csv $B > gpurun_out/ncu_bunny_final_list.log 2>&1
$B > gpurun_out/plain_bunny_final2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bounceKernel -s 8 -c 4 -o gpurun_out/prof_bunny_final $B > gpurun_out/ncu_bunny_final_full.log 2>&1
ls gpurun_out | wc -l
