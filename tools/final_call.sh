#!/bin/bash
# development helper, run ON the GPU box: full GPU test suite, every bench workload, smoke, refreshed ncu capture of the trace kernel
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/t_final.log 2>&1; tail -4 gpurun_out/t_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
for w in bunny dragon5 sphere dragon1m soup10m soup1m; do timeout 400 python bench.py --workload $w > gpurun_out/final_$w.log 2> gpurun_out/final_$w.err; tail -c 400 gpurun_out/final_$w.log; echo; done
timeout 300 python bench.py --impl reference > gpurun_out/final_ref_bunny.log 2>&1; tail -c 600 gpurun_out/final_ref_bunny.log
T="python bench.py --workload soup10m --steps 1 --warmup 3 --no-cpu-baseline"
$T > gpurun_out/plain_soup.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 3 -c 1 -o gpurun_out/prof_soup_v2 $T > gpurun_out/ncu_soup_full2.log 2>&1
ls gpurun_out | head -50
