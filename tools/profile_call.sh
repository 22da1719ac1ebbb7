#!/bin/bash
# development helper, run ON the GPU box (under gpurun): launch list + full ncu capture of the dominant kernels.
# every ncu run is preceded by the identical plain command, which must exit 0 (B200_PROFILING.md)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --spp-per-step 16 --no-cpu-baseline"
$B > gpurun_out/plain_bunny.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bunny.csv $B > gpurun_out/ncu_bunny_list.log 2>&1
$B > gpurun_out/plain_bunny2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bounceKernel -s 8 -c 4 -o gpurun_out/prof_bunny $B > gpurun_out/ncu_bunny_full.log 2>&1
T="python bench.py --workload soup10m --steps 1 --warmup 3 --no-cpu-baseline"
$T > gpurun_out/plain_soup.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 3 -c 1 -o gpurun_out/prof_soup $T > gpurun_out/ncu_soup_full.log 2>&1
D="python bench.py --workload dragon1m --steps 1 --warmup 3 --spp-per-step 16 --no-cpu-baseline"
$D > gpurun_out/plain_dragon1m.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bounceKernel -s 9 -c 2 -o gpurun_out/prof_dragon1m $D > gpurun_out/ncu_dragon1m_full.log 2>&1
for w in dragon5 dragon1m sphere; do python bench.py --workload $w --steps 10 > gpurun_out/b_$w.log 2> gpurun_out/b_$w.err; done
ls -la gpurun_out
