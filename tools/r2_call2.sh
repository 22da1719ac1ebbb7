#!/bin/bash
# round 2, call 2 (ON the GPU box): ncu of the traversal kernel, refined tree (per-lane leaves) vs reference tree (cooperative fat leaves)
O=gpurun_out/r2; mkdir -p $O
for r in 0 4; do
  M="python tools/microbench.py --modes fullstack --reps 2"
  KDPT_REFINE_LEAF=$r $M > $O/mb_d5_r$r.log 2>&1 && KDPT_REFINE_LEAF=$r ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 1 -c 1 -o $O/prof_d5_r$r -f $M > $O/ncu_d5_r$r.log 2>&1
  tail -1 $O/mb_d5_r$r.log
  M="python tools/microbench.py --modes fullstack --reps 2 --soup 1000000"
  KDPT_REFINE_LEAF=$r $M > $O/mb_s1_r$r.log 2>&1 && KDPT_REFINE_LEAF=$r ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 1 -c 1 -o $O/prof_s1_r$r -f $M > $O/ncu_s1_r$r.log 2>&1
  tail -1 $O/mb_s1_r$r.log
done
# reference-on-B200 row with the guarded copy
for a in "--scene cornell8.txt --obj stanford_bunny.obj" "--scene cornell.txt --obj dragon_5.obj" "--scene cornell.txt --obj dragon_5.obj --t1"; do
  timeout 600 python tools/ref_gpu_row.py --spp 8 $a > $O/refgpu_$(echo $a | md5sum | cut -c1-6).json 2> $O/refgpu_$(echo $a | md5sum | cut -c1-6).err; tail -c 1500 $O/refgpu_$(echo $a | md5sum | cut -c1-6).json; tail -3 $O/refgpu_$(echo $a | md5sum | cut -c1-6).err
done
ls -la $O | head -50
