#!/bin/bash
# round 2, call 11 (ON the GPU box): final-state suite, every bench workload, ncu traffic of one whole batch, 4-way table, start-up timings
O=gpurun_out/r2; mkdir -p $O
(time timeout 1800 python -m pytest tests -m gpu -x -q) > $O/t11.log 2>&1; tail -4 $O/t11.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke11.log 2>&1; tail -1 $O/smoke11.log
for b in 1 0; do echo "== KDPT_BIN_RAYS=$b"; KDPT_BIN_RAYS=$b python tools/microbench.py --modes fullstack 2>&1 | grep stack; KDPT_BIN_RAYS=$b python tools/microbench.py --modes fullstack --soup 1000000 2>&1 | grep stack; done
for w in dragon1m bunny dragon5 sphere soup10m soup1m; do timeout 600 python bench.py --workload $w > $O/final_$w.log 2> $O/final_$w.err; tail -c 300 $O/final_$w.log; echo; done
timeout 600 python bench.py --impl reference > $O/final_ref_dragon1m.log 2>&1; tail -c 400 $O/final_ref_dragon1m.log; echo
B="python bench.py --steps 2 --warmup 3 --spp-per-step 16 --no-cpu-baseline --no-extras"
$B > $O/plain_dragon1m3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"frontKernel|traceQueueKernel|finishQueueKernel" -s 24 -c 24 -o $O/prof_dragon1m_batch -f $B > $O/ncu_dragon1m_batch.log 2>&1
B="python bench.py --workload bunny --steps 2 --warmup 3 --spp-per-step 64 --no-cpu-baseline --no-extras"
$B > $O/plain_bunny3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"frontKernel|traceQueueKernel|finishQueueKernel" -s 12 -c 12 -o $O/prof_bunny_batch -f $B > $O/ncu_bunny_batch.log 2>&1
python tools/create_times.py > $O/create_times.txt 2>&1; cat $O/create_times.txt
# the reference's 4-way comparison (README: brute force / bounding box / kd-tree / short-stack) through the headless CLI, TESTINGMODE times
S=tests/golden/scenes; T=$(mktemp -d); cp $S/* $T/; gunzip -f $T/dragon_5.obj.gz
for sc in "cornell.txt sphere_low_1.obj" "cornell8.txt stanford_bunny.obj" "cornell.txt dragon_5.obj"; do set -- $sc
  for m in brute brute_bbox ref_t1 ref_t2 fullstack shortstack; do
    echo -n "$2 $m: "; ./kdtreepathtraceroptimization_b200/kdpt_render $T/$1 $T/$2 --res 800x800 --spp 16 --mode $m --testing --json 2>&1 | tail -1
  done
done > $O/four_way.txt 2>&1; cat $O/four_way.txt | cut -c1-330
