#!/bin/bash
# round 2, call 25 (ON the GPU box): two-level records from 256 MB of tree on — the 1.6 M dragon now walks them by default: its
# full-size parity test (4 096 sampled rays vs the oracle, bit for bit) on the new default
O=gpurun_out/r2; mkdir -p $O
(time timeout 100 python -m pytest tests/test_large_gpu.py -x -q -k "dense_dragon_1p6m and not reference_tree") > $O/t25.log 2>&1; grep -E "passed|failed|rror" $O/t25.log | tail -3; grep -E "^real" $O/t25.log
