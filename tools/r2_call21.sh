#!/bin/bash
# round 2, call 21 (ON the GPU box): what changed since call 17 — two-level records chosen for trees beyond the L2 (both node forms tested
# explicitly, the full-size 10 M / 1.6 M cases walk them by default), KDPT_KDVIZ, the drop-in shim with VIZKD mapped
O=gpurun_out/r2; mkdir -p $O
(time timeout 1200 python -m pytest tests/test_large_gpu.py tests/test_trace_gpu.py tests/test_dropin_gpu.py -x -q) > $O/t21a.log 2>&1; grep -E "passed|failed|rror" $O/t21a.log | tail -3; grep -E "^real" $O/t21a.log
(time timeout 600 python -m pytest tests/test_render_gpu.py -x -q -k "kdviz or node_forms or ray_cache") > $O/t21b.log 2>&1; grep -E "passed|failed|rror" $O/t21b.log | tail -3; grep -E "^real" $O/t21b.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
