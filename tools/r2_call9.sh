#!/bin/bash
# round 2, call 9 (ON a 2-GPU box): peer all-reduce test, torchrun bench at N=2 (sample sharding + NCCL on the bound buffer), N=1 beside it
O=gpurun_out/r2; mkdir -p $O
nvidia-smi -L
(time timeout 900 python -m pytest tests/test_render_gpu.py -x -q -k "peer_allreduce or sharding" -rs) > $O/t9_peer.log 2>&1; tail -6 $O/t9_peer.log
(time timeout 1800 python -m pytest tests/test_trace_gpu.py tests/test_render_gpu.py -x -q) > $O/t9.log 2>&1; tail -3 $O/t9.log
timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 10 > $O/n1_dragon1m.log 2> $O/n1_dragon1m.err; python -c "
import json; d=json.loads(open('$O/n1_dragon1m.log').read().strip().splitlines()[-1]); print('N=1 dragon1m %.1f Mrays/s e2e %.1f'%(d['value'],d['e2e']['value']))"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-extras --steps 10 > $O/n2_dragon1m.log 2> $O/n2_dragon1m.err; python -c "
import json; d=json.loads(open('$O/n2_dragon1m.log').read().strip().splitlines()[-1]); print('N=2 dragon1m %.1f Mrays/s e2e %.1f launches %d'%(d['value'],d['e2e']['value'],d['gpu_launches']))"
tail -3 $O/n2_dragon1m.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload soup10m --steps 5 --no-cpu-baseline > $O/n2_soup10m.log 2> $O/n2_soup10m.err; python -c "
import json; d=json.loads(open('$O/n2_soup10m.log').read().strip().splitlines()[-1]); print('N=2 soup10m %.1f Mrays/s e2e %.1f'%(d['value'],d['e2e']['value']))"
