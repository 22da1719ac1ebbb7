#!/bin/bash
# round 2, call 10 (ON an 8-GPU box): scaling of the default bench at N=8 and N=4, config 5 at 8 GPUs (no collective)
O=gpurun_out/r2; mkdir -p $O
nvidia-smi -L | wc -l; free -g | head -2; nproc
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline --no-extras --steps 10 > $O/n8_dragon1m.log 2> $O/n8_dragon1m.err; python -c "
import json; d=json.loads(open('$O/n8_dragon1m.log').read().strip().splitlines()[-1]); print('N=8 dragon1m %.1f Mrays/s e2e %.1f spp/s %.1f clocks %s'%(d['value'],d['e2e']['value'],d['spp_per_s'],d['clocks']))"
tail -2 $O/n8_dragon1m.err
timeout 900 $T --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --no-cpu-baseline --no-extras --steps 10 > $O/n4_dragon1m.log 2> $O/n4_dragon1m.err; python -c "
import json; d=json.loads(open('$O/n4_dragon1m.log').read().strip().splitlines()[-1]); print('N=4 dragon1m %.1f Mrays/s e2e %.1f'%(d['value'],d['e2e']['value']))"
timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 10 > $O/n1b_dragon1m.log 2>/dev/null; python -c "
import json; d=json.loads(open('$O/n1b_dragon1m.log').read().strip().splitlines()[-1]); print('N=1 dragon1m %.1f Mrays/s e2e %.1f'%(d['value'],d['e2e']['value']))"
avail=$(free -g | awk '/Mem:/ {print $7}')
if [ "$avail" -gt 200 ]; then
timeout 900 $T --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workload soup10m --steps 5 --no-cpu-baseline > $O/n8_soup10m.log 2> $O/n8_soup10m.err; python -c "
import json; d=json.loads(open('$O/n8_soup10m.log').read().strip().splitlines()[-1]); print('N=8 soup10m %.1f Mrays/s e2e %.1f'%(d['value'],d['e2e']['value']))"
else echo "skipping soup10m at 8 ranks: only $avail GB of host memory available"; fi
