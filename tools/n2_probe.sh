#!/bin/bash
# e2e of the host batch path on 2 GPUs under a few settings (one line each): bash tools/n2_probe.sh
for cfg in "SA_NOP=1" "SA_HOST_PACK=0" "SA_HOST_PIPELINE=slots"; do
  env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$cfg', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'])"
done
nproc; nvidia-smi topo -m | head -6
