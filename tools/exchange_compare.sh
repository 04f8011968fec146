#!/bin/bash
# bench.py at N GPUs with the best records exchanged over NCCL vs over peer memory (particle sharding, configs[1]).
#   tools/exchange_compare.sh <N>
N=${1:-2}
for ex in nccl p2p nccl p2p; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus $N --steps 200 --warmup 5 --quick --exchange $ex 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('$ex', d['n_gpus'], d['value'], d['ms_per_step'], d['gpu_launches'])"
done
