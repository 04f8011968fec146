#!/bin/bash
# Multi-GPU pass on one box: parity tests (sharded swarm and spectra bit-identical to one GPU, both exchanges), then the
# bench at the GPU counts given.    tools/scale_pass.sh <tag> "8 4"
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pso.py -m gpu -q --timeout=800 -k "multi or communicator or nccl or sharded" > gpurun_out/${TAG}_pytest_multi.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_multi.log
PORT=29700
for N in ${2:-8}; do
  for EX in p2p nccl; do
    PORT=$((PORT+1))
    Q="--quick"; [ "$EX" = p2p ] && Q=""
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 20 --warmup 5 --exchange $EX $Q > gpurun_out/${TAG}_bench_${N}gpu_$EX.log 2>&1
    grep "^{" gpurun_out/${TAG}_bench_${N}gpu_$EX.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('N=$N $EX', '%.4g evals/s' % d['value'], '%.4f ms/step' % d['ms_per_step'], 'launches', d['gpu_launches'],
          'identical', d.get('sharded_identical_on_all_ranks'), 'bit-identical to 1 GPU', d.get('bit_identical_to_one_gpu'))
    for k, v in d.get('secondary', {}).items():
        print('   ', k, '%.4g evals/s' % v['value'], '%.4f ms/step' % v['ms_per_step'], v.get('sharded_identical_on_all_ranks'))
"
  done
done
