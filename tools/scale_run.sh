#!/bin/bash
# Multi-GPU pass on ONE box with N GPUs: NCCL parity check of the particle-sharded swarm, then bench.py at every
# power of two up to N for BASELINE configs 2 (weak, particles), 3 (strong, spectra) and 4 (weak, 8,192 particles/GPU).
#   tools/scale_run.sh <tag> <N>
TAG=${1:-r01}
NMAX=${2:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=800 > gpurun_out/${TAG}_pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_multi.log
for W in c2 c3 c4; do
  : > gpurun_out/${TAG}_scale_${W}.jsonl
  N=1
  while [ $N -le $NMAX ]; do
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --workload $W --steps 50 --warmup 5 --quick >> gpurun_out/${TAG}_scale_${W}.jsonl 2>> gpurun_out/${TAG}_scale_${W}.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
        bench.py --gpus $N --workload $W --steps 50 --warmup 5 --quick >> gpurun_out/${TAG}_scale_${W}.jsonl 2>> gpurun_out/${TAG}_scale_${W}.err
    fi
    N=$((N * 2))
  done
done
tail -n 3 gpurun_out/${TAG}_pytest_multi.log
python - <<PY
import json
for w in ('c2', 'c3', 'c4'):
    base = None
    for line in open('gpurun_out/${TAG}_scale_%s.jsonl' % w):
        if not line.startswith('{'):
            continue
        d = json.loads(line)
        base = base or d['value']
        print(w, d['n_gpus'], '%.4g evals/s' % d['value'], 'x%.2f' % (d['value'] / base), '%.3f ms/step' % d['ms_per_step'], d['scaling'])
PY
