#!/bin/bash
# GPU pass: parity tests, bench, ncu launch list + one full capture of the dominant kernel.
#   tools/gpu_profile.sh <tag> [kernel-regex]
TAG=${1:-r01}
KREGEX=${2:-objective_uniform_kernel}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/${TAG}_bench.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_reference.log 2>&1; echo "exit $?" >> gpurun_out/${TAG}_bench_reference.log
timeout 300 python bench.py --steps 20 --warmup 3 --quick > gpurun_out/${TAG}_bench_short.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --quick > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --quick > gpurun_out/${TAG}_bench_short2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 5 -c 2 -o gpurun_out/${TAG}_prof python bench.py --steps 20 --warmup 3 --quick > gpurun_out/${TAG}_ncu_full.log 2>&1
for f in smoke pytest_gpu bench bench_reference; do tail -n 3 gpurun_out/${TAG}_$f.log; done
# extra captures: the fused swarm kernel (single-fit path) and the finish kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:swarm_fused_kernel -s 2 -c 1 -o gpurun_out/${TAG}_prof_fused python tools/fit_latency.py > gpurun_out/${TAG}_ncu_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:swarm_finish_kernel -s 5 -c 1 -o gpurun_out/${TAG}_prof_finish python bench.py --steps 20 --warmup 3 --quick > gpurun_out/${TAG}_ncu_finish.log 2>&1
