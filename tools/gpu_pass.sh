#!/bin/bash
# GPU pass: smoke, parity tests, bench (both arms), ncu launch list + full captures of the evaluation kernel per workload.
#   tools/gpu_pass.sh <tag> [skip-ncu | "wl1 wl2": workloads to capture, default "metric c2"]
#   tools/gpu_pass.sh <tag> "c3 c4" ncu-only       (only the captures: gpurun merges at most 64 MiB back, ~23 MB a report)
TAG=${1:-r02}
mkdir -p gpurun_out
if [ "$3" != ncu-only ]; then
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/${TAG}_bench.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_reference.log 2>&1; echo "exit $?" >> gpurun_out/${TAG}_bench_reference.log
for f in smoke pytest_gpu bench bench_reference; do tail -n 3 gpurun_out/${TAG}_$f.log | cut -c1-600; done
[ "$2" = skip-ncu ] && exit 0
timeout 300 python bench.py --steps 20 --warmup 3 --quick > gpurun_out/${TAG}_bench_short.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --quick > gpurun_out/${TAG}_ncu_launches.log 2>&1
fi
for WL in ${2:-metric c2}; do
  SRC=""; [ "$WL" = metric ] && SRC="--import-source on"
  timeout 300 python bench.py --steps 8 --warmup 3 --quick --workload $WL > gpurun_out/${TAG}_bench_short_$WL.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none $SRC -k regex:objective_stream_kernel -s 5 -c 1 -o gpurun_out/${TAG}_prof_$WL python bench.py --steps 8 --warmup 3 --quick --workload $WL > gpurun_out/${TAG}_ncu_full_$WL.log 2>&1
done
du -sh gpurun_out
ls -la gpurun_out/${TAG}_*
