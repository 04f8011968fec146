#!/bin/bash
# GPU pass without ncu: parity tests, then a short bench line per BASELINE workload.   tools/quick_check.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
for WL in metric c2 c3 c4; do
  timeout 600 python bench.py --quick --steps 20 --workload $WL > gpurun_out/${TAG}_bench_$WL.log 2>&1
  python - "$WL" "gpurun_out/${TAG}_bench_$WL.log" <<'PY'
import json, sys
for l in open(sys.argv[2]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print(sys.argv[1], "%.4g evals/s" % d["value"], "%.4f ms/step" % d["ms_per_step"], "kernel %.4f" % r["kernel_ms_per_launch"],
              "prepare %.4f" % r["prepare_ms_per_launch"], "e2e %.4g" % d["e2e"]["value"], d["config"]["kernel"])
PY
done
