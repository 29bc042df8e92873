#!/bin/bash
# quick GPU check: parity tests + bench of the main workloads.  Usage: gpurun -- bash tools/gpu_quick.sh [tag]
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/pytest.log
for wl in c2 c1b c4 c3; do
  python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_$wl.json
  python - <<PY
import json
d=json.load(open("$OUT/bench_$wl.json"))
print("$wl", "ms/step %.3f"%d["ms_per_step"], "e2e ms %.3f"%d["e2e"]["ms_per_step"], "Mpix/s %.1f"%d["value"], "launch_us %.2f"%d["roofline"]["launch_us"], "frac %.3f"%d["roofline"]["frac"], "launches", d["gpu_launches"]//d["steps"])
PY
done
