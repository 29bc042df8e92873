#!/bin/bash
OUT=gpurun_out/${1:-v13}
mkdir -p $OUT
timeout 300 python tools/level_timing.py 0 > $OUT/level_timing.txt 2>&1; cat $OUT/level_timing.txt
timeout 300 python tools/level_timing.py 1 > $OUT/level_timing_throughput.txt 2>&1; cat $OUT/level_timing_throughput.txt
timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_flow_gpu.py -m gpu -q -x 2>&1 | tail -2
for wl in c1b c2; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_$wl.json
  python - <<PY
import json
d=json.load(open("$OUT/bench_$wl.json")); print("$wl %.3f ms/step  %.1f Mpix/s  e2e %.3f ms" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
PY
done
