#!/bin/bash
# minimal A/B: solve tests, c2 + c3 bench, phase timeline
OUT=gpurun_out/${1:-quick}
mkdir -p $OUT
python -m pytest tests/test_stages_gpu.py tests/test_flow_gpu.py tests/test_slab_gpu.py -m gpu -q -x 2>&1 | tail -3 > $OUT/pytest.log; tail -1 $OUT/pytest.log
for wl in c2 c3 c1b; do
  python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_$wl.json
  python - <<PY
import json
d=json.load(open("$OUT/bench_$wl.json")); print("$wl %.3f ms/step  %.1f Mpix/s  e2e %.1f  launch_us %.1f frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["launch_us"], d["roofline"]["frac"]))
PY
done
(python tools/phase_timing.py 2048 2048 1 5; python tools/phase_timing.py 1024 1024 1 14 7) > $OUT/phase_timing.txt 2>&1; grep -E "outer|mean" $OUT/phase_timing.txt
