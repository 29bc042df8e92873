#!/bin/bash
OUT=gpurun_out/r01g
mkdir -p $OUT
timeout 300 python -m pytest tests/test_residual_gpu.py -m gpu -q 2>&1 | tail -30 > $OUT/pytest_res.log; tail -30 $OUT/pytest_res.log
timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_flow_gpu.py tests/test_slab_gpu.py tests/test_reference_gpu.py -m gpu -q -x 2>&1 | tail -3 > $OUT/pytest.log; tail -1 $OUT/pytest.log
for wl in c2 c3 c4; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_$wl.json
  python - <<PY
import json
d=json.load(open("$OUT/bench_$wl.json")); print("$wl %.3f ms/step  %.1f Mpix/s  e2e %.1f  launch_us %.1f frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["launch_us"], d["roofline"]["frac"]))
PY
done
(timeout 120 python tools/phase_timing.py 2048 2048 1 5; timeout 120 python tools/phase_timing.py 1024 1024 1 14 7) > $OUT/phase_timing.txt 2>&1; grep -E "outer|mean" $OUT/phase_timing.txt
