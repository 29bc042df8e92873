#!/bin/bash
# A/B of the bulk-copy prefetch: solve tests, then c2/c3/c4 with and without FLOW2D_NO_PREFETCH
OUT=gpurun_out/${1:-quick}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_stages_gpu.py tests/test_flow_gpu.py tests/test_slab_gpu.py tests/test_reference_gpu.py -m gpu -q -x 2>&1 | tail -3 > $OUT/pytest.log; tail -1 $OUT/pytest.log
for wl in c2 c3 c4; do
  for pf in 0 1; do
    if [ $pf = 0 ]; then export FLOW2D_NO_PREFETCH=1; else unset FLOW2D_NO_PREFETCH; fi
    timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_${wl}_pf$pf.json
    python - <<PY
import json
d=json.load(open("$OUT/bench_${wl}_pf$pf.json")); print("$wl prefetch=$pf %.3f ms/step  %.1f Mpix/s  e2e %.1f  launch_us %.1f frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["launch_us"], d["roofline"]["frac"]))
PY
  done
done
unset FLOW2D_NO_PREFETCH
(timeout 120 python tools/phase_timing.py 2048 2048 1 5; timeout 120 python tools/phase_timing.py 1024 1024 1 14 7) > $OUT/phase_timing.txt 2>&1; grep -E "outer|mean" $OUT/phase_timing.txt
