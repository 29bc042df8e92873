#!/bin/bash
# fused mid-level kernel: correctness under a strict timeout first, then A/B
OUT=gpurun_out/r01i
mkdir -p $OUT
timeout 300 python -m pytest tests/test_flow_gpu.py tests/test_stages_gpu.py -m gpu -q -x 2>&1 | tail -5 > $OUT/pytest1.log; tail -2 $OUT/pytest1.log
if ! grep -q "passed" $OUT/pytest1.log || grep -q "failed\|error" $OUT/pytest1.log; then echo "STOP: tests failed or hung"; exit 0; fi
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $OUT/pytest.log; tail -2 $OUT/pytest.log
for wl in c1b rub_c1a c3 ; do
  for fused in 0 1; do
    if [ $fused = 0 ]; then export FLOW2D_NO_FUSED_LEVELS=1; else unset FLOW2D_NO_FUSED_LEVELS; fi
    timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_${wl}_fused$fused.json
    python - <<PY
import json
d=json.load(open("$OUT/bench_${wl}_fused$fused.json")); print("$wl fused=$fused %.3f ms/step  e2e %.3f ms" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), d["launches_by_kernel"])
PY
  done
done
unset FLOW2D_NO_FUSED_LEVELS
timeout 300 python bench.py --workload c4 --streams 1 --pairs 1 --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_c4_single.json
python -c "
import json; d=json.load(open('$OUT/bench_c4_single.json')); print('c4 single %.3f ms' % d['ms_per_step'])"
