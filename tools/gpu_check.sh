#!/bin/bash
OUT=gpurun_out/${1:-v12}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $OUT/pytest.log; tail -1 $OUT/pytest.log
timeout 900 python tools/stress_parity.py 400 7 > $OUT/stress.log 2>&1; tail -3 $OUT/stress.log
for wl in c1b rub_c1a c3 c2; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_$wl.json
  python - <<PY
import json
d=json.load(open("$OUT/bench_$wl.json")); print("$wl %.3f ms/step  %.1f Mpix/s  e2e %.3f ms" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d["launches_by_kernel"])
PY
done
timeout 300 python bench.py --workload c4 --streams 1 --pairs 1 --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_c4_single.json
python -c "
import json; d=json.load(open('$OUT/bench_c4_single.json')); print('c4 single %.3f ms' % d['ms_per_step'])"
