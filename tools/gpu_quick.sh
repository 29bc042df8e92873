#!/bin/bash
# quick A/B after a kernel change: GPU tests, the four main workloads, the in-kernel phase timeline
OUT=gpurun_out/${1:-quick}
mkdir -p $OUT
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $OUT/pytest.log; tail -1 $OUT/pytest.log
for wl in c2 c1b c3 c4; do
  python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_$wl.json
  python - <<PY
import json
d=json.load(open("$OUT/bench_$wl.json")); print("$wl %.3f ms/step  %.1f Mpix/s  e2e %.1f  launch_us %.1f frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["launch_us"], d["roofline"]["frac"]))
PY
done
python bench.py --workload c4 --streams 1 --pairs 1 --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_c4_single.json
python -c "
import json; d=json.load(open('$OUT/bench_c4_single.json')); print('c4 single %.3f ms' % d['ms_per_step'])"
(python tools/phase_timing.py 2048 2048 1 5; python tools/phase_timing.py 1024 1024 1 14 7) > $OUT/phase_timing.txt 2>&1; cat $OUT/phase_timing.txt
