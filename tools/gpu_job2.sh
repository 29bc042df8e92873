#!/bin/bash
# second GPU job of the round: new tests, C3 capture, gradient C3, warp roofline, sequence driver
OUT=gpurun_out/r01d
mkdir -p $OUT
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $OUT/pytest.log; tail -2 $OUT/pytest.log
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_c2.json
python - <<PY
import json
d=json.load(open("$OUT/bench_c2.json")); print("c2", d["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline_warp"], d["launches_by_kernel"])
PY
for wl in c3g c1b; do
  python bench.py --impl reference --workload $wl --steps 2 --warmup 1 2>&1 | tail -1 > $OUT/bench_ref_$wl.json
  python bench.py --workload $wl --steps 5 --warmup 3 2>&1 | tail -1 > $OUT/bench_$wl.json
  python - <<PY
import json
r=json.load(open("$OUT/bench_ref_$wl.json")); d=json.load(open("$OUT/bench_$wl.json"))
print("$wl", "ref %.1f ms"%r.get("ms_per_step",0), "| ours %.2f ms e2e %.2f ms"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), d["roofline"]["kernel"], "warp frac %.3f"%d["roofline_warp"]["frac"], d["launches_by_kernel"])
PY
done
python tools/bench_sequence.py 32 8 > $OUT/sequence.json 2> $OUT/sequence.err; cat $OUT/sequence.json
python tools/bench_sequence.py 32 4 > $OUT/sequence_h4.json 2>> $OUT/sequence.err; cat $OUT/sequence_h4.json
ncu --set full --clock-control none --import-source on -k regex:solve_pass -s 2 -c 1 -o $OUT/solve_c3 python tools/profile_stages.py 4096 4096 > $OUT/ncu_full_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_small -s 100 -c 1 -o $OUT/solve_small python tools/profile_step.py c1b 0 1 > $OUT/ncu_full_small.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_kernel -c 1 -o $OUT/warp_4096 python tools/profile_stages.py 4096 4096 > $OUT/ncu_full_warp.log 2>&1
ls -la $OUT
