#!/bin/bash
OUT=gpurun_out/r01e
mkdir -p $OUT
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $OUT/pytest.log; tail -2 $OUT/pytest.log
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_c2.json
python bench.py --workload c3 --steps 5 --warmup 3 2>&1 | tail -1 > $OUT/bench_c3.json
python - <<PY
import json
for w in ("c2","c3"):
    d=json.load(open("$OUT/bench_%s.json"%w)); print(w, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline_warp"])
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/stages_4096.csv python tools/profile_stages.py 4096 4096 > $OUT/stages.log 2>&1
grep -E "warp_kernel|derivatives" $OUT/stages_4096.csv | grep duration
