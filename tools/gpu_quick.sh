#!/bin/bash
# One short gpurun call: GPU tests + the default bench (both arms).   Usage: gpurun -- bash tools/gpu_quick.sh [tag] [pytest -k expr]
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/gpu.txt; nproc >> $OUT/gpu.txt
if [ -n "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x -k "$2" 2>&1 | tail -15 > $OUT/pytest.log
else
  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $OUT/pytest.log
fi
tail -8 $OUT/pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $OUT/bench_ref.json
timeout 900 python bench.py 2> $OUT/bench.err | tail -1 > $OUT/bench.json
python - <<PY
import json
try:
    r=json.load(open("$OUT/bench_ref.json")); print("ref", r.get("value"), r.get("ms_per_step"))
except Exception as e: print("ref parse", e)
try:
    d=json.load(open("$OUT/bench.json"))
    print("ours value %.1f e2e %.1f ms/step %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]))
    print("roofline", d["roofline"]["kernel"], d["roofline"]["launch_us"], d["roofline"]["frac"], "warp", d["roofline_warp"]["frac"])
    print("shares", d["kernel_time_shares"])
    for k,v in d["extra"].items(): print(k, {x:v.get(x) for x in ("value","ms_per_step","error")}, (v.get("e2e") or {}).get("value"))
    print("cpu", d["cpu_baseline"])
except Exception as e:
    print("ours parse", e); print(open("$OUT/bench.err").read()[-3000:])
PY
