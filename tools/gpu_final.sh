#!/bin/bash
# Last call of round 2: everything with the final defaults on one box.  gpurun --timeout 480 -- bash tools/gpu_final.sh
OUT=gpurun_out/final
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
timeout 300 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -8 > $OUT/pytest.log; tail -2 $OUT/pytest.log
timeout 200 python bench.py 2>$OUT/bench.err | tail -1 > $OUT/bench.json
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench: %.1f Mpix/s, e2e %.1f, ms/step %.2f, frac %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]), d["launches_by_kernel"])
for k,v in (d.get("extra") or {}).items():
    print(k, {kk: v[kk] for kk in ("value","ms_per_step") if kk in v})
PY
timeout 150 python tools/ab_cluster.py $OUT/ab_quick.json quick 2> $OUT/ab_quick.log > /dev/null; grep -v "^[0-9]*x" $OUT/ab_quick.log | tail -20
timeout 90 ncu --set full --clock-control none --import-source on -k regex:solve_cluster -s 1 -c 1 -o $OUT/solve_cluster python tools/profile_cluster.py 64 64 > $OUT/ncu_cluster.log 2>&1
tail -2 $OUT/ncu_cluster.log
ls -la $OUT
