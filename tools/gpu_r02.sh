#!/bin/bash
# Round-2 evidence in ONE gpurun call: smoke, GPU tests, the default bench (both arms), ncu launch lists and
# full captures of the dominant kernels, per-kernel DRAM traffic.   Usage: gpurun -- bash tools/gpu_r02.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt
nproc >> $OUT/gpu.txt; lscpu | grep "Model name" >> $OUT/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
if [ -z "$SKIP_TESTS" ]; then
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 > $OUT/pytest.log; tail -2 $OUT/pytest.log
fi
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$OUT/bench_ref.err | tail -1 > $OUT/bench_ref.json
timeout 900 python bench.py 2>$OUT/bench.err | tail -1 > $OUT/bench.json
python - <<PY
import json
r=json.load(open("$OUT/bench_ref.json")); d=json.load(open("$OUT/bench.json"))
print("ref %.1f Mpix/s (%.1f ms)"%(r.get("value",0), r.get("ms_per_step",0)), "| ours %.1f Mpix/s, e2e %.1f, ms/step %.2f, launch_us %.1f frac %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["launch_us"], d["roofline"]["frac"]))
for k,v in (d.get("extra") or {}).items():
    print(k, {kk: v[kk] for kk in ("value","ms_per_step") if kk in v}, (v.get("roofline") or {}).get("frac"))
PY
for wl in c2 c1b c3g; do
  timeout 300 python bench.py --workload $wl --no-extra 2>/dev/null | tail -1 > $OUT/bench_$wl.json
done
timeout 300 python bench.py --workload c4 --streams 1 --pairs 1 --no-extra 2>/dev/null | tail -1 > $OUT/bench_c4_single.json
python - <<PY
import json
for wl in ("c2","c1b","c3g","c4_single"):
    try:
        d=json.load(open("$OUT/bench_%s.json"%wl)); print(wl, "%.1f Mpix/s e2e %.1f ms/step %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("launches_by_kernel"))
    except Exception as e: print(wl, "failed", e)
PY
# launch lists (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_c4.csv python tools/profile_step.py c4 0 1 > $OUT/ncu_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 78 -c 78 --csv --log-file $OUT/launches_c2.csv python tools/profile_step.py c2 1 1 > $OUT/ncu_c2.log 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:solve_pass3 -s 40 -c 1 -o $OUT/solve_c2 python tools/profile_step.py c2 0 1 > $OUT/ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_pass3 -s 540 -c 1 -o $OUT/solve_c4 python tools/profile_step.py c4 0 1 > $OUT/ncu_full_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_pass3 -s 2 -c 1 -o $OUT/solve_2048 python tools/profile_stages.py 2048 2048 > $OUT/ncu_full_2048.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_small -s 700 -c 1 -o $OUT/solve_small python tools/profile_step.py c4 0 1 > $OUT/ncu_full_small.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_kernel -c 1 -o $OUT/warp_2048 python tools/profile_stages.py 2048 2048 > $OUT/ncu_full_warp.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/stages_2048.csv python tools/profile_stages.py 2048 2048 > $OUT/stages.log 2>&1
mkdir -p $OUT/sanitizer
for t in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $t python tools/sanitize.py > $OUT/sanitizer/$t.log 2>&1
  grep -E "SUMMARY|^ok|done" $OUT/sanitizer/$t.log > $OUT/sanitizer/$t.txt
done
cat $OUT/sanitizer/*.txt | grep -E "SUMMARY|done"
timeout 300 python tools/bench_sequence.py 64 8 2>/dev/null | tail -1 > $OUT/sequence_h8.json
ls -la $OUT
