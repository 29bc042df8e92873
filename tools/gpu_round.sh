#!/bin/bash
# One gpurun call that refreshes everything under profiles/: GPU tests, bench (both arms, all
# single-GPU workloads), the sequence driver, ncu launch lists, full captures of the solve and warp
# kernels, per-kernel DRAM throughput, the in-kernel phase timeline and compute-sanitizer.
#   Usage: gpurun -- bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT $OUT/sanitizer
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt
nproc >> $OUT/gpu.txt; lscpu | grep "Model name" >> $OUT/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $OUT/pytest.log; tail -2 $OUT/pytest.log
for wl in c2 rub_c1b rub_c1a c1b c4 c3 c3g; do
  timeout 600 python bench.py --impl reference --workload $wl --steps 3 --warmup 1 2>&1 | tail -1 > $OUT/bench_ref_$wl.json
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_$wl.json
  python - <<PY
import json
r=json.load(open("$OUT/bench_ref_$wl.json")); d=json.load(open("$OUT/bench_$wl.json"))
print("$wl", "ref %.1f Mpix/s (%.1f ms)"%(r.get("value",0), r.get("ms_per_step",0)), "| ours %.1f Mpix/s, e2e %.1f Mpix/s, ms/step %.2f, launch_us %.1f frac %.3f warp frac %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["launch_us"], d["roofline"]["frac"], d["roofline_warp"]["frac"]))
PY
done
timeout 300 python bench.py --workload c4 --streams 1 --pairs 1 --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_c4_single.json
timeout 300 python bench.py --workload c1b --streams 4 --pairs 8 --steps 5 --warmup 3 2>&1 | tail -1 > $OUT/bench_c1b_batch.json
timeout 300 python tools/bench_sequence.py 32 4 > $OUT/sequence_h4.json 2> $OUT/sequence.err
timeout 300 python tools/bench_sequence.py 32 8 > $OUT/sequence_h8.json 2>> $OUT/sequence.err
# launch lists (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 78 -c 78 --csv --log-file $OUT/launches_c2.csv python tools/profile_step.py c2 1 1 > $OUT/ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_c1b.csv python tools/profile_step.py c1b 0 1 > $OUT/ncu_c1b.log 2>&1
# full captures: C2 later pass (L2-resident), 4096^2 later pass (DRAM-resident), one tiny level, one small pass, warp
ncu --set full --clock-control none --import-source on -k regex:solve_pass -s 40 -c 1 -o $OUT/solve_c2 python tools/profile_step.py c2 0 1 > $OUT/ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_pass -s 2 -c 1 -o $OUT/solve_c3 python tools/profile_stages.py 4096 4096 > $OUT/ncu_full_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_tiny -s 12 -c 1 -o $OUT/solve_tiny python tools/profile_step.py c1b 0 1 > $OUT/ncu_full_tiny.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_small -s 700 -c 1 -o $OUT/solve_small python tools/profile_step.py c1b 0 1 > $OUT/ncu_full_small.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_kernel -c 1 -o $OUT/warp_4096 python tools/profile_stages.py 4096 4096 > $OUT/ncu_full_warp.log 2>&1
# per-kernel DRAM traffic and time at 4096x4096
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/stages_4096.csv python tools/profile_stages.py 4096 4096 > $OUT/stages.log 2>&1
# in-kernel phase timeline
(timeout 120 python tools/phase_timing.py 584 388 1 5; timeout 120 python tools/phase_timing.py 2048 2048 1 5; timeout 120 python tools/phase_timing.py 1024 1024 1 14 7) > $OUT/phase_timing.txt 2>&1
# compute-sanitizer over the four solve paths, both data terms
for t in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $t python tools/sanitize.py > $OUT/sanitizer/$t.log 2>&1
  grep -E "SUMMARY|^ok|done" $OUT/sanitizer/$t.log > $OUT/sanitizer/$t.txt
done
cat $OUT/sanitizer/*.txt
ls -la $OUT
