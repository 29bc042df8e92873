#!/bin/bash
# One gpurun call: bench (both arms, several workloads) + ncu launch lists + one full capture of the
# dominant kernel.  Everything lands in gpurun_out/.   Usage: gpurun -- bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt
nproc >> $OUT/gpu.txt; lscpu | grep "Model name" >> $OUT/gpu.txt
for wl in c2 c1b c4 c3; do
  python bench.py --impl reference --workload $wl --steps 3 --warmup 1 2>&1 | tail -1 > $OUT/bench_ref_$wl.json
  python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_$wl.json
  echo "== $wl"; cat $OUT/bench_ref_$wl.json; cat $OUT/bench_$wl.json
done
# launch lists (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 103 -c 103 --csv --log-file $OUT/launches_c2.csv python tools/profile_step.py c2 1 1 > $OUT/ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_c1b.csv python tools/profile_step.py c1b 0 1 > $OUT/ncu_c1b.log 2>&1
# full capture of the dominant kernel: c2 (L2-resident) and c3 finest level (DRAM-resident)
ncu --set full --clock-control none --import-source on -k regex:solve_pass -s 110 -c 2 -o $OUT/solve_c2 python tools/profile_step.py c2 1 1 > $OUT/ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_pass -s 1340 -c 2 -o $OUT/solve_c3 python tools/profile_step.py c3 0 1 > $OUT/ncu_full_c3.log 2>&1
ls -la $OUT
