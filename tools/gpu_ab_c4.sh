#!/bin/bash
# A/B of scheduling knobs on the C4 batch (device + e2e Mpix/s).   Usage: gpurun -- bash tools/gpu_ab_c4.sh tag
OUT=gpurun_out/${1:-abc4}; mkdir -p $OUT
run() { # name, env..., args
  name=$1; shift
  env "$@" > /dev/null 2>&1 || true
}
for cfg in "base:0:0:0" "tm1:1:0:0" "s12:0:12:24" "s16:0:16:32" "tm1s16:1:16:32" "s4:0:4:16"; do
  IFS=: read name tm st pr <<< "$cfg"
  FLOW2D_BENCH_THROUGHPUT_MODE=$tm timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-extra $( [ $st -gt 0 ] && echo "--streams $st --pairs $pr" ) 2>$OUT/err_$name.txt | tail -1 > $OUT/bench_$name.json
  python -c "
import json; d=json.load(open('$OUT/bench_$name.json')); print('$name value %.1f e2e %.1f ms/step %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['schedule'])" || tail -3 $OUT/err_$name.txt
done
