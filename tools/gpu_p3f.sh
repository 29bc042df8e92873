#!/bin/bash
OUT=gpurun_out/${1:-p3f}; mkdir -p $OUT
for rep in 1 2; do
for v in model padall halo4; do
  unset FLOW2D_P3_PAD2ALL FLOW2D_P3_HALO4
  [ $v = padall ] && export FLOW2D_P3_PAD2ALL=1
  [ $v = halo4 ] && export FLOW2D_P3_HALO4=1
  for wl in c4 c4s; do
    args="--workload c4"; [ $wl = c4s ] && args="--workload c4 --streams 1 --pairs 1"
    timeout 300 python bench.py $args --steps 4 --warmup 2 --no-extra 2>$OUT/err_${wl}_$v.txt | tail -1 > $OUT/bench_${wl}_${v}_$rep.json
    python -c "
import json; d=json.load(open('$OUT/bench_${wl}_${v}_$rep.json')); print('$rep $wl $v value %.1f e2e %.1f ms/step %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step']))" || tail -3 $OUT/err_${wl}_$v.txt
  done
done
done
