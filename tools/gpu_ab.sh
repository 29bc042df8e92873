#!/bin/bash
# A/B of the tiled solve kernels (FLOW2D_SOLVE_V1=1 = first generation) + correctness of the new one.
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
true
timeout 1200 python -m pytest tests -m gpu -q -x -k "solve or compute_vs_oracle or rub_pair or native or slab or small_pass or full_size" 2>&1 | tail -15 > $OUT/pytest.log; tail -6 $OUT/pytest.log
for v in ${VARIANTS:-v2 v1}; do
  if [ $v = v1 ]; then export FLOW2D_SOLVE_V1=1; else unset FLOW2D_SOLVE_V1; fi
  (timeout 120 python tools/phase_timing.py 2048 2048 1 5; timeout 120 python tools/phase_timing.py 1024 1024 1 14 7) > $OUT/phase_$v.txt 2>&1
  cat $OUT/phase_$v.txt
  for wl in c2 c4 c3; do
    timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-extra 2>$OUT/err_${wl}_$v.txt | tail -1 > $OUT/bench_${wl}_$v.json
    python -c "
import json; d=json.load(open('$OUT/bench_${wl}_$v.json')); print('$v $wl value %.1f e2e %.1f ms/step %.3f launch_us %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_us']), d['kernel_time_shares'].get('ms_one_pair_alone'))" || tail -5 $OUT/err_${wl}_$v.txt
  done
  timeout 300 python bench.py --workload c4 --streams 1 --pairs 1 --steps 5 --warmup 3 --no-extra 2>/dev/null | tail -1 > $OUT/bench_c4single_$v.json
  python -c "
import json; d=json.load(open('$OUT/bench_c4single_$v.json')); print('$v c4 single ms/step %.3f'%d['ms_per_step'])"
done
