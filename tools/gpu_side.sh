#!/bin/bash
OUT=gpurun_out/${1:-side}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_flow_gpu.py tests/test_native_sizes_gpu.py tests/test_reference_gpu.py tests/test_ext_gpu.py tests/test_residual_gpu.py tests/test_cli_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee $OUT/pytest.log
for rep in 1 2; do
for v in side noside; do
  unset FLOW2D_NO_SIDE_PYRAMID
  [ $v = noside ] && export FLOW2D_NO_SIDE_PYRAMID=1
  for wl in c4 c4s c1b; do
    args="--workload $wl"; [ $wl = c4s ] && args="--workload c4 --streams 1 --pairs 1"
    timeout 300 python bench.py $args --steps 4 --warmup 2 --no-extra 2>$OUT/err_${wl}_$v.txt | tail -1 > $OUT/bench_${wl}_${v}_$rep.json
    python -c "
import json; d=json.load(open('$OUT/bench_${wl}_${v}_$rep.json')); print('$rep $wl $v value %.1f e2e %.1f ms/step %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step']))" || tail -3 $OUT/err_${wl}_$v.txt
  done
done
done
