#!/bin/bash
# all GPU tests with per-file time limits + single-GPU C5 / C4 numbers.   Usage: gpurun -- bash tools/gpu_check.sh tag
OUT=gpurun_out/${1:-chk}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -15 > $OUT/pytest.log; tail -6 $OUT/pytest.log
for wl in c5 c4; do
timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-extra 2>$OUT/err_$wl.txt | tail -1 > $OUT/bench_$wl.json
python -c "
import json; d=json.load(open('$OUT/bench_$wl.json')); print('$wl value %.1f e2e %.1f ms/step %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d.get('launches_by_kernel'))" || tail -5 $OUT/err_$wl.txt
done
