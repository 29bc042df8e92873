#!/bin/bash
# multi-GPU checks: slab test across real GPUs + C5 strong scaling + C4 weak scaling.  Usage: gpurun --gpus N -- bash tools/gpu_n.sh tag N
TAG=${1:-n2}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt
timeout 300 python -m pytest tests/test_slab_multigpu.py -m gpu -q -x -s 2>&1 | tail -30 > $OUT/pytest_multigpu.log; tail -12 $OUT/pytest_multigpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload c5 --steps 2 --warmup 1 2>$OUT/c5.err | tail -1 > $OUT/bench_c5_n$N.json
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_c5_n$N.json")); print("c5 N=$N ms/step %.1f e2e ms %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"])); print(d["schedule"]["sharding"])
except Exception as e:
    print("c5 parse", e); print(open("$OUT/c5.err").read()[-3000:])
PY
if [ "${3:-}" = "c4" ]; then
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 --no-extra 2>$OUT/c4.err | tail -1 > $OUT/bench_c4_n$N.json
python -c "
import json; d=json.load(open('$OUT/bench_c4_n$N.json')); print('c4 N=$N value %.1f e2e %.1f'%(d['value'], d['e2e']['value']))"
fi
