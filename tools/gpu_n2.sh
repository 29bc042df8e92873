#!/bin/bash
OUT=gpurun_out/n2_final
mkdir -p $OUT
timeout 300 python -m pytest tests/test_flow_gpu.py -m gpu -q -k launch_counts 2>&1 | tail -2
for wl in c2 c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/bench_${wl}_n2.json
  python -c "
import json; d=json.load(open('$OUT/bench_${wl}_n2.json')); print('$wl n2 %.1f Mpix/s  %.3f ms/step e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --workload c2 --steps 3 --warmup 1 2>&1 | tail -1 > $OUT/bench_ref_c2_n2.json; cut -c1-200 $OUT/bench_ref_c2_n2.json
timeout 600 python bench.py --workload c5 --steps 3 --warmup 1 2>&1 | tail -1 > $OUT/bench_c5_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c5 --steps 3 --warmup 1 2>&1 | tail -1 > $OUT/bench_c5_n2.json
python -c "
import json
for n in (1,2):
    d=json.load(open('$OUT/bench_c5_n%d.json'%n)); print('c5 n%d %.1f ms/step %.1f Mpix/s' % (n, d['ms_per_step'], d['value']))"
