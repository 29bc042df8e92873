"""The row-slab decomposition across REAL GPUs, one process per GPU (torchrun, NCCL for the wiring and the barrier;
the halo rows themselves travel as peer-mapped stores + flags, cuda-flow2d_b200/csrc/slab.cu): every rank's rows of the
flow must equal the single-GPU flow bit for bit.  Needs >= 2 GPUs (skipped otherwise; run with
`gpurun --gpus 2 -- python -m pytest tests/test_slab_multigpu.py -m gpu`), the 1-GPU box runs the same code with N
logical ranks on one device in tests/test_slab_gpu.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["FLOW2D_ROOT"])
import flow2d_loader
m = flow2d_loader.load()
from cuda_flow2d_b200 import synth, slab as S
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w, h = 320, 1400
for constancy in (0, 1):
    f0, f1, _, _ = synth.make_pair(w, h, 11, U0=(0.5, -0.3), U1=1.2, L=96.0)
    p = m.default_params(levels=10, outer=12, inner=5, alpha=20.0, median=5)
    fl = m.Flow2D(w, h, constancy=constancy, device=local)
    eu, ev = fl.compute(f0, f1, p)                     # single-GPU flow of the same frames
    S.connect_ipc(fl, dist, rank, world, "cuda:%d" % local, min_rows=64)
    d0, d1 = fl.to_container(f0, 0.0), fl.to_container(f1, 0.0)
    du, dv = fl.container(float("nan")), fl.container(float("nan"))
    for rep in range(2):
        torch.cuda.synchronize()
        dist.barrier()
        fl.compute_slab_device(d0, d1, du, dv, p)
        fl.slab_status()
        y0, y1 = fl.slab_rows()
        gu, gv = fl.from_container(du, w, h), fl.from_container(dv, w, h)
        ok = np.array_equal(gu[y0:y1], eu[y0:y1]) and np.array_equal(gv[y0:y1], ev[y0:y1])
        st = fl.slab_stats()
        print("rank %d constancy %d rep %d rows [%d,%d) exact=%s %s" % (rank, constancy, rep, y0, y1, ok, st), flush=True)
        assert ok and st["levels_slabbed"] >= 2 and st["exchanges"] > 0
    dist.barrier()
    fl.destroy()
dist.destroy_process_group()
print("SLAB_MULTIGPU_OK rank %d" % rank, flush=True)
'''


def test_slab_across_gpus_equals_single_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (this box has %d); the single-GPU variant is tests/test_slab_gpu.py" % n)
    world = 4 if n >= 4 else 2
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, FLOW2D_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    out = r.stdout.decode()
    print(out[-4000:])
    assert r.returncode == 0
    assert out.count("SLAB_MULTIGPU_OK") == world
