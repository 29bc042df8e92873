"""Calls every stage of the path once (or a few times) at one image size, for per-kernel ncu metrics:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file out.csv python tools/profile_stages.py 4096 4096
Kernels, in launch order: blur, resample x/y (restriction 0.9), resample x/y (prolongation), warp,
derivatives, solve_pass x3 (first pass S=1 | first pass S=5 | later pass S=5 via inner=10 -> 2 passes), add_median<5>."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import flow2d_loader  # noqa: E402

m = flow2d_loader.load()
from cuda_flow2d_b200 import synth  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else W
fl = m.Flow2D(W, H)
rng = np.random.default_rng(0)
f0 = torch.from_numpy(rng.uniform(0, 255, (H, W)).astype(np.float32))
f1 = torch.from_numpy(rng.uniform(0, 255, (H, W)).astype(np.float32))
u = torch.from_numpy(synth.smooth_random(W, H, 1, -2, 2))
v = torch.from_numpy(synth.smooth_random(W, H, 2, -2, 2))
c = lambda t: fl.to_container(t.numpy(), 0.0)
d0, d1, du_, dv_ = c(f0), c(f1), c(u), c(v)
o = [fl.container(0.0) for _ in range(6)]
torch.cuda.synchronize()
cw, ch = int(np.ceil(W * 0.9)), int(np.ceil(H * 0.9))
fl.stage_blur(d0, o[0], W, H, 1.5)
fl.stage_resample(d0, W, H, o[0], cw, ch)
fl.stage_resample(o[0], cw, ch, o[1], W, H)
fl.stage_warp(d0, d1, du_, dv_, o[0], W, H, 1.0, 1.0)
fl.stage_solve(d0, d1, du_, dv_, o[1], o[2], None, None, W, H, 1.0, 1.0, m.default_params(outer=1, inner=1, sweeps_per_pass=1))
fl.stage_solve(d0, d1, du_, dv_, o[1], o[2], None, None, W, H, 1.0, 1.0, m.default_params(outer=1, inner=5, sweeps_per_pass=5))
fl.stage_solve(d0, d1, du_, dv_, o[1], o[2], None, None, W, H, 1.0, 1.0, m.default_params(outer=1, inner=10, sweeps_per_pass=5))
fl.stage_add_median(du_, o[1], o[3], W, H, 5)
torch.cuda.synchronize()
print("done", W, H)
