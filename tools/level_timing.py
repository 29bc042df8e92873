"""Per-level cost of the solve as the scheduler runs it: flow2d_stage_solve on one level of a given size with the
C1b iteration counts (40 x 5), timed with CUDA events; prints us per outer iteration and which kernel was used.
Ground truth for the scheduler's time model (flow2d_api.cu: run_solve).   python tools/level_timing.py [throughput=0]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import flow2d_loader  # noqa: E402

m = flow2d_loader.load()
from cuda_flow2d_b200 import synth  # noqa: E402

throughput = int(sys.argv[1]) if len(sys.argv) > 1 else 0
sizes = [(12, 8), (24, 16), (36, 27), (40, 30), (48, 36), (59, 44), (72, 48), (96, 64), (128, 85), (160, 107), (200, 133), (240, 160),
         (280, 186), (330, 220), (400, 266), (473, 314), (584, 388), (800, 600), (1024, 1024)]
fl = m.Flow2D(1024, 1024)
stream = torch.cuda.Stream()
fl.set_stream(stream.cuda_stream)
for (w, h) in sizes:
    f0, f1, _, _ = synth.make_pair(w, h, 3, U1=1.0)
    c0, c1 = fl.to_container(np.pad(f0, ((0, 1024 - h), (0, 1024 - w))), 0.0), fl.to_container(np.pad(f1, ((0, 1024 - h), (0, 1024 - w))), 0.0)
    z0, z1 = fl.container(0.0), fl.container(0.0)
    du, dv = fl.container(0.0), fl.container(0.0)
    p = m.default_params(outer=40, inner=5, alpha=35.0)
    p.throughput_mode = throughput
    n0 = fl.launch_counts()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        fl.stage_solve(c0, c1, z0, z1, du, dv, None, None, w, h, 1.0, 1.0, p)  # warm-up
        n1 = fl.launch_counts()
        a.record(stream)
        for _ in range(5):
            fl.stage_solve(c0, c1, z0, z1, du, dv, None, None, w, h, 1.0, 1.0, p)
        b.record(stream)
    torch.cuda.synchronize()
    used = {k: n1.get(k, 0) - n0.get(k, 0) for k in n1 if k.startswith("solve") and n1.get(k, 0) > n0.get(k, 0)}
    us = a.elapsed_time(b) * 1e3 / 5
    print("%4dx%-4d %7d px  %8.1f us per level  %6.2f us per outer iteration  %s" % (w, h, w * h, us, us / 40, used), flush=True)
