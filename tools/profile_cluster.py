"""One level of the 1024^2 pyramid on the thread-block-cluster solve, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:solve_cluster -s 1 -c 1 -o rep python tools/profile_cluster.py 64 64
(the first launch warms up).  40 outer x 5 inner iterations, the reference's defaults."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import flow2d_loader  # noqa: E402

m = flow2d_loader.load()
from cuda_flow2d_b200 import synth  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else W
fl = m.Flow2D(W, H)
f0, f1, _, _ = synth.make_pair(W, H, 3, U1=1.5, L=48.0)
d = [fl.to_container(a) for a in (f0, f1, synth.smooth_random(W, H, 1, -2, 2), synth.smooth_random(W, H, 2, -2, 2))]
o = [fl.container(0.0) for _ in range(2)]
p = m.default_params(outer=40, inner=5)
for _ in range(3):
    fl.stage_solve(d[0], d[1], d[2], d[3], o[0], o[1], None, None, W, H, 1.0, 1.0, p)
torch.cuda.synchronize()
print("done", W, H, fl.launch_counts())
