"""Minimal driver for ncu: runs `warmup` + `steps` flow2d_compute_device steps of one bench workload and
nothing else on the GPU (no L2 flush kernels, no host API copies).
    ncu ... python tools/profile_step.py c2 1 1"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import flow2d_loader  # noqa: E402

m = flow2d_loader.load()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
warmup = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
over = {}
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    over[k] = int(v)
wl = bench.WORKLOADS[name]
f0, f1 = bench.make_frames(wl, 0)
fl = m.Flow2D(wl["w"], wl["h"])
p = m.default_params(**wl["cfg"], **over)
d0, d1 = fl.to_container(f0, 0.0), fl.to_container(f1, 0.0)
du, dv = fl.container(0.0), fl.container(0.0)
torch.cuda.synchronize()
for _ in range(warmup + steps):
    fl.compute_device(d0, d1, du, dv, p)
    torch.cuda.synchronize()
print("launches per step:", fl.stats()["kernel_launches"])
