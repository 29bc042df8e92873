"""Per-phase timeline of one solve_pass launch from in-kernel %globaltimer stamps (debug aid).
    python tools/phase_timing.py W H [outer inner sweeps_per_pass]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import flow2d_loader  # noqa: E402

m = flow2d_loader.load()
from cuda_flow2d_b200 import synth  # noqa: E402

W, H = int(sys.argv[1]), int(sys.argv[2])
outer = int(sys.argv[3]) if len(sys.argv) > 3 else 1
inner = int(sys.argv[4]) if len(sys.argv) > 4 else 5
spp = int(sys.argv[5]) if len(sys.argv) > 5 else 0
fl = m.Flow2D(W, H)
f0, f1, _, _ = synth.make_pair(W, H, 1)
u, v = synth.smooth_random(W, H, 1, -2, 2), synth.smooth_random(W, H, 2, -2, 2)
d = [fl.to_container(a, 0.0) for a in (f0, f1, u, v)]
o = [fl.container(0.0) for _ in range(2)]
stamps = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
p = m.default_params(outer=outer, inner=inner, sweeps_per_pass=spp, resident_levels=-1 if len(sys.argv) > 6 else 0)
for rep in range(3):
    fl.stage_solve(d[0], d[1], d[2], d[3], o[0], o[1], None, None, W, H, 1.0, 1.0, p)
torch.cuda.synchronize()
fl.debug_timing(stamps)
stamps.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
fl.stage_solve(d[0], d[1], d[2], d[3], o[0], o[1], None, None, W, H, 1.0, 1.0, p)
e1.record()
torch.cuda.synchronize()
s = stamps.cpu().numpy().reshape(-1, 8)
s = s[s[:, 0] > 0]
t0 = s[:, 0].min()
rel = (s[:, :6] - t0) / 1000.0
names = ["entry", "A loads+tensor+ksi", "B phi", "C weights", "D sweeps", "E stores"]
print("%dx%d outer=%d inner=%d: last launch had %d CTAs, stage_solve total %.1f us" % (W, H, outer, inner, len(s), e0.elapsed_time(e1) * 1e3))
print("CTA start spread: %.2f us; launch end %.2f us after first entry" % (rel[:, 0].max(), rel[:, 5].max()))
dur = np.diff(rel, axis=1)
for i in range(5):
    print("  %-22s mean %6.2f us   min %6.2f   max %6.2f" % (names[i + 1], dur[:, i].mean(), dur[:, i].min(), dur[:, i].max()))
print("  per-CTA total          mean %6.2f us   min %6.2f   max %6.2f" % ((rel[:, 5] - rel[:, 0]).mean(), (rel[:, 5] - rel[:, 0]).min(), (rel[:, 5] - rel[:, 0]).max()))
