"""Small end-to-end flows (Grey and Gradient, tiled / resident / tiny / small-pass / slab paths) for
compute-sanitizer:  compute-sanitizer --tool racecheck|memcheck|initcheck python tools/sanitize.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import flow2d_loader  # noqa: E402

m = flow2d_loader.load()
from cuda_flow2d_b200 import synth  # noqa: E402

for constancy in (0, 1):
    for (w, h), over in (((150, 110), dict(levels=12, outer=3, inner=5)),
                         ((150, 110), dict(levels=3, outer=2, inner=5, resident_levels=-1)),
                         ((96, 80), dict(levels=2, outer=1, inner=9, sweeps_per_pass=4, resident_levels=-1))):
        f0, f1, _, _ = synth.make_pair(w, h, 1)
        fl = m.Flow2D(w, h, constancy=constancy)
        u, v = fl.compute(f0, f1, m.default_params(**over))
        print("ok", constancy, w, h, over, float(abs(u).mean()))
        fl.destroy()
# the TMA-staged persistent pass (aligned and 2-column staging boxes, first and later passes) on a level large enough for
# several tiles per CTA, and the opt-in extensions (red-black SOR on tensor planes, convergence test, cascaded restriction)
f0, f1, _, _ = synth.make_pair(600, 420, 2)
lo = min(f0.min(), f1.min())
f0, f1 = f0 - lo, f1 - lo
for over in (dict(levels=2, outer=2, inner=5, resident_levels=-1), dict(levels=1, outer=1, inner=14, sweeps_per_pass=7, resident_levels=-1),
             dict(levels=3, outer=3, inner=4, scheme=1, omega=1.6, data_term=3, gamma=2.0, residual_tolerance=0.5, cascaded_restriction=1),
             dict(levels=3, outer=4, inner=5, residual_tolerance=0.5)):
    fl = m.Flow2D(600, 420)
    u, v = fl.compute(f0, f1, m.default_params(**over))
    print("ok", over, float(abs(u).mean()), fl.launch_counts())
    fl.destroy()
torch.cuda.synchronize()
print("done")
