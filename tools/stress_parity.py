"""Randomised parity sweep: random level sizes, parameters, data terms and input statistics (including inputs that
push the guarded fast-path divisions onto their slow path: constant frames, huge and tiny intensities, zero flow),
CUDA path vs the CPU oracle, bit for bit.  Not part of the test suite (it takes a GPU minute); run after kernel
changes:    python tools/stress_parity.py [cases=80] [seed=0]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import flow2d_loader  # noqa: E402

m = flow2d_loader.load()
from cuda_flow2d_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 80
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for c in range(cases):
    w, h = int(rng.integers(4, 420)), int(rng.integers(4, 330))
    constancy = int(rng.integers(0, 2))
    kind = ["synth", "noise", "const", "huge", "tiny", "steps"][int(rng.integers(0, 6))]
    if kind == "synth":
        f0, f1, _, _ = synth.make_pair(w, h, int(rng.integers(0, 1000)), U1=float(rng.uniform(0, 4)))
    elif kind == "noise":
        f0, f1 = (rng.uniform(0, 255, (h, w)).astype(np.float32) for _ in range(2))
    elif kind == "const":
        f0 = np.full((h, w), float(rng.uniform(0, 255)), np.float32)
        f1 = f0.copy() if rng.integers(0, 2) else np.full((h, w), float(rng.uniform(0, 255)), np.float32)
    elif kind == "huge":
        f0, f1 = (rng.uniform(0, 1, (h, w)).astype(np.float32) * np.float32(10.0 ** rng.integers(4, 13)) for _ in range(2))
    elif kind == "tiny":
        f0, f1 = (rng.uniform(0, 1, (h, w)).astype(np.float32) * np.float32(10.0 ** -rng.integers(6, 30)) for _ in range(2))
    else:
        f0 = (rng.integers(0, 2, (h, w)) * 255).astype(np.float32)
        f1 = np.roll(f0, (int(rng.integers(-3, 4)), int(rng.integers(-3, 4))), (0, 1))
    cfg = dict(levels=int(rng.integers(1, 60)), scale=float(rng.choice([0.5, 0.7, 0.8, 0.9, 0.95])), outer=int(rng.integers(1, 5)),
               inner=int(rng.integers(1, 12)), alpha=float(10.0 ** rng.uniform(-1, 2.5)), e_smooth=float(10.0 ** rng.uniform(-4, 0)),
               e_data=float(10.0 ** rng.uniform(-4, 3)), median=int(rng.choice([1, 3, 5, 7])), sigma=float(rng.choice([0.0, 0.6, 1.5, 3.0])))
    knobs = dict(sweeps_per_pass=int(rng.choice([0, 0, 1, 2, 3, 5, 7])), resident_levels=int(rng.choice([0, 0, 0, 2, -1])),
                 throughput_mode=int(rng.integers(0, 2)))
    fl = m.Flow2D(w, h, constancy=constancy)
    p = m.default_params(**cfg, **knobs)
    p.report_residuals = int(rng.integers(0, 2))
    u, v = fl.compute(f0, f1, p)
    ou, ov = O.compute_flow(f0, f1, O.make_params(constancy=constancy, **cfg))
    same = (np.array_equal(u, ou, equal_nan=True) and np.array_equal(v, ov, equal_nan=True))
    if not same:
        bad += 1
        d = np.nanmax(np.abs(u.astype(np.float64) - ou)) if np.isfinite(ou).any() else float("nan")
        print("MISMATCH case %d: %dx%d constancy %d %s %s %s  n_diff %d  max |du| %.3e" %
              (c, w, h, constancy, kind, cfg, knobs, int((u != ou).sum() + (v != ov).sum()), d), flush=True)
    fl.destroy()
print("stress_parity: %d cases, %d mismatches" % (cases, bad))
sys.exit(1 if bad else 0)
