"""Opt-in convergence diagnostics (residual norms by warp-shuffle reductions).  The
reference computes no norm, so the checker is the oracle's own extension (oracle_residual): same weights in
fp32, everything else in double; the only difference left is the summation order (rtol 1e-9)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def torch_():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.parametrize("w,h,hx,hy", [(96, 80, 1.0, 1.0), (200, 131, 1.37, 1.21), (33, 29, 2.0, 1.5), (300, 260, 1.0, 1.0)])
@pytest.mark.parametrize("constancy", [0, 1])
def test_stage_residual_vs_oracle(pkg, oracle, synth, torch_, w, h, hx, hy, constancy):
    f0, f1, _, _ = synth.make_pair(w, h, 11, U1=1.5)
    u = synth.smooth_random(w, h, 1, -1, 1)
    v = synth.smooth_random(w, h, 2, -1, 1)
    fl = pkg.Flow2D(w, h, constancy=constancy)
    p = pkg.default_params(outer=3, inner=7, alpha=20.0)
    c = [fl.to_container(a, 0.0) for a in (f0, f1, u, v)]
    du, dv, phi, ksi = (fl.container(0.0) for _ in range(4))
    fl.stage_solve(c[0], c[1], c[2], c[3], du, dv, phi, ksi, w, h, hx, hy, p)
    ru, rv = fl.stage_residual(c[0], c[1], c[2], c[3], du, dv, phi, ksi, w, h, hx, hy, p)
    op = oracle.make_params(outer=3, inner=7, alpha=20.0, constancy=constancy)
    odu, odv, ophi, oksi = oracle.solve_level(f0, f1, u, v, hx, hy, op)
    assert np.array_equal(fl.from_container(phi, w, h), ophi) and np.array_equal(fl.from_container(ksi, w, h), oksi)
    oru, orv = oracle.residual(f0, f1, u, v, odu, odv, ophi, oksi, hx, hy, 20.0, constancy)
    assert ru == pytest.approx(oru, rel=RTOL) and rv == pytest.approx(orv, rel=RTOL)
    assert ru > 0 and rv > 0


def test_residual_falls_with_more_sweeps(pkg, synth, torch_):
    """Jacobi converges slowly on the smooth modes, so the norm is only required to fall monotonically and to be far
    below its start after a few thousand sweeps of a small level."""
    w, h = 48, 40
    f0, f1, _, _ = synth.make_pair(w, h, 5, U1=1.0)
    fl = pkg.Flow2D(w, h)
    c0, c1 = fl.to_container(f0, 0.0), fl.to_container(f1, 0.0)
    z0, z1 = fl.container(0.0), fl.container(0.0)
    du, dv, phi, ksi = (fl.container(0.0) for _ in range(4))
    seen = []
    for inner in (2, 20, 200, 4000):
        p = pkg.default_params(outer=1, inner=inner, alpha=2.0)
        fl.stage_solve(c0, c1, z0, z1, du, dv, phi, ksi, w, h, 1.0, 1.0, p)
        seen.append(sum(fl.stage_residual(c0, c1, z0, z1, du, dv, phi, ksi, w, h, 1.0, 1.0, p)))
    assert seen[0] > seen[1] > seen[2] > seen[3]
    assert seen[3] < 0.05 * seen[0], seen


def test_level_residuals_of_a_flow_and_results_unchanged(pkg, oracle, synth, torch_):
    w, h = 160, 120
    f0, f1, _, _ = synth.make_pair(w, h, 9, U1=2.0)
    cfg = dict(levels=50, scale=0.8, outer=4, inner=5, alpha=20.0, median=3, sigma=1.0)
    fl = pkg.Flow2D(w, h)
    u0, v0 = fl.compute(f0, f1, pkg.default_params(**cfg))
    assert fl.level_residuals() == []
    p = pkg.default_params(**cfg)
    p.report_residuals = 1
    u1, v1 = fl.compute(f0, f1, p)
    assert np.array_equal(u0, u1) and np.array_equal(v0, v1)           # diagnostics never change the flow
    res = fl.level_residuals()
    assert len(res) == fl.stats()["levels_run"] == len(pkg.level_table(w, h, 0.8, 50))
    assert all(np.isfinite(a) and np.isfinite(b) and a >= 0 and b >= 0 for a, b in res)
    assert fl.launch_counts().get("residual") == len(res)
    # the finest level's number against the oracle: replay that level by hand
    u2, v2 = fl.compute(f0, f1, p)                                      # graph replay keeps reporting
    again = fl.level_residuals()                                        # (the order of the atomic adds may differ)
    assert len(again) == len(res) and all(a == pytest.approx(b, rel=1e-12) for ra, rb in zip(again, res) for a, b in zip(ra, rb))
    assert np.array_equal(u2, u1)
    ou, ov = oracle.compute_flow(f0, f1, oracle.make_params(**cfg))
    assert np.array_equal(u1, ou) and np.array_equal(v1, ov)


def test_level_times_and_results_unchanged(pkg, synth, torch_):
    """flow2d_params.report_level_times: the reference's per-level solve timer (cuda_operation_solve_2d.cpp:302-311)."""
    w, h = 320, 240
    f0, f1, _, _ = synth.make_pair(w, h, 9, U1=2.0)
    cfg = dict(levels=50, scale=0.8, outer=4, inner=5, alpha=20.0, median=3, sigma=1.0)
    fl = pkg.Flow2D(w, h)
    u0, v0 = fl.compute(f0, f1, pkg.default_params(**cfg))
    assert fl.level_times() == []
    p = pkg.default_params(**cfg)
    p.report_level_times = 1
    u1, v1 = fl.compute(f0, f1, p)
    assert np.array_equal(u0, u1) and np.array_equal(v0, v1)
    t = fl.level_times()
    assert len(t) == fl.stats()["levels_run"]
    assert all(lv > 0 and 0 < sv <= lv * 1.001 + 1e-3 for lv, sv in t), t
    assert sum(lv for lv, _ in t) <= fl.stats()["device_ms"] * 1.01 + 0.05
