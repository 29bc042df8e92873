"""The thread-block-cluster solve (csrc/solve_cluster.cu) against the oracle and against the other solve kernels, bit for bit.

Whole-level mode: a level of up to 16 x 1024 pixels runs ALL outer iterations on one cluster of 2-16 CTAs (halos through
distributed shared memory, barrier.cluster per sweep).  By default the scheduler uses it when four or more handles are alive
on the device (several frame pairs in flight: least SM time) and not for a lone handle (least latency; profiles/r02/cluster_ab);
FLOW2D_CLUSTER=2 forces it for every level that fits a cluster, =1 where a CTA's block fits 256 threads, =0 never.  Pass mode (opt-in, measured slower): a grid of clusters, one pass per launch,
128x64 / 128x128 regions.  The switches (FLOW2D_CLUSTER, FLOW2D_CLUSTER_PASS, ...) are read when a handle is created."""
import contextlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_():
    import torch
    assert torch.cuda.is_available()
    return torch


@contextlib.contextmanager
def _env(**kv):
    old = {k: os.environ.get(k) for k in kv}
    try:
        for k, v in kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _handle(pkg, w, h, constancy=0, **env):
    base = dict(FLOW2D_CLUSTER=None, FLOW2D_CLUSTER_PASS=None, FLOW2D_CLUSTER_COMPACT=None, FLOW2D_CLUSTER_MAX=None)
    base.update(env)
    with _env(**base):
        return pkg.Flow2D(max(w, 4), max(h, 4), constancy=constancy)


def _eq(a, b):
    return np.array_equal(a, b)  # (-0.0 == 0.0: the sign of a zero is the only tolerated difference)


def _solve_inputs(synth, w, h, seed):
    f0, f1, _, _ = synth.make_pair(w, h, seed, U1=1.5, L=48.0)
    u = synth.smooth_random(w, h, seed + 1, -2, 2)
    v = synth.smooth_random(w, h, seed + 2, -2, 2)
    return f0, f1, u, v


def _run_solve(torch, fl, f0, f1, u, v, w, h, hx, hy, params):
    d = [fl.to_container(a) for a in (f0, f1, u, v)]
    d_du, d_dv, d_phi, d_ksi = (fl.container(float("nan")) for _ in range(4))
    fl.stage_solve(d[0], d[1], d[2], d[3], d_du, d_dv, d_phi, d_ksi, w, h, hx, hy, params)
    torch.cuda.synchronize()
    return tuple(fl.from_container(t, w, h) for t in (d_du, d_dv, d_phi, d_ksi))


WHOLE_CASES = [
    # w, h, hx, hy, outer, inner
    (40, 30, 1.0, 1.0, 3, 5),          # 1200 px: the smallest levels above solve_tiny
    (33, 32, 1.3, 1.1, 2, 5),
    (64, 64, 1.0, 1.0, 4, 5),          # 16 x 256 threads
    (100, 90, 1.5, 1.5, 2, 7),         # 16 x 1024 threads, ragged blocks
    (128, 128, 1.0, 1.0, 2, 5),        # the largest square level
    (127, 128, 2.92, 2.425, 3, 1),     # one sweep per outer iteration
    (9, 1024, 1.0, 1.0, 2, 2),         # a strip: blocks of 9 x 64
    (1024, 16, 1.0, 1.0, 2, 3),
    (71, 113, 7.3, 3.4, 5, 4),
    (91, 82, 11.2, 12.5, 40, 5),       # a level of the 1024^2 pyramid with the reference's iteration counts
]


@pytest.mark.parametrize("w,h,hx,hy,outer,inner", WHOLE_CASES)
@pytest.mark.parametrize("constancy", [0, 1])
def test_cluster_whole_level_vs_oracle(pkg, oracle, synth, torch_, w, h, hx, hy, outer, inner, constancy):
    f0, f1, u, v = _solve_inputs(synth, w, h, 300 + w)
    fl = _handle(pkg, w, h, constancy, FLOW2D_CLUSTER=2)
    p = pkg.default_params(outer=outer, inner=inner, alpha=20.0)
    du, dv, phi, ksi = _run_solve(torch_, fl, f0, f1, u, v, w, h, hx, hy, p)
    counts = fl.launch_counts()
    assert counts.get("solve_cluster", 0) == 1 and not any(k.startswith("solve_") and k != "solve_cluster" for k in counts), counts
    op = oracle.make_params(outer=outer, inner=inner, alpha=20.0, constancy=constancy)
    edu, edv, ephi, eksi = oracle.solve_level(f0, f1, u, v, hx, hy, op)
    assert _eq(phi, ephi) and _eq(ksi, eksi)
    assert _eq(du, edu) and _eq(dv, edv)


@pytest.mark.parametrize("env", [dict(FLOW2D_CLUSTER_COMPACT=1), dict(FLOW2D_CLUSTER_MAX=8), dict(FLOW2D_CLUSTER_MAX=2)])
@pytest.mark.parametrize("w,h", [(45, 40), (64, 31), (90, 80)])
def test_cluster_shapes_agree(pkg, synth, torch_, env, w, h):
    """every cluster shape (few big CTAs, portable cluster sizes only, pairs) gives the bits of the other solve kernels"""
    f0, f1, u, v = _solve_inputs(synth, w, h, 11)
    p = pkg.default_params(outer=3, inner=5, alpha=20.0)
    ref = _run_solve(torch_, _handle(pkg, w, h, FLOW2D_CLUSTER=0), f0, f1, u, v, w, h, 1.4, 1.2, p)
    fl = _handle(pkg, w, h, FLOW2D_CLUSTER=2, **env)
    got = _run_solve(torch_, fl, f0, f1, u, v, w, h, 1.4, 1.2, p)
    if w * h <= int(env.get("FLOW2D_CLUSTER_MAX", 16)) * 1024:
        assert fl.launch_counts().get("solve_cluster", 0) == 1
    for k in range(4):
        assert _eq(got[k], ref[k])


def test_cluster_exact_variant(pkg, oracle, synth, torch_):
    """e_smooth = 0: flat cells make the argument of sqrt exactly zero; the plain IEEE variant runs from the start"""
    w, h = 60, 50
    f0, f1, u, v = _solve_inputs(synth, w, h, 5)
    f0[:, :20] = 3.0
    f1[:, :20] = 3.0
    u[:, :20] = 0.0
    v[:, :20] = 0.0
    fl = _handle(pkg, w, h, FLOW2D_CLUSTER=2)
    p = pkg.default_params(outer=3, inner=5, alpha=20.0, e_smooth=0.0)
    du, dv, phi, ksi = _run_solve(torch_, fl, f0, f1, u, v, w, h, 1.0, 1.0, p)
    assert fl.launch_counts().get("solve_cluster", 0) == 1
    op = oracle.make_params(outer=3, inner=5, alpha=20.0, e_smooth=0.0)
    edu, edv, ephi, eksi = oracle.solve_level(f0, f1, u, v, 1.0, 1.0, op)
    with np.errstate(invalid="ignore"):
        assert np.array_equal(du, edu, equal_nan=True) and np.array_equal(dv, edv, equal_nan=True)


def test_cluster_redo_vote_is_cluster_wide(pkg, oracle, synth, torch_):
    """A dividend outside the range of the branch-free division in ONE block of the cluster: every CTA repeats the outer
    iteration with IEEE divisions (a CTA that did not would leave the cluster barriers out of step)."""
    w, h = 96, 96
    f0, f1, u, v = _solve_inputs(synth, w, h, 8)
    f1 = f1.copy()
    f1[5, 7] = 1.0e16  # ft ~ 1e16 there: ft^2 > 2^100, outside the range of the branch-free sqrt, in that block only
                       # (everything stays finite: the oracle's result has no inf / nan)
    fl = _handle(pkg, w, h, FLOW2D_CLUSTER=2)
    p = pkg.default_params(outer=3, inner=5, alpha=20.0)
    du, dv, phi, ksi = _run_solve(torch_, fl, f0, f1, u, v, w, h, 1.0, 1.0, p)
    assert fl.launch_counts().get("solve_cluster", 0) == 1
    op = oracle.make_params(outer=3, inner=5, alpha=20.0)
    edu, edv, ephi, eksi = oracle.solve_level(f0, f1, u, v, 1.0, 1.0, op)
    assert np.isfinite(edu).all() and np.isfinite(eksi).all()
    assert _eq(du, edu) and _eq(dv, edv) and _eq(ksi, eksi) and _eq(phi, ephi)


@pytest.mark.parametrize("w,h,expect", [(30, 30, False), (36, 36, True), (64, 64, True), (60, 70, False), (91, 91, False)])
def test_cluster_mode_1_policy(pkg, oracle, synth, torch_, w, h, expect):
    """FLOW2D_CLUSTER=1: solve_tiny up to 1 024 px, the cluster up to 4 096 px (blocks of <= 256 threads), passes above"""
    f0, f1, u, v = _solve_inputs(synth, w, h, 77)
    fl = _handle(pkg, w, h, FLOW2D_CLUSTER=1)
    p = pkg.default_params(outer=4, inner=5, alpha=20.0)
    du, dv, phi, ksi = _run_solve(torch_, fl, f0, f1, u, v, w, h, 1.2, 1.3, p)
    assert (fl.launch_counts().get("solve_cluster", 0) == 1) == expect, fl.launch_counts()
    edu, edv, ephi, eksi = oracle.solve_level(f0, f1, u, v, 1.2, 1.3, oracle.make_params(outer=4, inner=5, alpha=20.0))
    assert _eq(du, edu) and _eq(dv, edv) and _eq(phi, ephi) and _eq(ksi, eksi)


def test_cluster_default_follows_the_handle_count(pkg, oracle, synth, torch_):
    """FLOW2D_CLUSTER unset: from four handles alive on the device on (frame pairs in flight at once: SM time counts) a level
    that fits a cluster runs on one; a lone handle keeps one launch per outer iteration (latency counts).  Same bits."""
    w, h = 91, 82
    f0, f1, u, v = _solve_inputs(synth, w, h, 31)
    p = pkg.default_params(outer=5, inner=5, alpha=20.0)
    edu, edv, ephi, eksi = oracle.solve_level(f0, f1, u, v, 1.2, 1.3, oracle.make_params(outer=5, inner=5, alpha=20.0))
    fl = _handle(pkg, w, h)
    extra = []
    try:
        while pkg.live_handles(0) < 4:
            extra.append(_handle(pkg, 8, 8))
        du, dv, phi, ksi = _run_solve(torch_, fl, f0, f1, u, v, w, h, 1.2, 1.3, p)
        assert fl.launch_counts().get("solve_cluster", 0) == 1, fl.launch_counts()
        assert _eq(du, edu) and _eq(dv, edv) and _eq(phi, ephi) and _eq(ksi, eksi)
    finally:
        for e in extra:
            e.destroy()
    fl.destroy()
    lone = _handle(pkg, w, h)  # (launch counts add up over the stage calls of a handle: a fresh one)
    if pkg.live_handles(0) < 4:  # (handles of earlier tests may still be waiting for the garbage collector)
        du, dv, phi, ksi = _run_solve(torch_, lone, f0, f1, u, v, w, h, 1.2, 1.3, p)
        counts = lone.launch_counts()
        assert counts.get("solve_cluster", 0) == 0 and any(k.startswith("solve") for k in counts), counts
        assert _eq(du, edu) and _eq(dv, edv) and _eq(phi, ephi) and _eq(ksi, eksi)


PASS_CASES = [(200, 150, 5), (333, 250, 3), (131, 67, 5), (260, 300, 7), (150, 140, 1)]


@pytest.mark.parametrize("w,h,inner", PASS_CASES)
@pytest.mark.parametrize("mode", [3, 4])
@pytest.mark.parametrize("constancy", [0, 1])
def test_cluster_pass_equals_tiled_pass(pkg, synth, torch_, w, h, inner, mode, constancy):
    """pass mode (forced: 3 = 128x64 regions of 16 x 512 threads, 4 = 128x128 of 16 x 1024) against the 64x48 tiled pass"""
    f0, f1, u, v = _solve_inputs(synth, w, h, 21)
    a_fl = _handle(pkg, w, h, constancy, FLOW2D_CLUSTER=0, FLOW2D_CLUSTER_PASS=mode)
    a = _run_solve(torch_, a_fl, f0, f1, u, v, w, h, 1.7, 1.4, pkg.default_params(outer=3, inner=inner, resident_levels=0))
    assert a_fl.launch_counts().get("solve_cluster", 0) == 3, a_fl.launch_counts()
    b = _run_solve(torch_, _handle(pkg, w, h, constancy, FLOW2D_CLUSTER=0), f0, f1, u, v, w, h, 1.7, 1.4,
                   pkg.default_params(outer=3, inner=inner, resident_levels=-1))
    for k in range(4):
        assert _eq(a[k], b[k])


@pytest.mark.parametrize("env", [dict(), dict(FLOW2D_CLUSTER=2), dict(FLOW2D_CLUSTER=2, FLOW2D_CLUSTER_PASS=1),
                                 dict(FLOW2D_CLUSTER=2, FLOW2D_CLUSTER_PASS=2), dict(FLOW2D_CLUSTER=2, FLOW2D_CLUSTER_COMPACT=1)])
def test_cluster_full_flow_rub_pair(pkg, oracle, rub, env):
    """the complete 47-level flow of the reference's own frame pair with the cluster kernels in the schedule, inside the
    replayed CUDA graph: bit-identical to the oracle (which is pinned to the reference build, tests/golden)"""
    f0, f1 = rub
    h, w = f0.shape
    cfg = dict(levels=50, outer=10, inner=5, alpha=35.0, sigma=1.5, median=5)
    fl = _handle(pkg, w, h, **env)
    extra = []
    while not env and pkg.live_handles(0) < 4:  # no switch set: the cluster schedule needs four handles alive on the device
        extra.append(_handle(pkg, 8, 8))
    u, v = fl.compute(f0, f1, pkg.default_params(**cfg))
    counts = fl.launch_counts()
    for e in extra:
        e.destroy()
    assert counts.get("solve_cluster", 0) >= 10, counts
    u2, v2 = fl.compute(f0, f1, pkg.default_params(**cfg))  # the replayed graph
    ou, ov = oracle.compute_flow(f0, f1, oracle.make_params(**cfg))
    assert _eq(u, ou) and _eq(v, ov)
    assert _eq(u2, ou) and _eq(v2, ov)


def test_cluster_flow_equals_default_flow_c4_pair(pkg, synth):
    """one 1024^2 pair with the reference's iteration counts (BASELINE configs[3]): cluster schedule == default schedule"""
    w = h = 1024
    f0, f1, _, _ = synth.make_pair(w, h, 3, U0=(0.5, -0.3), U1=0.7, L=64.0)
    cfg = dict(levels=50, outer=40, inner=5, alpha=20.0, sigma=1.0, median=5)
    base = _handle(pkg, w, h, FLOW2D_CLUSTER=0).compute(f0, f1, pkg.default_params(**cfg))
    fl = _handle(pkg, w, h, FLOW2D_CLUSTER=2)
    got = fl.compute(f0, f1, pkg.default_params(**cfg))
    assert fl.launch_counts().get("solve_cluster", 0) >= 10
    assert _eq(got[0], base[0]) and _eq(got[1], base[1])
    fl2 = _handle(pkg, w, h, FLOW2D_CLUSTER=2, FLOW2D_CLUSTER_PASS=1)
    got2 = fl2.compute(f0, f1, pkg.default_params(**cfg))
    assert _eq(got2[0], base[0]) and _eq(got2[1], base[1])
