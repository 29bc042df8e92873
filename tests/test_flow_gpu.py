"""End-to-end GPU parity of flow2d_compute / flow2d_compute_device against the CPU oracle and the
committed golden flows of the reference build (tests/golden/, generated on the GPU box by
tools/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import ROOT, epd_stats

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def torch_():
    import torch
    assert torch.cuda.is_available()
    return torch


CASES = [
    # w, h, params
    (96, 80, dict(levels=10, outer=10, inner=5, alpha=10.0, sigma=0.8, median=3)),
    (200, 150, dict(levels=50, outer=5, inner=5, alpha=35.0, sigma=1.5, median=5)),
    (64, 64, dict(levels=1, scale=0.5, outer=1, inner=40, alpha=0.25, e_smooth=1.0, e_data=1000.0, median=1, sigma=0.0)),  # C2-like
    (131, 67, dict(levels=6, outer=3, inner=12, alpha=5.0, sigma=0.45, median=7)),
]


@pytest.mark.parametrize("w,h,cfg", CASES)
@pytest.mark.parametrize("constancy", [0, 1])
def test_compute_vs_oracle(pkg, oracle, synth, torch_, w, h, cfg, constancy):
    f0, f1, _, _ = synth.make_pair(w, h, 77 + w, U0=(0.6, -0.4), U1=0.8, L=64.0)
    fl = pkg.Flow2D(w, h, constancy=constancy)
    u, v = fl.compute(f0, f1, pkg.default_params(**cfg))
    ou, ov = oracle.compute_flow(f0, f1, oracle.make_params(constancy=constancy, **cfg))
    s = epd_stats(u, v, ou, ov)
    print(s, fl.stats())
    assert np.all(u == ou) and np.all(v == ov), s


def test_rub_pair_vs_oracle_and_golden(pkg, oracle, rub, torch_):
    f0, f1 = rub
    fl = pkg.Flow2D(584, 388)
    u, v = fl.compute(f0, f1, pkg.default_params())
    ou, ov = oracle.compute_flow(f0, f1, oracle.make_params())
    s = epd_stats(u, v, ou, ov)
    print("C1b kernel vs oracle", s, fl.stats())
    assert np.all(u == ou) and np.all(v == ov), s
    g = os.path.join(ROOT, "tests", "golden", "rub_c1b_reference.npz")
    if os.path.exists(g):
        z = np.load(g)
        sg = epd_stats(u, v, z["u"], z["v"])
        print("C1b kernel vs committed reference golden", sg)
        assert sg["mean"] <= 1e-3 and sg["max"] <= 1e-2, sg


def test_compute_device_equals_compute(pkg, synth, torch_):
    w, h = 160, 120
    f0, f1, _, _ = synth.make_pair(w, h, 5)
    fl = pkg.Flow2D(w, h)
    p = pkg.default_params(levels=8, outer=4)
    u, v = fl.compute(f0, f1, p)
    d0, d1 = fl.to_container(f0, 0.0), fl.to_container(f1, 0.0)
    du, dv = fl.container(0.0), fl.container(0.0)
    s = torch_.cuda.Stream()
    fl.set_stream(s.cuda_stream)
    fl.compute_device(d0, d1, du, dv, p)
    s.synchronize()
    fl.set_stream(None)
    assert np.array_equal(fl.from_container(du, w, h), u) and np.array_equal(fl.from_container(dv, w, h), v)


def test_repeated_calls_are_deterministic(pkg, synth, torch_):
    w, h = 120, 90
    f0, f1, _, _ = synth.make_pair(w, h, 6)
    fl = pkg.Flow2D(w, h)
    p = pkg.default_params(levels=6, outer=3)
    a = fl.compute(f0, f1, p)
    g0, g1, _, _ = synth.make_pair(w, h, 8)
    fl.compute(g0, g1, p)  # different data in between: no state leaks from one pair to the next
    b = fl.compute(f0, f1, p)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_bad_parameters_are_errors(pkg, synth, torch_):
    fl = pkg.Flow2D(64, 64)
    f0, f1, _, _ = synth.make_pair(64, 64, 1)
    for bad in (dict(scale=1.0), dict(scale=0.0), dict(median=9), dict(median=0), dict(levels=0), dict(sigma=6.0)):
        with pytest.raises(pkg.Flow2DError):
            fl.compute(f0, f1, pkg.default_params(**bad))
    # and the handle is still usable afterwards
    fl.compute(f0, f1, pkg.default_params(levels=3, outer=2))


def test_full_size_properties(pkg, synth, torch_):
    """BASELINE sizes the oracle cannot finish quickly: size-independent properties instead.
    (1) schedule independence at 1024x1024, (2) a zero-motion pair gives exactly zero flow."""
    w = h = 1024
    f0, f1, _, _ = synth.make_pair(w, h, 1001, U0=(0.3, -0.2), U1=0.5, L=256.0)
    fl = pkg.Flow2D(w, h)
    c2 = dict(levels=1, scale=0.5, outer=1, inner=50, alpha=0.25, e_smooth=1.0, e_data=1000.0, median=1, sigma=0.0)
    a = fl.compute(f0, f1, pkg.default_params(sweeps_per_pass=5, **c2))
    b = fl.compute(f0, f1, pkg.default_params(sweeps_per_pass=2, **c2))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    z = fl.compute(f0, f0, pkg.default_params(levels=5, outer=2))
    assert not z[0].any() and not z[1].any()


def test_launch_counts_by_kernel(pkg, synth, torch_):
    """flow2d_last_launch_counts splits flow2d_last_stats' launch count by kernel; a replayed graph reports the same."""
    w, h = 200, 150
    f0, f1, _, _ = synth.make_pair(w, h, 2, U1=1.0)
    fl = pkg.Flow2D(w, h)
    p = pkg.default_params(levels=50, scale=0.8, outer=3, inner=5)
    fl.compute(f0, f1, p)
    counts = fl.launch_counts()
    assert sum(counts.values()) == fl.stats()["kernel_launches"]
    levels = fl.stats()["levels_run"]
    assert counts["warp"] == counts["derivatives"] == counts["add_median"] == levels
    # one x + one y launch per level for the two frames (on the side stream, ahead of the level loop) and one x + one y
    # launch per level for the two flow components; the finest level restricts nothing, the coarsest prolongates nothing
    assert counts["blur"] == 2 and counts["resample"] == 4 * (levels - 1)
    assert any(k.startswith("solve") for k in counts)
    fl.compute(f0, f1, p)
    assert fl.launch_counts() == counts


def test_graph_cache_keeps_alternating_buffers(pkg, synth, torch_):
    """A handle that alternates between a few (frames, flow) container sets -- bench.py's C4 batch, a frame ring -- must
    replay its captured graphs: captures == number of distinct sets, never a re-capture (flow2d_graph_stats)."""
    w, h = 160, 120
    p = pkg.default_params(levels=8, outer=3)
    fl = pkg.Flow2D(w, h)
    sets, expect = [], []
    for i in range(3):
        f0, f1, _, _ = synth.make_pair(w, h, 20 + i)
        expect.append(fl.compute(f0, f1, p))
        sets.append((fl.to_container(f0, 0.0), fl.to_container(f1, 0.0), fl.container(0.0), fl.container(0.0)))
    c0, _ = fl.graph_stats()
    for it in range(32):
        d0, d1, du, dv = sets[it % 3]
        fl.compute_device(d0, d1, du, dv, p)
    torch_.cuda.synchronize()
    c1, replays = fl.graph_stats()
    assert c1 - c0 == 3, (c0, c1)
    assert replays >= 29
    for (d0, d1, du, dv), (eu, ev) in zip(sets, expect):
        assert np.array_equal(fl.from_container(du, w, h), eu) and np.array_equal(fl.from_container(dv, w, h), ev)


def test_misaligned_containers_are_refused(pkg, synth, torch_):
    """Stage entry points take float4 paths: a pointer that is only 4-byte aligned is an INVALID_ARGUMENT, not a device fault."""
    w, h = 64, 48
    fl = pkg.Flow2D(w, h)
    big = torch_.zeros(h * fl.pitch + 8, dtype=torch_.float32, device="cuda")
    off = big[1:1 + h * fl.pitch].view(h, fl.pitch)  # 4-byte aligned only
    ok = [fl.container(0.0) for _ in range(5)]
    with pytest.raises(pkg.Flow2DError):
        fl.stage_warp(ok[0], ok[1], off, ok[2], ok[3], w, h, 1.0, 1.0)
    with pytest.raises(pkg.Flow2DError):
        fl.stage_blur(off, ok[0], w, h, 1.0)
    with pytest.raises(pkg.Flow2DError):
        fl.stage_median(ok[0], off, w, h, 5)
    with pytest.raises(pkg.Flow2DError):
        fl.stage_resample(off, w, h, ok[0], 32, 24)
    fl.stage_warp(ok[0], ok[1], ok[2], ok[3], ok[4], w, h, 1.0, 1.0)  # the handle still works
    torch_.cuda.synchronize()


@pytest.mark.parametrize("es,ed", [(0.0, 0.001), (0.001, 0.0)])
def test_zero_epsilon_is_exact(pkg, oracle, synth, torch_, es, ed):
    """equation_smoothness = 0 or equation_data = 0 are legal: the one-pixel kernels take their plain IEEE variant."""
    w, h = 96, 80
    f0, f1, _, _ = synth.make_pair(w, h, 31, U0=(0.4, -0.2), U1=0.6, L=48.0)
    cfg = dict(levels=6, outer=4, inner=5, alpha=10.0, sigma=0.8, median=3, e_smooth=es, e_data=ed)
    fl = pkg.Flow2D(w, h)
    u, v = fl.compute(f0, f1, pkg.default_params(**cfg))
    ou, ov = oracle.compute_flow(f0, f1, oracle.make_params(**cfg))
    same = (u == ou) | (np.isnan(u) & np.isnan(ou))
    assert same.all() and ((v == ov) | (np.isnan(v) & np.isnan(ov))).all()


def test_prepare_captures_the_schedule_ahead(pkg, oracle, synth, torch_):
    """flow2d_prepare: the graph of the host-API path is captured before the first call, which then only replays."""
    w, h = 200, 150
    f0, f1, _, _ = synth.make_pair(w, h, 4, U1=1.0)
    cfg = dict(levels=50, scale=0.8, outer=3, inner=5)
    fl = pkg.Flow2D(w, h)
    p = pkg.default_params(**cfg)
    fl.prepare(p)
    assert fl.graph_stats() == (1, 0)
    u, v = fl.compute(f0, f1, p)
    assert fl.graph_stats() == (1, 1)
    ou, ov = oracle.compute_flow(f0, f1, oracle.make_params(**cfg))
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
