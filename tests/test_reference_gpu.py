"""GPU parity tests against the reference's OWN CUDA build (oracle/_ref, compiled from
/root/reference by `make -C oracle ref`; the binaries travel to the GPU box, the sources do not).

Two things are checked on identical inputs, operator by operator and end to end:
  * the CPU oracle equals the reference kernels  -> this is what pins the oracle, and
  * the sm_100a kernels (through the C ABI) equal the reference kernels.
Index/selection work (resample cell ranges, warp taps, median) must be bit exact.  Floating point
is expected to be bit exact too (same expression trees); the hard gate is the north-star tolerance
(mean / max endpoint difference <= 1e-3 / 1e-2 px), and exactness is reported.
"""
import numpy as np
import pytest

from conftest import epd_stats

pytestmark = pytest.mark.gpu
F = np.float32


def _eq(a, b):
    return a.shape == b.shape and bool(np.all(a == b))


def _rep(name, got, exp):
    d = np.abs(got.astype(np.float64) - exp.astype(np.float64))
    return "%s: %d/%d differ, max abs %.3e" % (name, int((got != exp).sum()), got.size, float(d.max()))


@pytest.fixture(scope="module")
def torch_():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.mark.parametrize("w,h,sigma", [(64, 48, 1.5), (584, 388, 1.5), (132, 68, 0.45), (200, 152, 3.0)])
def test_blur_vs_reference(pkg, oracle, ref, torch_, w, h, sigma):
    rng = np.random.default_rng(1)
    img = rng.uniform(0, 255, (h, w)).astype(F)
    exp = ref.conv(img, sigma)
    assert _eq(oracle.blur(img, sigma), exp), _rep("oracle blur", oracle.blur(img, sigma), exp)
    fl = pkg.Flow2D(w, h)
    d_in, d_out = fl.to_container(img), fl.container(float("nan"))
    fl.stage_blur(d_in, d_out, w, h, sigma)
    torch_.cuda.synchronize()
    got = fl.from_container(d_out, w, h)
    assert _eq(got, exp), _rep("blur", got, exp)


@pytest.mark.parametrize("iw,ih,ow,oh", [(584, 388, 526, 350), (584, 388, 5, 4), (47, 31, 52, 35), (200, 150, 64, 43), (6, 4, 6, 4)])
def test_resample_vs_reference(pkg, oracle, ref, torch_, iw, ih, ow, oh):
    rng = np.random.default_rng(2)
    img = rng.uniform(0, 255, (ih, iw)).astype(F)
    exp = ref.resample(img, ow, oh)
    o = oracle.resample(img, ow, oh)
    assert _eq(o, exp), _rep("oracle resample", o, exp)
    fl = pkg.Flow2D(max(iw, ow), max(ih, oh))
    d_in, d_out = fl.to_container(img), fl.container(float("nan"))
    fl.stage_resample(d_in, iw, ih, d_out, ow, oh)
    torch_.cuda.synchronize()
    got = fl.from_container(d_out, ow, oh)
    assert _eq(got, exp), _rep("resample", got, exp)


@pytest.mark.parametrize("w,h,hx,hy", [(64, 48, 1.0, 1.0), (131, 67, 2.92, 2.425), (20, 16, 29.2, 24.25)])
def test_warp_vs_reference(pkg, oracle, synth, ref, torch_, w, h, hx, hy):
    rng = np.random.default_rng(3)
    f0 = rng.uniform(0, 255, (h, w)).astype(F)
    f1 = rng.uniform(0, 255, (h, w)).astype(F)
    u = (synth.smooth_random(w, h, 1, -3, 3) * hx).astype(F)
    v = (synth.smooth_random(w, h, 2, -3, 3) * hy).astype(F)
    exp = ref.warp(f0, f1, u, v, hx, hy)
    o = oracle.warp(f0, f1, u, v, hx, hy)
    assert _eq(o, exp), _rep("oracle warp", o, exp)
    fl = pkg.Flow2D(w, h)
    d = [fl.to_container(a) for a in (f0, f1, u, v)]
    d_out = fl.container(float("nan"))
    fl.stage_warp(d[0], d[1], d[2], d[3], d_out, w, h, hx, hy)
    torch_.cuda.synchronize()
    got = fl.from_container(d_out, w, h)
    assert _eq(got, exp), _rep("warp", got, exp)


@pytest.mark.parametrize("w,h,radius", [(64, 48, 5), (37, 29, 3), (131, 67, 7), (9, 5, 5), (40, 40, 1), (40, 40, 6)])
def test_median_vs_reference(pkg, oracle, ref, torch_, w, h, radius):
    rng = np.random.default_rng(4)
    img = rng.normal(0, 2, (h, w)).astype(F)
    exp = ref.median(img, radius)
    o, _ = oracle.median(img, radius)
    assert _eq(o, exp), _rep("oracle median", o, exp)
    fl = pkg.Flow2D(max(w, 4), max(h, 4))
    d_in, d_out = fl.to_container(img), fl.container(float("nan"))
    fl.stage_median(d_in, d_out, w, h, radius)
    torch_.cuda.synchronize()
    got = fl.from_container(d_out, w, h)
    assert _eq(got, exp), _rep("median", got, exp)


SOLVE = [
    # w, h, hx, hy, outer, inner
    (37, 29, 1.0, 1.0, 1, 1),
    (64, 48, 1.0, 1.0, 2, 5),
    (131, 67, 2.92, 2.425, 3, 5),
    (200, 150, 1.0, 1.0, 2, 5),
    (5, 4, 116.8, 97.0, 3, 5),
]


@pytest.mark.parametrize("w,h,hx,hy,outer,inner", SOLVE)
def test_solve_vs_reference_grey(pkg, oracle, synth, ref, torch_, w, h, hx, hy, outer, inner):
    f0, f1, _, _ = synth.make_pair(w, h, 300 + w, U1=1.5, L=48.0)
    u = synth.smooth_random(w, h, 11, -2, 2)
    v = synth.smooth_random(w, h, 12, -2, 2)
    alpha, es, ed = 20.0, 0.001, 0.001
    rdu, rdv, rphi, rksi = ref.solve(f0, f1, u, v, hx, hy, alpha, es, ed, outer, inner, 0)
    op = oracle.make_params(outer=outer, inner=inner, alpha=alpha)
    odu, odv, ophi, oksi = oracle.solve_level(f0, f1, u, v, hx, hy, op)
    msgs = [_rep("oracle " + n, a, b) for n, a, b in (("phi", ophi, rphi), ("ksi", oksi, rksi), ("du", odu, rdu), ("dv", odv, rdv))]
    print("\n".join(msgs))
    fl = pkg.Flow2D(max(w, 4), max(h, 4))
    d = [fl.to_container(a) for a in (f0, f1, u, v)]
    t = [fl.container(float("nan")) for _ in range(4)]
    fl.stage_solve(d[0], d[1], d[2], d[3], t[0], t[1], t[2], t[3], w, h, hx, hy, pkg.default_params(outer=outer, inner=inner, alpha=alpha))
    torch_.cuda.synchronize()
    gdu, gdv, gphi, gksi = (fl.from_container(x, w, h) for x in t)
    for name, got, exp in (("phi", gphi, rphi), ("ksi", gksi, rksi), ("du", gdu, rdu), ("dv", gdv, rdv)):
        print(_rep("kernel " + name, got, exp))
    # hard gate: tolerance; the bit-exactness status is printed above and asserted where it holds
    scale = max(1e-6, float(np.abs(rdu).max()), float(np.abs(rdv).max()))
    assert np.abs(gdu - rdu).max() <= 1e-4 * scale and np.abs(gdv - rdv).max() <= 1e-4 * scale
    assert np.abs(odu - rdu).max() <= 1e-4 * scale and np.abs(odv - rdv).max() <= 1e-4 * scale
    assert _eq(gphi, rphi) and _eq(gksi, rksi), "robust weights must be bit exact"
    assert _eq(gdu, rdu) and _eq(gdv, rdv), "Jacobi sweeps expected bit exact (same expression tree as the reference SASS)"
    assert _eq(odu, rdu) and _eq(odv, rdv), "oracle expected bit exact with the reference"


@pytest.mark.parametrize("w,h,hx,hy,outer,inner", [(64, 48, 1.0, 1.0, 2, 5), (128, 96, 1.5, 1.25, 3, 5), (131, 67, 2.92, 2.425, 2, 5)])
def test_solve_vs_reference_gradient(pkg, oracle, synth, ref, torch_, w, h, hx, hy, outer, inner):
    """Gradient constancy (solve_2d_grad): the reference result depends on its 16x8 CUDA tiling, which is
    reproduced, and next to partial blocks it reads uninitialised shared memory, which
    cannot be.  Sizes that are multiples of 16x8 must match everywhere; otherwise the comparison covers
    the pixels that the undefined cells (last column / row) cannot reach within outer*inner sweeps."""
    f0, f1, _, _ = synth.make_pair(w, h, 400 + w, U1=1.5, L=48.0)
    u = synth.smooth_random(w, h, 11, -2, 2)
    v = synth.smooth_random(w, h, 12, -2, 2)
    alpha = 5.0
    rdu, rdv, rphi, rksi = ref.solve(f0, f1, u, v, hx, hy, alpha, 0.001, 0.001, outer, inner, 1)
    fl = pkg.Flow2D(w, h, constancy=1)
    d = [fl.to_container(a) for a in (f0, f1, u, v)]
    t = [fl.container(float("nan")) for _ in range(4)]
    fl.stage_solve(d[0], d[1], d[2], d[3], t[0], t[1], t[2], t[3], w, h, hx, hy, pkg.default_params(outer=outer, inner=inner, alpha=alpha))
    torch_.cuda.synchronize()
    gdu, gdv = fl.from_container(t[0], w, h), fl.from_container(t[1], w, h)
    op = oracle.make_params(outer=outer, inner=inner, alpha=alpha, constancy=oracle.GRADIENT)
    odu, odv, _, _ = oracle.solve_level(f0, f1, u, v, hx, hy, op)
    assert _eq(gdu, odu) and _eq(gdv, odv), "kernel vs oracle (gradient)"
    reach = 0 if (w % 16 == 0 and h % 8 == 0) else outer * inner + 1
    ys, xs = slice(0, h - reach), slice(0, w - reach)
    print(_rep("gradient du, full frame", gdu, rdu))
    assert _eq(gdu[ys, xs], rdu[ys, xs]) and _eq(gdv[ys, xs], rdv[ys, xs])


C1B = dict(levels=50, scale=0.9, outer=40, inner=5, alpha=35.0, e_smooth=0.001, e_data=0.001, median=5, sigma=1.5)
C1A = dict(levels=20, scale=0.9, outer=20, inner=5, alpha=3.5, e_smooth=0.001, e_data=0.001, median=5, sigma=0.45)


@pytest.mark.parametrize("name,cfg", [("C1b", C1B), ("C1a", C1A)])
def test_rub_pair_end_to_end_vs_reference(pkg, oracle, rub, ref, torch_, name, cfg):
    """BASELINE.json configs[0]: the bundled rub pair with the reference's two default parameter sets."""
    f0, f1 = rub
    ru, rv, _ = ref.flow(f0, f1, cfg)
    fl = pkg.Flow2D(584, 388)
    u, v = fl.compute(f0, f1, pkg.default_params(**cfg))
    s = epd_stats(u, v, ru, rv)
    print(name, "kernel vs reference:", s, fl.stats())
    ou, ov = oracle.compute_flow(f0, f1, oracle.make_params(**cfg))
    so = epd_stats(ou, ov, ru, rv)
    print(name, "oracle vs reference:", so)
    assert s["mean"] <= 1e-3 and s["max"] <= 1e-2, s
    assert so["mean"] <= 1e-3 and so["max"] <= 1e-2, so
