"""GPU parity tests, one per reference operator: the sm_100a kernels (through the C ABI's per-stage
entry points) against the CPU oracle on the same seeded inputs.  Bar: bit exact (`==`; the only
tolerated difference is the sign of a zero), because both sides implement the reference's
expression tree operation for operation and every operation is IEEE round-to-nearest."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32

SIZES = [(64, 48), (37, 29), (5, 4), (17, 8), (16, 128), (131, 67), (200, 150)]


@pytest.fixture(scope="module")
def torch_():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _sync(torch):
    torch.cuda.synchronize()


def _eq(a, b):
    return a.shape == b.shape and bool(np.all(a == b))


def _report(name, got, exp):
    d = np.abs(got.astype(np.float64) - exp.astype(np.float64))
    return "%s: %d/%d differ, max abs %.3e" % (name, int((got != exp).sum()), got.size, float(d.max()))


@pytest.mark.parametrize("w,h", [(64, 48), (37, 29), (131, 67), (584, 388), (20, 16)])
@pytest.mark.parametrize("sigma", [0.45, 1.5, 3.0, 5.5, 0.2])
def test_blur(pkg, oracle, torch_, w, h, sigma):
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.uniform(0, 255, (h, w)).astype(F)
    fl = pkg.Flow2D(w, h)
    d_in, d_out = fl.to_container(img), fl.container(float("nan"))
    fl.stage_blur(d_in, d_out, w, h, sigma)
    _sync(torch_)
    got, exp = fl.from_container(d_out, w, h), oracle.blur(img, sigma)
    assert _eq(got, exp), _report("blur", got, exp)


@pytest.mark.parametrize("iw,ih,ow,oh", [(584, 388, 526, 350), (584, 388, 5, 4), (64, 48, 58, 44), (17, 13, 19, 15),
                                          (6, 4, 6, 4), (6, 4, 7, 5), (200, 150, 180, 135), (47, 31, 52, 35), (300, 200, 64, 43)])
def test_resample(pkg, oracle, torch_, iw, ih, ow, oh):
    rng = np.random.default_rng(iw + ih)
    img = rng.uniform(-3, 255, (ih, iw)).astype(F)
    fl = pkg.Flow2D(max(iw, ow, 4), max(ih, oh, 4))
    d_in, d_out = fl.to_container(img), fl.container(float("nan"))
    fl.stage_resample(d_in, iw, ih, d_out, ow, oh)
    _sync(torch_)
    got, exp = fl.from_container(d_out, ow, oh), oracle.resample(img, ow, oh)
    assert _eq(got, exp), _report("resample", got, exp)


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("hx,hy", [(1.0, 1.0), (2.92, 2.425), (116.8, 97.0)])
def test_warp(pkg, oracle, synth, torch_, w, h, hx, hy):
    rng = np.random.default_rng(w + 7 * h)
    f0 = rng.uniform(0, 255, (h, w)).astype(F)
    f1 = rng.uniform(0, 255, (h, w)).astype(F)
    u = (synth.smooth_random(w, h, 1, -3, 3) * hx).astype(F)
    v = (synth.smooth_random(w, h, 2, -3, 3) * hy).astype(F)
    u[0, 0] = np.nan
    u[h // 2, w // 2] = 0.0   # exactly on a grid point
    v[h // 2, w // 2] = 0.0
    fl = pkg.Flow2D(max(w, 4), max(h, 4))
    d = [fl.to_container(a) for a in (f0, f1, u, v)]
    d_out = fl.container(float("nan"))
    fl.stage_warp(d[0], d[1], d[2], d[3], d_out, w, h, hx, hy)
    _sync(torch_)
    got, exp = fl.from_container(d_out, w, h), oracle.warp(f0, f1, u, v, hx, hy)
    assert _eq(got, exp), _report("warp", got, exp)


@pytest.mark.parametrize("w,h", SIZES + [(9, 5)])
@pytest.mark.parametrize("radius", [1, 3, 4, 5, 7])
def test_median_and_add_median(pkg, oracle, torch_, w, h, radius):
    rng = np.random.default_rng(w * 31 + h + radius)
    a = rng.normal(0, 2, (h, w)).astype(F)
    b = rng.normal(0, 0.3, (h, w)).astype(F)
    a[rng.integers(0, h, 5), rng.integers(0, w, 5)] = 7.5  # ties
    fl = pkg.Flow2D(max(w, 4), max(h, 4))
    d_a, d_b, d_o = fl.to_container(a), fl.to_container(b), fl.container(float("nan"))
    fl.stage_median(d_a, d_o, w, h, radius)
    _sync(torch_)
    got, (exp, rc) = fl.from_container(d_o, w, h), oracle.median(a, radius)
    assert rc == 0 and _eq(got, exp), _report("median", got, exp)
    fl.stage_add_median(d_a, d_b, d_o, w, h, radius)
    _sync(torch_)
    got, (exp, rc) = fl.from_container(d_o, w, h), oracle.median(a + b, radius)
    assert _eq(got, exp), _report("add_median", got, exp)
    fl.stage_add(d_a, d_b, w, h)
    _sync(torch_)
    assert _eq(fl.from_container(d_a, w, h), a + b)


@pytest.mark.parametrize("radius", [0, 2, 9, 11])
def test_median_unsupported_radius_is_an_error(pkg, torch_, radius):
    fl = pkg.Flow2D(16, 16)
    with pytest.raises(pkg.Flow2DError) as e:
        fl.stage_median(fl.container(0.0), fl.container(0.0), 16, 16, radius)
    assert e.value.code == -5


def _solve_inputs(synth, w, h, seed):
    f0, f1, _, _ = synth.make_pair(w, h, seed, U1=1.5, L=48.0)
    u = synth.smooth_random(w, h, seed + 1, -2, 2)
    v = synth.smooth_random(w, h, seed + 2, -2, 2)
    return f0, f1, u, v


def _run_solve(pkg, torch, fl, f0, f1, u, v, w, h, hx, hy, params):
    d = [fl.to_container(a) for a in (f0, f1, u, v)]
    d_du, d_dv, d_phi, d_ksi = (fl.container(float("nan")) for _ in range(4))
    fl.stage_solve(d[0], d[1], d[2], d[3], d_du, d_dv, d_phi, d_ksi, w, h, hx, hy, params)
    torch.cuda.synchronize()
    return tuple(fl.from_container(t, w, h) for t in (d_du, d_dv, d_phi, d_ksi))


SOLVE_CASES = [
    # w, h, hx, hy, outer, inner
    (37, 29, 1.0, 1.0, 2, 3),          # resident (single CTA)
    (5, 4, 116.8, 97.0, 3, 5),         # coarsest rub level: solve_tiny (one thread per pixel)
    (32, 32, 1.0, 1.0, 3, 5),          # largest square solve_tiny level
    (38, 26, 15.4, 14.9, 4, 5),        # solve_tiny, ragged
    (2, 2, 1.0, 1.0, 2, 2),            # smallest level the path accepts
    (59, 46, 1.25, 1.5, 2, 5),         # largest resident level
    (60, 46, 1.0, 1.0, 2, 5),          # one column too wide for resident mode -> tiled
    (59, 47, 1.0, 1.0, 2, 5),          # one row too high for resident mode -> tiled
    (131, 67, 2.92, 2.425, 2, 5),      # tiled, ragged
    (200, 150, 1.0, 1.0, 3, 5),        # tiled, several CTAs
    (96, 120, 1.0, 1.0, 1, 12),        # inner > sweeps per pass: phi/ksi stored and reloaded
    (64, 64, 1.0, 1.0, 2, 1),          # one sweep per pass
    (100, 90, 1.5, 1.5, 1, 7),
]


@pytest.mark.parametrize("w,h,hx,hy,outer,inner", SOLVE_CASES)
@pytest.mark.parametrize("constancy", [0, 1])
def test_solve_vs_oracle(pkg, oracle, synth, torch_, w, h, hx, hy, outer, inner, constancy):
    f0, f1, u, v = _solve_inputs(synth, w, h, 100 + w)
    fl = pkg.Flow2D(max(w, 4), max(h, 4), constancy=constancy)
    p = pkg.default_params(outer=outer, inner=inner, alpha=20.0)
    du, dv, phi, ksi = _run_solve(pkg, torch_, fl, f0, f1, u, v, w, h, hx, hy, p)
    op = oracle.make_params(outer=outer, inner=inner, alpha=20.0, constancy=constancy)
    edu, edv, ephi, eksi = oracle.solve_level(f0, f1, u, v, hx, hy, op)
    assert _eq(phi, ephi), _report("phi", phi, ephi)
    assert _eq(ksi, eksi), _report("ksi", ksi, eksi)
    assert _eq(du, edu), _report("du", du, edu)
    assert _eq(dv, edv), _report("dv", dv, edv)


@pytest.mark.parametrize("sweeps", [1, 2, 3, 4, 5, 6, 7])
def test_solve_is_independent_of_the_schedule(pkg, synth, torch_, sweeps):
    """Jacobi is tiling independent: every sweeps-per-pass / resident choice gives the same bits."""
    w, h = 150, 100
    f0, f1, u, v = _solve_inputs(synth, w, h, 7)
    fl = pkg.Flow2D(w, h)
    base = _run_solve(pkg, torch_, fl, f0, f1, u, v, w, h, 1.0, 1.0, pkg.default_params(outer=2, inner=7, sweeps_per_pass=5))
    got = _run_solve(pkg, torch_, fl, f0, f1, u, v, w, h, 1.0, 1.0, pkg.default_params(outer=2, inner=7, sweeps_per_pass=sweeps))
    assert _eq(got[0], base[0]) and _eq(got[1], base[1])


@pytest.mark.parametrize("w,h,inner", [(131, 67, 5), (200, 150, 5), (333, 250, 3), (64, 64, 1), (100, 90, 7)])
def test_small_pass_equals_solve_pass(pkg, synth, torch_, w, h, inner):
    """Mid-size levels: the one-thread-per-pixel pass (32x32 regions, chosen automatically) and the
    64x48-region solve_pass (resident_levels = -1 forces it) must give the same bits."""
    f0, f1, u, v = _solve_inputs(synth, w, h, 21)
    fl = pkg.Flow2D(w, h)
    a = _run_solve(pkg, torch_, fl, f0, f1, u, v, w, h, 1.7, 1.4, pkg.default_params(outer=3, inner=inner, resident_levels=0))
    b = _run_solve(pkg, torch_, fl, f0, f1, u, v, w, h, 1.7, 1.4, pkg.default_params(outer=3, inner=inner, resident_levels=-1))
    for k in range(4):
        assert _eq(a[k], b[k])


@pytest.mark.parametrize("w,h", [(48, 40), (30, 24), (12, 9)])
def test_tiny_resident_and_tiled_agree(pkg, synth, torch_, w, h):
    """resident_levels: 0 = automatic (solve_tiny for <= 1024 px, else resident solve_pass, else tiled),
    2 = resident solve_pass even for tiny levels, -1 = always tiled."""
    f0, f1, u, v = _solve_inputs(synth, w, h, 9)
    fl = pkg.Flow2D(w, h)
    r = [_run_solve(pkg, torch_, fl, f0, f1, u, v, w, h, 1.3, 1.1, pkg.default_params(outer=4, inner=5, resident_levels=m))
         for m in (0, 2, -1)]
    for other in r[1:]:
        for k in range(4):
            assert _eq(r[0][k], other[k])
